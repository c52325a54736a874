#!/bin/sh
# Final 1-GPU round: the GPU suite, smoke(), the default bench line of both arms exactly as the driver runs it, and the four
# named BASELINE workloads (both arms on the full configuration).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/rfinal_tests.log
grep -E "passed|failed|error" gpurun_out/rfinal_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/r2_bench_final_n1_reference.json 2> gpurun_out/rfinal_ref.err; echo "reference arm rc=$?"
python bench.py > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/rfinal_ours.err; echo "our arm rc=$?"
for w in lastfm delicious_tags_tgcn amazon_book_ngcf gowalla_dgcf; do
  python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2_bench_$w.json 2> gpurun_out/rfinal_$w.err; echo "$w rc=$?"
done
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_final_n1.json") if l.startswith("{")][-1])
r = d["roofline"]
print(f"default: {d['ms_per_step']:.1f} ms/step value {d['value']:.0f} e2e {d['e2e']['value']:.0f} fwd {r['ms_per_launch']:.2f} rows {r.get('fwd_last_layer_rows_ms')} bwd {r['bwd_launch_ms']} frac {r['frac']:.3f} dram_frac {r.get('dram_frac')} eval {d['eval']['ms']:.2f} ms {d['eval']['users_per_s']:.0f} users/s clocks {d['clocks']}")
for w in ("lastfm", "delicious_tags_tgcn", "amazon_book_ngcf", "gowalla_dgcf"):
    try:
        x = json.loads([l for l in open(f"gpurun_out/r2_bench_{w}.json") if l.startswith("{")][-1])
        print(f"{w}: {x['ms_per_step']:.3f} ms/step graphed {x['graphed_step']} e2e {x['e2e']['value']:.0f} cpu {x.get('cpu_baseline', {}).get('ms_per_step')} eval {[round(x['eval'][k]['users_per_s']) for k in ('topk_only', 'with_auc')]}")
    except Exception as e:
        print(w, "FAILED", e)
PY
