#!/bin/sh
# 8-GPU box: multi-GPU parity (every exchange path vs the single-GPU step), then the scaling bench at N = 8 and N = 4.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py --out gpurun_out/r2_multi_gpu_parity_n8.jsonl > gpurun_out/n8_parity.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/n8_parity.log | cut -c1-300
for n in ${NS:-8 4}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
      bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err
  echo "bench n=$n rc=$?"
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2_scale_n{n}.json") if l.startswith("{")][-1])
    print(f"N={n}: {d['ms_per_step']:.2f} ms/step  value {d['value']:.0f}  e2e {d['e2e']['value']:.0f}  check {d['check']['last_loss']} {d['check']['param_abs_sum']:.6f}")
    for r in d.get("per_rank", []):
        print("   ", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()})
    print("   eval", {k: d.get("eval", {}).get(k) for k in ("users_per_s", "ms", "users")})
except Exception as e:
    print(f"N={n}: FAILED {e}")
PY
done
