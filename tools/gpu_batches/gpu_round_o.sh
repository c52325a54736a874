#!/bin/sh
# Round O (2 GPUs): parity of every multi-GPU mode (incl. split forward / backward partitions), then the N = 2 bench with
# and without the split partitions.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    tests/multi_gpu_check.py --out gpurun_out/r2_multi_gpu_parity_n2.jsonl > gpurun_out/n2_parity.log 2>&1
echo "parity rc=$?"; grep -E "MULTI|Error|error" gpurun_out/n2_parity.log | head -5; grep -o '"mode": "[a-z-]*"' gpurun_out/n2_parity.log | tr '\n' ' '
for sp in 1 0; do
  TAGREC_SPLIT_PARTITION=$sp timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$sp \
      bench.py --gpus 2 --steps 5 --warmup 3 --eval-users 0 > gpurun_out/ro_n2_sp$sp.json 2> gpurun_out/ro_n2_sp$sp.err
  python - $sp <<'PY'
import json, sys
sp = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/ro_n2_sp{sp}.json") if l.startswith("{")][-1])
    print(f"N=2 split={sp}: {d['ms_per_step']:.2f} ms/step  check {d['check']['last_loss']} {d['check']['param_abs_sum']:.6f}  fwd {[round(r['fwd_ms'],2) for r in d['per_rank']]} bwd {[round(r['bwd_ms'],2) for r in d['per_rank']]} bounds {d['partition']['bounds']} bwd {d['partition'].get('bounds_bwd')}")
except Exception as e:
    print(f"N=2 split={sp}: FAILED {e}"); print(open(f"gpurun_out/ro_n2_sp{sp}.err").read()[-1500:])
PY
done
