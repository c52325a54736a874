#!/bin/sh
# Round I (2 GPUs): the whole GPU suite (incl. multi-GPU parity with the Adam epilogue), then the N = 2 bench with and
# without the epilogue form.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/ri_tests.log
grep -E "passed|failed|error" gpurun_out/ri_tests.log
cp gpurun_out/multi_gpu_parity_n2.jsonl gpurun_out/r2_multi_gpu_parity_n2.jsonl 2>/dev/null
for ep in 1 0; do
  TAGREC_ADAM_EPILOGUE=$ep timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$ep \
      bench.py --gpus 2 --steps 5 --warmup 3 --eval-users 0 > gpurun_out/ri_n2_ep$ep.json 2> gpurun_out/ri_n2_ep$ep.err
  python - $ep <<'PY'
import json, sys
ep = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/ri_n2_ep{ep}.json") if l.startswith("{")][-1])
    print(f"N=2 epilogue={ep}: {d['ms_per_step']:.2f} ms/step  check {d['check']['last_loss']} {d['check']['param_abs_sum']:.6f}  adam_ms {[r['adam_ms'] for r in d['per_rank']]} bwd {[r['bwd_ms'] for r in d['per_rank']]}")
except Exception as e:
    print(f"N=2 epilogue={ep}: FAILED {e}")
PY
done
