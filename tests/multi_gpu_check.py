"""Run under torchrun on N >= 2 GPUs: the node-range sharded LightGCN step (K1 on row blocks + NCCL all-gather)
equals the single-GPU step on the same inputs.  Exit code 0 and 'MULTI_GPU_OK' on success.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import tagrec_b200 as T
    from tagrec_b200.distributed import shard_graph
    U, I = 30000, 6000
    rng = np.random.RandomState(0)
    u = np.r_[rng.randint(0, U, 400000), rng.permutation(U)[:9000]]
    i = np.r_[rng.randint(0, I, 400000), np.zeros(9000, dtype=np.int64)]           # item 0 is a long row
    key = np.unique(u.astype(np.int64) * I + i)
    ui = (key // I, key % I)
    T.set_config("lightgcn", use_tag=False, reg=1e-3, dim_layer_list=[64, 64, 64], device=dev, init_device=dev)
    full = T.build_csr(U, I, ui, "bi_norm", dev)
    batch = torch.tensor(np.stack([ui[0][:2048], ui[1][:2048], rng.randint(0, I, 2048)], 1), device=dev)

    def run(graph):
        class D:
            num = {"user": U, "item": I}
            prebuilt_adj = graph
        torch.manual_seed(5)
        m = T.LightGCN(D)
        m.train()
        lossx = m.loss(batch)
        sum(lossx).backward()
        fw = torch.cat([t.detach() for t in m.forward()])
        return [x.item() for x in lossx], torch.cat([p.grad for p in m.embed]), fw

    l1, g1, f1 = run(full)
    l2, g2, f2 = run(shard_graph(full, rank, world))
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    errs = (abs(l1[0] - l2[0]) / abs(l1[0]), abs(l1[1] - l2[1]) / abs(l1[1]), rel(g2, g1), rel(f2, f1))
    ok = all(e < 1e-5 for e in errs)
    # replicas must be bit-identical across ranks (each row is produced by exactly one rank)
    ref = g2.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(ref, g2))
    flags = torch.tensor([int(ok), int(same)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"world={world} errs(loss,reg,grad,final)={errs} replicas_identical={bool(flags[1])}")
        print("MULTI_GPU_OK" if flags.min().item() == 1 else "MULTI_GPU_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if flags.min().item() == 1 else 1)


if __name__ == "__main__":
    main()
