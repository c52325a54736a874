"""CPU, world_size 2, gloo: the host-side logic of the sharded path (row partition, CSR row blocks, in-place
all-gather of row blocks) reproduces the unsharded propagation.  The per-block SpMM is done by the oracle here —
on a GPU box the same code drives K1 (tests/multi_gpu_check.py)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import blocks, nums
from oracle import adjacency as OA
from oracle import propagation as OP

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny.npz")


def test_partition_balances_nnz():
    from tagrec_b200.distributed import partition_rows
    rng = np.random.RandomState(0)
    deg = np.r_[rng.randint(0, 5, 1000), [5000], rng.randint(0, 50, 200)]
    rowptr = np.r_[0, np.cumsum(deg)]
    for world in (1, 2, 4, 8):
        b = partition_rows(torch.tensor(rowptr), world)
        assert b[0] == 0 and b[-1] == len(deg) and all(x <= y for x, y in zip(b, b[1:]))
        per = [rowptr[b[i + 1]] - rowptr[b[i]] for i in range(world)]
        assert sum(per) == rowptr[-1]
        assert max(per) <= rowptr[-1] / world + 5000 + 50          # within one (largest) row of the ideal share


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tagrec_b200.distributed import RowComm, partition_rows, slice_csr
    g = dict(np.load(GOLDEN))
    U, I, _, _ = nums(g)
    n, rowptr, col, val = OA.creat_adj(U, I, blocks(g)[0], "bi_norm")
    bounds = partition_rows(torch.tensor(rowptr), world)
    comm = RowComm(bounds, rank, world)
    lo, hi = comm.lo, comm.hi
    rp, c, v = slice_csr(torch.tensor(rowptr), torch.tensor(col), torch.tensor(val), lo, hi)
    assert int(rp[0]) == 0 and rp.numel() == hi - lo + 1
    e0 = torch.cat([torch.tensor(g["lgcn_param_embed.0"]), torch.tensor(g["lgcn_param_embed.1"])]).double()
    # forward exactly as functional.lightgcn_forward_layers orders it: local block, then all-gather
    x, acc = e0, e0.clone()
    nl = 3
    for k in range(nl):
        y = torch.zeros_like(e0)
        # local rows of A times the full table (square CSR with only this block's rows filled)
        rp_full = np.zeros(n + 1, dtype=np.int64)
        rp_full[lo + 1:hi + 1] = rp.numpy()[1:]
        rp_full[hi + 1:] = rp_full[hi]
        y_full = OP.spmm(rp_full, c.numpy(), v.numpy(), x)
        y[lo:hi] = y_full[lo:hi]
        acc[lo:hi] += OP.row_normalise(y[lo:hi])[0]
        if k < nl - 1:
            comm.all_gather_rows(y)
        x = y
    acc[lo:hi] /= (nl + 1)
    comm.all_gather_rows(acc)
    ref, _ = OP.lightgcn_forward((rowptr, col, val), e0, nl)
    err = float((acc - ref).abs().max())
    if rank == 0:
        np.save(out, np.array([err, comm.bytes_moved]))
    dist.destroy_process_group()


def test_sharded_forward_world2_gloo(tmp_path):
    out = str(tmp_path / "res.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    err, moved = np.load(out)
    assert err < 1e-12
    assert moved > 0


class _FixedScores(torch.nn.Module):
    """Stand-in model for the CPU test of Basic_test's HOST logic (user sharding, chunking, all-reduce, division by
    the user count): eval_topk / eval_auc — device kernels in the product — are played by the oracle here."""

    def __init__(self, scores):
        super().__init__()
        self.scores = scores.numpy()

    def eval_topk(self, users, k, train_ptr, train_items, path="auto"):
        from oracle import metrics as OM
        u = users.numpy()
        ms = OM.mask_train(self.scores[u], u, train_ptr.numpy(), train_items.numpy())
        return torch.as_tensor(OM.topk_ids(ms, k).astype(np.int32)), None

    def eval_auc(self, users, train_ptr, train_items, test_ptr, test_items, out=None):
        from oracle import metrics as OM
        u = users.numpy()
        ms = OM.mask_train(self.scores[u], u, train_ptr.numpy(), train_items.numpy())
        tp, ti = test_ptr.numpy(), test_items.numpy()
        out += torch.tensor([sum(OM.auc_one(ms[r], ti[tp[x]:tp[x + 1]]) for r, x in enumerate(u)), float(len(u))],
                            dtype=torch.float64)
        return out


def _oracle_metric_sums(users, topk_ids, test_ptr, test_items, ks, out=None):
    """tagrec_eval_metrics (a device kernel) played by the oracle."""
    from oracle import metrics as OM
    r = OM.ranking_metrics(topk_ids.numpy(), users.numpy(), test_ptr.numpy(), test_items.numpy(), list(ks))
    out += torch.tensor([r["recall"], r["precision"], r["hr"], r["ndcg"]], dtype=torch.float64)
    return out


def _eval_worker(rank, world, port, out):
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import tagrec_b200 as T
    from helpers import user_lists
    g = dict(np.load(GOLDEN))
    U, I, _, _ = nums(g)
    T.set_config("lightgcn", test_batch=7, topks=[5, 20], device=torch.device("cpu"), has_val=False)

    class D:
        pass
    d = D()
    d.num = {"user": U, "item": I}
    d.user_items = {"train": user_lists(g, "train"), "test": user_lists(g, "test")}
    scores = torch.sigmoid(torch.tensor(g["lgcn_fwd_0"]) @ torch.tensor(g["lgcn_fwd_1"]).T)
    T.basic_test.metric_sums = _oracle_metric_sums
    res = T.Basic_test(d).run(_FixedScores(scores))
    if rank == 0:
        np.save(out, np.array([res[k][j] for k in ("recall", "precision", "hr", "ndcg") for j in range(2)] + res["auc"],
                              dtype=np.float64))
    if world > 1:
        dist.destroy_process_group()


def test_sharded_evaluation_world2_gloo(tmp_path):
    """Basic_test under torch.distributed: users sharded over 2 ranks + all-reduce == single process == the
    reference's epoch_test (golden eval_* of the tiny dataset)."""
    outs = []
    for world in (1, 2):
        out = str(tmp_path / f"eval{world}.npy")
        port = 31500 + os.getpid() % 2000 + world
        if world == 1:
            _eval_worker(0, 1, port, out)
        else:
            mp.spawn(_eval_worker, args=(world, port, out), nprocs=world, join=True)
        outs.append(np.load(out))
    assert np.allclose(outs[0], outs[1], rtol=0, atol=1e-12)
    g = dict(np.load(GOLDEN))
    want = np.r_[[g[f"eval_{k}"][j] for k in ("recall", "precision", "hr", "ndcg") for j in range(2)], g["eval_auc"]]
    assert np.allclose(outs[0], want, atol=1e-6), (outs[0], want)
