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


class _CpuGraph:
    """A row block of the normalised adjacency on the CPU: the fields distributed.* / NGCF touch, with K1 (a device
    kernel in the product) played by a dense-index torch reference below."""

    def __init__(self, n, rowptr, col, val, val_t, row_offset, comm, num_list):
        self.n, self.rowptr, self.col, self.val, self.val_t = n, rowptr, col, val, val_t
        self.row_offset, self.comm, self.num_list, self.n_rows = row_offset, comm, num_list, rowptr.numel() - 1


def _oracle_spmm_raw(graph, x, out=None, transposed=False, beta=0.0):
    """tagrec_spmm on a row block: out[row_offset + r] = sum_j val[j] x[col[j]] (other rows untouched)."""
    if out is None:
        out = torch.empty_like(x)
    deg = graph.rowptr[1:] - graph.rowptr[:-1]
    rows = torch.repeat_interleave(torch.arange(graph.n_rows), deg)
    v = (graph.val_t if transposed else graph.val).to(x.dtype)
    y = torch.zeros((graph.n_rows, x.shape[1]), dtype=x.dtype)
    y.index_add_(0, rows, v[:, None] * x[graph.col.long()])
    out[graph.row_offset:graph.row_offset + graph.n_rows] = y
    return out


def _ngcf_worker(rank, world, port, out):
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import tagrec_b200 as T
    from tagrec_b200 import adj as A
    from tagrec_b200.distributed import RowComm, partition_rows, slice_csr
    A.spmm_raw = _oracle_spmm_raw
    g = dict(np.load(GOLDEN))
    U, I, _, _ = nums(g)
    n, rowptr, col, val = OA.creat_adj(U, I, blocks(g)[0], "ngcf")
    # the 'ngcf' normalisation is not symmetric: values of A^T through the transposed CSR
    import scipy.sparse as sp
    at = sp.csr_matrix((val, col, rowptr), shape=(n, n)).T.tocsr()
    at.sort_indices()
    assert np.array_equal(at.indptr, rowptr) and np.array_equal(at.indices, col)
    rp, c, v, vt = (torch.tensor(np.asarray(x)) for x in (rowptr, col, val, at.data.astype(np.float32)))
    bounds = partition_rows(rp, world)
    comm = RowComm(bounds, rank, world) if world > 1 else None
    lo, hi = (comm.lo, comm.hi) if comm else (0, n)
    rpl, cl, vl = slice_csr(rp, c, v, lo, hi)
    vtl = vt[int(rp[lo]):int(rp[hi])].clone()
    graph = _CpuGraph(n, rpl, cl, vl, vtl, lo, comm, [U, I])
    T.set_config("ngcf", use_tag=False, reg=1e-3, dim_layer_list=[32, 16], device=torch.device("cpu"))

    class D:
        num = {"user": U, "item": I}
        prebuilt_adj = graph
    torch.manual_seed(11)
    m = T.NGCF(D).double()
    m.train()
    final = m._final_table()
    rng = np.random.RandomState(3)
    bu, bi, bj = rng.randint(0, U, 64), rng.randint(0, I, 64), rng.randint(0, I, 64)
    fu, fi, fj = final[bu], final[U + bi], final[U + bj]
    loss = torch.nn.functional.softplus(-((fu * fi).sum(1) - (fu * fj).sum(1))).mean() + 1e-3 * (fu.pow(2).sum() + fi.pow(2).sum())
    loss.backward()
    if rank == 0:
        np.savez(out, loss=loss.item(), final=final.detach().numpy(),
                 **{f"g_{k}": p.grad.numpy() for k, p in m.named_parameters()})
    if world > 1:
        # replicas bit-identical: every gradient row / dense gradient is produced once and shared
        for k, p in m.named_parameters():
            r0 = p.grad.clone()
            dist.broadcast(r0, src=0)
            assert torch.equal(r0, p.grad), k
        dist.destroy_process_group()


def test_sharded_ngcf_world2_gloo(tmp_path):
    """NGCF on a node-range sharded graph (distributed.ShardedSpMMFn / GatherRowsFn / RowOwnedParamFn /
    AllReduceGradFn: K1 on row blocks, all-gather of layer outputs, all-gather of embedding-gradient rows, ALL-REDUCE
    of the dense weight gradients) == the unsharded model: propagated table, loss and the gradient of every parameter;
    replicas bit-identical.  K1 is played by a torch reference on the CPU (GPU: tests/multi_gpu_check.py)."""
    res = []
    for world in (1, 2):
        out = str(tmp_path / f"ngcf{world}.npz")
        port = 33500 + os.getpid() % 2000 + world
        if world == 1:
            _ngcf_worker(0, 1, port, out)
        else:
            mp.spawn(_ngcf_worker, args=(world, port, out), nprocs=world, join=True)
        res.append(dict(np.load(out)))
    a, b = res
    assert abs(a["loss"] - b["loss"]) < 1e-12 * abs(a["loss"])
    assert np.abs(a["final"] - b["final"]).max() < 1e-12
    for k in a:
        if k.startswith("g_"):
            assert np.abs(a[k] - b[k]).max() <= 1e-10 * max(np.abs(a[k]).max(), 1e-300), k
    assert any(k.startswith("g_mat.W1") for k in a) and np.abs(a["g_mat.W1_0"]).max() > 0


def test_cut_by_cost_and_feedback_scaling():
    """Host logic of the measured partition (distributed.row_costs / cut_by_cost): contiguous ranges of equal modelled
    cost; scaling one range's cost (a rank measured slower than the mean) moves its cuts inward; degenerate inputs stay
    monotone."""
    from tagrec_b200.distributed import cut_by_cost, row_costs

    class G:
        pass
    rng = np.random.RandomState(1)
    deg = np.r_[rng.randint(1, 200, 4000), rng.randint(200, 4000, 400)]
    g = G()
    g.rowptr = torch.tensor(np.r_[0, np.cumsum(deg)])
    tb, tw = [0, 4000, 4400], [3e-11, 1e-11]                      # user rows cost 3x per entry
    cost = row_costs(g, tb, tw)
    assert cost.shape[0] == 4400 and float(cost.min()) > 0
    for world in (2, 4, 8):
        b = cut_by_cost(cost, world)
        assert b[0] == 0 and b[-1] == 4400 and all(x <= y for x, y in zip(b, b[1:]))
        per = [float(cost[b[i]:b[i + 1]].sum()) for i in range(world)]
        assert max(per) - min(per) <= 2 * float(cost.max())          # within one (largest) row of each other
    b4 = cut_by_cost(cost, 4)
    slow = cost.clone()
    slow[b4[1]:b4[2]] *= 1.3                                         # rank 1 measured 30 % over the mean
    n4 = cut_by_cost(slow, 4)
    assert n4[2] - n4[1] < b4[2] - b4[1]                             # it gets fewer rows
    assert cut_by_cost(torch.ones(3, dtype=torch.float64), 8)[-1] == 3      # more ranks than rows: still monotone
