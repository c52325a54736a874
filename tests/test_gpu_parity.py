"""GPU parity tests: the CUDA path (through the C ABI) vs the oracle and the golden fixtures from the reference.

Tolerances (north_star): bit-exact CSR / top-K ids; <= 1e-5 relative (to the tensor's max magnitude) for fp32
embeddings, loss and gradients; <= 1e-4 for Recall/NDCG.
"""
import numpy as np
import pytest
import torch

import tagrec_b200 as T
from helpers import blocks, coalesced, nums, relerr, user_lists
from oracle import adjacency as OA
from oracle import metrics as OM
from oracle import propagation as OP

pytestmark = pytest.mark.gpu
NORMS = ("bi_norm", "si_norm", "si_norm_self", "ngcf", "plain")
TOL = 1e-5


def dev():
    return torch.device("cuda:0")


def make_data(g, tags=False):
    class D:
        pass
    import scipy.sparse as sp
    U, I, Tg, W = nums(g)
    ui, ut, it = blocks(g)
    d = D()
    d.num = {"user": U, "item": I, "tag": Tg, "weight": W}
    coo = lambda rc, shape: sp.coo_matrix((np.ones(len(rc[0])), rc), dtype=np.float32, shape=shape)
    d.ui_adj, d.ut_adj, d.it_adj = coo(ui, (U, I)), coo(ut, (U, Tg)), coo(it, (I, Tg))
    d.user_items = {"train": user_lists(g, "train"), "test": user_lists(g, "test")}
    d.edge_index = {"train": g["edge_index_train"]}
    return d


# ------------------------------------------------------------------------------------------------------------ K0
@pytest.mark.parametrize("use_tag", [False, True])
@pytest.mark.parametrize("nt", NORMS)
def test_k0_csr_bit_exact_vs_reference(tiny, use_tag, nt):
    U, I, Tg, _ = nums(tiny)
    ui, ut, it = blocks(tiny)
    g = T.build_csr(U, I, ui, nt, dev(), Tg, ut if use_tag else None, it if use_tag else None)
    tag = f"adj_{'uit' if use_tag else 'ui'}_{nt}"
    grow, gcol, gval = coalesced(g.n, tiny[tag + "_row"], tiny[tag + "_col"], tiny[tag + "_val"])
    assert np.array_equal(g.row_ids().cpu().numpy(), grow)
    assert np.array_equal(g.col.cpu().numpy(), gcol)
    assert np.array_equal(g.val.cpu().numpy().view(np.uint32), gval.view(np.uint32))


def test_k0_transposed_values(tiny):
    U, I, _, _ = nums(tiny)
    g = T.build_csr(U, I, blocks(tiny)[0], "ngcf", dev())
    import scipy.sparse as sp
    a = sp.csr_matrix((g.val.cpu().numpy(), g.col.cpu().numpy(), g.rowptr.cpu().numpy()), shape=g.shape)
    at = sp.csr_matrix((g.val_t.cpu().numpy(), g.col.cpu().numpy(), g.rowptr.cpu().numpy()), shape=g.shape)
    assert abs(a.T - at).max() == 0


def test_k0_empty_and_isolated():
    g = T.build_csr(3, 4, (np.array([1]), np.array([2])), "bi_norm", dev())
    assert g.rowptr.cpu().tolist() == [0, 0, 1, 1, 1, 1, 2, 2]
    assert g.col.cpu().tolist() == [5, 1] and g.val.cpu().tolist() == [1.0, 1.0]


# ------------------------------------------------------------------------------------------------------------ K1
def random_graph(U, I, E, seed, hub=0):
    rng = np.random.RandomState(seed)
    u = rng.randint(0, U, E)
    i = rng.randint(0, I, E)
    if hub:   # one item connected to `hub` users -> exercises the long-row chunk path
        u = np.r_[u, rng.permutation(U)[:hub]]
        i = np.r_[i, np.zeros(hub, dtype=np.int64)]
    key = np.unique(u.astype(np.int64) * I + i)
    return key // I, key % I


@pytest.mark.parametrize("dim", [32, 64, 128])
@pytest.mark.parametrize("hub", [0, 9000])
def test_k1_spmm_plain_and_transpose(dim, hub):
    U, I = 12000, 3000
    ui = random_graph(U, I, 60000, 1, hub)
    for nt in ("bi_norm", "ngcf"):
        g = T.build_csr(U, I, ui, nt, dev())
        assert (g.n_long > 0) == (hub > 0)
        x = torch.randn(g.n, dim, device=dev())
        csr = (g.rowptr.cpu().numpy(), g.col.cpu().numpy(), g.val.cpu().numpy())
        y = T.spmm_raw(g, x)
        ref = OP.spmm(*csr, x.cpu().double())
        assert relerr(y.cpu().numpy(), ref.numpy()) < TOL
        yt = T.spmm_raw(g, x, transposed=True)
        reft = OP.spmm_t(*csr, x.cpu().double())
        assert relerr(yt.cpu().numpy(), reft.numpy()) < TOL
        # scratch rows / counters are left clean for the next launch
        y2 = T.spmm_raw(g, x)
        assert torch.equal(y, y2) or relerr(y2.cpu().numpy(), ref.numpy()) < TOL


def test_k1_autograd_matches_transpose():
    U, I = 500, 700
    g = T.build_csr(U, I, random_graph(U, I, 5000, 2), "si_norm", dev())
    x = torch.randn(g.n, 64, device=dev(), requires_grad=True)
    w = torch.randn(g.n, 64, device=dev())
    (T.split_mm(g, x) * w).sum().backward()
    csr = (g.rowptr.cpu().numpy(), g.col.cpu().numpy(), g.val.cpu().numpy())
    assert relerr(x.grad.cpu().numpy(), OP.spmm_t(*csr, w.cpu().double()).numpy()) < TOL


@pytest.mark.parametrize("tag,use_tag,kind", [("lgcn", False, "softplus"), ("lgcn_tag", True, "softplus"),
                                               ("lgcn_logsig", False, "logsigmoid")])
def test_lightgcn_forward_loss_grad_vs_reference(tiny, tag, use_tag, kind):
    """model.forward / model.loss + backward == the reference's outputs on identical inputs."""
    T.set_config("lightgcn", use_tag=use_tag, reg=1e-3, dim_layer_list=[64, 64, 64], device=dev(),
                 mul_loss_func=kind)
    model = T.LightGCN(make_data(tiny)).to(dev())
    ne = 3 if use_tag else 2
    with torch.no_grad():
        for k in range(ne):
            model.embed[k].copy_(torch.tensor(tiny[f"{tag}_param_embed.{k}"]))
    assert list(model.state_dict().keys()) == [f"embed.{k}" for k in range(ne)]
    model.train()
    fw = model.forward()
    for k in range(ne):
        assert relerr(fw[k].detach().cpu().numpy(), tiny[f"{tag}_fwd_{k}"]) < TOL
    lossx = model.loss(torch.tensor(tiny[f"{tag}_batch"], device=dev()))
    assert isinstance(lossx, tuple) and len(lossx) == 2 and lossx[0].dim() == 0
    assert abs(lossx[0].item() - tiny[f"{tag}_loss"][0]) < TOL * abs(tiny[f"{tag}_loss"][0])
    assert abs(lossx[1].item() - tiny[f"{tag}_loss"][1]) < TOL * abs(tiny[f"{tag}_loss"][1])
    sum(lossx).backward()
    for k in range(ne):
        assert relerr(model.embed[k].grad.cpu().numpy(), tiny[f"{tag}_grad_embed.{k}"]) < TOL
    # differentiable forward() (separate autograd node) gives the same gradient
    model.zero_grad()
    fw = model.forward()
    b = torch.tensor(tiny[f"{tag}_batch"], device=dev())
    U = model.num_list[0]
    fu, fp, fn = fw[0][b[:, 0]], fw[1][b[:, 1]], fw[1][b[:, 2]]
    loss = OP.bpr_loss(fu, fp, fn, kind)
    eu, ei = model.get_ego_embed()[:2]
    reg = 1e-3 * OP.l2reg(eu[b[:, 0]], ei[b[:, 1]], ei[b[:, 2]])
    (loss + reg).backward()
    for k in range(ne):
        assert relerr(model.embed[k].grad.cpu().numpy(), tiny[f"{tag}_grad_embed.{k}"]) < TOL
    model.eval()
    with torch.no_grad():
        r = model.predict_rating(torch.tensor(tiny[f"{tag}_pred_users"], device=dev()))
    assert relerr(r.cpu().numpy(), tiny[f"{tag}_pred"]) < TOL


def test_lightgcn_long_rows_and_upstream_scale():
    """Hub rows (chunked path) + non-unit upstream gradients, vs the fp64 oracle."""
    U, I = 9000, 500
    ui = random_graph(U, I, 30000, 3, hub=8000)
    T.set_config("lightgcn", use_tag=False, reg=1e-2, dim_layer_list=[64, 64], device=dev())

    class D:
        num = {"user": U, "item": I}
    import scipy.sparse as sp
    D.ui_adj = sp.coo_matrix((np.ones(len(ui[0])), ui), dtype=np.float32, shape=(U, I))
    torch.manual_seed(1)
    model = T.LightGCN(D).to(dev())
    assert model.norm_adj.n_long >= 1
    rng = np.random.RandomState(0)
    batch = np.stack([ui[0][:512], ui[1][:512], rng.randint(0, I, 512)], 1).astype(np.int64)
    batch[:64, 1] = 0                                   # the hub item, many times in one batch
    lossx = model.loss(torch.tensor(batch, device=dev()))
    (2.0 * lossx[0] + 3.0 * lossx[1]).backward()
    g = model.norm_adj
    csr = (g.rowptr.cpu().numpy(), g.col.cpu().numpy(), g.val.cpu().numpy())
    e0 = torch.cat([p.detach().cpu().double() for p in model.embed])
    final, raw = OP.lightgcn_forward(csr, e0, 2)
    loss, reg, gf, ge = OP.bpr_forward_backward(final, e0, batch, U, 1e-2, "softplus")
    g0 = OP.lightgcn_backward(csr, raw, 2.0 * gf, 2) + 3.0 * ge
    got = torch.cat([p.grad.cpu().double() for p in model.embed])
    assert relerr(got.numpy(), g0.numpy()) < TOL
    assert abs(lossx[0].item() - loss.item()) < TOL and abs(lossx[1].item() - reg.item()) < TOL * reg.item()


def test_training_trajectory_vs_reference(tiny):
    """3+1 Adam steps through Basic_train's epoch_training on the reference's fixed triple file: same per-step
    losses and same parameters afterwards (incl. the tail batch being trained twice)."""
    T.set_config("lightgcn", use_tag=False, reg=1e-3, dim_layer_list=[64, 64, 64], device=dev(), train_batch=64, lr=0.01)
    model = T.LightGCN(make_data(tiny)).to(dev())
    with torch.no_grad():
        for k in range(2):
            model.embed[k].copy_(torch.tensor(tiny[f"lgcn_param_embed.{k}"]))
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    triples = torch.tensor(tiny["train_triples"], device=dev())

    class Fixed(T.Abstract_training_data):
        batch_size = 64

        def reset(self):
            self.all_train_data = triples
    model.train()
    losses = T.epoch_training(Fixed(), model.loss, opt)
    assert len(losses) == len(tiny["train_losses"])
    assert np.allclose(losses, tiny["train_losses"], rtol=2e-5, atol=0)
    for k in range(2):
        assert relerr(model.embed[k].detach().cpu().numpy(), tiny[f"train_after_embed.{k}"]) < 1e-4


# ------------------------------------------------------------------------------------------------------------ K2
@pytest.mark.parametrize("dim", [32, 64, 128])
@pytest.mark.parametrize("kind", ["softplus", "logsigmoid"])
def test_k2_bpr_vs_oracle(dim, kind):
    from tagrec_b200.functional import bpr_fwd_bwd
    U, I, B = 300, 200, 777
    rng = np.random.RandomState(5)
    final = torch.randn(U + I, dim) * 0.5
    ego = torch.randn(U + I, dim) * 0.5
    batch = np.stack([rng.randint(0, 20, B), rng.randint(0, 10, B), rng.randint(0, I, B)], 1).astype(np.int64)
    batch[::2, 0] = batch[1::2, 0][:len(batch[::2])] if B % 2 == 0 else batch[::2, 0]   # same user in both halves of a warp
    loss, reg, gf, ge = OP.bpr_forward_backward(final.double(), ego.double(), batch, U, 0.05, kind)
    fd, ed = final.to(dev()), ego.to(dev())
    g_final, g_reg = torch.zeros_like(fd), torch.zeros_like(ed)
    out = torch.empty(2, device=dev())
    bpr_fwd_bwd(torch.tensor(batch, device=dev()), U, fd, ed, 0.05, kind, g_final, g_reg, out)
    assert abs(out[0].item() - loss.item()) < TOL * abs(loss.item())
    assert abs(out[1].item() - reg.item()) < TOL * abs(reg.item())
    assert relerr(g_final.cpu().numpy(), gf.numpy()) < TOL
    assert relerr(g_reg.cpu().numpy(), ge.numpy()) < TOL


# ------------------------------------------------------------------------------------------------------------ K3
def near_tie_ok(ms_row, ids_a, ids_b, k, tol=1e-6):
    """Sets may differ only through items whose score is within tol (relative) of the k-th score."""
    a, b = set(ids_a.tolist()), set(ids_b.tolist())
    if a == b:
        return True
    kth = np.sort(ms_row)[::-1][k - 1]
    return all(abs(ms_row[i] - kth) <= tol * max(1.0, abs(kth)) for i in a ^ b)


def test_k3_topk_and_metrics_vs_reference(medium):
    T.set_config("lightgcn", use_tag=False, reg=0.0, dim_layer_list=[64, 64, 64], device=dev(), test_batch=16,
                 topks=[10, 20])
    d = make_data(medium)
    model = T.LightGCN(d).to(dev())
    with torch.no_grad():
        for k in range(2):
            model.embed[k].copy_(torch.tensor(medium[f"lgcn_param_embed.{k}"]))
    model.eval()
    U, I, _, _ = nums(medium)
    users = medium["eval_users"]
    tp, ti = medium["train_ptr"], medium["train_items"]
    ptr_, items = T.bpr_training_data.user_items_to_csr(d.user_items["train"], U)
    ids, scores = model.eval_topk(torch.tensor(users, device=dev()), 20, torch.tensor(ptr_, device=dev()),
                                  torch.tensor(items, device=dev()).int())
    ids, scores = ids.cpu().numpy(), scores.cpu().numpy()
    # (a) against the (-score, id) order of the reference's own masked predict_rating output
    ref_ids, ref_scores = medium["eval_top40_ids"], medium["eval_top40_scores"]
    exact = 0
    for r in range(len(users)):
        row = np.full(I, -np.inf)
        row[ref_ids[r]] = ref_scores[r]
        if not np.array_equal(ids[r], ref_ids[r, :20]):
            kth = ref_scores[r, 19]
            diff = set(ids[r]) ^ set(ref_ids[r, :20])
            assert all(abs(row[i] - kth) <= 1e-6 for i in diff if np.isfinite(row[i])), f"user row {r}: not a near-tie"
        else:
            exact += 1
    assert exact >= len(users) - 3, f"only {exact}/{len(users)} rows identical"
    assert np.allclose(scores[:, 0], ref_scores[:, 0], atol=1e-6)
    # (b) metrics through the drop-in Basic_test vs the reference's epoch_test
    res = T.Basic_test(d).run(model)
    for k in ("recall", "precision", "hr", "ndcg"):
        assert np.allclose(res[k], medium[f"eval_{k}"], atol=1e-4), (k, res[k], medium[f"eval_{k}"])


def test_k3_masked_items_fill_the_tail():
    """A user who interacted with all but 3 items: top-5 = the 3 free items then masked ids in id order with
    score -1024 (basic_test.py:47)."""
    from tagrec_b200.eval_ops import topk_scores
    I = 300
    ut = torch.randn(2, 64, device=dev()) * 0.1          # small dots: fp32 sigmoid stays strictly monotone
    it = torch.randn(I, 64, device=dev()) * 0.1
    free = [7, 100, 250]
    train0 = [i for i in range(I) if i not in free]
    ptr_ = torch.tensor([0, len(train0), len(train0)], device=dev())
    items = torch.tensor(train0, device=dev(), dtype=torch.int32)
    ids, sc = topk_scores(torch.tensor([0, 1], device=dev()), ut, it, ptr_, items, 5)
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    assert set(ids[0, :3]) == set(free) and list(ids[0, 3:]) == [0, 1]
    assert np.all(sc[0, 3:] == -1024.0) and np.all(sc[0, :3] > 0)
    dense = torch.sigmoid(ut[1] @ it.T).cpu().numpy()
    assert list(ids[1]) == list(np.lexsort((np.arange(I), -dense))[:5])


@pytest.mark.parametrize("nu,n_item,dim,k", [(1, 50, 64, 20), (130, 5000, 64, 20), (70, 1000, 256, 100), (64, 129, 32, 5)])
def test_k3_shapes_vs_oracle(nu, n_item, dim, k):
    from tagrec_b200.eval_ops import topk_scores
    g = torch.Generator().manual_seed(nu + n_item)
    ut = torch.randn(nu + 3, dim, generator=g)
    it = torch.randn(n_item, dim, generator=g)
    rng = np.random.RandomState(0)
    users = rng.permutation(nu + 3)[:nu]
    train = {int(u): sorted(rng.choice(n_item, rng.randint(0, min(30, n_item - k)), replace=False).tolist())
             for u in range(nu + 3)}
    ptr_, items = T.bpr_training_data.user_items_to_csr(train, nu + 3)
    ids, _ = topk_scores(torch.tensor(users, device=dev()), ut.to(dev()), it.to(dev()), torch.tensor(ptr_, device=dev()),
                         torch.tensor(items, device=dev()).int(), k)
    scores = (ut[users].double() @ it.double().T).numpy()
    ms = OM.mask_train(scores, users, ptr_, items)
    ref = OM.topk_ids(ms, k)
    ids = ids.cpu().numpy()
    for r in range(nu):
        assert near_tie_ok(ms[r], ids[r], ref[r], k), f"row {r}"


def _random_eval_case(nu, n_item, k, seed, scale=1.0, heavy=False, n_tables=None):
    g = torch.Generator().manual_seed(seed)
    n_tab = n_tables or (nu + 5)
    ut = torch.randn(n_tab, 64, generator=g) * scale
    it = torch.randn(n_item, 64, generator=g) * scale
    rng = np.random.RandomState(seed)
    users = rng.permutation(n_tab)[:nu]
    train = {}
    for u in range(n_tab):
        hi = max(1, min(n_item - k, (n_item // 2) if (heavy and u % 7 == 0) else 40))
        train[u] = sorted(rng.choice(n_item, rng.randint(0, hi), replace=False).tolist())
    ptr_, items = T.bpr_training_data.user_items_to_csr(train, n_tab)
    return (torch.tensor(users, device=dev()), ut.to(dev()), it.to(dev()), torch.tensor(ptr_, device=dev()),
            torch.tensor(items, device=dev()).int())


@pytest.mark.parametrize("nu,n_item,k,scale,heavy", [
    (1, 50, 20, 1.0, False), (64, 129, 5, 1.0, False), (130, 5000, 20, 1.0, True), (300, 40000, 20, 0.1, True),
    (129, 1025, 64, 1.0, False), (257, 3000, 100, 1.0, False), (700, 20000, 10, 3.0, False)])
def test_k3_tensor_core_path_equals_fp32_path(nu, n_item, k, scale, heavy):
    """tcgen05 TF32 filter + exact fp32 re-score (csrc/eval_tc.cu) returns the SAME ids and scores as the exact
    fp32 CUDA-core path — bit-exact, including item-id tie-breaks, partial tiles, several item splits, users with
    very long train rows and both CTA shapes (1 / 2 user halves)."""
    from tagrec_b200.eval_ops import topk_scores
    users, ut, it, ptr_, items = _random_eval_case(nu, n_item, k, 1000 + nu + n_item, scale, heavy)
    ids_a, sc_a = topk_scores(users, ut, it, ptr_, items, k, path="fp32")
    ids_b, sc_b = topk_scores(users, ut, it, ptr_, items, k, path="tf32")
    torch.cuda.synchronize()
    assert torch.equal(ids_a, ids_b)
    assert torch.equal(sc_a, sc_b)


def test_k3_tensor_core_ties_and_masked_tail():
    """Duplicate item rows (exact score ties -> lower id first) and a user with fewer than k un-masked items."""
    from tagrec_b200.eval_ops import topk_scores
    I = 1000
    g = torch.Generator().manual_seed(5)
    it = torch.randn(I, 64, generator=g) * 0.1
    it[500:1000] = it[0:500]                                   # every item has an exact twin 500 ids later
    ut = torch.randn(3, 64, generator=g) * 0.1
    free = [3, 503, 777]
    train0 = [i for i in range(I) if i not in free]
    ptr_ = torch.tensor([0, len(train0), len(train0), len(train0)], device=dev())
    items = torch.tensor(train0, device=dev(), dtype=torch.int32)
    users = torch.tensor([0, 1, 2], device=dev())
    ids, sc = topk_scores(users, ut.to(dev()), it.to(dev()), ptr_, items, 8, path="tf32")
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    dense = (ut.double() @ it.double().T).numpy()
    assert set(ids[0, :3]) == set(free) and list(ids[0, 3:]) == [0, 1, 2, 4, 5]
    assert np.all(sc[0, 3:] == -1024.0)
    for r in (1, 2):
        want = np.lexsort((np.arange(I), -dense[r]))[:8]
        assert list(ids[r]) == list(want)
        assert all(ids[r, j] + 500 == ids[r, j + 1] for j in range(0, 8, 2))      # twins adjacent, lower id first


def test_k3_tensor_core_large_vs_torch():
    """16 K-item / 2 K-user case against torch fp64 scores: sets equal up to near-ties."""
    from tagrec_b200.eval_ops import topk_scores
    users, ut, it, ptr_, items = _random_eval_case(2000, 16000, 20, 77, 0.3, True, n_tables=2100)
    ids, _ = topk_scores(users, ut, it, ptr_, items, 20, path="tf32")
    scores = (ut[users].double() @ it.double().T).cpu().numpy()
    ms = OM.mask_train(scores, users.cpu().numpy(), ptr_.cpu().numpy(), items.cpu().numpy())
    ref = OM.topk_ids(ms, 20)
    ids = ids.cpu().numpy()
    for r in range(len(users)):
        assert near_tie_ok(ms[r], ids[r], ref[r], 20), f"row {r}"


# ------------------------------------------------------------------------------------------------------- sampler
def test_device_sampler_properties(medium):
    U, I, _, _ = nums(medium)
    T.set_config("lightgcn", train_batch=64, sampler="device", device=dev(), seed=7)
    d = make_data(medium)
    s = T.BPR_training_data(d, None)
    a = s.all_train_data.cpu().numpy()
    s.reset()
    b = s.all_train_data.cpu().numpy()
    e = medium["edge_index_train"]
    # every positive edge exactly once per epoch (a permutation), negatives never in the user's train set
    key = lambda x: np.sort(x[:, 0] * I + x[:, 1])
    assert np.array_equal(key(a), key(e)) and np.array_equal(key(b), key(e))
    train = d.user_items["train"]
    assert all(n not in train[u] for u, _, n in a) and all(0 <= n < I for _, _, n in a)
    assert not np.array_equal(a, b) and not np.array_equal(a[:, :2], e)      # reshuffled and re-drawn every epoch
    # negatives are uniform over the allowed items: mean id close to the allowed mean
    assert abs(a[:, 2].mean() / I - 0.5) < 0.05


# ---------------------------------------------------------------------------------------------------------- Adam
def test_fused_adam_vs_torch():
    from tagrec_b200._lib import lib, ptr, stream_ptr, check
    p = torch.randn(1000, 64, device=dev())
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=0.01)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn_like(p)
        ref.grad = g.clone()
        opt.step()
        check(lib().tagrec_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), 0.01, 0.9, 0.999, 1e-8, 0.0, step,
                                     stream_ptr()), "adam")
    assert relerr(p.cpu().numpy(), ref.detach().cpu().numpy()) < 1e-6


# ----------------------------------------------------------------------------------------------------- multi-GPU
def test_multi_gpu_sharded_step_matches_single():
    """Only on boxes with >= 2 GPUs (gpurun --gpus 2): torchrun tests/multi_gpu_check.py."""
    import os
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
