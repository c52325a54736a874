"""Pins the CPU oracle against outputs of the UNMODIFIED reference (tests/golden/*.npz, made by make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import adjacency as OA
from oracle import metrics as OM
from oracle import propagation as OP
from oracle import sampler as OS
from helpers import blocks, coalesced, nums, relerr, user_lists

NORMS = ("bi_norm", "si_norm", "si_norm_self", "ngcf", "plain")


def test_reference_coo_order_fact(tiny):
    """Documented reference behaviour: the COO creat_adj emits is row-major, but the column order INSIDE a row is
    a scipy artefact (ascending after an even number of sparse products, descending after an odd number);
    torch.sparse.mm coalesces (sorts by row, col) before multiplying, so the canonical form is ascending."""
    for nt, asc in (("bi_norm", True), ("plain", True), ("si_norm", False)):
        r, c = tiny[f"adj_ui_{nt}_row"], tiny[f"adj_ui_{nt}_col"]
        assert np.all(np.diff(r) >= 0)
        same_row = np.diff(r) == 0
        d = np.diff(c)[same_row]
        assert np.all(d > 0) if asc else np.all(d < 0)


@pytest.mark.parametrize("use_tag", [False, True])
@pytest.mark.parametrize("nt", NORMS)
def test_csr_bit_exact(tiny, use_tag, nt):
    U, I, T, _ = nums(tiny)
    ui, ut, it = blocks(tiny)
    n, rowptr, col, val = OA.creat_adj(U, I, ui, nt, T, ut if use_tag else None, it if use_tag else None)
    tag = f"adj_{'uit' if use_tag else 'ui'}_{nt}"
    assert n == U + I + (T if use_tag else 0)
    grow, gcol, gval = coalesced(n, tiny[tag + "_row"], tiny[tag + "_col"], tiny[tag + "_val"])
    assert np.array_equal(OA.row_ids(rowptr), grow)
    assert np.array_equal(col, gcol)
    assert val.dtype == np.float32
    assert np.array_equal(val.view(np.uint32), gval.view(np.uint32)), "values not bit-exact"


def test_row_folds(tiny):
    U, I, _, _ = nums(tiny)
    n, rowptr, _, _ = OA.creat_adj(U, I, blocks(tiny)[0], "bi_norm")
    folds = OA.fold_rows(n, 3)
    assert [b - a for a, b in folds] == list(tiny["adj_fold3_rows"])
    assert [int(rowptr[b] - rowptr[a]) for a, b in folds] == list(tiny["adj_fold3_nnz"])


def _lightgcn_case(g, tag, use_tag, dtype):
    U, I, T, _ = nums(g)
    ui, ut, it = blocks(g)
    n, rowptr, col, val = OA.creat_adj(U, I, ui, "bi_norm", T, ut if use_tag else None, it if use_tag else None)
    names = [f"{tag}_param_embed.{k}" for k in range(3 if use_tag else 2)]
    e0 = torch.cat([torch.tensor(g[k]) for k in names]).to(dtype)
    return (rowptr, col, val), e0, U, I


@pytest.mark.parametrize("tag,use_tag,kind", [("lgcn", False, "softplus"), ("lgcn_tag", True, "softplus"),
                                               ("lgcn_logsig", False, "logsigmoid")])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_lightgcn_forward_loss_grad(tiny, tag, use_tag, kind, dtype):
    csr, e0, U, I = _lightgcn_case(tiny, tag, use_tag, dtype)
    loss, reg, g0, final = OP.lightgcn_loss_and_grad(csr, e0, tiny[f"{tag}_batch"], U, 3, 1e-3, kind)
    ref_final = np.concatenate([tiny[f"{tag}_fwd_{k}"] for k in range(3 if use_tag else 2)])
    assert relerr(final.numpy(), ref_final) < 2e-6
    assert abs(loss.item() - tiny[f"{tag}_loss"][0]) < 1e-6 * abs(tiny[f"{tag}_loss"][0])
    assert abs(reg.item() - tiny[f"{tag}_loss"][1]) < 1e-5 * abs(tiny[f"{tag}_loss"][1])
    ref_g = np.concatenate([tiny[f"{tag}_grad_embed.{k}"] for k in range(3 if use_tag else 2)])
    assert relerr(g0.numpy(), ref_g) < 5e-6


def test_lightgcn_predict(tiny):
    csr, e0, U, I = _lightgcn_case(tiny, "lgcn", False, torch.float32)
    final, _ = OP.lightgcn_forward(csr, e0, 3)
    r = OP.predict_rating(final[:U], final[U:], tiny["lgcn_pred_users"])
    assert relerr(r.numpy(), tiny["lgcn_pred"]) < 1e-6


@pytest.mark.parametrize("tag,use_tag", [("ngcf", False), ("ngcf_tag", True)])
def test_ngcf_forward_and_autograd(tiny, tag, use_tag):
    U, I, T, _ = nums(tiny)
    ui, ut, it = blocks(tiny)
    n, rowptr, col, val = OA.creat_adj(U, I, ui, "ngcf", T, ut if use_tag else None, it if use_tag else None)
    nemb = 3 if use_tag else 2
    embs = [torch.tensor(tiny[f"{tag}_param_embed.{k}"], requires_grad=True) for k in range(nemb)]
    mats = {k.split("mat.")[1]: torch.tensor(v, requires_grad=True) for k, v in tiny.items()
            if k.startswith(f"{tag}_param_mat.")}
    out = OP.ngcf_forward((rowptr, col, val), torch.cat(embs), mats, 3)
    ref = np.concatenate([tiny[f"{tag}_fwd_{k}"] for k in range(nemb)])
    assert out.shape[1] == 256
    assert relerr(out.detach().numpy(), ref) < 2e-6
    b = torch.tensor(tiny[f"{tag}_batch"])
    fu, fp, fn = out[b[:, 0]], out[U + b[:, 1]], out[U + b[:, 2]]
    loss = OP.bpr_loss(fu, fp, fn, "logsigmoid")
    reg = 1e-3 * OP.l2reg(fu, fp, fn)                      # ngcf.py:102-103 reg on PROPAGATED rows
    assert abs(loss.item() - tiny[f"{tag}_loss"][0]) < 1e-6
    assert abs(reg.item() - tiny[f"{tag}_loss"][1]) < 1e-7
    (loss + reg).backward()
    for k in range(nemb):
        assert relerr(embs[k].grad.numpy(), tiny[f"{tag}_grad_embed.{k}"]) < 1e-5
    for k, v in mats.items():
        assert relerr(v.grad.numpy(), tiny[f"{tag}_grad_mat.{k}"]) < 1e-5


def test_mt19937_primitives(tiny):
    rng = OS.MT19937(99)
    assert [OS.randint(rng, 0, 1000) for _ in range(64)] == list(tiny["sampler_randint_1000"])
    assert [OS.randint(rng, 0, 17632) for _ in range(64)] == list(tiny["sampler_randint_17632"])
    assert np.array_equal(OS.shuffle_index(rng, 50), tiny["sampler_shuffle_50"])
    # and against the live numpy of this host
    np.random.seed(12345)
    rng = OS.MT19937(12345)
    assert [np.random.randint(0, 77) for _ in range(200)] == [OS.randint(rng, 0, 77) for _ in range(200)]


@pytest.mark.parametrize("name", ["tiny", "medium"])
def test_sampler_bit_exact(name, request):
    g = request.getfixturevalue(name)
    _, I, _, _ = nums(g)
    train = user_lists(g, "train")
    rng = OS.MT19937(2020)
    first = OS.sample_epoch(rng, g["edge_index_train"], train, I)
    second = OS.sample_epoch(rng, g["edge_index_train"], train, I)
    assert np.array_equal(first, g["sampler_first"])
    assert np.array_equal(second, g["sampler_second"])
    sizes = [b - a for a, b in OS.mini_batches(len(second), 64)]
    assert sizes == list(g["sampler_batch_sizes"])


def test_minibatch_tail_quirk():
    assert OS.mini_batches(212, 64) == [(0, 64), (64, 128), (128, 212), (192, 212)]
    assert OM.minibatch_slices(32, 16) == [(0, 16), (16, 32), (32, 32)]      # the empty batch of SURVEY A13


def test_eval_metrics_tiny(tiny):
    ms = tiny["eval_masked_scores"]
    users = tiny["eval_users"]
    res, _ = OM.epoch_test(ms, users, tiny["test_ptr"], tiny["test_items"], list(tiny["eval_topks"]))
    for k in ("recall", "precision", "hr", "ndcg", "auc"):
        assert np.allclose(res[k], tiny[f"eval_{k}"], rtol=0, atol=1e-7 if k == "precision" else 1e-12), k


def test_eval_pipeline_medium(medium):
    """forward -> predict -> mask -> top-K -> metrics, all oracle, vs the reference's epoch_test."""
    U, I, _, _ = nums(medium)
    n, rowptr, col, val = OA.creat_adj(U, I, blocks(medium)[0], "bi_norm")
    e0 = torch.cat([torch.tensor(medium["lgcn_param_embed.0"]), torch.tensor(medium["lgcn_param_embed.1"])])
    final, _ = OP.lightgcn_forward((rowptr, col, val), e0, 3)
    users = medium["eval_users"]
    scores = OP.predict_rating(final[:U], final[U:], users).numpy()
    ms = OM.mask_train(scores, users, medium["train_ptr"], medium["train_items"])
    ks = list(medium["eval_topks"])
    res, top = OM.epoch_test(ms, users, medium["test_ptr"], medium["test_items"], ks)
    for k in ("recall", "precision", "hr", "ndcg"):
        assert np.allclose(res[k], medium[f"eval_{k}"], atol=1e-4), k          # north_star: Recall/NDCG@20 1e-4
    assert abs(res["auc"][0] - medium["eval_auc"][0]) < 1e-4
    # top-20 id sets vs the (-score,id) order of the reference's own scores
    ref = medium["eval_top40_ids"][:, :20]
    same = sum(set(a) == set(b) for a, b in zip(top[:, :20], ref))
    assert same >= len(users) - 2, f"{len(users) - same} users differ in top-20 set"


# ------------------------------------------------------------------------------ DGCF / DisenGCN / TGCN oracles
def _edges(g, use_tag):
    U, I, Tg, _ = nums(g)
    ui, ut, it = blocks(g)
    n, rowptr, col, _ = OA.creat_adj(U, I, ui, "plain", Tg if use_tag else 0, ut if use_tag else None,
                                     it if use_tag else None)
    head = torch.as_tensor(np.repeat(np.arange(n), np.diff(rowptr)), dtype=torch.int64)
    return n, head, torch.as_tensor(np.asarray(col), dtype=torch.int64)


def _bpr_autograd(final_u, final_i, reg_u, reg_i, batch, reg, kind):
    b = torch.as_tensor(batch, dtype=torch.int64)
    loss = OP.bpr_loss(final_u[b[:, 0]], final_i[b[:, 1]], final_i[b[:, 2]], kind)
    return loss, reg * OP.l2reg(reg_u[b[:, 0]], reg_i[b[:, 1]], reg_i[b[:, 2]])


def test_dgcf_oracle_vs_reference(tiny):
    """oracle.routing.dgcf_forward + autograd == model/dgcf.py (forward, loss tuple, embedding gradients)."""
    from oracle import routing as OR
    U, I, _, _ = nums(tiny)
    n, head, tail = _edges(tiny, False)
    emb = [torch.tensor(tiny[f"dgcf_param_embed.{k}"], requires_grad=True) for k in range(2)]
    final = OR.dgcf_forward(head, tail, n, torch.cat(emb), 3, 2)
    fu, fi = final[:U], final[U:]
    assert relerr(fu.detach().numpy(), tiny["dgcf_fwd_0"]) < 1e-6 and relerr(fi.detach().numpy(), tiny["dgcf_fwd_1"]) < 1e-6
    loss, reg = _bpr_autograd(fu, fi, emb[0], emb[1], tiny["dgcf_batch"], 1e-3, "softplus")
    assert abs(loss.item() - tiny["dgcf_loss"][0]) < 1e-6 and abs(reg.item() - tiny["dgcf_loss"][1]) < 1e-9
    (loss + reg).backward()
    for k in range(2):
        assert relerr(emb[k].grad.numpy(), tiny[f"dgcf_grad_embed.{k}"]) < 1e-5


def test_disengcn_oracle_vs_reference(tiny):
    """oracle.routing.disengcn_forward == model/disengcn.py on the tripartite graph; gradients against the float64
    run of the reference (the float32 ones are ill-conditioned, see tests/golden/make_golden_fp64.py)."""
    from oracle import routing as OR
    U, I, Tg, _ = nums(tiny)
    n, head, tail = _edges(tiny, True)
    truth = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "routing_fp64.npz")))
    for dtype, fwd_tol in ((torch.float32, 1e-6), (torch.float64, 1e-12)):
        emb = [torch.tensor(tiny[f"disengcn_param_embed.{k}"], dtype=dtype, requires_grad=True) for k in range(3)]
        ws = [(torch.tensor(tiny[f"disengcn_param_layer.{k}.W"], dtype=dtype, requires_grad=True),
               torch.tensor(tiny[f"disengcn_param_layer.{k}.b"], dtype=dtype, requires_grad=True)) for k in range(3)]
        final = OR.disengcn_forward(head, tail, n, torch.cat(emb), ws, 2)
        parts = torch.split(final, [U, I, Tg])
        want = [truth[f"disengcn_fwd64_{k}"] if dtype == torch.float64 else tiny[f"disengcn_fwd_{k}"] for k in range(3)]
        for k in range(3):
            assert relerr(parts[k].detach().numpy(), want[k]) < fwd_tol
        loss, reg = _bpr_autograd(parts[0], parts[1], parts[0], parts[1], tiny["disengcn_batch"], 1e-3, "softplus")
        (loss + reg).backward()
        if dtype == torch.float64:
            for k in range(3):
                assert relerr(emb[k].grad.numpy(), truth[f"disengcn_grad64_embed.{k}"]) < 1e-9
                assert relerr(ws[k][0].grad.numpy(), truth[f"disengcn_grad64_layer.{k}.W"]) < 1e-9
                assert relerr(ws[k][1].grad.numpy(), truth[f"disengcn_grad64_layer.{k}.b"]) < 1e-9


def test_tgcn_oracle_vs_reference(tiny, tiny_tgcn):
    """oracle.tgcn.tgcn_forward + autograd == model/tgcn.py with the reference's own neighbour tables."""
    from oracle import tgcn as OT
    U = nums(tiny)[0]
    P = {k[len("tgcn_param_"):]: torch.tensor(v, requires_grad=True) for k, v in tiny_tgcn.items()
         if k.startswith("tgcn_param_")}
    tables = [(tiny_tgcn[f"tgcn_nbr_{n}"], tiny_tgcn[f"tgcn_nbw_{n}"]) for n in ("ui", "ut", "iu", "it", "tu", "ti")]
    fu, fi, ft = OT.tgcn_forward(P, tables, 2, 5)
    for k, t in enumerate((fu, fi, ft)):
        assert relerr(t.detach().numpy(), tiny_tgcn[f"tgcn_fwd_{k}"]) < 1e-6
    loss, reg = _bpr_autograd(fu, fi, fu, fi, tiny_tgcn["tgcn_batch"], 1e-3, "logsigmoid")
    assert abs(loss.item() - tiny_tgcn["tgcn_loss"][0]) < 1e-6 and abs(reg.item() - tiny_tgcn["tgcn_loss"][1]) < 1e-8
    (loss + reg).backward()
    for name, p in P.items():
        want = tiny_tgcn[f"tgcn_grad_{name}"]
        got = p.grad.numpy() if p.grad is not None else np.zeros_like(want)
        assert relerr(got, want) < 2e-4, name            # O(1e-9) second-layer attention gradients: fp32 noise
