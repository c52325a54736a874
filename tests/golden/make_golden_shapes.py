#!/usr/bin/env python
"""Goldens on the BASELINE.json shapes (C1-C4), produced by RUNNING THE UNMODIFIED REFERENCE on CPU.

    C1  LightGCN            LastFM-shaped        1 892 x 17 632, ~93 K interactions      (BASELINE configs[0])
    C2  LightGCN use_tag    Delicious-tags-shaped tripartite 1 867 x 69 223 x 40 897      (configs[1])
    C2  TGCN (k = 25, [64,64])  same graph                                                (configs[1])
    C3  NGCF [64,64,64]     Amazon-book-shaped   52 643 x 91 599, ~3 M interactions       (configs[2])
    C4  DGCF                Gowalla-shaped       29 858 x 40 981, ~1 M interactions       (configs[3])

Runs only in the build container (needs /root/reference, read-only; same shims as make_golden.py).  The datasets are
NOT stored: they come from tagrec_b200.data.synth_named(name, seed=2020) (numpy RandomState — the parity test on the
GPU box regenerates them and checks the stored edge checksum), and the initial parameters come from
torch.manual_seed(2020) + Xavier in the reference's creation order (checked through stored checksums).  To keep the
fixtures small, every [N, d] table (propagated embeddings, embedding gradients) is stored as 256 sampled rows plus
float64 checksums (sum, sum of squares, sum of |x|) over the whole table; small parameter gradients are stored whole.

    python tests/golden/make_golden_shapes.py [c1 c2 c2_tgcn c3 c4]       # writes tests/golden/shape_<name>.npz
"""
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
_argv = sys.argv[1:]
import make_golden as MG  # noqa: E402  (installs the shims, imports the reference)
import torch  # noqa: E402
from make_golden import BT, CFG, DGCF, LightGCN, NGCF, TGCN, Args, _batch, init_seed, set_cfg  # noqa: E402

torch.set_num_threads(os.cpu_count())
SAMPLE_ROWS = 256
BATCH = 2048


def load_data_module():
    """tagrec_b200/data.py alone (numpy/scipy only) — importing the package would pull the CUDA library in."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("tagrec_data", os.path.join(ROOT, "tag-aware-recommendation_b200", "data.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


DATA = load_data_module()


def checksums(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([a.sum(), (a * a).sum(), np.abs(a).sum()], dtype=np.float64)


def edge_checksum(ds):
    e = ds.edge_index["train"].astype(np.int64)
    out = [len(e), int((e[:, 0] * 1000003 + e[:, 1]).sum() % (1 << 61))]
    if ds.uit_data is not None:
        t = ds.uit_data.astype(np.int64)
        out += [len(t), int((t[:, 0] * 1000003 + t[:, 1] * 10007 + t[:, 2]).sum() % (1 << 61))]
    return np.array(out, dtype=np.int64)


def table_entry(out, key, arr, rng):
    arr = np.asarray(arr)
    out[key + "_shape"] = np.array(arr.shape, dtype=np.int64)
    out[key + "_sums"] = checksums(arr)
    if arr.ndim == 2 and arr.shape[0] > 4 * SAMPLE_ROWS:
        rows = np.sort(rng.choice(arr.shape[0], SAMPLE_ROWS, replace=False))
        out[key + "_rows"] = rows.astype(np.int64)
        out[key + "_vals"] = arr[rows].copy()
    else:
        out[key + "_full"] = arr.copy()


def truth64(m, bt, tuple_batch, out):
    """The SAME reference model (same float32 parameter values, same adjacency values) run once more in float64:
    the yardstick for gradients that are long float32 reductions on both sides (weight gradients summed over 1e5
    nodes, scatter-added embedding gradients).  Stored for the rows / tensors the float32 golden stores."""
    import copy
    t0 = time.time()
    torch.set_default_dtype(torch.float64)
    try:
        m64 = copy.deepcopy(m).double()
        if hasattr(m64, "norm_adj") and torch.is_tensor(m64.norm_adj):
            m64.norm_adj = m64.norm_adj.double()
        m64.train()
        m64.zero_grad()
        lossx = m64.loss((bt, None)) if tuple_batch else m64.loss(bt)
        out["loss64"] = np.array([x.item() for x in lossx], dtype=np.float64)
        sum(lossx).backward()
        for k, p in m64.named_parameters():
            g = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().numpy()
            if f"grad_{k}_rows" in out:
                out[f"grad64_{k}_vals"] = g[out[f"grad_{k}_rows"]].copy()
            else:
                out[f"grad64_{k}_full"] = g.copy()
            ref32 = out.get(f"grad_{k}_vals", out.get(f"grad_{k}_full"))
            g64 = out.get(f"grad64_{k}_vals", out.get(f"grad64_{k}_full"))
            scale = np.abs(g64).max()
            if scale > 0:
                print(f"    {k}: reference fp32 vs fp64 {np.abs(ref32 - g64).max() / scale:.2e}", flush=True)
    finally:
        torch.set_default_dtype(torch.float32)
    print(f"  float64 pass {time.time() - t0:.1f} s", flush=True)


def run_model(ds, name, cls, use_tag, tag, extra=None, tuple_batch=False, with64=False):
    """loss / grads / propagated tables / top-20 of 512 users from the reference class at its seeded initial state."""
    t0 = time.time()
    set_cfg(name, use_tag=use_tag, reg=1e-4, train_batch=BATCH, test_batch=512, topks=[20], **(extra or {}))
    init_seed(2020)
    with np.errstate(divide="ignore"):
        m = cls(ds)
    rng = np.random.RandomState(123)
    out = {"edges": edge_checksum(ds)}
    for k, v in m.state_dict().items():
        out[f"param_{k}_sums"] = checksums(v.detach().numpy())
    m.train()
    fw = m.forward()
    for k, t in enumerate(fw):
        table_entry(out, f"fwd_{k}", t.detach().numpy(), rng)
    batch = _batch(ds, np.random.RandomState(7), BATCH)
    out["batch"] = batch
    bt = torch.tensor(batch, dtype=torch.long)
    lossx = m.loss((bt, None)) if tuple_batch else m.loss(bt)
    out["loss"] = np.array([x.item() for x in lossx], dtype=np.float64)
    m.zero_grad()
    sum(lossx).backward()
    for k, p in m.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        table_entry(out, f"grad_{k}", g.detach().numpy(), rng)
    if with64:
        truth64(m, bt, tuple_batch, out)
    # top-20 of 512 users (dict-key order prefix of the test users), masked like basic_test.py:37-47, (-score, id) order
    m.eval()
    users = list(ds.user_items["test"].keys())[:512]
    with torch.no_grad():
        r = m.predict_rating(torch.tensor(users, dtype=torch.long)).clone()
    for row, u in enumerate(users):
        r[row, ds.user_items["train"].get(u, [])] = -(1 << 10)
    rs = r.numpy()
    order = np.lexsort((np.broadcast_to(np.arange(rs.shape[1]), rs.shape), -rs.astype(np.float64)), axis=1)[:, :24]
    out["top_users"] = np.array(users, dtype=np.int64)
    out["top24_ids"] = order.astype(np.int32)
    out["top24_scores"] = np.take_along_axis(rs, order, 1)
    print(f"  {tag}: model pass {time.time() - t0:.1f} s, loss {out['loss']}", flush=True)
    return out, m


def quantise(p):
    """int8 grid of one table: q in [-127, 127], scale = max|p| / 127 (float32).  The canonical parameter value is
    float32(q) * scale, computed identically by the reference side here and by the parity test."""
    scale = np.float32(np.abs(p).max() / 127.0)
    q = np.clip(np.rint(p / scale), -127, 127).astype(np.int8)
    return q, scale


def trained_state(ds, m, out, epochs=6):
    """A TRAINED-like parameter state that can be stored small: the reference model is trained for a few epochs through
    its own loop (BPR_training_data with cpu_core = 1, training/basic_train.py epoch_training, Adam lr 0.01), then every
    embedding table is snapped to an int8 grid (1 byte per element in the fixture) and the snapped values are loaded back
    into the reference model — the metrics below are the reference's results FOR THESE EXACT PARAMETERS."""
    from training.basic_train import epoch_training
    from train_data.bpr_training_data import BPR_training_data
    t0 = time.time()
    set_cfg("lightgcn", use_tag=False, reg=1e-4, train_batch=BATCH, lr=0.01, cpu_core=1)
    opt = torch.optim.Adam(m.parameters(), lr=0.01)
    sampler = BPR_training_data(ds, Args())
    m.train()
    for ep in range(epochs):
        losses = epoch_training(sampler, m.loss, opt)
    with torch.no_grad():
        for k, p in enumerate(m.embed):
            q, scale = quantise(p.detach().numpy())
            out[f"trained_q_{k}"], out[f"trained_scale_{k}"] = q, np.array(scale, dtype=np.float32)
            p.copy_(torch.from_numpy(q.astype(np.float32) * scale))
    print(f"  trained {epochs} epochs in {time.time() - t0:.1f} s, last epoch mean loss {np.mean(losses):.4f}", flush=True)


def top_all(ds, m, out, prefix):
    """(-score, id)-ordered top-24 of every test user from the reference's masked predict_rating (basic_test.py:37-47)."""
    m.eval()
    users = list(ds.user_items["test"].keys())
    ids, scores = [], []
    with torch.no_grad():
        for s in range(0, len(users), 512):
            ub = users[s:s + 512]
            r = m.predict_rating(torch.tensor(ub, dtype=torch.long)).clone()
            for row, u in enumerate(ub):
                r[row, ds.user_items["train"].get(u, [])] = -(1 << 10)
            rs = r.numpy()
            order = np.lexsort((np.broadcast_to(np.arange(rs.shape[1]), rs.shape), -rs.astype(np.float64)), axis=1)[:, :24]
            ids.append(order.astype(np.int32))
            scores.append(np.take_along_axis(rs, order, 1))
    out[prefix + "_users"] = np.array(users, dtype=np.int64)
    out[prefix + "_top24_ids"], out[prefix + "_top24_scores"] = np.concatenate(ids), np.concatenate(scores)
    rng = np.random.RandomState(5)
    with torch.no_grad():
        for k, t in enumerate(m.forward()):
            table_entry(out, f"{prefix}_fwd_{k}", t.numpy(), rng)


def eval_all(ds, m, out):
    """training/basic_test.py:30-80 over ALL test users (sklearn AUC per user: minutes on the bigger shapes)."""
    t0 = time.time()
    set_cfg("lightgcn", test_batch=512, topks=[10, 20])
    m.eval()
    res = BT.epoch_test(m, ds.user_items["train"], ds.user_items["test"], Args())
    out["eval_topks"] = np.array([10, 20], dtype=np.int64)
    for k, v in res.items():
        out[f"eval_{k}"] = np.array(v, dtype=np.float64)
    print(f"  epoch_test {time.time() - t0:.1f} s: {res}", flush=True)


def trajectory(ds, out, steps=5):
    """`steps` Adam steps of LightGCN through training/basic_train.py's epoch_training body on fixed batches."""
    from training.basic_train import epoch_training
    from train_data.bpr_training_data import BPR_training_data
    set_cfg("lightgcn", use_tag=False, reg=1e-4, train_batch=BATCH, lr=0.001)
    init_seed(2020)
    with np.errstate(divide="ignore"):
        m = LightGCN(ds)
    opt = torch.optim.Adam(m.parameters(), lr=CFG["lr"])
    triples = _batch(ds, np.random.RandomState(11), BATCH * steps)

    class Fixed:
        batch_size = BATCH

        def reset(self):
            self.all_train_data = torch.tensor(triples, dtype=torch.long)

        mini_batch = BPR_training_data.mini_batch

    m.train()
    losses = epoch_training(Fixed(), m.loss, opt)
    out["traj_triples"] = triples
    out["traj_losses"] = np.array(losses, dtype=np.float64)


def save(name, out):
    p = os.path.join(HERE, f"shape_{name}.npz")
    np.savez_compressed(p, **out)
    print(f"{p}: {os.path.getsize(p) // 1024} KiB", flush=True)


def main():
    which = _argv or ["c1", "c2", "c2_tgcn", "c3", "c4"]
    os.chdir("/tmp")
    if "c1" in which:
        print("C1 lastfm / LightGCN", flush=True)
        ds = DATA.synth_named("lastfm")
        out, m = run_model(ds, "lightgcn", LightGCN, False, "c1")
        trained_state(ds, m, out)
        eval_all(ds, m, out)
        top_all(ds, m, out, "trained")
        trajectory(ds, out)
        save("c1_lightgcn", out)
    if "c2" in which or "c2_tgcn" in which:
        ds = DATA.synth_bipartite(seed=2020, **DATA.SHAPES["delicious_tags"])
        if "c2" in which:
            print("C2 delicious_tags / LightGCN use_tag", flush=True)
            out, m = run_model(ds, "lightgcn", LightGCN, True, "c2")
            save("c2_lightgcn_tag", out)
        if "c2_tgcn" in which:
            print("C2 delicious_tags / TGCN", flush=True)
            # neighbour tables: the restated builder capped at the k = 25 columns the model reads (inputs of this
            # golden, regenerated from np.random.seed(2020) by the test); the reference class consumes them unchanged
            def tables(self=None):
                np.random.seed(2020)
                return DATA.get_all_neighbor(ds, width=25)
            ds.get_all_neighbor = tables
            out, m = run_model(ds, "tgcn", TGCN, True, "c2_tgcn", extra=dict(dim_layer_list=[64, 64], neighbor_k=25),
                               with64=True)
            save("c2_tgcn", out)
    if "c3" in which:
        print("C3 amazon_book / NGCF", flush=True)
        ds = DATA.synth_named("amazon_book")
        out, m = run_model(ds, "ngcf", NGCF, False, "c3", with64=True)
        save("c3_ngcf", out)
    if "c4" in which:
        print("C4 gowalla / DGCF", flush=True)
        ds = DATA.synth_named("gowalla")
        out, m = run_model(ds, "dgcf", DGCF, False, "c4", tuple_batch=True)
        save("c4_dgcf", out)


if __name__ == "__main__":
    main()
