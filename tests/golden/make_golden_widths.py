#!/usr/bin/env python
"""Goldens for the reference's DEFAULT layer widths, dim_layer_list = [64, 32, 16] (utility/utils.py:39), on the tiny
dataset of make_golden.py: NGCF (176-d concatenated tables) and TGCN (64 -> 64 -> 32 -> 16, also 176-d), produced by the
unmodified reference classes.  Build container only.   python tests/golden/make_golden_widths.py -> tiny_widths.npz
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
import torch  # noqa: E402
from make_golden import NGCF, TGCN, Args, _batch, init_seed, set_cfg  # noqa: E402


def run(d, name, cls, tag, extra):
    set_cfg(name, use_tag=True, reg=1e-3, **extra)
    init_seed(2020)
    with np.errstate(divide="ignore"):
        m = cls(d)
    out = {}
    if name == "tgcn":
        names = ["ui", "ut", "iu", "it", "tu", "ti"]
        for n, (idx, w) in zip(names, m.all_sample):
            out[f"{tag}_nbr_{n}"] = np.asarray(idx, dtype=np.int64)
            out[f"{tag}_nbw_{n}"] = np.asarray(w, dtype=np.int64)
    for k, v in m.state_dict().items():
        out[f"{tag}_param_{k}"] = v.detach().numpy().copy()
    m.train()
    for k, t in enumerate(m.forward()):
        out[f"{tag}_fwd_{k}"] = t.detach().numpy().copy()
    batch = _batch(d, np.random.RandomState(7), 48)
    out[f"{tag}_batch"] = batch
    lossx = m.loss(torch.tensor(batch, dtype=torch.long))
    out[f"{tag}_loss"] = np.array([x.item() for x in lossx], dtype=np.float64)
    m.zero_grad()
    sum(lossx).backward()
    for k, p in m.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        out[f"{tag}_grad_{k}"] = g.detach().numpy().copy()
    m.eval()
    users = torch.tensor([0, 3, 5, d.num["user"] - 1, 3], dtype=torch.long)
    out[f"{tag}_pred_users"] = users.numpy()
    with torch.no_grad():
        out[f"{tag}_pred"] = m.predict_rating(users).numpy().copy()
    # float64 truth: the same reference class, same parameters and batch, run in double.  The weight / bias gradients
    # are float32 sums over every node on both sides; the parity bar for them is stated against this truth.
    torch.set_default_dtype(torch.float64)
    try:
        import copy
        m64 = copy.deepcopy(m).double()
        if hasattr(m64, "norm_adj") and torch.is_tensor(m64.norm_adj):
            m64.norm_adj = m64.norm_adj.double()
        m64.zero_grad()
        m64.train()
        lossx = m64.loss(torch.tensor(batch, dtype=torch.long))
        out[f"{tag}_loss64"] = np.array([x.item() for x in lossx], dtype=np.float64)
        sum(lossx).backward()
        for k, p in m64.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            out[f"{tag}_grad64_{k}"] = g.detach().numpy().copy()
            e = np.abs(out[f"{tag}_grad_{k}"].astype(np.float64) - out[f"{tag}_grad64_{k}"]).max()
            print(f"{tag} {k}: max|g| {np.abs(out[f'{tag}_grad64_{k}']).max():.3e}  reference fp32 vs fp64 {e / max(np.abs(out[f'{tag}_grad64_{k}']).max(), 1e-300):.2e}")
    finally:
        torch.set_default_dtype(torch.float32)
    return out


def main():
    os.chdir("/tmp")
    d = MG.make_dataset(seed=1, U=40, I=60, T=25, n_edge=420, n_uit=500)        # == the dataset of tiny.npz
    from data.tgcn_load import TGCN_load
    d.args = Args()
    d.cpu_core = 1
    d.get_all_neighbor = types.MethodType(TGCN_load.get_all_neighbor, d)
    out = {}
    out.update(run(d, "ngcf", NGCF, "ngcf_w", dict(dim_layer_list=[64, 32, 16])))
    out.update(run(d, "tgcn", TGCN, "tgcn_w", dict(dim_layer_list=[64, 32, 16], neighbor_k=5)))
    p = os.path.join(HERE, "tiny_widths.npz")
    np.savez_compressed(p, **out)
    print(p, os.path.getsize(p) // 1024, "KiB")


if __name__ == "__main__":
    main()
