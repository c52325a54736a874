#!/usr/bin/env python
"""KGAT goldens from the unmodified reference (model/kgat.py, train_data/transe_training_data.py) on the tiny dataset
of make_golden.py.  Two variants:
  kgat_stock   the stock overlay (agg_type 'bi_agg', utility/config.py:54-60) with TGCN_load.create_edge's [2, E] edge
               arrays: the reference then returns the ego tables (kgat.py:99) — forward / loss / transe_loss / every
               gradient, plus the first batches of KGAT_training_data (numpy seed 2020);
  kgat_inter   agg_type 'bi_inter' with [E, 2] edge arrays: the intended model (relation-aware attention -> row softmax
               -> bi-interaction layers) — forward (256-d), loss, every gradient (attention NOT detached).
Build container only.   python tests/golden/make_golden_kgat.py -> tiny_kgat.npz
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402
import torch  # noqa: E402
from make_golden import Args, _batch, init_seed, set_cfg  # noqa: E402
from model.kgat import KGAT  # noqa: E402
from train_data.transe_training_data import KGAT_training_data  # noqa: E402
from data.tgcn_load import TGCN_load  # noqa: E402


def run(d, tag, agg_type, edges_e2):
    set_cfg("kgat", use_tag=True, reg=1e-3, cor_reg=1e-3, agg_type=agg_type, dim_layer_list=[64, 64, 64], transe_batch=32)
    init_seed(2020)
    stock = types.MethodType(TGCN_load.create_edge, d)
    d.create_edge = (lambda: {k: np.ascontiguousarray(v.T) for k, v in stock().items()}) if edges_e2 else stock
    m = KGAT(d)
    out = {}
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
    if not edges_e2:
        d.create_edge = stock
        np.random.seed(2020)
        td = KGAT_training_data(d, Args())
        batches = []
        for i, b in enumerate(td.mini_batch()):
            batches.append(b.numpy().copy())
            if i == 2:
                break
        out[f"{tag}_kg_batches"] = np.stack(batches)
        out[f"{tag}_kg_tot_inter"] = np.array([td.tot_inter], dtype=np.int64)
        lossx = m.transe_loss(torch.tensor(batches[0], dtype=torch.long))
        out[f"{tag}_transe_loss"] = np.array([x.item() for x in lossx], dtype=np.float64)
        m.zero_grad()
        sum(lossx).backward()
        for k, p in m.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            out[f"{tag}_transe_grad_{k}"] = g.detach().numpy().copy()
    return out


def main():
    os.chdir("/tmp")
    d = MG.make_dataset(seed=1, U=40, I=60, T=25, n_edge=420, n_uit=500)        # == the dataset of tiny.npz
    out = {}
    out.update(run(d, "kgat_stock", "bi_agg", False))
    out.update(run(d, "kgat_inter", "bi_inter", True))
    p = os.path.join(HERE, "tiny_kgat.npz")
    np.savez_compressed(p, **out)
    print(p, os.path.getsize(p) // 1024, "KiB")
    print({k: v for k, v in out.items() if k.endswith("_loss")})


if __name__ == "__main__":
    main()
