"""Golden stream of the reference's DGCF_training_data (train_data/bpr_training_data.py:47-84) on the tiny dataset.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_dgcf_sampler.py
Writes tests/golden/dgcf_sampler.npz: the (data, cor) batches of one epoch after random.seed(5); np.random.seed(5),
train_batch 16, cor_batch 10, with and without tags.
"""
import collections
import collections.abc
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from helpers import nums, user_lists  # noqa: E402

stub = types.ModuleType("tensorboardX")
stub.SummaryWriter = object
sys.modules["tensorboardX"] = stub
collections.Iterable = collections.abc.Iterable
sys.argv = ["golden", "--model", "dgcf"]
sys.path.insert(0, "/root/reference")
os.chdir("/tmp")
from utility.word import CFG  # noqa: E402
from train_data.bpr_training_data import DGCF_training_data  # noqa: E402

g = dict(np.load(os.path.join(HERE, "tiny.npz")))
U, I, Tg, _ = nums(g)


class D:
    pass


d = D()
d.num = {"user": U, "item": I, "tag": Tg}
d.user_items = {"train": user_lists(g, "train")}
d.edge_index = {"train": g["edge_index_train"]}
out = {}
for use_tag in (False, True):
    CFG.update(train_batch=16, use_tag=use_tag, cor_batch=10, device=torch.device("cpu"))
    random.seed(5)
    np.random.seed(5)
    s = DGCF_training_data(d, None)
    batches = list(s.mini_batch())
    tag = "tag" if use_tag else "notag"
    out[f"{tag}_data"] = np.stack([b[0].numpy() for b in batches])
    out[f"{tag}_cor"] = np.stack([b[1].numpy() for b in batches])
    # small-population branch: fewer users than the batch -> np.random.choice (train_data/utils.py:62-63)
    CFG.update(train_batch=64)
    random.seed(6)
    np.random.seed(6)
    s = DGCF_training_data(d, None)
    out[f"{tag}_small_data"] = np.stack([b[0].numpy() for b in s.mini_batch()])
np.savez_compressed(os.path.join(HERE, "dgcf_sampler.npz"), **out)
print({k: v.shape for k, v in out.items()})
