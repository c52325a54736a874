"""Golden stream of the reference's TransTag_training_data (train_data/transe_training_data.py:42-70) on the tiny dataset.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_transtag.py
Writes tests/golden/transtag_sampler.npz: all_train_data of two consecutive epochs after np.random.seed(2020),
cpu_core 1, pool None (a fresh forked worker per call: both epochs draw the same negatives).
"""
import collections
import collections.abc
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from helpers import nums  # noqa: E402

stub = types.ModuleType("tensorboardX")
stub.SummaryWriter = object
sys.modules["tensorboardX"] = stub
collections.Iterable = collections.abc.Iterable
sys.argv = ["golden", "--model", "tgcn"]
sys.path.insert(0, "/root/reference")
os.chdir("/tmp")
from utility.word import CFG  # noqa: E402
from train_data.transe_training_data import TransTag_training_data  # noqa: E402

g = dict(np.load(os.path.join(HERE, "tiny.npz")))
U, I, Tg, _ = nums(g)


class D:
    pass


class A:
    pool = None


d = D()
d.num = {"user": U, "item": I, "tag": Tg}
d.uit_data = g["uit_data"]
CFG.update(transtag_batch=64, cpu_core=1, device=torch.device("cpu"))
np.random.seed(2020)
r = TransTag_training_data(d, A())
first = r.all_train_data.numpy().copy()
r.reset()
second = r.all_train_data.numpy().copy()
np.savez_compressed(os.path.join(HERE, "transtag_sampler.npz"), first=first, second=second)
print(first.shape, np.array_equal(first, second))
