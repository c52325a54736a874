"""BPR sampler — drop-in for train_data/bpr_training_data.py:12-45 + train_data/abstract.py.

``BPR_training_data(data, args)`` / ``.reset()`` / ``.mini_batch()`` / ``.all_train_data`` / ``.tot_inter`` keep the
reference's meaning.  Sampling itself runs in libtagrec_b200.so:
  * CFG['sampler'] == 'device'  (default): one Philox kernel over all positive edges, on the training device;
  * CFG['sampler'] == 'mt19937': the numpy-legacy stream restated in C++ — bit-exact with the reference run with
    ``--cpu_core 1``; it reads and advances numpy's GLOBAL RandomState exactly as the reference does (the parent's
    generator moves only by the shuffle).
"""
import time

import numpy as np
import torch

from . import config
from ._lib import check, lib, ptr, stream_ptr


class Abstract_training_data:
    """train_data/abstract.py:4-23."""

    def __init__(self, args=None):
        cfg = config.current()
        self.device = cfg['device']
        self.cpu_core = cfg["cpu_core"]
        self.all_train_data = None

    def get_all_training_data(self):
        raise NotImplementedError

    def reset(self):
        self.all_train_data = self.get_all_training_data()

    def mini_batch(self):
        # abstract.py:17-23: the tail shorter than a batch is merged into the previous one AND yielded again
        # (the loop keeps going) — reproduced on purpose
        n = self.all_train_data.shape[0]
        for i in range(0, n, self.batch_size):
            if i + 2 * self.batch_size > n:
                yield self.all_train_data[i:]
            else:
                yield self.all_train_data[i:i + self.batch_size]


def user_items_to_csr(user_items, n_user, sort=True):
    """dict u -> list of items  ->  (ptr int64 [n_user+1], items int64) with ascending items per user."""
    ptr_ = np.zeros(n_user + 1, dtype=np.int64)
    for u, its in user_items.items():
        ptr_[u + 1] = len(its)
    ptr_ = np.cumsum(ptr_)
    flat = np.empty(int(ptr_[-1]), dtype=np.int64)
    for u, its in user_items.items():
        a = np.asarray(its, dtype=np.int64)
        flat[ptr_[u]:ptr_[u + 1]] = np.sort(a) if sort else a
    return ptr_, flat


class BPR_training_data(Abstract_training_data):
    def __init__(self, data, args=None):
        super().__init__(args)
        cfg = config.current()
        self.batch_size = cfg['train_batch']
        self.mode = cfg.get('sampler', 'device')
        self.seed = int(cfg.get('seed', 2020))
        self.num = int(data.num['item'])
        self.num_user = int(data.num['user'])
        self.train_ui = data.user_items['train']
        self.pos_inter = np.ascontiguousarray(data.edge_index['train'], dtype=np.int64)
        self.args = args
        self.epoch = 0
        csr = getattr(data, "train_csr", None)
        if csr is None:
            csr = user_items_to_csr(self.train_ui, self.num_user)
        self._ptr_h, self._items_h = csr
        if self.mode == "device":
            dev = self.device
            self._edges_d = torch.as_tensor(self.pos_inter, device=dev)
            self._ptr_d = torch.as_tensor(self._ptr_h, device=dev)
            self._items_d = torch.as_tensor(self._items_h, device=dev).to(torch.int32)
        start = time.time()
        self.all_train_data = self.get_all_training_data()       # bpr_training_data.py:23 (discarded by reset())
        self.tot_inter = self.all_train_data.shape[0] // self.batch_size
        print(f"BPR_training_data producer tot_inter: {self.tot_inter},"
              f"[all_training_data spend time:{time.time()-start}]")

    def get_all_training_data(self):
        e = self.pos_inter.shape[0]
        if self.mode == "mt19937":
            kind, key, pos, has_gauss, cached = np.random.get_state()
            state = np.empty(625, dtype=np.uint32)
            state[:624], state[624] = key, pos
            out = np.empty((e, 3), dtype=np.int64)
            check(lib().tagrec_sample_bpr_host(ptr(state), ptr(self.pos_inter), e, ptr(self._ptr_h), ptr(self._items_h),
                                               self.num, ptr(out)), "tagrec_sample_bpr_host")
            np.random.set_state((kind, state[:624].copy(), int(state[624]), has_gauss, cached))
            return torch.as_tensor(out, dtype=torch.long, device=self.device)
        out = torch.empty((e, 3), dtype=torch.int64, device=self.device)
        check(lib().tagrec_sample_bpr_device(ptr(self._edges_d), e, ptr(self._ptr_d), ptr(self._items_d), self.num,
                                             self.seed, self.epoch, ptr(out), stream_ptr(out.device)),
              "tagrec_sample_bpr_device")
        self.epoch += 1
        return out
