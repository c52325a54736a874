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


class DGCF_training_data(Abstract_training_data):
    """Drop-in for train_data/bpr_training_data.py:47-84 (the NGCF-style sampler DGCF / DisenGCN / DisenHAN use,
    com.py:35,46,57): every batch = ``train_batch`` sampled users with one positive and one negative item each, plus
    the (unused) ``cor`` index sample; ``tot_inter = E // B + 1`` batches per epoch; ``reset()`` is a no-op;
    ``mini_batch()`` yields ``(data, cor)`` tuples (models unpack them, dgcf.py:116).

    CFG['sampler'] == 'mt19937' restates the reference's host procedure call for call — ``random.sample`` for the
    users and the ``cor`` indices, ``np.random.choice`` for the positive, ``np.random.randint`` rejection for the
    negative (train_data/utils.py:58-78) — so seeding ``random`` and ``np.random`` reproduces its stream bit for
    bit.  'device' (default) draws users/positives with torch's CUDA generator and the negatives with the Philox
    rejection kernel (``tagrec_sample_bpr_device``)."""

    def __init__(self, data, args=None):
        super().__init__(args)
        cfg = config.current()
        self.batch_size = cfg['train_batch']
        self.cor_batch = cfg.get('cor_batch', 100)
        self.use_tag = cfg['use_tag']
        self.mode = cfg.get('sampler', 'device')
        self.seed = int(cfg.get('seed', 2020))
        self.num_item = int(data.num['item'])
        self.num_user = int(data.num['user'])
        self.num_tag = int(data.num.get('tag', 0)) if isinstance(data.num, dict) else int(data.num['tag'])
        self.train_ui = data.user_items['train']
        self.pos_inter = data.edge_index['train']
        self.tot_inter = self.pos_inter.shape[0] // self.batch_size + 1
        self.calls = 0
        if self.mode == "device":
            dev = self.device
            ptr_h, items_h = user_items_to_csr(self.train_ui, self.num_user)
            self._ptr_d = torch.as_tensor(ptr_h, device=dev)
            self._items_d = torch.as_tensor(items_h, device=dev).to(torch.int32)
            self._users_d = torch.as_tensor(np.fromiter(self.train_ui.keys(), dtype=np.int64), device=dev)
            self._gen = torch.Generator(device=dev)
            self._gen.manual_seed(self.seed)
        start = time.time()
        self.mini_sample()
        print(f"DGCF_training_data producer,tot_inter:{self.tot_inter},cor_batch:{self.cor_batch},"
              f"[mini_sample time:{time.time()-start}]")

    def _host_sample(self):
        import random
        all_user = list(self.train_ui.keys())
        if len(all_user) > self.batch_size:
            sample_user = random.sample(all_user, self.batch_size)
        else:
            sample_user = np.random.choice(all_user, self.batch_size)
        rows = []
        for u in sample_user:
            pos_list = self.train_ui[u]
            pos_i = np.random.choice(pos_list)
            while True:
                neg_i = np.random.randint(0, self.num_item)
                if neg_i not in pos_list:
                    break
            rows.append([u, pos_i, neg_i])
        data = torch.tensor(np.array(rows), dtype=torch.long, device=self.device)
        cor = [random.sample(list(range(self.num_user)), self.cor_batch),
               random.sample(list(range(self.num_item)), self.cor_batch)]
        if self.use_tag:
            cor.append(random.sample(list(range(self.num_tag)), self.cor_batch))
        return data, torch.tensor(np.stack(cor), dtype=torch.long, device=self.device)

    def _device_sample(self):
        dev, b, g = self.device, self.batch_size, self._gen
        nu = self._users_d.numel()
        if nu > b:
            users = self._users_d[torch.randperm(nu, device=dev, generator=g)[:b]]
        else:
            users = self._users_d[torch.randint(0, nu, (b,), device=dev, generator=g)]
        deg = self._ptr_d[users + 1] - self._ptr_d[users]
        off = torch.minimum((torch.rand(b, device=dev, generator=g, dtype=torch.float64) * deg).long(), deg - 1)
        pos = self._items_d[self._ptr_d[users] + off].long()
        edges = torch.stack([users, pos], 1).contiguous()
        out = torch.empty((b, 3), dtype=torch.int64, device=dev)
        check(lib().tagrec_sample_bpr_device(ptr(edges), b, ptr(self._ptr_d), ptr(self._items_d), self.num_item,
                                             self.seed, self.calls, ptr(out), stream_ptr(dev)),
              "tagrec_sample_bpr_device")
        sizes = [self.num_user, self.num_item] + ([self.num_tag] if self.use_tag else [])
        # (the reference's random.sample raises when a population is smaller than cor_batch; here: with replacement)
        cb = self.cor_batch
        cor = torch.stack([torch.randperm(n, device=dev, generator=g)[:cb] if n >= cb
                           else torch.randint(0, n, (cb,), device=dev, generator=g) for n in sizes])
        return out, cor

    def mini_sample(self):
        self.calls += 1
        return self._host_sample() if self.mode != "device" else self._device_sample()

    def reset(self):
        pass

    def mini_batch(self):
        for _ in range(0, self.tot_inter):
            yield self.mini_sample()


class TransTag_training_data(Abstract_training_data):
    """Drop-in for train_data/transe_training_data.py:42-70 — TGCN's second training phase (com.py:68-70): for every
    (user, tag, item) triple of ``data.uit_data`` one negative item that the user never tagged with that tag; no
    shuffle; re-sampled by ``reset()`` every epoch; batches of ``transtag_batch`` rows ``[u, t, i+, i-]``.

    CFG['sampler'] == 'mt19937': numpy-legacy stream restated in C++ (``tagrec_sample_neg_tail_host``), bit-exact with
    the reference at ``cpu_core == 1`` — a forked worker samples from a COPY of the global generator, so (as in the
    reference) the parent's state does not move and every epoch draws the same negatives unless something else
    advanced it.  'device': Philox rejection kernel (rows come out in a pseudo-random order)."""

    def __init__(self, data, args=None):
        super().__init__(args)
        cfg = config.current()
        self.args = args
        self.batch_size = cfg['transtag_batch']
        self.mode = cfg.get('sampler', 'device')
        self.seed = int(cfg.get('seed', 2020))
        self.num = int(data.num['item'])
        self.uti_data = np.ascontiguousarray(np.asarray(data.uit_data)[:, [0, 2, 1]], dtype=np.int64)
        # (u, t) groups -> ascending tails: the reference's u_t_dict (train_data/utils.py:40-46) as a CSR
        n_tag = int(self.uti_data[:, 1].max()) + 1 if len(self.uti_data) else 1
        key = self.uti_data[:, 0] * n_tag + self.uti_data[:, 1]
        uniq, self._group = np.unique(key, return_inverse=True)
        order = np.lexsort((self.uti_data[:, 2], self._group))
        self._gitems = np.ascontiguousarray(self.uti_data[order, 2])
        self._gptr = np.zeros(len(uniq) + 1, dtype=np.int64)
        np.cumsum(np.bincount(self._group, minlength=len(uniq)), out=self._gptr[1:])
        self._group = np.ascontiguousarray(self._group, dtype=np.int64)
        self.epoch = 0
        if self.mode == "device":
            dev = self.device
            e = len(self.uti_data)
            self._edges_d = torch.as_tensor(np.stack([self._group, np.arange(e)], 1), device=dev)
            self._gptr_d = torch.as_tensor(self._gptr, device=dev)
            self._gitems_d = torch.as_tensor(self._gitems, device=dev).to(torch.int32)
            self._uti_d = torch.as_tensor(self.uti_data, device=dev)
        start = time.time()
        self.all_train_data = self.get_all_training_data()
        self.tot_inter = self.all_train_data.shape[0] // self.batch_size
        print(f"TransTag_training_data producer, tot_inter: {self.tot_inter},"
              f"[all_training_data time:{time.time()-start}]")

    def get_all_training_data(self):
        e = len(self.uti_data)
        if self.mode != "device":
            kind, keyw, pos, has_gauss, cached = np.random.get_state()
            state = np.empty(625, dtype=np.uint32)
            state[:624], state[624] = keyw, pos
            neg = np.empty(e, dtype=np.int64)
            check(lib().tagrec_sample_neg_tail_host(ptr(state), ptr(self._group), e, ptr(self._gptr), ptr(self._gitems),
                                                    self.num, ptr(neg)), "tagrec_sample_neg_tail_host")
            out = np.concatenate([self.uti_data, neg[:, None]], axis=1)
            return torch.as_tensor(out, dtype=torch.long, device=self.device)
        tri = torch.empty((e, 3), dtype=torch.int64, device=self.device)
        check(lib().tagrec_sample_bpr_device(ptr(self._edges_d), e, ptr(self._gptr_d), ptr(self._gitems_d), self.num,
                                             self.seed + 1, self.epoch, ptr(tri), stream_ptr(tri.device)),
              "tagrec_sample_bpr_device")
        self.epoch += 1
        return torch.cat([self._uti_d[tri[:, 1]], tri[:, 2:3]], dim=1)
