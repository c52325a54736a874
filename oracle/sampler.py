"""Oracle: BPR negative sampler and batch iterator, bit-exact restatement of the numpy-legacy streams.

Restates train_data/bpr_training_data.py:12-45, train_data/utils.py:5-28,52-55 and
train_data/abstract.py:14-23 for ``cpu_core == 1`` (the only reproducible setting).

The arithmetic lives in numpy's legacy ``RandomState`` (third-party, unpinned by the reference; pinned here to the
image's numpy 2.3.5): MT19937 seeded by ``init_genrand``; ``randint(0, n)`` = masked rejection on ONE 32-bit
output per attempt (``random_bounded_uint64_fill`` -> ``buffered_bounded_masked_uint32``); ``shuffle`` = descending
Fisher-Yates with ``random_interval`` (same masked rejection).  Pure python: small cases only.
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import numpy as np


class MT19937:
    """Mersenne Twister 32-bit stream identical to ``np.random.seed(int)`` / ``next_uint32``."""
    N, M = 624, 397

    def __init__(self, seed):
        mt = [0] * self.N
        mt[0] = seed & 0xFFFFFFFF
        for i in range(1, self.N):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.mt, self.pos = mt, self.N

    def clone(self):
        c = object.__new__(MT19937)
        c.mt, c.pos = list(self.mt), self.pos
        return c

    def _twist(self):
        mt, N, M = self.mt, self.N, self.M
        for k in range(N):
            y = (mt[k] & 0x80000000) | (mt[(k + 1) % N] & 0x7FFFFFFF)
            mt[k] = mt[(k + M) % N] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        self.pos = 0

    def next_uint32(self):
        if self.pos == self.N:
            self._twist()
        y = self.mt[self.pos]
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def _mask(mx):
    m = mx
    for s in (1, 2, 4, 8, 16, 32):
        m |= m >> s
    return m


def bounded(rng, mx):
    """numpy random_interval / buffered_bounded_masked_uint32: uniform in [0, mx], mx < 2^32."""
    if mx == 0:
        return 0
    m = _mask(mx)
    while True:
        v = rng.next_uint32() & m
        if v <= mx:
            return v


def randint(rng, low, high):
    """np.random.randint(low, high) scalar draw (train_data/utils.py:23)."""
    return low + bounded(rng, high - 1 - low)


def shuffle_index(rng, n):
    """np.random.shuffle(np.arange(n)) (train_data/utils.py:52-55)."""
    idx = np.arange(n)
    for i in range(n - 1, 0, -1):
        j = bounded(rng, i)
        idx[i], idx[j] = idx[j], idx[i]
    return idx


def sample_epoch(rng, edge_index, train_ui, num_item):
    """One ``get_all_training_data()`` with cpu_core=1 (bpr_training_data.py:29-45).

    The single forked worker starts from a COPY of the parent's generator state and the parent's state is not
    advanced by it; the parent then draws the shuffle from that same state.  ``rng`` is advanced in
    place exactly as the parent's generator is (by the shuffle only)."""
    worker = rng.clone()
    out = np.empty((len(edge_index), 3), dtype=np.int64)
    for k, (u, pos) in enumerate(edge_index):
        pos_set = train_ui[int(u)]
        while True:                                                # train_data/utils.py:22-26
            j = randint(worker, 0, num_item)
            if j not in pos_set:
                out[k] = (u, pos, j)
                break
    return out[shuffle_index(rng, len(out))]


def mini_batches(n_rows, batch_size):
    """abstract.py:17-23 — (start, stop) slices; when i + 2B > E the slice is [i:] AND the loop goes on, so the
    last E mod B rows are yielded twice."""
    out = []
    for i in range(0, n_rows, batch_size):
        if i + 2 * batch_size > n_rows:
            out.append((i, n_rows))
        else:
            out.append((i, i + batch_size))
    return out


# ---------------------------------------------------------------------------------------------------------------
# bench.py's CPU arm: the reference's reset() AS IT EXECUTES (numpy's own generator, Python loop per edge, a pool of
# forked workers over equal chunks, vstack, shuffle) — train_data/bpr_training_data.py:29-45 + train_data/utils.py:5-28,52-55.
def _sample_neg_chunk(args):
    pos_inter, data_dict, num_item = args
    import numpy as np
    data = []
    for u, pos_i in pos_inter:                                    # train_data/utils.py:19-26
        while True:
            idx = np.random.randint(0, num_item)
            if idx not in data_dict[u]:
                data.append([u, pos_i, idx])
                break
    return np.array(data)


def reference_reset(pos_inter, train_ui, num_item, cpu_core):
    """One ``BPR_training_data.reset()`` the reference's way; returns the (E, 3) int64 array.  ``cpu_core`` forked
    workers (multiprocessing.Pool, like bpr_training_data.py:37-39) or an in-process loop when cpu_core == 1."""
    import multiprocessing

    import numpy as np
    size = len(pos_inter) // cpu_core                             # train_data/utils.py:5-16 split_data
    chunks = [pos_inter[i * size:(len(pos_inter) if i == cpu_core - 1 else (i + 1) * size)] for i in range(cpu_core)]
    jobs = [(c, train_ui, num_item) for c in chunks]
    if cpu_core == 1:
        results = [_sample_neg_chunk(jobs[0])]
    else:
        with multiprocessing.get_context("fork").Pool(cpu_core) as pool:
            results = pool.map(_sample_neg_chunk, jobs)
    data = np.vstack(results)
    idx = np.arange(len(data))                                    # train_data/utils.py:52-55 shuffle
    np.random.shuffle(idx)
    return data[idx]
