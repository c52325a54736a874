// BPR negative sampler.  Replaces train_data/bpr_training_data.py:29-45 + train_data/utils.py:19-28,52-55
// (a Python loop per edge inside forked workers, list-scan membership, np.vstack, np.random.shuffle, H2D).
//
// Two modes:
//  * tagrec_sample_bpr_host   — parity mode, HOST code: the numpy-legacy MT19937 stream restated in C++
//    (init_genrand seeding, one 32-bit output per masked-rejection attempt of randint, descending Fisher-Yates
//    shuffle with the same bounded draw).  Bit-exact with the reference for cpu_core == 1: the
//    negatives are drawn from a COPY of the generator (the forked worker), the caller's state is advanced only by
//    the shuffle.
//  * tagrec_sample_bpr_device — throughput mode, sm_100a kernel: Philox4x32-10 counter RNG keyed by
//    (seed, epoch, edge, attempt), the same masked rejection for an unbiased item, membership by binary search in
//    the user's ascending train row, and the epoch shuffle as an on-the-fly Feistel permutation of the edge index
//    (no sort, no second pass).  Statistically equivalent to the reference, not the same stream.
#include <vector>

#include "common.cuh"

namespace tagrec {

// ------------------------------------------------------------------------------------------------ MT19937 (host)
struct Mt {
    uint32_t* mt;   // 624 words
    uint32_t* pos;  // index of the next word, 624 == needs a twist
    void twist() {
        for (int k = 0; k < 624; ++k) {
            const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        *pos = 0;
    }
    uint32_t next() {
        if (*pos >= 624) twist();
        uint32_t y = mt[(*pos)++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    // numpy random_interval / buffered_bounded_masked_uint32: uniform in [0, mx], mx < 2^32
    uint32_t bounded(uint32_t mx) {
        if (mx == 0) return 0;
        uint32_t mask = mx;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        while ((v = next() & mask) > mx) {}
        return v;
    }
};

// ------------------------------------------------------------------------------------------------ Philox (device)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

__device__ __forceinline__ uint32_t mix32(uint32_t x, uint32_t k) {
    x ^= k;
    x *= 0x9E3779B1u; x ^= x >> 15;
    x *= 0x85EBCA77u; x ^= x >> 13;
    x *= 0xC2B2AE3Du; x ^= x >> 16;
    return x;
}

// Bijection on [0, n): 4-round Feistel network on 2*h bits with cycle walking.
__device__ __forceinline__ uint64_t feistel_perm(uint64_t i, uint64_t n, int h, uint64_t key) {
    const uint64_t half_mask = (1ull << h) - 1ull;
    do {
        uint64_t l = i >> h, r = i & half_mask;
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            const uint64_t f = mix32((uint32_t)r ^ (uint32_t)(r >> 32), (uint32_t)(key >> (8 * round)) + 0x632BE5ABu * round);
            const uint64_t nl = r, nr = (l ^ f) & half_mask;
            l = nl; r = nr;
        }
        i = (l << h) | r;
    } while (i >= n);
    return i;
}

__global__ void __launch_bounds__(256)
sample_bpr_kernel(const int64_t* __restrict__ edges, int64_t e, const int64_t* __restrict__ train_ptr,
                  const int32_t* __restrict__ train_items, uint32_t num_item, uint64_t seed, uint64_t epoch, int h,
                  int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e) return;
    const uint64_t src = feistel_perm((uint64_t)i, (uint64_t)e, h, seed * 0x9E3779B97F4A7C15ull + epoch);
    const int64_t u = edges[2 * src], pos = edges[2 * src + 1];
    const int64_t lo0 = train_ptr[u], hi0 = train_ptr[u + 1];
    const uint32_t mx = num_item - 1;
    uint32_t mask = mx;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    int64_t neg = -1;
    for (uint32_t blk = 0; neg < 0; ++blk) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)src, (uint32_t)(src >> 32), (uint32_t)epoch, blk), key);
        const uint32_t draws[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const uint32_t cand = draws[d] & mask;
            if (neg >= 0 || cand > mx) continue;
            int64_t lo = lo0, hi = hi0;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if ((uint32_t)__ldg(train_items + mid) < cand) lo = mid + 1; else hi = mid;
            }
            if (!(lo < hi0 && (uint32_t)__ldg(train_items + lo) == cand)) neg = cand;
        }
        if (blk > (1u << 20)) neg = 0;   // a user that interacted with every item: cannot happen in valid data
    }
    out[3 * i] = u;
    out[3 * i + 1] = pos;
    out[3 * i + 2] = neg;
}

}  // namespace tagrec

using namespace tagrec;

extern "C" void tagrec_mt19937_seed(uint32_t seed, uint32_t* state) {
    // init_genrand — what np.random.seed(int) does
    state[0] = seed;
    for (uint32_t i = 1; i < 624; ++i) state[i] = 1812433253u * (state[i - 1] ^ (state[i - 1] >> 30)) + i;
    state[624] = 624;
}

extern "C" int tagrec_sample_bpr_host(uint32_t* state, const int64_t* edges, int64_t e, const int64_t* train_ptr,
                                      const int64_t* train_items_sorted, int64_t num_item, int64_t* triples_out) {
    TAGREC_REQUIRE(state && edges && train_ptr && train_items_sorted && triples_out, "null pointer");
    TAGREC_REQUIRE(num_item > 0 && num_item <= 0xffffffffll, "num_item out of range");
    // the forked worker: a copy of the generator (bpr_training_data.py:37-39 with cpu_core == 1)
    std::vector<uint32_t> wstate(state, state + 625);
    Mt worker{wstate.data(), wstate.data() + 624};
    std::vector<int64_t> tmp((size_t)e * 3);
    for (int64_t k = 0; k < e; ++k) {
        const int64_t u = edges[2 * k];
        const int64_t* lo = train_items_sorted + train_ptr[u];
        const int64_t* hi = train_items_sorted + train_ptr[u + 1];
        int64_t neg;
        for (;;) {                                               // train_data/utils.py:22-26
            neg = (int64_t)worker.bounded((uint32_t)(num_item - 1));
            const int64_t *a = lo, *b = hi;
            while (a < b) {
                const int64_t* mid = a + (b - a) / 2;
                if (*mid < neg) a = mid + 1; else b = mid;
            }
            if (!(a < hi && *a == neg)) break;
        }
        tmp[3 * k] = u;
        tmp[3 * k + 1] = edges[2 * k + 1];
        tmp[3 * k + 2] = neg;
    }
    // the parent: np.random.shuffle(arange(e)) then data[index] (train_data/utils.py:52-55)
    Mt parent{state, state + 624};
    std::vector<int64_t> idx((size_t)e);
    for (int64_t k = 0; k < e; ++k) idx[k] = k;
    for (int64_t i = e - 1; i >= 1; --i) {
        const int64_t j = (int64_t)parent.bounded((uint32_t)i);
        std::swap(idx[i], idx[j]);
    }
    for (int64_t k = 0; k < e; ++k) {
        triples_out[3 * k] = tmp[3 * idx[k]];
        triples_out[3 * k + 1] = tmp[3 * idx[k] + 1];
        triples_out[3 * k + 2] = tmp[3 * idx[k] + 2];
    }
    return TAGREC_OK;
}

extern "C" int tagrec_sample_neg_tail_host(const uint32_t* state, const int64_t* group, int64_t e,
                                           const int64_t* group_ptr, const int64_t* group_items_sorted, int64_t num,
                                           int64_t* neg_out) {
    TAGREC_REQUIRE(state && group && group_ptr && group_items_sorted && neg_out, "null pointer");
    TAGREC_REQUIRE(num > 0 && num <= 0xffffffffll, "num out of range");
    // the forked worker of transe_training_data.py:61-66 with cpu_core == 1: a COPY of the generator, the parent's
    // state is not advanced (there is no shuffle in this sampler)
    std::vector<uint32_t> wstate(state, state + 625);
    Mt worker{wstate.data(), wstate.data() + 624};
    for (int64_t k = 0; k < e; ++k) {
        const int64_t* lo = group_items_sorted + group_ptr[group[k]];
        const int64_t* hi = group_items_sorted + group_ptr[group[k] + 1];
        int64_t neg;
        for (;;) {                                               // train_data/utils.py:31-37 sample_neg_tail
            neg = (int64_t)worker.bounded((uint32_t)(num - 1));
            const int64_t *a = lo, *b = hi;
            while (a < b) {
                const int64_t* mid = a + (b - a) / 2;
                if (*mid < neg) a = mid + 1; else b = mid;
            }
            if (!(a < hi && *a == neg)) break;
        }
        neg_out[k] = neg;
    }
    return TAGREC_OK;
}

extern "C" int tagrec_sample_bpr_device(const int64_t* edges, int64_t e, const int64_t* train_ptr,
                                        const int32_t* train_items, int64_t num_item, uint64_t seed, uint64_t epoch,
                                        int64_t* triples_out, void* stream) {
    TAGREC_REQUIRE(edges && train_ptr && train_items && triples_out, "null pointer");
    TAGREC_REQUIRE(num_item > 0 && num_item <= 0x7fffffffll, "num_item out of range");
    if (e == 0) return TAGREC_OK;
    int bits = 1;
    while ((1ll << bits) < e) ++bits;
    const int h = (bits + 1) / 2;
    TAGREC_LAUNCH(sample_bpr_kernel, (unsigned)((e + 255) / 256), 256, 0, stream, edges, e, train_ptr, train_items,
                  (uint32_t)num_item, seed, epoch, h, triples_out);
    return TAGREC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// TGCN neighbour tables on the device.  Replaces data/utils.py:87-106 (all_neighbor_sample) + data/tgcn_load.py:41-53:
// a Python loop over every row with ``matrix[i].toarray()`` and np.random.choice.  One thread per row of one relation
// (row type a -> column type b = the stored entries of the row whose column id lies in [col_lo, col_hi) of the
// tripartite CSR): `width` neighbour ids (+1; 0 = padding, only in empty rows) drawn WITH replacement when the row has
// fewer than `width` entries, a uniformly random `width`-subset in random order otherwise (Floyd's algorithm + a
// Fisher-Yates pass), and the matching integer edge weights.  Philox4x32-10 keyed by (seed, relation, row):
// statistically the reference's tables, not numpy's stream (the host builder in data.py keeps the stream-identical form).
namespace tagrec {
constexpr int kMaxNbrWidth = 64;

__device__ __forceinline__ uint32_t bounded_u32(uint32_t r, uint32_t n) {       // uniform in [0, n), n < 2^31
    return (uint32_t)(((uint64_t)r * (uint64_t)n) >> 32);
}

__global__ void __launch_bounds__(128)
neighbor_table_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ weight,
                      int64_t row_begin, int64_t n_rows, int32_t col_lo, int32_t col_hi, int width, uint64_t seed,
                      uint32_t relation, int64_t* __restrict__ ids, int64_t* __restrict__ wts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const int64_t r = row_begin + i;
    int64_t lo = __ldg(rowptr + r), hi = __ldg(rowptr + r + 1);
    {   // the row's entries with column in [col_lo, col_hi): two lower bounds over the ascending columns
        int64_t a = lo, b = hi;
        while (a < b) { const int64_t m = (a + b) >> 1; if (__ldg(col + m) < col_lo) a = m + 1; else b = m; }
        const int64_t s = a;
        b = hi;
        while (a < b) { const int64_t m = (a + b) >> 1; if (__ldg(col + m) < col_hi) a = m + 1; else b = m; }
        lo = s; hi = a;
    }
    const int64_t x = hi - lo;
    int64_t* out_i = ids + i * width;
    int64_t* out_w = wts + i * width;
    if (x == 0) {
        for (int s = 0; s < width; ++s) { out_i[s] = 0; out_w[s] = 0; }
        return;
    }
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    uint32_t blk = 0;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    int have = 0;
    auto next = [&]() -> uint32_t {
        if (have == 0) {
            rnd = philox4x32_10(make_uint4((uint32_t)r, (uint32_t)(r >> 32), relation, blk++), key);
            have = 4;
        }
        const uint32_t v = have == 4 ? rnd.x : have == 3 ? rnd.y : have == 2 ? rnd.z : rnd.w;
        --have;
        return v;
    };
    int32_t pick[kMaxNbrWidth];
    if (x < width) {                                    // with replacement (np.random.choice(ids, max_deg))
        for (int s = 0; s < width; ++s) pick[s] = (int32_t)bounded_u32(next(), (uint32_t)x);
    } else {                                            // without replacement: Floyd, then shuffle the order
        int cnt = 0;
        for (int64_t j = x - width; j < x; ++j) {
            int32_t t = (int32_t)bounded_u32(next(), (uint32_t)(j + 1));
            bool seen = false;
            for (int q = 0; q < cnt; ++q) seen |= (pick[q] == t);
            pick[cnt++] = seen ? (int32_t)j : t;
        }
        for (int s = width - 1; s > 0; --s) {
            const int t = (int)bounded_u32(next(), (uint32_t)(s + 1));
            const int32_t tmp = pick[s]; pick[s] = pick[t]; pick[t] = tmp;
        }
    }
    for (int s = 0; s < width; ++s) {
        const int64_t e = lo + pick[s];
        out_i[s] = (int64_t)(__ldg(col + e) - col_lo) + 1;
        out_w[s] = (int64_t)llrintf(__ldg(weight + e));
    }
}
}  // namespace tagrec

extern "C" int tagrec_neighbor_table(const int64_t* rowptr, const int32_t* col, const float* weight, int64_t row_begin,
                                     int64_t n_rows, int32_t col_lo, int32_t col_hi, int width, uint64_t seed,
                                     uint32_t relation, int64_t* ids, int64_t* wts, void* stream) {
    TAGREC_REQUIRE(rowptr && col && weight && ids && wts, "null pointer");
    TAGREC_REQUIRE(width >= 1 && width <= tagrec::kMaxNbrWidth, "table width must be in 1..64");
    TAGREC_REQUIRE(col_hi >= col_lo && row_begin >= 0 && n_rows >= 0, "bad row / column range");
    if (n_rows == 0) return TAGREC_OK;
    TAGREC_LAUNCH(tagrec::neighbor_table_kernel, (unsigned)((n_rows + 127) / 128), 128, 0, stream, rowptr, col, weight,
                  row_begin, n_rows, col_lo, col_hi, width, seed, relation, ids, wts);
    return TAGREC_OK;
}
