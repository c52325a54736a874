// K0 — block adjacency -> CSR on device, bit-exact with the reference's scipy path.  sm_100a.
//
// Replaces model/help/adj.py:7-35 (lil_matrix block assembly: minutes for 1 M edges, impossible at 1 B),
// adj.py:90-110 (value computation of bi_norm / si_norm) and adj.py:144-150 (sp2tensor).
//
// Pipeline (all device, one stream):
//   1. emit one 64-bit key (row << 32 | col) per directed entry of the block matrix (both orientations of every
//      block, + the diagonal when self loops are requested)                                   [emit_keys_kernel]
//   2. LSD radix sort of the keys on the col bits then the row bits (cub::DeviceRadixSort — a CUDA-toolkit
//      primitive; the sort is a one-off set-up step, not part of the per-step hot loop)
//   3. run-length encode equal keys: multiplicity = the integer edge weight the reference gets from summing
//      duplicates in COO->LIL (data/utils.py:50-53)                                            [cub RLE]
//   4. decode keys to col / weight, derive rowptr from row changes, weighted float32 degree    [decode_kernel,
//                                                                                               degree_kernel]
//   5. (host, numpy) dpow = np.power(degree, -0.5 | -1)      — adj.py:93,105, see tagrec_b200.h
//   6. val = (dpow[row]*w)*dpow[col] with two separately rounded multiplies == scipy's D*A*D    [normalise_kernel]
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>

#include "common.cuh"

namespace tagrec {

struct Blocks {
    const int64_t *ui_row, *ui_col, *ut_row, *ut_col, *it_row, *it_col;
    int64_t e_ui, e_ut, e_it, n_user, n_item, n_tag, n_diag;
};

__device__ __forceinline__ uint64_t make_key(int64_t r, int64_t c) { return ((uint64_t)r << 32) | (uint64_t)c; }

__global__ void emit_keys_kernel(Blocks b, uint64_t* __restrict__ keys) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t e_all = b.e_ui + b.e_ut + b.e_it;
    const int64_t off_i = b.n_user, off_t = b.n_user + b.n_item;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e_all + b.n_diag; i += stride) {
        if (i >= e_all) {  // diagonal
            const int64_t d = i - e_all;
            keys[2 * e_all + d] = make_key(d, d);
            continue;
        }
        int64_t r, c;
        if (i < b.e_ui) {
            r = b.ui_row[i];
            c = b.ui_col[i] + off_i;
        } else if (i < b.e_ui + b.e_ut) {
            r = b.ut_row[i - b.e_ui];
            c = b.ut_col[i - b.e_ui] + off_t;
        } else {
            r = b.it_row[i - b.e_ui - b.e_ut] + off_i;
            c = b.it_col[i - b.e_ui - b.e_ut] + off_t;
        }
        keys[2 * i] = make_key(r, c);
        keys[2 * i + 1] = make_key(c, r);
    }
}

// One thread per unique entry: col, weight; the first entry of every row (and the gap of empty rows before it)
// writes rowptr.
__global__ void decode_kernel(const uint64_t* __restrict__ ukeys, const int32_t* __restrict__ counts,
                              const int32_t* __restrict__ n_runs, int64_t n, int64_t* __restrict__ rowptr,
                              int32_t* __restrict__ col, float* __restrict__ weight) {
    const int64_t nnz = *n_runs;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += stride) {
        const uint64_t k = ukeys[j];
        const int64_t r = (int64_t)(k >> 32);
        col[j] = (int32_t)(k & 0xffffffffu);
        weight[j] = (float)counts[j];
        const int64_t rprev = j == 0 ? -1 : (int64_t)(ukeys[j - 1] >> 32);
        for (int64_t rr = rprev + 1; rr <= r; ++rr) rowptr[rr] = j;
        if (j == nnz - 1)
            for (int64_t rr = r + 1; rr <= n; ++rr) rowptr[rr] = nnz;
    }
    if (nnz == 0)
        for (int64_t rr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; rr <= n; rr += stride) rowptr[rr] = 0;
}

// degree[r] = float32 sum of the row's weights (adj.py:92: adj.sum(1) on float32 CSR; integer-valued, exact).
// self_loops == 2 ('ngcf'): the diagonal is not part of the degree.
__global__ void degree_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                              const float* __restrict__ weight, int64_t n, int self_loops, float* __restrict__ degree) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += nwarps) {
        const int64_t s = rowptr[r], e = rowptr[r + 1];
        float acc = 0.f;
        for (int64_t j = s + lane; j < e; j += 32)
            if (!(self_loops == 2 && col[j] == r)) acc += weight[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) degree[r] = acc;
    }
}

__global__ void normalise_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                 const float* __restrict__ weight, const float* __restrict__ dpow, int64_t n, int mode,
                                 int self_loops, float* __restrict__ val) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += nwarps) {
        const int64_t s = rowptr[r], e = rowptr[r + 1];
        const float dr = (mode == 0 || mode == 1) ? dpow[r] : 1.f;
        for (int64_t j = s + lane; j < e; j += 32) {
            const int c = col[j];
            const float w = weight[j];
            float v;
            if (mode == 0) {
                v = __fmul_rn(__fmul_rn(dr, w), dpow[c]);          // (D*A)*D, two roundings (adj.py:97)
            } else if (mode == 1) {
                v = (self_loops == 2 && c == r) ? 1.f : __fmul_rn(dr, w);
            } else if (mode == 2) {
                v = (self_loops == 2 && c == r) ? 1.f : __fmul_rn(dpow[c], w);
            } else {
                v = w;
            }
            val[j] = v;
        }
    }
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct WsLayout {
    size_t keys_a, keys_b, counts, n_runs, cub, total;
};

static WsLayout layout(int64_t m) {
    WsLayout w{};
    size_t sort_bytes = 0, rle_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, m, 0, 32);
    cub::DeviceRunLengthEncode::Encode(nullptr, rle_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                       (int32_t*)nullptr, (int32_t*)nullptr, m);
    size_t off = 0;
    w.keys_a = off; off += align_up((size_t)m * 8);
    w.keys_b = off; off += align_up((size_t)m * 8);
    w.counts = off; off += align_up((size_t)m * 4);
    w.n_runs = off; off += 256;
    w.cub = off;    off += align_up(sort_bytes > rle_bytes ? sort_bytes : rle_bytes);
    w.total = off;
    return w;
}

static int bits_for(int64_t n) {
    int b = 1;
    while ((1ll << b) < n) ++b;
    return b;
}

}  // namespace tagrec

using namespace tagrec;

extern "C" size_t tagrec_csr_workspace_bytes(int64_t n_directed) {
    return layout(n_directed > 0 ? n_directed : 1).total;
}

extern "C" int tagrec_csr_build_structure(const int64_t* ui_row, const int64_t* ui_col, int64_t e_ui,
                                          const int64_t* ut_row, const int64_t* ut_col, int64_t e_ut,
                                          const int64_t* it_row, const int64_t* it_col, int64_t e_it, int64_t n_user,
                                          int64_t n_item, int64_t n_tag, int self_loops, void* workspace,
                                          size_t workspace_bytes, int64_t* rowptr, int32_t* col, float* weight,
                                          int64_t cap, float* degree, int64_t* nnz_host, void* stream) {
    TAGREC_REQUIRE(e_ui >= 0 && e_ut >= 0 && e_it >= 0, "negative edge count");
    TAGREC_REQUIRE(e_ui == 0 || (ui_row && ui_col), "ui arrays missing");
    TAGREC_REQUIRE(e_ut == 0 || (ut_row && ut_col), "ut arrays missing");
    TAGREC_REQUIRE(e_it == 0 || (it_row && it_col), "it arrays missing");
    TAGREC_REQUIRE(rowptr && col && weight && degree && nnz_host && workspace, "output / workspace pointer missing");
    const bool tags = (e_ut + e_it) > 0 || n_tag > 0;
    const int64_t n = n_user + n_item + (tags ? n_tag : 0);
    TAGREC_REQUIRE(n > 0 && n < (1ll << 31), "node count out of range (int32 columns)");
    const int64_t n_diag = self_loops ? n : 0;
    const int64_t m = 2 * (e_ui + e_ut + e_it) + n_diag;
    if (cap < m) return fail(TAGREC_ENOMEM, "col/weight capacity below number of directed entries", __FILE__, __LINE__);
    cudaStream_t st = (cudaStream_t)stream;
    if (m == 0) {
        TAGREC_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(n + 1) * 8, st));
        TAGREC_CUDA(cudaMemsetAsync(degree, 0, (size_t)n * 4, st));
        TAGREC_CUDA(cudaStreamSynchronize(st));
        *nnz_host = 0;
        return TAGREC_OK;
    }
    const WsLayout w = layout(m);
    if (workspace_bytes < w.total) return fail(TAGREC_ENOMEM, "workspace too small", __FILE__, __LINE__);
    char* ws = static_cast<char*>(workspace);
    uint64_t* keys_a = reinterpret_cast<uint64_t*>(ws + w.keys_a);
    uint64_t* keys_b = reinterpret_cast<uint64_t*>(ws + w.keys_b);
    int32_t* counts = reinterpret_cast<int32_t*>(ws + w.counts);
    int32_t* n_runs = reinterpret_cast<int32_t*>(ws + w.n_runs);
    void* cub_ws = ws + w.cub;
    size_t cub_bytes = w.total - w.cub;

    Blocks b{ui_row, ui_col, ut_row, ut_col, it_row, it_col, e_ui, e_ut, e_it, n_user, n_item, tags ? n_tag : 0, n_diag};
    const int grid = kSMs * 8;
    TAGREC_LAUNCH(emit_keys_kernel, grid, 256, 0, st, b, keys_a);
    const int nb = bits_for(n);
    // LSD: column bits first, then row bits (stable) -> ascending (row, col)
    TAGREC_CUDA(cub::DeviceRadixSort::SortKeys(cub_ws, cub_bytes, keys_a, keys_b, m, 0, nb, st));
    TAGREC_CUDA(cub::DeviceRadixSort::SortKeys(cub_ws, cub_bytes, keys_b, keys_a, m, 32, 32 + nb, st));
    TAGREC_CUDA(cub::DeviceRunLengthEncode::Encode(cub_ws, cub_bytes, keys_a, keys_b, counts, n_runs, m, st));
    TAGREC_LAUNCH(decode_kernel, grid, 256, 0, st, keys_b, counts, n_runs, n, rowptr, col, weight);
    TAGREC_LAUNCH(degree_kernel, grid, 256, 0, st, rowptr, col, weight, n, self_loops, degree);
    int32_t runs_h = 0;
    TAGREC_CUDA(cudaMemcpyAsync(&runs_h, n_runs, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    TAGREC_CUDA(cudaStreamSynchronize(st));
    *nnz_host = runs_h;
    return TAGREC_OK;
}

extern "C" int tagrec_csr_normalise(const int64_t* rowptr, const int32_t* col, const float* weight, const float* dpow,
                                    int64_t n, int mode, int self_loops, float* val, void* stream) {
    TAGREC_REQUIRE(rowptr && col && weight && val, "null pointer");
    TAGREC_REQUIRE(mode >= 0 && mode <= 3, "mode must be 0..3");
    TAGREC_REQUIRE(mode == 3 || dpow, "dpow missing");
    if (n == 0) return TAGREC_OK;
    TAGREC_LAUNCH(normalise_kernel, kSMs * 8, 256, 0, stream, rowptr, col, weight, dpow, n, mode, self_loops, val);
    return TAGREC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Column-window bounds of selected rows (plan builder of the column-blocked K1, adj.py): bounds[w * n_sel + i] = index
// of the first stored entry of row rows[i] whose column id is >= w * window  (w = 0 .. n_win; ascending columns per
// row make it a lower bound).  One thread per (row, boundary).
namespace tagrec {
__global__ void __launch_bounds__(256)
window_bounds_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ rows,
                     int64_t n_sel, int64_t window, int n_win, int64_t* __restrict__ bounds) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_sel * (n_win + 1)) return;
    const int64_t w = idx / n_sel, i = idx - w * n_sel;
    const int64_t r = rows[i];
    int64_t lo = __ldg(rowptr + r), hi = __ldg(rowptr + r + 1);
    const int64_t key = w * window;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(col + mid) < key) lo = mid + 1; else hi = mid;
    }
    bounds[idx] = lo;
}
}  // namespace tagrec

extern "C" int tagrec_csr_window_bounds(const int64_t* rowptr, const int32_t* col, const int32_t* rows, int64_t n_sel,
                                        int64_t window, int n_win, int64_t* bounds, void* stream) {
    TAGREC_REQUIRE(rowptr && col && rows && bounds, "null pointer");
    TAGREC_REQUIRE(window > 0 && n_win >= 1 && n_sel >= 0, "bad window / count");
    if (n_sel == 0) return TAGREC_OK;
    const int64_t total = n_sel * (n_win + 1);
    const int64_t grid = (total + 255) / 256;
    TAGREC_REQUIRE(grid < (1ll << 31), "grid too large");
    TAGREC_LAUNCH(tagrec::window_bounds_kernel, (unsigned)grid, 256, 0, stream, rowptr, col, rows, n_sel, window, n_win, bounds);
    return TAGREC_OK;
}
