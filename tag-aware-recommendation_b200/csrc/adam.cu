// Fused dense Adam — one pass over param / grad / exp_avg / exp_avg_sq (28 B per element) instead of the ~10
// multi-tensor passes of torch.optim.Adam (com.py:25).  Same update rule as torch's single-tensor Adam
// (amsgrad=False, maximize=False):  m.lerp_(g, 1-b1);  v = b2*v + (1-b2)*g*g;
// p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
#include "common.cuh"

namespace tagrec {

__global__ void __launch_bounds__(256)
adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
            int64_t n4, float* ps, const float* gs, float* ms, float* vs, int64_t n, float b1, float b2, float eps,
            float wd, float step_size, float inv_sqrt_bc2, const float* __restrict__ scal, Mirror pm) {
    if (scal) {     // capturable mode: the bias corrections of THIS step live in device memory
        step_size = __ldg(scal);
        inv_sqrt_bc2 = __ldg(scal + 1);
    }
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        adam_update(pp, gg, mm, vv, b1, b2, eps, wd, step_size, inv_sqrt_bc2);
    };
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = p[i], mm = m[i], vv = v[i];
        const float4 gg = __ldcs(g + i);
        upd(pp.x, gg.x, mm.x, vv.x);
        upd(pp.y, gg.y, mm.y, vv.y);
        upd(pp.z, gg.z, mm.z, vv.z);
        upd(pp.w, gg.w, mm.w, vv.w);
        store_row(reinterpret_cast<float*>(p), pm, i, pp);      // pm.n > 0: the new value goes to every rank's copy
        m[i] = mm; v[i] = vv;
    }
    // scalar tail
    for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        upd(ps[i], gs[i], ms[i], vs[i]);
}

}  // namespace tagrec

using namespace tagrec;

extern "C" int tagrec_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, float lr, float beta1,
                                float beta2, float eps, float weight_decay, int64_t step, void* stream) {
    TAGREC_REQUIRE(param && grad && m && v, "null pointer");
    TAGREC_REQUIRE(step >= 1, "step counts from 1");
    if (n == 0) return TAGREC_OK;
    const bool aligned = (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
    const int64_t n4 = aligned ? n / 4 : 0;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    int64_t blocks = (n4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    TAGREC_LAUNCH(adam_kernel, (unsigned)blocks, 256, 0, stream, reinterpret_cast<float4*>(param),
                  reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n4,
                  param, grad, m, v, n, beta1, beta2, eps, weight_decay, step_size, inv_sqrt_bc2,
                  static_cast<const float*>(nullptr), Mirror{});
    return TAGREC_OK;
}

// Owner-sharded update (multi-GPU, no reference equivalent): the same step on a contiguous segment of a parameter table
// this rank owns; the new parameter values are stored through `param_mirror` (tagrec_mirror_t over the SEGMENT's first
// element on every rank / its multicast address) so that every rank's replica of the table is updated by its owner,
// while exp_avg / exp_avg_sq stay local.  n must be a multiple of 4 and all pointers 16-byte aligned.
extern "C" int tagrec_adam_step_mirror(float* param, const float* grad, float* m, float* v, int64_t n, float lr,
                                       float beta1, float beta2, float eps, float weight_decay, int64_t step,
                                       const tagrec_mirror_t* param_mirror, void* stream) {
    TAGREC_REQUIRE(param && grad && m && v, "null pointer");
    TAGREC_REQUIRE(step >= 1, "step counts from 1");
    TAGREC_REQUIRE(n % 4 == 0 && (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) == 0,
                   "segment must be 16-byte aligned and a multiple of 4 floats");
    Mirror pm;
    if (int rc = set_mirror(pm, param_mirror)) return rc;
    if (n == 0) return TAGREC_OK;
    const int64_t n4 = n / 4;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    TAGREC_LAUNCH(adam_kernel, (unsigned)blocks, 256, 0, stream, reinterpret_cast<float4*>(param),
                  reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n4,
                  param, grad, m, v, n, beta1, beta2, eps, weight_decay, step_size, inv_sqrt_bc2,
                  static_cast<const float*>(nullptr), pm);
    return TAGREC_OK;
}

namespace tagrec {
// step += 1;  scal = { lr / (1 - b1^step), 1 / sqrt(1 - b2^step) }   (one thread; CUDA-graph capturable)
__global__ void adam_scalars_kernel(int64_t* step, float lr, float b1, float b2, float* scal) {
    const int64_t t = *step + 1;
    *step = t;
    scal[0] = (float)((double)lr / (1.0 - pow((double)b1, (double)t)));
    scal[1] = (float)(1.0 / sqrt(1.0 - pow((double)b2, (double)t)));
}
}  // namespace tagrec

extern "C" int tagrec_adam_advance(int64_t* step_dev, float lr, float beta1, float beta2, float* scal_dev,
                                   void* stream) {
    TAGREC_REQUIRE(step_dev && scal_dev, "null pointer");
    TAGREC_LAUNCH(adam_scalars_kernel, 1, 1, 0, stream, step_dev, lr, beta1, beta2, scal_dev);
    return TAGREC_OK;
}

extern "C" int tagrec_adam_step_dev(float* param, const float* grad, float* m, float* v, int64_t n, float beta1,
                                    float beta2, float eps, float weight_decay, const float* scal_dev, void* stream) {
    TAGREC_REQUIRE(param && grad && m && v && scal_dev, "null pointer");
    if (n == 0) return TAGREC_OK;
    const bool aligned = (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
    const int64_t n4 = aligned ? n / 4 : 0;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    TAGREC_LAUNCH(adam_kernel, (unsigned)blocks, 256, 0, stream, reinterpret_cast<float4*>(param),
                  reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n4,
                  param, grad, m, v, n, beta1, beta2, eps, weight_decay, 0.f, 0.f, scal_dev, Mirror{});
    return TAGREC_OK;
}
