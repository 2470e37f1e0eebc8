// Fused Adam over the flat parameter bucket (reference: th.optim.Adam in
// training/trainer.py:33,114-116 -- lr, betas (0.9, 0.999), eps 1e-8, no weight
// decay).  One launch for all 54 tensors; the step counter lives on the device so
// the update is CUDA-graph replayable.  SURVEY 8f rank 1.
#include "../../include/marlc.h"
#include "common.cuh"

namespace marlc {

// Grid-stride, 128-bit accesses.  The bias corrections need double-precision pow (th.optim.Adam computes
// 1 - beta**step in Python doubles); they are computed by ONE thread per block - every thread doing two
// fp64 pows made this 27 MB streaming kernel take 35 us on B200's token fp64 units.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long n, float lr, float b1, float b2, float eps,
                                                   float grad_scale, const long long* __restrict__ step) {
    __shared__ float s_bc[2];
    if (threadIdx.x == 0) {
        const double t = (double)(step[0] + 1);
        s_bc[0] = (float)(1.0 - pow((double)b1, t));
        s_bc[1] = (float)(1.0 - pow((double)b2, t));
    }
    __syncthreads();
    const float bc1 = s_bc[0], bc2 = s_bc[1];
    const float step_size = lr / bc1, sq_bc2 = sqrtf(bc2);
    auto upd = [&](float& pi, float gi, float& mi, float& vi) {
        gi *= grad_scale;
        mi = b1 * mi + (1.f - b1) * gi;
        vi = b2 * vi + (1.f - b2) * gi * gi;
        const float denom = sqrtf(vi) / sq_bc2 + eps;
        pi -= step_size * (mi / denom);
    };
    const long n4 = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0) ? (n >> 2) : 0;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 pp = p4[i], mm = m4[i], vv = v4[i];
        const float4 gg = g4[i];
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
    for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float pi = p[i], mi = m[i], vi = v[i];
        upd(pi, g[i], mi, vi);
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}
__global__ void adam_tick_kernel(long long* step) { step[0] += 1; }

}  // namespace marlc

extern "C" int marlc_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                               float lr, float beta1, float beta2, float eps, float grad_scale, int64_t* step,
                               void* stream) {
    using namespace marlc;
    MARLC_CHECK(params && grads && exp_avg && exp_avg_sq && step, "adam_step: null pointer");
    if (n <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const long blocks = std::max<long>(1, std::min<long>((n / 4 + 255) / 256, 148L * 8));
    adam_kernel<<<(unsigned)blocks, 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, (long)n, lr, beta1,
                                                           beta2, eps, grad_scale, (const long long*)step);
    MARLC_LAUNCH_CHECK();
    adam_tick_kernel<<<1, 1, 0, s>>>((long long*)step);
    MARLC_LAUNCH_CHECK();
    return 0;
}
