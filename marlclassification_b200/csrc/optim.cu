// Fused Adam over the flat parameter bucket (reference: th.optim.Adam in
// training/trainer.py:33,114-116 -- lr, betas (0.9, 0.999), eps 1e-8, no weight
// decay).  One launch for all 54 tensors; the step counter lives on the device so
// the update is CUDA-graph replayable.  SURVEY 8f rank 1.
#include "../../include/marlc.h"
#include "common.cuh"

namespace marlc {

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long n, float lr, float b1, float b2, float eps, float grad_scale,
                            const long long* __restrict__ step) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double t = (double)(step[0] + 1);
    const float bc1 = (float)(1.0 - pow((double)b1, t));
    const float bc2 = (float)(1.0 - pow((double)b2, t));
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] -= (lr / bc1) * (mi / denom);
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
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, (long)n, lr, beta1,
                                                           beta2, eps, grad_scale, (const long long*)step);
    MARLC_LAUNCH_CHECK();
    adam_tick_kernel<<<1, 1, 0, s>>>((long long*)step);
    MARLC_LAUNCH_CHECK();
    return 0;
}
