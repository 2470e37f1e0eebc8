// Device-side body of the fused gather + CNN forward, shared by cnn.cu (stand-alone
// kernel) and chain.cu (the fused per-step "pre" kernel).
#pragma once
#include "kernels.cuh"

namespace marlc {

constexpr float GN_EPS = 1e-5f;

struct CnnFwdArgs {
    CnnDesc d;
    const float* img;
    const int* pos;
    const float* patch;
    float* y_save[MAX_CNN_LAYERS];
    float* out;
    long ldo;
    int B, H, W, M, bufsz;
};

// One window per call; `sm` holds 2 * a.bufsz floats.  All threads of the CTA participate.
__device__ __forceinline__ void cnn_fwd_block(const CnnFwdArgs& a, const int m, float* sm) {
    float* in = sm;
    float* out = sm + a.bufsz;
    const CnnDesc& d = a.d;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int f = d.f, ff = f * f;

    // ---- load the window (gather fused; MnistCnn keeps channel 0 only: cin[0] < img_c)
    if (a.patch) {
        const float* src = a.patch + (long)m * d.img_c * ff;
        for (int e = tid; e < d.cin[0] * ff; e += nt) in[e] = src[e];
    } else {
        const int b = m % a.B;
        const int py = a.pos[2 * m], px = a.pos[2 * m + 1];
        const float* src = a.img + (long)b * d.img_c * a.H * a.W;
        for (int e = tid; e < d.cin[0] * ff; e += nt) {
            const int c = e / ff, i = (e / f) % f, j = e % f;
            in[e] = __ldg(src + ((long)c * a.H + py + i) * a.W + px + j);
        }
    }
    __syncthreads();

    for (int l = 0; l < d.L; ++l) {
        const int ci_n = d.cin[l], co_n = d.cout[l], hi = d.hin[l], ho = d.hout[l];
        const int npos = ho * ho, total = co_n * npos;
        const float* __restrict__ w = d.w[l];
        const float* __restrict__ bias = d.b[l];
        float* ys = a.y_save[l] ? a.y_save[l] + (long)m * total : nullptr;
        for (int idx = tid; idx < total; idx += nt) {
            const int co = idx / npos, oy = (idx / ho) % ho, ox = idx % ho;
            const float* wb = w + (long)co * ci_n * 9;
            float acc = bias[co];
            for (int ci = 0; ci < ci_n; ++ci) {
                const float* xin = in + ci * hi * hi;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int iy = 2 * oy - 1 + ky;
                    if (iy < 0 || iy >= hi) continue;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int ix = 2 * ox - 1 + kx;
                        if (ix < 0 || ix >= hi) continue;
                        acc = fmaf(wb[ci * 9 + ky * 3 + kx], xin[iy * hi + ix], acc);
                    }
                }
            }
            out[idx] = acc;
            if (ys) ys[idx] = acc;
        }
        __syncthreads();
        // GroupNorm + SiLU in place; the channels of one group are contiguous
        const int G = d.groups[l], cpg = co_n / G, ng = cpg * npos;
        const float inv = 1.0f / (float)ng;
        for (int g = warp; g < G; g += nwarps) {
            float* base = out + g * ng;
            float s = 0.f;
            for (int e = lane; e < ng; e += 32) s += base[e];
            const float mean = warp_sum(s) * inv;
            float v = 0.f;
            for (int e = lane; e < ng; e += 32) { float dd = base[e] - mean; v += dd * dd; }
            const float rstd = 1.0f / sqrtf(warp_sum(v) * inv + GN_EPS);
            for (int e = lane; e < ng; e += 32) {
                const int c = g * cpg + e / npos;
                base[e] = siluf_((base[e] - mean) * rstd * d.gn_w[l][c] + d.gn_b[l][c]);
            }
        }
        __syncthreads();
        float* t = in; in = out; out = t;
    }
    float* o = a.out + (long)m * a.ldo;
    for (int e = tid; e < d.out_size; e += nt) o[e] = in[e];
}


inline int cnn_max_act(const CnnDesc& d) {
    int mx = d.cin[0] * d.f * d.f;
    for (int l = 0; l < d.L; ++l) mx = max(mx, d.cout[l] * d.hout[l] * d.hout[l]);
    return mx;
}

}  // namespace marlc
