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
    float* out_lo = nullptr;  // optional: tf32_lo(out), same leading dimension (pre-split 3xTF32 operand)
    long ldo;
    int B, H, W, M;
    int padsz;  // floats of one zero-bordered activation buffer: max_l cin_l * (hin_l + 2)^2
    int ysz;    // floats of the raw conv-output buffer: max_l cout_l * hout_l^2
    int wbuf;   // floats of the weight staging buffer
};

__host__ __device__ inline int cnn_wpitch(int wrow) { return wrow | 1; }  // odd pitch: conflict-free rows

// One window per call.  `sm` holds 2*padsz + ysz + wbuf floats.  All threads of the CTA participate.
// Activations live in shared memory with a one-pixel zero border (no bounds checks in the 3x3
// stride-2 taps); each layer's weights are staged in shared memory (odd pitch) before use, so
// the inner loops never wait on global memory.
__device__ __forceinline__ void cnn_fwd_block(const CnnFwdArgs& a, const int m, float* sm) {
    float* in = sm;
    float* nxt = sm + a.padsz;
    float* yb = nxt + a.padsz;
    float* ws = yb + a.ysz;
    const CnnDesc& d = a.d;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int f = d.f, ff = f * f;

    // ---- weight prefetch: if every layer's weights fit in the staging buffer together, ALL of them
    //      are requested now with cp.async (one commit group per layer) and stream in while the
    //      window is gathered and the earlier layers compute; otherwise layers are staged one by
    //      one (synchronously, in output-channel chunks) right before use.
    int woff[MAX_CNN_LAYERS];
    bool prefetched;
    {
        int tot = 0;
        for (int l = 0; l < d.L; ++l) { woff[l] = tot; tot += d.cout[l] * (((d.cin[l] * 9 + 3) & ~3) + 4); }
        prefetched = tot <= a.wbuf;
        if (prefetched) {
            for (int l = 0; l < d.L; ++l) {
                const int wrow = d.cin[l] * 9, Pp = ((wrow + 3) & ~3) + 4, rows = d.cout[l];
                const float* __restrict__ w = d.w[l];
                float* dst = ws + woff[l];
                if ((wrow & 3) == 0 && ((reinterpret_cast<uintptr_t>(w) & 15) == 0)) {
                    const int w4 = wrow >> 2;
                    for (int i = tid; i < rows * w4; i += nt) {
                        const int r = i / w4, c4 = i - r * w4;
                        const uint32_t da = (uint32_t)__cvta_generic_to_shared(dst + r * Pp + 4 * c4);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(w + (long)r * wrow + 4 * c4) : "memory");
                    }
                } else {
                    for (int i = tid; i < rows * wrow; i += nt) {
                        const int r = i / wrow, c1 = i - r * wrow;
                        const uint32_t da = (uint32_t)__cvta_generic_to_shared(dst + r * Pp + c1);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(da), "l"(w + i) : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
        }
    }
    // ---- load the window into the zero-bordered buffer (gather fused; MnistCnn keeps channel 0
    //      only: cin[0] < img_c)
    {
        const int hp = f + 2, c0 = d.cin[0];
        for (int e = tid; e < c0 * hp * hp; e += nt) in[e] = 0.f;
        __syncthreads();
        if (a.patch) {
            const float* src = a.patch + (long)m * d.img_c * ff;
            for (int e = tid; e < c0 * ff; e += nt) {
                const int c = e / ff, i = (e / f) % f, j = e % f;
                in[c * hp * hp + (i + 1) * hp + j + 1] = src[e];
            }
        } else {
            const int b = m % a.B;
            const int py = a.pos[2 * m], px = a.pos[2 * m + 1];
            const float* src = a.img + (long)b * d.img_c * a.H * a.W;
            for (int e = tid; e < c0 * ff; e += nt) {
                const int c = e / ff, i = (e / f) % f, j = e % f;
                in[c * hp * hp + (i + 1) * hp + j + 1] = __ldg(src + ((long)c * a.H + py + i) * a.W + px + j);
            }
        }
    }
    __syncthreads();

    for (int l = 0; l < d.L; ++l) {
        const int ci_n = d.cin[l], co_n = d.cout[l], hi = d.hin[l], ho = d.hout[l];
        const int hp = hi + 2, hp2 = hp * hp, npos = ho * ho, total = co_n * npos;
        const int wrow = ci_n * 9, P = prefetched ? ((wrow + 3) & ~3) + 4 : cnn_wpitch(wrow);
        const int cc = prefetched ? co_n : min(co_n, a.wbuf / P);
        const float* wl_s = prefetched ? ws + woff[l] : ws;
        if (prefetched) {  // wait for this layer's commit group (groups complete in order)
            switch (d.L - 1 - l) {
                case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
                case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
                case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
                case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
                case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
                default: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
            }
        }
        const float* __restrict__ w = d.w[l];
        const float* __restrict__ bias = d.b[l];
        float* ys = a.y_save[l] ? a.y_save[l] + (long)m * total : nullptr;
        const bool last = (l + 1 == d.L);
        const int hop = ho + 2;
        if (!last)
            for (int e = tid; e < co_n * hop * hop; e += nt) nxt[e] = 0.f;  // border of the next input
        for (int co0 = 0; co0 < co_n; co0 += cc) {
            const int ncur = min(cc, co_n - co0);
            if (!prefetched) stage_rows(ws, w + (long)co0 * wrow, ncur, wrow, P);
            __syncthreads();
            for (int idx = tid; idx < ncur * npos; idx += nt) {
                const int c = idx / npos, pos = idx - c * npos, oy = pos / ho, ox = pos - oy * ho;
                const float* wq = wl_s + c * P;
                const float* x = in + (2 * oy) * hp + 2 * ox;
                float a0 = bias[co0 + c], a1 = 0.f, a2 = 0.f;
                for (int ci = 0; ci < ci_n; ++ci, wq += 9, x += hp2) {
                    a0 = fmaf(wq[0], x[0], a0); a0 = fmaf(wq[1], x[1], a0); a0 = fmaf(wq[2], x[2], a0);
                    a1 = fmaf(wq[3], x[hp], a1); a1 = fmaf(wq[4], x[hp + 1], a1); a1 = fmaf(wq[5], x[hp + 2], a1);
                    a2 = fmaf(wq[6], x[2 * hp], a2); a2 = fmaf(wq[7], x[2 * hp + 1], a2); a2 = fmaf(wq[8], x[2 * hp + 2], a2);
                }
                const float acc = a0 + a1 + a2;
                yb[(co0 + c) * npos + pos] = acc;
                if (ys) ys[(co0 + c) * npos + pos] = acc;
            }
            __syncthreads();
        }
        // GroupNorm + SiLU; the channels of one group are contiguous in yb
        const int G = d.groups[l], cpg = co_n / G, ng = cpg * npos;
        const float inv = 1.0f / (float)ng;
        float* og = a.out + (long)m * a.ldo;
        float* ogl = a.out_lo ? a.out_lo + (long)m * a.ldo : nullptr;
        for (int g = warp; g < G; g += nwarps) {
            const float* base = yb + g * ng;
            float s = 0.f;
            for (int e = lane; e < ng; e += 32) s += base[e];
            const float mean = warp_sum(s) * inv;
            float v = 0.f;
            for (int e = lane; e < ng; e += 32) { float dd = base[e] - mean; v += dd * dd; }
            const float rstd = 1.0f / sqrtf(warp_sum(v) * inv + GN_EPS);
            for (int e = lane; e < ng; e += 32) {
                const int c = g * cpg + e / npos, pos = e % npos;
                const float o = siluf_((base[e] - mean) * rstd * d.gn_w[l][c] + d.gn_b[l][c]);
                if (last) { og[c * npos + pos] = o; if (ogl) ogl[c * npos + pos] = tf32_lo(o); }
                else nxt[c * hop * hop + (pos / ho + 1) * hop + (pos % ho) + 1] = o;
            }
        }
        __syncthreads();
        float* t = in; in = nxt; nxt = t;
    }
}

// shared-memory plan (floats) for cnn_fwd_block
inline void cnn_fwd_plan(const CnnDesc& d, int* padsz, int* ysz, int* wbuf) {
    int p = 0, y = 0, wmax = 0;
    for (int l = 0; l < d.L; ++l) {
        p = max(p, d.cin[l] * (d.hin[l] + 2) * (d.hin[l] + 2));
        y = max(y, d.cout[l] * d.hout[l] * d.hout[l]);
        wmax = max(wmax, d.cout[l] * cnn_wpitch(d.cin[l] * 9));
    }
    *padsz = (p + 3) & ~3;
    *ysz = (y + 3) & ~3;
    int wall = 0;  // all layers at once, rows padded to a multiple of 4 floats + 4 (cp.async alignment)
    for (int l = 0; l < d.L; ++l) wall += d.cout[l] * (((d.cin[l] * 9 + 3) & ~3) + 4);
    // <= 100 KB of staged weights: everything prefetched if it fits, else channel chunks per layer
    *wbuf = wall <= 25 * 1024 ? wall : min(wmax, 24 * 1024);
}
inline size_t cnn_fwd_smem_bytes(const CnnFwdArgs& a) { return sizeof(float) * ((size_t)2 * a.padsz + a.ysz + a.wbuf); }

inline int cnn_max_act(const CnnDesc& d) {
    int mx = d.cin[0] * d.f * d.f;
    for (int l = 0; l < d.L; ++l) mx = max(mx, d.cout[l] * d.hout[l] * d.hout[l]);
    return mx;
}

}  // namespace marlc
