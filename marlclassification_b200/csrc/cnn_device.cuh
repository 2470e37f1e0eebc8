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
__device__ __forceinline__ bool cnn_fwd_t_ok(const CnnFwdArgs& a, int nt);
__device__ __forceinline__ void cnn_fwd_block_t(const CnnFwdArgs& a, const int m, float* sm);

__device__ __forceinline__ void cnn_fwd_block(const CnnFwdArgs& a, const int m, float* sm) {
    if (cnn_fwd_t_ok(a, blockDim.x)) { cnn_fwd_block_t(a, m, sm); return; }
    float* in = sm;
    float* nxt = sm + a.padsz;
    float* yb = nxt + a.padsz;
    float* ws = yb + a.ysz;
    const CnnDesc& d = a.d;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int f = d.f, ff = f * f;
#ifdef MARLC_CNN_TRACE  // timeline of window 0 (cycles since entry)
    __shared__ long long cnn_tr[24];
    const long long t_entry = clock64();
    int tr_i = 0;
#define CNN_TRACE() do { if (m == 0 && tid == 0 && tr_i < 24) cnn_tr[tr_i] = clock64() - t_entry; ++tr_i; } while (0)
#else
#define CNN_TRACE() do { } while (0)
#endif

    // ---- weight prefetch: if every layer's weights fit in the staging buffer together, ALL of them
    //      are requested now and stream in while the window is gathered and the earlier layers
    //      compute; otherwise layers are staged one by one (synchronously, in output-channel chunks)
    //      right before use.  A weight row whose byte size is a multiple of 16 travels as ONE bulk
    //      copy (cp.async.bulk, completion on a per-layer mbarrier): one instruction per row instead of
    //      one 16-byte cp.async per 4 floats - issuing those took 6300 cycles of a 41000-cycle block
    //      (in-kernel trace).  Other rows (the 27-float rows of the first layer) use 4-byte cp.async.
    __shared__ __align__(8) uint64_t wbar[MAX_CNN_LAYERS];
    __shared__ float s_red[32];  // per-warp partial sums of the GroupNorm statistics
    int woff[MAX_CNN_LAYERS];
    bool prefetched;
    unsigned bulk_mask = 0;  // layers whose weights arrive through wbar[l]
    {
        int tot = 0;
        for (int l = 0; l < d.L; ++l) { woff[l] = tot; tot += d.cout[l] * (((d.cin[l] * 9 + 3) & ~3) + 4); }
        prefetched = tot <= a.wbuf;
        if (prefetched) {
            for (int l = 0; l < d.L; ++l)
                if (((d.cin[l] * 9) & 3) == 0 && ((reinterpret_cast<uintptr_t>(d.w[l]) & 15) == 0)) bulk_mask |= 1u << l;
            if (tid == 0) {
                for (int l = 0; l < d.L; ++l)
                    if (bulk_mask >> l & 1) {
                        const uint32_t ba = (uint32_t)__cvta_generic_to_shared(&wbar[l]);
                        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba));
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba),
                                     "r"((uint32_t)(d.cout[l] * d.cin[l] * 9 * 4)) : "memory");
                    }
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();  // barriers initialised and armed before any copy can complete on them
            for (int l = 0; l < d.L; ++l) {
                const int wrow = d.cin[l] * 9, Pp = ((wrow + 3) & ~3) + 4, rows = d.cout[l];
                const float* __restrict__ w = d.w[l];
                float* dst = ws + woff[l];
                if (bulk_mask >> l & 1) {
                    const uint32_t ba = (uint32_t)__cvta_generic_to_shared(&wbar[l]);
                    for (int r = tid; r < rows; r += nt) {
                        const uint32_t da = (uint32_t)__cvta_generic_to_shared(dst + r * Pp);
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(da),
                                     "l"(w + (long)r * wrow), "r"((uint32_t)(wrow * 4)), "r"(ba) : "memory");
                    }
                } else {
                    for (int r = warp; r < rows; r += nwarps)
                        for (int c1 = lane; c1 < wrow; c1 += 32) {
                            const uint32_t da = (uint32_t)__cvta_generic_to_shared(dst + r * Pp + c1);
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(da), "l"(w + (long)r * wrow + c1) : "memory");
                        }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");  // (empty group for bulk layers: keeps the group count = layer index)
            }
        }
    }
    CNN_TRACE();  // 0: weight prefetch issued
    // ---- load the window into the zero-bordered buffer (gather fused; MnistCnn keeps channel 0
    //      only: cin[0] < img_c)
    {
        const int hp = f + 2, c0 = d.cin[0];
        for (int e = tid; e < c0 * hp * hp; e += nt) in[e] = 0.f;
        __syncthreads();
        pdl_wait();  // weights are launch constants; the window position / patch comes from the previous kernel
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
    CNN_TRACE();  // 1: window gathered

    for (int l = 0; l < d.L; ++l) {
        const int ci_n = d.cin[l], co_n = d.cout[l], hi = d.hin[l], ho = d.hout[l];
        const int hp = hi + 2, hp2 = hp * hp, npos = ho * ho, total = co_n * npos;
        const int wrow = ci_n * 9, P = prefetched ? ((wrow + 3) & ~3) + 4 : cnn_wpitch(wrow);
        const int cc = prefetched ? co_n : min(co_n, a.wbuf / P);
        const float* wl_s = prefetched ? ws + woff[l] : ws;
        if (prefetched && (bulk_mask >> l & 1)) {
            const uint32_t ba = (uint32_t)__cvta_generic_to_shared(&wbar[l]);
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(ba) : "memory");
            } while (!done);
        } else if (prefetched) {  // wait for this layer's commit group (groups complete in order)
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
            CNN_TRACE();  // 2+3l: weights of layer l available
            for (int idx = tid; idx < ncur * npos; idx += nt) {
                const int c = idx / npos, pos = idx - c * npos, oy = pos / ho, ox = pos - oy * ho;
                const float* wq = wl_s + c * P;
                const float* x = in + (2 * oy) * hp + 2 * ox;
                float a0 = bias[co0 + c], a1 = 0.f, a2 = 0.f;
#pragma unroll 4  // keeps the loads of the next input channels in flight under the FMA chains
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
        CNN_TRACE();  // 3+3l: convolution done
        // GroupNorm + SiLU; the channels of one group are contiguous in yb
        const int G = d.groups[l], cpg = co_n / G, ng = cpg * npos;
        const float inv = 1.0f / (float)ng;
        float* og = a.out + (long)m * a.ldo;
        float* ogl = a.out_lo ? a.out_lo + (long)m * a.ldo : nullptr;
        // All warps take part: with G < warps, warps/G warps share one group (partial sums meet in
        // shared memory); with one warp per group only G of the 8 warps worked and the two-group first
        // layer took 6800 cycles (in-kernel trace).
        const int wpg = (G < nwarps && nwarps % G == 0) ? nwarps / G : 1;  // warps per group
        for (int g0 = 0; g0 < G; g0 += (wpg > 1 ? G : nwarps)) {
            const int g = wpg > 1 ? warp / wpg : g0 + warp;
            const int sw = wpg > 1 ? warp % wpg : 0;
            const bool act = g < G;
            const float* base = yb + (act ? g : 0) * ng;
            float s = 0.f;
            if (act) for (int e = sw * 32 + lane; e < ng; e += wpg * 32) s += base[e];
            s = warp_sum(s);
            if (wpg > 1) {
                if (lane == 0) s_red[warp] = s;
                __syncthreads();
                s = 0.f;
                for (int i = 0; i < wpg; ++i) s += s_red[g * wpg + i];
                __syncthreads();
            }
            const float mean = s * inv;
            float v = 0.f;
            if (act) for (int e = sw * 32 + lane; e < ng; e += wpg * 32) { const float dd = base[e] - mean; v += dd * dd; }
            v = warp_sum(v);
            if (wpg > 1) {
                if (lane == 0) s_red[warp] = v;
                __syncthreads();
                v = 0.f;
                for (int i = 0; i < wpg; ++i) v += s_red[g * wpg + i];
            }
            const float rstd = 1.0f / sqrtf(v * inv + GN_EPS);
            if (act)
                for (int e = sw * 32 + lane; e < ng; e += wpg * 32) {
                    const int c = g * cpg + e / npos, pos = e % npos;
                    const float o = siluf_((base[e] - mean) * rstd * d.gn_w[l][c] + d.gn_b[l][c]);
                    if (last) { og[c * npos + pos] = o; if (ogl) ogl[c * npos + pos] = tf32_lo(o); }
                    else nxt[c * hop * hop + (pos / ho + 1) * hop + (pos % ho) + 1] = o;
                }
            if (wpg > 1) break;  // every group was handled in this single pass
        }
        __syncthreads();
        CNN_TRACE();  // 4+3l: GroupNorm + SiLU done
        float* t = in; in = nxt; nxt = t;
    }
#ifdef MARLC_CNN_TRACE
    if (m == 0 && tid == 0) {
        printf("cnn trace:");
        for (int i = 0; i < tr_i && i < 24; ++i) printf(" %lld", cnn_tr[i]);
        printf(" | end %lld\n", clock64() - t_entry);
    }
#endif
}


// ---------------------------------------------------------------------------------------------
// Register-tiled forward block (used when the engine provides transposed weights wT[l] =
// [cin*9][cout] and all layers fit in the staging buffer).  Measured on the previous version with
// an in-kernel trace (C2, 41 000 cycles per window): 15 % issuing ~6000 16-byte cp.async, 38 % in
// conv loops that were bound by the shared-memory pipe (2 LDS per FMA, one warp-wide LDS per cycle
// per SM), 30 % in GroupNorm passes run by G of the 8 warps with integer divisions and global
// gamma/beta loads per element.  Here:
//   * each layer's weights arrive with ONE bulk copy (cp.async.bulk + mbarrier);
//   * a work item is 4 consecutive output channels x 1 output position: the 4 weights of a tap are
//     one 128-bit shared load, the input pixel one 32-bit load -> 0.5 LDS per FMA; the input
//     channels are split over `ks` adjacent lanes when there are fewer items than threads and
//     recombined with shuffles;
//   * conv results stay in registers through GroupNorm: statistics via shared-memory atomics
//     (two-pass, as the reference), gamma/beta requested before the statistics, no divisions.
// ---------------------------------------------------------------------------------------------
constexpr int CNN_T_MAXE = 10;  // max outputs per thread kept in registers through GroupNorm

__device__ __forceinline__ bool cnn_fwd_t_ok(const CnnFwdArgs& a, int nt) {
    const CnnDesc& d = a.d;
    if (!d.wT[0]) return false;
    int tot = 0;
    for (int l = 0; l < d.L; ++l) {
        if (!d.wT[l] || (d.cout[l] & 3) || d.groups[l] > 32) return false;
        {   // GroupNorm register budget: elements of one group per lane of its warps
            const int nw = nt >> 5, G = d.groups[l], ngl = (d.cout[l] / G) * d.hout[l] * d.hout[l];
            const int wpg = G <= nw ? nw / G : 1;
            if ((G <= nw && nw % G != 0) || (ngl + wpg * 32 - 1) / (wpg * 32) > CNN_T_MAXE) return false;
        }
        tot += d.cout[l] * d.cin[l] * 9;
    }
    return tot <= a.wbuf;
}

__device__ __forceinline__ void cnn_fwd_block_t(const CnnFwdArgs& a, const int m, float* sm) {
    float* in = sm;
    float* nxt = sm + a.padsz;
    float* yb = nxt + a.padsz;  // conv outputs [kslice][c][pos]: one plane per input-channel slice
    float* ws = yb + a.ysz;
    const CnnDesc& d = a.d;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int f = d.f, ff = f * f;
    __shared__ __align__(8) uint64_t wbar[MAX_CNN_LAYERS];
    __shared__ float s_red[32];  // per-warp partial sums of the GroupNorm statistics
#ifdef MARLC_CNN_TRACE
    __shared__ long long cnn_tr[24];
    const long long t_entry = clock64();
    int tr_i = 0;
#define CNN_TRACE_T() do { if (m == 0 && tid == 0 && tr_i < 24) cnn_tr[tr_i] = clock64() - t_entry; ++tr_i; } while (0)
#else
#define CNN_TRACE_T() do { } while (0)
#endif
    int woff[MAX_CNN_LAYERS];
    {
        int tot = 0;
        for (int l = 0; l < d.L; ++l) { woff[l] = tot; tot += d.cout[l] * d.cin[l] * 9; }
    }
    if (tid == 0) {
        for (int l = 0; l < d.L; ++l) {
            const uint32_t ba = (uint32_t)__cvta_generic_to_shared(&wbar[l]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int l = 0; l < d.L; ++l) {  // one bulk copy per layer, all in flight from the first cycle
            const uint32_t ba = (uint32_t)__cvta_generic_to_shared(&wbar[l]);
            const uint32_t bytes = (uint32_t)(d.cout[l] * d.cin[l] * 9 * 4);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(ws + woff[l])), "l"(d.wT[l]), "r"(bytes), "r"(ba) : "memory");
        }
    }
    CNN_TRACE_T();  // 0: weight copies issued
    // ---- window -> zero-bordered buffer (gather fused; MnistCnn keeps channel 0 only)
    {
        const int hp = f + 2, c0 = d.cin[0];
        for (int e = tid; e < c0 * hp * hp; e += nt) in[e] = 0.f;
        __syncthreads();  // (also publishes the mbarrier initialisation to the waiting threads)
        pdl_wait();  // weights are launch constants; the window position / patch comes from the previous kernel
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
    CNN_TRACE_T();  // 1: window gathered

    for (int l = 0; l < d.L; ++l) {
        const int ci_n = d.cin[l], co_n = d.cout[l], hi = d.hin[l], ho = d.hout[l];
        const int hp = hi + 2, hp2 = hp * hp, npos = ho * ho, hop = ho + 2, total = co_n * npos;
        const int G = d.groups[l], cpg = co_n / G, ncg = co_n >> 2;
        const bool last = (l + 1 == d.L);
        const float* __restrict__ wl = ws + woff[l];
        const int items = ncg * npos;
        int ks = 1;  // input channels split over ks groups of threads, one partial plane of yb each
        while (ks < 8 && items * ks * 2 <= nt && ks * 2 <= ci_n && (ks * 2) * total <= a.ysz) ks <<= 1;
        if (!last)
            for (int e = tid; e < co_n * hop * hop; e += nt) nxt[e] = 0.f;  // border of the next input
        {  // this layer's weights have landed?
            const uint32_t ba = (uint32_t)__cvta_generic_to_shared(&wbar[l]);
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(ba) : "memory");
            } while (!done);
        }
        CNN_TRACE_T();  // 2+3l: weights available
        // ---- convolution.  Work index = kslice * items + pos * ncg + cg: the lanes of a warp differ in
        //      the channel group first (their 128-bit weight loads cover distinct banks) and share the
        //      input pixel whenever they share the position (broadcast).
        const float ritems = 1.0f / (float)items, rncg = 1.0f / (float)ncg, rho_c = 1.0f / (float)ho;
        for (int wk = tid; wk < items * ks; wk += nt) {
            const int kslice = __float2int_rz(((float)wk + 0.5f) * ritems), item = wk - kslice * items;
            const int pos = __float2int_rz(((float)item + 0.5f) * rncg), c4 = (item - pos * ncg) << 2;
            const int oy = __float2int_rz(((float)pos + 0.5f) * rho_c), ox = pos - oy * ho;
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kslice == 0) s4 = *reinterpret_cast<const float4*>(d.b[l] + c4);
            const float* x = in + kslice * hp2 + (2 * oy) * hp + 2 * ox;
            const float* wq = wl + (long)(kslice * 9) * co_n + c4;
#pragma unroll 2
            for (int ci = kslice; ci < ci_n; ci += ks, x += ks * hp2, wq += ks * 9 * co_n) {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float xv = x[ky * hp + kx];
                        const float4 w4 = *reinterpret_cast<const float4*>(wq + (ky * 3 + kx) * co_n);
                        s4.x = fmaf(xv, w4.x, s4.x); s4.y = fmaf(xv, w4.y, s4.y);
                        s4.z = fmaf(xv, w4.z, s4.z); s4.w = fmaf(xv, w4.w, s4.w);
                    }
            }
            float* y4 = yb + kslice * total + c4 * npos + pos;
            y4[0] = s4.x; y4[npos] = s4.y; y4[2 * npos] = s4.z; y4[3 * npos] = s4.w;
        }
        __syncthreads();
        CNN_TRACE_T();  // 3+3l: convolution done
        // ---- GroupNorm + SiLU with the values held in registers.  Warps are bound to groups: with
        //      G <= warps, warps/G warps share a group (their partial sums meet in s_red: no atomics -
        //      a float or 64-bit atomicAdd on shared memory is a compare-and-swap loop and serialised
        //      for thousands of cycles on the few group counters); with G > warps each warp owns G/warps
        //      whole groups.  Index arithmetic uses exact float reciprocals ((e + 0.5) / n is never
        //      within rounding distance of an integer for e < 2^16) and the activation the MUFU
        //      exponential / reciprocal: integer divisions and accurate expf made this phase ~250
        //      dependent instructions per element.
        const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
        const int ng = cpg * npos;
        const int wpg = G <= nwarps ? nwarps / G : 1;          // warps per group
        const int gpw = G <= nwarps ? 1 : (G + nwarps - 1) / nwarps;  // groups per warp
        float* ysl = a.y_save[l] ? a.y_save[l] + (long)m * total : nullptr;
        float* og = a.out + (long)m * a.ldo;
        float* ogl = a.out_lo ? a.out_lo + (long)m * a.ldo : nullptr;
        const float inv = 1.0f / (float)ng, rnpos = 1.0f / (float)npos, rho = 1.0f / (float)ho;
        for (int gi = 0; gi < gpw; ++gi) {
            const int g = G <= nwarps ? warp / wpg : warp * gpw + gi;
            const int sub = G <= nwarps ? warp % wpg : 0;
            const bool gok = g < G;
            float xv[CNN_T_MAXE], gam[CNN_T_MAXE], bet[CNN_T_MAXE];
            int ce[CNN_T_MAXE];
            float s1 = 0.f;
#pragma unroll
            for (int k = 0; k < CNN_T_MAXE; ++k) {
                const int el = sub * 32 + lane + k * wpg * 32;  // index inside the group
                if (k * wpg * 32 >= ng) break;  // block-uniform
                const bool ok = gok && el < ng;
                const int e = g * ng + el;
                xv[k] = 0.f; ce[k] = 0;
                if (ok) {
                    ce[k] = __float2int_rz(((float)e + 0.5f) * rnpos);
                    for (int sl = 0; sl < ks; ++sl) xv[k] += yb[sl * total + e];  // combine the input-channel slices
                    gam[k] = __ldg(d.gn_w[l] + ce[k]); bet[k] = __ldg(d.gn_b[l] + ce[k]);  // consumed after the statistics
                    if (ysl) ysl[e] = xv[k];  // saved for backward (coalesced)
                }
                s1 += xv[k];
            }
            CNN_TRACE_T();
            s1 = warp_sum(s1);
            if (wpg > 1) {
                if (lane == 0) s_red[warp] = s1;
                __syncthreads();
                s1 = 0.f;
                for (int i = 0; i < wpg; ++i) s1 += s_red[(warp / wpg) * wpg + i];
                __syncthreads();
            }
            CNN_TRACE_T();
            const float mean = s1 * inv;
            float s2 = 0.f;
#pragma unroll
            for (int k = 0; k < CNN_T_MAXE; ++k) {
                const int el = sub * 32 + lane + k * wpg * 32;
                if (k * wpg * 32 >= ng) break;
                if (gok && el < ng) { const float dd = xv[k] - mean; s2 += dd * dd; }
            }
            s2 = warp_sum(s2);
            if (wpg > 1) {
                if (lane == 0) s_red[warp] = s2;
                __syncthreads();
                s2 = 0.f;
                for (int i = 0; i < wpg; ++i) s2 += s_red[(warp / wpg) * wpg + i];
            }
            CNN_TRACE_T();
            const float rstd = rsqrtf(s2 * inv + GN_EPS);
#pragma unroll
            for (int k = 0; k < CNN_T_MAXE; ++k) {
                const int el = sub * 32 + lane + k * wpg * 32;
                if (k * wpg * 32 >= ng) break;
                if (gok && el < ng) {
                    const int e = g * ng + el, c = ce[k], pos = e - c * npos;
                    const float z = (xv[k] - mean) * rstd * gam[k] + bet[k];
                    const float o = __fdividef(z, 1.0f + __expf(-z));  // SiLU
                    if (last) { og[e] = o; if (ogl) ogl[e] = tf32_lo(o); }
                    else {
                        const int oy = __float2int_rz(((float)pos + 0.5f) * rho), ox = pos - oy * ho;
                        nxt[c * hop * hop + (oy + 1) * hop + ox + 1] = o;
                    }
                }
            }
        }
        __syncthreads();
        CNN_TRACE_T();  // 4+3l: GroupNorm + SiLU done
        float* t = in; in = nxt; nxt = t;
    }
#ifdef MARLC_CNN_TRACE
    if (m == 0 && tid == 0) {
        printf("cnn trace (tiled):");
        for (int i = 0; i < tr_i && i < 24; ++i) printf(" %lld", cnn_tr[i]);
        printf(" | end %lld\n", clock64() - t_entry);
    }
#endif
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
    *ysz = (max(y, 1024) + 3) & ~3;  // >= 4 x 256: room for the input-channel-slice planes of the tiled block
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
