// Row-wise / element-wise kernels of the agent networks (HBM/L2-bound work):
// LayerNorm+SiLU forward/backward, column sums, cross-agent message mean,
// position features, LSTM cell point-wise forward/backward.
#include "common.cuh"

namespace marlc {

constexpr float LN_EPS = 1e-5f;  // nn.LayerNorm default, used by every block of the reference

static inline int row_grid(int R, int warps_per_block) {
    int blocks = (R + warps_per_block - 1) / warps_per_block;
    return max(1, min(blocks, MARLC_SMS * 8));
}

// ---------------------------------------------------------------------------
// S = SiLU(LayerNorm(Y) * gamma + beta), one warp per row
// (message.py:26-33,42-49; state.py:13-17; policy.py:11-14,22-25; prediction.py:10-13)
// ---------------------------------------------------------------------------
__global__ void ln_silu_fwd_kernel(const float* __restrict__ Y, long ldy, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ S, long lds, int R, int N) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const float invN = 1.0f / (float)N;
    for (int r = warp; r < R; r += nwarps) {
        const float* y = Y + (long)r * ldy;
        float s = 0.f;
        for (int n = lane; n < N; n += 32) s += y[n];
        const float mean = warp_sum(s) * invN;
        float v = 0.f;
        for (int n = lane; n < N; n += 32) {
            float d = y[n] - mean;
            v += d * d;
        }
        const float rstd = 1.0f / sqrtf(warp_sum(v) * invN + LN_EPS);
        float* o = S + (long)r * lds;
        for (int n = lane; n < N; n += 32) o[n] = siluf_((y[n] - mean) * rstd * gamma[n] + beta[n]);
    }
}

int ln_silu_fwd(const float* Y, long ldy, const float* gamma, const float* beta, float* S, long lds, int R, int N,
                cudaStream_t s) {
    if (R <= 0) return 0;
    ln_silu_fwd_kernel<<<row_grid(R, 8), 256, 0, s>>>(Y, ldy, gamma, beta, S, lds, R, N);
    MARLC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------
// Backward of the block above.  Statistics are recomputed from the saved
// pre-norm Y (cheaper than storing mean/rstd: the row is read anyway).
// Column sums for dgamma / dbeta / dbias are reduced per CTA in shared memory,
// then one atomicAdd per column per CTA.
// ---------------------------------------------------------------------------
__global__ void ln_silu_bwd_kernel(const float* __restrict__ dS, long ldds, const float* __restrict__ Y, long ldy,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ dY, long lddy, float* __restrict__ dgamma,
                                   float* __restrict__ dbeta, float* __restrict__ dbias, int R, int N) {
    extern __shared__ float acc[];  // [3][N]
    for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const float invN = 1.0f / (float)N;
    for (int r = warp; r < R; r += nwarps) {
        const float* y = Y + (long)r * ldy;
        const float* ds = dS + (long)r * ldds;
        float s = 0.f;
        for (int n = lane; n < N; n += 32) s += y[n];
        const float mean = warp_sum(s) * invN;
        float v = 0.f;
        for (int n = lane; n < N; n += 32) {
            float d = y[n] - mean;
            v += d * d;
        }
        const float rstd = 1.0f / sqrtf(warp_sum(v) * invN + LN_EPS);
        // c1 = mean(dxhat), c2 = mean(dxhat * xhat)
        float c1 = 0.f, c2 = 0.f;
        for (int n = lane; n < N; n += 32) {
            float xh = (y[n] - mean) * rstd;
            float dz = ds[n] * silu_grad_(xh * gamma[n] + beta[n]);
            float dxh = dz * gamma[n];
            c1 += dxh;
            c2 += dxh * xh;
            atomicAdd(&acc[n], dz * xh);
            atomicAdd(&acc[N + n], dz);
        }
        c1 = warp_sum(c1) * invN;
        c2 = warp_sum(c2) * invN;
        float* dy = dY + (long)r * lddy;
        for (int n = lane; n < N; n += 32) {
            float xh = (y[n] - mean) * rstd;
            float dz = ds[n] * silu_grad_(xh * gamma[n] + beta[n]);
            float g = rstd * (dz * gamma[n] - c1 - xh * c2);
            dy[n] = g;
            atomicAdd(&acc[2 * N + n], g);
        }
    }
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        if (dgamma) atomicAdd(&dgamma[n], acc[n]);
        if (dbeta) atomicAdd(&dbeta[n], acc[N + n]);
        if (dbias) atomicAdd(&dbias[n], acc[2 * N + n]);
    }
}

// ---------------------------------------------------------------------------
// Register-accumulating variant used on the hot path (N <= 32*NC):
//   * each lane owns columns lane, lane+32, ...; dgamma / dbeta / dbias partial sums stay in
//     registers across all rows of the warp, then one shared atomic per lane-column per warp and
//     one global atomic per column per CTA (the kernel above does three shared atomics per element);
//   * optionally fuses the preceding tiny product dS = dOut[R,No] . W3[No,N] of a head's final
//     Linear (No = nb_action, 1 or nb_class): W3 is staged in shared memory once per CTA and the
//     row of dOut is broadcast with shuffles, so dS never exists in memory.
// ---------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(256) ln_silu_bwd_reg_kernel(const float* __restrict__ dS, long ldds,
                                                              const float* __restrict__ dOut, int No,
                                                              const float* __restrict__ W3,
                                                              const float* __restrict__ Y, long ldy,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float* __restrict__ dY,
                                                              long lddy, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, float* __restrict__ dbias,
                                                              int R, int N) {
    extern __shared__ float sm[];
    float* acc = sm;          // [3][N]
    float* w3s = sm + 3 * N;  // [No][N] when fused
    for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) acc[i] = 0.f;
    if (W3)
        for (int i = threadIdx.x; i < No * N; i += blockDim.x) w3s[i] = __ldg(W3 + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const float invN = 1.0f / (float)N;
    float gam[NC], bet[NC], ag[NC], ab[NC], abi[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        const int n = lane + 32 * j;
        gam[j] = n < N ? gamma[n] : 0.f;
        bet[j] = n < N ? beta[n] : 0.f;
        ag[j] = ab[j] = abi[j] = 0.f;
    }
    for (int r = warp; r < R; r += nwarps) {
        float y[NC], ds[NC];
        const float* yr = Y + (long)r * ldy;
#pragma unroll
        for (int j = 0; j < NC; ++j) y[j] = (lane + 32 * j < N) ? yr[lane + 32 * j] : 0.f;
        if (W3) {
#pragma unroll
            for (int j = 0; j < NC; ++j) ds[j] = 0.f;
            const float* dor = dOut + (long)r * No;
            for (int k0 = 0; k0 < No; k0 += 32) {
                const float mine = (k0 + lane < No) ? dor[k0 + lane] : 0.f;
                const int kn = min(32, No - k0);
                for (int k = 0; k < kn; ++k) {
                    const float d = __shfl_sync(0xffffffffu, mine, k);
                    const float* wr = w3s + (k0 + k) * N + lane;
#pragma unroll
                    for (int j = 0; j < NC; ++j)
                        if (lane + 32 * j < N) ds[j] = fmaf(d, wr[32 * j], ds[j]);
                }
            }
        } else {
            const float* dsr = dS + (long)r * ldds;
#pragma unroll
            for (int j = 0; j < NC; ++j) ds[j] = (lane + 32 * j < N) ? dsr[lane + 32 * j] : 0.f;
        }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NC; ++j) s += y[j];  // padding columns hold 0
        const float mean = warp_sum(s) * invN;
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < NC; ++j)
            if (lane + 32 * j < N) { const float d = y[j] - mean; v += d * d; }
        const float rstd = 1.0f / sqrtf(warp_sum(v) * invN + LN_EPS);
        float c1 = 0.f, c2 = 0.f, xh[NC], dz[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) {
            const bool ok = lane + 32 * j < N;
            xh[j] = ok ? (y[j] - mean) * rstd : 0.f;
            dz[j] = ok ? ds[j] * silu_grad_(xh[j] * gam[j] + bet[j]) : 0.f;
            const float dxh = dz[j] * gam[j];
            c1 += dxh;
            c2 += dxh * xh[j];
            ag[j] += dz[j] * xh[j];
            ab[j] += dz[j];
        }
        c1 = warp_sum(c1) * invN;
        c2 = warp_sum(c2) * invN;
        float* dy = dY + (long)r * lddy;
#pragma unroll
        for (int j = 0; j < NC; ++j)
            if (lane + 32 * j < N) {
                const float g = rstd * (dz[j] * gam[j] - c1 - xh[j] * c2);
                dy[lane + 32 * j] = g;
                abi[j] += g;
            }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        const int n = lane + 32 * j;
        if (n < N) {
            atomicAdd(&acc[n], ag[j]);
            atomicAdd(&acc[N + n], ab[j]);
            atomicAdd(&acc[2 * N + n], abi[j]);
        }
    }
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        if (dgamma) atomicAdd(&dgamma[n], acc[n]);
        if (dbeta) atomicAdd(&dbeta[n], acc[N + n]);
        if (dbias) atomicAdd(&dbias[n], acc[2 * N + n]);
    }
}

template <int NC>
static int launch_ln_bwd_reg(const float* dS, long ldds, const float* dOut, int No, const float* W3, const float* Y,
                             long ldy, const float* gamma, const float* beta, float* dY, long lddy, float* dgamma,
                             float* dbeta, float* dbias, int R, int N, cudaStream_t s) {
    const size_t smem = sizeof(float) * ((size_t)3 * N + (W3 ? (size_t)No * N : 0));
    MARLC_CHECK(smem <= 200 * 1024, "ln_silu_bwd: %zu B of shared memory (No=%d, N=%d)", smem, No, N);
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        MARLC_CUDA(cudaFuncSetAttribute(ln_silu_bwd_reg_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    // one CTA per SM when the staged weights are large (their staging cost is per CTA), else two
    const int blocks = max(1, min((R + 7) / 8, MARLC_SMS * (W3 && No * N > 8192 ? 1 : 2)));
    ln_silu_bwd_reg_kernel<NC><<<blocks, 256, smem, s>>>(dS, ldds, dOut, No, W3, Y, ldy, gamma, beta, dY, lddy, dgamma,
                                                         dbeta, dbias, R, N);
    MARLC_LAUNCH_CHECK();
    return 0;
}

// dS given (W3 == nullptr), or fused dS = dOut[R,No] . W3[No,N]
int ln_silu_bwd_fused(const float* dS, long ldds, const float* dOut, int No, const float* W3, const float* Y, long ldy,
                      const float* gamma, const float* beta, float* dY, long lddy, float* dgamma, float* dbeta,
                      float* dbias, int R, int N, cudaStream_t s) {
    if (R <= 0) return 0;
#define MARLC_LNB(NC) \
    return launch_ln_bwd_reg<NC>(dS, ldds, dOut, No, W3, Y, ldy, gamma, beta, dY, lddy, dgamma, dbeta, dbias, R, N, s)
    if (N <= 32) MARLC_LNB(1);
    if (N <= 64) MARLC_LNB(2);
    if (N <= 128) MARLC_LNB(4);
    if (N <= 256) MARLC_LNB(8);
    if (N <= 384) MARLC_LNB(12);
    if (N <= 512) MARLC_LNB(16);
    if (N <= 1024) MARLC_LNB(32);
#undef MARLC_LNB
    MARLC_FAIL("ln_silu_bwd_fused: N=%d too wide", N);
}

int ln_silu_bwd(const float* dS, long ldds, const float* Y, long ldy, const float* gamma, const float* beta, float* dY,
                long lddy, float* dgamma, float* dbeta, float* dbias, int R, int N, cudaStream_t s) {
    if (R <= 0) return 0;
    if (N <= 1024)
        return ln_silu_bwd_fused(dS, ldds, nullptr, 0, nullptr, Y, ldy, gamma, beta, dY, lddy, dgamma, dbeta, dbias, R, N, s);
    MARLC_CHECK(3 * N * sizeof(float) <= 48 * 1024, "ln_silu_bwd: N=%d too wide", N);
    int blocks = max(1, min((R + 7) / 8, MARLC_SMS * 2));
    ln_silu_bwd_kernel<<<blocks, 256, 3 * N * sizeof(float), s>>>(dS, ldds, Y, ldy, gamma, beta, dY, lddy, dgamma,
                                                                  dbeta, dbias, R, N);
    MARLC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------
// out[n] += sum_r X[r,n]   (bias gradients)
// ---------------------------------------------------------------------------
__global__ void colsum_kernel(const float* __restrict__ X, long ldx, float* __restrict__ out, float* __restrict__ out2,
                              int R, int N, int rows_per_block) {
    __shared__ float red[8][33];
    const int n = blockIdx.x * 32 + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(R, r0 + rows_per_block);
    float s = 0.f;
    if (n < N)
        for (int r = r0 + threadIdx.y; r < r1; r += 8) s += X[(long)r * ldx + n];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        atomicAdd(&out[n], t);
        if (out2) atomicAdd(&out2[n], t);  // e.g. bias_ih and bias_hh of an LSTM cell share their gradient
    }
}

int colsum_add(const float* X, long ldx, float* out, int R, int N, cudaStream_t s) {
    return colsum_add2(X, ldx, out, nullptr, R, N, s);
}
int colsum_add2(const float* X, long ldx, float* out, float* out2, int R, int N, cudaStream_t s) {
    if (R <= 0 || N <= 0) return 0;
    int gy = max(1, min((R + 63) / 64, 128));
    int rpb = (R + gy - 1) / gy;
    dim3 grid((N + 31) / 32, gy);
    colsum_kernel<<<grid, dim3(32, 8), 0, s>>>(X, ldx, out, out2, R, N, rpb);
    MARLC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------
// Cross-agent message mean (message.py:5-17): coll[a,b,:] = (sum_a' m[a',b,:] - m[a,b,:]) / (Na-1)
// Lanes are split (agent-lane, feature-lane); partial sums over agents are
// combined with warp shuffles.  The operator is symmetric, so the same kernel is
// its own adjoint in the backward sweep.
// ---------------------------------------------------------------------------
template <int W>  // feature lanes per warp (power of two <= 32); 32/W agent lanes
__global__ void msg_mean_kernel(const float* __restrict__ msg, float* __restrict__ coll, int Na, int Nb, int n) {
    constexpr int AS = 32 / W;
    const int lane = threadIdx.x & 31, js = lane % W, as = lane / W;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int chunks = (n + W - 1) / W;
    if (warp >= Nb * chunks) return;
    const int b = warp / chunks, j = (warp % chunks) * W + js;
    const long stride = (long)Nb * n;
    const float* base = msg + (long)b * n + j;
    float part = 0.f;
    if (j < n)
        for (int a = as; a < Na; a += AS) part += base[a * stride];
#pragma unroll
    for (int o = W; o < 32; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (j < n) {
        const float inv = Na > 1 ? 1.0f / (float)(Na - 1) : 0.f;
        float* out = coll + (long)b * n + j;
        for (int a = as; a < Na; a += AS) out[a * stride] = Na > 1 ? (part - base[a * stride]) * inv : 0.f;
    }
}

int msg_mean(const float* msg, float* coll, int Na, int Nb, int n, cudaStream_t s) {
    if (Na <= 0 || Nb <= 0 || n <= 0) return 0;
    int W = 32;
    while (W > 1 && W / 2 >= n) W >>= 1;
    if (W == 32 && Na >= 2) W = 16;  // keep >= 2 agent lanes so the reduction runs across lanes
    int chunks = (n + W - 1) / W;
    int warps = Nb * chunks;
    int blocks = (warps * 32 + 255) / 256;
#define MM(Wv) msg_mean_kernel<Wv><<<blocks, 256, 0, s>>>(msg, coll, Na, Nb, n)
    switch (W) {
        case 32: MM(32); break;
        case 16: MM(16); break;
        case 8: MM(8); break;
        case 4: MM(4); break;
        case 2: MM(2); break;
        default: MM(1); break;
    }
#undef MM
    MARLC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------
// lambda = SiLU(LN(npos W^T + b))  with d = 2 inputs (state.py:7-17); warp per row
// ---------------------------------------------------------------------------
__global__ void pos_features_kernel(const float* __restrict__ npos, const float* __restrict__ W,
                                    const float* __restrict__ b, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, float* __restrict__ y_pre,
                                    float* __restrict__ out, long ldo, int R, int nd) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const float inv = 1.0f / (float)nd;
    for (int r = warp; r < R; r += nwarps) {
        const float p0 = npos[2 * r], p1 = npos[2 * r + 1];
        float s = 0.f;
        for (int j = lane; j < nd; j += 32) {
            // same association order as addmm: (x0*w0 + x1*w1) + b
            float y = fmaf(p1, W[2 * j + 1], p0 * W[2 * j]) + b[j];
            y_pre[(long)r * nd + j] = y;
            s += y;
        }
        const float mean = warp_sum(s) * inv;
        float v = 0.f;
        for (int j = lane; j < nd; j += 32) {
            float d = y_pre[(long)r * nd + j] - mean;
            v += d * d;
        }
        const float rstd = 1.0f / sqrtf(warp_sum(v) * inv + LN_EPS);
        for (int j = lane; j < nd; j += 32)
            out[(long)r * ldo + j] = siluf_((y_pre[(long)r * nd + j] - mean) * rstd * gamma[j] + beta[j]);
    }
}

int pos_features_fwd(const float* npos, const float* W, const float* b, const float* gamma, const float* beta,
                     float* y_pre, float* out, long ldo, int R, int nd, cudaStream_t s) {
    if (R <= 0) return 0;
    pos_features_kernel<<<row_grid(R, 8), 256, 0, s>>>(npos, W, b, gamma, beta, y_pre, out, ldo, R, nd);
    MARLC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------
// LSTM cell point-wise part (recurrent.py:30 -> nn.LSTMCell; gate order i,f,g,o)
// gates holds pre-activations on entry, activations (kept for backward) on exit.
// ---------------------------------------------------------------------------
__global__ void lstm_cell_fwd_kernel(float* __restrict__ gates, const float* __restrict__ c_prev,
                                     float* __restrict__ c_new, float* __restrict__ h_new, int M, int n) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)M * n) return;
    const int m = idx / n, j = idx % n;
    float* g = gates + (long)m * 4 * n;
    const float i = sigmoidf_(g[j]), f = sigmoidf_(g[n + j]), gg = tanhf(g[2 * n + j]), o = sigmoidf_(g[3 * n + j]);
    const float c = f * c_prev[idx] + i * gg;
    g[j] = i; g[n + j] = f; g[2 * n + j] = gg; g[3 * n + j] = o;
    c_new[idx] = c;
    h_new[idx] = o * tanhf(c);
}

int lstm_cell_fwd(float* gates, const float* c_prev, float* c_new, float* h_new, int M, int n, cudaStream_t s) {
    long tot = (long)M * n;
    if (tot <= 0) return 0;
    lstm_cell_fwd_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(gates, c_prev, c_new, h_new, M, n);
    MARLC_LAUNCH_CHECK();
    return 0;
}

__global__ void lstm_cell_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ dh2,
                                     const float* __restrict__ dc_next, const float* __restrict__ gates,
                                     const float* __restrict__ c_prev, const float* __restrict__ c_new,
                                     float* __restrict__ dgates, float* __restrict__ dc_prev, int M, int n) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)M * n) return;
    const int m = idx / n, j = idx % n;
    const float* g = gates + (long)m * 4 * n;
    const float i = g[j], f = g[n + j], gg = g[2 * n + j], o = g[3 * n + j];
    float dhv = dh ? dh[idx] : 0.f;
    if (dh2) dhv += dh2[idx];
    const float tc = tanhf(c_new[idx]);
    const float dc = (dc_next ? dc_next[idx] : 0.f) + dhv * o * (1.f - tc * tc);
    float* dg = dgates + (long)m * 4 * n;
    dg[j] = dc * gg * i * (1.f - i);
    dg[n + j] = dc * c_prev[idx] * f * (1.f - f);
    dg[2 * n + j] = dc * i * (1.f - gg * gg);
    dg[3 * n + j] = dhv * tc * o * (1.f - o);
    dc_prev[idx] = dc * f;
}

int lstm_cell_bwd(const float* dh, const float* dh2, const float* dc_next, const float* gates, const float* c_prev,
                  const float* c_new, float* dgates, float* dc_prev, int M, int n, cudaStream_t s) {
    long tot = (long)M * n;
    if (tot <= 0) return 0;
    lstm_cell_bwd_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(dh, dh2, dc_next, gates, c_prev, c_new, dgates,
                                                                       dc_prev, M, n);
    MARLC_LAUNCH_CHECK();
    return 0;
}

__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}
int add_inplace(float* dst, const float* src, long n, cudaStream_t s) {
    if (n <= 0) return 0;
    add_inplace_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dst, src, n);
    MARLC_LAUNCH_CHECK();
    return 0;
}

}  // namespace marlc
