// Generic fp32 GEMM on CUDA cores (FFMA), arbitrary strides, grouped + split-K.
//
// Role: the exact-fp32 path.  It serves (a) ragged shapes the TMA/tcgen05 path
// cannot address (row strides that are not multiples of 16 B, e.g. the
// reference's own test sizes 23/22/21...), (b) tiny-N products (policy nA=4,
// critic N=1) and (c) the strict-fp32 parity mode.  The tensor-core path
// (gemm_tc.cu) replaces it for the LSTM gate / head / dX / dW GEMMs when shapes
// allow.
#include "common.cuh"

namespace marlc {

constexpr int BM = 64, BN = 64, BK = 16;

struct GemmLaunch {
    GemmProblem p[4];
    int count;
    int splits;  // split-K factor (atomicAdd epilogue when > 1)
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmLaunch g) {
    const int pi = blockIdx.z / g.splits, sp = blockIdx.z % g.splits;
    const GemmProblem& P = g.p[pi];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m0 >= P.M || n0 >= P.N) return;

    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int pass = 0; pass < 2; ++pass) {
        const float* A = pass ? P.A2 : P.A;
        const float* B = pass ? P.B2 : P.B;
        const int K = pass ? P.K2 : P.K;
        if (A == nullptr || K <= 0) continue;
        const long sam = pass ? P.sam2 : P.sam, sak = pass ? P.sak2 : P.sak;
        const long sbk = pass ? P.sbk2 : P.sbk, sbn = pass ? P.sbn2 : P.sbn;
        // split-K range (multiple of BK)
        int kb = 0, ke = K;
        if (g.splits > 1) {
            int chunk = ((K + g.splits - 1) / g.splits + BK - 1) / BK * BK;
            kb = sp * chunk;
            ke = min(K, kb + chunk);
        }
        for (int k0 = kb; k0 < ke; k0 += BK) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int e = tid + i * 256;
                int m, k;
                if (sak == 1) { m = e / BK; k = e % BK; } else { k = e / BM; m = e % BM; }
                float v = 0.f;
                if (m0 + m < P.M && k0 + k < ke) v = A[(long)(m0 + m) * sam + (long)(k0 + k) * sak];
                As[k][m] = v;
                int n;
                if (sbk == 1) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
                v = 0.f;
                if (n0 + n < P.N && k0 + k < ke) v = B[(long)(k0 + k) * sbk + (long)(n0 + n) * sbn];
                Bs[k][n] = v;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= P.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= P.N) continue;
            float v = acc[i][j];
            if (sp == 0) {
                if (P.bias) v += P.bias[n];
                if (P.bias2) v += P.bias2[n];
            }
            float* c = P.C + (long)m * P.ldc + n;
            if (g.splits > 1) atomicAdd(c, v);
            else if (P.accumulate) *c += v;
            else *c = v;
        }
    }
}

static int launch(const GemmLaunch& g, cudaStream_t s) {
    int maxM = 0, maxN = 0;
    for (int i = 0; i < g.count; ++i) {
        maxM = max(maxM, g.p[i].M);
        maxN = max(maxN, g.p[i].N);
    }
    if (maxM <= 0 || maxN <= 0) return 0;
    dim3 grid((maxN + BN - 1) / BN, (maxM + BM - 1) / BM, g.count * g.splits);
    gemm_simt_kernel<<<grid, 256, 0, s>>>(g);
    ++g_simt_gemm_launches;
    MARLC_LAUNCH_CHECK();
    return 0;
}

int gemm_group(const GemmGroup& grp, cudaStream_t s) {
    GemmLaunch g;
    g.count = grp.count;
    g.splits = 1;
    for (int i = 0; i < grp.count; ++i) g.p[i] = grp.p[i];
    return launch(g, s);
}

static GemmProblem blank() {
    GemmProblem p;
    p.A = p.B = p.A2 = p.B2 = p.bias = p.bias2 = nullptr;
    p.sam = p.sak = p.sbk = p.sbn = p.sam2 = p.sak2 = p.sbk2 = p.sbn2 = 0;
    p.C = nullptr;
    p.ldc = 0;
    p.M = p.N = p.K = p.K2 = 0;
    p.accumulate = 0;
    return p;
}

int gemm_nt(const float* X, long ldx, const float* W, long ldw, const float* bias, float* Y, long ldy, int M, int N,
            int K, int accumulate, cudaStream_t s) {
    if (N == 1 && M >= 256) return rowdot(X, ldx, W, bias, Y, ldy, M, K, accumulate, s);  // critic head: a row-wise dot
    GemmLaunch g;
    g.count = 1;
    g.splits = 1;
    GemmProblem p = blank();
    p.A = X; p.sam = ldx; p.sak = 1;
    p.B = W; p.sbk = 1; p.sbn = ldw;
    p.bias = bias;
    p.C = Y; p.ldc = ldy;
    p.M = M; p.N = N; p.K = K;
    p.accumulate = accumulate;
    g.p[0] = p;
    return launch(g, s);
}

int gemm_nn(const float* dY, long lddy, const float* W, long ldw, float* dX, long lddx, int M, int N, int K,
            int accumulate, cudaStream_t s) {
    // dX[m,k] = sum_n dY[m,n] W[n,k]: reduction dim is N, output width is K
    GemmLaunch g;
    g.count = 1;
    g.splits = 1;
    GemmProblem p = blank();
    p.A = dY; p.sam = lddy; p.sak = 1;
    p.B = W; p.sbk = ldw; p.sbn = 1;
    p.C = dX; p.ldc = lddx;
    p.M = M; p.N = K; p.K = N;
    p.accumulate = accumulate;
    g.p[0] = p;
    return launch(g, s);
}

int gemm_tn(const float* dY, long lddy, const float* X, long ldx, float* dW, long lddw, int R, int N, int K,
            int accumulate, cudaStream_t s) {
    // dW[n,k] = sum_r dY[r,n] X[r,k]: reduction over rows R (split-K with atomics).
    {   // a handful of outputs per row: one coalesced pass over the tall operand (skinny.cu)
        bool done = false;
        MARLC_TRY(tn_skinny(dY, lddy, X, ldx, dW, lddw, R, N, K, accumulate, s, &done));
        if (done) return 0;
    }
    GemmLaunch g;
    g.count = 1;
    GemmProblem p = blank();
    p.A = dY; p.sam = 1; p.sak = lddy;
    p.B = X; p.sbk = ldx; p.sbn = 1;
    p.C = dW; p.ldc = lddw;
    p.M = N; p.N = K; p.K = R;
    p.accumulate = accumulate;
    int tiles = ((N + BM - 1) / BM) * ((K + BN - 1) / BN);
    int splits = (2 * MARLC_SMS + tiles - 1) / tiles;
    int max_splits = (R + 8 * BK - 1) / (8 * BK);
    splits = max(1, min(splits, max_splits));
    g.splits = splits;
    if (splits > 1 && !accumulate) {
        // atomics need a zeroed destination
        if (lddw == K) {
            MARLC_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)N * K, s));
        } else {
            MARLC_CUDA(cudaMemset2DAsync(dW, sizeof(float) * lddw, 0, sizeof(float) * K, N, s));
        }
    }
    g.p[0] = p;
    return launch(g, s);
}

}  // namespace marlc
