// Environment kernels: patch gather, integer position transition, normalised
// positions, episode initial state, and the fused policy-head -> sample ->
// log-prob -> transition step (reference: core/environment.py, core/agent.py).
#include "common.cuh"
#include "kernels.cuh"

namespace marlc {

// ---------------------------------------------------------------------------
// K1 patch gather (environment.py:96-126):
//   obs[a,b,c,i,j] = img[b,c,pos[a,b,0]+i,pos[a,b,1]+j]
// One CTA per (a,b) window.  Source rows are read as ALIGNED 128-bit vectors
// (aligned-down start, covering the f-float segment) into shared memory, then
// the contiguous C*f*f output is written as 128-bit vectors.  Falls back to
// scalar accesses when W or the patch size is not a multiple of 4.
// ---------------------------------------------------------------------------
template <typename PosT>
__global__ void __launch_bounds__(128)
patch_gather_kernel(const float* __restrict__ img, const PosT* __restrict__ pos, float* __restrict__ obs, int B, int C,
                    int H, int W, int f, int vec_ok) {
    extern __shared__ __align__(16) float stage[];  // [C*f][pitch]
    const int m = blockIdx.x;                        // a*B + b
    const int b = m % B;
    const int py = (int)pos[2 * (long)m], px = (int)pos[2 * (long)m + 1];
    const int rows = C * f;
    const float* src = img + (long)b * C * H * W;
    float* dst = obs + (long)m * rows * f;
    if (vec_ok) {
        const int x0 = px & ~3;
        const int nv = (px + f - x0 + 3) >> 2;  // float4 per row segment
        const int pitch = ((f + 3) & ~3) + 8;   // floats, multiple of 4
        for (int e = threadIdx.x; e < rows * nv; e += blockDim.x) {
            const int r = e / nv, v = e % nv;
            const int c = r / f, i = r % f;
            const int x = x0 + 4 * v;
            const float* p = src + ((long)c * H + (py + i)) * W + x;
            float4 val;
            if (x + 3 < W) val = __ldg(reinterpret_cast<const float4*>(p));
            else {  // last vector of an image row may overhang (never dereference past the row)
                val.x = x < W ? p[0] : 0.f; val.y = x + 1 < W ? p[1] : 0.f;
                val.z = x + 2 < W ? p[2] : 0.f; val.w = 0.f;
            }
            *reinterpret_cast<float4*>(&stage[r * pitch + 4 * v]) = val;
        }
        __syncthreads();
        const int off = px - x0;
        const int total4 = (rows * f) >> 2;
        for (int e = threadIdx.x; e < total4; e += blockDim.x) {
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = 4 * e + q, r = idx / f, j = idx % f;
                o[q] = stage[r * pitch + off + j];
            }
            reinterpret_cast<float4*>(dst)[e] = make_float4(o[0], o[1], o[2], o[3]);
        }
    } else {
        for (int e = threadIdx.x; e < rows * f; e += blockDim.x) {
            const int r = e / f, j = e % f, c = r / f, i = r % f;
            dst[e] = __ldg(src + ((long)c * H + (py + i)) * W + px + j);
        }
    }
}

template <typename PosT>
static int patch_gather_t(const float* img, const PosT* pos, float* obs, int Na, int B, int C, int H, int W, int f,
                          cudaStream_t s) {
    MARLC_CHECK(f >= 1 && f <= H && f <= W, "patch_gather: window f=%d does not fit %dx%d", f, H, W);
    const int M = Na * B;
    if (M <= 0) return 0;
    int vec_ok = (W % 4 == 0) && ((C * f * f) % 4 == 0) && (((uintptr_t)img & 15) == 0) && (((uintptr_t)obs & 15) == 0);
    const int pitch = ((f + 3) & ~3) + 8;
    size_t smem = vec_ok ? sizeof(float) * (size_t)C * f * pitch : 0;
    if (smem > 200 * 1024) { vec_ok = 0; smem = 0; }
    if (smem > 48 * 1024)
        MARLC_CUDA(cudaFuncSetAttribute(patch_gather_kernel<PosT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    patch_gather_kernel<PosT><<<M, 128, smem, s>>>(img, pos, obs, B, C, H, W, f, vec_ok);
    MARLC_LAUNCH_CHECK();
    return 0;
}
int patch_gather_i64(const float* img, const int64_t* pos, float* obs, int Na, int B, int C, int H, int W, int f,
                     cudaStream_t s) {
    return patch_gather_t<int64_t>(img, pos, obs, Na, B, C, H, W, f, s);
}
int patch_gather_i32(const float* img, const int* pos, float* obs, int Na, int B, int C, int H, int W, int f,
                     cudaStream_t s) {
    return patch_gather_t<int>(img, pos, obs, Na, B, C, H, W, f, s);
}

// ---------------------------------------------------------------------------
// K2 transition (environment.py:56-66, 128-150) in pure integer arithmetic:
// the move is applied only if every dim stays inside (0 <= p+m, p+m+f < S).
// Also emits normalised positions float(p)/float(S) (environment.py:74-81).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void apply_move(int& py, int& px, int my, int mx, int f, int H, int W) {
    const int ny = py + my, nx = px + mx;
    const bool ok = (ny >= 0) & (ny + f < H) & (nx >= 0) & (nx + f < W);
    if (ok) { py = ny; px = nx; }
}

__global__ void transition_i64_kernel(int64_t* __restrict__ pos, const int64_t* __restrict__ act,
                                      const int64_t* __restrict__ table, int nA, int M, int f, int H, int W,
                                      float* __restrict__ npos, int* __restrict__ err) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    int py = (int)pos[2 * m], px = (int)pos[2 * m + 1];
    const int64_t a = act[m];
    if (a < 0 || a >= nA) { if (err) atomicExch(err, 1); }
    else apply_move(py, px, (int)table[2 * a], (int)table[2 * a + 1], f, H, W);
    pos[2 * m] = py; pos[2 * m + 1] = px;
    if (npos) { npos[2 * m] = (float)py / (float)H; npos[2 * m + 1] = (float)px / (float)W; }
}

int transition_i64(int64_t* pos, const int64_t* act, const int64_t* table, int nA, int M, int f, int H, int W,
                   float* npos, int* err, cudaStream_t s) {
    if (M <= 0) return 0;
    transition_i64_kernel<<<(M + 127) / 128, 128, 0, s>>>(pos, act, table, nA, M, f, H, W, npos, err);
    MARLC_LAUNCH_CHECK();
    return 0;
}

__global__ void normalized_positions_kernel(const int64_t* __restrict__ pos, float* __restrict__ out, int M, int H, int W) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    out[2 * m] = (float)pos[2 * m] / (float)H;
    out[2 * m + 1] = (float)pos[2 * m + 1] / (float)W;
}
int normalized_positions_i64(const int64_t* pos, float* out, int M, int H, int W, cudaStream_t s) {
    if (M <= 0) return 0;
    normalized_positions_kernel<<<(M + 127) / 128, 128, 0, s>>>(pos, out, M, H, W);
    MARLC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------
// Episode initial state.  Either copies injected values or draws them:
// positions ~ U{0..S_d-f-1} (environment.py:33-43), recurrent state ~ N(0,1)
// (models.py:148-159).  Random streams are Philox(seed) with the episode
// counter read from device memory so CUDA-graph replays draw fresh numbers.
// ---------------------------------------------------------------------------
__global__ void episode_init_kernel(EpisodeInitArgs a) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t episode = a.rng_state ? a.rng_state[1] : 0;
    const Philox ph(a.rng_state ? a.rng_state[0] : 0);
    if (idx < a.M) {
        int py, px;
        if (a.pos0) { py = (int)a.pos0[2 * idx]; px = (int)a.pos0[2 * idx + 1]; }
        else {
            uint4 r = ph((uint64_t)idx, (episode << 8) | 1);
            py = (int)(((uint64_t)r.x * (uint64_t)(a.H - a.f)) >> 32);
            px = (int)(((uint64_t)r.y * (uint64_t)(a.W - a.f)) >> 32);
        }
        a.pos[2 * idx] = py; a.pos[2 * idx + 1] = px;
        a.npos[2 * idx] = (float)py / (float)a.H;
        a.npos[2 * idx + 1] = (float)px / (float)a.W;
    }
    // hidden states: 4 tensors, element idx over max size
    for (int k = 0; k < 4; ++k) {
        const long n = (long)a.M * a.width[k];
        if (idx >= n) continue;
        float v;
        if (a.hidden0[k]) v = a.hidden0[k][idx];
        else {
            uint4 r = ph((uint64_t)idx, (episode << 8) | (2 + k));
            // Box-Muller
            float u1 = fmaxf(u01(r.x), 1.0f / 16777216.0f), u2 = u01(r.y);
            v = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
        }
        a.hidden[k][idx] = v;
    }
    for (long i = idx; i < (long)a.M * a.n_m; i += (long)gridDim.x * blockDim.x) a.msg0[i] = 0.f;
}

int episode_init(const EpisodeInitArgs& a, cudaStream_t s) {
    long n = a.M;
    for (int k = 0; k < 4; ++k) n = max(n, (long)a.M * a.width[k]);
    episode_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
    MARLC_LAUNCH_CHECK();
    return 0;
}

__global__ void rng_advance_kernel(uint64_t* rng_state) { rng_state[1] += 1; }
int rng_advance(uint64_t* rng_state, cudaStream_t s) {
    rng_advance_kernel<<<1, 1, 0, s>>>(rng_state);
    MARLC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------
// K7 policy tail (policy.py:15-16 + agent.py:51-61 + environment.py:56-66), one
// warp per row: logits = s1 W3^T + b3 -> softmax -> sample (or injected action)
// -> log p[a] -> integer transition -> next normalised position.
// ---------------------------------------------------------------------------
constexpr int MAX_ACTIONS = 16;

__global__ void policy_act_kernel(PolicyActArgs a) {
    const int lane = threadIdx.x & 31;
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= a.M) return;
    const float* x = a.s1 + (long)m * a.nl;
    float logit[MAX_ACTIONS];
#pragma unroll 1
    for (int j = 0; j < a.nA; ++j) {
        const float* w = a.W3 + (long)j * a.nl;
        float s = 0.f;
        for (int k = lane; k < a.nl; k += 32) s = fmaf(x[k], w[k], s);
        logit[j] = warp_sum(s) + a.b3[j];
    }
    float mx = -INFINITY;
    for (int j = 0; j < a.nA; ++j) mx = fmaxf(mx, logit[j]);
    float den = 0.f;
    for (int j = 0; j < a.nA; ++j) { logit[j] = expf(logit[j] - mx); den += logit[j]; }
    const float inv = 1.0f / den;
    int act;
    if (a.act_in) act = (int)a.act_in[m];
    else {
        const uint64_t episode = a.rng_state[1];
        const Philox ph(a.rng_state[0]);
        uint4 r = ph((uint64_t)a.t * (uint64_t)a.M + (uint64_t)m, (episode << 8) | 7);
        const float u = u01(r.x);
        float cdf = 0.f;
        act = a.nA - 1;
        for (int j = 0; j < a.nA; ++j) {
            cdf += logit[j] * inv;
            if (u < cdf) { act = j; break; }
        }
    }
    if (lane == 0) {
        float pa = 0.f;
        for (int j = 0; j < a.nA; ++j) {
            const float p = logit[j] * inv;
            a.probs[(long)m * a.nA + j] = p;
            if (j == act) pa = p;
        }
        a.logp[m] = logf(pa);
        a.act_out[m] = act;
        int py = a.pos_in[2 * m], px = a.pos_in[2 * m + 1];
        if (act >= 0 && act < a.nA) apply_move(py, px, a.moves[2 * act], a.moves[2 * act + 1], a.f, a.H, a.W);
        a.pos_out[2 * m] = py; a.pos_out[2 * m + 1] = px;
        a.step_pos[2 * (long)m] = py; a.step_pos[2 * (long)m + 1] = px;
        a.npos_out[2 * m] = (float)py / (float)a.H;
        a.npos_out[2 * m + 1] = (float)px / (float)a.W;
    }
}

int policy_act(const PolicyActArgs& a, cudaStream_t s) {
    MARLC_CHECK(a.nA <= MAX_ACTIONS, "policy_act: nb_action=%d > %d", a.nA, MAX_ACTIONS);
    if (a.M <= 0) return 0;
    policy_act_kernel<<<(a.M * 32 + 127) / 128, 128, 0, s>>>(a);
    MARLC_LAUNCH_CHECK();
    return 0;
}

}  // namespace marlc
