// Environment kernels: patch gather, integer position transition, normalised
// positions, episode initial state, and the fused policy-head -> sample ->
// log-prob -> transition step (reference: core/environment.py, core/agent.py).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

namespace marlc {

// ---------------------------------------------------------------------------
// K1 patch gather (environment.py:96-126):
//   obs[a,b,c,i,j] = img[b,c,pos[a,b,0]+i,pos[a,b,1]+j]
// (The first version staged aligned 128-bit row reads through shared memory, one CTA per window:
// 2.4 TB/s at 65 536 windows.  Two block barriers and the index math per window dominated.)
// ---------------------------------------------------------------------------
// One WARP per window, no shared memory, no block barrier: lane-strided over the contiguous
// C*f*f output (perfectly coalesced 128 B stores); each warp-load covers 32 consecutive output
// elements = 32/f source row segments (coalesced within a segment).  All iterations' loads are
// issued before the first store (UNROLL independent requests per lane) so a window costs about
// one memory round trip.  Measured 2.4 -> see profiles/README.md for the achieved GB/s.
// The three divisions per element (e -> c, i, j) were ~60 of the ~75 instructions per element as
// runtime `/`; FastDiv (common.cuh) + 32-bit offsets + a branch-free tail: 88 -> 45 us at 65 536 windows.
template <typename PosT, int UNROLL, bool FAST>
__global__ void __launch_bounds__(256)
patch_gather_kernel(const float* __restrict__ img, const PosT* __restrict__ pos, float* __restrict__ obs, int Na, int B,
                    int C, int H, int W, int f, FastDiv dff, FastDiv df) {
    const int lane = threadIdx.x & 31;
    // consecutive warps take the agents of ONE image (b-major): an image is then read in a burst
    // while its DRAM pages are open / its sectors are in L2, instead of once per agent row
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= Na * B) return;
    const int b = g / Na, m = (g - b * Na) * B + b;  // output row m = a*B + b
    const int py = (int)pos[2 * (long)m], px = (int)pos[2 * (long)m + 1];
    const int ff = f * f, total = C * ff;
    const int plane = H * W;  // offsets inside one image fit 32 bits (host-checked)
    const float* src = img + (long)b * C * plane + (py * W + px);
    float* dst = obs + (long)m * total;
    for (int e0 = lane; e0 < total; e0 += 32 * UNROLL) {
        float v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            // tail lanes re-read the last element instead of branching around the load
            const int e = min(e0 + 32 * u, total - 1);
            const int c = FAST ? dff.div_nz(e) : e / ff;
            const int r = e - c * ff;
            const int i = FAST ? df.div_nz(r) : r / f;
            v[u] = __ldg(src + (c * plane + i * W + (r - i * f)));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            if (e0 + 32 * u < total) dst[e0 + 32 * u] = v[u];
    }
}

// MARLC_GATHER_DIV=1 selects the runtime-division body (kept for A/B timing, scripts/gather_ab.py).
static bool gather_use_div() {
    const char* e = getenv("MARLC_GATHER_DIV");
    return e && e[0] == '1';
}

template <typename PosT>
static int patch_gather_t(const float* img, const PosT* pos, float* obs, int Na, int B, int C, int H, int W, int f,
                          cudaStream_t s) {
    MARLC_CHECK(f >= 1 && f <= H && f <= W, "patch_gather: window f=%d does not fit %dx%d", f, H, W);
    const int M = Na * B;
    if (M <= 0) return 0;
    const int blocks = (M + 7) / 8;  // 8 warps = 8 windows per CTA
    const long ff = (long)f * f, total = (long)C * ff;
    MARLC_CHECK((long)C * H * W < 0x7fffffffl, "patch_gather: one image must hold fewer than 2^31 elements");
    const bool fast = f >= 2 && FastDiv::exact_up_to(total, ff) && !gather_use_div();
    const FastDiv dff((unsigned)max(ff, 2l)), df((unsigned)max(f, 2));
#define MARLC_GATHER_LAUNCH(U, F) \
    patch_gather_kernel<PosT, U, F><<<blocks, 256, 0, s>>>(img, pos, obs, Na, B, C, H, W, f, dff, df)
    if (total <= 32 * 8) { if (fast) MARLC_GATHER_LAUNCH(8, true); else MARLC_GATHER_LAUNCH(8, false); }
    else { if (fast) MARLC_GATHER_LAUNCH(16, true); else MARLC_GATHER_LAUNCH(16, false); }
#undef MARLC_GATHER_LAUNCH
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
    // one 16-byte load / store per agent for the (y, x) pair, 8-byte for the action and the float2
    longlong2 p = reinterpret_cast<const longlong2*>(pos)[m];
    int py = (int)p.x, px = (int)p.y;
    const int64_t a = act[m];
    if (a < 0 || a >= nA) { if (err) atomicExch(err, 1); }
    else apply_move(py, px, (int)__ldg(table + 2 * a), (int)__ldg(table + 2 * a + 1), f, H, W);
    p.x = py; p.y = px;
    reinterpret_cast<longlong2*>(pos)[m] = p;
    if (npos) reinterpret_cast<float2*>(npos)[m] = make_float2((float)py / (float)H, (float)px / (float)W);
}

int transition_i64(int64_t* pos, const int64_t* act, const int64_t* table, int nA, int M, int f, int H, int W,
                   float* npos, int* err, cudaStream_t s) {
    if (M <= 0) return 0;
    MARLC_CHECK(((uintptr_t)pos & 15) == 0 && ((uintptr_t)npos & 7) == 0, "transition: positions must be 16-byte aligned");
    transition_i64_kernel<<<(M + 255) / 256, 256, 0, s>>>(pos, act, table, nA, M, f, H, W, npos, err);
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
