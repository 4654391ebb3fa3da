// Per-warp action-value selection of one (b, t) sample, shared by the fused mixer kernels.
// algorithm/q_learner.py:100-117 -- the same arithmetic as q_select_kernel (select_td.cu), arranged for a warp that owns
// the sample: the [N, A] slabs are read coalesced into shared memory (availability mask folded in, q_targets masked in
// place like the reference does), then lane = agent scans its row.
#pragma once
#include "common.cuh"
#include "../../include/marl_b200.h"

namespace marl {

constexpr float kNegBig = -9999999.0f;   // algorithm/q_learner.py:105,112,126

struct SelectArgs {
    const float* q; const float* qn; float* qt; const float* avail_next; const long long* u;
    float* q_out; float* qt_out; long long* a_star;
};

inline SelectArgs select_args(const marl_select_fused* s, const long long* u, float* q_chosen, float* q_tc) {
    return SelectArgs{s->q_evals, s->q_evals_next, s->q_targets, s->avail_u_next, u, q_chosen, q_tc, s->a_star};
}
inline bool select_ok(const marl_select_fused* s) { return s->q_evals && s->q_targets && s->avail_u_next; }
// dynamic shared memory the staging needs for `warps` warps per CTA
inline size_t select_smem(int warps, int N, int A) { return (size_t)warps * 2 * N * A * sizeof(float); }
constexpr size_t kSelectSmemMax = 160 * 1024;

// stage: 2 * N * A floats of this warp; qc / tc: N floats each of this warp (q_chosen, q_targets_chosen on return)
__device__ __forceinline__ void warp_select(const SelectArgs& s, long long m, int N, int A, int lane, float* stage,
                                            float* qc, float* tc) {
    const int NA = N * A;
    const long long o = m * NA;
    float* sn = stage;
    float* st = stage + NA;
    for (int i = lane; i < NA; i += 32) {
        const bool off = s.avail_next[o + i] == 0.0f;
        float v = s.qt[o + i];
        if (off) { v = kNegBig; s.qt[o + i] = v; }                     // in place, q_learner.py:105
        st[i] = v;
        if (s.qn) sn[i] = off ? kNegBig : s.qn[o + i];                 // :112
    }
    __syncwarp();
    for (int n = lane; n < N; n += 32) {
        const long long i = m * N + n;
        const float c = s.q[i * A + s.u[i]];                           // :100
        int best = 0;
        float tmax = 0.f, tsel = 0.f;
        if (s.qn) {
            float bv = 0.f;
            for (int a = 0; a < A; ++a) {
                const float v = sn[n * A + a];
                if (a == 0 || v > bv) { bv = v; best = a; }            // first maximum wins (th.argmax on CPU)
            }
        }
        for (int a = 0; a < A; ++a) {
            const float v = st[n * A + a];
            if (a == 0 || v > tmax) tmax = v;
            if (a == best) tsel = v;
        }
        const float t = s.qn ? tsel : tmax;                            // :114 / :117
        qc[n] = c; tc[n] = t;
        s.q_out[i] = c; s.qt_out[i] = t;
        if (s.a_star) s.a_star[i] = s.qn ? best : -1;
    }
    __syncwarp();
}

// dL/dh through the agents' head for a sample whose dL/dq has one non-zero per agent row (g[n] at action u[n]):
// dq . W2 is that row of W2 scaled -- what the [M*N, H, A] dgrad GEMM would compute
__device__ __forceinline__ void warp_dhext(const float* __restrict__ fc2_w, const long long* __restrict__ u, long long m,
                                           int N, int lane, const float* g, float* __restrict__ dhext) {
#pragma unroll 4
    for (int n = 0; n < N; ++n) {
        const float gn = g[n];
        const float2 w = __ldg((const float2*)(fc2_w + __ldg(u + m * N + n) * MARL_H) + lane);
        ((float2*)(dhext + (m * N + n) * MARL_H))[lane] = make_float2(gn * w.x, gn * w.y);
    }
}

}  // namespace marl
