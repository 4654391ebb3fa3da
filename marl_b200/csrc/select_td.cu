// Action-value selection (gather / masked argmax / target gather), TD loss, VDN, batch ingest.
#include "common.cuh"
#include "../../include/marl_b200.h"
#include "profile.h"
#include "select.cuh"

namespace marl {

__global__ void __launch_bounds__(256) q_select_kernel(
    int rows, int A, const float* __restrict__ q_evals, const long long* __restrict__ u,
    const float* __restrict__ q_evals_next, float* __restrict__ q_targets,
    const float* __restrict__ avail_next, const float* __restrict__ avail,
    float* __restrict__ q_chosen, long long* __restrict__ a_star, float* __restrict__ q_tc,
    float* __restrict__ max_q_evals, float* __restrict__ q_targets_max, float* __restrict__ a_star_onehot) {
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const long long o = (long long)i * A;
    if (q_chosen) q_chosen[i] = q_evals[o + u[i]];                     // q_learner.py:100
    int best = 0;
    float tmax = 0.f, tsel = 0.f;
    if (q_evals_next) {                                                // double-Q, q_learner.py:110-114
        float bv = 0.f;
        for (int a = 0; a < A; ++a) {
            float v = (avail_next[o + a] == 0.0f) ? kNegBig : q_evals_next[o + a];
            if (a == 0 || v > bv) { bv = v; best = a; }                // first maximum wins (th.argmax on CPU)
        }
    }
    for (int a = 0; a < A; ++a) {
        float v = q_targets[o + a];
        if (avail_next[o + a] == 0.0f) { v = kNegBig; q_targets[o + a] = v; }   // in place, q_learner.py:105
        if (a == 0 || v > tmax) tmax = v;
        if (a == best) tsel = v;
    }
    q_tc[i] = q_evals_next ? tsel : tmax;                              // q_learner.py:114 / :117
    if (a_star) a_star[i] = q_evals_next ? best : -1;
    if (a_star_onehot)                                                 // q_learner.py:140-143
        for (int a = 0; a < A; ++a) a_star_onehot[o + a] = (q_evals_next && a == best) ? 1.0f : 0.0f;
    if (q_targets_max) q_targets_max[i] = tmax;                        // q_learner.py:150
    if (max_q_evals) {                                                 // q_learner.py:125-127
        float m = 0.f;
        for (int a = 0; a < A; ++a) {
            float v = (avail[o + a] == 0.0f) ? kNegBig : q_evals[o + a];
            if (a == 0 || v > m) m = v;
        }
        max_q_evals[i] = m;
    }
}

// Epsilon-greedy action of every (env, agent) row: controller/share_params.py:62-72.  Unavailable actions are masked
// to -inf and the first maximum wins (th.argmax); rows whose host-drawn `explore` flag is set take the host-drawn
// random available action instead (the reference draws np.random.uniform() / np.random.choice per agent; drawing them
// on the host in the reference's order keeps the action stream bit-exact under a seed).
__global__ void __launch_bounds__(256) epsgreedy_kernel(int rows, int A, const float* __restrict__ q,
                                                        const float* __restrict__ avail, const unsigned char* __restrict__ explore,
                                                        const long long* __restrict__ random_action,
                                                        long long* __restrict__ action, float* __restrict__ onehot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const long long o = (long long)i * A;
    int best = 0;
    float bv = -INFINITY;
    for (int a = 0; a < A; ++a) {
        const float v = (avail && avail[o + a] == 0.0f) ? -INFINITY : q[o + a];
        if (v > bv) { bv = v; best = a; }                     // strict: first maximum; all masked -> 0 like th.argmax
    }
    if (explore && explore[i]) best = (int)random_action[i];
    action[i] = best;
    if (onehot)
        for (int a = 0; a < A; ++a) onehot[o + a] = (a == best) ? 1.0f : 0.0f;
}

// Block-wide sum of two values; thread 0 of the block adds them to out[0], out[1].
__device__ __forceinline__ void block_accumulate2(float a, float b, float* out) {
    __shared__ float sa[32], sb[32];
    a = warp_sum(a); b = warp_sum(b);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (l == 0) { sa[w] = a; sb[w] = b; }
    __syncthreads();
    if (w == 0) {
        a = l < nw ? sa[l] : 0.f; b = l < nw ? sb[l] : 0.f;
        a = warp_sum(a); b = warp_sum(b);
        if (l == 0) { atomicAdd(out, a); atomicAdd(out + 1, b); }
    }
}

__device__ __forceinline__ float td_grad(float q_tot, float q_tot_t, float r, float term, float padded, float gamma,
                                         float& sq, float& msk) {
    const float mask = 1.0f - padded;                                  // q_learner.py:83
    const float y = r + gamma * q_tot_t * (1.0f - term);               // :165
    const float d = mask * (y - q_tot);                                // :166-167
    sq = d * d; msk = mask;
    return -2.0f * mask * d;                                           // d(sum d^2)/d q_tot
}

__global__ void __launch_bounds__(256) td_loss_kernel(int M, const float* q_tot, const float* q_tot_t, const float* r,
                                                      const float* term, const float* padded, float gamma,
                                                      float* dq_tot, float* scalars) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    float sq = 0.f, msk = 0.f;
    if (m < M) {
        float g = td_grad(q_tot[m], q_tot_t[m], r[m], term[m], padded[m], gamma, sq, msk);
        if (dq_tot) dq_tot[m] = g;
    }
    block_accumulate2(sq, msk, scalars);
}

__global__ void __launch_bounds__(256) vdn_td_kernel(int M, int N, int A, const float* q_chosen, const float* q_tc,
                                                     const long long* u, const float* r, const float* term,
                                                     const float* padded, float gamma, float* q_tot, float* q_tot_t,
                                                     float* dq, float* scalars) {
    pdl_enter();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    float sq = 0.f, msk = 0.f;
    if (m < M) {
        float a = 0.f, b = 0.f;
        for (int n = 0; n < N; ++n) { a += q_chosen[m * N + n]; b += q_tc[m * N + n]; }   // mixer.py:16
        q_tot[m] = a; q_tot_t[m] = b;
        const float g = td_grad(a, b, r[m], term[m], padded[m], gamma, sq, msk);
        if (dq)
            for (int n = 0; n < N; ++n) {
                const long long o = ((long long)m * N + n) * A;
                const int ua = (int)u[m * N + n];
                for (int c = 0; c < A; ++c) dq[o + c] = (c == ua) ? g : 0.0f;
            }
    }
    block_accumulate2(sq, msk, scalars);
}

// The same with one warp per sample, for the learner: optionally does the action-value selection itself (warp_select) and
// emits dL/dh through the agents' head (warp_dhext), so that neither q_select nor the fc2 dgrad GEMM is launched.
constexpr int kVdnWarps = 8;
constexpr int kVdnMaxAgents = 64;
struct VdnFusedArgs {
    int M, N, A;
    const float* q_chosen; const float* q_tc; const long long* u;
    const float* r; const float* term; const float* padded; float gamma;
    float* q_tot; float* q_tot_t; float* dq; float* scalars;
    const float* fc2_w; float* dhext;
    SelectArgs sel;
};

__global__ void __launch_bounds__(kVdnWarps * 32) vdn_td_fused_kernel(VdnFusedArgs a) {
    pdl_enter();
    __shared__ float ssel[kVdnWarps][3][kVdnMaxAgents];
    extern __shared__ __align__(16) float ssel_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, N = a.N, A = a.A;
    const float* heads = nullptr;
    if (a.sel.q && a.sel.heads) {
        float* h = ssel_dyn + kVdnWarps * select_warp_floats(N, A, true);
        stage_heads(a.sel, A, h);
        heads = h;
        __syncthreads();
    }
    float sq = 0.f, msk = 0.f;
    for (long long m = blockIdx.x * kVdnWarps + warp; m < a.M; m += (long long)gridDim.x * kVdnWarps) {
        const float* q = a.q_chosen + m * N;
        const float* qt = a.q_tc + m * N;
        if (a.sel.q) {
            warp_select(a.sel, m, N, A, lane, ssel_dyn + warp * select_warp_floats(N, A, a.sel.heads), ssel[warp][0], ssel[warp][1], heads);
            q = ssel[warp][0]; qt = ssel[warp][1];
        }
        float tot = 0.f, tot_t = 0.f;
        for (int n = 0; n < N; ++n) { tot += q[n]; tot_t += qt[n]; }                     // mixer.py:16, agent order
        float s1, s2;
        const float g = td_grad(tot, tot_t, a.r[m], a.term[m], a.padded[m], a.gamma, s1, s2);
        if (lane == 0) { a.q_tot[m] = tot; a.q_tot_t[m] = tot_t; sq += s1; msk += s2; }
        if (a.dq)
            for (int i = lane; i < N * A; i += 32) {
                const int n = i / A, c = i - n * A;
                a.dq[m * N * A + i] = (c == (int)a.u[m * N + n]) ? g : 0.0f;
            }
        if (a.dhext) {
            for (int n = lane; n < N; n += 32) ssel[warp][2][n] = g;
            __syncwarp();
            warp_dhext(a.fc2_w, a.u, m, N, lane, ssel[warp][2], a.dhext);
        }
        __syncwarp();
    }
    block_accumulate2(sq, msk, a.scalars);
}

struct IngestKey { const void* src; void* dst; int inner; int kind; };   // kind 0: f64->f32, 1: f64->i64, 2: i64->i64, 3: f32->f32
struct IngestArgs { IngestKey k[11]; int B, L, T_src; const long long* idx; };   // idx (nullable): source episode of output episode b

__global__ void __launch_bounds__(256) ingest_kernel(IngestArgs a) {
    const IngestKey key = a.k[blockIdx.y];
    const long long per_b = (long long)a.L * key.inner, total = (long long)a.B * per_b;
    const long long src_b = (long long)a.T_src * key.inner;
    const long long tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    // 128-bit paths need 16-byte aligned bases as well as quad-sized rows (a caller's tensor view may start anywhere)
    const bool al16 = ((reinterpret_cast<uintptr_t>(key.src) | reinterpret_cast<uintptr_t>(key.dst)) & 15) == 0;
    if (a.idx && key.kind == 3 && al16 && (per_b & 3) == 0 && (src_b & 3) == 0) {
        // gather of sampled episodes (device replay buffer): 128-bit copies, a quad never straddles two episodes
        const long long qb = per_b >> 2, nq = (long long)a.B * qb;
        const float4* s = (const float4*)key.src; float4* d = (float4*)key.dst;
        for (long long i = tid0; i < nq; i += stride) {
            const long long b = i / qb, rem = i - b * qb;
            d[i] = __ldg(s + a.idx[b] * (src_b >> 2) + rem);
        }
        return;
    }
    if (!a.idx && al16 && a.T_src == a.L && (total & 3) == 0 && (key.kind == 0 || key.kind == 3)) {
        // no truncation: straight vectorised cast / copy
        const long long nq = total >> 2;
        if (key.kind == 0) {
            const double2* s = (const double2*)key.src; float4* d = (float4*)key.dst;
            for (long long i = tid0; i < nq; i += stride) {
                const double2 u0 = s[2 * i], u1 = s[2 * i + 1];
                d[i] = make_float4((float)u0.x, (float)u0.y, (float)u1.x, (float)u1.y);
            }
        } else {
            const float4* s = (const float4*)key.src; float4* d = (float4*)key.dst;
            for (long long i = tid0; i < nq; i += stride) d[i] = s[i];
        }
        return;
    }
    for (long long i = tid0; i < total; i += stride) {
        const long long b = i / per_b, rem = i - b * per_b, si = (a.idx ? a.idx[b] : b) * src_b + rem;
        if (key.kind == 0) ((float*)key.dst)[i] = (float)((const double*)key.src)[si];
        else if (key.kind == 1) ((long long*)key.dst)[i] = (long long)((const double*)key.src)[si];   // trunc toward zero
        else if (key.kind == 2) ((long long*)key.dst)[i] = ((const long long*)key.src)[si];
        else ((float*)key.dst)[i] = ((const float*)key.src)[si];
    }
}

}  // namespace marl

using namespace marl;

extern "C" int marl_q_select(const marl_dims* d, const float* q_evals, const long long* u, const float* q_evals_next,
                             float* q_targets, const float* avail_u_next, const float* avail_u, float* q_chosen,
                             long long* a_star, float* q_targets_chosen, float* max_q_evals, float* q_targets_max,
                             float* a_star_onehot, void* stream) {
    if (!d || !q_targets || !avail_u_next || !q_targets_chosen) return MARL_EINVAL;
    if (q_chosen && (!q_evals || !u)) return MARL_EINVAL;
    if (max_q_evals && (!avail_u || !q_evals)) return MARL_EINVAL;
    const int rows = d->B * d->L * d->N;
    if (rows <= 0) return MARL_OK;
    pdl_scope(rows);
    { ProfScope ps_("q_select_kernel", (cudaStream_t)stream); launch_pdl(q_select_kernel, dim3((rows + 255) / 256), dim3(256), 0, (cudaStream_t)stream, rows, d->A, q_evals, u, q_evals_next, q_targets,
        avail_u_next, avail_u, q_chosen, a_star, q_targets_chosen, max_q_evals, q_targets_max, a_star_onehot); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

extern "C" int marl_epsgreedy_select(int rows, int A, const float* q, const float* avail, const unsigned char* explore,
                                     const long long* random_action, long long* action, float* onehot, void* stream) {
    if (rows < 0 || A <= 0 || !q || !action || (explore && !random_action)) return MARL_EINVAL;
    if (rows == 0) return MARL_OK;
    { ProfScope ps_("epsgreedy_kernel", (cudaStream_t)stream);
      epsgreedy_kernel<<<(rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rows, A, q, avail, explore, random_action, action, onehot); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

extern "C" int marl_td_loss(int M, const float* q_tot, const float* q_tot_target, const float* r, const float* terminated,
                            const float* padded, float gamma, float* dq_tot, float* scalars, void* stream) {
    if (M < 0 || !q_tot || !q_tot_target || !r || !terminated || !padded || !scalars) return MARL_EINVAL;
    if (M == 0) return MARL_OK;
    { ProfScope ps_("td_loss_kernel", (cudaStream_t)stream); td_loss_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(M, q_tot, q_tot_target, r, terminated, padded, gamma,
                                                                     dq_tot, scalars); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

extern "C" int marl_vdn_td_fwd_bwd(const marl_dims* d, float* q_chosen, float* q_targets_chosen,
                                   const long long* u, const float* r, const float* terminated, const float* padded,
                                   float gamma, float* q_tot, float* q_tot_target, float* dq, float* scalars,
                                   const float* fc2_w, float* dhext, const marl_select_fused* sel, void* stream) {
    if (!d || !q_chosen || !q_targets_chosen || !r || !terminated || !padded || !q_tot || !q_tot_target || !scalars)
        return MARL_EINVAL;
    if ((dq || dhext || sel) && !u) return MARL_EINVAL;
    if (dhext && (!fc2_w || ((uintptr_t)fc2_w & 7) || ((uintptr_t)dhext & 7))) return MARL_EINVAL;
    if (sel && !select_ok(sel)) return MARL_EINVAL;
    const int M = d->B * d->L;
    if (M <= 0) return MARL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    pdl_scope((long long)M * d->N);
    if (!sel && !dhext) {
        ProfScope ps_("vdn_td_kernel", st);
        launch_pdl(vdn_td_kernel, dim3((M + 255) / 256), dim3(256), 0, st, M, d->N, d->A, (const float*)q_chosen,
                   (const float*)q_targets_chosen, u, r, terminated, padded, gamma, q_tot, q_tot_target, dq, scalars);
    } else {
        if (d->N > kVdnMaxAgents) return MARL_EINVAL;
        const size_t dyn = sel ? select_smem(kVdnWarps, d->N, d->A, sel->hidden_evals != nullptr) : 0;
        if (dyn > kSelectSmemMax) return MARL_EINVAL;
        if (dyn > 40 * 1024) {
            static bool attr_set = false;
            if (!attr_set) { cudaFuncSetAttribute(vdn_td_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelectSmemMax); attr_set = true; }
        }
        VdnFusedArgs a{};
        a.M = M; a.N = d->N; a.A = d->A; a.q_chosen = q_chosen; a.q_tc = q_targets_chosen; a.u = u;
        a.r = r; a.term = terminated; a.padded = padded; a.gamma = gamma;
        a.q_tot = q_tot; a.q_tot_t = q_tot_target; a.dq = dq; a.scalars = scalars; a.fc2_w = fc2_w; a.dhext = dhext;
        if (sel) a.sel = select_args(sel, u, q_chosen, q_targets_chosen);
        int blocks = (M + kVdnWarps - 1) / kVdnWarps;
        blocks = blocks > 4 * kNumSMs ? 4 * kNumSMs : blocks;
        ProfScope ps_("vdn_td_fused_kernel", st);
        launch_pdl(vdn_td_fused_kernel, dim3(blocks), dim3(kVdnWarps * 32), dyn, st, a);
    }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

extern "C" int marl_ingest_f64(const marl_episode_f64* s, int T_src, const marl_dims* d, const marl_episode_f32* o,
                               void* stream) {
    if (!s || !d || !o || T_src < d->L || d->B <= 0 || d->L <= 0) return MARL_EINVAL;
    IngestArgs a{};
    a.B = d->B; a.L = d->L; a.T_src = T_src;
    const int NO = d->N * d->O, NA = d->N * d->A;
    a.k[0] = {s->o, o->o, NO, 0};
    a.k[1] = {s->u, o->u, d->N, s->u_is_int64 ? 2 : 1};
    a.k[2] = {s->s, o->s, d->S, 0};
    a.k[3] = {s->r, o->r, 1, 0};
    a.k[4] = {s->o_next, o->o_next, NO, 0};
    a.k[5] = {s->s_next, o->s_next, d->S, 0};
    a.k[6] = {s->avail_u, o->avail_u, NA, 0};
    a.k[7] = {s->avail_u_next, o->avail_u_next, NA, 0};
    a.k[8] = {s->u_onehot, o->u_onehot, NA, 0};
    a.k[9] = {s->padded, o->padded, 1, 0};
    a.k[10] = {s->terminated, o->terminated, 1, 0};
    for (int i = 0; i < 11; ++i)
        if (!a.k[i].src || !a.k[i].dst) return MARL_EINVAL;
    long long biggest = (long long)d->B * d->L * (NO > NA ? NO : NA);
    int bx = (int)((biggest + 4095) / 4096);
    if (bx < 1) bx = 1;
    if (bx > kNumSMs) bx = kNumSMs;
    { ProfScope ps_("ingest_kernel", (cudaStream_t)stream); ingest_kernel<<<dim3(bx, 11), 256, 0, (cudaStream_t)stream>>>(a); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

static int ingest_f32_impl(const marl_episode_f32* s, int T_src, const long long* idx, const marl_dims* d,
                           const marl_episode_f32* o, void* stream) {
    if (!s || !d || !o || T_src < d->L || d->B <= 0 || d->L <= 0) return MARL_EINVAL;
    IngestArgs a{};
    a.B = d->B; a.L = d->L; a.T_src = T_src; a.idx = idx;
    const int NO = d->N * d->O, NA = d->N * d->A;
    a.k[0] = {s->o, o->o, NO, 3};
    a.k[1] = {s->u, o->u, d->N, 2};
    a.k[2] = {s->s, o->s, d->S, 3};
    a.k[3] = {s->r, o->r, 1, 3};
    a.k[4] = {s->o_next, o->o_next, NO, 3};
    a.k[5] = {s->s_next, o->s_next, d->S, 3};
    a.k[6] = {s->avail_u, o->avail_u, NA, 3};
    a.k[7] = {s->avail_u_next, o->avail_u_next, NA, 3};
    a.k[8] = {s->u_onehot, o->u_onehot, NA, 3};
    a.k[9] = {s->padded, o->padded, 1, 3};
    a.k[10] = {s->terminated, o->terminated, 1, 3};
    for (int i = 0; i < 11; ++i)
        if (!a.k[i].src || !a.k[i].dst) return MARL_EINVAL;
    long long biggest = (long long)d->B * d->L * (NO > NA ? NO : NA);
    int bx = (int)((biggest + 4095) / 4096);
    if (bx < 1) bx = 1;
    if (bx > kNumSMs) bx = kNumSMs;
    { ProfScope ps_("ingest_kernel", (cudaStream_t)stream); ingest_kernel<<<dim3(bx, 11), 256, 0, (cudaStream_t)stream>>>(a); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

extern "C" int marl_ingest_f32(const marl_episode_f32* s, int T_src, const marl_dims* d, const marl_episode_f32* o,
                               void* stream) {
    return ingest_f32_impl(s, T_src, nullptr, d, o, stream);
}

extern "C" int marl_replay_gather_f32(const marl_episode_f32* ring, int T_ring, const long long* episode_idx,
                                      const marl_dims* d, const marl_episode_f32* o, void* stream) {
    if (!episode_idx) return MARL_EINVAL;
    return ingest_f32_impl(ring, T_ring, episode_idx, d, o, stream);
}
