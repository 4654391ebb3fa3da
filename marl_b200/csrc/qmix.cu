// QMIX monotonic mixer (network/mixer.py:21-80), forward + backward, optionally fused with the
// target mixer and the TD loss (algorithm/q_learner.py:161-168).
//
//   hy = s . wcat^T + bcat        one GEMM for the four hyper-network heads (linear_fwd), C = N*E+3E
//   one warp per (b,t) sample, lane = embed index e (E = 32):
//     hid_e = elu(sum_n q_n |w1[n,e]| + b1_e);  q_tot = sum_e hid_e |w2_e| + (sum_e relu(hb2_e) wb2_e + bb2)
//   the agents x embed contraction lives in registers, the embed reduction is a warp shuffle;
//   backward in the same warp emits the per-sample gradient row dhy[C] that feeds the
//   weight-gradient GEMM dwcat += dhy^T . s.
#include "linear.h"
#include "forkjoin.h"
#include "../../include/marl_b200.h"
#include "profile.h"

namespace marl {

constexpr int E = MARL_QMIX_EMBED;
constexpr int kQmixWarps = 8;
constexpr int kQmixMaxAgents = 64;

enum { QMIX_FWD = 0, QMIX_BWD = 1, QMIX_TD = 2 };

struct QmixMixArgs {
    int M, N, A, mode;
    const float* hy;  const float* q;  const float* wb2;  const float* bb2;            // eval side
    const float* hy_t; const float* q_t; const float* wb2_t; const float* bb2_t;       // target side (TD)
    const float* r; const float* term; const float* padded; float gamma;
    const float* dq_tot_in;       // BWD
    float* q_tot; float* q_tot_t;
    float* dhy;                   // [M,C]
    float* dq_small;              // [M,N] or null
    const long long* u; float* dq_dense;   // [M,N,A] or null
    float* g_wb2; float* g_bb2; float* scalars;
};

// forward of one sample for one lane; returns q_tot (all lanes) and the lane's intermediates
__device__ __forceinline__ float qmix_forward_lane(const float* __restrict__ y, const float* __restrict__ q, int N,
                                                   float wb2e, float bb2, int lane, float& pre, float& hid,
                                                   float& w2raw, float& hb2) {
    float acc = y[N * E + lane];                                   // hyper_b1
    for (int n = 0; n < N; ++n) acc = fmaf(q[n], fabsf(y[n * E + lane]), acc);   // bmm(q, |w1|) + b1, mixer.py:64-70
    pre = acc;
    hid = acc > 0.0f ? acc : (expf(acc) - 1.0f);                   // F.elu
    w2raw = y[N * E + E + lane];
    hb2 = y[N * E + 2 * E + lane];
    float part = hid * fabsf(w2raw) + fmaxf(hb2, 0.0f) * wb2e;     // bmm(hidden, |w2|) + hyper_b2(s), mixer.py:72-78
    return warp_sum(part) + bb2;
}

__global__ void __launch_bounds__(kQmixWarps * 32) qmix_mix_kernel(QmixMixArgs a) {
    __shared__ float sdq[kQmixWarps][kQmixMaxAgents];
    __shared__ float sred[kQmixWarps][E + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = a.N, C = N * E + 3 * E;
    const float wb2e = a.wb2[lane], bb2 = a.bb2[0];
    float wb2te = 0.f, bb2t = 0.f;
    if (a.mode == QMIX_TD) { wb2te = a.wb2_t[lane]; bb2t = a.bb2_t[0]; }
    float acc_wb2 = 0.f, acc_bb2 = 0.f, acc_sq = 0.f, acc_mask = 0.f;
    for (int m = blockIdx.x * kQmixWarps + warp; m < a.M; m += gridDim.x * kQmixWarps) {
        const float* y = a.hy + (long long)m * C;
        const float* q = a.q + (long long)m * N;
        float pre, hid, w2raw, hb2;
        const float tot = qmix_forward_lane(y, q, N, wb2e, bb2, lane, pre, hid, w2raw, hb2);
        if (lane == 0 && a.q_tot) a.q_tot[m] = tot;
        if (a.mode == QMIX_FWD) continue;
        float G;
        if (a.mode == QMIX_TD) {
            float p2, h2, w2t, hbt;
            const float tot_t = qmix_forward_lane(a.hy_t + (long long)m * C, a.q_t + (long long)m * N, N, wb2te, bb2t,
                                                  lane, p2, h2, w2t, hbt);
            if (lane == 0 && a.q_tot_t) a.q_tot_t[m] = tot_t;
            const float mask = 1.0f - a.padded[m];                                   // q_learner.py:83
            const float yt = a.r[m] + a.gamma * tot_t * (1.0f - a.term[m]);          // :165
            const float d = mask * (yt - tot);                                       // :166-167
            acc_sq += d * d; acc_mask += mask;                                       // (lane-uniform)
            G = -2.0f * mask * d;
        } else {
            G = a.dq_tot_in[m];
        }
        // backward
        float* dy = a.dhy + (long long)m * C;
        const float dpre = G * fabsf(w2raw) * (pre > 0.0f ? 1.0f : (hid + 1.0f));   // elu'(x) = exp(x) for x <= 0
        const float sgn2 = (w2raw > 0.0f) ? 1.0f : (w2raw < 0.0f ? -1.0f : 0.0f);
        dy[N * E + lane] = dpre;                                                    // d hyper_b1
        dy[N * E + E + lane] = G * hid * sgn2;                                      // d hyper_w2 (through abs)
        dy[N * E + 2 * E + lane] = (hb2 > 0.0f) ? G * wb2e : 0.0f;                  // d hyper_b2.0 (through relu)
        acc_wb2 += G * fmaxf(hb2, 0.0f);
        acc_bb2 += G;
        for (int n = 0; n < N; ++n) {
            const float w1raw = y[n * E + lane];
            const float sgn1 = (w1raw > 0.0f) ? 1.0f : (w1raw < 0.0f ? -1.0f : 0.0f);
            dy[n * E + lane] = dpre * q[n] * sgn1;                                  // d hyper_w1 (through abs)
            const float dqn = warp_sum(dpre * fabsf(w1raw));
            if (lane == 0) sdq[warp][n] = dqn;
        }
        __syncwarp();
        if (a.dq_small)
            for (int n = lane; n < N; n += 32) a.dq_small[(long long)m * N + n] = sdq[warp][n];
        if (a.dq_dense) {
            const int A = a.A;
            for (int i = lane; i < N * A; i += 32) {
                const int n = i / A, c = i - n * A;
                a.dq_dense[(long long)m * N * A + i] = (c == (int)a.u[(long long)m * N + n]) ? sdq[warp][n] : 0.0f;
            }
        }
        __syncwarp();
    }
    if (a.mode == QMIX_FWD) return;
    // block reduction of the hyper_b2.2 gradient and the loss scalars
    sred[warp][lane] = acc_wb2;
    if (lane == 0) sred[warp][E] = acc_bb2;
    __syncthreads();
    if (warp == 0) {
        float v = 0.f, b = 0.f;
        for (int w = 0; w < kQmixWarps; ++w) { v += sred[w][lane]; b += sred[w][E]; }
        if (a.g_wb2) atomicAdd(a.g_wb2 + lane, v);
        if (lane == 0 && a.g_bb2) atomicAdd(a.g_bb2, b);
    }
    if (a.mode == QMIX_TD) {
        __syncthreads();
        if (lane == 0) { sred[warp][0] = acc_sq; sred[warp][1] = acc_mask; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float v = 0.f, b = 0.f;
            for (int w = 0; w < kQmixWarps; ++w) { v += sred[w][0]; b += sred[w][1]; }
            atomicAdd(a.scalars, v); atomicAdd(a.scalars + 1, b);
        }
    }
}

static int hyper_fwd(int M, int N, int S, const marl_qmix_params* p, const float* s, float* hy, cudaStream_t st) {
    LinearFwd f{};
    f.in = plain_operand(s, S, S);
    f.w = p->wcat; f.ldw = S; f.bias = p->bcat;
    f.y = hy; f.ldy = N * E + 3 * E; f.M = M; f.N = N * E + 3 * E; f.batch = 1;
    return linear_fwd(f, st);
}

static int hyper_wgrad(int M, int N, int S, const float* s, const float* dhy, const marl_qmix_grads* g, cudaStream_t st) {
    LinearWgrad w{};
    w.dy = dhy; w.lddy = N * E + 3 * E; w.in = plain_operand(s, S, S);
    w.dw = g->wcat; w.ldw = S; w.db = g->bcat; w.M = M; w.N = N * E + 3 * E; w.batch = 1;
    return linear_wgrad(w, st);
}

static int mix_grid(int M) {
    int blocks = (M + kQmixWarps - 1) / kQmixWarps;
    return blocks < 1 ? 1 : (blocks > 4 * kNumSMs ? 4 * kNumSMs : blocks);
}

}  // namespace marl

using namespace marl;

static bool qmix_params_ok(const marl_qmix_params* p) { return p && p->wcat && p->bcat && p->wb2 && p->bb2; }

extern "C" int marl_qmix_fwd(int M, int N, int S, const marl_qmix_params* p, const float* q, const float* s, float* hy,
                             float* q_tot, void* stream) {
    if (M < 0 || N < 1 || N > kQmixMaxAgents || S < 1 || !qmix_params_ok(p) || !q || !s || !hy || !q_tot) return MARL_EINVAL;
    if (M == 0) return MARL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = hyper_fwd(M, N, S, p, s, hy, st);
    if (rc) return rc;
    QmixMixArgs a{};
    a.M = M; a.N = N; a.mode = QMIX_FWD; a.hy = hy; a.q = q; a.wb2 = p->wb2; a.bb2 = p->bb2; a.q_tot = q_tot;
    { ProfScope ps_("qmix_mix_kernel", st); qmix_mix_kernel<<<mix_grid(M), kQmixWarps * 32, 0, st>>>(a); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

extern "C" int marl_qmix_bwd(int M, int N, int S, const marl_qmix_params* p, const float* q, const float* s,
                             const float* hy, const float* dq_tot, float* dhy, float* dq, const marl_qmix_grads* g,
                             void* stream) {
    if (M < 0 || N < 1 || N > kQmixMaxAgents || S < 1 || !qmix_params_ok(p) || !q || !s || !hy || !dq_tot || !dhy || !g)
        return MARL_EINVAL;
    if (M == 0) return MARL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    QmixMixArgs a{};
    a.M = M; a.N = N; a.mode = QMIX_BWD; a.hy = hy; a.q = q; a.wb2 = p->wb2; a.bb2 = p->bb2;
    a.dq_tot_in = dq_tot; a.dhy = dhy; a.dq_small = dq; a.g_wb2 = g->wb2; a.g_bb2 = g->bb2;
    { ProfScope ps_("qmix_mix_kernel", st); qmix_mix_kernel<<<mix_grid(M), kQmixWarps * 32, 0, st>>>(a); }
    MARL_LAUNCH_CHECK();
    return hyper_wgrad(M, N, S, s, dhy, g, st);
}

extern "C" int marl_qmix_td_fwd_bwd(const marl_dims* d, const marl_qmix_params* p, const marl_qmix_params* pt,
                                    const float* s, const float* s_next, const float* q_chosen, const float* q_tc,
                                    const long long* u, const float* r, const float* terminated, const float* padded,
                                    float gamma, float* hy, float* hy_target, float* dhy, float* q_tot, float* q_tot_target,
                                    float* dq, const marl_qmix_grads* g, float* scalars, int flags, void* stream) {
    if (!d || !qmix_params_ok(p) || !qmix_params_ok(pt) || !s || !s_next || !q_chosen || !q_tc || !r || !terminated ||
        !padded || !hy || !hy_target || !dhy || !g || !scalars)
        return MARL_EINVAL;
    if (dq && !u) return MARL_EINVAL;
    if (d->N < 1 || d->N > kQmixMaxAgents || d->S < 1) return MARL_EINVAL;
    const int M = d->B * d->L;
    if (M <= 0) return MARL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (!(flags & 1)) {
        ForkJoin fj(st, 2);      // eval and target hyper-networks side by side
        if ((rc = hyper_fwd(M, d->N, d->S, p, s, hy, fj.lane(0)))) return rc;
        if ((rc = hyper_fwd(M, d->N, d->S, pt, s_next, hy_target, fj.lane(1)))) return rc;
        fj.join();
    }
    QmixMixArgs a{};
    a.M = M; a.N = d->N; a.A = d->A; a.mode = QMIX_TD;
    a.hy = hy; a.q = q_chosen; a.wb2 = p->wb2; a.bb2 = p->bb2;
    a.hy_t = hy_target; a.q_t = q_tc; a.wb2_t = pt->wb2; a.bb2_t = pt->bb2;
    a.r = r; a.term = terminated; a.padded = padded; a.gamma = gamma;
    a.q_tot = q_tot; a.q_tot_t = q_tot_target; a.dhy = dhy; a.u = u; a.dq_dense = dq;
    a.g_wb2 = g->wb2; a.g_bb2 = g->bb2; a.scalars = scalars;
    { ProfScope ps_("qmix_mix_kernel", st); qmix_mix_kernel<<<mix_grid(M), kQmixWarps * 32, 0, st>>>(a); }
    MARL_LAUNCH_CHECK();
    if (flags & 2) return MARL_OK;
    return hyper_wgrad(M, d->N, d->S, s, dhy, g, st);
}

extern "C" int marl_qmix_hyper_fwd(int M, int N, int S, const marl_qmix_params* p, const float* s, float* hy, void* stream) {
    if (M < 0 || N < 1 || N > kQmixMaxAgents || S < 1 || !qmix_params_ok(p) || !s || !hy) return MARL_EINVAL;
    if (M == 0) return MARL_OK;
    return hyper_fwd(M, N, S, p, s, hy, (cudaStream_t)stream);
}

extern "C" int marl_qmix_hyper_wgrad(int M, int N, int S, const float* s, const float* dhy, const marl_qmix_grads* g,
                                     void* stream) {
    if (M < 0 || N < 1 || N > kQmixMaxAgents || S < 1 || !s || !dhy || !g || !g->wcat || !g->bcat) return MARL_EINVAL;
    if (M == 0) return MARL_OK;
    return hyper_wgrad(M, N, S, s, dhy, g, (cudaStream_t)stream);
}
