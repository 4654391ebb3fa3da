// QPLEX duplex-dueling mixer: DMAQer + DMAQ_SI_Weight (network/mixer.py:85-288), forward + backward.
//
// The reference evaluates 2 two-layer MLPs (hyper_w_final, V) and 3*num_kernel three-layer MLPs
// (key / agents / action extractors) as ~200 tiny cuBLAS calls per mixer call.  Here the layers are
// concatenated (the host lays the parameters out for it) so that one mixer call is 6 GEMM launches:
//   L1s : h1[:, 0:Ws]       = relu(s        . w1s^T + b1s)   Ws = 2*he + 2*K*ae  (hyper_w_final.0 | V.0 | key.k.0 | agents.k.0)
//   L1a : h1[:, Ws:Ws+K*ae] = relu([s | a]  . w1a^T + b1a)                       (action.k.0; [s | a] is never materialised)
//   L2  : h2[:, z*ae:(z+1)*ae] = relu(h1[:, 2he + z*ae ...] . w2[z]^T + b2[z])    batched over the 3K extractors
//   L3k : o3[:, k]          = h2[:, k*ae ...] . w3k[k] + b3k[k]                   batched over K   (key.k.4)
//   L3n : o3[:, K + z*N ..] = h2[:, (K+z)*ae ...] . w3n[z]^T + b3n[z]             batched over 2K  (agents.k.4 | action.k.4)
//   LFV : wv[:, z*N ..]     = h1[:, z*he ...] . wfv[z]^T + bfv[z]                 batched over 2   (hyper_w_final.2 | V.2)
// followed by one warp-per-sample kernel (lane = agent) doing the transformation, the lambda heads and
// the dueling sums; its backward emits d(o3), d(wv), dq and the GEMM chain is walked in reverse.
#include "linear.h"
#include "forkjoin.h"
#include "../../include/marl_b200.h"
#include "profile.h"

namespace marl {

struct QplexMixArgs {
    int M, N, K;
    const float* wv; const float* o3; const float* q; const float* max_q;
    const float* dv_tot; const float* da_tot;   // both null: forward only
    float* v_tot; float* a_tot; float* q_tot;
    float* dwv; float* do3; float* dq;
    int weighted_head, minus_one, with_adv;
};

__global__ void __launch_bounds__(256) qplex_mix_kernel(QplexMixArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = a.N, K = a.K, W3 = K + 2 * K * N;
    for (int m = blockIdx.x * 8 + warp; m < a.M; m += gridDim.x * 8) {
        const bool act = lane < N;
        const float* wv = a.wv + (long long)m * 2 * N;
        const float* o3 = a.o3 + (long long)m * W3;
        const float wraw = act ? wv[lane] : 0.f, v = act ? wv[N + lane] : 0.f;
        const float w = fabsf(wraw) + 1e-10f;                               // mixer.py:269-270
        const float q = act ? a.q[(long long)m * N + lane] : 0.f;
        const float qh = a.weighted_head ? (w * q + v) : q;                 // mixer.py:275-276
        const float vt = warp_sum(act ? qh : 0.f);                          // calc_v, mixer.py:218-220
        float adv = 0.f, lam = 0.f, at = 0.f;
        if (a.with_adv) {
            const float mq = act ? a.max_q[(long long)m * N + lane] : 0.f;
            const float mh = a.weighted_head ? (w * mq + v) : mq;           // mixer.py:279-281
            adv = qh - mh;                                                  // detached, mixer.py:237
            for (int k = 0; k < K; ++k) {                                   // DMAQ_SI_Weight.forward, mixer.py:159-169
                const float key = fabsf(o3[k]) + 1e-10f;
                const float ag = act ? sigmoidf_acc(o3[K + k * N + lane]) : 0.f;
                const float ac = act ? sigmoidf_acc(o3[K + K * N + k * N + lane]) : 0.f;
                lam += key * ag * ac;
            }
            at = warp_sum(act ? adv * (a.minus_one ? (lam - 1.0f) : lam) : 0.f);   // mixer.py:243-246
        }
        if (lane == 0) {
            if (a.v_tot) a.v_tot[m] = vt;
            if (a.a_tot) a.a_tot[m] = at;
            if (a.q_tot) a.q_tot[m] = vt + at;
        }
        if (!a.dv_tot && !a.da_tot) continue;
        const float G = a.dv_tot ? a.dv_tot[m] : 0.f;
        const float Ga = a.da_tot ? a.da_tot[m] : 0.f;
        float* dwv = a.dwv + (long long)m * 2 * N;
        if (act) {
            const float sgn = wraw > 0.f ? 1.f : (wraw < 0.f ? -1.f : 0.f);
            if (a.weighted_head) { dwv[lane] = G * q * sgn; dwv[N + lane] = G; a.dq[(long long)m * N + lane] = G * w; }
            else { dwv[lane] = 0.f; dwv[N + lane] = 0.f; a.dq[(long long)m * N + lane] = G; }
        }
        if (a.with_adv) {
            float* do3 = a.do3 + (long long)m * W3;
            const float dlam = Ga * adv;                                     // only the lambda heads see a_tot's gradient
            for (int k = 0; k < K; ++k) {
                const float kraw = o3[k];
                const float key = fabsf(kraw) + 1e-10f;
                const float ag = act ? sigmoidf_acc(o3[K + k * N + lane]) : 0.f;
                const float ac = act ? sigmoidf_acc(o3[K + K * N + k * N + lane]) : 0.f;
                const float dkey = warp_sum(act ? dlam * ag * ac : 0.f);
                if (lane == 0) do3[k] = dkey * (kraw > 0.f ? 1.f : (kraw < 0.f ? -1.f : 0.f));
                if (act) {
                    do3[K + k * N + lane] = dlam * key * ac * ag * (1.0f - ag);
                    do3[K + K * N + k * N + lane] = dlam * key * ag * ac * (1.0f - ac);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) scatter_dq_kernel(int rows, int A, const float* dq_small, const long long* u, float* dq) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int ua = (int)u[i];
    const float g = dq_small[i];
    for (int c = 0; c < A; ++c) dq[(long long)i * A + c] = (c == ua) ? g : 0.0f;
}

// `layers` = adv_hypernet_layers (network/mixer.py:115-145): Linear layers per lambda extractor.
//   3 (default): L1 -> ReLU -> L2 -> ReLU -> L3 ; 2: L1 -> ReLU -> L3 (the first hidden layer feeds the output layer: h2 is h1's
//   extractor block) ; 1: one Linear straight from the state / [state | actions] (no extractor columns in h1 at all).
struct QplexLayout { int he, ae, K, N, A, S, Ws, W1, W2, W3, layers; };
static QplexLayout layout(const marl_qplex_dims* d) {
    QplexLayout l{d->he, d->ae, d->K, d->N, d->A, d->S, 0, 0, 0, 0, d->layers ? d->layers : 3};
    const int ext = l.layers == 1 ? 0 : d->K * d->ae;          // hidden columns per extractor family
    l.Ws = 2 * d->he + 2 * ext;
    l.W1 = l.Ws + ext;
    l.W2 = 3 * ext;
    l.W3 = d->K + 2 * d->K * d->N;
    return l;
}
// where the output layers of the extractors read their input: h2 (3 layers) or the extractor block of h1 (2 layers)
static const float* ext_hidden(const QplexLayout& l, const marl_qplex_ws* ws, int& pitch) {
    if (l.layers == 3) { pitch = l.W2; return ws->h2; }
    pitch = l.W1; return ws->h1 + 2 * l.he;
}
static float* ext_hidden_mut(const QplexLayout& l, const marl_qplex_ws* ws, int& pitch) {
    if (l.layers == 3) { pitch = l.W2; return ws->h2; }
    pitch = l.W1; return ws->h1 + 2 * l.he;
}

static int qplex_forward_gemms(int M, const QplexLayout& l, const marl_qplex_params* p, const float* s, const float* actions,
                               const marl_qplex_ws* ws, bool with_adv, cudaStream_t st) {
    int rc;
    {   // L1s (all heads fed by the state); without the advantage stream only hyper_w_final.0 | V.0 are needed
        LinearFwd f{};
        f.in = plain_operand(s, l.S, l.S);
        f.w = p->w1s; f.ldw = l.S; f.bias = p->b1s; f.y = ws->h1; f.ldy = l.W1;
        f.M = M; f.N = with_adv ? l.Ws : 2 * l.he; f.relu = 1; f.batch = 1;
        if ((rc = linear_fwd(f, st))) return rc;
    }
    {   // LFV
        LinearFwd f{};
        f.in = plain_operand(ws->h1, l.W1, l.he, l.he);
        f.w = p->wfv; f.ldw = l.he; f.w_bs = (long long)l.N * l.he; f.bias = p->bfv; f.b_bs = l.N;
        f.y = ws->wv; f.ldy = 2 * l.N; f.y_bs = l.N; f.M = M; f.N = l.N; f.batch = 2;
        if ((rc = linear_fwd(f, st))) return rc;
    }
    if (!with_adv) return MARL_OK;
    if (l.layers == 1) {
        // single-Linear extractors: o3 = [key_k(s) | agents_k(s) | action_k([s | a])] in two GEMMs (w3k holds the K key rows
        // followed by the K*N agents rows, w3n the K*N action rows)
        LinearFwd f{};
        f.in = plain_operand(s, l.S, l.S);
        f.w = p->w3k; f.ldw = l.S; f.bias = p->b3k; f.y = ws->o3; f.ldy = l.W3; f.M = M; f.N = l.K + l.K * l.N; f.batch = 1;
        if ((rc = linear_fwd(f, st))) return rc;
        LinearFwd g{};
        LinOperand in = plain_operand(s, l.S, l.S);
        in.x2 = actions; in.ldx2 = l.N * l.A; in.K2 = l.N * l.A;
        g.in = in;
        g.w = p->w3n; g.ldw = l.S + l.N * l.A; g.bias = p->b3n; g.y = ws->o3 + l.K + l.K * l.N; g.ldy = l.W3;
        g.M = M; g.N = l.K * l.N; g.batch = 1;
        return linear_fwd(g, st);
    }
    {   // L1a on the virtual concatenation [s | actions]
        LinearFwd f{};
        LinOperand in = plain_operand(s, l.S, l.S);
        in.x2 = actions; in.ldx2 = l.N * l.A; in.K2 = l.N * l.A;
        f.in = in;
        f.w = p->w1a; f.ldw = l.S + l.N * l.A; f.bias = p->b1a; f.y = ws->h1 + l.Ws; f.ldy = l.W1;
        f.M = M; f.N = l.K * l.ae; f.relu = 1; f.batch = 1;
        if ((rc = linear_fwd(f, st))) return rc;
    }
    if (l.layers == 3) {   // L2, batched over the 3K extractors
        LinearFwd f{};
        f.in = plain_operand(ws->h1 + 2 * l.he, l.W1, l.ae, l.ae);
        f.w = p->w2; f.ldw = l.ae; f.w_bs = (long long)l.ae * l.ae; f.bias = p->b2; f.b_bs = l.ae;
        f.y = ws->h2; f.ldy = l.W2; f.y_bs = l.ae; f.M = M; f.N = l.ae; f.relu = 1; f.batch = 3 * l.K;
        if ((rc = linear_fwd(f, st))) return rc;
    }
    int hp;
    const float* hx = ext_hidden(l, ws, hp);
    {   // L3k
        LinearFwd f{};
        f.in = plain_operand(hx, hp, l.ae, l.ae);
        f.w = p->w3k; f.ldw = l.ae; f.w_bs = l.ae; f.bias = p->b3k; f.b_bs = 1;
        f.y = ws->o3; f.ldy = l.W3; f.y_bs = 1; f.M = M; f.N = 1; f.batch = l.K;
        if ((rc = linear_fwd(f, st))) return rc;
    }
    {   // L3n
        LinearFwd f{};
        f.in = plain_operand(hx + l.K * l.ae, hp, l.ae, l.ae);
        f.w = p->w3n; f.ldw = l.ae; f.w_bs = (long long)l.N * l.ae; f.bias = p->b3n; f.b_bs = l.N;
        f.y = ws->o3 + l.K; f.ldy = l.W3; f.y_bs = l.N; f.M = M; f.N = l.N; f.batch = 2 * l.K;
        if ((rc = linear_fwd(f, st))) return rc;
    }
    return MARL_OK;
}

static bool qplex_ok(const marl_qplex_dims* d, const marl_qplex_params* p, const marl_qplex_ws* ws) {
    if (!(d && p && ws && d->N >= 1 && d->N <= 32 && d->K >= 1 && d->he >= 1 && d->ae >= 1 && d->S >= 1 && d->A >= 1)) return false;
    const int layers = d->layers ? d->layers : 3;
    if (layers < 1 || layers > 3) return false;
    if (!(p->w1s && p->b1s && p->w3k && p->b3k && p->w3n && p->b3n && p->wfv && p->bfv && ws->h1 && ws->o3 && ws->wv)) return false;
    if (layers >= 2 && !(p->w1a && p->b1a)) return false;
    if (layers == 3 && !(p->w2 && p->b2 && ws->h2)) return false;
    return true;
}

}  // namespace marl

using namespace marl;

extern "C" int marl_qplex_fwd(int M, const marl_qplex_dims* d, const marl_qplex_params* p, const float* q, const float* s,
                              const float* actions, const float* max_q, const marl_qplex_ws* ws, float* v_tot,
                              float* a_tot, float* q_tot, void* stream) {
    if (M < 0 || !qplex_ok(d, p, ws) || !q || !s) return MARL_EINVAL;
    const bool with_adv = actions != nullptr;
    if (with_adv && !max_q) return MARL_EINVAL;
    if (M == 0) return MARL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const QplexLayout l = layout(d);
    int rc = qplex_forward_gemms(M, l, p, s, actions, ws, with_adv, st);
    if (rc) return rc;
    QplexMixArgs a{};
    a.M = M; a.N = l.N; a.K = l.K; a.wv = ws->wv; a.o3 = ws->o3; a.q = q; a.max_q = max_q;
    a.v_tot = v_tot; a.a_tot = a_tot; a.q_tot = q_tot;
    a.weighted_head = d->weighted_head; a.minus_one = d->is_minus_one; a.with_adv = with_adv;
    int blocks = (M + 7) / 8; if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    { ProfScope ps_("qplex_mix_kernel", st); qplex_mix_kernel<<<blocks, 256, 0, st>>>(a); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

extern "C" int marl_qplex_bwd(int M, const marl_qplex_dims* d, const marl_qplex_params* p, const float* q, const float* s,
                              const float* actions, const float* max_q, const marl_qplex_ws* ws, const float* dv_tot,
                              const float* da_tot, const marl_qplex_ws* dws, float* dq, const marl_qplex_grads* g,
                              void* stream) {
    if (M < 0 || !qplex_ok(d, p, ws) || !q || !s || (!dv_tot && !da_tot) || !dws || !dq || !g) return MARL_EINVAL;
    if (!dws->h1 || !dws->o3 || !dws->wv || ((d->layers ? d->layers : 3) == 3 && !dws->h2)) return MARL_EINVAL;
    if (da_tot && !actions) return MARL_EINVAL;
    const bool with_adv = actions != nullptr && da_tot != nullptr;
    if (with_adv && !max_q) return MARL_EINVAL;
    if (M == 0) return MARL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const QplexLayout l = layout(d);
    int rc;
    {
        QplexMixArgs a{};
        a.M = M; a.N = l.N; a.K = l.K; a.wv = ws->wv; a.o3 = ws->o3; a.q = q; a.max_q = max_q; a.dv_tot = dv_tot; a.da_tot = da_tot;
        a.dwv = dws->wv; a.do3 = dws->o3; a.dq = dq;
        a.weighted_head = d->weighted_head; a.minus_one = d->is_minus_one; a.with_adv = with_adv;
        int blocks = (M + 7) / 8; if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
        { ProfScope ps_("qplex_mix_kernel", st); qplex_mix_kernel<<<blocks, 256, 0, st>>>(a); }
        MARL_LAUNCH_CHECK();
    }
    // The data gradients form the dependent chain (LFV, L3k, L3n, L2: small, latency-bound products); every weight gradient only
    // reads what the chain has already produced, so they run on a forked lane BESIDE it instead of between its links (eager
    // timeline of config 3, profiles/r2c_cfg3_timeline.txt: 1.34 ms of alternating weight / data gradients on one stream).
    ForkJoin fw(st, 2);
    cudaStream_t wl = fw.lane(1);
    {   // LFV backward: weights, then dh1[:, 0:2he] (through relu)
        LinearWgrad w{};
        w.dy = dws->wv; w.lddy = 2 * l.N; w.dy_bs = l.N; w.in = plain_operand(ws->h1, l.W1, l.he, l.he);
        w.dw = g->wfv; w.ldw = l.he; w.dw_bs = (long long)l.N * l.he; w.db = g->bfv; w.db_bs = l.N; w.M = M; w.N = l.N; w.batch = 2;
        if ((rc = linear_wgrad(w, wl))) return rc;
        LinearDgrad dg{};
        dg.dy = dws->wv; dg.lddy = 2 * l.N; dg.dy_bs = l.N; dg.w = p->wfv; dg.ldw = l.he; dg.w_bs = (long long)l.N * l.he;
        dg.dx = dws->h1; dg.lddx = l.W1; dg.dx_bs = l.he; dg.relu_src = ws->h1; dg.ldrs = l.W1; dg.rs_bs = l.he;
        dg.M = M; dg.N = l.N; dg.K = l.he; dg.batch = 2;
        if ((rc = linear_dgrad(dg, st))) return rc;
    }
    if (with_adv && l.layers == 1) {
        // single-Linear extractors: only weight gradients (their inputs, the state and the actions, need none)
        LinearWgrad w{};
        w.dy = dws->o3; w.lddy = l.W3; w.in = plain_operand(s, l.S, l.S);
        w.dw = g->w3k; w.ldw = l.S; w.db = g->b3k; w.M = M; w.N = l.K + l.K * l.N; w.batch = 1;
        if ((rc = linear_wgrad(w, wl))) return rc;
        LinearWgrad w2{};
        LinOperand in = plain_operand(s, l.S, l.S);
        in.x2 = actions; in.ldx2 = l.N * l.A; in.K2 = l.N * l.A;
        w2.dy = dws->o3 + l.K + l.K * l.N; w2.lddy = l.W3; w2.in = in;
        w2.dw = g->w3n; w2.ldw = l.S + l.N * l.A; w2.db = g->b3n; w2.M = M; w2.N = l.K * l.N; w2.batch = 1;
        if ((rc = linear_wgrad(w2, wl))) return rc;
    } else if (with_adv) {
        int hp, dhp;
        const float* hx = ext_hidden(l, ws, hp);
        float* dhx = ext_hidden_mut(l, dws, dhp);
        {   // L3k / L3n backward
            LinearWgrad w{};
            w.dy = dws->o3; w.lddy = l.W3; w.dy_bs = 1; w.in = plain_operand(hx, hp, l.ae, l.ae);
            w.dw = g->w3k; w.ldw = l.ae; w.dw_bs = l.ae; w.db = g->b3k; w.db_bs = 1; w.M = M; w.N = 1; w.batch = l.K;
            if ((rc = linear_wgrad(w, wl))) return rc;
            LinearDgrad dg{};
            dg.dy = dws->o3; dg.lddy = l.W3; dg.dy_bs = 1; dg.w = p->w3k; dg.ldw = l.ae; dg.w_bs = l.ae;
            dg.dx = dhx; dg.lddx = dhp; dg.dx_bs = l.ae; dg.relu_src = hx; dg.ldrs = hp; dg.rs_bs = l.ae;
            dg.M = M; dg.N = 1; dg.K = l.ae; dg.batch = l.K;
            if ((rc = linear_dgrad(dg, st))) return rc;
            LinearWgrad w2{};
            w2.dy = dws->o3 + l.K; w2.lddy = l.W3; w2.dy_bs = l.N; w2.in = plain_operand(hx + l.K * l.ae, hp, l.ae, l.ae);
            w2.dw = g->w3n; w2.ldw = l.ae; w2.dw_bs = (long long)l.N * l.ae; w2.db = g->b3n; w2.db_bs = l.N;
            w2.M = M; w2.N = l.N; w2.batch = 2 * l.K;
            if ((rc = linear_wgrad(w2, wl))) return rc;
            LinearDgrad d2{};
            d2.dy = dws->o3 + l.K; d2.lddy = l.W3; d2.dy_bs = l.N; d2.w = p->w3n; d2.ldw = l.ae; d2.w_bs = (long long)l.N * l.ae;
            d2.dx = dhx + l.K * l.ae; d2.lddx = dhp; d2.dx_bs = l.ae; d2.relu_src = hx + l.K * l.ae; d2.ldrs = hp;
            d2.rs_bs = l.ae; d2.M = M; d2.N = l.N; d2.K = l.ae; d2.batch = 2 * l.K;
            if ((rc = linear_dgrad(d2, st))) return rc;
        }
        if (l.layers == 3) {   // L2 backward
            LinearWgrad w{};
            w.dy = dws->h2; w.lddy = l.W2; w.dy_bs = l.ae; w.in = plain_operand(ws->h1 + 2 * l.he, l.W1, l.ae, l.ae);
            w.dw = g->w2; w.ldw = l.ae; w.dw_bs = (long long)l.ae * l.ae; w.db = g->b2; w.db_bs = l.ae;
            w.M = M; w.N = l.ae; w.batch = 3 * l.K;
            ForkJoin f2(st, 2);            // (same side stream: behind the weight gradients above, and behind the dh2 the chain just wrote)
            if ((rc = linear_wgrad(w, f2.lane(1)))) return rc;
            LinearDgrad dg{};
            dg.dy = dws->h2; dg.lddy = l.W2; dg.dy_bs = l.ae; dg.w = p->w2; dg.ldw = l.ae; dg.w_bs = (long long)l.ae * l.ae;
            dg.dx = dws->h1 + 2 * l.he; dg.lddx = l.W1; dg.dx_bs = l.ae; dg.relu_src = ws->h1 + 2 * l.he; dg.ldrs = l.W1;
            dg.rs_bs = l.ae; dg.M = M; dg.N = l.ae; dg.K = l.ae; dg.batch = 3 * l.K;
            if ((rc = linear_dgrad(dg, st))) return rc;
        }
        {   // L1a weights
            LinearWgrad w{};
            LinOperand in = plain_operand(s, l.S, l.S);
            in.x2 = actions; in.ldx2 = l.N * l.A; in.K2 = l.N * l.A;
            w.dy = dws->h1 + l.Ws; w.lddy = l.W1; w.in = in;
            w.dw = g->w1a; w.ldw = l.S + l.N * l.A; w.db = g->b1a; w.M = M; w.N = l.K * l.ae; w.batch = 1;
            ForkJoin f3(st, 2);            // dh1 is complete: this one beside the L1s weight gradient below
            if ((rc = linear_wgrad(w, f3.lane(1)))) return rc;
        }
    }
    {   // L1s weights (only the heads that were evaluated)
        LinearWgrad w{};
        w.dy = dws->h1; w.lddy = l.W1; w.in = plain_operand(s, l.S, l.S);
        w.dw = g->w1s; w.ldw = l.S; w.db = g->b1s; w.M = M; w.N = with_adv ? l.Ws : 2 * l.he; w.batch = 1;
        if ((rc = linear_wgrad(w, st))) return rc;
    }
    // (measured and dropped: leaving the lane forked past this call so that the last weight gradients run beside the BPTT
    // recurrence -- a deferred join behind marl_agent_unroll_bwd; 3524 vs 3518 us at config 3, 2895 vs 2877 us with the same
    // change in qtran.cu at config 4: what the products gain the recurrence loses to them)
    fw.join();                             // every lane above is the same side stream: one join covers them
    return MARL_OK;
}

extern "C" int marl_scatter_dq(const marl_dims* d, const float* dq_small, const long long* u, float* dq, void* stream) {
    if (!d || !dq_small || !u || !dq) return MARL_EINVAL;
    const int rows = d->B * d->L * d->N;
    if (rows <= 0) return MARL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    { ProfScope ps_("scatter_dq_kernel", st); scatter_dq_kernel<<<(rows + 255) / 256, 256, 0, st>>>(rows, d->A, dq_small, u, dq); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}
