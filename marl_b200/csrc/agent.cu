// Shared-parameter recurrent agent (fc1 -> ReLU -> GRUCell -> fc2), unrolled over the episode.
// Replaces the per-timestep Python loop of controller/share_params.py:125-168 around
// network/q_network.py:16-21 and its autograd reversal.
//
//   phase A (all B*L*N rows at once):   x  = relu(fc1([obs | last_action | agent_id]))   (linear_fwd)
//                                       gi = W_ih x + b_ih                               (linear_fwd)
//   phase B (sequential in t):          persistent CTAs own R rows (b,n) each for all L steps;
//                                       a 64x4 thread grid keeps its slice of W_hh in registers,
//                                       h lives in shared memory, one __syncthreads per step.
//   phase C (all rows):                 q = W2 h + b2                                     (linear_fwd)
//
// Independent unrolls (eval on o, target on o_next) run in different CTAs of the same launch;
// dependent ones (the double-Q eval unroll on o_next that starts from the final hidden of the
// eval unroll on o, algorithm/q_learner.py:96,110) are chained inside the same CTA.
#include "linear.h"
#include "../../include/marl_b200.h"
#include "profile.h"

namespace marl {

constexpr int kMaxStreams = 4;
constexpr int kGruThreads = 256;   // 64 hidden units x 4-way split of the reduction

struct GruSegment {
    const float* gi;      // [B,L,N,3H]
    const float* w_hh;    // [3H,H]
    const float* b_hh;    // [3H]
    float* hidden;        // [B,L,N,H]
    float* gates;         // [B,L,N,4H] or null
    float* h_last;        // [B*N,H] or null
};

struct GruFwdArgs {
    GruSegment seg[kMaxStreams];
    int chain_start[kMaxStreams];
    int chain_len[kMaxStreams];
    const float* h0[kMaxStreams];
    int n_chains, B, L, N;
};

template <int R>
__global__ void __launch_bounds__(kGruThreads) gru_unroll_fwd_kernel(GruFwdArgs a) {
    __shared__ __align__(16) float hs[2][R][MARL_H];
    const int tid = threadIdx.x, ks = tid & 3, j = tid >> 2;
    const int chain = blockIdx.y;
    const int rows = a.B * a.N;
    const int row0 = blockIdx.x * R;
    const int my_row = row0 + ks;
    const bool owner = (ks < R) && (my_row < rows);
    const float* h0 = a.h0[chain];
    for (int idx = tid; idx < R * MARL_H; idx += kGruThreads) {
        int rr = idx / MARL_H, jj = idx % MARL_H, row = row0 + rr;
        hs[0][rr][jj] = (h0 && row < rows) ? h0[(long long)row * MARL_H + jj] : 0.0f;
    }
    __syncthreads();
    int cur = 0;
    const int L = a.L, N = a.N;
    const long long base = owner ? ((long long)(my_row / N) * L * N + (my_row % N)) : 0;

    for (int sg = a.chain_start[chain]; sg < a.chain_start[chain] + a.chain_len[chain]; ++sg) {
        const GruSegment S = a.seg[sg];
        // W_hh slice in registers: gate g, unit j, k = 4*(4*i+ks)+c  (float4-interleaved split of the 64-long dot)
        float w[3][16];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 v = __ldg(reinterpret_cast<const float4*>(S.w_hh + (long long)(g * MARL_H + j) * MARL_H + 4 * (4 * i + ks)));
                w[g][4 * i + 0] = v.x; w[g][4 * i + 1] = v.y; w[g][4 * i + 2] = v.z; w[g][4 * i + 3] = v.w;
            }
        const float bh_r = __ldg(S.b_hh + j), bh_z = __ldg(S.b_hh + MARL_H + j), bh_n = __ldg(S.b_hh + 2 * MARL_H + j);
        float gi_r = 0.f, gi_z = 0.f, gi_n = 0.f;
        if (owner) {
            const float* p = S.gi + base * MARL_G;
            gi_r = __ldg(p + j); gi_z = __ldg(p + MARL_H + j); gi_n = __ldg(p + 2 * MARL_H + j);
        }
        for (int t = 0; t < L; ++t) {
            float nx_r = 0.f, nx_z = 0.f, nx_n = 0.f;
            if (owner && t + 1 < L) {   // software prefetch of the next step's input gates
                const float* p = S.gi + (base + (long long)(t + 1) * N) * MARL_G;
                nx_r = __ldg(p + j); nx_z = __ldg(p + MARL_H + j); nx_n = __ldg(p + 2 * MARL_H + j);
            }
            float acc[3][R];
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int rr = 0; rr < R; ++rr) acc[g][rr] = 0.0f;
#pragma unroll
            for (int rr = 0; rr < R; ++rr)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 hv = *reinterpret_cast<const float4*>(&hs[cur][rr][4 * (4 * i + ks)]);
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        acc[g][rr] = fmaf(w[g][4 * i + 0], hv.x, acc[g][rr]);
                        acc[g][rr] = fmaf(w[g][4 * i + 1], hv.y, acc[g][rr]);
                        acc[g][rr] = fmaf(w[g][4 * i + 2], hv.z, acc[g][rr]);
                        acc[g][rr] = fmaf(w[g][4 * i + 3], hv.w, acc[g][rr]);
                    }
                }
            float ar = 0.f, az = 0.f, an = 0.f;
#pragma unroll
            for (int rr = 0; rr < R; ++rr)
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    float v = acc[g][rr];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    if (rr == ks) { if (g == 0) ar = v; else if (g == 1) az = v; else an = v; }
                }
            if (owner) {
                // GRUCell (torch/nn/modules/rnn.py; aten gru_cell): gate order r, z, n
                const float gh_n = an + bh_n;
                const float r = sigmoidf_acc((ar + bh_r) + gi_r);
                const float z = sigmoidf_acc((az + bh_z) + gi_z);
                const float n = tanhf(gi_n + r * gh_n);
                const float hold = hs[cur][ks][j];
                const float hnew = __fadd_rn(__fmul_rn(hold - n, z), n);
                hs[cur ^ 1][ks][j] = hnew;
                const long long idx = base + (long long)t * N;
                S.hidden[idx * MARL_H + j] = hnew;
                if (S.gates) {
                    float* gp = S.gates + idx * (4 * MARL_H);
                    gp[j] = r; gp[MARL_H + j] = z; gp[2 * MARL_H + j] = n; gp[3 * MARL_H + j] = gh_n;
                }
            }
            gi_r = nx_r; gi_z = nx_z; gi_n = nx_n;
            __syncthreads();
            cur ^= 1;
        }
        if (S.h_last && owner) S.h_last[(long long)my_row * MARL_H + j] = hs[cur][ks][j];
    }
}

struct GruBwdArgs {
    const float* gates;    // [B,L,N,4H]
    const float* hidden;   // [B,L,N,H]
    const float* dh_ext;   // [B,L,N,H] or null
    const float* dh_ext2;  // [B,L,N,H] or null (external dL/dhidden)
    const float* w_hh;     // [3H,H]
    const float* h0;       // [B*N,H] or null (zeros)
    float* dgi;            // [B,L,N,3H]
    float* dgh;            // [B,L,N,3H]
    float* dh0;            // [B*N,H] or null
    int B, L, N;
};

template <int R>
__global__ void __launch_bounds__(kGruThreads) gru_unroll_bwd_kernel(GruBwdArgs a) {
    __shared__ __align__(16) float sg[2][R][MARL_G];
    const int tid = threadIdx.x, ks = tid & 3, j = tid >> 2;
    const int rows = a.B * a.N;
    const int row0 = blockIdx.x * R;
    const int my_row = row0 + ks;
    const bool owner = (ks < R) && (my_row < rows);
    const int L = a.L, N = a.N;
    const long long base = owner ? ((long long)(my_row / N) * L * N + (my_row % N)) : 0;
    // column j of W_hh, rows m = 4*(4*i+ks)+c, i = 0..11   (dh_prev[j] += sum_m dgh[m] * W_hh[m, j])
    float wt[48];
#pragma unroll
    for (int i = 0; i < 12; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) wt[4 * i + c] = __ldg(a.w_hh + (long long)(4 * (4 * i + ks) + c) * MARL_H + j);
    for (int idx = tid; idx < 2 * R * MARL_G; idx += kGruThreads) (&sg[0][0][0])[idx] = 0.0f;
    __syncthreads();

    float dh_carry = 0.0f;
    float g_r = 0.f, g_z = 0.f, g_n = 0.f, g_hn = 0.f, hprev = 0.f, dhe = 0.f;
    auto fetch = [&](int t, float& r, float& z, float& n, float& hn, float& hp, float& de) {
        const long long idx = base + (long long)t * N;
        const float* gp = a.gates + idx * (4 * MARL_H);
        r = __ldg(gp + j); z = __ldg(gp + MARL_H + j); n = __ldg(gp + 2 * MARL_H + j); hn = __ldg(gp + 3 * MARL_H + j);
        hp = t > 0 ? __ldg(a.hidden + (idx - N) * MARL_H + j) : (a.h0 ? __ldg(a.h0 + (long long)my_row * MARL_H + j) : 0.0f);
        de = 0.0f;
        if (a.dh_ext) de += __ldg(a.dh_ext + idx * MARL_H + j);
        if (a.dh_ext2) de += __ldg(a.dh_ext2 + idx * MARL_H + j);
    };
    if (owner) fetch(L - 1, g_r, g_z, g_n, g_hn, hprev, dhe);
    int cur = 0;
    for (int t = L - 1; t >= 0; --t) {
        float p_r = 0.f, p_z = 0.f, p_n = 0.f, p_hn = 0.f, p_hp = 0.f, p_de = 0.f;
        if (owner && t > 0) fetch(t - 1, p_r, p_z, p_n, p_hn, p_hp, p_de);
        float dh_dir = 0.0f;
        if (owner) {
            // h' = (h - n) z + n ; n = tanh(gi_n + r (W_hn h + b_hn)) ; r,z = sigmoid(...)
            const float dh = dh_carry + dhe;
            const float dn_pre = dh * (1.0f - g_z) * (1.0f - g_n * g_n);
            const float dz_pre = dh * (hprev - g_n) * g_z * (1.0f - g_z);
            const float dr_pre = dn_pre * g_hn * g_r * (1.0f - g_r);
            const float dghn = dn_pre * g_r;
            dh_dir = dh * g_z;
            const long long idx = base + (long long)t * N;
            float* pi = a.dgi + idx * MARL_G;
            float* ph = a.dgh + idx * MARL_G;
            pi[j] = dr_pre; pi[MARL_H + j] = dz_pre; pi[2 * MARL_H + j] = dn_pre;
            ph[j] = dr_pre; ph[MARL_H + j] = dz_pre; ph[2 * MARL_H + j] = dghn;
            sg[cur][ks][j] = dr_pre; sg[cur][ks][MARL_H + j] = dz_pre; sg[cur][ks][2 * MARL_H + j] = dghn;
        }
        __syncthreads();
        float mine = 0.0f;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            float part = 0.0f;
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                float4 v = *reinterpret_cast<const float4*>(&sg[cur][rr][4 * (4 * i + ks)]);
                part = fmaf(wt[4 * i + 0], v.x, part);
                part = fmaf(wt[4 * i + 1], v.y, part);
                part = fmaf(wt[4 * i + 2], v.z, part);
                part = fmaf(wt[4 * i + 3], v.w, part);
            }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (rr == ks) mine = part;
        }
        dh_carry = dh_dir + mine;
        g_r = p_r; g_z = p_z; g_n = p_n; g_hn = p_hn; hprev = p_hp; dhe = p_de;
        cur ^= 1;
    }
    if (a.dh0 && owner) a.dh0[(long long)my_row * MARL_H + j] = dh_carry;
}

static int pick_rows_per_cta(int rows, int n_chains) {
    // All CTAs must be co-resident (the chains are L steps long): <= 2 CTAs per SM.
    for (int R = 1; R <= 4; R *= 2)
        if (((rows + R - 1) / R) * n_chains <= 2 * kNumSMs) return R;
    return 4;
}

static int check_dims(const marl_dims* d) {
    if (!d || d->B <= 0 || d->L <= 0 || d->N <= 0 || d->A <= 0 || d->O <= 0 || d->S < 0) return MARL_EINVAL;
    return MARL_OK;
}

static LinOperand agent_input(const marl_dims* d, const float* obs, const float* onehot, int shift, int full) {
    if (full) return plain_operand(obs, d->O + d->A + d->N, d->O + d->A + d->N);
    LinOperand in{};
    in.x = obs; in.ldx = d->O; in.K1 = d->O; in.x_bs = 0;
    in.x2 = onehot; in.ldx2 = d->A; in.K2 = d->A; in.x2_bs = 0;
    in.x2_shift = shift ? d->N : 0;
    in.x2_period = d->L * d->N;
    in.onehot_mod = d->N;
    return in;
}

}  // namespace marl

using namespace marl;

extern "C" int marl_agent_unroll_fwd(const marl_dims* d, const marl_unroll_stream* s, int n_streams, void* stream) {
    if (check_dims(d) || !s || n_streams < 1 || n_streams > kMaxStreams) return MARL_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int rows_total = d->B * d->L * d->N;
    const int I = d->O + d->A + d->N;
    for (int i = 0; i < n_streams; ++i) {
        if (!s[i].obs || (!s[i].onehot && !s[i].full_input) || !s[i].hidden || !s[i].x || !s[i].gi) return MARL_EINVAL;
        if (s[i].h0_from >= i) return MARL_EINVAL;
    }
    // phase A
    for (int i = 0; i < n_streams; ++i) {
        LinearFwd f{};
        f.in = agent_input(d, s[i].obs, s[i].onehot, s[i].shift_onehot, s[i].full_input);
        f.w = s[i].params.fc1_w; f.ldw = I; f.bias = s[i].params.fc1_b;
        f.y = s[i].x; f.ldy = MARL_H; f.M = rows_total; f.N = MARL_H; f.relu = 1; f.batch = 1;
        int rc = linear_fwd(f, st);
        if (rc) return rc;
        LinearFwd g{};
        g.in = plain_operand(s[i].x, MARL_H, MARL_H);
        g.w = s[i].params.w_ih; g.ldw = MARL_H; g.bias = s[i].params.b_ih;
        g.y = s[i].gi; g.ldy = MARL_G; g.M = rows_total; g.N = MARL_G; g.batch = 1;
        rc = linear_fwd(g, st);
        if (rc) return rc;
    }
    // phase B: build chains (a stream whose h0_from == j continues chain of j; j must be a chain tail)
    GruFwdArgs ga{};
    ga.B = d->B; ga.L = d->L; ga.N = d->N;
    int order[kMaxStreams], n_ordered = 0;
    bool used[kMaxStreams] = {};
    for (int i = 0; i < n_streams; ++i) {
        if (s[i].h0_from >= 0) continue;
        int c = ga.n_chains++;
        ga.chain_start[c] = n_ordered;
        ga.h0[c] = s[i].h0;
        int cur = i;
        while (cur >= 0) {
            order[n_ordered++] = cur; used[cur] = true;
            int nxt = -1;
            for (int k = cur + 1; k < n_streams; ++k)
                if (!used[k] && s[k].h0_from == cur) { nxt = k; break; }
            cur = nxt;
        }
        ga.chain_len[c] = n_ordered - ga.chain_start[c];
    }
    if (n_ordered != n_streams) return MARL_EINVAL;   // two successors of one stream are not supported
    for (int k = 0; k < n_streams; ++k) {
        const marl_unroll_stream& u = s[order[k]];
        ga.seg[k] = GruSegment{u.gi, u.params.w_hh, u.params.b_hh, u.hidden, u.gates, u.h_last};
    }
    const int rows = d->B * d->N;
    const int R = pick_rows_per_cta(rows, ga.n_chains);
    dim3 grid((rows + R - 1) / R, ga.n_chains);
    if (R == 1) { ProfScope ps_("gru_unroll_fwd_kernel", st); gru_unroll_fwd_kernel<1><<<grid, kGruThreads, 0, st>>>(ga); }
    else if (R == 2) { ProfScope ps_("gru_unroll_fwd_kernel", st); gru_unroll_fwd_kernel<2><<<grid, kGruThreads, 0, st>>>(ga); }
    else { ProfScope ps_("gru_unroll_fwd_kernel", st); gru_unroll_fwd_kernel<4><<<grid, kGruThreads, 0, st>>>(ga); }
    MARL_LAUNCH_CHECK();
    // phase C
    for (int i = 0; i < n_streams; ++i) {
        if (!s[i].q) continue;
        LinearFwd f{};
        f.in = plain_operand(s[i].hidden, MARL_H, MARL_H);
        f.w = s[i].params.fc2_w; f.ldw = MARL_H; f.bias = s[i].params.fc2_b;
        f.y = s[i].q; f.ldy = d->A; f.M = rows_total; f.N = d->A; f.batch = 1;
        int rc = linear_fwd(f, st);
        if (rc) return rc;
    }
    return MARL_OK;
}

extern "C" int marl_agent_unroll_bwd(const marl_dims* d, const marl_unroll_bwd* a, void* stream) {
    if (check_dims(d) || !a || !a->hidden || !a->x || !a->gates || !a->dgi || !a->dgh || !a->dx) return MARL_EINVAL;
    if (a->dq && !a->dhext) return MARL_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const int rows_total = d->B * d->L * d->N;
    const int I = d->O + d->A + d->N;
    int rc;
    if (a->dq) {
        // q = W2 h + b2:  dh_ext = dq . W2 ;  dW2 += dq^T h ; db2 += colsum(dq)
        LinearDgrad g{};
        g.dy = a->dq; g.lddy = d->A; g.w = a->params.fc2_w; g.ldw = MARL_H; g.w_col0 = 0;
        g.dx = a->dhext; g.lddx = MARL_H; g.M = rows_total; g.N = d->A; g.K = MARL_H; g.batch = 1;
        if ((rc = linear_dgrad(g, st))) return rc;
        LinearWgrad w{};
        w.dy = a->dq; w.lddy = d->A; w.in = plain_operand(a->hidden, MARL_H, MARL_H);
        w.dw = a->grads.fc2_w; w.ldw = MARL_H; w.db = a->grads.fc2_b; w.M = rows_total; w.N = d->A; w.batch = 1;
        if ((rc = linear_wgrad(w, st))) return rc;
    }
    GruBwdArgs ga{a->gates, a->hidden, a->dq ? a->dhext : nullptr, a->dhidden, a->params.w_hh, a->h0, a->dgi, a->dgh, a->dh0,
                  d->B, d->L, d->N};
    const int rows = d->B * d->N;
    const int R = pick_rows_per_cta(rows, 1);
    dim3 grid((rows + R - 1) / R);
    if (R == 1) { ProfScope ps_("gru_unroll_bwd_kernel", st); gru_unroll_bwd_kernel<1><<<grid, kGruThreads, 0, st>>>(ga); }
    else if (R == 2) { ProfScope ps_("gru_unroll_bwd_kernel", st); gru_unroll_bwd_kernel<2><<<grid, kGruThreads, 0, st>>>(ga); }
    else { ProfScope ps_("gru_unroll_bwd_kernel", st); gru_unroll_bwd_kernel<4><<<grid, kGruThreads, 0, st>>>(ga); }
    MARL_LAUNCH_CHECK();
    {   // dW_hh += dgh^T . h_{t-1} (hidden shifted by one step, zeros at t = 0) ; db_hh
        LinearWgrad w{};
        w.dy = a->dgh; w.lddy = MARL_G;
        LinOperand in{};
        in.x = nullptr; in.K1 = 0; in.x2 = a->hidden; in.ldx2 = MARL_H; in.K2 = MARL_H;
        in.x2_shift = d->N; in.x2_period = d->L * d->N; in.onehot_mod = 0;
        w.in = in;
        w.dw = a->grads.w_hh; w.ldw = MARL_H; w.db = a->grads.b_hh; w.M = rows_total; w.N = MARL_G; w.batch = 1;
        if ((rc = linear_wgrad(w, st))) return rc;
        if (a->h0) {   // the t = 0 rows see h0 instead of zeros: one [N x H] problem per episode
            LinearWgrad w0{};
            w0.dy = a->dgh; w0.lddy = MARL_G; w0.dy_bs = (long long)d->L * d->N * MARL_G;
            w0.in = plain_operand(a->h0, MARL_H, MARL_H, (long long)d->N * MARL_H);
            w0.dw = a->grads.w_hh; w0.ldw = MARL_H; w0.db = nullptr; w0.M = d->N; w0.N = MARL_G; w0.batch = d->B;
            if ((rc = linear_wgrad(w0, st))) return rc;
        }
    }
    {   // dW_ih += dgi^T . x ; db_ih
        LinearWgrad w{};
        w.dy = a->dgi; w.lddy = MARL_G; w.in = plain_operand(a->x, MARL_H, MARL_H);
        w.dw = a->grads.w_ih; w.ldw = MARL_H; w.db = a->grads.b_ih; w.M = rows_total; w.N = MARL_G; w.batch = 1;
        if ((rc = linear_wgrad(w, st))) return rc;
    }
    {   // dx = (dgi . W_ih) * (x > 0)
        LinearDgrad g{};
        g.dy = a->dgi; g.lddy = MARL_G; g.w = a->params.w_ih; g.ldw = MARL_H; g.w_col0 = 0;
        g.dx = a->dx; g.lddx = MARL_H; g.relu_src = a->x; g.ldrs = MARL_H;
        g.M = rows_total; g.N = MARL_G; g.K = MARL_H; g.batch = 1;
        if ((rc = linear_dgrad(g, st))) return rc;
    }
    {   // dW1 += dx^T . [obs | last_action | agent_id] ; db1
        LinearWgrad w{};
        w.dy = a->dx; w.lddy = MARL_H; w.in = agent_input(d, a->obs, a->onehot, a->shift_onehot, a->full_input);
        w.dw = a->grads.fc1_w; w.ldw = I; w.db = a->grads.fc1_b; w.M = rows_total; w.N = MARL_H; w.batch = 1;
        if ((rc = linear_wgrad(w, st))) return rc;
    }
    return MARL_OK;
}
