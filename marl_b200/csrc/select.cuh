// Per-warp action-value selection of one (b, t) sample, shared by the fused mixer kernels.
// algorithm/q_learner.py:100-117 -- the same arithmetic as q_select_kernel (select_td.cu), arranged for a warp that owns
// the sample: the [N, A] slabs are read coalesced into shared memory (availability mask folded in, q_targets masked in
// place like the reference does), then lane = agent scans its row.
#pragma once
#include "common.cuh"
#include "../../include/marl_b200.h"

#ifndef MIX_SEL_STAMP
#define MIX_SEL_STAMP(tag) do { } while (0)
#endif

namespace marl {

constexpr float kNegBig = -9999999.0f;   // algorithm/q_learner.py:105,112,126

struct SelectArgs {
    float* q; float* qn; float* qt; const float* avail_next; const long long* u;
    float* q_out; float* qt_out; long long* a_star;
    // heads != 0: q / qn / qt are OUTPUTS too, computed here from the agents' hidden states (q = W2 h + b2,
    // network/q_network.py:21-22) -- the three [B*L*N, A, H] head GEMMs of the unroll are then not launched
    int heads;
    const float* h_e; const float* h_t; const float* h_n;          // hidden of eval(o), target(o_next), eval(o_next)
    const float* w2; const float* b2; const float* w2t; const float* b2t;
};

inline SelectArgs select_args(const marl_select_fused* s, const long long* u, float* q_chosen, float* q_tc) {
    SelectArgs a{};
    a.q = s->q_evals; a.qn = s->q_evals_next; a.qt = s->q_targets; a.avail_next = s->avail_u_next; a.u = u;
    a.q_out = q_chosen; a.qt_out = q_tc; a.a_star = s->a_star;
    a.heads = s->hidden_evals != nullptr;
    a.h_e = s->hidden_evals; a.h_t = s->hidden_targets; a.h_n = s->hidden_evals_next;
    a.w2 = s->fc2_w; a.b2 = s->fc2_b; a.w2t = s->fc2_w_target; a.b2t = s->fc2_b_target;
    return a;
}
inline bool select_ok(const marl_select_fused* s) {
    if (!s->q_evals || !s->q_targets || !s->avail_u_next) return false;
    if (!s->hidden_evals) return true;
    if (!s->hidden_targets || !s->fc2_w || !s->fc2_b || !s->fc2_w_target || !s->fc2_b_target) return false;
    if ((s->q_evals_next != nullptr) != (s->hidden_evals_next != nullptr)) return false;
    return (((uintptr_t)s->hidden_evals | (uintptr_t)s->hidden_targets | (uintptr_t)s->hidden_evals_next |
             (uintptr_t)s->fc2_w | (uintptr_t)s->fc2_w_target) & 15) == 0;
}
#ifndef MARL_HEADS_MMA
#define MARL_HEADS_MMA 1
#endif
// x = hi + lo, both rounded to TF32 (integer add + mask = round to nearest, ties away: umma.cuh tf32_rna)
__device__ __forceinline__ void head_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
}
// d[16x8] += a[16x8] . b[8x8]   (row.col, fp32 accumulate; fragment layout of the PTX ISA: lane = 4 g + t holds a(g, t), a(g+8, t),
// a(g, t+4), a(g+8, t+4); b(t, g), b(t+4, g); d(g, 2t), d(g, 2t+1), d(g+8, 2t), d(g+8, 2t+1))
__device__ __forceinline__ void head_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
constexpr int kHeadLd = MARL_H + 4;    // padded row of the staged head weights / hidden slabs (floats)
// floats of one warp's staging area: [N, A] slabs (2, or 3 with heads) rounded up to 16 bytes, then (heads) the sample's
// three hidden slabs [N][kHeadLd]
__host__ __device__ inline size_t select_warp_floats(int N, int A, bool heads) {
    const size_t slabs = (((size_t)(heads ? 3 : 2) * N * A) + 3) & ~(size_t)3;
    return slabs + (heads ? 3 * (size_t)N * kHeadLd : 0);
}
// dynamic shared memory for `warps` warps per CTA (+ both head matrices and biases when heads)
inline size_t select_smem(int warps, int N, int A, bool heads) {
    return ((size_t)warps * select_warp_floats(N, A, heads) + (heads ? 2 * (size_t)A * (kHeadLd + 1) : 0)) * sizeof(float);
}
constexpr size_t kSelectSmemMax = 160 * 1024;

// Block-wide, once per CTA (followed by __syncthreads()): both head matrices + biases into shared memory (read through
// L1 instead, the 2 A distinct rows a warp touches per load made the kernel 20 us slower).
// Layout at `dst` (16-byte aligned): w2 [A][kHeadLd] | w2t [A][kHeadLd] | b2 [A] | b2t [A]
__device__ __forceinline__ void stage_heads(const SelectArgs& s, int A, float* dst) {
    for (int i = threadIdx.x; i < A * MARL_H; i += blockDim.x) {
        const int a = i / MARL_H, k = i - a * MARL_H;
        dst[a * kHeadLd + k] = s.w2[i];
        dst[(A + a) * kHeadLd + k] = s.w2t[i];
    }
    for (int i = threadIdx.x; i < A; i += blockDim.x) {
        dst[2 * A * kHeadLd + i] = s.b2[i];
        dst[2 * A * kHeadLd + A + i] = s.b2t[i];
    }
}

// The sample's three hidden slabs [N][kHeadLd] -> this warp's staging area, asynchronously (cp.async): issued before the CTA stages
// the head matrices so that the two global round trips overlap; warp_select(.., prestaged = true) waits for them.
__device__ __forceinline__ void warp_stage_hidden_async(const SelectArgs& s, long long m, int N, int A, int lane, float* stage) {
    constexpr int Q = MARL_H / 4, LQ = kHeadLd / 4;
    const int NA = N * A;
    float4* she = (float4*)(stage + ((3 * NA + 3) & ~3));
    float4* sht = she + N * LQ;
    float4* shn = sht + N * LQ;
    const float4* ge = (const float4*)(s.h_e + m * N * MARL_H);
    const float4* gt = (const float4*)(s.h_t + m * N * MARL_H);
    const float4* gn = s.h_n ? (const float4*)(s.h_n + m * N * MARL_H) : ge;
    for (int i = lane; i < N * Q; i += 32) {
        const int n = i / Q, k = i - n * Q;
        cp_async16(she + n * LQ + k, ge + i); cp_async16(sht + n * LQ + k, gt + i); cp_async16(shn + n * LQ + k, gn + i);
    }
    cp_async_commit();
}

// stage: 2 * N * A floats of this warp; qc / tc: N floats each of this warp (q_chosen, q_targets_chosen on return)
// heads: the staged head matrices (stage_heads) when s.heads, else unused.
__device__ __forceinline__ void warp_select(const SelectArgs& s, long long m, int N, int A, int lane, float* stage,
                                            float* qc, float* tc, const float* heads, bool prestaged = false) {
    const int NA = N * A;
    const long long o = m * NA;
    float* sn = stage;
    float* st = stage + NA;
    float* se = stage + 2 * NA;                                        // only with heads
    if (s.heads) {
        // the sample's hidden rows first, all loads in flight at once (rows padded: agents land in different banks)
        constexpr int Q = MARL_H / 4, LQ = kHeadLd / 4;
        float4* she = (float4*)(stage + ((3 * NA + 3) & ~3));
        float4* sht = she + N * LQ;
        float4* shn = sht + N * LQ;
        if (!prestaged) warp_stage_hidden_async(s, m, N, A, lane, stage);
#if MARL_HEADS_MMA
        // the availability mask of the sample: its global round trip overlaps the wait for the hidden rows
        float av_pre[2];
#pragma unroll
        for (int it = 0; it < 2; ++it) av_pre[it] = lane + 32 * it < NA ? s.avail_next[o + lane + 32 * it] : 1.0f;
#endif
        cp_async_wait<0>();
        __syncwarp();
        MIX_SEL_STAMP(31);
#if MARL_HEADS_MMA
        // The three head products of the sample, q = W2 h + b2 for its 3 N hidden rows, on the tensor cores (mma.sync m16n8k8,
        // TF32 operands split hi + lo in registers, lo.hi + hi.lo in their own accumulator, hi.hi in the main one: the 3xTF32
        // scheme of linear.cu).  The slabs [e | t | n] are one row-major [3N][kHeadLd] matrix = the A operand (fragment loads
        // hit 32 different banks: row stride 68), a head matrix row = a column of B.  Both weight sets run over every row
        // tile that holds one of their rows; a row keeps the product with its own set.  The raw products land in the [N, A] slabs; a second pass, lane = slab entry, masks them and writes q / qt / qn
        // with coalesced stores.  The FFMA form of this loop (below) was bound by shared-memory wavefronts: 112 128-bit loads per
        // lane and sample (tools/mix_trace.py: 10.5 of the kernel's 32 us).
        {
            const int R3 = 3 * N, g = lane >> 2, t = lane & 3;
            const float* hrows = (const float*)she;
            const float* bias = heads + 2 * A * kHeadLd;
            for (int m0 = 0; m0 < R3; m0 += 16) {
                const int ra = m0 + g, rb = m0 + g + 8;
                const float* pa = hrows + min(ra, R3 - 1) * kHeadLd + t;
                const float* pb = hrows + min(rb, R3 - 1) * kHeadLd + t;
                // rows of the target net: slab 1 = [N, 2N); rows of the online net: the rest
                const bool has_t = m0 < 2 * N && m0 + 16 > N, has_e = m0 < N || m0 + 16 > 2 * N;
                for (int set = 0; set < 2; ++set) {           // (one weight set at a time: both at once is 32 accumulators and spills)
                    if (set ? !has_t : !has_e) continue;
                    for (int c0 = 0; c0 < A; c0 += 16) {
                        const bool two = c0 + 8 < A;
                        const float* w0 = heads + (set * A + min(c0 + g, A - 1)) * kHeadLd + t;
                        const float* w1 = heads + (set * A + min(c0 + 8 + g, A - 1)) * kHeadLd + t;
                        float am[2][4] = {}, ac[2][4] = {};                               // per column tile: main, corrections
#pragma unroll
                        for (int ks = 0; ks < MARL_H / 8; ++ks) {
                            uint32_t ah[4], al[4];
                            head_split(pa[8 * ks], ah[0], al[0]); head_split(pb[8 * ks], ah[1], al[1]);
                            head_split(pa[8 * ks + 4], ah[2], al[2]); head_split(pb[8 * ks + 4], ah[3], al[3]);
#pragma unroll
                            for (int jj = 0; jj < 2; ++jj) {
                                if (jj && !two) continue;
                                const float* w = jj ? w1 : w0;
                                uint32_t bh[2], bl[2];
                                head_split(w[8 * ks], bh[0], bl[0]); head_split(w[8 * ks + 4], bh[1], bl[1]);
                                head_mma(ac[jj], al, bh); head_mma(ac[jj], ah, bl); head_mma(am[jj], ah, bh);
                            }
                        }
                        MIX_SEL_STAMP(33);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int r = h ? rb : ra;
                            if (r >= R3) continue;
                            const int slab = r / N, n = r - slab * N;          // 0: eval(o), 1: target(o_next), 2: eval(o_next)
                            if ((slab == 1) != (set == 1)) continue;
                            float* dst = (slab == 0 ? se : slab == 1 ? st : sn) + n * A;
#pragma unroll
                            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int a = c0 + 8 * jj + 2 * t + e;
                                    if (a < A) dst[a] = (am[jj][2 * h + e] + ac[jj][2 * h + e]) + bias[set * A + a];
                                }
                        }
                    }
                }
            }
            __syncwarp();
            MIX_SEL_STAMP(34);
            for (int i = lane, it = 0; i < NA; i += 32, ++it) {
                const bool off = (it < 2 ? av_pre[it & 1] : s.avail_next[o + i]) == 0.0f;
                s.q[o + i] = se[i];
                const float qt = off ? kNegBig : st[i];                                      // q_learner.py:105
                s.qt[o + i] = qt; st[i] = qt;
                if (s.qn) { const float qn = sn[i]; s.qn[o + i] = qn; sn[i] = off ? kNegBig : qn; }     // :110-112 (the masked copy is not kept)
            }
        }
#else
        // one lane = one agent and TWO actions: the agent's three hidden rows are read once for both, 7 instead of 10 128-bit
        // shared-memory loads per 24 FMAs (the loop is bound by shared-memory bandwidth, tools/mix_trace.py)
        const int A2 = (A + 1) >> 1;
        for (int i = lane; i < N * A2; i += 32) {
            const int n = i / A2, ap = i - n * A2, a0 = 2 * ap, a1 = a0 + 1 < A ? a0 + 1 : a0;
            const float4* w0 = (const float4*)(heads + a0 * kHeadLd);
            const float4* w1 = (const float4*)(heads + a1 * kHeadLd);
            const float4* v0 = (const float4*)(heads + (A + a0) * kHeadLd);
            const float4* v1 = (const float4*)(heads + (A + a1) * kHeadLd);
            const float4* he = she + n * LQ;
            const float4* ht = sht + n * LQ;
            const float4* hn = shn + n * LQ;
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 ce0 = z4, cn0 = z4, ct0 = z4, ce1 = z4, cn1 = z4, ct1 = z4;      // 24 independent FMA chains
#pragma unroll 2
            for (int k = 0; k < Q; ++k) {
                const float4 e4 = he[k], t4 = ht[k], n4 = hn[k];
                const float4 a4 = w0[k], b4 = v0[k];
                ce0.x = fmaf(a4.x, e4.x, ce0.x); ce0.y = fmaf(a4.y, e4.y, ce0.y); ce0.z = fmaf(a4.z, e4.z, ce0.z); ce0.w = fmaf(a4.w, e4.w, ce0.w);
                cn0.x = fmaf(a4.x, n4.x, cn0.x); cn0.y = fmaf(a4.y, n4.y, cn0.y); cn0.z = fmaf(a4.z, n4.z, cn0.z); cn0.w = fmaf(a4.w, n4.w, cn0.w);
                ct0.x = fmaf(b4.x, t4.x, ct0.x); ct0.y = fmaf(b4.y, t4.y, ct0.y); ct0.z = fmaf(b4.z, t4.z, ct0.z); ct0.w = fmaf(b4.w, t4.w, ct0.w);
                const float4 c4 = w1[k], d4 = v1[k];
                ce1.x = fmaf(c4.x, e4.x, ce1.x); ce1.y = fmaf(c4.y, e4.y, ce1.y); ce1.z = fmaf(c4.z, e4.z, ce1.z); ce1.w = fmaf(c4.w, e4.w, ce1.w);
                cn1.x = fmaf(c4.x, n4.x, cn1.x); cn1.y = fmaf(c4.y, n4.y, cn1.y); cn1.z = fmaf(c4.z, n4.z, cn1.z); cn1.w = fmaf(c4.w, n4.w, cn1.w);
                ct1.x = fmaf(d4.x, t4.x, ct1.x); ct1.y = fmaf(d4.y, t4.y, ct1.y); ct1.z = fmaf(d4.z, t4.z, ct1.z); ct1.w = fmaf(d4.w, t4.w, ct1.w);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int a = h ? a1 : a0;
                if (h && a1 == a0) break;                                  // odd A: the last lane of an agent has one action
                const float4 ce = h ? ce1 : ce0, cn = h ? cn1 : cn0, ct = h ? ct1 : ct0;
                float qe = (ce.x + ce.y) + (ce.z + ce.w), qn = (cn.x + cn.y) + (cn.z + cn.w), qt = (ct.x + ct.y) + (ct.z + ct.w);
                const float be = heads[2 * A * kHeadLd + a], bt = heads[2 * A * kHeadLd + A + a];
                qe += be; qn += be; qt += bt;
                const int j = n * A + a;
                const bool off = s.avail_next[o + j] == 0.0f;
                if (off) qt = kNegBig;                                     // q_learner.py:105
                s.q[o + j] = qe; s.qt[o + j] = qt;
                se[j] = qe; st[j] = qt;
                if (s.qn) { s.qn[o + j] = qn; sn[j] = off ? kNegBig : qn; }     // :110-112 (the masked copy is not kept)
            }
        }
#endif
    } else {
        for (int i = lane; i < NA; i += 32) {
            const bool off = s.avail_next[o + i] == 0.0f;
            float v = s.qt[o + i];
            if (off) { v = kNegBig; s.qt[o + i] = v; }                 // in place, q_learner.py:105
            st[i] = v;
            if (s.qn) sn[i] = off ? kNegBig : s.qn[o + i];             // :112
        }
    }
    __syncwarp();
    MIX_SEL_STAMP(32);
    for (int n = lane; n < N; n += 32) {
        const long long i = m * N + n;
        const float c = s.heads ? se[n * A + (int)s.u[i]] : s.q[i * A + s.u[i]];   // :100
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
