// Input layers of the recurrent agent for ALL unrolls of a step in one persistent launch:
//     x  = relu(fc1([obs | last_action | agent_id]))      (network/q_network.py:17)
//     gi = W_ih x + b_ih                                  (the input half of the GRUCell, network/q_network.py:19)
// for every (episode, step, agent) row of every stream (eval on o, target on o_next, double-Q eval on o_next:
// controller/share_params.py:125-168).
//
// One CTA per SM walks 128-row tiles.  Roles (17 warps):
//   8 producer warps   global -> registers (two k-tiles ahead) -> hi/lo TF32 split -> shared-memory ring (3 stages)
//   1 MMA warp         GEMM 1: acc1[128 x 64]  = in . W1^T     tcgen05.mma kind::tf32, operands from the ring
//                      GEMM 2: acc2[128 x 192] = x . W_ih^T    A operand read from TENSOR MEMORY, B = W_ih resident in
//                                                              shared memory for the CTA's whole life
//   4 "E" warps        acc1 -> +b1 -> ReLU -> x to global; x split into hi/lo and written back to TMEM as GEMM 2's A
//   4 "F" warps        acc2 -> +b_ih -> gi to global (staged through shared memory: 128-byte coalesced rows)
// so x never makes the round trip through HBM/L2 between the two layers, and the four stages overlap across tiles.
//
// 3xTF32 (see linear.cu).  GEMM 1 streams its reduction, so the two 2^-11 correction products go to their own
// accumulator; GEMM 2 has both operands resident and issues ALL its correction products first: the accumulator is still
// ~2^-11 of its final magnitude while they land, so the fp32-accumulate truncation they add is negligible and the eight
// main products see the same number of updates as with a separate correction accumulator.
#include "front.h"
#include "umma.cuh"
#include "tgemm.h"
#include "profile.h"
#include "../../include/marl_b200.h"

namespace marl {

constexpr int FM = 128, FN1 = MARL_H, FN2 = MARL_G, FK = 16;
constexpr int kFrontStages = 2;
// k-tiles a producer thread keeps in flight in registers.  With two, a phase trace (tools/front_trace.py) showed ~0.9 us per
// k-tile = ~1.8 us of loaded memory latency per fetch: the kernel is bound by bytes in flight, not by conversion or MMAs.
#ifndef MARL_FRONT_DEPTH
#define MARL_FRONT_DEPTH 3
#endif
constexpr int kFrontDepth = MARL_FRONT_DEPTH;
constexpr int kProd = 256;                                   // producer threads
constexpr int kFrontThreads = kProd + 32 + 128 + 128;        // + MMA warp + E warps + F warps
constexpr int kWarpMma = kProd / 32, kWarpE = kWarpMma + 1, kWarpF = kWarpE + 4;
constexpr int FA_PITCH = FM * 4 + 8, FB_PITCH = FN1 * 4 + 8, FW_PITCH = FN2 * 4 + 8;   // canonical K-major layout, see linear.cu
constexpr int kStagePitch = 36;                              // floats per staged row: conflict-free 128-bit stores and loads
constexpr int kStageE = FN1 + 4;                             // same for the [32 rows][64 cols] block of x

// TMEM columns
constexpr uint32_t kColAcc1 = 0, kColCorr1 = 64, kColXhi = 128, kColXlo = 192, kColAcc2 = 256, kFrontTmemCols = 512;

struct FrontStage {
    float a_hi[(FK / 4) * FA_PITCH];
    float a_lo[(FK / 4) * FA_PITCH];
    float b_hi[(FK / 4) * FB_PITCH];
    float b_lo[(FK / 4) * FB_PITCH];
};

struct alignas(128) FrontSmem {
    float wih_hi[(MARL_H / 4) * FW_PITCH];
    float wih_lo[(MARL_H / 4) * FW_PITCH];
    FrontStage st[kFrontStages];
    float stage_e[4][32 * kStageE];
    float stage_f[4][32 * kStagePitch];
    float b1[FN1];
    float bih[FN2];
    FrontSet set;
    FrontStream s[kFrontMaxStreams];
    uint64_t full[kFrontStages], empty[kFrontStages];
    uint64_t acc1_full, acc1_free, xa_ready, acc2_full, acc2_free;
    uint32_t tmem_base;
};
constexpr size_t kFrontSmemBytes = sizeof(FrontSmem) + 128;
static_assert(kFrontSmemBytes <= 227 * 1024, "front kernel shared memory");

constexpr uint32_t front_idesc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(FM >> 4) << 24); }

__device__ __forceinline__ void fmma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void fmma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void fcommit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// (a pointer read from shared memory is generic to the compiler; a generic store would fence the LDS traffic around it)
__device__ __forceinline__ void st_global4(float* p, const float4& v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// the warp's staged [32 rows][32 cols] block -> global rows of 128 contiguous bytes (four rows per instruction)
template <int PITCH>
__device__ __forceinline__ void staged_rows_out(const float* stg, float* out, int ld, int row0, int rows, int col0, const float* bias, bool relu) {
    const int lane = threadIdx.x & 31, c4 = (lane & 7) * 4, rsub = lane >> 3;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) b = *reinterpret_cast<const float4*>(bias + col0 + c4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = 4 * i + rsub;
        float4 t = *reinterpret_cast<const float4*>(stg + r * PITCH + c4);
        t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
        if (relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
        if (row0 + r < rows) st_global4(out + (long long)(row0 + r) * ld + col0 + c4, t);
    }
}

// debug phase trace of CTA 0: (tag, globaltimer ns) pairs, 255 per role (0 producer, 1 MMA, 2 E, 3 F)
#define FT_STAMP(role, tag)                                                                                            \
    do {                                                                                                               \
        if (trace && tn < 255) {                                                                                       \
            long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                       \
            trace[(role) * 512 + 2 * tn] = (tag); trace[(role) * 512 + 2 * tn + 1] = t_; ++tn; trace[(role) * 512 + 510] = tn; \
        }                                                                                                              \
    } while (0)

__global__ void __launch_bounds__(kFrontThreads, 1) agent_front_kernel(const FrontArgs a) {
    extern __shared__ unsigned char front_smem_raw[];
    FrontSmem& sm = *reinterpret_cast<FrontSmem*>((reinterpret_cast<uintptr_t>(front_smem_raw) + 127) & ~(uintptr_t)127);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    long long* trace = (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == kWarpMma || warp == kWarpE || warp == kWarpF)) ? a.trace : nullptr;
    int tn = 0;

    // ---- which parameter set this CTA serves, copied to shared memory (dynamically indexed reads of the kernel
    // parameter bank cost ~200 cycles each)
    int p = 0;
    for (int k = 1; k < a.n_sets; ++k) if ((int)blockIdx.x >= a.set[k].cta0) p = k;
    if (tid == 0) {
        sm.set = a.set[p];
        for (int k = 0; k < kFrontMaxStreams; ++k) sm.s[k] = a.s[k];
#pragma unroll
        for (int s = 0; s < kFrontStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&sm.full[s])), "r"(kProd));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.empty[s])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.acc1_full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(smem_u32(&sm.acc1_free)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(smem_u32(&sm.xa_ready)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.acc2_full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(smem_u32(&sm.acc2_free)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpMma) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kFrontTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    FT_STAMP(warp == 0 ? 0 : warp == kWarpMma ? 1 : warp == kWarpE ? 2 : 3, 9000);
    pdl_wait();                                          // nothing above touched global memory
    FT_STAMP(warp == 0 ? 0 : warp == kWarpMma ? 1 : warp == kWarpE ? 2 : 3, 9001);
    const uint32_t tmem = sm.tmem_base;
    const int I = a.I, rows = a.rows, tps = a.tiles_per_stream;
    const int nk = (I + FK - 1) / FK;
    const int n_ctas = sm.set.n_ctas, cta = (int)blockIdx.x - sm.set.cta0;
    const int set_tiles = sm.set.n_streams * tps;
    const int n_my = cta < set_tiles ? (set_tiles - cta + n_ctas - 1) / n_ctas : 0;

    // ---- resident operands: W_ih as hi / lo TF32 halves in the K-major UMMA layout, the two bias vectors
    {
        const float* w = sm.set.w_ih;
        for (int idx = tid; idx < FN2 * (MARL_H / 4); idx += kFrontThreads) {
            const int n = idx >> 4, c = idx & 15;
            const float4 v = __ldg(reinterpret_cast<const float4*>(w + (long long)n * MARL_H + 4 * c));
            float4 h, l;
            tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y); tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
            *reinterpret_cast<float4*>(sm.wih_hi + c * FW_PITCH + n * 4) = h;
            *reinterpret_cast<float4*>(sm.wih_lo + c * FW_PITCH + n * 4) = l;
        }
        if (tid < FN1) sm.b1[tid] = sm.set.b1 ? __ldg(sm.set.b1 + tid) : 0.f;
        if (tid < FN2) sm.bih[tid] = sm.set.b_ih ? __ldg(sm.set.b_ih + tid) : 0.f;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
    FT_STAMP(warp == 0 ? 0 : warp == kWarpMma ? 1 : warp == kWarpE ? 2 : 3, 9002);

    if (warp < kWarpMma) {
        // =========================================================== producers
        const int ai[2] = {tid >> 2, (tid + kProd) >> 2}, ar = (tid & 3) * 4;        // A quads: rows ai[l], reduction offset ar
        const int bj = tid >> 2;                                                      // B quad: row bj (< 64), same ar
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const OpMat B{sm.set.w1, I, FN1, I, sm.set.vec_w1 != 0};
        const OpMat::Row brow = B.row(bj);
        OpLin A{};
        OpLin::Row arow[2];
        int f_it = -1, f_kt = nk - 1;                    // fetch cursor (tile iteration, k-tile); advanced before each fetch
        float4 ra[kFrontDepth][2], rb[kFrontDepth];
        auto fetch = [&](int set) {
            if (++f_kt == nk) {
                f_kt = 0; ++f_it;
                if (f_it < n_my) {
                    const int t = cta + f_it * n_ctas;
                    const FrontStream& S = sm.s[sm.set.stream[t / tps]];
                    const int m0 = (t % tps) * FM;
                    A = OpLin{S.in, 0, rows, I, S.vec_in != 0, S.vec_in != 0 && vec2_ok(S.in)};
                    arow[0] = A.row(m0 + ai[0]); arow[1] = A.row(m0 + ai[1]);
                }
            }
            if (f_it >= n_my) return;
            const int r = f_kt * FK + ar;
            ra[set][0] = r < I ? A.quad_at(arow[0], r) : zero4;
            ra[set][1] = r < I ? A.quad_at(arow[1], r) : zero4;
            rb[set] = r < I ? B.quad_at(brow, r) : zero4;
        };
        auto put = [&](float* hi, float* lo, int off, const float4& v) {
            float4 h, l;
            tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y); tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
            *reinterpret_cast<float4*>(hi + off) = h;
            *reinterpret_cast<float4*>(lo + off) = l;
        };
        const int total = n_my * nk;
#pragma unroll
        for (int d = 0; d < kFrontDepth; ++d) fetch(d);
        for (int g0 = 0; g0 < total; g0 += kFrontDepth) {
#pragma unroll
            for (int d = 0; d < kFrontDepth; ++d) {                         // (unrolled: the register set index must be static)
                const int g = g0 + d;
                if (g >= total) break;
                const int s = g % kFrontStages, use = g / kFrontStages;
                FrontStage& st = sm.st[s];
                mbar_wait(&sm.empty[s], (uint32_t)((use & 1) ^ 1));      // the MMAs that read this stage last time are done
                FT_STAMP(0, g);
                put(st.a_hi, st.a_lo, (ar >> 2) * FA_PITCH + ai[0] * 4, ra[d][0]);
                put(st.a_hi, st.a_lo, (ar >> 2) * FA_PITCH + ai[1] * 4, ra[d][1]);
                put(st.b_hi, st.b_lo, (ar >> 2) * FB_PITCH + bj * 4, rb[d]);
                FT_STAMP(0, 2000 + g);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
                mbar_arrive(&sm.full[s]);
                FT_STAMP(0, 1000 + g);
                fetch(d);                                                    // k-tile g + kFrontDepth
                FT_STAMP(0, 3000 + g);
            }
        }
    } else if (warp == kWarpMma) {
        // =========================================================== MMA issue (lane 0)
        const uint32_t id1 = front_idesc(FN1), id2 = front_idesc(FN2);
        auto gemm2 = [&](int j) {
            mbar_wait(&sm.xa_ready, (uint32_t)(j & 1));                    // x(j) hi / lo are in TMEM
            mbar_wait(&sm.acc2_free, (uint32_t)((j & 1) ^ 1));             // gi(j-1) has been read out of acc2
            tc_fence_after();
            FT_STAMP(1, 2000 + j);
            if (lane == 0) {
                const uint32_t wh = smem_u32(sm.wih_hi), wl = smem_u32(sm.wih_lo);
#pragma unroll
                for (int k8 = 0; k8 < MARL_H / 8; ++k8) {
                    const uint32_t bo = (uint32_t)(2 * k8) * FW_PITCH * 4;
                    fmma_ts(tmem + kColAcc2, tmem + kColXlo + 8 * k8, umma_desc(wh + bo, FW_PITCH * 4, 128), id2, k8 ? 1u : 0u);
                    fmma_ts(tmem + kColAcc2, tmem + kColXhi + 8 * k8, umma_desc(wl + bo, FW_PITCH * 4, 128), id2, 1u);
                }
#pragma unroll
                for (int k8 = 0; k8 < MARL_H / 8; ++k8) {
                    const uint32_t bo = (uint32_t)(2 * k8) * FW_PITCH * 4;
                    fmma_ts(tmem + kColAcc2, tmem + kColXhi + 8 * k8, umma_desc(wh + bo, FW_PITCH * 4, 128), id2, 1u);
                }
                fcommit(&sm.acc2_full);
            }
            FT_STAMP(1, 3000 + j);
            __syncwarp();
        };
        int g = 0;
        for (int it = 0; it < n_my; ++it) {
            mbar_wait(&sm.acc1_free, (uint32_t)((it & 1) ^ 1));            // x(it-1) has been read out of acc1
            tc_fence_after();
            for (int kt = 0; kt < nk; ++kt, ++g) {
                const int s = g % kFrontStages, use = g / kFrontStages;
                mbar_wait(&sm.full[s], (uint32_t)(use & 1));
                tc_fence_after();
                FT_STAMP(1, g);
                if (lane == 0) {
                    const FrontStage& st = sm.st[s];
#pragma unroll
                    for (int k8 = 0; k8 < FK / 8; ++k8) {
                        const uint32_t ao = (uint32_t)(2 * k8) * FA_PITCH * 4, bo = (uint32_t)(2 * k8) * FB_PITCH * 4;
                        const uint64_t dah = umma_desc(smem_u32(st.a_hi) + ao, FA_PITCH * 4, 128), dal = umma_desc(smem_u32(st.a_lo) + ao, FA_PITCH * 4, 128);
                        const uint64_t dbh = umma_desc(smem_u32(st.b_hi) + bo, FB_PITCH * 4, 128), dbl = umma_desc(smem_u32(st.b_lo) + bo, FB_PITCH * 4, 128);
                        const uint32_t first = (kt | k8) ? 1u : 0u;
                        fmma_ss(tmem + kColCorr1, dal, dbh, id1, first);
                        fmma_ss(tmem + kColCorr1, dah, dbl, id1, 1u);
                        fmma_ss(tmem + kColAcc1, dah, dbh, id1, first);
                    }
                    fcommit(&sm.empty[s]);
                    if (kt == nk - 1) fcommit(&sm.acc1_full);
                }
                __syncwarp();
            }
            gemm2(it);
        }
    } else if (warp < kWarpF) {
        // =========================================================== E: acc1 -> x -> (TMEM hi/lo, global)
        const int q = warp & 3;                                            // TMEM lane quarter this warp may touch
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        float* stg = sm.stage_e[q];
        for (int it = 0; it < n_my; ++it) {
            const int t = cta + it * n_ctas;
            const FrontStream& S = sm.s[sm.set.stream[t / tps]];
            float* xout = S.store_x ? S.x : nullptr;
            const int m0 = (t % tps) * FM;
            mbar_wait(&sm.acc1_full, (uint32_t)(it & 1));
            tc_fence_after();
            FT_STAMP(2, it);
            if (it > 0) mbar_wait(&sm.acc2_full, (uint32_t)((it - 1) & 1));   // GEMM 2 of the previous tile has read x(it-1) from TMEM
            tc_fence_after();
            FT_STAMP(2, 5500 + it);
            // 16 columns at a time (registers: 17 warps are allocated as 20, i.e. 96 per thread): acc1 -> +b1 -> ReLU -> hi / lo
            // back into TMEM; the fp32 values are parked in the warp's staging block and leave for global memory AFTER the
            // hand-over, off the MMA warp's critical path
#pragma unroll 1
            for (int c = 0; c < FN1 / 16; ++c) {
                uint32_t m[16], k[16];
                tmem_ld16(tq + kColAcc1 + 16 * c, m);
                tmem_ld16(tq + kColCorr1 + 16 * c, k);
                tmem_wait_ld();
#pragma unroll
                for (int e4 = 0; e4 < 16; e4 += 4) {
                    float xv[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        xv[e] = fmaxf(__uint_as_float(m[e4 + e]) + __uint_as_float(k[e4 + e]) + sm.b1[16 * c + e4 + e], 0.f);
                        float hh, ll;
                        tf32_split(xv[e], hh, ll);
                        m[e4 + e] = __float_as_uint(hh); k[e4 + e] = __float_as_uint(ll);
                    }
                    if (xout) *reinterpret_cast<float4*>(stg + lane * kStageE + 16 * c + e4) = make_float4(xv[0], xv[1], xv[2], xv[3]);
                }
                tmem_st16(tq + kColXhi + 16 * c, m);
                tmem_st16(tq + kColXlo + 16 * c, k);
            }
            tc_fence_before();
            mbar_arrive(&sm.acc1_free);
            FT_STAMP(2, 6000 + it);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&sm.xa_ready);
            FT_STAMP(2, 1000 + it);
            if (xout) {
                __syncwarp();
                staged_rows_out<kStageE>(stg, xout, FN1, m0 + q * 32, rows, 0, nullptr, false);
                staged_rows_out<kStageE>(stg + 32, xout, FN1, m0 + q * 32, rows, 32, nullptr, false);
                __syncwarp();
            }
            FT_STAMP(2, 2000 + it);
        }
    } else {
        // =========================================================== F: acc2 -> gi
        const int q = warp & 3;
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        float* stg = sm.stage_f[q];
        for (int it = 0; it < n_my; ++it) {
            const int t = cta + it * n_ctas;
            float* gout = sm.s[sm.set.stream[t / tps]].gi;
            const int m0 = (t % tps) * FM;
            mbar_wait(&sm.acc2_full, (uint32_t)(it & 1));
            tc_fence_after();
            FT_STAMP(3, it);
#pragma unroll 1
            for (int c = 0; c < FN2 / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(tq + kColAcc2 + 32 * c, v);
                tmem_wait_ld();
                if (it == 1) FT_STAMP(3, 5000 + c);
                if (c == FN2 / 32 - 1) { tc_fence_before(); mbar_arrive(&sm.acc2_free); }
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4)
                    *reinterpret_cast<float4*>(stg + lane * kStagePitch + 4 * g4) =
                        make_float4(__uint_as_float(v[4 * g4]), __uint_as_float(v[4 * g4 + 1]), __uint_as_float(v[4 * g4 + 2]), __uint_as_float(v[4 * g4 + 3]));
                __syncwarp();
                if (it == 1) FT_STAMP(3, 6000 + c);
                staged_rows_out<kStagePitch>(stg, gout, FN2, m0 + q * 32, rows, 32 * c, sm.bih, false);
                __syncwarp();
                if (it == 1) FT_STAMP(3, 7000 + c);
            }
            FT_STAMP(3, 1000 + it);
        }
    }
    FT_STAMP(warp == 0 ? 0 : warp == kWarpMma ? 1 : warp == kWarpE ? 2 : 3, 9003);
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kFrontTmemCols));
}

bool front_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("MARL_B200_FRONT"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Groups the streams into parameter sets and deals the CTAs of the grid to the sets in proportion to their tiles.
// a.s[0 .. n_streams), a.rows, a.I are filled by the caller, together with one FrontSet per stream in a.set[i]
// (w1 / b1 / w_ih / b_ih of stream i); this merges equal sets.
bool front_plan(FrontArgs& a, int n_streams) {
    if (!front_enabled() || n_streams < 1 || n_streams > kFrontMaxStreams || a.rows <= 0) return false;
    if (a.I <= 0 || a.I > 256) return false;            // GEMM 1 keeps one main accumulator: <= 32 accumulator updates
    for (int i = 0; i < n_streams; ++i) {
        if (!aligned16(a.s[i].gi) || (a.s[i].x && !aligned16(a.s[i].x))) return false;
        if (!aligned16(a.set[i].w_ih)) return false;
        const LinOperand& o = a.s[i].in;
        a.s[i].vec_in = (o.K1 == 0) ? 0 : (o.K1 >= 4 && o.x && aligned16(o.x) && (o.ldx & 3) == 0 && (o.K1 & 3) == 0);
        a.set[i].vec_w1 = aligned16(a.set[i].w1) && (a.I & 3) == 0;
    }
    FrontSet sets[kFrontMaxStreams];
    int ns = 0;
    for (int i = 0; i < n_streams; ++i) {
        int hit = -1;
        for (int k = 0; k < ns; ++k)
            if (sets[k].w1 == a.set[i].w1 && sets[k].b1 == a.set[i].b1 && sets[k].w_ih == a.set[i].w_ih && sets[k].b_ih == a.set[i].b_ih) hit = k;
        if (hit < 0) { sets[ns] = a.set[i]; sets[ns].n_streams = 0; hit = ns++; }
        sets[hit].stream[sets[hit].n_streams++] = i;
    }
    a.tiles_per_stream = (a.rows + FM - 1) / FM;
    const long long total = (long long)n_streams * a.tiles_per_stream;
    const int grid = (int)(total < kNumSMs ? total : kNumSMs);
    int ctas[kFrontMaxStreams];
    for (int k = 0; k < ns; ++k) ctas[k] = 1;
    for (int left = grid - ns; left > 0; --left) {       // the next CTA goes to the set with the most tiles per CTA
        int best = 0;
        for (int k = 1; k < ns; ++k)
            if ((double)sets[k].n_streams / ctas[k] > (double)sets[best].n_streams / ctas[best]) best = k;
        ++ctas[best];
    }
    int c0 = 0;
    for (int k = 0; k < ns; ++k) { sets[k].cta0 = c0; sets[k].n_ctas = ctas[k]; c0 += ctas[k]; a.set[k] = sets[k]; }
    a.n_sets = ns;
    return true;
}

int front_launch(const FrontArgs& a, int prio, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(agent_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFrontSmemBytes); attr_set = true; }
    int grid = 0;
    for (int k = 0; k < a.n_sets; ++k) grid += a.set[k].n_ctas;
    ProfScope ps_("agent_front_kernel", st);
    FrontArgs b = a;
    b.trace = trace_buffer();
    launch_pdl_prio(prio, agent_front_kernel, dim3(grid), dim3(kFrontThreads), kFrontSmemBytes, st, b);
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

}  // namespace marl
