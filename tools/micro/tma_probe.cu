// Probe: TMA tensor-map loads (SWIZZLE_128B boxes of 32 floats) feeding tcgen05.mma kind::tf32 straight from the
// swizzled tile, K-major and MN-major operands, raw fp32 words as the "hi" part (the tensor core ignores the low 13
// mantissa bits) plus an element-wise "lo" pass in the same layout, and the A operand sourced from TMEM.
// Validates the descriptor conventions used by marl_b200/csrc/tgemm.cu against a CPU reference.
//
//   tma_probe <mode> <N> <K> <split>     mode: kk | km | mm | ts     split: 1 | 3
//     kk: A[128,K] row-major (K-major),   B[N,K] row-major (K-major)        D = A . B^T       (forward)
//     km: A[128,K] row-major (K-major),   B[K,N] row-major (MN-major)       D = A . B         (data gradient)
//     mm: A[K,128] row-major (MN-major),  B[K,N] row-major (MN-major)       D = A^T . B       (weight gradient)
//     ts: like kk, A copied to TMEM first (tcgen05.st) and read from there by the MMA
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// SWIZZLE_128B descriptor: start address, LBO, SBO (>> 4), version 1 (bit 46), layout type 2 (bits 61..63)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;     // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B (MN-major 32-bit operands)
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0; int spins = 0;
    while (!done && spins < 20000000) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        ++spins;
    }
    return done != 0;
}
__device__ __forceinline__ float tf32_rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float tf32_lo(float x) { return tf32_rna(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u)); }

constexpr int M = 128;

struct Params {
    int N, K, split, a_mn, b_mn, ts, variant;
};

// smem: a_raw | a_lo | b_raw | b_lo, each region 1024-byte aligned.
//   K-major operand [ROWS, K]:  K/32 boxes of {32 floats, ROWS}: box kb at kb*ROWS*128 B, row r at r*128 B (swizzled)
//   MN-major operand [K, COLS]: COLS/32 boxes of {32 floats, K rows}: box mb at mb*K*128 B, k row at k*128 B (swizzled)
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                             float* C, Params p, int* err) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base;
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int N = p.N, K = p.K;
    const uint32_t a_bytes = M * K * 4, b_bytes = N * K * 4;
    float* a_raw = reinterpret_cast<float*>(base);
    float* a_lo = reinterpret_cast<float*>(base + a_bytes);
    float* b_raw = reinterpret_cast<float*>(base + 2 * a_bytes);
    float* b_lo = reinterpret_cast<float*>(base + 2 * a_bytes + b_bytes);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_tma)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_tma)), "r"(a_bytes + b_bytes) : "memory");
        if (!p.a_mn) for (int kb = 0; kb < K / 32; ++kb) tma_load_2d(reinterpret_cast<unsigned char*>(a_raw) + kb * M * 128, &mapA, kb * 32, 0, &bar_tma);
        else for (int mb = 0; mb < M / 32; ++mb) tma_load_2d(reinterpret_cast<unsigned char*>(a_raw) + mb * K * 128, &mapA, mb * 32, 0, &bar_tma);
        if (!p.b_mn) for (int kb = 0; kb < K / 32; ++kb) tma_load_2d(reinterpret_cast<unsigned char*>(b_raw) + kb * N * 128, &mapB, kb * 32, 0, &bar_tma);
        else for (int nb = 0; nb < N / 32; ++nb) tma_load_2d(reinterpret_cast<unsigned char*>(b_raw) + nb * K * 128, &mapB, nb * 32, 0, &bar_tma);
    }
    if (!mbar_wait(&bar_tma, 0)) { if (tid == 0) *err = 1; }
    // element-wise lo pass in the tile's own (swizzled) layout
    // (split 4: the raw words are replaced by their rounded TF32 value as well -> unbiased hi, one more store)
    auto pass = [&](float* raw, float* lo, int n4) {
        for (int i = tid; i < n4; i += 128) {
            const float4 v = reinterpret_cast<const float4*>(raw)[i];
            if (p.split == 4) {
                const float4 h = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
                reinterpret_cast<float4*>(raw)[i] = h;
                reinterpret_cast<float4*>(lo)[i] = make_float4(tf32_rna(v.x - h.x), tf32_rna(v.y - h.y), tf32_rna(v.z - h.z), tf32_rna(v.w - h.w));
            } else {
                reinterpret_cast<float4*>(lo)[i] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
            }
        }
    };
    pass(a_raw, a_lo, M * K / 4);
    pass(b_raw, b_lo, N * K / 4);
    // TMEM layout: D at columns [0, N), A (ts mode) raw at [256, 256+K), lo at [256+K, 256+2K)
    const uint32_t tA = tmem + 256, tAlo = tmem + 256 + K;
    if (p.ts) {
        // thread (warp w, lane l) owns TMEM lane 32w+l = row of A; un-swizzle the K-major tile: chunk c of row r sits at c ^ (r & 7)
        const int r = tid;
        for (int k0 = 0; k0 < K; k0 += 8) {
            uint32_t hi[8], lo[8];
            for (int h = 0; h < 2; ++h) {
                const int k = k0 + 4 * h, kb = k / 32, c = (k % 32) / 4;
                const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<unsigned char*>(a_raw) + kb * M * 128 + r * 128 + ((c ^ (r & 7)) << 4));
                hi[4 * h + 0] = __float_as_uint(v.x); hi[4 * h + 1] = __float_as_uint(v.y); hi[4 * h + 2] = __float_as_uint(v.z); hi[4 * h + 3] = __float_as_uint(v.w);
                // (the lo words come from the lo pass above: with split 4 the raw tile already holds the ROUNDED hi values)
                const float4 w = *reinterpret_cast<const float4*>(reinterpret_cast<unsigned char*>(a_lo) + kb * M * 128 + r * 128 + ((c ^ (r & 7)) << 4));
                lo[4 * h + 0] = __float_as_uint(w.x); lo[4 * h + 1] = __float_as_uint(w.y);
                lo[4 * h + 2] = __float_as_uint(w.z); lo[4 * h + 3] = __float_as_uint(w.w);
            }
            const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         ::"r"(tA + lane_base + k0), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         ::"r"(tAlo + lane_base + k0), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = make_idesc(M, N, p.a_mn, p.b_mn);
        uint32_t acc = 0;
        for (int k8 = 0; k8 < K / 8; ++k8) {
            // K-major: box kb = k8/4, +32 B per k8 inside the 128-byte swizzle row; LBO unused (1), SBO = 8 rows = 1024 B
            // MN-major: 8 k rows = 1024 B per k8; LBO = distance between 32-wide MN boxes = K*128 B, SBO = 1024 B
            const uint32_t a_off = p.a_mn ? (uint32_t)k8 * 1024 : (uint32_t)(k8 / 4) * M * 128 + (k8 % 4) * 32;
            const uint32_t b_off = p.b_mn ? (uint32_t)k8 * 1024 : (uint32_t)(k8 / 4) * N * 128 + (k8 % 4) * 32;
            // MN-major 32-bit operands: SWIZZLE_128B_BASE32B, atoms of 4 k rows x 128 B: SBO = 512 B between 4-row groups
            uint32_t a_lbo = p.a_mn ? K * 128 : 16, b_lbo = p.b_mn ? K * 128 : 16;
            uint32_t a_sbo = p.a_mn ? 512 : 1024, b_sbo = p.b_mn ? 512 : 1024;
            if (p.variant == 1) { if (p.a_mn) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; } if (p.b_mn) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; } }
            const uint32_t a_lt = p.a_mn ? 1 : 2, b_lt = p.b_mn ? 1 : 2;
            const uint64_t dah = make_desc(smem_u32(a_raw) + a_off, a_lbo, a_sbo, a_lt), dal = make_desc(smem_u32(a_lo) + a_off, a_lbo, a_sbo, a_lt);
            const uint64_t dbh = make_desc(smem_u32(b_raw) + b_off, b_lbo, b_sbo, b_lt), dbl = make_desc(smem_u32(b_lo) + b_off, b_lbo, b_sbo, b_lt);
            if (p.ts) {
                if (p.split >= 3) {
                    mma_ts(tmem, tAlo + k8 * 8, dbh, idesc, acc); acc = 1;
                    mma_ts(tmem, tA + k8 * 8, dbl, idesc, acc);
                }
                mma_ts(tmem, tA + k8 * 8, dbh, idesc, acc); acc = 1;
            } else {
                if (p.split >= 3) {
                    mma_ss(tmem, dal, dbh, idesc, acc); acc = 1;
                    mma_ss(tmem, dah, dbl, idesc, acc);
                }
                mma_ss(tmem, dah, dbh, idesc, acc); acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    if (!mbar_wait(&bar_mma, 0)) { if (tid == 0) *err = 2; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int row = warp * 32 + (tid & 31);
        for (int j = 0; j < 32 && c0 + j < N; ++j) C[row * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiled enc, float* g, int rows, int cols, int box_rows, bool atom32 = false) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t est[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

static double trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return (double)y; }

int main(int argc, char** argv) {
    if (argc < 5) { printf("usage: tma_probe mode N K split\n"); return 1; }
    const char* mode = argv[1];
    Params p{};
    p.N = atoi(argv[2]); p.K = atoi(argv[3]); p.split = atoi(argv[4]); p.variant = argc > 5 ? atoi(argv[5]) : 0;
    p.a_mn = !strcmp(mode, "mm"); p.b_mn = !strcmp(mode, "mm") || !strcmp(mode, "km"); p.ts = !strcmp(mode, "ts");
    const int N = p.N, K = p.K;
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    EncodeTiled enc = (EncodeTiled)fn;
    // logical A[m][k], B[n][k]; stored row-major as [M,K] (K-major) or [K,M] (MN-major)
    std::vector<float> A(M * K), B(N * K), Ag(M * K), Bg(N * K), C(M * N);
    srand(1);
    for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) Ag[p.a_mn ? k * M + m : m * K + k] = A[m * K + k];
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bg[p.b_mn ? k * N + n : n * K + k] = B[n * K + k];
    float *dA, *dB, *dC; int* derr;
    CK(cudaMalloc(&dA, Ag.size() * 4)); CK(cudaMalloc(&dB, Bg.size() * 4)); CK(cudaMalloc(&dC, C.size() * 4)); CK(cudaMalloc(&derr, 4));
    CK(cudaMemcpy(dA, Ag.data(), Ag.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bg.data(), Bg.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0, C.size() * 4)); CK(cudaMemset(derr, 0, 4));
    CUtensorMap mA = p.a_mn ? make_map(enc, dA, K, M, K, true) : make_map(enc, dA, M, K, M);
    CUtensorMap mB = p.b_mn ? make_map(enc, dB, K, N, K, true) : make_map(enc, dB, N, K, N);
    const size_t smem = (size_t)(2 * M * K + 2 * N * K) * 4 + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe<<<1, 128, smem>>>(mA, mB, dC, p, derr);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    int herr; CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    double worst = 0, worst_t = 0, scale = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0, ref_t = 0;
            for (int k = 0; k < K; ++k) {
                ref += (double)A[m * K + k] * (double)B[n * K + k];
                ref_t += trunc_tf32(A[m * K + k]) * trunc_tf32(B[n * K + k]);
            }
            worst = fmax(worst, fabs(ref - C[m * N + n])); worst_t = fmax(worst_t, fabs(ref_t - C[m * N + n]));
            scale = fmax(scale, fabs(ref));
        }
    printf("mode=%s N=%d K=%d split=%d variant=%d : rel err vs exact %.3e, vs truncated-input product %.3e, timeout=%d  C[0][0..1]=%.5f %.5f C[1][0]=%.5f\n",
           mode, N, K, p.split, p.variant, worst / scale, worst_t / scale, herr, C[0], C[1], C[N]);
    return 0;
}
