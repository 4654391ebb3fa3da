// Micro-benchmark: cycles per tcgen05.mma kind::tf32 (K = 8) by tile shape, operand source and accumulator dependence.
//   mma_rate           prints a table; operands are zeros in shared memory / TMEM (timing only)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
struct P { int M, N, ts, nacc, reps, b_mn; };
__global__ void __launch_bounds__(128) bench(P p, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bar; __shared__ uint32_t tmem_base;
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(base)[i] = 0.f;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.b_mn << 16) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
        const uint32_t a0 = smem_u32(base), b0 = smem_u32(base + 16384);
        long long t0 = clock64();
        for (int r = 0; r < p.reps; ++r) {
            const int j = r & 3;
            const uint64_t ad = make_desc(a0 + j * 32, 16, 1024, 2);
            const uint64_t bd = p.b_mn ? make_desc(b0 + j * 1024, 4096, 512, 1) : make_desc(b0 + j * 32, 16, 1024, 2);
            const uint32_t d = tmem + (uint32_t)((r % p.nacc) * p.N);
            if (p.ts) mma_ts(d, tmem + 384 + j * 8, bd, idesc, r >= p.nacc); else mma_ss(d, ad, bd, idesc, r >= p.nacc);
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main(int argc, char** argv) {
    const int grid = argc > 1 ? atoi(argv[1]) : 1;
    long long* d; CK(cudaMalloc(&d, 16));
    CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const int reps = 256;
    printf("%4s %4s %3s %5s %5s | issue cyc/MMA | done cyc/MMA | TF32 FLOP/clk/SM\n", "M", "N", "ts", "nacc", "b_mn");
    for (int M : {128})
        for (int N : {32, 64, 96, 128, 192, 256})
            for (int ts = 0; ts < 2; ++ts)
                for (int nacc : {1})
                    for (int b_mn = 0; b_mn < 2; ++b_mn) {
                        if (nacc * N > 384) continue;
                        if (b_mn && (N % 32)) continue;
                        if (ts && M == 128 && (N % 16)) continue;
                        P p{M, N, ts, nacc, reps, b_mn};
                        bench<<<grid, 128, 64 * 1024>>>(p, d);
                        CK(cudaDeviceSynchronize());
                        long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
                        printf("%4d %4d %3d %5d %5d | %13.1f | %12.1f | %8.0f\n", M, N, ts, nacc, b_mn, (double)h[0] / reps, (double)h[1] / reps,
                               2.0 * M * N * 8 * reps / (double)h[1]);
                    }
    return 0;
}
