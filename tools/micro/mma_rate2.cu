// Micro-benchmark 2: cycles per tcgen05.mma kind::tf32 (K = 8) with the issue loop stripped to the bare instruction:
// descriptors precomputed, accumulate flag constant, 8 MMAs per loop iteration.  mma_rate.cu's loop rebuilt the descriptors and
// took a modulo per MMA, which bounds a single issuing thread at ~190 cycles whatever the shape.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
struct P { int M, N, mode, nacc, reps; };   // mode 0: SS tf32, 1: TS tf32, 2: SS bf16 (K = 16)
__global__ void __launch_bounds__(128) bench(P p, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bar; __shared__ uint32_t tmem_base;
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<float*>(base)[i] = 0.f;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t fmt = p.mode == 2 ? 1u : 2u;      // bf16 : tf32
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
        const uint32_t a0 = smem_u32(base), b0 = smem_u32(base + 32768);
        uint64_t ad[4], bd[4]; uint32_t at[4], dd[2];
        for (int j = 0; j < 4; ++j) { ad[j] = make_desc(a0 + j * 32, 16, 1024, 2); bd[j] = make_desc(b0 + j * 32, 16, 1024, 2); at[j] = tmem + 448 + j * 8; }
        dd[0] = tmem; dd[1] = tmem + (p.nacc > 1 ? p.N : 0);
        long long t0 = clock64();
        if (p.mode == 0)      for (int r = 0; r < p.reps; r += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) mma_ss(dd[u & 1], ad[u & 3], bd[u & 3], idesc); }
        else if (p.mode == 1) for (int r = 0; r < p.reps; r += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) mma_ts(dd[u & 1], at[u & 3], bd[u & 3], idesc); }
        else                  for (int r = 0; r < p.reps; r += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) mma_f16(dd[u & 1], ad[u & 3], bd[u & 3], idesc); }
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
    CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    const int reps = 512;
    printf("grid %d\n%4s %4s %5s %5s | issue cyc/MMA | done cyc/MMA | FLOP/clk/SM\n", grid, "M", "N", "mode", "nacc");
    for (int M : {64, 128})
        for (int N : {16, 32, 64, 128, 192, 256})
            for (int mode = 0; mode < 3; ++mode)
                for (int nacc : {1, 2}) {
                    if (nacc * N > 448) continue;
                    if (M == 128 && (N % 16)) continue;
                    P p{M, N, mode, nacc, reps};
                    bench<<<grid, 128, 100 * 1024>>>(p, d);
                    CK(cudaDeviceSynchronize());
                    long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
                    const int K = mode == 2 ? 16 : 8;
                    printf("%4d %4d %5s %5d | %13.1f | %12.1f | %8.0f\n", M, N, mode == 0 ? "ss" : mode == 1 ? "ts" : "bf16", nacc, (double)h[0] / reps, (double)h[1] / reps,
                           2.0 * M * N * K * reps / (double)h[1]);
                }
    return 0;
}
