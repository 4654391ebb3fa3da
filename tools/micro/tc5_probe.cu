// Probe: tcgen05.mma kind::tf32 with hand-built shared-memory descriptors (no swizzle), K-major and MN-major
// operands, 1xTF32 and 3xTF32, accumulators in TMEM, read back with tcgen05.ld.  Validates the descriptor
// conventions used by marl_b200/csrc/linear.cu against a CPU reference.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;     // version = 1 (Blackwell)
    return d;                    // layout_type = 0 (no swizzle), base_offset = 0, lbo_mode = 0
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                 // c_format = F32
    d |= 2u << 7;                 // a_format = TF32
    d |= 2u << 10;                // b_format = TF32
    d |= (uint32_t)a_mn << 15;    // a_major (0 = K-major)
    d |= (uint32_t)b_mn << 16;    // b_major
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr int M = 128;

// smem layouts (floats):  K-major operand  [kc][row][4]  : off(row,k) = (k/4)*ROWS*4 + row*4 + k%4   (LBO = ROWS*16 B, SBO = 128 B)
//                         MN-major operand [mn/4][k][4]  : off(mn,k)  = (mn/4)*KT*4 + k*4 + mn%4     (LBO = 128 B,      SBO = KT*16 B)
__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* C, int N, int K, int three, int b_mn, int* err, int swap) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    float* a_hi = smem;                 // 128 x K
    float* a_lo = a_hi + M * K;
    float* b_hi = a_lo + M * K;         // N x K
    float* b_lo = b_hi + N * K;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * K; i += 128) {
        const int r = i / K, k = i % K;
        const float x = A[r * K + k];
        const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
        const int off = (k / 4) * M * 4 + r * 4 + (k % 4);
        a_hi[off] = x; a_lo[off] = x - hi;
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        const float x = B[n * K + k];
        const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
        const int off = b_mn ? ((n / 4) * K * 4 + k * 4 + (n % 4)) : ((k / 4) * N * 4 + n * 4 + (k % 4));
        b_hi[off] = x; b_lo[off] = x - hi;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(M, N, 0, b_mn);
        uint32_t accumulate = 0;
        for (int k8 = 0; k8 < K / 8; ++k8) {
            const uint32_t a_off = (uint32_t)(2 * k8) * M * 16;                                   // two 16-byte k-chunks per MMA
            const uint32_t b_off = b_mn ? (uint32_t)k8 * 128 : (uint32_t)(2 * k8) * N * 16;
            const uint32_t a_lbo = M * 16, a_sbo = 128;
            uint32_t b_lbo = b_mn ? 128 : N * 16, b_sbo = b_mn ? K * 16 : 128;
            if (swap && b_mn) { uint32_t t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
            const uint64_t dah = make_desc(smem_u32(a_hi) + a_off, a_lbo, a_sbo), dal = make_desc(smem_u32(a_lo) + a_off, a_lbo, a_sbo);
            const uint64_t dbh = make_desc(smem_u32(b_hi) + b_off, b_lbo, b_sbo), dbl = make_desc(smem_u32(b_lo) + b_off, b_lbo, b_sbo);
            if (three) {
                mma_tf32(tmem, dal, dbh, idesc, accumulate); accumulate = 1;
                mma_tf32(tmem, dah, dbl, idesc, accumulate);
            }
            mma_tf32(tmem, dah, dbh, idesc, accumulate); accumulate = 1;
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // wait for the MMAs (bounded spin: never hang the GPU)
    {
        uint32_t done = 0; int spins = 0;
        while (!done && spins < 20000000) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
            ++spins;
        }
        if (!done && tid == 0) *err = 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w owns TMEM lanes 32w..32w+31 (= rows), 32 columns at a time
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
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

int main() {
    const int K = 32;
    int* derr; CK(cudaMalloc(&derr, 4)); CK(cudaMemset(derr, 0, 4));
    for (int N : {64, 32, 8}) {
        std::vector<float> A(M * K), B(N * K), C(M * N);
        srand(1);
        for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
        for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
        float *dA, *dB, *dC;
        CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dC, C.size() * 4));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        const size_t smem = (size_t)(2 * M * K + 2 * N * K) * 4;
        CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int b_mn = 0; b_mn < 3; ++b_mn)
            for (int three = 0; three < 2; ++three) {
                const int swap = b_mn == 2;
                CK(cudaMemset(dC, 0, C.size() * 4));
                probe<<<1, 128, smem>>>(dA, dB, dC, N, K, three, b_mn ? 1 : 0, derr, swap);
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
                int herr; CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
                double worst = 0, scale = 0;
                for (int m = 0; m < M; ++m)
                    for (int n = 0; n < N; ++n) {
                        double ref = 0;
                        for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)B[n * K + k];
                        worst = fmax(worst, fabs(ref - C[m * N + n])); scale = fmax(scale, fabs(ref));
                    }
                printf("N=%3d b_mn_major=%d three=%d : max abs err %.3e (scale %.3f, rel %.3e) timeout=%d   C[0..3]= %.4f %.4f %.4f %.4f | C[1][0]=%.4f\n", N, b_mn, three, worst,
                       scale, worst / scale, herr, C[0], C[1], C[2], C[3], C[N]);
                if (b_mn == 0 && three == 1) { double r0=0,r1=0,r4=0; for (int k=0;k<K;++k){r0+=A[k]*B[k]; r1+=A[k]*B[K+k]; r4+=A[K+k]*B[k];} printf("   ref C[0][0]=%.4f C[0][1]=%.4f C[1][0]=%.4f\n", r0, r1, r4); }
            }
        cudaFree(dA); cudaFree(dB); cudaFree(dC);
    }
    return 0;
}
