// Micro-benchmark: legacy mma.sync TF32 (m16n8k8) and BF16 (m16n8k16) issue rate on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void tf32_kernel(float* out, int iters) {
    float c[8][4] = {};
    unsigned a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f800000u};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 1.2345f) out[0] = s;
}
__global__ void bf16_kernel(float* out, int iters) {
    float c[8][4] = {};
    unsigned a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f803f80u};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 1.2345f) out[0] = s;
}
int main() {
    float* d; cudaMalloc(&d, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 16; warps *= 2) {
        int iters = 20000, blocks = 148 * 4;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); tf32_kernel<<<blocks, warps * 32>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 16 * 8 * 8 * 8.0 * iters * warps * blocks;
        printf("tf32 m16n8k8  warps/cta %2d: %.1f TFLOP/s\n", warps, fl / ms / 1e9);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); bf16_kernel<<<blocks, warps * 32>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 2.0 * 16 * 8 * 16 * 8.0 * iters * warps * blocks;
        printf("bf16 m16n8k16 warps/cta %2d: %.1f TFLOP/s\n", warps, fl / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
