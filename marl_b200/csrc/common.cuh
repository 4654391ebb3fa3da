// Shared device helpers for libmarl_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MARL_H 64            // rnn_hidden_dim  (common/arguments.py:88)
#define MARL_G (3 * MARL_H)  // GRU gate width (r,z,n)

#define MARL_OK 0
#define MARL_EINVAL (-1)

#define MARL_LAUNCH_CHECK()                                   \
    do {                                                      \
        cudaError_t e__ = cudaGetLastError();                 \
        if (e__ != cudaSuccess) return (int)e__;              \
    } while (0)

namespace marl {

constexpr int kNumSMs = 148;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// ------------------------------------------------------------------------------------------
// Linear-layer operand description shared by the forward / data-gradient / weight-gradient
// GEMM kernels.  The logical input row is the K-wide concatenation
//     [ x (K1 cols) | x2 (K2 cols, optionally shifted by `x2_shift` rows within a period) |
//       one-hot(m % onehot_mod) (onehot_mod cols) ]
// which is how the agent input [obs | last_action | agent_id] (share_params.py:84-112) and the
// QTRAN [hidden | action] / [state | encoding] concatenations are consumed without ever being
// materialised.
// ------------------------------------------------------------------------------------------
struct LinOperand {
    const float* x;   int ldx;  int K1;
    const float* x2;  int ldx2; int K2;
    int x2_shift;     // rows; element (m, K1+k) reads x2[m - shift] and is 0 when (m % x2_period) < shift
    int x2_period;
    int onehot_mod;   // 0 = none
    long long x_bs, x2_bs;   // blockIdx.z strides (batched problems)
};

__device__ __forceinline__ float lin_load(const LinOperand& a, int z, int m, int k) {
    if (k < a.K1) return __ldg(a.x + (long long)z * a.x_bs + (long long)m * a.ldx + k);
    k -= a.K1;
    if (k < a.K2) {
        int mm = m;
        if (a.x2_shift) {
            if ((m % a.x2_period) < a.x2_shift) return 0.0f;
            mm = m - a.x2_shift;
        }
        return __ldg(a.x2 + (long long)z * a.x2_bs + (long long)mm * a.ldx2 + k);
    }
    k -= a.K2;
    return (m % a.onehot_mod) == k ? 1.0f : 0.0f;
}

__host__ __device__ __forceinline__ int lin_width(const LinOperand& a) { return a.K1 + a.K2 + a.onehot_mod; }

}  // namespace marl
