// Dense-layer primitives: forward, data-gradient and weight-gradient GEMMs (FP32 FMA,
// 64x64x16 shared-memory tiles, 4x4 register micro-tiles).  They replace the cuBLAS
// addmm calls behind nn.Linear in the reference (network/q_network.py:17,20;
// network/mixer.py:45-55,117-145,200-206,365-375,399-409) and their autograd duals.
#include "linear.h"
#include "profile.h"

namespace marl {

template <bool A_RED_CONTIG, bool B_RED_CONTIG, class FA, class FB>
__device__ __forceinline__ void gemm_mainloop(TileSmem& s, float (&acc)[TM][TN], FA fa, FB fb,
                                              int i0, int j0, int kbeg, int kend) {
    const int tid = threadIdx.x, ty = tid / (BN / TN), tx = tid % (BN / TN);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int l = 0; l < (BM * BK) / GEMM_THREADS; ++l) {
            int idx = tid + l * GEMM_THREADS;
            int i, kk;
            if (A_RED_CONTIG) { i = idx / BK; kk = idx % BK; } else { i = idx % BM; kk = idx / BM; }
            s.a[kk][i] = (k0 + kk < kend) ? fa(i0 + i, k0 + kk) : 0.0f;
        }
#pragma unroll
        for (int l = 0; l < (BN * BK) / GEMM_THREADS; ++l) {
            int idx = tid + l * GEMM_THREADS;
            int j, kk;
            if (B_RED_CONTIG) { j = idx / BK; kk = idx % BK; } else { j = idx % BN; kk = idx / BN; }
            s.b[kk][j] = (k0 + kk < kend) ? fb(j0 + j, k0 + kk) : 0.0f;
        }
        __syncthreads();
        tile_fma(s, acc, ty, tx);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(GEMM_THREADS) linear_fwd_kernel(LinearFwd a) {
    __shared__ TileSmem s;
    const int z = blockIdx.z, m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int K = lin_width(a.in);
    const float* w = a.w + (long long)z * a.w_bs;
    auto fa = [&](int m, int k) { return m < a.M ? lin_load(a.in, z, m, k) : 0.0f; };
    auto fb = [&](int n, int k) { return n < a.N ? __ldg(w + (long long)n * a.ldw + k) : 0.0f; };
    float acc[TM][TN] = {};
    gemm_mainloop<true, true>(s, acc, fa, fb, m0, n0, 0, K);
    const int ty = threadIdx.x / (BN / TN), tx = threadIdx.x % (BN / TN);
    const float* bias = a.bias ? a.bias + (long long)z * a.b_bs : nullptr;
    float* y = a.y + (long long)z * a.y_bs;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            if (n >= a.N) continue;
            float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.0f);
            if (a.relu) v = fmaxf(v, 0.0f);
            float* dst = y + (long long)m * a.ldy + n;
            *dst = a.accumulate ? (*dst + v) : v;
        }
    }
}

__global__ void __launch_bounds__(GEMM_THREADS) linear_dgrad_kernel(LinearDgrad a) {
    __shared__ TileSmem s;
    const int z = blockIdx.z, m0 = blockIdx.x * BM, k0 = blockIdx.y * BN;
    const float* dy = a.dy + (long long)z * a.dy_bs;
    const float* w = a.w + (long long)z * a.w_bs + a.w_col0;
    auto fa = [&](int m, int n) { return m < a.M ? __ldg(dy + (long long)m * a.lddy + n) : 0.0f; };
    auto fb = [&](int k, int n) { return k < a.K ? __ldg(w + (long long)n * a.ldw + k) : 0.0f; };
    float acc[TM][TN] = {};
    gemm_mainloop<true, false>(s, acc, fa, fb, m0, k0, 0, a.N);
    const int ty = threadIdx.x / (BN / TN), tx = threadIdx.x % (BN / TN);
    float* dx = a.dx + (long long)z * a.dx_bs;
    const float* rs = a.relu_src ? a.relu_src + (long long)z * a.rs_bs : nullptr;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int k = k0 + tx * TN + j;
            if (k >= a.K) continue;
            float v = acc[i][j];
            if (rs && !(__ldg(rs + (long long)m * a.ldrs + k) > 0.0f)) v = 0.0f;
            float* dst = dx + (long long)m * a.lddx + k;
            *dst = a.accumulate ? (*dst + v) : v;
        }
    }
}

__global__ void __launch_bounds__(GEMM_THREADS) linear_wgrad_kernel(LinearWgrad a, int splits, int chunk) {
    __shared__ TileSmem s;
    const int zb = blockIdx.z / splits, sp = blockIdx.z % splits;
    const int i0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
    const int K = lin_width(a.in);
    const int Kout = K + (a.db ? 1 : 0);
    const float* dy = a.dy + (long long)zb * a.dy_bs;
    const int mbeg = sp * chunk, mend = min(a.M, mbeg + chunk);
    if (mbeg >= mend) return;
    auto fa = [&](int i, int m) { return i < a.N ? __ldg(dy + (long long)m * a.lddy + i) : 0.0f; };
    auto fb = [&](int j, int m) { return j < K ? lin_load(a.in, zb, m, j) : (j == K && j < Kout ? 1.0f : 0.0f); };
    float acc[TM][TN] = {};
    gemm_mainloop<false, false>(s, acc, fa, fb, i0, j0, mbeg, mend);
    const int ty = threadIdx.x / (BN / TN), tx = threadIdx.x % (BN / TN);
    float* dw = a.dw + (long long)zb * a.dw_bs;
    float* db = a.db ? a.db + (long long)zb * a.db_bs : nullptr;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int n = i0 + ty * TM + i;
        if (n >= a.N) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int k = j0 + tx * TN + j;
            if (k < K) atomicAdd(dw + (long long)n * a.ldw + k, acc[i][j]);
            else if (k == K && db) atomicAdd(db + n, acc[i][j]);
        }
    }
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

int linear_fwd(const LinearFwd& a, cudaStream_t st) {
    if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return MARL_OK;
    dim3 grid(cdiv(a.M, BM), cdiv(a.N, BN), a.batch);
    { ProfScope ps_("linear_fwd_kernel", st); linear_fwd_kernel<<<grid, GEMM_THREADS, 0, st>>>(a); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

int linear_dgrad(const LinearDgrad& a, cudaStream_t st) {
    if (a.M <= 0 || a.K <= 0 || a.batch <= 0) return MARL_OK;
    dim3 grid(cdiv(a.M, BM), cdiv(a.K, BN), a.batch);
    { ProfScope ps_("linear_dgrad_kernel", st); linear_dgrad_kernel<<<grid, GEMM_THREADS, 0, st>>>(a); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

int linear_wgrad(const LinearWgrad& a, cudaStream_t st) {
    if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return MARL_OK;
    const int Kout = lin_width(a.in) + (a.db ? 1 : 0);
    const int tiles = cdiv(a.N, BM) * cdiv(Kout, BN) * a.batch;
    int splits = cdiv(4 * kNumSMs, tiles);
    splits = max(1, min(splits, cdiv(a.M, 128)));
    int chunk = cdiv(cdiv(a.M, splits), BK) * BK;
    splits = cdiv(a.M, chunk);
    dim3 grid(cdiv(a.N, BM), cdiv(Kout, BN), a.batch * splits);
    { ProfScope ps_("linear_wgrad_kernel", st); linear_wgrad_kernel<<<grid, GEMM_THREADS, 0, st>>>(a, splits, chunk); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

}  // namespace marl
