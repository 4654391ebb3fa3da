// Dense-layer primitives: forward, data-gradient and weight-gradient GEMMs.
// FP32 FMA, 128x64x16 shared-memory tiles, 8x4 register micro-tiles, 128-bit global loads where the
// operand allows it, and register double-buffering of the next k-tile.  They replace the cuBLAS
// addmm calls behind nn.Linear in the reference (network/q_network.py:17,20;
// network/mixer.py:45-55,117-145,200-206,365-375,399-409) and their autograd duals.
#include "linear.h"
#include "profile.h"

namespace marl {

constexpr int GM = 128, GN = 64, GK = 16, GT = 256;   // CTA tile and thread count
constexpr int LDA = GM + 4, LDB = GN + 4;

struct GemmSmem {
    float a[GK][LDA];
    float b[GK][LDB];
};

// ---- operand fetchers: one float4 (4 consecutive elements along the contiguous dim) per call ---------

// Operand described by a LinOperand: logical matrix [rows, width], contiguous along the column index.
struct OpLin {
    LinOperand o; int z; int rows; int width; bool vec; bool vec2;
    __device__ __forceinline__ float4 quad(int i, int r) const {   // elements (i, r..r+3)
        if (i >= rows) return make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec && r + 3 < o.K1)
            return __ldg(reinterpret_cast<const float4*>(o.x + (long long)z * o.x_bs + (long long)i * o.ldx + r));
        if (vec2 && r >= o.K1 && r + 3 < o.K1 + o.K2) {
            int ii = i;
            if (o.x2_shift) {
                if ((i % o.x2_period) < o.x2_shift) return make_float4(0.f, 0.f, 0.f, 0.f);
                ii = i - o.x2_shift;
            }
            return __ldg(reinterpret_cast<const float4*>(o.x2 + (long long)z * o.x2_bs + (long long)ii * o.ldx2 + (r - o.K1)));
        }
        float4 v;
        v.x = r + 0 < width ? lin_load(o, z, i, r + 0) : 0.f;
        v.y = r + 1 < width ? lin_load(o, z, i, r + 1) : 0.f;
        v.z = r + 2 < width ? lin_load(o, z, i, r + 2) : 0.f;
        v.w = r + 3 < width ? lin_load(o, z, i, r + 3) : 0.f;
        return v;
    }
};

// Plain row-major matrix P[rows, cols] (ld), fetched along its contiguous (column) dimension.
struct OpMat {
    const float* p; int ld; int rows; int cols; bool vec;
    __device__ __forceinline__ float4 quad(int row, int col) const {   // elements (row, col..col+3)
        if (row >= rows) return make_float4(0.f, 0.f, 0.f, 0.f);
        const float* q = p + (long long)row * ld + col;
        if (vec && col + 3 < cols) return __ldg(reinterpret_cast<const float4*>(q));
        float4 v;
        v.x = col + 0 < cols ? __ldg(q + 0) : 0.f;
        v.y = col + 1 < cols ? __ldg(q + 1) : 0.f;
        v.z = col + 2 < cols ? __ldg(q + 2) : 0.f;
        v.w = col + 3 < cols ? __ldg(q + 3) : 0.f;
        return v;
    }
};

// second-source 128-bit loads: x2 aligned, pitches and region starts multiples of 4 floats
__host__ __device__ __forceinline__ bool vec2_ok(const LinOperand& o) {
    return o.K2 >= 4 && o.x2 && (reinterpret_cast<uintptr_t>(o.x2) & 15) == 0 && (o.ldx2 & 3) == 0 &&
           (o.K1 & 3) == 0 && (o.K2 & 3) == 0 && (o.x2_bs & 3) == 0;
}

__device__ __forceinline__ void fma_tile(const GemmSmem& s, float (&acc)[8][4], int ty, int tx) {
#pragma unroll
    for (int k = 0; k < GK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&s.a[k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&s.a[k][64 + ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&s.b[k][tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// Per-thread share of one k-tile: A tile GM x GK = 512 quads (2 per thread), B tile GN x GK = 256 quads (1).
//   RED operand (memory contiguous along the reduction): quad = 4 consecutive reduction elements of
//       row i, scattered to s[r..r+3][i];
//   otherwise: quad = rows i..i+3 at reduction index r, stored with one float4 to s[r][i..i+3].
template <bool A_RED, bool B_RED>
struct TileMap {
    int ai[2], ar[2], bj, br;
    __device__ __forceinline__ TileMap() {
        const int tid = threadIdx.x;
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            const int q = tid + l * GT;
            if (A_RED) { ai[l] = q >> 2; ar[l] = (q & 3) * 4; } else { ai[l] = (q & 31) * 4; ar[l] = q >> 5; }
        }
        if (B_RED) { bj = tid >> 2; br = (tid & 3) * 4; } else { bj = (tid & 15) * 4; br = tid >> 4; }
    }
    __device__ __forceinline__ void stash(GemmSmem& s, const float4 (&ra)[2], const float4& rb) const {
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            if (A_RED) {
                s.a[ar[l] + 0][ai[l]] = ra[l].x; s.a[ar[l] + 1][ai[l]] = ra[l].y;
                s.a[ar[l] + 2][ai[l]] = ra[l].z; s.a[ar[l] + 3][ai[l]] = ra[l].w;
            } else {
                *reinterpret_cast<float4*>(&s.a[ar[l]][ai[l]]) = ra[l];
            }
        }
        if (B_RED) {
            s.b[br + 0][bj] = rb.x; s.b[br + 1][bj] = rb.y; s.b[br + 2][bj] = rb.z; s.b[br + 3][bj] = rb.w;
        } else {
            *reinterpret_cast<float4*>(&s.b[br][bj]) = rb;
        }
    }
};

// Main loop; BIAS additionally accumulates the column sums of the A tile (threads 0..GM-1) into bsum.
template <bool A_RED, bool B_RED, bool BIAS, class FA, class FB>
__device__ __forceinline__ void gemm_loop(GemmSmem& s, float (&acc)[8][4], FA fa, FB fb, int i0, int j0, int rbeg,
                                          int rend, float& bsum) {
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const TileMap<A_RED, B_RED> tm;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 ra[2], rb;
    auto fetch = [&](int r0) {
#pragma unroll
        for (int l = 0; l < 2; ++l) ra[l] = (r0 + tm.ar[l] < rend) ? fa(i0 + tm.ai[l], r0 + tm.ar[l]) : zero4;
        rb = (r0 + tm.br < rend) ? fb(j0 + tm.bj, r0 + tm.br) : zero4;
    };
    fetch(rbeg);
    for (int r0 = rbeg; r0 < rend; r0 += GK) {
        tm.stash(s, ra, rb);
        __syncthreads();
        if (r0 + GK < rend) fetch(r0 + GK);      // next tile's global loads overlap this tile's FMAs
        if (BIAS && tid < GM) {
#pragma unroll
            for (int k = 0; k < GK; ++k) bsum += s.a[k][tid];
        }
        fma_tile(s, acc, ty, tx);
        __syncthreads();
    }
}

// y[M,N] (+)= act(in . w^T + bias)
template <bool VEC_A, bool VEC_B>
__global__ void __launch_bounds__(GT) linear_fwd_kernel(LinearFwd a) {
    __shared__ GemmSmem s;
    const int z = blockIdx.z, m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
    const int K = lin_width(a.in);
    const OpLin A{a.in, z, a.M, K, VEC_A, VEC_A && vec2_ok(a.in)};                                        // rows m, reduction k (contiguous)
    const OpMat B{a.w + (long long)z * a.w_bs, a.ldw, a.N, K, VEC_B};             // rows n, reduction k (contiguous)
    auto fa = [&](int m, int k) { return A.quad(m, k); };
    auto fb = [&](int n, int k) { return B.quad(n, k); };
    float acc[8][4] = {};
    float unused = 0.f;
    gemm_loop<true, true, false>(s, acc, fa, fb, m0, n0, 0, K, unused);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    const float* bias = a.bias ? a.bias + (long long)z * a.b_bs : nullptr;
    float* y = a.y + (long long)z * a.y_bs;
    float bv[4];
    const float bmul = a.bias_mul != 0.f ? a.bias_mul : 1.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { const int n = n0 + tx * 4 + j; bv[j] = (bias && n < a.N) ? bmul * __ldg(bias + n) : 0.f; }
    const bool vec_out = ((a.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && !a.accumulate;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
        if (m >= a.M) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[j] = acc[i][j] + bv[j]; if (a.relu) v[j] = fmaxf(v[j], 0.f); }
        float* dst = y + (long long)m * a.ldy + n0 + tx * 4;
        if (vec_out && n0 + tx * 4 + 3 < a.N) {
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n0 + tx * 4 + j < a.N) dst[j] = a.accumulate ? (dst[j] + v[j]) : v[j];
        }
    }
}

// dx[M,K] (+)= (dy[M,N] . w[N, col0:col0+K]) * (relu_src > 0)
template <bool VEC_A, bool VEC_B>
__global__ void __launch_bounds__(GT) linear_dgrad_kernel(LinearDgrad a) {
    __shared__ GemmSmem s;
    const int z = blockIdx.z, m0 = blockIdx.x * GM, k0 = blockIdx.y * GN;
    const OpMat A{a.dy + (long long)z * a.dy_bs, a.lddy, a.M, a.N, VEC_A};               // rows m, reduction n (contiguous)
    const OpMat B{a.w + (long long)z * a.w_bs + a.w_col0, a.ldw, a.N, a.K, VEC_B};       // rows n (reduction), cols k
    auto fa = [&](int m, int n) { return A.quad(m, n); };
    auto fb = [&](int k, int n) { return B.quad(n, k); };
    float acc[8][4] = {};
    float unused = 0.f;
    gemm_loop<true, false, false>(s, acc, fa, fb, m0, k0, 0, a.N, unused);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float* dx = a.dx + (long long)z * a.dx_bs;
    const float* rs = a.relu_src ? a.relu_src + (long long)z * a.rs_bs : nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k >= a.K) continue;
            float v = acc[i][j];
            if (rs && !(__ldg(rs + (long long)m * a.ldrs + k) > 0.0f)) v = 0.0f;
            float* dst = dx + (long long)m * a.lddx + k;
            *dst = a.accumulate ? (*dst + v) : v;
        }
    }
}

// dw[N, 0:K] += dy^T . in ; db[N] += colsum(dy).  Split over the M rows (blockIdx.z), atomics on the output.
template <bool VEC_A, bool VEC_B>
__global__ void __launch_bounds__(GT) linear_wgrad_kernel(LinearWgrad a, int splits, int chunk) {
    __shared__ GemmSmem s;
    const int zb = blockIdx.z / splits, sp = blockIdx.z % splits;
    const int i0 = blockIdx.x * GM, j0 = blockIdx.y * GN;
    const int K = lin_width(a.in);
    const int mbeg = sp * chunk, mend = min(a.M, mbeg + chunk);
    if (mbeg >= mend) return;
    const OpMat A{a.dy + (long long)zb * a.dy_bs, a.lddy, a.M, a.N, VEC_A};      // rows m (reduction), cols n
    const OpLin B{a.in, zb, a.M, K, VEC_B, VEC_B && vec2_ok(a.in)};                                     // rows m (reduction), cols k
    auto fa = [&](int n, int m) { return A.quad(m, n); };
    auto fb = [&](int k, int m) { return B.quad(m, k); };
    float acc[8][4] = {};
    float bsum = 0.f;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    if (a.db != nullptr && blockIdx.y == 0) {
        gemm_loop<false, false, true>(s, acc, fa, fb, i0, j0, mbeg, mend, bsum);
        if (tid < GM && i0 + tid < a.N)
            atomicAdd(a.db + (long long)zb * a.db_bs + i0 + tid, (a.db_mul != 0.f ? a.db_mul : 1.0f) * bsum);
    } else {
        gemm_loop<false, false, false>(s, acc, fa, fb, i0, j0, mbeg, mend, bsum);
    }
    float* dw = a.dw + (long long)zb * a.dw_bs;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = i0 + (i >> 2) * 64 + ty * 4 + (i & 3);
        if (n >= a.N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = j0 + tx * 4 + j;
            if (k < K) atomicAdd(dw + (long long)n * a.ldw + k, acc[i][j]);
        }
    }
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// 128-bit loads are legal when the base is 16-byte aligned and row pitch / region start are multiples of 4 floats
static bool vec_ok_lin(const LinOperand& o) {
    if (o.K1 == 0) return vec2_ok(o);      // only a second source (e.g. the shifted hidden state)
    return o.K1 >= 4 && o.x && aligned16(o.x) && (o.ldx & 3) == 0 && (o.K1 & 3) == 0 && ((o.x_bs & 3) == 0);
}
static bool vec_ok_mat(const float* p, int ld, long long bs, int col0 = 0) {
    return p && aligned16(p + col0) && (ld & 3) == 0 && ((bs & 3) == 0);
}

#define MARL_DISPATCH2(KERNEL, VA, VB, GRID, ST, ...)                                              \
    do {                                                                                           \
        if (VA && VB) KERNEL<true, true><<<GRID, GT, 0, ST>>>(__VA_ARGS__);                        \
        else if (VA) KERNEL<true, false><<<GRID, GT, 0, ST>>>(__VA_ARGS__);                        \
        else if (VB) KERNEL<false, true><<<GRID, GT, 0, ST>>>(__VA_ARGS__);                        \
        else KERNEL<false, false><<<GRID, GT, 0, ST>>>(__VA_ARGS__);                               \
    } while (0)

int linear_fwd(const LinearFwd& a, cudaStream_t st) {
    if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return MARL_OK;
    dim3 grid(cdiv(a.M, GM), cdiv(a.N, GN), a.batch);
    const bool va = vec_ok_lin(a.in), vb = vec_ok_mat(a.w, a.ldw, a.w_bs);
    { ProfScope ps_("linear_fwd_kernel", st); MARL_DISPATCH2(linear_fwd_kernel, va, vb, grid, st, a); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

int linear_dgrad(const LinearDgrad& a, cudaStream_t st) {
    if (a.M <= 0 || a.K <= 0 || a.batch <= 0) return MARL_OK;
    dim3 grid(cdiv(a.M, GM), cdiv(a.K, GN), a.batch);
    const bool va = vec_ok_mat(a.dy, a.lddy, a.dy_bs), vb = vec_ok_mat(a.w, a.ldw, a.w_bs, a.w_col0);
    { ProfScope ps_("linear_dgrad_kernel", st); MARL_DISPATCH2(linear_dgrad_kernel, va, vb, grid, st, a); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

int linear_wgrad(const LinearWgrad& a, cudaStream_t st) {
    if (a.M <= 0 || a.N <= 0 || a.batch <= 0) return MARL_OK;
    const int K = lin_width(a.in);
    const int tiles = cdiv(a.N, GM) * cdiv(K, GN) * a.batch;
    int splits = cdiv(2 * kNumSMs, tiles);
    splits = max(1, min(splits, cdiv(a.M, 256)));
    int chunk = cdiv(cdiv(a.M, splits), GK) * GK;
    splits = cdiv(a.M, chunk);
    dim3 grid(cdiv(a.N, GM), cdiv(K, GN), a.batch * splits);
    const bool va = vec_ok_mat(a.dy, a.lddy, a.dy_bs), vb = vec_ok_lin(a.in);
    { ProfScope ps_("linear_wgrad_kernel", st); MARL_DISPATCH2(linear_wgrad_kernel, va, vb, grid, st, a, splits, chunk); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

}  // namespace marl
