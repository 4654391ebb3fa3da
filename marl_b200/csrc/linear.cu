// Dense-layer primitives: forward, data-gradient and weight-gradient GEMMs.
// 128x64x16 shared-memory tiles on the tensor cores: mma.sync m16n8k8 TF32 with the 3xTF32 split
// (fp32-level accuracy, see mma_tile), fp32 accumulation, 128-bit global loads where the operand allows it,
// register double-buffering of the next k-tile.  They replace the cuBLAS
// addmm calls behind nn.Linear in the reference (network/q_network.py:17,20;
// network/mixer.py:45-55,117-145,200-206,365-375,399-409) and their autograd duals.
#include "linear.h"
#include "profile.h"

namespace marl {

constexpr int GM = 128, GN = 64, GK = 16, GT = 256;   // CTA tile and thread count

// Shared-memory tile of one operand: ROWS output rows/cols x GK reduction elements.
//   RED operand (memory contiguous along the reduction): stored [row][k], pitch GK+4
//   otherwise (contiguous along the output index):        stored [k][row], pitch ROWS+8
// Both pitches make the per-lane fragment reads of mma.m16n8k8 (row = lane/4, k = lane%4) hit 32
// distinct banks, and let the global->shared copy use one 128-bit store per fetched quad.
template <bool RED, int ROWS>
struct OperandTile {
    static constexpr int LD = RED ? (GK + 4) : (ROWS + 8);
    float v[RED ? ROWS * LD : GK * LD];
    __device__ __forceinline__ void store_quad(int i, int r, const float4& q) {
        if (RED) *reinterpret_cast<float4*>(&v[i * LD + r]) = q;      // (i, r..r+3)
        else *reinterpret_cast<float4*>(&v[r * LD + i]) = q;          // (i..i+3, r)
    }
    __device__ __forceinline__ float* quad_ptr(int i, int r) { return RED ? &v[i * LD + r] : &v[r * LD + i]; }
    __device__ __forceinline__ float at(int i, int k) const { return RED ? v[i * LD + k] : v[k * LD + i]; }
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
        // ragged / virtual part: row-level work (shift test, agent id) once per quad
        const float* xr = o.x ? o.x + (long long)z * o.x_bs + (long long)i * o.ldx : nullptr;
        const bool x2_live = o.K2 > 0 && !(o.x2_shift && (i % o.x2_period) < o.x2_shift);
        const float* x2r = x2_live ? o.x2 + (long long)z * o.x2_bs + (long long)(i - o.x2_shift) * o.ldx2 - o.K1 : nullptr;
        const int hot = o.onehot_mod ? o.K1 + o.K2 + (i % o.onehot_mod) : -1;
        float e[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int k = r + c;
            float val = 0.f;
            if (k < o.K1) val = __ldg(xr + k);
            else if (k < o.K1 + o.K2) { if (x2_live) val = __ldg(x2r + k); }
            else if (k == hot) val = 1.0f;
            e[c] = val;
        }
        return make_float4(e[0], e[1], e[2], e[3]);
    }
    // Fills dst (16 B of shared memory) with elements (i, r..r+3): cp.async when the source is a 16-byte
    // aligned run of a real tensor, a plain store otherwise (ragged edge, one-hot columns, zero padding).
    __device__ __forceinline__ void fill(float* dst, int i, int r) const {
        if (i < rows) {
            if (vec && r + 3 < o.K1) {
                cp_async16(dst, o.x + (long long)z * o.x_bs + (long long)i * o.ldx + r);
                return;
            }
            if (vec2 && r >= o.K1 && r + 3 < o.K1 + o.K2 && !(o.x2_shift && (i % o.x2_period) < o.x2_shift)) {
                cp_async16(dst, o.x2 + (long long)z * o.x2_bs + (long long)(i - o.x2_shift) * o.ldx2 + (r - o.K1));
                return;
            }
        }
        *reinterpret_cast<float4*>(dst) = quad(i, r);
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
    __device__ __forceinline__ void fill(float* dst, int row, int col) const {
        if (vec && row < rows && col + 3 < cols) cp_async16(dst, p + (long long)row * ld + col);
        else *reinterpret_cast<float4*>(dst) = quad(row, col);
    }
};

// second-source 128-bit loads: x2 aligned, pitches and region starts multiples of 4 floats
__host__ __device__ __forceinline__ bool vec2_ok(const LinOperand& o) {
    return o.K2 >= 4 && o.x2 && (reinterpret_cast<uintptr_t>(o.x2) & 15) == 0 && (o.ldx2 & 3) == 0 &&
           (o.K1 & 3) == 0 && (o.K2 & 3) == 0 && (o.x2_bs & 3) == 0;
}

__device__ __forceinline__ unsigned to_tf32(float x) { unsigned u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return u; }

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Accumulators of one warp: its 32x32 slice of the 128x64 CTA tile as 2 (m16) x 4 (n8) mma tiles.
struct WarpAcc {
    float c[2][4][4];
    __device__ __forceinline__ WarpAcc() {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) c[i][j][k] = 0.f;
    }
    // g(row_in_tile, col_in_tile, v0, v1): the thread's outputs as 16 pairs of horizontally adjacent values
    template <class G>
    __device__ __forceinline__ void for_each_pair(G g) const {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int r0 = (warp & 3) * 32 + (lane >> 2), c0 = (warp >> 2) * 32 + 2 * (lane & 3);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) g(r0 + mt * 16 + h * 8, c0 + nt * 8, c[mt][nt][2 * h], c[mt][nt][2 * h + 1]);
    }
    // f(row_in_tile, col_in_tile, value) for each of the 32 outputs this thread owns
    template <class F>
    __device__ __forceinline__ void for_each(F f) const {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int r0 = (warp & 3) * 32 + (lane >> 2), c0 = (warp >> 2) * 32 + 2 * (lane & 3);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                f(r0 + mt * 16, c0 + nt * 8, c[mt][nt][0]);
                f(r0 + mt * 16, c0 + nt * 8 + 1, c[mt][nt][1]);
                f(r0 + mt * 16 + 8, c0 + nt * 8, c[mt][nt][2]);
                f(r0 + mt * 16 + 8, c0 + nt * 8 + 1, c[mt][nt][3]);
            }
    }
};

// One k-tile on the tensor cores with the 3xTF32 split (x = hi + lo, both TF32):
//   a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi   -- fp32-level accuracy (the dropped a_lo.b_lo is ~2^-22 relative),
// which is what the 1e-5 parity gate on losses and gradients needs; a single TF32 pass (2^-11) would not do.
template <bool A_RED, bool B_RED>
__device__ __forceinline__ void mma_tile(const OperandTile<A_RED, GM>& A, const OperandTile<B_RED, GN>& B, WarpAcc& acc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ar = (warp & 3) * 32 + (lane >> 2), bn = (warp >> 2) * 32 + (lane >> 2), kq = lane & 3;
#pragma unroll
    for (int k8 = 0; k8 < GK; k8 += 8) {
        unsigned ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const float v[4] = {A.at(ar + mt * 16, k8 + kq), A.at(ar + mt * 16 + 8, k8 + kq),
                                A.at(ar + mt * 16, k8 + kq + 4), A.at(ar + mt * 16 + 8, k8 + kq + 4)};
#pragma unroll
            for (int e = 0; e < 4; ++e) { ah[mt][e] = to_tf32(v[e]); al[mt][e] = to_tf32(v[e] - __uint_as_float(ah[mt][e])); }
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const float v[2] = {B.at(bn + nt * 8, k8 + kq), B.at(bn + nt * 8, k8 + kq + 4)};
#pragma unroll
            for (int e = 0; e < 2; ++e) { bh[nt][e] = to_tf32(v[e]); bl[nt][e] = to_tf32(v[e] - __uint_as_float(bh[nt][e])); }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                mma_tf32(acc.c[mt][nt], al[mt], bh[nt]);
                mma_tf32(acc.c[mt][nt], ah[mt], bl[nt]);
                mma_tf32(acc.c[mt][nt], ah[mt], bh[nt]);
            }
    }
}

// Per-thread share of the global->shared copy of one k-tile: A tile GM x GK = 512 quads (2 per thread),
// B tile GN x GK = 256 quads (1 per thread).  RED: quad = 4 consecutive reduction elements of row i;
// otherwise quad = rows i..i+3 at reduction index r.
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
};

template <bool A_RED, bool B_RED>
struct GemmSmem {
    OperandTile<A_RED, GM> a;
    OperandTile<B_RED, GN> b;
};

constexpr int kStages = 3;   // cp.async ring: two k-tiles in flight while one is on the tensor cores

// Main loop.  FA(dst, i, r) / FB(dst, j, r) fill 16 bytes of shared memory with the operand quad at
// (row i or rows i..i+3, reduction r).  BIAS additionally accumulates the column sums of the (non-RED)
// A tile into bsum (threads 0..GM-1).  One __syncthreads per k-tile.
template <bool A_RED, bool B_RED, bool BIAS, class FA, class FB>
__device__ __forceinline__ void gemm_loop(GemmSmem<A_RED, B_RED>* s, WarpAcc& acc, FA fa, FB fb, int i0, int j0, int rbeg,
                                          int rend, float& bsum) {
    const int tid = threadIdx.x;
    const TileMap<A_RED, B_RED> tm;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nk = (rend - rbeg + GK - 1) / GK;
    auto load = [&](int kt) {
        if (kt < nk) {
            GemmSmem<A_RED, B_RED>& st = s[kt % kStages];
            const int r0 = rbeg + kt * GK;
#pragma unroll
            for (int l = 0; l < 2; ++l) {
                float* dst = st.a.quad_ptr(tm.ai[l], tm.ar[l]);
                if (r0 + tm.ar[l] < rend) fa(dst, i0 + tm.ai[l], r0 + tm.ar[l]);
                else *reinterpret_cast<float4*>(dst) = zero4;
            }
            float* dst = st.b.quad_ptr(tm.bj, tm.br);
            if (r0 + tm.br < rend) fb(dst, j0 + tm.bj, r0 + tm.br);
            else *reinterpret_cast<float4*>(dst) = zero4;
        }
        cp_async_commit();
    };
#pragma unroll
    for (int kt = 0; kt < kStages - 1; ++kt) load(kt);
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<kStages - 2>();            // this thread's copies of tile kt have landed
        __syncthreads();                         // ... everybody's; and everyone is done with tile kt-1
        load(kt + kStages - 1);                  // refill the stage tile kt-1 occupied
        const GemmSmem<A_RED, B_RED>& st = s[kt % kStages];
        if (BIAS && tid < GM) {
#pragma unroll
            for (int k = 0; k < GK; ++k) bsum += st.a.at(tid, k);
        }
        mma_tile<A_RED, B_RED>(st.a, st.b, acc);
    }
    cp_async_wait<0>();
}

// y[M,N] (+)= act(in . w^T + bias)
template <bool VEC_A, bool VEC_B>
__global__ void __launch_bounds__(GT) linear_fwd_kernel(LinearFwd a) {
    __shared__ GemmSmem<true, true> s[kStages];
    const int z = blockIdx.z, m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
    const int K = lin_width(a.in);
    const OpLin A{a.in, z, a.M, K, VEC_A, VEC_A && vec2_ok(a.in)};               // rows m, reduction k (contiguous)
    const OpMat B{a.w + (long long)z * a.w_bs, a.ldw, a.N, K, VEC_B};             // rows n, reduction k (contiguous)
    auto fa = [&](float* dst, int m, int k) { A.fill(dst, m, k); };
    auto fb = [&](float* dst, int n, int k) { B.fill(dst, n, k); };
    WarpAcc acc;
    float unused = 0.f;
    gemm_loop<true, true, false>(s, acc, fa, fb, m0, n0, 0, K, unused);
    const float* bias = a.bias ? a.bias + (long long)z * a.b_bs : nullptr;
    float* y = a.y + (long long)z * a.y_bs;
    const float bmul = a.bias_mul != 0.f ? a.bias_mul : 1.0f;
    const bool pair_ok = ((a.ldy & 1) == 0) && ((reinterpret_cast<uintptr_t>(y) & 7) == 0) && !a.accumulate;
    acc.for_each_pair([&](int i, int j, float v0, float v1) {
        const int m = m0 + i, n = n0 + j;
        if (m >= a.M || n >= a.N) return;
        const bool two = n + 1 < a.N;
        if (bias) { v0 += bmul * __ldg(bias + n); if (two) v1 += bmul * __ldg(bias + n + 1); }
        if (a.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        float* dst = y + (long long)m * a.ldy + n;
        if (pair_ok && two) { *reinterpret_cast<float2*>(dst) = make_float2(v0, v1); return; }
        dst[0] = a.accumulate ? dst[0] + v0 : v0;
        if (two) dst[1] = a.accumulate ? dst[1] + v1 : v1;
    });
}

// dx[M,K] (+)= (dy[M,N] . w[N, col0:col0+K]) * (relu_src > 0)
template <bool VEC_A, bool VEC_B>
__global__ void __launch_bounds__(GT) linear_dgrad_kernel(LinearDgrad a) {
    __shared__ GemmSmem<true, false> s[kStages];
    const int z = blockIdx.z, m0 = blockIdx.x * GM, k0 = blockIdx.y * GN;
    const OpMat A{a.dy + (long long)z * a.dy_bs, a.lddy, a.M, a.N, VEC_A};               // rows m, reduction n (contiguous)
    const OpMat B{a.w + (long long)z * a.w_bs + a.w_col0, a.ldw, a.N, a.K, VEC_B};       // rows n (reduction), cols k
    auto fa = [&](float* dst, int m, int n) { A.fill(dst, m, n); };
    auto fb = [&](float* dst, int k, int n) { B.fill(dst, n, k); };
    WarpAcc acc;
    float unused = 0.f;
    gemm_loop<true, false, false>(s, acc, fa, fb, m0, k0, 0, a.N, unused);
    float* dx = a.dx + (long long)z * a.dx_bs;
    const float* rs = a.relu_src ? a.relu_src + (long long)z * a.rs_bs : nullptr;
    acc.for_each([&](int i, int j, float v) {
        const int m = m0 + i, k = k0 + j;
        if (m < a.M && k < a.K) {
            if (rs && !(__ldg(rs + (long long)m * a.ldrs + k) > 0.0f)) v = 0.0f;
            float* dst = dx + (long long)m * a.lddx + k;
            *dst = a.accumulate ? (*dst + v) : v;
        }
    });
}

// dw[N, 0:K] += dy^T . in ; db[N] += colsum(dy).  Split over the M rows (blockIdx.z), atomics on the output.
template <bool VEC_A, bool VEC_B>
__global__ void __launch_bounds__(GT) linear_wgrad_kernel(LinearWgrad a, int splits, int chunk) {
    __shared__ GemmSmem<false, false> s[kStages];
    const int zb = blockIdx.z / splits, sp = blockIdx.z % splits;
    const int i0 = blockIdx.x * GM, j0 = blockIdx.y * GN;
    const int K = lin_width(a.in);
    const int mbeg = sp * chunk, mend = min(a.M, mbeg + chunk);
    if (mbeg >= mend) return;
    const OpMat A{a.dy + (long long)zb * a.dy_bs, a.lddy, a.M, a.N, VEC_A};      // rows m (reduction), cols n
    const OpLin B{a.in, zb, a.M, K, VEC_B, VEC_B && vec2_ok(a.in)};             // rows m (reduction), cols k
    auto fa = [&](float* dst, int n, int m) { A.fill(dst, m, n); };
    auto fb = [&](float* dst, int k, int m) { B.fill(dst, m, k); };
    WarpAcc acc;
    float bsum = 0.f;
    const int tid = threadIdx.x;
    if (a.db != nullptr && blockIdx.y == 0) {
        gemm_loop<false, false, true>(s, acc, fa, fb, i0, j0, mbeg, mend, bsum);
        if (tid < GM && i0 + tid < a.N)
            atomicAdd(a.db + (long long)zb * a.db_bs + i0 + tid, (a.db_mul != 0.f ? a.db_mul : 1.0f) * bsum);
    } else {
        gemm_loop<false, false, false>(s, acc, fa, fb, i0, j0, mbeg, mend, bsum);
    }
    float* dw = a.dw + (long long)zb * a.dw_bs;
    acc.for_each([&](int i, int j, float v) {
        const int n = i0 + i, k = j0 + j;
        if (n < a.N && k < K) atomicAdd(dw + (long long)n * a.ldw + k, v);
    });
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
