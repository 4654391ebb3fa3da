// Internal (non-ABI): TMA-fed, warp-specialised, persistent tcgen05 GEMM used by the dense-layer primitives.
// One launch executes a GROUP of independent problems (e.g. the input layers of the three agent unrolls, or the
// data gradient + the weight gradients behind the BPTT kernel), each cut into work items of one 128 x BN output tile
// and one slice of the reduction.
#pragma once
#include <cuda.h>          // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint)
#include "linear.h"

namespace marl {

constexpr int TG_MAXP = 6;              // problems per grouped launch

// everything but the tensor maps: copied to shared memory once per CTA (register-indexed reads of the kernel parameter
// bank, c[0x0][R + off], cost ~200 cycles each and the role loops are full of them otherwise)
struct TGScalars {
    int Md, Nd, Kd;                      // D[Md, Nd] = sum_k A(m,k) B(n,k), k < Kd
    int BN;                              // N tile: a multiple of 32 up to 256
    int a_mn, b_mn;                      // 0: operand stored [rows, k] (K-major), 1: stored [k, rows] (MN-major)
    int m_tiles, n_tiles, k_splits, kb_per_split, kb_total;
    int item0, n_items;                  // work items [item0, item0 + n_items) of the group
    int nmain;                           // main accumulators dealt round-robin (fp32 accumulate truncation, see linear.cu)
    int merge_corr;                      // short reductions: the 2^-11 correction products share the main accumulator
    int epi;                             // 0: y = act(acc + bias) ; 1: dx = acc * (relu_src > 0) ; 2: partial tile to scratch
    int out_tma;                         // epi 0, rows on the UMMA rows, TMA-addressable output: the epilogue leaves by cp.async.bulk.tensor stores
    int transposed;                      // epi 0/1 computed as D^T (features on the 128 UMMA rows, data rows on N): out[n, m]
    float* out; int ldo; int accumulate; int relu;
    const float* bias; float bias_mul;
    const float* relu_src; int ldrs;
    float* partial;                      // epi 2: [k_split][n_tile][m_tile][128][BN]
    // ---- converter hooks (values written into the RAW tile before the hi/lo pass) ----
    int a_row_shift, a_row_period;       // MN-major A: rows come from (k - shift); rows with (k % period) < shift are zero
    int fill_on;                         // composite agent input [obs | last action | agent id | (1)]: 1 = in operand A, 2 = in operand B (K-major)
    int fill_col0, fill_A, fill_N, fill_ones, fill_shift, fill_period, fill_rows;
    const float* fill_onehot;            // [rows, fill_A]
    const float* b_ptr; int ldb;         // MN-major B gathered by the converters (pitch not TMA-able), Nd <= 32
};

struct alignas(64) TGProblem {
    CUtensorMap mapA, mapB;
    CUtensorMap mapOut;                  // output [rows, cols] as 32 x 32 SWIZZLE_128B boxes (TMA store epilogue), when out_tma
    TGScalars s;
};

struct TGGroup {
    TGProblem p[TG_MAXP];
    int n, total_items;
    int stage_bytes, n_stages, b_off;    // smem ring geometry (B tile at b_off inside a stage)
    int cols_per_buf, n_bufs;            // TMEM accumulator buffers
    int debug;                           // MARL_TGEMM_DEBUG bits: 1 no epilogue stores, 2 no convert pass, 4 no MMAs (timing experiments)
};

// deterministic second stage of the split reductions: dw[n, m] += sum_s partial[s][..][m][n] ; db[n] += mul * (ones row)
struct TGReduceJob {
    const float* partial; int k_splits, m_tiles, n_tiles, BN;
    float* dw; int ldw; int K_in, N_out;
    float* db; float db_mul;
};
struct TGReduceGroup { TGReduceJob j[TG_MAXP]; int n; };

bool tgemm_enabled();
// debug phase trace (marl_tgemm_trace): device buffer of 4 roles x 512 words, or nullptr when tracing is off
long long* trace_buffer();
// library-owned view of a caller-provided scratch arena (marl_set_scratch); nullptr when it does not fit
float* tgemm_scratch(size_t bytes);

// Builders: return false when the operands do not qualify (alignment / pitch / shape) -- the caller then uses the
// register-staged kernels of linear.cu.
class TGBuilder {
public:
    TGBuilder();
    bool add_fwd(const LinearFwd& a);
    bool add_dgrad(const LinearDgrad& a);
    bool add_wgrad(const LinearWgrad& a);          // queues the matching reduce job
    bool empty() const { return g_.n == 0; }
    int launch(cudaStream_t st);                   // GEMM group (+ nothing else)
    int launch_reduce(cudaStream_t st);            // the queued reduce jobs of this builder
    void move_reduce_to(TGBuilder& other);         // let another builder's launch_reduce() take this builder's jobs
    bool has_reduce() const { return r_.n > 0; }
private:
    bool push(TGProblem& p);
    TGGroup g_;
    TGReduceGroup r_;
    int max_bn_, max_cols_;
};

}  // namespace marl
