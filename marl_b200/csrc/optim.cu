// Global-norm gradient clipping fused with the RMSprop / Adam update over ONE flat parameter
// buffer.  Replaces th.nn.utils.clip_grad_norm_ + optimizer.step() (algorithm/q_learner.py:172-173),
// i.e. torch/nn/utils/clip_grad.py and torch/optim/{rmsprop,adam}.py with default hyper-parameters.
// Two launches, deterministic: (1) per-block partial sums of squares, (2) every block reduces the
// partials in the same fixed order, then updates its slice.
#include "common.cuh"
#include "../../include/marl_b200.h"
#include "profile.h"
#include <cstring>
#include <cooperative_groups.h>

namespace marl {

constexpr int kOptBlocks = kNumSMs;
constexpr int kOptThreads = 256;

__global__ void __launch_bounds__(kOptThreads) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ partials,
                                                            int* step_counter) {
    __shared__ float sw[kOptThreads / 32];
    if (step_counter && blockIdx.x == 0 && threadIdx.x == 0) *step_counter += 1;   // graph-replayable Adam step count
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * kOptThreads + threadIdx.x; i < n; i += (long long)kOptBlocks * kOptThreads) {
        const float v = g[i];
        acc = fmaf(v, v, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < kOptThreads / 32; ++w) t += sw[w];
        partials[blockIdx.x] = t;
    }
}

// returns scale = (1/mask_sum) * min(1, max_norm/(total_norm+1e-6)); writes loss/norm from block 0
__device__ __forceinline__ float clip_scale(const float* partials, const float* scalars, float max_norm, float* loss_out) {
    __shared__ float s_scale;
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int b = 0; b < kOptBlocks; ++b) t += partials[b];
        const float inv = 1.0f / scalars[1];
        const float total_norm = sqrtf(t) * inv;
        float coef = max_norm / (total_norm + 1e-6f);
        coef = coef > 1.0f ? 1.0f : coef;              // torch.clamp(clip_coef, max=1.0)
        s_scale = inv * coef;
        if (blockIdx.x == 0 && loss_out) { loss_out[0] = scalars[0] * inv; loss_out[1] = total_norm; }
    }
    __syncthreads();
    return s_scale;
}

__global__ void __launch_bounds__(kOptThreads) clip_rmsprop_kernel(float* __restrict__ p, float* __restrict__ g,
                                                                   float* __restrict__ sq, long long n,
                                                                   const float* __restrict__ partials,
                                                                   const float* __restrict__ scalars, float max_norm,
                                                                   float lr, float alpha, float eps, float* loss_out) {
    const float scale = clip_scale(partials, scalars, max_norm, loss_out);
    for (long long i = (long long)blockIdx.x * kOptThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kOptThreads) {
        const float gr = g[i] * scale;
        g[i] = gr;
        const float v = alpha * sq[i] + (1.0f - alpha) * gr * gr;   // square_avg.mul_(alpha).addcmul_(g, g, 1-alpha)
        sq[i] = v;
        p[i] = p[i] - lr * (gr / (sqrtf(v) + eps));                 // p.addcdiv_(g, sqrt(v)+eps, -lr)
    }
}

__global__ void __launch_bounds__(kOptThreads) clip_adam_kernel(float* __restrict__ p, float* __restrict__ g,
                                                                float* __restrict__ m1, float* __restrict__ m2, long long n,
                                                                const float* __restrict__ partials,
                                                                const float* __restrict__ scalars, float max_norm,
                                                                float lr, float beta1, float beta2, float eps, int step,
                                                                const int* step_counter, float* loss_out) {
    const float scale = clip_scale(partials, scalars, max_norm, loss_out);
    // bias corrections in double, as torch does with python floats (torch/optim/adam.py)
    const int stp = step_counter ? *step_counter : step;
    const float step_size = (float)((double)lr / (1.0 - pow((double)beta1, (double)stp)));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)stp));
    for (long long i = (long long)blockIdx.x * kOptThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kOptThreads) {
        const float gr = g[i] * scale;
        g[i] = gr;
        const float a = m1[i] + (1.0f - beta1) * (gr - m1[i]);      // exp_avg.lerp_(grad, 1-beta1)
        const float b = beta2 * m2[i] + (1.0f - beta2) * gr * gr;   // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1-beta2)
        m1[i] = a; m2[i] = b;
        const float denom = sqrtf(b) / bc2_sqrt + eps;
        p[i] = p[i] - step_size * (a / denom);
    }
}

// Small parameter sets (<= kFusedMax): one CTA does norm + clip + update in a single launch.
constexpr long long kFusedMax = 4096;   // one CTA only pays off for tiny parameter sets (latency-bound loop)
constexpr int kFusedThreads = 1024;

template <bool ADAM>
__global__ void __launch_bounds__(kFusedThreads) clip_step_fused_kernel(float* __restrict__ p, float* __restrict__ g,
                                                                       float* __restrict__ m1, float* __restrict__ m2, long long n,
                                                                       const float* __restrict__ scalars, float max_norm, float lr,
                                                                       float c1, float c2, float eps, int step, int* step_counter,
                                                                       float* loss_out) {
    __shared__ float sw[kFusedThreads / 32];
    __shared__ float s_scale;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < n; i += kFusedThreads) { const float v = g[i]; acc = fmaf(v, v, acc); }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < kFusedThreads / 32; ++w) t += sw[w];
        const float inv = 1.0f / scalars[1];
        const float total_norm = sqrtf(t) * inv;
        float coef = max_norm / (total_norm + 1e-6f);
        coef = coef > 1.0f ? 1.0f : coef;
        s_scale = inv * coef;
        if (loss_out) { loss_out[0] = scalars[0] * inv; loss_out[1] = total_norm; }
        if (ADAM && step_counter) *step_counter += 1;
    }
    __syncthreads();
    const float scale = s_scale;
    float step_size = 0.f, bc2_sqrt = 1.f;
    if (ADAM) {
        const int stp = step_counter ? *step_counter : step;
        step_size = (float)((double)lr / (1.0 - pow((double)c1, (double)stp)));
        bc2_sqrt = (float)sqrt(1.0 - pow((double)c2, (double)stp));
    }
    for (long long i = threadIdx.x; i < n; i += kFusedThreads) {
        const float gr = g[i] * scale;
        g[i] = gr;
        if (ADAM) {
            const float a = m1[i] + (1.0f - c1) * (gr - m1[i]);
            const float b = c2 * m2[i] + (1.0f - c2) * gr * gr;
            m1[i] = a; m2[i] = b;
            p[i] = p[i] - step_size * (a / (sqrtf(b) / bc2_sqrt + eps));
        } else {
            const float v = c1 * m1[i] + (1.0f - c1) * gr * gr;
            m1[i] = v;
            p[i] = p[i] - lr * (gr / (sqrtf(v) + eps));
        }
    }
}

// Mid-size parameter sets (cfg 2: 62,894 floats): ONE launch of one thread-block cluster.  Each of the 8 CTAs reduces
// its slice of the squared gradient, the partial sums are exchanged through distributed shared memory (every CTA
// reads the 8 partials in the same order, so the norm is bit-identical in all of them and deterministic), then each
// CTA clips and updates its slice.  Replaces the sumsq + clip/update pair of launches on the step's critical path.
constexpr long long kClusterMax = 1 << 18;
constexpr int kClusterCtas = 8, kClusterThreads = 1024;

template <bool ADAM>
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads)
clip_step_cluster_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m1, float* __restrict__ m2, long long n,
                         const float* __restrict__ scalars, float max_norm, float lr, float c1, float c2, float eps, int step,
                         int* step_counter, float* loss_out) {
    pdl_enter();
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float sw[kClusterThreads / 32];
    __shared__ float s_part, s_scale;
    const unsigned rank = cluster.block_rank();
    // 128-bit accesses over the first n4 quads (all four buffers come from 16-byte aligned flat allocations), the
    // <= 3 tail elements are handled by the first threads of CTA 0
    const long long n4 = n >> 2, q0 = (long long)rank * kClusterThreads + threadIdx.x, qstride = (long long)kClusterCtas * kClusterThreads;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float acc = 0.f;
#pragma unroll 2
    for (long long q = q0; q < n4; q += qstride) {
        const float4 v = g4[q];
        acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
    const long long tail = (n4 << 2) + threadIdx.x;
    const bool has_tail = rank == 0 && tail < n;
    if (has_tail) { const float v = g[tail]; acc = fmaf(v, v, acc); }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < kClusterThreads / 32; ++w) t += sw[w];
        s_part = t;
        if (ADAM && step_counter && rank == 0) *step_counter += 1;      // graph-replayable Adam step count
    }
    cluster.sync();                                  // partials (and the step count) visible cluster-wide
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (unsigned r = 0; r < (unsigned)kClusterCtas; ++r) t += *cluster.map_shared_rank(&s_part, r);
        const float inv = 1.0f / scalars[1];
        const float total_norm = sqrtf(t) * inv;
        float coef = max_norm / (total_norm + 1e-6f);
        coef = coef > 1.0f ? 1.0f : coef;              // torch.clamp(clip_coef, max=1.0)
        s_scale = inv * coef;
        if (rank == 0 && loss_out) { loss_out[0] = scalars[0] * inv; loss_out[1] = total_norm; }
    }
    cluster.sync();                                  // nobody leaves (or overwrites s_part) while a peer still reads it
    const float scale = s_scale;
    float step_size = 0.f, bc2_sqrt = 1.f;
    if (ADAM) {
        const int stp = step_counter ? *step_counter : step;
        step_size = (float)((double)lr / (1.0 - pow((double)c1, (double)stp)));
        bc2_sqrt = (float)sqrt(1.0 - pow((double)c2, (double)stp));
    }
    auto update = [&](float& pv, float& gv, float& av, float& bv) {
        const float gr = gv * scale;
        gv = gr;
        if (ADAM) {
            const float a = av + (1.0f - c1) * (gr - av);              // exp_avg.lerp_(grad, 1-beta1)
            const float b = c2 * bv + (1.0f - c2) * gr * gr;           // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1-beta2)
            av = a; bv = b;
            pv = pv - step_size * (a / (sqrtf(b) / bc2_sqrt + eps));
        } else {
            const float v = c1 * av + (1.0f - c1) * gr * gr;           // square_avg.mul_(alpha).addcmul_(g, g, 1-alpha)
            av = v;
            pv = pv - lr * (gr / (sqrtf(v) + eps));                    // p.addcdiv_(g, sqrt(v)+eps, -lr)
        }
    };
    float4* p4 = reinterpret_cast<float4*>(p); float4* gw4 = reinterpret_cast<float4*>(g);
    float4* a4 = reinterpret_cast<float4*>(m1); float4* b4 = reinterpret_cast<float4*>(m2);
#pragma unroll 2
    for (long long q = q0; q < n4; q += qstride) {
        float4 pv = p4[q], gv = gw4[q], av = a4[q], bv = ADAM ? b4[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        update(pv.x, gv.x, av.x, bv.x); update(pv.y, gv.y, av.y, bv.y); update(pv.z, gv.z, av.z, bv.z); update(pv.w, gv.w, av.w, bv.w);
        p4[q] = pv; gw4[q] = gv; a4[q] = av;
        if (ADAM) b4[q] = bv;
    }
    if (has_tail) {
        float pv = p[tail], gv = g[tail], av = m1[tail], bv = ADAM ? m2[tail] : 0.f;
        update(pv, gv, av, bv);
        p[tail] = pv; g[tail] = gv; m1[tail] = av;
        if (ADAM) m2[tail] = bv;
    }
}


// ---- data parallel: all-reduce over NVLink peer memory fused into the same cluster launch ---------------------------
// Every rank's flat [grad | loss_sum | mask_sum] buffer lives in cudaIpc-shared memory (marl_peer_*).  The kernel
//   1. signals "my gradient is complete" into every peer's flag array and waits for all peers (system-scope
//      release / acquire on 32-bit epoch counters, one slot per (phase, rank));
//   2. reads the W buffers with cache-volatile loads and sums them in rank order -- the same order on every rank, so
//      the replicas stay bit-identical -- keeping the sums in registers (n <= 2^18: <= 8 quads per thread);
//   3. computes the global norm like the single-GPU kernel (partials through distributed shared memory);
//   4. signals "done reading" / waits, so that nobody overwrites a gradient a peer still reads;
//   5. clips, updates its replica and leaves the clipped global gradient in its own buffer.
// One launch replaces {graph 1 end, ncclAllReduce, graph 2 start, clip + optimiser}; the step is one CUDA graph again.
constexpr int kPeerMaxWorld = MARL_PEER_MAX_WORLD;
constexpr int kPeerQuads = (int)(kClusterMax / 4 / (kClusterCtas * kClusterThreads));      // 8
struct PeerArgs {
    int world, rank;
    const float* g[kPeerMaxWorld];       // rank r's [grad n | loss_sum | mask_sum]; g[rank] is the local buffer
    unsigned* flags[kPeerMaxWorld];      // rank r's flag array [2][kPeerMaxWorld]
    unsigned* epoch;                     // local: launches so far
    int* error;                          // local: set to 1 when a peer did not arrive in time
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_peer4(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer(const float* p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
// CTA 0, threads t < world: tell rank t that (phase, me) reached epoch e, then wait until rank t told me the same
__device__ __forceinline__ void peer_barrier(const PeerArgs& pa, int phase, unsigned e) {
    if (threadIdx.x < (unsigned)pa.world) {
        const int t = threadIdx.x;
        __threadfence_system();
        st_release_sys(pa.flags[t] + phase * kPeerMaxWorld + pa.rank, e);
        const unsigned* mine = pa.flags[pa.rank] + phase * kPeerMaxWorld + t;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(mine) - e) < 0) {
            if (clock64() - t0 > (56LL << 30)) { *pa.error = 1; break; }     // ~30 s: a peer died; fail loudly on the host
        }
        __threadfence_system();
    }
}

template <bool ADAM>
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads)
clip_step_peer_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m1, float* __restrict__ m2, long long n,
                      float max_norm, float lr, float c1, float c2, float eps, int* step_counter, float* loss_out,
                      PeerArgs pa) {
    pdl_enter();
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float sw[kClusterThreads / 32];
    __shared__ float s_part, s_scale;
    const unsigned rank = cluster.block_rank();
    const unsigned e = *pa.epoch + 1;
    if (rank == 0) { peer_barrier(pa, 0, e); __syncthreads(); }          // every rank's backward is complete
    cluster.sync();
    const long long n4 = n >> 2, q0 = (long long)rank * kClusterThreads + threadIdx.x, qstride = (long long)kClusterCtas * kClusterThreads;
    float4 gs[kPeerQuads];
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < kPeerQuads; ++k) {
        const long long q = q0 + k * qstride;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < n4)
            for (int r = 0; r < pa.world; ++r) {
                const float4 x = ld_peer4(reinterpret_cast<const float4*>(pa.g[r]) + q);
                v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
            }
        gs[k] = v;
        acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
    float loss_sum = 0.f, mask_sum = 0.f;            // the two tail scalars, summed the same way (every thread: L1-free loads, 2 W values)
    if (threadIdx.x == 0)
        for (int r = 0; r < pa.world; ++r) { loss_sum += ld_peer(pa.g[r] + n); mask_sum += ld_peer(pa.g[r] + n + 1); }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < kClusterThreads / 32; ++w) t += sw[w];
        s_part = t;
        if (ADAM && step_counter && rank == 0) *step_counter += 1;
    }
    cluster.sync();                                  // all CTAs are past their peer reads; partials visible
    if (rank == 0) { peer_barrier(pa, 1, e); __syncthreads(); }          // ... on every rank: gradients may be overwritten from here on
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (unsigned r = 0; r < (unsigned)kClusterCtas; ++r) t += *cluster.map_shared_rank(&s_part, r);
        const float inv = 1.0f / mask_sum;
        const float total_norm = sqrtf(t) * inv;
        float coef = max_norm / (total_norm + 1e-6f);
        coef = coef > 1.0f ? 1.0f : coef;
        s_scale = inv * coef;
        if (rank == 0) {
            if (loss_out) { loss_out[0] = *pa.error ? __int_as_float(0x7fc00000) : loss_sum * inv; loss_out[1] = total_norm; }
            g[n] = loss_sum; g[n + 1] = mask_sum;   // what the all-reduce would have left in the tail
            *pa.epoch = e;
        }
    }
    cluster.sync();
    const float scale = s_scale;
    float step_size = 0.f, bc2_sqrt = 1.f;
    if (ADAM) {
        const int stp = *step_counter;
        step_size = (float)((double)lr / (1.0 - pow((double)c1, (double)stp)));
        bc2_sqrt = (float)sqrt(1.0 - pow((double)c2, (double)stp));
    }
    auto update = [&](float& pv, float& gv, float& av, float& bv) {
        const float gr = gv * scale;
        gv = gr;
        if (ADAM) {
            const float a = av + (1.0f - c1) * (gr - av);
            const float b = c2 * bv + (1.0f - c2) * gr * gr;
            av = a; bv = b;
            pv = pv - step_size * (a / (sqrtf(b) / bc2_sqrt + eps));
        } else {
            const float v = c1 * av + (1.0f - c1) * gr * gr;
            av = v;
            pv = pv - lr * (gr / (sqrtf(v) + eps));
        }
    };
    float4* p4 = reinterpret_cast<float4*>(p); float4* gw4 = reinterpret_cast<float4*>(g);
    float4* a4 = reinterpret_cast<float4*>(m1); float4* b4 = reinterpret_cast<float4*>(m2);
#pragma unroll
    for (int k = 0; k < kPeerQuads; ++k) {
        const long long q = q0 + k * qstride;
        if (q < n4) {
            float4 pv = p4[q], gv = gs[k], av = a4[q], bv = ADAM ? b4[q] : make_float4(0.f, 0.f, 0.f, 0.f);
            update(pv.x, gv.x, av.x, bv.x); update(pv.y, gv.y, av.y, bv.y); update(pv.z, gv.z, av.z, bv.z); update(pv.w, gv.w, av.w, bv.w);
            p4[q] = pv; gw4[q] = gv; a4[q] = av;
            if (ADAM) b4[q] = bv;
        }
    }
}

}  // namespace marl

using namespace marl;

extern "C" int marl_optim_partials(void) { return kOptBlocks; }

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int opt_grid(long long n) {
    long long b = (n + kOptThreads - 1) / kOptThreads;
    return (int)(b < 1 ? 1 : (b > kOptBlocks ? kOptBlocks : b));
}

extern "C" int marl_clip_rmsprop_step(float* params, float* grads, float* square_avg, long long n, const float* scalars,
                                      float max_norm, float lr, float alpha, float eps, float* partials, float* loss_out,
                                      void* stream) {
    if (!params || !grads || !square_avg || n <= 0 || !scalars || !partials) return MARL_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= kFusedMax) {
        { ProfScope ps_("clip_step_fused_kernel", st);
          clip_step_fused_kernel<false><<<1, kFusedThreads, 0, st>>>(params, grads, square_avg, nullptr, n, scalars, max_norm,
                                                                     lr, alpha, 0.f, eps, 0, nullptr, loss_out); }
        MARL_LAUNCH_CHECK();
        return MARL_OK;
    }
    if (n <= kClusterMax && aligned16(params) && aligned16(grads) && aligned16(square_avg)) {
        { ProfScope ps_("clip_step_cluster_kernel", st);
          launch_pdl(clip_step_cluster_kernel<false>, dim3(kClusterCtas), dim3(kClusterThreads), 0, st, params, grads, square_avg,
                     (float*)nullptr, n, scalars, max_norm, lr, alpha, 0.f, eps, 0, (int*)nullptr, loss_out); }
        MARL_LAUNCH_CHECK();
        return MARL_OK;
    }
    { ProfScope ps_("sumsq_kernel", st); sumsq_kernel<<<kOptBlocks, kOptThreads, 0, st>>>(grads, n, partials, nullptr); }
    MARL_LAUNCH_CHECK();
    { ProfScope ps_("clip_rmsprop_kernel", st); clip_rmsprop_kernel<<<opt_grid(n), kOptThreads, 0, st>>>(params, grads, square_avg, n, partials, scalars, max_norm, lr,
                                                             alpha, eps, loss_out); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

extern "C" int marl_clip_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                                   const float* scalars, float max_norm, float lr, float beta1, float beta2, float eps,
                                   int step, int* step_counter, float* partials, float* loss_out, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || n <= 0 || !scalars || !partials) return MARL_EINVAL;
    if (!step_counter && step < 1) return MARL_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= kFusedMax) {
        { ProfScope ps_("clip_step_fused_kernel", st);
          clip_step_fused_kernel<true><<<1, kFusedThreads, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, scalars, max_norm,
                                                                    lr, beta1, beta2, eps, step, step_counter, loss_out); }
        MARL_LAUNCH_CHECK();
        return MARL_OK;
    }
    if (n <= kClusterMax && aligned16(params) && aligned16(grads) && aligned16(exp_avg) && aligned16(exp_avg_sq)) {
        { ProfScope ps_("clip_step_cluster_kernel", st);
          launch_pdl(clip_step_cluster_kernel<true>, dim3(kClusterCtas), dim3(kClusterThreads), 0, st, params, grads, exp_avg, exp_avg_sq,
                     n, scalars, max_norm, lr, beta1, beta2, eps, step, step_counter, loss_out); }
        MARL_LAUNCH_CHECK();
        return MARL_OK;
    }
    { ProfScope ps_("sumsq_kernel", st); sumsq_kernel<<<kOptBlocks, kOptThreads, 0, st>>>(grads, n, partials, step_counter); }
    MARL_LAUNCH_CHECK();
    { ProfScope ps_("clip_adam_kernel", st); clip_adam_kernel<<<opt_grid(n), kOptThreads, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, partials, scalars,
                                                          max_norm, lr, beta1, beta2, eps, step, step_counter, loss_out); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

// ---- peer memory + the fused all-reduce / optimiser step ------------------------------------------------------------
extern "C" int marl_peer_alloc(size_t bytes, void** ptr) {
    if (!ptr || bytes == 0) return MARL_EINVAL;
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*ptr, 0, bytes);
    return e == cudaSuccess ? MARL_OK : (int)e;
}
extern "C" int marl_peer_free(void* ptr) { return ptr ? (int)cudaFree(ptr) : MARL_OK; }
extern "C" int marl_peer_export(const void* ptr, unsigned char* handle) {
    if (!ptr || !handle) return MARL_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == MARL_PEER_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr));
    if (e != cudaSuccess) return (int)e;
    memcpy(handle, &h, sizeof(h));
    return MARL_OK;
}
extern "C" int marl_peer_open(const unsigned char* handle, void** ptr) {
    if (!handle || !ptr) return MARL_EINVAL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}
extern "C" int marl_peer_close(void* ptr) { return ptr ? (int)cudaIpcCloseMemHandle(ptr) : MARL_OK; }

extern "C" int marl_clip_step_peer(int adam, float* params, float* grads, float* m1, float* m2, long long n, float max_norm,
                                   float lr, float c1, float c2, float eps, int* step_counter, float* loss_out,
                                   const marl_peer_group* pg, void* stream) {
    if (!params || !grads || !m1 || (adam && (!m2 || !step_counter)) || n <= 0 || !pg) return MARL_EINVAL;
    if (pg->world < 1 || pg->world > kPeerMaxWorld || pg->rank < 0 || pg->rank >= pg->world || !pg->epoch || !pg->error)
        return MARL_EINVAL;
    if (n > kClusterMax || (n & 3) || !aligned16(params) || !aligned16(grads) || !aligned16(m1) || (adam && !aligned16(m2)))
        return MARL_EINVAL;
    PeerArgs pa{};
    pa.world = pg->world; pa.rank = pg->rank; pa.epoch = pg->epoch; pa.error = pg->error;
    for (int r = 0; r < pg->world; ++r) {
        if (!pg->grads[r] || !pg->flags[r] || !aligned16(pg->grads[r])) return MARL_EINVAL;
        pa.g[r] = pg->grads[r]; pa.flags[r] = pg->flags[r];
    }
    if (pg->grads[pg->rank] != grads) return MARL_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps_("clip_step_peer_kernel", st);
    if (adam)
        launch_pdl(clip_step_peer_kernel<true>, dim3(kClusterCtas), dim3(kClusterThreads), 0, st, params, grads, m1, m2, n, max_norm,
                   lr, c1, c2, eps, step_counter, loss_out, pa);
    else
        launch_pdl(clip_step_peer_kernel<false>, dim3(kClusterCtas), dim3(kClusterThreads), 0, st, params, grads, m1, (float*)nullptr, n,
                   max_norm, lr, c1, c2, eps, (int*)nullptr, loss_out, pa);
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}
