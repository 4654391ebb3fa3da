// Batched TwoAgentsMatrixGame (env/single_state_matrix_game.py:5-120 of the reference):
// n_envs independent one-step cooperative matrix games per launch, each writing one
// episode record in the ReplayBuffer layout (common/replaybuffer.py:19-30, T = 1).
// HBM-bound: 16 B of actions in, 124 B of episode record out per env-step.
#include "common.cuh"
#include "../../include/marl_b200.h"
#include "profile.h"

namespace marl {

struct Payoff { double v[9]; };

// 4 envs per thread so that every key is written with 128-bit stores.
template <typename ActT>
__global__ void __launch_bounds__(256) matrix_game_step_kernel(
    Payoff pay, const ActT* __restrict__ actions, long long n_envs, float obs_value,
    float* __restrict__ o, float* __restrict__ s, long long* __restrict__ u, float* __restrict__ r,
    float* __restrict__ o_next, float* __restrict__ s_next, float* __restrict__ avail_u,
    float* __restrict__ avail_u_next, float* __restrict__ u_onehot, float* __restrict__ padded,
    float* __restrict__ terminated, double* __restrict__ r64) {
    __shared__ double sp[9];
    if (threadIdx.x < 9) sp[threadIdx.x] = pay.v[threadIdx.x];
    __syncthreads();
    const long long n_quads = n_envs >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float4 ov = make_float4(obs_value, obs_value, obs_value, obs_value);
    const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f), zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += stride) {
        int a[8];
        if (sizeof(ActT) == 8) {
            const longlong2* ap = reinterpret_cast<const longlong2*>(actions) + q * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) { longlong2 t = __ldg(ap + i); a[2 * i] = (int)t.x; a[2 * i + 1] = (int)t.y; }
        } else {
            const int4* ap = reinterpret_cast<const int4*>(actions) + q * 2;
            int4 t0 = __ldg(ap), t1 = __ldg(ap + 1);
            a[0] = t0.x; a[1] = t0.y; a[2] = t0.z; a[3] = t0.w; a[4] = t1.x; a[5] = t1.y; a[6] = t1.z; a[7] = t1.w;
        }
        double rw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rw[i] = sp[a[2 * i] * 3 + a[2 * i + 1]];   // step(): payoff_table[a0, a1]
        reinterpret_cast<float4*>(r)[q] = make_float4((float)rw[0], (float)rw[1], (float)rw[2], (float)rw[3]);
        if (r64) {
            reinterpret_cast<double2*>(r64)[2 * q] = make_double2(rw[0], rw[1]);
            reinterpret_cast<double2*>(r64)[2 * q + 1] = make_double2(rw[2], rw[3]);
        }
        reinterpret_cast<float4*>(s)[q] = ov;
        reinterpret_cast<float4*>(s_next)[q] = ov;
        reinterpret_cast<float4*>(padded)[q] = zero4;
        reinterpret_cast<float4*>(terminated)[q] = one4;
        reinterpret_cast<float4*>(o)[2 * q] = ov;
        reinterpret_cast<float4*>(o)[2 * q + 1] = ov;
        reinterpret_cast<float4*>(o_next)[2 * q] = ov;
        reinterpret_cast<float4*>(o_next)[2 * q + 1] = ov;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            reinterpret_cast<longlong2*>(u)[4 * q + i] = make_longlong2(a[2 * i], a[2 * i + 1]);
        float oh[24];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) oh[3 * i + c] = (a[i] == c) ? 1.f : 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            reinterpret_cast<float4*>(u_onehot)[6 * q + i] = make_float4(oh[4 * i], oh[4 * i + 1], oh[4 * i + 2], oh[4 * i + 3]);
            reinterpret_cast<float4*>(avail_u)[6 * q + i] = one4;
            reinterpret_cast<float4*>(avail_u_next)[6 * q + i] = one4;
        }
    }
    // scalar tail (n_envs % 4)
    for (long long e = (n_quads << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_envs; e += stride) {
        int a0 = (int)actions[2 * e], a1 = (int)actions[2 * e + 1];
        double rw = sp[a0 * 3 + a1];
        r[e] = (float)rw;
        if (r64) r64[e] = rw;
        s[e] = obs_value; s_next[e] = obs_value; padded[e] = 0.f; terminated[e] = 1.f;
        o[2 * e] = o[2 * e + 1] = obs_value;
        o_next[2 * e] = o_next[2 * e + 1] = obs_value;
        u[2 * e] = a0; u[2 * e + 1] = a1;
        for (int c = 0; c < 3; ++c) {
            u_onehot[6 * e + c] = (a0 == c) ? 1.f : 0.f;
            u_onehot[6 * e + 3 + c] = (a1 == c) ? 1.f : 0.f;
        }
        for (int c = 0; c < 6; ++c) { avail_u[6 * e + c] = 1.f; avail_u_next[6 * e + c] = 1.f; }
    }
}

__global__ void matrix_game_validate_kernel(const long long* a64, const int* a32, long long n, int* bad) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < 2 * n; i += stride) {
        long long v = a64 ? a64[i] : (long long)a32[i];
        if (v < 0 || v > 2) atomicExch(bad, 1);
    }
}

}  // namespace marl

extern "C" int marl_matrix_game_step(const double* payoff_host, const void* actions, int action_bytes,
                                     long long n_envs, float obs_value, const marl_episode_f32* out,
                                     double* r64, void* stream) {
    using namespace marl;
    if (!payoff_host || !actions || !out || n_envs < 0 || (action_bytes != 4 && action_bytes != 8)) return MARL_EINVAL;
    if (n_envs == 0) return MARL_OK;
    Payoff p;
    for (int i = 0; i < 9; ++i) p.v[i] = payoff_host[i];
    long long quads = (n_envs + 3) / 4;
    long long want = (quads + 255) / 256;
    int blocks = (int)(want < 1 ? 1 : (want > 8LL * kNumSMs ? 8LL * kNumSMs : want));
    if (blocks > kNumSMs) blocks = (blocks / kNumSMs) * kNumSMs;   // whole waves
    cudaStream_t st = (cudaStream_t)stream;
    if (action_bytes == 8)
        { ProfScope ps_("matrix_game_step_kernel", st); matrix_game_step_kernel<long long><<<blocks, 256, 0, st>>>(p, (const long long*)actions, n_envs, obs_value,
            out->o, out->s, out->u, out->r, out->o_next, out->s_next, out->avail_u, out->avail_u_next,
            out->u_onehot, out->padded, out->terminated, r64); }
    else
        { ProfScope ps_("matrix_game_step_kernel", st); matrix_game_step_kernel<int><<<blocks, 256, 0, st>>>(p, (const int*)actions, n_envs, obs_value,
            out->o, out->s, out->u, out->r, out->o_next, out->s_next, out->avail_u, out->avail_u_next,
            out->u_onehot, out->padded, out->terminated, r64); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}

extern "C" int marl_matrix_game_validate_actions(const void* actions, int action_bytes, long long n_envs,
                                                 int* bad_flag_device, void* stream) {
    using namespace marl;
    if (!actions || !bad_flag_device || (action_bytes != 4 && action_bytes != 8)) return MARL_EINVAL;
    if (n_envs == 0) return MARL_OK;
    int blocks = (int)((2 * n_envs + 255) / 256);
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    { ProfScope ps_("matrix_game_validate_kernel", (cudaStream_t)stream); matrix_game_validate_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        action_bytes == 8 ? (const long long*)actions : nullptr, action_bytes == 4 ? (const int*)actions : nullptr,
        n_envs, bad_flag_device); }
    MARL_LAUNCH_CHECK();
    return MARL_OK;
}
