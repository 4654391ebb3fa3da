// Batched TwoAgentsMatrixGame (env/single_state_matrix_game.py:5-120 of the reference):
// n_envs independent one-step cooperative matrix games per launch, each writing one
// episode record in the ReplayBuffer layout (common/replaybuffer.py:19-30, T = 1).
// HBM-bound: 16 B of actions in, 124 B of episode record out per env-step.
#include "common.cuh"
#include "../../include/marl_b200.h"
#include "profile.h"

namespace marl {

struct Payoff { double v[9]; };

// One CTA steps a tile of kEnvTile environments: the actions are staged in shared memory, then every
// key of the episode record is written as one contiguous run with 128-bit stores (each warp instruction
// covers 512 contiguous bytes), whatever the key's per-env width (1, 2 or 6 values).
constexpr int kEnvTile = 2048;
constexpr int kEnvThreads = 256;

__device__ __forceinline__ void fill_run(float* __restrict__ base, long long first, long long count, float value) {
    // [first, first+count) floats, first % 4 == 0 and count % 4 == 0 guaranteed by the caller for full tiles
    const float4 v = make_float4(value, value, value, value);
    float4* p = reinterpret_cast<float4*>(base + first);
    for (long long q = threadIdx.x; q < (count >> 2); q += kEnvThreads) p[q] = v;
}

template <typename ActT>
__global__ void __launch_bounds__(kEnvThreads) matrix_game_step_kernel(
    Payoff pay, const ActT* __restrict__ actions, long long n_envs, float obs_value,
    float* __restrict__ o, float* __restrict__ s, long long* __restrict__ u, float* __restrict__ r,
    float* __restrict__ o_next, float* __restrict__ s_next, float* __restrict__ avail_u,
    float* __restrict__ avail_u_next, float* __restrict__ u_onehot, float* __restrict__ padded,
    float* __restrict__ terminated, double* __restrict__ r64, int T /* envs per tile, multiple of 256, <= kEnvTile */) {
    __shared__ double sp[9];
    __shared__ unsigned char sa[2 * kEnvTile];
    if (threadIdx.x < 9) sp[threadIdx.x] = pay.v[threadIdx.x];
    const long long n_tiles = (n_envs + T - 1) / T;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long e0 = tile * (long long)T;
        const int ne = (int)min((long long)T, n_envs - e0);
        __syncthreads();
        // stage the tile's actions (2 per env)
        if (ne == T) {
            if (sizeof(ActT) == 8) {
                const longlong2* ap = reinterpret_cast<const longlong2*>(actions + 2 * e0);
                for (int i = threadIdx.x; i < T; i += kEnvThreads) {
                    const longlong2 t = __ldg(ap + i);
                    sa[2 * i] = (unsigned char)t.x; sa[2 * i + 1] = (unsigned char)t.y;
                }
            } else {
                const int4* ap = reinterpret_cast<const int4*>(actions + 2 * e0);
                for (int i = threadIdx.x; i < T / 2; i += kEnvThreads) {
                    const int4 t = __ldg(ap + i);
                    sa[4 * i] = (unsigned char)t.x; sa[4 * i + 1] = (unsigned char)t.y;
                    sa[4 * i + 2] = (unsigned char)t.z; sa[4 * i + 3] = (unsigned char)t.w;
                }
            }
        } else {
            for (int i = threadIdx.x; i < 2 * ne; i += kEnvThreads) sa[i] = (unsigned char)actions[2 * e0 + i];
        }
        __syncthreads();
        if (ne == T) {
            // constant keys: obs / state (get_obs, get_state or the ones of get_episodes), masks, flags
            fill_run(s, e0, T, obs_value);
            fill_run(s_next, e0, T, obs_value);
            fill_run(padded, e0, T, 0.0f);
            fill_run(terminated, e0, T, 1.0f);
            fill_run(o, 2 * e0, 2 * T, obs_value);
            fill_run(o_next, 2 * e0, 2 * T, obs_value);
            fill_run(avail_u, 6 * e0, 6 * T, 1.0f);
            fill_run(avail_u_next, 6 * e0, 6 * T, 1.0f);
            // reward: payoff_table[a0, a1] (step(), float64 in the reference)
            for (int q = threadIdx.x; q < T / 4; q += kEnvThreads) {
                double rw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) rw[k] = sp[sa[8 * q + 2 * k] * 3 + sa[8 * q + 2 * k + 1]];
                reinterpret_cast<float4*>(r + e0)[q] = make_float4((float)rw[0], (float)rw[1], (float)rw[2], (float)rw[3]);
                if (r64) {
                    reinterpret_cast<double2*>(r64 + e0)[2 * q] = make_double2(rw[0], rw[1]);
                    reinterpret_cast<double2*>(r64 + e0)[2 * q + 1] = make_double2(rw[2], rw[3]);
                }
            }
            // actions as int64 [env, 2]: one env (16 B) per lane
            for (int i = threadIdx.x; i < T; i += kEnvThreads)
                reinterpret_cast<longlong2*>(u + 2 * e0)[i] = make_longlong2(sa[2 * i], sa[2 * i + 1]);
            // one-hot of both actions: 6 floats per env, written as a flat run of float4
            for (int q = threadIdx.x; q < 6 * T / 4; q += kEnvThreads) {
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int f = 4 * q + k, agent = f / 3, c = f - 3 * agent;     // f = env*6 + agent_in_env*3 + c
                    v[k] = (sa[agent] == c) ? 1.0f : 0.0f;
                }
                reinterpret_cast<float4*>(u_onehot + 6 * e0)[q] = make_float4(v[0], v[1], v[2], v[3]);
            }
        } else {
            // ragged last tile: scalar path
            for (int i = threadIdx.x; i < ne; i += kEnvThreads) {
                const long long e = e0 + i;
                const int a0 = sa[2 * i], a1 = sa[2 * i + 1];
                const double rw = sp[a0 * 3 + a1];
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
    // tile: as large as possible (long contiguous runs) while still giving every SM a couple of tiles
    long long per = (n_envs + 2 * kNumSMs - 1) / (2 * kNumSMs);
    int T = (int)((per + 255) / 256 * 256);
    if (T < 256) T = 256;
    if (T > kEnvTile) T = kEnvTile;
    long long want = (n_envs + T - 1) / T;
    int blocks = (int)(want < 1 ? 1 : (want > 8LL * kNumSMs ? 8LL * kNumSMs : want));
    if (blocks > kNumSMs) blocks = (blocks / kNumSMs) * kNumSMs;   // whole waves of the 148 SMs
    cudaStream_t st = (cudaStream_t)stream;
    if (action_bytes == 8)
        { ProfScope ps_("matrix_game_step_kernel", st); matrix_game_step_kernel<long long><<<blocks, 256, 0, st>>>(p, (const long long*)actions, n_envs, obs_value,
            out->o, out->s, out->u, out->r, out->o_next, out->s_next, out->avail_u, out->avail_u_next,
            out->u_onehot, out->padded, out->terminated, r64, T); }
    else
        { ProfScope ps_("matrix_game_step_kernel", st); matrix_game_step_kernel<int><<<blocks, 256, 0, st>>>(p, (const int*)actions, n_envs, obs_value,
            out->o, out->s, out->u, out->r, out->o_next, out->s_next, out->avail_u, out->avail_u_next,
            out->u_onehot, out->padded, out->terminated, r64, T); }
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
