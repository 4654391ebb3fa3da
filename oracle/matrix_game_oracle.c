/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference matrix game
 * (env/single_state_matrix_game.py:27-40 step(), :81-120 episode record layout), used by the
 * parity tests and as the single-core CPU baseline of bench.py.  Never linked into the product. */
#include <stdint.h>

/* One env step per entry: reward = payoff[a0][a1] (float64), terminated = 1.  Writes the same
 * 11-key episode record (T = 1) that the CUDA kernel emits: fp32 arrays, int64 actions. */
void mg_oracle_step(const double* payoff /*[9]*/, const int64_t* actions /*[n][2]*/, int64_t n, float obs_value,
                    float* o, float* s, int64_t* u, float* r, float* o_next, float* s_next, float* avail_u,
                    float* avail_u_next, float* u_onehot, float* padded, float* terminated, double* r64) {
    for (int64_t e = 0; e < n; ++e) {
        const int64_t a0 = actions[2 * e], a1 = actions[2 * e + 1];
        const double rew = payoff[a0 * 3 + a1];
        r[e] = (float)rew;
        if (r64) r64[e] = rew;
        s[e] = obs_value; s_next[e] = obs_value;
        padded[e] = 0.0f; terminated[e] = 1.0f;
        for (int i = 0; i < 2; ++i) { o[2 * e + i] = obs_value; o_next[2 * e + i] = obs_value; }
        u[2 * e] = a0; u[2 * e + 1] = a1;
        for (int c = 0; c < 3; ++c) {
            u_onehot[6 * e + c] = (a0 == c) ? 1.0f : 0.0f;
            u_onehot[6 * e + 3 + c] = (a1 == c) ? 1.0f : 0.0f;
        }
        for (int c = 0; c < 6; ++c) { avail_u[6 * e + c] = 1.0f; avail_u_next[6 * e + c] = 1.0f; }
    }
}
