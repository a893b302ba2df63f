/*
 * hlynr_rollout.h -- C ABI of the device-side pieces of PPO rollout collection that consume the tensor API of the
 * simulator (BASELINE config 5 "train.py PPO rollout collection").  The policy network itself is the
 * caller's (torch); these are the two non-GEMM steps stable-baselines3 runs in Python/NumPy around env.step:
 *
 *   reference interface (third party: stable-baselines3 2.x, see oracle/sb3_post.py)             replaced by
 *   ------------------------------------------------------------------------------------------  ------------------------
 *   OnPolicyAlgorithm.collect_rollouts: "if done and infos[idx]['TimeLimit.truncated']:          hlynr_bootstrap_timeouts
 *       rewards[idx] += gamma * policy.predict_values(terminal_observation)"
 *   RolloutBuffer.compute_returns_and_advantage (GAE(lambda), called at rl_system/scripts/       hlynr_gae
 *       train_flat_ppo.py:431-448 through model.learn)
 *
 * Conventions as in hlynr.h.
 */
#ifndef HLYNR_ROLLOUT_H
#define HLYNR_ROLLOUT_H

#include "hlynr.h"

#ifdef __cplusplus
extern "C" {
#endif

/* rewards_dev[env] += gamma * terminal_values_dev[r] for every record r < min(*counter_dev, rows) of the step's done list
 * that was truncated but not terminated (info['TimeLimit.truncated']).  terminal_values_dev has `rows` entries (the value
 * net evaluated on the first `rows` stacked terminal observations).  If the step finished more than `rows` episodes,
 * *overflow_dev (int32, device, may be NULL) is incremented by the number of truncated-and-not-terminated records left out
 * (timeouts that went without their bootstrap value): callers pass rows = N, or check *overflow_dev once per rollout. */
int hlynr_bootstrap_timeouts(float* rewards_dev, const HlynrDoneRecord* records_dev, const int32_t* counter_dev,
                             int32_t rows, const float* terminal_values_dev, double gamma, int32_t* overflow_dev,
                             int device, void* stream);

/* GAE(lambda) over a [T, N] rollout, float32 arithmetic in SB3's order (one thread per env, reverse scan over T):
 *   rewards_dev, values_dev, episode_starts_dev : float[T, N] (episode_starts = 1.0 where the step begins an episode)
 *   last_values_dev float[N], last_dones_dev uint8[N] : value / done flag of the observation after the last step
 *   advantages_dev, returns_dev : float[T, N] outputs (returns = advantages + values) */
int hlynr_gae(const float* rewards_dev, const float* values_dev, const float* episode_starts_dev,
              const float* last_values_dev, const uint8_t* last_dones_dev, int64_t T, int64_t N, double gamma,
              double gae_lambda, float* advantages_dev, float* returns_dev, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HLYNR_ROLLOUT_H */
