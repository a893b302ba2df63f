// hlynr_rollout.cu -- timeout bootstrapping and GAE(lambda) on the device (C ABI: include/hlynr_rollout.h).
// Restated from stable-baselines3 2.x: common/on_policy_algorithm.py collect_rollouts, common/buffers.py
// RolloutBuffer.compute_returns_and_advantage (float32 arrays, Python-float gamma / gae_lambda).
#include <cuda_runtime.h>

#include "../../include/hlynr_rollout.h"

extern "C" int hlynr_internal_fail(const char* fmt, ...);
#define fail hlynr_internal_fail

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

__global__ void bootstrap_kernel(float* __restrict__ rewards, const HlynrDoneRecord* __restrict__ recs, const int32_t* __restrict__ counter,
                                 int32_t rows, const float* __restrict__ tv, float gamma, int32_t* overflow) {
    const int32_t count = *counter;
    const int32_t m = count < rows ? count : rows;
    for (int32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        const uint32_t fl = recs[r].flags;
        if ((fl & HLYNR_DONE_TRUNCATED) && !(fl & HLYNR_DONE_TERMINATED)) {
            const int32_t e = recs[r].env;  // one record per env per step: no write conflict
            rewards[e] = __fadd_rn(rewards[e], __fmul_rn(gamma, tv[r]));
        }
    }
    // records beyond `rows` got no value: count the ones that NEEDED one (truncated and not terminated) -- a non-zero *overflow
    // means timeouts went without their gamma * V(terminal_observation)
    if (overflow && count > rows) {
        int32_t missed = 0;
        for (int32_t r = rows + blockIdx.x * blockDim.x + threadIdx.x; r < count; r += gridDim.x * blockDim.x) {
            const uint32_t fl = recs[r].flags;
            missed += ((fl & HLYNR_DONE_TRUNCATED) && !(fl & HLYNR_DONE_TERMINATED)) ? 1 : 0;
        }
        if (missed) atomicAdd(overflow, missed);
    }
}

// one thread per env, reverse scan over T; every access is coalesced over N
__global__ void __launch_bounds__(256) gae_kernel(const float* __restrict__ rew, const float* __restrict__ val, const float* __restrict__ starts,
                                                  const float* __restrict__ last_val, const uint8_t* __restrict__ last_done, int64_t T,
                                                  int64_t N, float gamma, float gamma_lambda, float* __restrict__ adv, float* __restrict__ ret) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    float next_val = last_val[e];
    float next_nonterm = 1.f - (last_done[e] ? 1.f : 0.f);
    float lam = 0.f;
    for (int64_t t = T - 1; t >= 0; --t) {
        const int64_t k = t * N + e;
        const float v = val[k];
        // delta = r + gamma * next_values * next_non_terminal - values;  last = delta + gamma * lambda * next_non_terminal * last
        const float delta = __fsub_rn(__fadd_rn(rew[k], __fmul_rn(__fmul_rn(gamma, next_val), next_nonterm)), v);
        lam = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lambda, next_nonterm), lam));  // gamma * gae_lambda is a Python-float product
        adv[k] = lam;
        ret[k] = __fadd_rn(lam, v);
        next_val = v;
        next_nonterm = 1.f - starts[k];
    }
}

}  // namespace

extern "C" {

int hlynr_bootstrap_timeouts(float* rewards_dev, const HlynrDoneRecord* records_dev, const int32_t* counter_dev, int32_t rows,
                             const float* terminal_values_dev, double gamma, int32_t* overflow_dev, int device, void* stream) {
    if (!rewards_dev || !records_dev || !counter_dev || !terminal_values_dev) return fail("hlynr_bootstrap_timeouts: null argument");
    if (rows <= 0) return fail("hlynr_bootstrap_timeouts: rows must be positive");
    DeviceGuard g(device);
    const int blocks = rows < 256 * 64 ? (rows + 255) / 256 : 64;
    bootstrap_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(rewards_dev, records_dev, counter_dev, rows, terminal_values_dev, (float)gamma, overflow_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("hlynr_bootstrap_timeouts: %s", cudaGetErrorString(e));
    return 0;
}

int hlynr_gae(const float* rewards_dev, const float* values_dev, const float* episode_starts_dev, const float* last_values_dev,
              const uint8_t* last_dones_dev, int64_t T, int64_t N, double gamma, double gae_lambda, float* advantages_dev,
              float* returns_dev, int device, void* stream) {
    if (!rewards_dev || !values_dev || !episode_starts_dev || !last_values_dev || !last_dones_dev || !advantages_dev || !returns_dev)
        return fail("hlynr_gae: null argument");
    if (T <= 0 || N <= 0) return fail("hlynr_gae: T and N must be positive");
    DeviceGuard g(device);
    gae_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rewards_dev, values_dev, episode_starts_dev, last_values_dev,
                                                                            last_dones_dev, T, N, (float)gamma, (float)(gamma * gae_lambda),
                                                                            advantages_dev, returns_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("hlynr_gae: %s", cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
