// hlynr_post.cu -- on-device VecFrameStack + VecNormalize (C ABI: include/hlynr_post.h).
//
// Semantics restated from stable-baselines3 2.x (oracle/sb3_post.py cites the source files):
//   stacked[e] = concat(frame[t-k+1], ..., frame[t]); a finished env's stack is zeroed before its reset observation
//   is inserted; info['terminal_observation'] = concat(previous_stack[26:], terminal_obs);
//   obs_rms.update(stacked batch) (batch mean/var merged into the running float64 moments), then
//   out = clip((stacked - mean) / sqrt(var + eps), -clip, clip) as float32; terminal observations likewise.
//
// HBM traffic per env-step (n_stack = 4): 104 B column-sum read of the new frame + 416 B ring read + 416 B output
// write (+ 9 B age/done flags, 16 B returns) -- the explicit roll/copy of SB3 (another 832 B) never happens, and the
// step kernel writes its observation straight into the ring slot.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include <cuda_runtime.h>

#include "../../include/hlynr_post.h"

extern "C" int hlynr_internal_fail(const char* fmt, ...);  // hlynr_capi.cu: sets the thread's hlynr_last_error()
#define fail hlynr_internal_fail
#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

namespace {

constexpr int OD = HLYNR_OBS_DIM;  // 26
constexpr int KMAX = HLYNR_POST_MAX_STACK;
constexpr int DMAX = OD * KMAX;

// device statistics block (doubles)
struct PostStats {
    double S[KMAX][OD], Q[KMAX][OD];        // per-lag column sums / sums of squares of the current stacked batch (lag 0 = newest)
    double col[OD], colq[OD];               // scratch: sums over the new frame
    double corr[KMAX][OD], corrq[KMAX][OD]; // scratch: frames of the envs that finished in this step (subtracted)
    double ret_sum, ret_sq;                 // scratch: sums over the discounted returns
    double mean[DMAX], var[DMAX], count;    // obs_rms (feature f = j*26 + c, j = 0 oldest)
    double ret_mean, ret_var, ret_count;    // ret_rms
    double chk_S[KMAX][OD], chk_Q[KMAX][OD];  // check_sums scratch
    double inv_std[DMAX];                   // 1 / sqrt(var + eps): the normalise kernels compute in float64 like SB3 (obs f32 - mean f64)
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

__device__ __forceinline__ int ring_slot(int newest, int lag, int k) { int s = newest - lag; return s < 0 ? s + k : s; }

// ---- sums over the new frame: thread-constant channel, 16 envs per 416-thread block iteration, coalesced ----
// Every reduction of this file is two-stage and atomic-free: blocks write partial sums to partials[block][word], one
// block per word then folds them in a fixed order, so the statistics (and with them the normalised observations) are
// bit-reproducible from run to run.
constexpr int PW_COL = 0, PW_COLQ = OD, PW_RET = 2 * OD, PW_RETQ = 2 * OD + 1, PW_CORR = 2 * OD + 2;
constexpr int PART_WORDS = PW_CORR + 2 * (KMAX - 1) * OD;   // + corr[(k-1)][26], corrq[(k-1)][26]

__global__ void __launch_bounds__(16 * OD) colsum_kernel(const float* __restrict__ frame, int64_t n, double* __restrict__ partials) {
    __shared__ double sh[16 * OD];
    const int t = threadIdx.x, c = t % OD, r = t / OD;
    double s = 0.0, q = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * 16 + r; e < n; e += (int64_t)gridDim.x * 16) {
        const double v = (double)frame[e * OD + c];
        s += v; q += v * v;
    }
    sh[t] = s;
    __syncthreads();
    double* mine = partials + (size_t)blockIdx.x * PART_WORDS;
    if (r == 0) { double a = 0.0; for (int i = 0; i < 16; ++i) a += sh[i * OD + c]; mine[PW_COL + c] = a; }
    __syncthreads();
    sh[t] = q;
    __syncthreads();
    if (r == 0) { double a = 0.0; for (int i = 0; i < 16; ++i) a += sh[i * OD + c]; mine[PW_COLQ + c] = a; }
}

// second stage: block w folds word w of the partial rows (fixed order) into its PostStats field
__global__ void __launch_bounds__(256) fold_kernel(const double* __restrict__ partials, int nb_col, int nb_ret, int nb_corr, int per,
                                                   PostStats* st) {
    __shared__ double sh[256];
    const int w = blockIdx.x < 2 * OD + 2 ? blockIdx.x : PW_CORR + (blockIdx.x - (2 * OD + 2));
    const int nblocks = w < 2 * OD ? nb_col : (w < PW_CORR ? nb_ret : nb_corr);
    double a = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) a += partials[(size_t)b * PART_WORDS + w];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    if (w < 2 * OD) (&st->col[0])[w] = sh[0];                       // col[26], colq[26] are adjacent
    else if (w < PW_CORR) (&st->ret_sum)[w - 2 * OD] = sh[0];        // ret_sum, ret_sq
    else if (w - PW_CORR < per) (&st->corr[1][0])[w - PW_CORR] = sh[0];
    else (&st->corrq[1][0])[w - PW_CORR - per] = sh[0];
}

// ---- VecNormalize._update_reward: returns = returns * gamma + reward; ret_rms sums; returns[done] = 0 ----
// Also the per-env part of VecFrameStack: age_new = done ? 0 : min(age_old + 1, k - 1) (how many older frames of the
// ring belong to the env's current episode).
__global__ void __launch_bounds__(256) returns_kernel(double* __restrict__ ret, const float* __restrict__ rew,
                                                      const uint8_t* __restrict__ term, const uint8_t* __restrict__ trunc, int64_t n,
                                                      double gamma, double* __restrict__ partials, int training,
                                                      const uint8_t* __restrict__ age_in, uint8_t* __restrict__ age_out, int k) {
    double s = 0.0, q = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        double r = ret[e];
        if (training) { r = r * gamma + (double)rew[e]; s += r; q += r * r; }
        const bool done = (term[e] | trunc[e]) != 0;
        ret[e] = done ? 0.0 : r;  // returns[dones] = 0 happens whether or not the statistics move
        const int a = (int)age_in[e] + 1;
        age_out[e] = (uint8_t)(done ? 0 : (a < k - 1 ? a : k - 1));
    }
    if (!training) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    __shared__ double shs[8], shq[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { shs[w] = s; shq[w] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += shs[i]; b += shq[i]; }
        partials[(size_t)blockIdx.x * PART_WORDS + PW_RET] = a; partials[(size_t)blockIdx.x * PART_WORDS + PW_RETQ] = b;
    }
}

// ---- frames of finished envs leave the stack: accumulate what has to be subtracted from the shifted sums ----
// thread = (record lane r, old lag d in [0, k-2], channel c); old lag d becomes lag d+1 for the envs that go on.
__global__ void __launch_bounds__(512) done_corr_kernel(const HlynrDoneRecord* __restrict__ recs, const int32_t* __restrict__ counter,
                                                        int32_t cap, const float* __restrict__ frames, int64_t plane,
                                                        const uint8_t* __restrict__ age_old, int k, int newest_old,
                                                        double* __restrict__ partials) {
    const int per = (k - 1) * OD;                 // threads per record
    const int lanes = blockDim.x / per;           // records per block iteration
    const int t = threadIdx.x;
    if (t >= lanes * per) return;
    const int r = t / per, f = t % per, d = f / OD, c = f % OD;
    int32_t count = *counter;
    if (count > cap) count = cap;
    double s = 0.0, q = 0.0;
    const float* src = frames + (int64_t)ring_slot(newest_old, d, k) * plane;
    for (int32_t i = blockIdx.x * lanes + r; i < count; i += gridDim.x * lanes) {
        const int32_t e = recs[i].env;
        if (d <= (int)age_old[e]) { const double v = (double)src[(int64_t)e * OD + c]; s += v; q += v * v; }
    }
    // fold the record lanes of this block, then one partial row per block: [corr (k-1)*26 | corrq (k-1)*26]
    __shared__ double sh[512];
    sh[t] = s;
    __syncthreads();
    double* mine = partials + (size_t)blockIdx.x * PART_WORDS + PW_CORR;
    if (r == 0) { double a = 0.0; for (int i = 0; i < lanes; ++i) a += sh[i * per + f]; mine[f] = a; }
    __syncthreads();
    sh[t] = q;
    __syncthreads();
    if (r == 0) { double a = 0.0; for (int i = 0; i < lanes; ++i) a += sh[i * per + f]; mine[per + f] = a; }
}

// ---- shift the per-lag sums, merge the batch moments into obs_rms / ret_rms (RunningMeanStd.update_from_moments) ----
__global__ void merge_kernel(PostStats* st, int k, double n, int shift, int training, int reset_all, int with_returns, double eps) {
    const int f = threadIdx.x, D = k * OD;
    if (shift && f < OD) {  // the per-lag sums follow the ring on every reset / step
        const int c = f;
        for (int d = k - 1; d >= 1; --d) {
            st->S[d][c] = reset_all ? 0.0 : st->S[d - 1][c] - st->corr[d][c];
            st->Q[d][c] = reset_all ? 0.0 : st->Q[d - 1][c] - st->corrq[d][c];
            st->corr[d][c] = 0.0; st->corrq[d][c] = 0.0;
        }
        st->S[0][c] = st->col[c]; st->Q[0][c] = st->colq[c];
        st->col[c] = 0.0; st->colq[c] = 0.0;
    }
    const double count = st->count;
    __syncthreads();
    if (f < D) {
        if (training) {
            const int j = f / OD, c = f % OD, d = k - 1 - j;
            const double bm = st->S[d][c] / n;
            double bv = st->Q[d][c] / n - bm * bm;
            if (bv < 0.0) bv = 0.0;
            const double delta = bm - st->mean[f], tot = count + n;
            const double m2 = st->var[f] * count + bv * n + delta * delta * count * n / tot;
            st->mean[f] = st->mean[f] + delta * n / tot;
            st->var[f] = m2 / tot;
        }
        st->inv_std[f] = 1.0 / sqrt(st->var[f] + eps);
    }
    __syncthreads();
    if (f == 0 && training) {
        st->count = count + n;
        if (with_returns) {
            const double bm = st->ret_sum / n;
            double bv = st->ret_sq / n - bm * bm;
            if (bv < 0.0) bv = 0.0;
            const double c0 = st->ret_count, delta = bm - st->ret_mean, tot = c0 + n;
            const double m2 = st->ret_var * c0 + bv * n + delta * delta * c0 * n / tot;
            st->ret_mean = st->ret_mean + delta * n / tot;
            st->ret_var = m2 / tot;
            st->ret_count = tot;
        }
        st->ret_sum = 0.0; st->ret_sq = 0.0;
    }
}

// ---- stacked (+ normalised) observation: thread-constant feature PAIR (26 is even, so a float2 never straddles two
// ring slots and every load/store is 8-byte aligned), 4 envs in flight per thread, coalesced [N, 26k] output ----
#define NORM_UNROLL 8
__global__ void __launch_bounds__(512) normalize_kernel(const float* __restrict__ frames, int64_t plane, const uint8_t* __restrict__ age,
                                                        const PostStats* __restrict__ st, float* __restrict__ out, int64_t n, int k,
                                                        int newest, float clip, int raw) {
    const int D = k * OD, P = D / 2, lanes = blockDim.x / P, t = threadIdx.x;
    if (t >= lanes * P) return;
    const int r = t / P, p = t % P, f = 2 * p, j = f / OD, c = f % OD, lag = k - 1 - j;
    // SB3 evaluates (obs_f32 - mean_f64) / sqrt(var + eps) in float64 and rounds the result to float32.  Here the mean is
    // split into a float pair (hi + lo), so the cancellation obs - mean keeps ~48 bits; the product and the clip are float:
    // the result is within 2 float ulps of SB3's, without touching the FP64 pipe.
    const float m0h = (float)st->mean[f], m0l = (float)(st->mean[f] - (double)m0h), i0 = (float)st->inv_std[f];
    const float m1h = (float)st->mean[f + 1], m1l = (float)(st->mean[f + 1] - (double)m1h), i1 = (float)st->inv_std[f + 1];
    const float* src = frames + (int64_t)ring_slot(newest, lag, k) * plane + c;
    const int64_t stride = (int64_t)gridDim.x * lanes;
    for (int64_t e0 = (int64_t)blockIdx.x * lanes + r; e0 < n; e0 += stride * NORM_UNROLL) {
        float2 v[NORM_UNROLL];
        int ag[NORM_UNROLL];
        bool ok[NORM_UNROLL];
#pragma unroll
        for (int u = 0; u < NORM_UNROLL; ++u) {  // frame and age loads are independent: all in flight together; a slot
            const int64_t e = e0 + u * stride;   // older than the env's episode holds stale but readable data
            ok[u] = e < n;
            v[u] = make_float2(0.f, 0.f);
            ag[u] = 0;
            if (ok[u]) { v[u] = __ldcs(reinterpret_cast<const float2*>(src + e * OD)); ag[u] = (int)__ldg(age + e); }
        }
#pragma unroll
        for (int u = 0; u < NORM_UNROLL; ++u) {
            if (!ok[u]) continue;
            float2 w = lag <= ag[u] ? v[u] : make_float2(0.f, 0.f);
            if (!raw) {
                w.x = fminf(fmaxf(((w.x - m0h) - m0l) * i0, -clip), clip);
                w.y = fminf(fmaxf(((w.y - m1h) - m1l) * i1, -clip), clip);
            }
            __stcs(reinterpret_cast<float2*>(out + (e0 + u * stride) * D + f), w);
        }
    }
}

// ---- info['terminal_observation'] of finished envs: concat(previous_stack[26:], terminal_obs), normalised ----
__global__ void __launch_bounds__(512) terminal_kernel(const HlynrDoneRecord* __restrict__ recs, const int32_t* __restrict__ counter, int32_t cap,
                                                       const float* __restrict__ frames, int64_t plane, const uint8_t* __restrict__ age_old,
                                                       const PostStats* __restrict__ st, float* __restrict__ out, int k, int newest_old,
                                                       float clip) {
    const int D = k * OD, lanes = blockDim.x / D, t = threadIdx.x;
    if (t >= lanes * D) return;
    const int r = t / D, f = t % D, j = f / OD, c = f % OD;
    const double mean = st->mean[f], inv = st->inv_std[f];
    int32_t count = *counter;
    if (count > cap) count = cap;
    for (int32_t i = blockIdx.x * lanes + r; i < count; i += gridDim.x * lanes) {
        float v;
        if (j == k - 1) v = recs[i].terminal_obs[c];
        else {
            const int32_t e = recs[i].env;
            const int lag_old = k - 2 - j;
            v = lag_old <= (int)age_old[e] ? frames[(int64_t)ring_slot(newest_old, lag_old, k) * plane + (int64_t)e * OD + c] : 0.f;
        }
        out[(int64_t)i * D + f] = (float)fmin(fmax(((double)v - mean) * inv, -(double)clip), (double)clip);
    }
}

// ---- VecNormalize.normalize_obs on caller rows ----
__global__ void __launch_bounds__(512) normalize_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows, int D,
                                                             const PostStats* __restrict__ st, float clip) {
    const int lanes = blockDim.x / D, t = threadIdx.x;
    if (t >= lanes * D) return;
    const int r = t / D, f = t % D;
    const double mean = st->mean[f], inv = st->inv_std[f];
    for (int64_t e = (int64_t)blockIdx.x * lanes + r; e < rows; e += (int64_t)gridDim.x * lanes)
        out[e * D + f] = (float)fmin(fmax(((double)in[e * D + f] - mean) * inv, -(double)clip), (double)clip);
}

// ---- test hook: per-lag sums recomputed from the ring ----
__global__ void __launch_bounds__(512) recompute_sums_kernel(const float* __restrict__ frames, int64_t plane, const uint8_t* __restrict__ age,
                                                             int64_t n, int k, int newest, PostStats* st) {
    const int D = k * OD, lanes = blockDim.x / D, t = threadIdx.x;
    if (t >= lanes * D) return;
    const int r = t / D, f = t % D, d = f / OD, c = f % OD;  // here f enumerates (lag d, channel c)
    const float* src = frames + (int64_t)ring_slot(newest, d, k) * plane + c;
    double s = 0.0, q = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * lanes + r; e < n; e += (int64_t)gridDim.x * lanes)
        if (d <= (int)age[e]) { const double v = (double)src[e * OD]; s += v; q += v * v; }
    atomicAdd(&st->chk_S[d][c], s); atomicAdd(&st->chk_Q[d][c], q);
}

}  // namespace

struct hlynr_post {
    int64_t n = 0, plane = 0;   // plane = floats per ring slot (n * 26 rounded up to 16-byte multiples)
    int device = 0, k = 1, D = OD;
    double clip = 10.0, eps = 1e-8, gamma = 0.99;
    float* frames = nullptr;
    uint8_t* age[2] = {nullptr, nullptr};
    int age_cur = 0;
    double* returns = nullptr;
    PostStats* st = nullptr;
    double* partials = nullptr; // [max_blocks][PART_WORDS] first-stage sums
    uint32_t t = 0;             // frames written so far; target slot = t % k, newest = (t - 1) % k
    bool sums_valid = false;    // the per-lag sums S/Q describe the current ring (maintained by every reset/step that can)
    int sm_count = 148;
    int64_t launches = 0;
};

static int blocks_for(const hlynr_post* p, int64_t items, int per_block) {
    int64_t b = (items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)p->sm_count * 8;
    if (b > cap) b = cap;
    return b < 1 ? 1 : (int)b;
}

extern "C" {

int hlynr_post_create(int64_t n_envs, int device, int n_stack, double clip_obs, double epsilon, double gamma, hlynr_post_t** out) {
    if (!out) return fail("hlynr_post_create: null argument");
    if (n_envs <= 0) return fail("hlynr_post_create: n_envs must be positive");
    if (n_stack < 1 || n_stack > KMAX) return fail("hlynr_post_create: n_stack must be in [1, %d]", KMAX);
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail("hlynr_post_create: device %d not available", device);
    DeviceGuard g(device);
    hlynr_post* p = new (std::nothrow) hlynr_post();
    if (!p) return fail("hlynr_post_create: out of host memory");
    p->n = n_envs; p->device = device; p->k = n_stack; p->D = OD * n_stack;
    p->clip = clip_obs; p->eps = epsilon; p->gamma = gamma;
    p->plane = ((n_envs * OD + 31) / 32) * 32;  // every slot starts 128-byte aligned
    cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaError_t e = cudaMalloc(&p->frames, sizeof(float) * p->plane * n_stack);
    if (e == cudaSuccess) e = cudaMalloc(&p->age[0], n_envs);
    if (e == cudaSuccess) e = cudaMalloc(&p->age[1], n_envs);
    if (e == cudaSuccess) e = cudaMalloc(&p->returns, sizeof(double) * n_envs);
    if (e == cudaSuccess) e = cudaMalloc(&p->st, sizeof(PostStats));
    if (e == cudaSuccess) e = cudaMalloc(&p->partials, sizeof(double) * PART_WORDS * ((size_t)p->sm_count * 8 + 64));
    if (e == cudaSuccess) e = cudaMemset(p->partials, 0, sizeof(double) * PART_WORDS * ((size_t)p->sm_count * 8 + 64));
    if (e != cudaSuccess) { int r = fail("hlynr_post_create: cudaMalloc failed: %s", cudaGetErrorString(e)); hlynr_post_destroy(p); return r; }
    cudaMemset(p->frames, 0, sizeof(float) * p->plane * n_stack);
    cudaMemset(p->age[0], 0, n_envs); cudaMemset(p->age[1], 0, n_envs);
    cudaMemset(p->returns, 0, sizeof(double) * n_envs);
    PostStats* h = new (std::nothrow) PostStats();
    if (!h) { hlynr_post_destroy(p); return fail("hlynr_post_create: out of host memory"); }
    memset(h, 0, sizeof(*h));
    for (int f = 0; f < DMAX; ++f) { h->var[f] = 1.0; h->inv_std[f] = 1.0 / sqrt(1.0 + epsilon); }
    h->count = 1e-4; h->ret_var = 1.0; h->ret_count = 1e-4;  // RunningMeanStd(epsilon=1e-4)
    e = cudaMemcpy(p->st, h, sizeof(*h), cudaMemcpyHostToDevice);
    delete h;
    if (e != cudaSuccess) { int r = fail("hlynr_post_create: %s", cudaGetErrorString(e)); hlynr_post_destroy(p); return r; }
    *out = p;
    return 0;
}

void hlynr_post_destroy(hlynr_post_t* p) {
    if (!p) return;
    DeviceGuard g(p->device);
    cudaFree(p->frames); cudaFree(p->age[0]); cudaFree(p->age[1]); cudaFree(p->returns); cudaFree(p->st); cudaFree(p->partials);
    delete p;
}

int hlynr_post_obs_dim(const hlynr_post_t* p, int* out) { if (!p || !out) return fail("null argument"); *out = p->D; return 0; }
int hlynr_post_launch_count(const hlynr_post_t* p, int64_t* out) { if (!p || !out) return fail("null argument"); *out = p->launches; return 0; }

int hlynr_post_note_replayed_steps(hlynr_post_t* p, int64_t steps, int64_t launches) {
    if (!p) return fail("hlynr_post_note_replayed_steps: null handle");
    p->t = (uint32_t)((int64_t)p->t + steps); p->launches += launches;  // negative: undo the bookkeeping of recorded-only calls
    return 0;
}

int hlynr_post_obs_target(hlynr_post_t* p, float** obs_dev) {
    if (!p || !obs_dev) return fail("hlynr_post_obs_target: null argument");
    *obs_dev = p->frames + (int64_t)(p->t % (uint32_t)p->k) * p->plane;
    return 0;
}

static int launch_normalize(hlynr_post* p, float* out, int raw, cudaStream_t st) {
    const int P = p->D / 2;
    const int lanes = 512 / P > 0 ? 512 / P : 1;
    const int newest = (int)((p->t - 1) % (uint32_t)p->k);
    normalize_kernel<<<blocks_for(p, (p->n + NORM_UNROLL - 1) / NORM_UNROLL, lanes), lanes * P, 0, st>>>(
        p->frames, p->plane, p->age[p->age_cur], p->st, out, p->n, p->k, newest, (float)p->clip, raw);
    CK(cudaGetLastError());
    p->launches += 1;
    return 0;
}

int hlynr_post_reset(hlynr_post_t* p, float* out_dev, int training, void* stream) {
    if (!p || !out_dev) return fail("hlynr_post_reset: null argument");
    DeviceGuard g(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const float* frame = p->frames + (int64_t)(p->t % (uint32_t)p->k) * p->plane;
    p->t += 1;
    CK(cudaMemsetAsync(p->returns, 0, sizeof(double) * p->n, st));  // VecNormalize.reset: returns = zeros
    const int nb_col = blocks_for(p, p->n, 16);
    colsum_kernel<<<nb_col, 16 * OD, 0, st>>>(frame, p->n, p->partials);
    fold_kernel<<<2 * OD, 256, 0, st>>>(p->partials, nb_col, 0, 0, 0, p->st);
    merge_kernel<<<1, 256, 0, st>>>(p->st, p->k, (double)p->n, 1, training, 1, 0, p->eps);
    CK(cudaGetLastError());
    p->launches += 3;
    p->sums_valid = true;
    CK(cudaMemsetAsync(p->age[p->age_cur], 0, p->n, st));  // VecFrameStack.reset: only the newest frame is valid
    return launch_normalize(p, out_dev, 0, st);
}

int hlynr_post_step(hlynr_post_t* p, const float* reward_dev, const uint8_t* terminated_dev, const uint8_t* truncated_dev,
                    const HlynrDoneRecord* records_dev, const int32_t* counter_dev, int32_t capacity, float* out_dev,
                    float* terminal_out_dev, int training, void* stream) {
    if (!p || !reward_dev || !terminated_dev || !truncated_dev || !out_dev) return fail("hlynr_post_step: null argument");
    if (p->t == 0) return fail("hlynr_post_step: call hlynr_post_reset first");
    const bool can_maintain = p->k == 1 || (records_dev && counter_dev);
    if (training && !can_maintain) return fail("hlynr_post_step: training with n_stack > 1 needs the done list of the step");
    if (training && !p->sums_valid)
        return fail("hlynr_post_step: the column sums are stale (earlier steps ran without a done list); call hlynr_post_check_sums(resync=1) first");
    if (terminal_out_dev && (!records_dev || !counter_dev)) return fail("hlynr_post_step: terminal_out_dev needs the done list of the step");
    DeviceGuard g(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int newest_old = (int)((p->t - 1) % (uint32_t)p->k);
    const float* frame = p->frames + (int64_t)(p->t % (uint32_t)p->k) * p->plane;
    const uint8_t* age_old = p->age[p->age_cur];
    p->t += 1;
    const int nb_ret = blocks_for(p, p->n, 256);
    returns_kernel<<<nb_ret, 256, 0, st>>>(p->returns, reward_dev, terminated_dev, truncated_dev, p->n, p->gamma, p->partials, training,
                                           age_old, p->age[p->age_cur ^ 1], p->k);
    p->launches += 1;
    if (can_maintain && p->sums_valid) {  // the sums follow the ring on every step; obs_rms / ret_rms only move when training
        const int nb_col = blocks_for(p, p->n, 16);
        colsum_kernel<<<nb_col, 16 * OD, 0, st>>>(frame, p->n, p->partials);
        p->launches += 1;
        const int per = (p->k - 1) * OD;
        if (p->k > 1) {
            const int lanes = 512 / per;
            done_corr_kernel<<<64, lanes * per, 0, st>>>(records_dev, counter_dev, capacity, p->frames, p->plane, age_old, p->k, newest_old,
                                                         p->partials);
            p->launches += 1;
        }
        fold_kernel<<<2 * OD + 2 + 2 * per, 256, 0, st>>>(p->partials, nb_col, training ? nb_ret : 0, 64, per, p->st);
        p->launches += 1;
        merge_kernel<<<1, 256, 0, st>>>(p->st, p->k, (double)p->n, 1, training, 0, training, p->eps);
        p->launches += 1;
    } else {
        p->sums_valid = false;
    }
    if (terminal_out_dev) {
        const int lanes = 512 / p->D > 0 ? 512 / p->D : 1;
        terminal_kernel<<<64, lanes * p->D, 0, st>>>(records_dev, counter_dev, capacity, p->frames, p->plane, age_old, p->st, terminal_out_dev,
                                                     p->k, newest_old, (float)p->clip);
        p->launches += 1;
    }
    CK(cudaGetLastError());
    p->age_cur ^= 1;  // returns_kernel wrote the new ages; done_corr / terminal above used the old ones
    return launch_normalize(p, out_dev, 0, st);
}

int hlynr_post_original(hlynr_post_t* p, float* out_dev, void* stream) {
    if (!p || !out_dev) return fail("hlynr_post_original: null argument");
    if (p->t == 0) return fail("hlynr_post_original: no observation yet");
    DeviceGuard g(p->device);
    return launch_normalize(p, out_dev, 1, (cudaStream_t)stream);
}

int hlynr_post_normalize(hlynr_post_t* p, const float* stacked_dev, int64_t rows, float* out_dev, void* stream) {
    if (!p || !stacked_dev || !out_dev) return fail("hlynr_post_normalize: null argument");
    if (rows <= 0) return 0;
    DeviceGuard g(p->device);
    const int lanes = 512 / p->D > 0 ? 512 / p->D : 1;
    normalize_rows_kernel<<<blocks_for(p, rows, lanes), lanes * p->D, 0, (cudaStream_t)stream>>>(stacked_dev, out_dev, rows, p->D, p->st, (float)p->clip);
    CK(cudaGetLastError());
    p->launches += 1;
    return 0;
}

int hlynr_post_get_stats(hlynr_post_t* p, double* mean, double* var, double* count, double* ret_mean, double* ret_var,
                         double* ret_count, void* stream) {
    if (!p) return fail("hlynr_post_get_stats: null handle");
    DeviceGuard g(p->device);
    PostStats* h = new (std::nothrow) PostStats();
    if (!h) return fail("out of host memory");
    cudaError_t e = cudaMemcpyAsync(h, p->st, sizeof(*h), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) { delete h; return fail("hlynr_post_get_stats: %s", cudaGetErrorString(e)); }
    if (mean) memcpy(mean, h->mean, sizeof(double) * p->D);
    if (var) memcpy(var, h->var, sizeof(double) * p->D);
    if (count) *count = h->count;
    if (ret_mean) *ret_mean = h->ret_mean;
    if (ret_var) *ret_var = h->ret_var;
    if (ret_count) *ret_count = h->ret_count;
    delete h;
    return 0;
}

int hlynr_post_set_stats(hlynr_post_t* p, const double* mean, const double* var, double count, double ret_mean, double ret_var,
                         double ret_count, void* stream) {
    if (!p || !mean || !var) return fail("hlynr_post_set_stats: null argument");
    DeviceGuard g(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(p->st->mean, mean, sizeof(double) * p->D, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(p->st->var, var, sizeof(double) * p->D, cudaMemcpyHostToDevice, st));
    const double tail[4] = {count, ret_mean, ret_var, ret_count};
    CK(cudaMemcpyAsync(&p->st->count, &tail[0], sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(&p->st->ret_mean, &tail[1], sizeof(double) * 3, cudaMemcpyHostToDevice, st));
    merge_kernel<<<1, 256, 0, st>>>(p->st, p->k, (double)p->n, 0, 0, 0, 0, p->eps);  // refresh inv_std
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));  // `tail` is a stack array
    p->launches += 1;
    return 0;
}

int hlynr_post_check_sums(hlynr_post_t* p, int resync, double* max_abs_diff, void* stream) {
    if (!p || !max_abs_diff) return fail("hlynr_post_check_sums: null argument");
    if (p->t == 0) return fail("hlynr_post_check_sums: no observation yet");
    DeviceGuard g(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(p->st->chk_S, 0, sizeof(double) * KMAX * OD * 2, st));
    const int lanes = 512 / p->D > 0 ? 512 / p->D : 1;
    recompute_sums_kernel<<<blocks_for(p, p->n, lanes), lanes * p->D, 0, st>>>(p->frames, p->plane, p->age[p->age_cur], p->n, p->k,
                                                                              (int)((p->t - 1) % (uint32_t)p->k), p->st);
    CK(cudaGetLastError());
    p->launches += 1;
    PostStats* h = new (std::nothrow) PostStats();
    if (!h) return fail("out of host memory");
    cudaError_t e = cudaMemcpyAsync(h, p->st, sizeof(*h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { delete h; return fail("hlynr_post_check_sums: %s", cudaGetErrorString(e)); }
    double m = 0.0;
    for (int d = 0; d < p->k; ++d)
        for (int c = 0; c < OD; ++c) {
            m = fmax(m, fabs(h->S[d][c] - h->chk_S[d][c]));
            m = fmax(m, fabs(h->Q[d][c] - h->chk_Q[d][c]));
        }
    *max_abs_diff = m;
    delete h;
    if (resync) {
        CK(cudaMemcpyAsync(p->st->S, p->st->chk_S, sizeof(double) * KMAX * OD, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(p->st->Q, p->st->chk_Q, sizeof(double) * KMAX * OD, cudaMemcpyDeviceToDevice, st));
        p->sums_valid = true;
    }
    return 0;
}

}  // extern "C"
