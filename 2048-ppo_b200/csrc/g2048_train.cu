// g2048_train.cu -- rewards-to-go / advantage scan and the fused PPO-clip + critic + entropy
// loss (forward and analytic backward w.r.t. logits and value).  sm_100a only.
//
// Reference (file:line in RobotSail/2048-PPO):
//   reward, discounted scan, normalisation, advantage   train.py:698-772
//   batch moments for the EMA update                    train.py:732-739, 898-901
//   loss                                                train.py:497-554
//
// Both kernels are single-pass HBM-bound reductions over time-major [T,B] rollout buffers
// (index t*B + b, so a warp reads 32 consecutive envs).  The scan runs in float64 like the
// reference's Python floats and rounds once to float32 (train.py:399-400).
#include <cfloat>
#include <cmath>
#include "g2048_device.cuh"
#include "g2048_host.h"
#include "g2048_loss.cuh"

namespace g2048 {

constexpr uint32_t ROLL_VALID = 0x80u;   // flags bit 7: this [t,b] slot holds a recorded move

// ------------------------------------------------------------------ block reduction (f64)
template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* __restrict__ partials) {
    __shared__ double sh[32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if (lane == 0) sh[warp][k] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = lane < nwarps ? sh[lane][k] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) partials[size_t(blockIdx.x) * NV + k] = x;
        }
    }
}

// fixed-order final sum of the per-block partials (deterministic, no atomics)
template <int NV>
__global__ void final_reduce_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ out) {
    __shared__ double sh[256];
    for (int k = 0; k < NV; ++k) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nblocks; i += blockDim.x) s += partials[size_t(i) * NV + k];
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (int(threadIdx.x) < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[k] = sh[0];
        __syncthreads();
    }
}

// ------------------------------------------------------------------ rewards-to-go + advantage
// One thread per env column, walking t = T-1 .. 0.
//   reward_t = w_points*points + w_mono*(gamma*mono_after - mono_before) + w_empt*(gamma*empt_after - empt_before)
//   with mono_after = empt_after = 0 on the move that ended the game (train.py:318-322)
//   G_t = reward_t + gamma*G_{t+1}, restarting from 0 after a DONE move or an invalid slot
//   g_norm = (G - mu_c) / (sd + 1e-8),  adv = g_norm - V_rollout
// partials: [gridDim.x][3] = {sum G, sum G^2, count} over valid slots.
__global__ void __launch_bounds__(128)
rtg_adv_kernel(const int32_t* __restrict__ points, const uint64_t* __restrict__ shaping,
               const uint8_t* __restrict__ flags, const float* __restrict__ value, int T, int64_t B, double gamma,
               double w_points, double w_mono, double w_empt, double mu_c, double inv_sd, const float* __restrict__ bootstrap,
               float* __restrict__ reward_out, float* __restrict__ g_raw_out, float* __restrict__ g_norm_out,
               float* __restrict__ adv_out, double* __restrict__ partials) {
    int64_t b = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    double acc[3] = {0.0, 0.0, 0.0};
    if (b < B) {
        // return-to-go after the last slot: 0 as in the reference (an episode cut by max_steps, train.py:724-728), or the
        // caller's estimate for a game that continues past the buffer; a DONE move or an invalid slot resets it either way
        double g = bootstrap ? double(bootstrap[b]) : 0.0;
#pragma unroll 4
        for (int t = T - 1; t >= 0; --t) {
            const int64_t i = int64_t(t) * B + b;
            const uint32_t f = flags[i];
            if (!(f & ROLL_VALID)) {
                g = 0.0;
                if (reward_out) reward_out[i] = 0.f;
                if (g_raw_out) g_raw_out[i] = 0.f;
                g_norm_out[i] = 0.f;
                adv_out[i] = 0.f;
                continue;
            }
            const uint64_t s = shaping[i];
            const bool done = f & FLAG_DONE;
            if (done) g = 0.0;
            const double mono_b = double(uint32_t(s) & 63u), mono_a = done ? 0.0 : double(uint32_t(s >> 6) & 63u);
            const double empt_b = double(uint32_t(s >> 12) & 31u), empt_a = done ? 0.0 : double(uint32_t(s >> 17) & 31u);
            const double r = double(points[i]) * w_points +
                             (w_mono * (gamma * mono_a - mono_b) + w_empt * (gamma * empt_a - empt_b));
            g = r + gamma * g;
            const double gn = (g - mu_c) * inv_sd;
            if (reward_out) reward_out[i] = float(r);
            if (g_raw_out) g_raw_out[i] = float(g);
            g_norm_out[i] = float(gn);
            adv_out[i] = float(gn - double(value[i]));
            acc[0] += g;
            acc[1] += g * g;
            acc[2] += 1.0;
        }
    }
    block_reduce_store<3>(acc, partials);
}

// ------------------------------------------------------------------ PPO-clip + critic + entropy
// Per sample (train.py:497-554), with l = logits, m = legal mask, a = action, A = advantage,
// R = normalised return-to-go, lp_old = rollout log-prob of a:
//   lp    = log_softmax(l masked to -inf)           rho = exp(clamp(lp[a] - lp_old, -20, 20))
//   ppo   = min(A*rho, A*clamp(rho, 1-eps, 1+eps))
//   q     = softmax(clamp(masked l, -20, 20))  (illegal logits become -20 inside the partition sum)
//   H     = -sum_{legal} q log q
//   vl    = smooth_l1(V, R)  (beta = 1)
//   u     = ppo - c_v*vl + beta_ent*H ;  loss = -(1/N) sum u
// Writes d loss / d logits and d loss / d V (already scaled by 1/N) and per-block partial sums
// {sum ppo, sum vl, sum H, count}.
__global__ void __launch_bounds__(256)
ppo_loss_kernel(const float4* __restrict__ logits, const float* __restrict__ value, const float* __restrict__ old_logp,
                int old_logp_stride, const uint8_t* __restrict__ actions, const uint8_t* __restrict__ legal,
                const uint8_t* __restrict__ flags, const float* __restrict__ adv, const float* __restrict__ g_norm,
                int64_t n, float clip_eps, float c_v, float beta_ent, float inv_n, float4* __restrict__ dlogits,
                float* __restrict__ dvalue, double* __restrict__ partials) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (flags && !(flags[i] & ROLL_VALID)) {
            dlogits[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            dvalue[i] = 0.f;
            continue;
        }
        const float4 l4 = logits[i];
        const float l[4] = {l4.x, l4.y, l4.z, l4.w};
        const uint32_t a = actions[i] & 3u;
        const float lp_old = old_logp[i * old_logp_stride + (old_logp_stride == 4 ? a : 0)];
        float gl[4], dv_unit, ppo, vl, H;
        ppo_sample(l, legal[i] & 15u, a, lp_old, adv[i], value[i], g_norm[i], clip_eps, beta_ent, ppo, vl, H, gl, dv_unit);
#pragma unroll
        for (int k = 0; k < 4; ++k) gl[k] *= -inv_n;
        dlogits[i] = make_float4(gl[0], gl[1], gl[2], gl[3]);
        dvalue[i] = inv_n * c_v * dv_unit;
        acc[0] += double(ppo);
        acc[1] += double(vl);
        acc[2] += double(H);
        acc[3] += 1.0;
    }
    block_reduce_store<4>(acc, partials);
}

// ------------------------------------------------------------------ KL(old || new) statistic (train.py:577-597, logged only)
// Per sample: sum over the legal moves of p_old (log p_old - log p_new), both distributions the masked softmax of their logits.
// partials: [gridDim.x][3] = {sum KL, count, max KL} over valid slots; the final kernel sums the first two in a fixed order
// and takes the maximum of the third.
__device__ __forceinline__ void masked_log_softmax4(const float (&l)[4], uint32_t legal, float (&lp)[4]) {
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (legal >> k & 1u) m = fmaxf(m, l[k]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (legal >> k & 1u) s += expf(l[k] - m);
    const float lse = m + logf(s);
#pragma unroll
    for (int k = 0; k < 4; ++k) lp[k] = l[k] - lse;
}
__global__ void __launch_bounds__(256)
masked_kl_kernel(const float4* __restrict__ old_logits, const float4* __restrict__ new_logits, const uint8_t* __restrict__ legal,
                 const uint8_t* __restrict__ flags, int64_t n, float* __restrict__ kl_out, double* __restrict__ partials) {
    double acc[2] = {0.0, 0.0};
    double mx = -INFINITY;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t lg = legal[i] & 15u;
        if ((flags && !(flags[i] & ROLL_VALID)) || lg == 0u) {
            if (kl_out) kl_out[i] = 0.f;
            continue;
        }
        const float4 o4 = old_logits[i], n4 = new_logits[i];
        const float lo_in[4] = {o4.x, o4.y, o4.z, o4.w}, ln_in[4] = {n4.x, n4.y, n4.z, n4.w};
        float lo[4], ln[4];
        masked_log_softmax4(lo_in, lg, lo);
        masked_log_softmax4(ln_in, lg, ln);
        float kl = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (lg >> k & 1u) kl += expf(lo[k]) * (lo[k] - ln[k]);
        if (kl_out) kl_out[i] = kl;
        acc[0] += double(kl);
        acc[1] += 1.0;
        mx = fmax(mx, double(kl));
    }
    __shared__ double shm[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) shm[warp] = mx;
    __syncthreads();
    if (warp == 0) {
        double x = lane < nwarps ? shm[lane] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
        if (lane == 0) partials[size_t(gridDim.x) * 2 + blockIdx.x] = x;
    }
    block_reduce_store<2>(acc, partials);
}
__global__ void final_reduce_kl_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ out) {
    __shared__ double sh[256];
    for (int k = 0; k < 2; ++k) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nblocks; i += blockDim.x) s += partials[size_t(i) * 2 + k];
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (int(threadIdx.x) < o) sh[threadIdx.x] += sh[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[k] = sh[0];
        __syncthreads();
    }
    double m = -INFINITY;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) m = fmax(m, partials[size_t(nblocks) * 2 + i]);
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (int(threadIdx.x) < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[2] = out[1] > 0.0 ? sh[0] : 0.0;
}

}  // namespace g2048

using namespace g2048;

extern "C" {

constexpr int64_t MAX_REDUCE_BLOCKS = 16384;

int64_t g2048_reduce_workspace_bytes(void) { return int64_t(sizeof(double)) * 4 * MAX_REDUCE_BLOCKS; }

int g2048_rtg_advantage(const int32_t* points, const uint64_t* shaping, const uint8_t* flags, const float* value,
                        int32_t T, int64_t B, double gamma, double w_points, double w_mono, double w_empt,
                        double mu_corrected, double stddev, float* reward_out, float* g_raw_out, float* g_norm_out,
                        float* adv_out, double* stats_out, void* workspace, void* stream) {
    return g2048_rtg_advantage_bootstrap(points, shaping, flags, value, nullptr, T, B, gamma, w_points, w_mono, w_empt, mu_corrected,
                                         stddev, reward_out, g_raw_out, g_norm_out, adv_out, stats_out, workspace, stream);
}

int g2048_rtg_advantage_bootstrap(const int32_t* points, const uint64_t* shaping, const uint8_t* flags, const float* value,
                                  const float* bootstrap, int32_t T, int64_t B, double gamma, double w_points, double w_mono,
                                  double w_empt, double mu_corrected, double stddev, float* reward_out, float* g_raw_out,
                                  float* g_norm_out, float* adv_out, double* stats_out, void* workspace, void* stream) {
    G2048_REQUIRE(T >= 0 && B >= 0, "g2048_rtg_advantage: negative shape");
    G2048_REQUIRE(stats_out != nullptr, "g2048_rtg_advantage: stats_out is NULL");
    cudaStream_t st = cudaStream_t(stream);
    if (T == 0 || B == 0) {       // empty rollout: nothing to read (the data pointers may be NULL)
        G2048_CHECK_CUDA(cudaMemsetAsync(stats_out, 0, 3 * sizeof(double), st));
        return G2048_OK;
    }
    G2048_REQUIRE(points && shaping && flags && value && g_norm_out && adv_out && workspace,
                  "g2048_rtg_advantage: NULL pointer argument");
    const int threads = 128;
    const int64_t blocks = (B + threads - 1) / threads;
    double* partials = static_cast<double*>(workspace);
    if (blocks > MAX_REDUCE_BLOCKS)
        return fail(G2048_EINVAL, "g2048_rtg_advantage: B=%lld exceeds the %lld-column limit of one call", (long long)B,
                    (long long)(MAX_REDUCE_BLOCKS * threads));
    rtg_adv_kernel<<<unsigned(blocks), threads, 0, st>>>(points, shaping, flags, value, T, B, gamma, w_points, w_mono,
                                                         w_empt, mu_corrected, 1.0 / (stddev + 1e-8), bootstrap, reward_out,
                                                         g_raw_out, g_norm_out, adv_out, partials);
    G2048_CHECK_LAUNCH("rtg_adv_kernel");
    final_reduce_kernel<3><<<1, 256, 0, st>>>(partials, int(blocks), stats_out);
    G2048_CHECK_LAUNCH("final_reduce_kernel");
    return G2048_OK;
}

int g2048_ppo_loss(const float* logits, const float* value, const float* old_logp, int32_t old_logp_stride,
                   const uint8_t* actions, const uint8_t* legal, const uint8_t* flags, const float* adv,
                   const float* g_norm, int64_t n, float clip_eps, float c_v, float beta_ent, float inv_n,
                   float* dlogits, float* dvalue, double* stats_out, void* workspace, void* stream) {
    G2048_REQUIRE(n >= 0, "g2048_ppo_loss: n < 0");
    G2048_REQUIRE(stats_out != nullptr, "g2048_ppo_loss: stats_out is NULL");
    G2048_REQUIRE(old_logp_stride == 1 || old_logp_stride == 4, "g2048_ppo_loss: old_logp_stride must be 1 or 4");
    cudaStream_t st = cudaStream_t(stream);
    if (n == 0) {                 // empty batch: the data pointers may be NULL
        G2048_CHECK_CUDA(cudaMemsetAsync(stats_out, 0, 4 * sizeof(double), st));
        return G2048_OK;
    }
    G2048_REQUIRE(logits && value && old_logp && actions && legal && adv && g_norm && dlogits && dvalue && workspace,
                  "g2048_ppo_loss: NULL pointer argument");
    const int threads = 256;
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t cap = int64_t(num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    if (blocks > MAX_REDUCE_BLOCKS) blocks = MAX_REDUCE_BLOCKS;
    double* partials = static_cast<double*>(workspace);
    ppo_loss_kernel<<<unsigned(blocks), threads, 0, st>>>(reinterpret_cast<const float4*>(logits), value, old_logp,
                                                          old_logp_stride, actions, legal, flags, adv, g_norm, n,
                                                          clip_eps, c_v, beta_ent, inv_n,
                                                          reinterpret_cast<float4*>(dlogits), dvalue, partials);
    G2048_CHECK_LAUNCH("ppo_loss_kernel");
    final_reduce_kernel<4><<<1, 256, 0, st>>>(partials, int(blocks), stats_out);
    G2048_CHECK_LAUNCH("final_reduce_kernel");
    return G2048_OK;
}

int g2048_masked_kl(const float* old_logits, const float* new_logits, const uint8_t* legal, const uint8_t* flags, int64_t n,
                    float* kl_out, double* stats_out, void* workspace, void* stream) {
    G2048_REQUIRE(n >= 0, "g2048_masked_kl: n < 0");
    G2048_REQUIRE(stats_out != nullptr, "g2048_masked_kl: stats_out is NULL");
    cudaStream_t st = cudaStream_t(stream);
    if (n == 0) {                 // empty batch: the data pointers may be NULL
        G2048_CHECK_CUDA(cudaMemsetAsync(stats_out, 0, 3 * sizeof(double), st));
        return G2048_OK;
    }
    G2048_REQUIRE(old_logits && new_logits && legal && workspace, "g2048_masked_kl: NULL pointer argument");
    G2048_REQUIRE(((reinterpret_cast<uintptr_t>(old_logits) | reinterpret_cast<uintptr_t>(new_logits)) & 15) == 0,
                  "g2048_masked_kl: the [n,4] logits must be 16-byte aligned");
    const int threads = 256;
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t cap = int64_t(num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    double* partials = static_cast<double*>(workspace);
    masked_kl_kernel<<<unsigned(blocks), threads, 0, st>>>(reinterpret_cast<const float4*>(old_logits),
                                                           reinterpret_cast<const float4*>(new_logits), legal, flags, n, kl_out, partials);
    G2048_CHECK_LAUNCH("masked_kl_kernel");
    final_reduce_kl_kernel<<<1, 256, 0, st>>>(partials, int(blocks), stats_out);
    G2048_CHECK_LAUNCH("final_reduce_kl_kernel");
    return G2048_OK;
}

}  // extern "C"
