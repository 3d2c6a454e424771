// g2048_urm_train.cu -- the GameURM-specific pieces of the policy update, forward AND hand-written backward:
//   attention over the 16 cells of a board (4 heads, head_dim 16, no mask)        game.py:1296-1317
//   ConvSwiGLU inner chain  silu(gate) * up -> depthwise conv (k = 2) -> silu     game.py:1264-1276
//   hidden = rms_norm(hidden + branch)                                             game.py:1223-1229, 1345-1350
// The four projections of a block run on the split-operand tcgen05 GEMM / weight-gradient kernels (g2048_linear.cu); these
// kernels are the rest of a block, so that a GameURM update step runs on this library in both directions (torch autograd is only
// the tape).  They are HBM-bound streaming kernels in fp32; nothing is saved between the passes except the op inputs -- the
// backward kernels recompute the softmax / SiLUs.  Parameter gradients of the conv are reduced in a fixed order (deterministic).
// sm_100a only.
#include <cmath>
#include "g2048_host.h"

namespace g2048 {
namespace urmt {

constexpr int SEQ = 16, NHEAD = 4, HD = 16, H = 64, INTER = 120;

// hardware ex2 / rcp approximations (~1e-6 relative): the IEEE divide and expf were most of the ConvSwiGLU kernels' instructions
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
// d silu(x) / dx = s (1 + x (1 - s)), s = sigmoid(x)
__device__ __forceinline__ float dsiluf_(float x) {
    const float s = sigmoidf_(x);
    return s * fmaf(x, 1.0f - s, 1.0f);
}

// ---------------------------------------------------------------------------------------------------------------- attention
// One block = one env: 64 threads = (head, query token).  The env's q | k | v rows (16 x 192 floats) are staged in shared memory.
// qkv layout per token: [q (64) | k (64) | v (64)], head h at columns 16 h .. 16 h + 15 of each third (game.py:1306-1307).
struct AttnTile {
    float qkv[SEQ][3 * H + 4];            // + 4: rows 16 bytes apart from a multiple of 32 banks
};

__device__ __forceinline__ void load_env_rows(float* dst, int dst_stride, const float* __restrict__ src, int row_floats, int tid, int nthreads) {
    for (int i = tid; i < SEQ * row_floats / 4; i += nthreads) {
        const int r = (i * 4) / row_floats, c = (i * 4) % row_floats;
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
        *reinterpret_cast<float4*>(dst + r * dst_stride + c) = v;
    }
}

// softmax row of query i of head h over the 16 keys (scale 1/sqrt(head_dim) = 0.25)
__device__ __forceinline__ void prob_row(const AttnTile& T, int h, int i, float (&p)[SEQ]) {
    float q[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] = T.qkv[i][h * HD + d] * 0.25f;
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < SEQ; ++j) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) s = fmaf(q[d], T.qkv[j][H + h * HD + d], s);
        p[j] = s;
        mx = fmaxf(mx, s);
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < SEQ; ++j) {
        p[j] = expf(p[j] - mx);
        den += p[j];
    }
    const float inv = 1.0f / den;
#pragma unroll
    for (int j = 0; j < SEQ; ++j) p[j] *= inv;
}

__global__ void __launch_bounds__(64) attn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, int64_t B) {
    __shared__ __align__(16) AttnTile T;
    const int tid = threadIdx.x, h = tid >> 4, i = tid & 15;
    for (int64_t env = blockIdx.x; env < B; env += gridDim.x) {
        __syncthreads();
        load_env_rows(&T.qkv[0][0], 3 * H + 4, qkv + env * SEQ * 3 * H, 3 * H, tid, 64);
        __syncthreads();
        float p[SEQ];
        prob_row(T, h, i, p);
        float o[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
        for (int j = 0; j < SEQ; ++j)
#pragma unroll
            for (int d = 0; d < HD; ++d) o[d] = fmaf(p[j], T.qkv[j][2 * H + h * HD + d], o[d]);
        float4* dst = reinterpret_cast<float4*>(out + (env * SEQ + i) * H + h * HD);
#pragma unroll
        for (int d = 0; d < HD; d += 4) dst[d / 4] = make_float4(o[d], o[d + 1], o[d + 2], o[d + 3]);
    }
}

// dQ_i = 0.25 sum_j dS_ij K_j, dK_j = 0.25 sum_i dS_ij Q_i, dV_j = sum_i P_ij dO_i, with dS_ij = P_ij (dP_ij - sum_k P_ik dP_ik),
// dP_ij = dO_i . V_j.  Thread (h, i) works out row i of P and dS (kept in shared memory), then column i of both.
struct AttnBwdTile {
    AttnTile t;
    float dout[SEQ][H + 4];
    float P[NHEAD][SEQ][SEQ + 1];
    float dS[NHEAD][SEQ][SEQ + 1];
};

__global__ void __launch_bounds__(64, 8) attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dout,
                                                         float* __restrict__ dqkv, int64_t B) {
    __shared__ __align__(16) AttnBwdTile T;
    const int tid = threadIdx.x, h = tid >> 4, i = tid & 15;
    for (int64_t env = blockIdx.x; env < B; env += gridDim.x) {
        __syncthreads();
        load_env_rows(&T.t.qkv[0][0], 3 * H + 4, qkv + env * SEQ * 3 * H, 3 * H, tid, 64);
        load_env_rows(&T.dout[0][0], H + 4, dout + env * SEQ * H, H, tid, 64);
        __syncthreads();
        // row i of head h, key by key (the j loops are NOT unrolled: scores, probabilities and dP pass through this thread's own
        // rows of the shared-memory P / dS arrays instead of 3 x 16 registers with every K / V load hoisted above them)
        float q[HD], g[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            q[d] = T.t.qkv[i][h * HD + d] * 0.25f;
            g[d] = T.dout[i][h * HD + d];
        }
        float mx = -INFINITY;
#pragma unroll 2
        for (int j = 0; j < SEQ; ++j) {
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) {
                s = fmaf(q[d], T.t.qkv[j][H + h * HD + d], s);
                dp = fmaf(g[d], T.t.qkv[j][2 * H + h * HD + d], dp);
            }
            T.P[h][i][j] = s;
            T.dS[h][i][j] = dp;
            mx = fmaxf(mx, s);
        }
        float den = 0.f, delta = 0.f;
#pragma unroll 4
        for (int j = 0; j < SEQ; ++j) {
            const float e = expf(T.P[h][i][j] - mx);
            T.P[h][i][j] = e;
            den += e;
            delta = fmaf(e, T.dS[h][i][j], delta);
        }
        const float inv = 1.0f / den;
        delta *= inv;
        float dq[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) dq[d] = 0.f;
#pragma unroll 2
        for (int j = 0; j < SEQ; ++j) {
            const float pj = T.P[h][i][j] * inv;
            const float ds = pj * (T.dS[h][i][j] - delta);
            T.P[h][i][j] = pj;
            T.dS[h][i][j] = ds;
#pragma unroll
            for (int d = 0; d < HD; ++d) dq[d] = fmaf(ds, T.t.qkv[j][H + h * HD + d], dq[d]);
        }
        float* row = dqkv + (env * SEQ + i) * 3 * H;
#pragma unroll
        for (int d = 0; d < HD; d += 4)
            *reinterpret_cast<float4*>(row + h * HD + d) = make_float4(dq[d] * 0.25f, dq[d + 1] * 0.25f, dq[d + 2] * 0.25f, dq[d + 3] * 0.25f);
        __syncthreads();
        // column i: key / value token i
        float dk[HD], dv[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) dk[d] = dv[d] = 0.f;
#pragma unroll 2
        for (int r = 0; r < SEQ; ++r) {
            const float ds = T.dS[h][r][i], pr = T.P[h][r][i];
#pragma unroll
            for (int d = 0; d < HD; ++d) {
                dk[d] = fmaf(ds, T.t.qkv[r][h * HD + d], dk[d]);
                dv[d] = fmaf(pr, T.dout[r][h * HD + d], dv[d]);
            }
        }
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            *reinterpret_cast<float4*>(row + H + h * HD + d) = make_float4(dk[d] * 0.25f, dk[d + 1] * 0.25f, dk[d + 2] * 0.25f, dk[d + 3] * 0.25f);
            *reinterpret_cast<float4*>(row + 2 * H + h * HD + d) = make_float4(dv[d], dv[d + 1], dv[d + 2], dv[d + 3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------- ConvSwiGLU
// x = silu(gate) * up; c[t] = w0 x[t-1] + w1 x[t] + b (the conv of game.py:1253-1260 is kernel 2, padding 1, trimmed to 16
// outputs: out[t] sees x[t-1] and x[t]); y = silu(c).  gate / up / y: [B, 16, 120]; conv weight [120, 2] (w0 = [:, 0]).
// One thread = (env, channel), walking the 16 tokens; a warp touches 32 consecutive channels of one token row (coalesced).
constexpr int SW_THREADS = 128;          // channels (120 used)

__global__ void __launch_bounds__(SW_THREADS) swiglu_fwd_kernel(const float* __restrict__ gate, const float* __restrict__ up,
                                                                const float* __restrict__ cw, const float* __restrict__ cb,
                                                                float* __restrict__ y, int64_t B) {
    const int c = threadIdx.x;
    if (c >= INTER) return;
    const float w0 = cw[2 * c], w1 = cw[2 * c + 1], b = cb[c];
    for (int64_t env = blockIdx.x; env < B; env += gridDim.x) {
        const int64_t base = env * SEQ * INTER + c;
        float prev = 0.f;
#pragma unroll
        for (int t = 0; t < SEQ; ++t) {
            const float x = siluf_(__ldg(gate + base + t * INTER)) * __ldg(up + base + t * INTER);
            y[base + t * INTER] = siluf_(fmaf(w0, prev, fmaf(w1, x, b)));
            prev = x;
        }
    }
}

// dc[t] = dy[t] silu'(c[t]); dx[t] = w1 dc[t] + w0 dc[t+1]; dgate = dx up silu'(gate); dup = dx silu(gate);
// dw0 = sum dc[t] x[t-1], dw1 = sum dc[t] x[t], db = sum dc[t]  (per-block partials [3][128], reduced by conv_reduce_kernel)
__global__ void __launch_bounds__(SW_THREADS) swiglu_bwd_kernel(const float* __restrict__ gate, const float* __restrict__ up,
                                                                const float* __restrict__ cw, const float* __restrict__ cb,
                                                                const float* __restrict__ dy, float* __restrict__ dgate,
                                                                float* __restrict__ dup, float* __restrict__ partials, int64_t B) {
    const int c = threadIdx.x;
    float a0 = 0.f, a1 = 0.f, ab = 0.f;
    if (c < INTER) {
        const float w0 = cw[2 * c], w1 = cw[2 * c + 1], b = cb[c];
        for (int64_t env = blockIdx.x; env < B; env += gridDim.x) {
            const int64_t base = env * SEQ * INTER + c;
            float x[SEQ], dc[SEQ];
            float prev = 0.f;
#pragma unroll
            for (int t = 0; t < SEQ; ++t) {
                x[t] = siluf_(__ldg(gate + base + t * INTER)) * __ldg(up + base + t * INTER);
                const float cc = fmaf(w0, prev, fmaf(w1, x[t], b));
                dc[t] = __ldg(dy + base + t * INTER) * dsiluf_(cc);
                a0 = fmaf(dc[t], prev, a0);
                a1 = fmaf(dc[t], x[t], a1);
                ab += dc[t];
                prev = x[t];
            }
#pragma unroll
            for (int t = 0; t < SEQ; ++t) {
                const float dx = fmaf(w1, dc[t], t + 1 < SEQ ? w0 * dc[t + 1] : 0.f);
                const float g = __ldg(gate + base + t * INTER), u = __ldg(up + base + t * INTER);
                dgate[base + t * INTER] = dx * u * dsiluf_(g);
                dup[base + t * INTER] = dx * siluf_(g);
            }
        }
    }
    float* pp = partials + size_t(blockIdx.x) * 3 * SW_THREADS;
    pp[c] = a0;
    pp[SW_THREADS + c] = a1;
    pp[2 * SW_THREADS + c] = ab;
}

// fixed-order sum of the per-block partials into dcw [120, 2], dcb [120]
// (block k = which sum; four independent accumulators keep loads in flight, combined in a fixed order)
__global__ void conv_reduce_kernel(const float* __restrict__ partials, int nblocks, float* __restrict__ dcw, float* __restrict__ dcb) {
    const int c = threadIdx.x, k = blockIdx.x;
    if (c >= INTER) return;
    const float* pp = partials + k * SW_THREADS + c;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    int i = 0;
    for (; i + 4 <= nblocks; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] += __ldg(pp + size_t(i + u) * 3 * SW_THREADS);
    }
    for (; i < nblocks; ++i) a[0] += __ldg(pp + size_t(i) * 3 * SW_THREADS);
    const float sum = (a[0] + a[1]) + (a[2] + a[3]);
    if (k == 0) dcw[2 * c] = sum;
    else if (k == 1) dcw[2 * c + 1] = sum;
    else dcb[c] = sum;
}

// ---------------------------------------------------------------------------------------------------------------- RMS norm + residual
// y = s * rsqrt(mean(s^2) + eps), s = x + r (game.py:1223-1229); rs = the row's rsqrt factor, saved for the backward.
// backward (the same gradient flows to x and r): ds = rs (dy - y mean(dy . y)).
// One warp = one row of 64 floats (two per lane, coalesced float2).
__global__ void __launch_bounds__(256) norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ r, float* __restrict__ y,
                                                       float* __restrict__ rs, int64_t rows, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nw = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = w0; row < rows; row += nw) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(x + row * H) + lane), b = __ldg(reinterpret_cast<const float2*>(r + row * H) + lane);
        const float s0 = a.x + b.x, s1 = a.y + b.y;
        float sq = fmaf(s0, s0, s1 * s1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float f = rsqrtf(sq * (1.0f / H) + eps);
        reinterpret_cast<float2*>(y + row * H)[lane] = make_float2(s0 * f, s1 * f);
        if (lane == 0) rs[row] = f;
    }
}

__global__ void __launch_bounds__(256) norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ rs, const float* __restrict__ dy,
                                                       float* __restrict__ ds, int64_t rows) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, nw = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = w0; row < rows; row += nw) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(y + row * H) + lane), g = __ldg(reinterpret_cast<const float2*>(dy + row * H) + lane);
        float dot = fmaf(v.x, g.x, v.y * g.y);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        const float m = dot * (1.0f / H), f = __ldg(rs + row);
        reinterpret_cast<float2*>(ds + row * H)[lane] = make_float2(f * (g.x - v.x * m), f * (g.y - v.y * m));
    }
}

static int grid_for(int64_t units, int per_sm) {
    const int64_t cap = int64_t(num_sms()) * per_sm;
    return int(units < cap ? (units > 0 ? units : 1) : cap);
}

}  // namespace urmt
}  // namespace g2048

using namespace g2048;

extern "C" {

int g2048_urm_attn_fwd(const float* qkv, float* out, int64_t B, void* stream) {
    G2048_REQUIRE(B >= 0, "g2048_urm_attn_fwd: negative batch");
    if (B == 0) return G2048_OK;
    G2048_REQUIRE(qkv && out, "g2048_urm_attn_fwd: NULL pointer argument");
    urmt::attn_fwd_kernel<<<urmt::grid_for(B, 16), 64, 0, cudaStream_t(stream)>>>(qkv, out, B);
    G2048_CHECK_LAUNCH("urmt::attn_fwd_kernel");
    return G2048_OK;
}

int g2048_urm_attn_bwd(const float* qkv, const float* dout, float* dqkv, int64_t B, void* stream) {
    G2048_REQUIRE(B >= 0, "g2048_urm_attn_bwd: negative batch");
    if (B == 0) return G2048_OK;
    G2048_REQUIRE(qkv && dout && dqkv, "g2048_urm_attn_bwd: NULL pointer argument");
    urmt::attn_bwd_kernel<<<urmt::grid_for(B, 8), 64, 0, cudaStream_t(stream)>>>(qkv, dout, dqkv, B);
    G2048_CHECK_LAUNCH("urmt::attn_bwd_kernel");
    return G2048_OK;
}

int g2048_urm_swiglu_fwd(const float* gate, const float* up, const float* conv_w, const float* conv_b, float* y, int64_t B, void* stream) {
    G2048_REQUIRE(B >= 0, "g2048_urm_swiglu_fwd: negative batch");
    if (B == 0) return G2048_OK;
    G2048_REQUIRE(gate && up && conv_w && conv_b && y, "g2048_urm_swiglu_fwd: NULL pointer argument");
    urmt::swiglu_fwd_kernel<<<urmt::grid_for(B, 16), urmt::SW_THREADS, 0, cudaStream_t(stream)>>>(gate, up, conv_w, conv_b, y, B);
    G2048_CHECK_LAUNCH("urmt::swiglu_fwd_kernel");
    return G2048_OK;
}

int64_t g2048_urm_swiglu_workspace_floats(void) { return int64_t(num_sms()) * 8 * 3 * urmt::SW_THREADS; }

int g2048_urm_swiglu_bwd(const float* gate, const float* up, const float* conv_w, const float* conv_b, const float* dy, float* dgate,
                         float* dup, float* dconv_w, float* dconv_b, float* workspace, int64_t B, void* stream) {
    G2048_REQUIRE(B >= 0, "g2048_urm_swiglu_bwd: negative batch");
    G2048_REQUIRE(gate && up && conv_w && conv_b && dy && dgate && dup && dconv_w && dconv_b && workspace,
                  "g2048_urm_swiglu_bwd: NULL pointer argument");
    const int grid = B == 0 ? 1 : urmt::grid_for(B, 8);
    urmt::swiglu_bwd_kernel<<<grid, urmt::SW_THREADS, 0, cudaStream_t(stream)>>>(gate, up, conv_w, conv_b, dy, dgate, dup, workspace, B);
    G2048_CHECK_LAUNCH("urmt::swiglu_bwd_kernel");
    urmt::conv_reduce_kernel<<<3, urmt::SW_THREADS, 0, cudaStream_t(stream)>>>(workspace, grid, dconv_w, dconv_b);
    G2048_CHECK_LAUNCH("urmt::conv_reduce_kernel");
    return G2048_OK;
}

int g2048_urm_norm_fwd(const float* x, const float* r, float* y, float* rs, int64_t rows, float eps, void* stream) {
    G2048_REQUIRE(rows >= 0, "g2048_urm_norm_fwd: negative row count");
    if (rows == 0) return G2048_OK;
    G2048_REQUIRE(x && r && y && rs, "g2048_urm_norm_fwd: NULL pointer argument");
    urmt::norm_fwd_kernel<<<urmt::grid_for((rows + 7) / 8, 8), 256, 0, cudaStream_t(stream)>>>(x, r, y, rs, rows, eps);
    G2048_CHECK_LAUNCH("urmt::norm_fwd_kernel");
    return G2048_OK;
}

int g2048_urm_norm_bwd(const float* y, const float* rs, const float* dy, float* ds, int64_t rows, void* stream) {
    G2048_REQUIRE(rows >= 0, "g2048_urm_norm_bwd: negative row count");
    if (rows == 0) return G2048_OK;
    G2048_REQUIRE(y && rs && dy && ds, "g2048_urm_norm_bwd: NULL pointer argument");
    urmt::norm_bwd_kernel<<<urmt::grid_for((rows + 7) / 8, 8), 256, 0, cudaStream_t(stream)>>>(y, rs, dy, ds, rows);
    G2048_CHECK_LAUNCH("urmt::norm_bwd_kernel");
    return G2048_OK;
}

}  // extern "C"
