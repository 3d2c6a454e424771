// g2048_update.cu -- fused elementwise part of the policy update: y = res + ReLU(LayerNorm(z)),
// forward and backward, for the GameMLP trunk (game.py:1038-1046 ResidualBlock, 1069-1073 stem).
//
// torch autograd keeps the GEMMs (cuBLAS) and the graph; these two kernels replace the chain
// LayerNorm -> ReLU -> (Dropout p=0) -> residual add and its backward (ATen's layer-norm kernels
// plus elementwise launches were 50 % of the update, profiles/r01_launches_bench.csv).
// HBM-bound streaming kernels: one warp per row, float4 accesses, rows are contiguous (h*4 bytes).
//   forward : read z (+res), write y, mean, rstd          -> (2|3)*h*4 + 8 bytes / row
//   backward: read z, g, mean, rstd, write dz             -> 3*h*4 + 8 bytes / row
//             dgamma / dbeta: per-lane register partials over the rows of a warp, reduced per block in
//             shared memory, written as per-block partials and summed in a fixed order (deterministic).
#include "g2048_host.h"

namespace g2048 {

constexpr int LN_WARPS = 8;
constexpr int LN_MAX_CHUNKS = 2;     // float4 chunks per lane: h <= 256

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(LN_WARPS * 32)
ln_relu_res_fwd_kernel(const float* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta,
                       const float* __restrict__ res, float* __restrict__ y, float* __restrict__ mean_out,
                       float* __restrict__ rstd_out, int64_t n, int h, float eps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nch = h >> 2;
    float4 g[LN_MAX_CHUNKS], b[LN_MAX_CHUNKS];
#pragma unroll
    for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
        const int j = lane + 32 * c;
        g[c] = j < nch ? reinterpret_cast<const float4*>(gamma)[j] : make_float4(0, 0, 0, 0);
        b[c] = j < nch ? reinterpret_cast<const float4*>(beta)[j] : make_float4(0, 0, 0, 0);
    }
    const float inv_h = 1.0f / float(h);
    for (int64_t row = int64_t(blockIdx.x) * LN_WARPS + warp; row < n; row += int64_t(gridDim.x) * LN_WARPS) {
        const float4* zr = reinterpret_cast<const float4*>(z + row * h);
        float4 v[LN_MAX_CHUNKS];
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
            const int j = lane + 32 * c;
            v[c] = j < nch ? __ldg(zr + j) : make_float4(0, 0, 0, 0);
            s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
        }
        const float mean = warp_sum(s) * inv_h;
        float q = 0.f;
#pragma unroll
        for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
            const int j = lane + 32 * c;
            if (j < nch) {
                v[c].x -= mean; v[c].y -= mean; v[c].z -= mean; v[c].w -= mean;
                q += (v[c].x * v[c].x + v[c].y * v[c].y) + (v[c].z * v[c].z + v[c].w * v[c].w);
            }
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_h + eps);
#pragma unroll
        for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
            const int j = lane + 32 * c;
            if (j < nch) {
                float4 o;
                o.x = fmaxf(fmaf(v[c].x * rstd, g[c].x, b[c].x), 0.f);
                o.y = fmaxf(fmaf(v[c].y * rstd, g[c].y, b[c].y), 0.f);
                o.z = fmaxf(fmaf(v[c].z * rstd, g[c].z, b[c].z), 0.f);
                o.w = fmaxf(fmaf(v[c].w * rstd, g[c].w, b[c].w), 0.f);
                if (res) {
                    const float4 r = __ldg(reinterpret_cast<const float4*>(res + row * h) + j);
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                reinterpret_cast<float4*>(y + row * h)[j] = o;
            }
        }
        if (lane == 0) {
            mean_out[row] = mean;
            rstd_out[row] = rstd;
        }
    }
}

__global__ void __launch_bounds__(LN_WARPS * 32)
ln_relu_res_bwd_kernel(const float* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta,
                       const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                       const float* __restrict__ gout, float* __restrict__ dz, float* __restrict__ partials,
                       int64_t n, int h) {
    __shared__ float red[LN_WARPS][2][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nch = h >> 2;
    float4 g[LN_MAX_CHUNKS], b[LN_MAX_CHUNKS], dg[LN_MAX_CHUNKS], db[LN_MAX_CHUNKS];
#pragma unroll
    for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
        const int j = lane + 32 * c;
        g[c] = j < nch ? reinterpret_cast<const float4*>(gamma)[j] : make_float4(0, 0, 0, 0);
        b[c] = j < nch ? reinterpret_cast<const float4*>(beta)[j] : make_float4(0, 0, 0, 0);
        dg[c] = make_float4(0, 0, 0, 0);
        db[c] = make_float4(0, 0, 0, 0);
    }
    const float inv_h = 1.0f / float(h);
    for (int64_t row = int64_t(blockIdx.x) * LN_WARPS + warp; row < n; row += int64_t(gridDim.x) * LN_WARPS) {
        const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
        float4 xh[LN_MAX_CHUNKS], a[LN_MAX_CHUNKS];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
            const int j = lane + 32 * c;
            xh[c] = make_float4(0, 0, 0, 0);
            a[c] = make_float4(0, 0, 0, 0);
            if (j < nch) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(z + row * h) + j);
                const float4 go = __ldg(reinterpret_cast<const float4*>(gout + row * h) + j);
                xh[c] = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
                float4 dy;   // gradient through ReLU: passes where the pre-activation is positive
                dy.x = fmaf(xh[c].x, g[c].x, b[c].x) > 0.f ? go.x : 0.f;
                dy.y = fmaf(xh[c].y, g[c].y, b[c].y) > 0.f ? go.y : 0.f;
                dy.z = fmaf(xh[c].z, g[c].z, b[c].z) > 0.f ? go.z : 0.f;
                dy.w = fmaf(xh[c].w, g[c].w, b[c].w) > 0.f ? go.w : 0.f;
                dg[c].x = fmaf(dy.x, xh[c].x, dg[c].x); dg[c].y = fmaf(dy.y, xh[c].y, dg[c].y);
                dg[c].z = fmaf(dy.z, xh[c].z, dg[c].z); dg[c].w = fmaf(dy.w, xh[c].w, dg[c].w);
                db[c].x += dy.x; db[c].y += dy.y; db[c].z += dy.z; db[c].w += dy.w;
                a[c] = make_float4(dy.x * g[c].x, dy.y * g[c].y, dy.z * g[c].z, dy.w * g[c].w);
                s1 += (a[c].x + a[c].y) + (a[c].z + a[c].w);
                s2 += (a[c].x * xh[c].x + a[c].y * xh[c].y) + (a[c].z * xh[c].z + a[c].w * xh[c].w);
            }
        }
        const float m1 = warp_sum(s1) * inv_h, m2 = warp_sum(s2) * inv_h;
#pragma unroll
        for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
            const int j = lane + 32 * c;
            if (j < nch) {
                float4 o;
                o.x = rstd * (a[c].x - m1 - xh[c].x * m2);
                o.y = rstd * (a[c].y - m1 - xh[c].y * m2);
                o.z = rstd * (a[c].z - m1 - xh[c].z * m2);
                o.w = rstd * (a[c].w - m1 - xh[c].w * m2);
                reinterpret_cast<float4*>(dz + row * h)[j] = o;
            }
        }
    }
    // per-block reduction of the column sums, then one partial row per block
#pragma unroll
    for (int c = 0; c < LN_MAX_CHUNKS; ++c) {
        const int j = lane + 32 * c;
        if (j < nch) {
            reinterpret_cast<float4*>(red[warp][0])[j] = dg[c];
            reinterpret_cast<float4*>(red[warp][1])[j] = db[c];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * h; i += blockDim.x) {
        const int which = i / h, col = i % h;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) s += red[w][which][col];
        partials[(size_t(blockIdx.x) * 2 + which) * h + col] = s;
    }
}

// dgamma[h], dbeta[h] = fixed-order sum of the per-block partials
__global__ void ln_param_grad_reduce_kernel(const float* __restrict__ partials, int nblocks, int h,
                                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * h) return;
    const int which = i / h, col = i % h;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += partials[(size_t(b) * 2 + which) * h + col];
    (which == 0 ? dgamma : dbeta)[col] = s;
}

static int ln_grid(int64_t n) {
    int64_t blocks = (n + LN_WARPS - 1) / LN_WARPS;
    const int64_t cap = int64_t(num_sms()) * 8;
    return int(blocks < cap ? blocks : cap);
}

}  // namespace g2048

using namespace g2048;

extern "C" {

int64_t g2048_ln_workspace_floats(int32_t h) { return int64_t(148 * 8 + 64) * 2 * h; }

int g2048_ln_relu_res_fwd(const float* z, const float* gamma, const float* beta, const float* res, float* y,
                          float* mean, float* rstd, int64_t n, int32_t h, float eps, void* stream) {
    G2048_REQUIRE(n >= 0, "g2048_ln_relu_res_fwd: n < 0");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(z && gamma && beta && y && mean && rstd, "g2048_ln_relu_res_fwd: NULL pointer argument");
    if (h < 4 || h > 256 || (h & 3)) return fail(G2048_ESHAPE, "g2048_ln_relu_res_fwd: h=%d must be a multiple of 4 in [4,256]", h);
    ln_relu_res_fwd_kernel<<<ln_grid(n), LN_WARPS * 32, 0, cudaStream_t(stream)>>>(z, gamma, beta, res, y, mean, rstd, n, h, eps);
    G2048_CHECK_LAUNCH("ln_relu_res_fwd_kernel");
    return G2048_OK;
}

int g2048_ln_relu_res_bwd(const float* z, const float* gamma, const float* beta, const float* mean, const float* rstd,
                          const float* gout, float* dz, float* dgamma, float* dbeta, float* workspace, int64_t n,
                          int32_t h, void* stream) {
    G2048_REQUIRE(n >= 0, "g2048_ln_relu_res_bwd: n < 0");
    G2048_REQUIRE(dgamma && dbeta, "g2048_ln_relu_res_bwd: NULL gradient output");
    if (h < 4 || h > 256 || (h & 3)) return fail(G2048_ESHAPE, "g2048_ln_relu_res_bwd: h=%d must be a multiple of 4 in [4,256]", h);
    cudaStream_t st = cudaStream_t(stream);
    G2048_REQUIRE(n == 0 || (z && gamma && beta && mean && rstd && gout && dz && workspace),
                  "g2048_ln_relu_res_bwd: NULL pointer argument");
    if (n == 0) {
        G2048_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * h, st));
        G2048_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * h, st));
        return G2048_OK;
    }
    const int grid = ln_grid(n);
    ln_relu_res_bwd_kernel<<<grid, LN_WARPS * 32, 0, st>>>(z, gamma, beta, mean, rstd, gout, dz, workspace, n, h);
    G2048_CHECK_LAUNCH("ln_relu_res_bwd_kernel");
    ln_param_grad_reduce_kernel<<<(2 * h + 127) / 128, 128, 0, st>>>(workspace, grid, h, dgamma, dbeta);
    G2048_CHECK_LAUNCH("ln_param_grad_reduce_kernel");
    return G2048_OK;
}

}  // extern "C"
