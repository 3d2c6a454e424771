// g2048_rollout.cuh -- pieces shared by the fp32 (FFMA) and bf16 (tcgen05) fused rollout kernels:
// launch parameters, the packed-weight layout and the per-env policy + env-step tail.
#pragma once
#include <cmath>
#include "g2048_device.cuh"
#include "g2048_host.h"

namespace g2048 {

constexpr int RO_TILE = 128;          // envs per tile
constexpr int RO_CONSUMERS = 256;     // GEMM threads
constexpr int RO_THREADS = RO_CONSUMERS + 32;   // + producer warp
constexpr int RO_KC = 16;             // k rows per weight chunk
constexpr int RO_STAGES = 4;
constexpr uint64_t RESET_KEY_TWEAK = 0x9E3779B97F4A7C15ull;   // key stream of the auto-reset draws

struct RolloutParams {
    int64_t B;
    int32_t T, hidden, layers, auto_reset;
    uint64_t seed, env0, ctr0;
    const float* packed;
    const uint32_t* lut;
    uint64_t* boards;              // in/out [B]
    uint8_t* alive;                // in/out [B] or NULL
    const uint8_t* forced_actions; // [T,B] or NULL
    uint64_t* rec_boards;
    uint8_t* rec_actions;
    uint8_t* rec_legal;
    float* rec_logp;               // [T,B,4]
    float* rec_value;
    int32_t* rec_points;
    uint64_t* rec_shaping;
    uint8_t* rec_flags;
    float* rec_entropy;            // or NULL
    int32_t* sched;                // [ceil(B/128)] hand-off flags of the horizon segments (x3 kernel), zeroed before the launch; or NULL
    int32_t segs;                  // horizon segments per tile (1 = a CTA plays a tile's whole horizon)
};

// packed weight layout (floats), HP = padded hidden
__host__ __device__ inline int64_t pk_stem_w(int) { return 0; }
__host__ __device__ inline int64_t pk_stem_b0(int HP) { return int64_t(16) * HP; }
__host__ __device__ inline int64_t pk_stem_g(int HP) { return int64_t(17) * HP; }
__host__ __device__ inline int64_t pk_stem_beta(int HP) { return int64_t(18) * HP; }
__host__ __device__ inline int64_t pk_layer(int HP, int l) { return int64_t(19) * HP + int64_t(l) * HP * (HP + 2); }
__host__ __device__ inline int64_t pk_heads(int HP, int L) { return pk_layer(HP, L); }
__host__ __device__ inline int64_t pk_total(int HP, int L) { return pk_heads(HP, L) + 5 * HP + 8; }

__host__ __device__ inline int padded_hidden(int h) {
    if (h <= 64) return 64;
    if (h <= 128) return 128;
    if (h <= 192) return 192;
    if (h <= 208) return 208;
    return -1;
}


// bf16 operand images for the tensor-core kernel follow the fp32 section (128-byte aligned):
// stem image [HP rows x 128 B] (K = 16 exponents, k-step 0), then one [KB x HP rows x 128 B]
// image per residual block, all in the 128B-swizzled K-major layout of g2048_tc.cuh.
__host__ __device__ inline int kblocks_of(int HP) { return (HP + 63) / 64; }
__host__ __device__ inline int64_t pk_img_base(int HP, int L) { return (pk_total(HP, L) + 31) / 32 * 32; }   // floats
__host__ __device__ inline int64_t img_stem_bytes(int HP) { return int64_t(HP) * 128; }
__host__ __device__ inline int64_t img_layer_bytes(int HP) { return int64_t(kblocks_of(HP)) * HP * 128; }
__host__ __device__ inline int64_t pk_total_with_images(int HP, int L) {
    return pk_img_base(HP, L) + (img_stem_bytes(HP) + int64_t(L) * img_layer_bytes(HP)) / 4;
}

// Split-fp16 operand images for the fp32-grade tensor-core kernel (g2048_rollout_x3.cu) follow the bf16 images
// (256-byte aligned): k-blocks of 16 input features in the 32-byte-swizzled K-major layout of g2048_tc.cuh
// (sw32_offset), each block = [hi part: HP rows x 32 B | lo part: HP rows x 32 B], w = hi + lo in fp16 (22 mantissa
// bits).  Block 0 = the stem (16 exponent columns), then HP/16 blocks per residual block in k order.
__host__ __device__ inline int64_t x3_block_bytes(int HP) { return int64_t(HP) * 64; }
__host__ __device__ inline int64_t pk_x3_base(int HP, int L) { return (pk_total_with_images(HP, L) + 63) / 64 * 64; }   // floats
__host__ __device__ inline int64_t x3_blocks(int HP, int L) { return 1 + int64_t(L) * (HP / 16); }
__host__ __device__ inline int64_t pk_total_all(int HP, int L) { return pk_x3_base(HP, L) + x3_blocks(HP, L) * x3_block_bytes(HP) / 4; }

// Start-of-step bookkeeping for one env: legal mask of the current board; a terminal board is
// reset at once (auto_reset) or the env goes idle.  game.py:103-119, 942-950.
__device__ __forceinline__ uint32_t begin_step(const RolloutParams& p, int64_t env, uint64_t ctr, Board& board, bool& alive) {
    if (!alive) return 0u;
    uint32_t lm = legal_mask(board);
    if (lm == 0u) {
        if (p.auto_reset) {
            board = reset_board(env_draws(p.seed ^ RESET_KEY_TWEAK, p.env0 + uint64_t(env), ctr));
            lm = legal_mask(board);
        } else {
            alive = false;
        }
    }
    return lm;
}

// Masked log-softmax, categorical sample, env step and the [t, env] record for one env.
// o[0..3] = action logits, o[4] = value.  train.py:266-326.
__device__ __forceinline__ void policy_env_step(const RolloutParams& p, const LutGlobal& lut, int t, int64_t env,
                                                uint64_t ctr, uint32_t lm, const float (&o)[5], Board& board, bool& alive) {
    const int64_t ri = int64_t(t) * p.B + env;
    if (!alive) {
        p.rec_flags[ri] = 0;
        p.rec_boards[ri] = pack_board(board);
        p.rec_actions[ri] = 0;
        p.rec_legal[ri] = 0;
        p.rec_value[ri] = 0.f;
        p.rec_points[ri] = 0;
        p.rec_shaping[ri] = 0;
        reinterpret_cast<float4*>(p.rec_logp)[ri] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.rec_entropy) p.rec_entropy[ri] = 0.f;
        return;
    }
    // masked log-softmax (train.py:271-274, 326)
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if ((lm >> j) & 1u) mx = fmaxf(mx, o[j]);
    float e[4], se = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        e[j] = ((lm >> j) & 1u) ? expf(o[j] - mx) : 0.f;
        se += e[j];
    }
    const float lse = mx + logf(se);
    float lp[4], ent = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        lp[j] = ((lm >> j) & 1u) ? o[j] - lse : -INFINITY;
        const float pj = e[j] / se;
        if (pj > 0.f) ent -= pj * logf(pj);          // train.py:290-291
    }
    const U4 d = env_draws(p.seed, p.env0 + uint64_t(env), ctr);
    uint32_t a;
    if (p.forced_actions) {
        a = p.forced_actions[ri] & 3u;
    } else {
        // inverse-CDF categorical sample over the legal actions, 24-bit uniform from Philox word 2
        const float thr = float(d.z >> 8) * (1.0f / 16777216.0f) * se;
        float cum = 0.f;
        a = 31u - uint32_t(__clz(int(lm)));          // last legal action (round-off guard)
        bool found = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            cum += e[j];
            if (!found && ((lm >> j) & 1u) && thr < cum) {
                a = uint32_t(j);
                found = true;
            }
        }
    }
    const StepOut so = env_step<true>(board, a, d.x, d.y, lut);   // train.py:294
    p.rec_boards[ri] = pack_board(board);
    p.rec_actions[ri] = uint8_t(a);
    p.rec_legal[ri] = uint8_t(lm);
    reinterpret_cast<float4*>(p.rec_logp)[ri] = make_float4(lp[0], lp[1], lp[2], lp[3]);
    p.rec_value[ri] = o[4];
    p.rec_points[ri] = so.points;
    p.rec_shaping[ri] = uint64_t(so.shape_lo) | uint64_t(so.shape_hi) << 32;
    p.rec_flags[ri] = uint8_t(so.flags | 0x80u);
    if (p.rec_entropy) p.rec_entropy[ri] = ent;
    board = so.board;
    if (so.flags & FLAG_DONE) {
        if (p.auto_reset) board = reset_board(env_draws(p.seed ^ RESET_KEY_TWEAK, p.env0 + uint64_t(env), ctr));
        else alive = false;
    }
}

// bf16 / tcgen05 variant (g2048_rollout_tc.cu)
int launch_rollout_tc(const RolloutParams& p, int HP, cudaStream_t st);
// split-fp16 / tcgen05 variant, fp32-grade (g2048_rollout_x3.cu)
int launch_rollout_x3(const RolloutParams& p, int HP, cudaStream_t st);

}  // namespace g2048
