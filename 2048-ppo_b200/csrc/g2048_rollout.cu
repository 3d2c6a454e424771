// g2048_rollout.cu -- persistent fused actor-critic rollout for the GameMLP policy.
//
// One launch plays T steps of B environments: board -> model input -> stem / residual blocks /
// action + value heads -> masked log-softmax -> categorical sample (Philox) -> env step (move,
// merge points, shaping, Philox spawn, done / auto-reset) -> one record per [t, env].
// Nothing but the records (and the final boards) touches HBM; weights stream L2 -> shared
// memory through bulk async copies (cp.async.bulk + mbarrier ring) issued by a producer warp.
//
// Reference (file:line in RobotSail/2048-PPO):
//   play_game_for_episode            train.py:213-345   (the loop this kernel batches)
//   to_model_format                  game.py:92-101     (the constant row/col inputs are folded
//                                                        into a per-model stem bias)
//   GameMLP.forward / ResidualBlock  game.py:1033-1046, 1145-1220 (eval mode: dropout off)
//   masked softmax + sampling        train.py:266-291, 326
//   StepData fields                  train.py:299-326
//
// Work decomposition: a CTA owns tiles of 128 envs and runs each tile through all T steps
// (boards stay in registers).  The GEMMs are fp32 FFMA with an 8 (envs) x TN (units) register
// tile per thread: 256 consumer threads = 16 env-groups x 16 unit-groups; activations live in
// shared memory k-major (X[k][env]) so both operands are conflict-free 128-bit loads.
// fp32 accumulation keeps the recorded log-probs / values within ~1e-6 of the torch model.
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "g2048_rollout.cuh"

namespace g2048 {

// ------------------------------------------------------------------ weight packing
struct PackSrc {
    const float *stem_w, *stem_g, *stem_b;
    const float* blk_w[8];
    const float* blk_g[8];
    const float* blk_b[8];
    const float *act_w, *act_b, *val_w, *val_b;
};

__global__ void pack_mlp_kernel(PackSrc s, int h, int HP, int L, float* __restrict__ out) {
    const int64_t total = pk_total(HP, L);
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        float v = 0.f;
        if (i < pk_stem_b0(HP)) {                       // Wt_stem[k][n] = stem.0.weight[n][3k]
            int k = int(i / HP), n = int(i % HP);
            if (n < h) v = s.stem_w[n * 48 + 3 * k];
        } else if (i < pk_stem_g(HP)) {                 // b0[n] = sum_cells W[n][3c+1]*row/3 + W[n][3c+2]*col/3
            int n = int(i - pk_stem_b0(HP));
            if (n < h) {
                float acc = 0.f;
                for (int c = 0; c < 16; ++c) {
                    acc = fmaf(s.stem_w[n * 48 + 3 * c + 1], pos_feature(c >> 2), acc);
                    acc = fmaf(s.stem_w[n * 48 + 3 * c + 2], pos_feature(c & 3), acc);
                }
                v = acc;
            }
        } else if (i < pk_stem_beta(HP)) {
            int n = int(i - pk_stem_g(HP));
            if (n < h) v = s.stem_g[n];
        } else if (i < pk_layer(HP, 0)) {
            int n = int(i - pk_stem_beta(HP));
            if (n < h) v = s.stem_b[n];
        } else if (i < pk_heads(HP, L)) {
            int64_t r = i - pk_layer(HP, 0);
            int l = int(r / (int64_t(HP) * (HP + 2)));
            r -= int64_t(l) * HP * (HP + 2);
            if (r < int64_t(HP) * HP) {                 // Wt[k][n] = W[n][k]
                int k = int(r / HP), n = int(r % HP);
                if (k < h && n < h) v = s.blk_w[l][n * h + k];
            } else if (r < int64_t(HP) * (HP + 1)) {
                int n = int(r - int64_t(HP) * HP);
                if (n < h) v = s.blk_g[l][n];
            } else {
                int n = int(r - int64_t(HP) * (HP + 1));
                if (n < h) v = s.blk_b[l][n];
            }
        } else {
            int64_t r = i - pk_heads(HP, L);
            if (r < 5 * HP) {
                int j = int(r / HP), n = int(r % HP);
                if (n < h) v = j < 4 ? s.act_w[j * h + n] : s.val_w[n];
            } else {
                int j = int(r - 5 * HP);
                v = j < 4 ? s.act_b[j] : (j == 4 ? s.val_b[0] : 0.f);
            }
        }
        out[i] = v;
    }
}

// bf16 operand images for the tensor-core kernel (layout: g2048_tc.cuh sw128_offset)
__global__ void pack_mlp_images_kernel(PackSrc s, int h, int HP, int L, uint8_t* __restrict__ img) {
    const int KB = kblocks_of(HP);
    const int64_t stem_elems = int64_t(HP) * 64, layer_elems = int64_t(KB) * HP * 64;
    const int64_t total = stem_elems + int64_t(L) * layer_elems;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        float v = 0.f;
        uint8_t* base;
        int n, k;
        if (i < stem_elems) {
            n = int(i / 64);
            k = int(i % 64);
            if (n < h && k < 16) v = s.stem_w[n * 48 + 3 * k];
            base = img;
        } else {
            const int64_t r = i - stem_elems;
            const int l = int(r / layer_elems);
            const int64_t e = r % layer_elems;
            n = int(e / (int64_t(KB) * 64));
            k = int(e % (int64_t(KB) * 64));
            if (n < h && k < h) v = s.blk_w[l][n * h + k];
            base = img + img_stem_bytes(HP) + int64_t(l) * img_layer_bytes(HP);
        }
        const uint32_t blk = uint32_t(k) >> 6, kk = uint32_t(k) & 63u;
        const uint32_t off = blk * uint32_t(HP) * 128u + uint32_t(n >> 3) * 1024u + uint32_t(n & 7) * 128u +
                             (((kk >> 3) ^ uint32_t(n & 7)) << 4) + (kk & 7u) * 2u;
        *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16(v);
    }
}

// split-fp16 operand images for the fp32-grade tensor-core kernel (layout: g2048_rollout.cuh x3_*, g2048_tc.cuh sw32_offset)
__global__ void pack_mlp_x3_kernel(PackSrc s, int h, int HP, int L, uint8_t* __restrict__ img) {
    const int NB = HP / 16;
    const int64_t total = x3_blocks(HP, L) * HP * 16;
    const size_t part = size_t(HP) * 32;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int b = int(i / (HP * 16)), rem = int(i % (HP * 16)), n = rem / 16, kk = rem % 16;
        float v = 0.f;
        if (b == 0) {
            if (n < h) v = s.stem_w[n * 48 + 3 * kk];
        } else {
            const int l = (b - 1) / NB, k = ((b - 1) % NB) * 16 + kk;
            if (n < h && k < h) v = s.blk_w[l][size_t(n) * h + k];
        }
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        uint8_t* blk = img + size_t(b) * 2 * part;
        const uint32_t off = uint32_t(n) * 32u + uint32_t((((kk >> 3) ^ (n >> 2)) & 1) << 4) + uint32_t(kk & 7) * 2u;
        *reinterpret_cast<__half*>(blk + off) = hi;
        *reinterpret_cast<__half*>(blk + part + off) = lo;
    }
}

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t sm_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mb_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sm_u32(bar)) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(sm_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     sm_u32(dst)),
                 "l"(src), "r"(bytes), "r"(sm_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(RO_CONSUMERS) : "memory"); }

// ------------------------------------------------------------------ kernel
template <int HP>
struct Tile {
    static constexpr int NJ = HP / 64;            // float4 column groups per thread
    static constexpr int NS = (HP - 64 * NJ) / 16; // scalar columns per thread
    static constexpr int TN = 4 * NJ + NS;
    static constexpr int CHUNK_FLOATS = RO_KC * HP;
    static constexpr int CHUNKS_PER_LAYER = HP / RO_KC;
    // unit index of this thread's j-th column
    __device__ static __forceinline__ int col(int ng, int j) {
        return j < 4 * NJ ? 64 * (j >> 2) + 4 * ng + (j & 3) : 64 * NJ + 16 * (j - 4 * NJ) + ng;
    }
};

template <int HP>
struct RolloutSmem {
    float X[HP][RO_TILE];                          // activations, k-major
    float W[RO_STAGES][RO_KC * HP];                // weight chunk ring
    float red[16][RO_TILE];                        // LayerNorm partials
    float stat[2][RO_TILE];                        // mean, rstd
    float headw[5 * HP + 8];
    float headp[RO_TILE][8];
    uint64_t full[RO_STAGES], empty[RO_STAGES];
};

template <int HP>
__device__ __forceinline__ void gemm_chunks(RolloutSmem<HP>& S, int nchunks, int mg, int ng, uint32_t& cons_it,
                                            float (&acc)[8][Tile<HP>::TN], int k0) {
    using TL = Tile<HP>;
    for (int c = 0; c < nchunks; ++c, ++cons_it) {
        const int stage = cons_it % RO_STAGES;
        mb_wait(&S.full[stage], (cons_it / RO_STAGES) & 1u);
        const float* Wc = S.W[stage];
#pragma unroll 4
        for (int kk = 0; kk < RO_KC; ++kk) {
            const int k = k0 + c * RO_KC + kk;
            const float4 xa = *reinterpret_cast<const float4*>(&S.X[k][4 * mg]);
            const float4 xb = *reinterpret_cast<const float4*>(&S.X[k][64 + 4 * mg]);
            const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            float w[TL::TN];
#pragma unroll
            for (int j = 0; j < TL::NJ; ++j) {
                const float4 wv = *reinterpret_cast<const float4*>(&Wc[kk * HP + 64 * j + 4 * ng]);
                w[4 * j + 0] = wv.x;
                w[4 * j + 1] = wv.y;
                w[4 * j + 2] = wv.z;
                w[4 * j + 3] = wv.w;
            }
#pragma unroll
            for (int j = 0; j < TL::NS; ++j) w[4 * TL::NJ + j] = Wc[kk * HP + 64 * TL::NJ + 16 * j + ng];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TL::TN; ++j) acc[i][j] = fmaf(x[i], w[j], acc[i][j]);
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mb_arrive(&S.empty[stage]);
    }
}

// LayerNorm (eps 1e-5, biased variance, two-pass) + ReLU (+ residual) of the tile's pre-activations,
// written back into X.  game.py:1038-1046, 1069-1073.
template <int HP, bool RESIDUAL>
__device__ __forceinline__ void layernorm_relu_store(RolloutSmem<HP>& S, float (&acc)[8][Tile<HP>::TN], int mg, int ng,
                                                     int h, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta) {
    using TL = Tile<HP>;
    const int tid = threadIdx.x;
    const float inv_h = 1.0f / float(h);
    int m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = (i < 4 ? 4 * mg + i : 64 + 4 * mg + (i - 4));
    // mean
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < TL::TN; ++j) s += (TL::col(ng, j) < h) ? acc[i][j] : 0.f;
        S.red[ng][m[i]] = s;
    }
    consumer_sync();   // also: every thread is past the GEMM, X may be overwritten below
    if (tid < RO_TILE) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) s += S.red[g][tid];
        S.stat[0][tid] = s * inv_h;
    }
    consumer_sync();
    // variance
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float mu = S.stat[0][m[i]];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < TL::TN; ++j) {
            acc[i][j] -= mu;
            s += (TL::col(ng, j) < h) ? acc[i][j] * acc[i][j] : 0.f;
        }
        S.red[ng][m[i]] = s;
    }
    consumer_sync();
    if (tid < RO_TILE) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) s += S.red[g][tid];
        S.stat[1][tid] = 1.0f / sqrtf(s * inv_h + 1e-5f);
    }
    consumer_sync();
#pragma unroll
    for (int j = 0; j < TL::TN; ++j) {
        const int n = TL::col(ng, j);
        const bool real = n < h;
        const float g = real ? __ldg(gamma + n) : 0.f, b = real ? __ldg(beta + n) : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float y = fmaxf(fmaf(acc[i][j] * S.stat[1][m[i]], g, b), 0.f);
            if (RESIDUAL) y += S.X[n][m[i]];
            S.X[n][m[i]] = real ? y : 0.f;
        }
    }
    consumer_sync();
}

template <int HP>
__global__ void __launch_bounds__(RO_THREADS, 1) rollout_mlp_kernel(RolloutParams p) {
    using TL = Tile<HP>;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    RolloutSmem<HP>& S = *reinterpret_cast<RolloutSmem<HP>*>(smem_raw);
    const int tid = threadIdx.x;
    const int L = p.layers, h = p.hidden;
    const int64_t ntiles = (p.B + RO_TILE - 1) / RO_TILE;
    const int64_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int chunks_per_step = 1 + L * TL::CHUNKS_PER_LAYER;

    if (tid == 0) {
        for (int s = 0; s < RO_STAGES; ++s) {
            mb_init(&S.full[s], 1);
            mb_init(&S.empty[s], RO_CONSUMERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 5 * HP + 8; i += RO_THREADS) S.headw[i] = p.packed[pk_heads(HP, L) + i];
    __syncthreads();

    if (tid >= RO_CONSUMERS) {
        // ---------------- producer warp: stream the weight chunks of every (tile, step) in order
        if (tid == RO_CONSUMERS) {
            const uint64_t total = uint64_t(my_tiles) * uint64_t(p.T) * uint64_t(chunks_per_step);
            uint32_t c_in_step = 0;
            for (uint64_t it = 0; it < total; ++it) {
                const int stage = int(it % RO_STAGES);
                if (it >= RO_STAGES) mb_wait(&S.empty[stage], uint32_t((it / RO_STAGES) - 1) & 1u);
                const float* src;
                if (c_in_step == 0) src = p.packed + pk_stem_w(HP);
                else {
                    const int l = (c_in_step - 1) / TL::CHUNKS_PER_LAYER, c = (c_in_step - 1) % TL::CHUNKS_PER_LAYER;
                    src = p.packed + pk_layer(HP, l) + int64_t(c) * TL::CHUNK_FLOATS;
                }
                mb_expect_tx(&S.full[stage], TL::CHUNK_FLOATS * 4);
                bulk_g2s(S.W[stage], src, TL::CHUNK_FLOATS * 4, &S.full[stage]);
                if (++c_in_step == uint32_t(chunks_per_step)) c_in_step = 0;
            }
        }
        return;
    }

    // ---------------- consumers
    const int mg = tid & 15, ng = tid >> 4;
    const LutGlobal lut{p.lut};
    uint32_t cons_it = 0;
    for (int64_t tl = 0; tl < my_tiles; ++tl) {
        const int64_t tile = blockIdx.x + tl * gridDim.x;
        const int64_t env = tile * RO_TILE + tid;               // meaningful for tid < RO_TILE
        const bool owner = tid < RO_TILE && env < p.B;
        Board board = {0u, 0u};
        bool alive = false;
        if (owner) {
            board = make_board(p.boards[env]);
            alive = p.alive ? p.alive[env] != 0 : true;
        }
        for (int t = 0; t < p.T; ++t) {
            const uint64_t ctr = p.ctr0 + uint64_t(t);
            uint32_t lm = 0;
            if (tid < RO_TILE) {
                if (owner) lm = begin_step(p, env, ctr, board, alive);
                // model input: the 16 exponents (row/col features are folded into the stem bias)
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint32_t w = k < 8 ? board.lo : board.hi;
                    S.X[k][tid] = float((w >> (4 * (k & 7))) & 15u);
                }
            }
            consumer_sync();

            float acc[8][TL::TN];
            // ---- stem: Linear(48->h, no bias) + LayerNorm + ReLU
#pragma unroll
            for (int j = 0; j < TL::TN; ++j) {
                const float b0 = __ldg(p.packed + pk_stem_b0(HP) + TL::col(ng, j));
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i][j] = b0;
            }
            gemm_chunks<HP>(S, 1, mg, ng, cons_it, acc, 0);
            layernorm_relu_store<HP, false>(S, acc, mg, ng, h, p.packed + pk_stem_g(HP), p.packed + pk_stem_beta(HP));
            // ---- residual blocks: x + ReLU(LayerNorm(Linear(h->h, no bias)(x)))   (dropout off: eval)
            for (int l = 0; l < L; ++l) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < TL::TN; ++j) acc[i][j] = 0.f;
                gemm_chunks<HP>(S, TL::CHUNKS_PER_LAYER, mg, ng, cons_it, acc, 0);
                const float* lw = p.packed + pk_layer(HP, l) + int64_t(HP) * HP;
                layernorm_relu_store<HP, true>(S, acc, mg, ng, h, lw, lw + HP);
            }
            // ---- heads: 4 action logits + value, each thread half of the units of one env
            {
                const int m = tid & (RO_TILE - 1), half = tid >> 7;
                float o[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
                const int n0 = half * (HP / 2);
#pragma unroll 4
                for (int n = n0; n < n0 + HP / 2; ++n) {
                    const float x = S.X[n][m];
#pragma unroll
                    for (int j = 0; j < 5; ++j) o[j] = fmaf(S.headw[j * HP + n], x, o[j]);
                }
                if (half == 1) {
#pragma unroll
                    for (int j = 0; j < 5; ++j) S.headp[m][j] = o[j];
                }
                consumer_sync();
                if (half == 0) {
#pragma unroll
                    for (int j = 0; j < 5; ++j) o[j] += S.headp[m][j] + S.headw[5 * HP + j];
                }
                // ---- policy + env step (one thread per env)
                if (owner) policy_env_step(p, lut, t, env, ctr, lm, o, board, alive);
            }
            // X is rewritten by the owners at the top of the next step; all head reads of X are
            // ordered before it by the consumer_sync() inside this step's head phase for half 0/1
            // and by the sync after the model-input write for everyone else.
            consumer_sync();
        }
        if (owner) {
            p.boards[env] = pack_board(board);
            if (p.alive) p.alive[env] = alive ? 1 : 0;
        }
    }
}

template <int HP>
int launch_rollout(const RolloutParams& p, cudaStream_t st) {
    const int smem = int(sizeof(RolloutSmem<HP>));
    auto kern = rollout_mlp_kernel<HP>;
    G2048_CHECK_CUDA(ensure_smem(kern, smem));
    const int64_t ntiles = (p.B + RO_TILE - 1) / RO_TILE;
    const int grid = int(ntiles < num_sms() ? ntiles : num_sms());
    kern<<<grid, RO_THREADS, smem, st>>>(p);
    G2048_CHECK_LAUNCH("rollout_mlp_kernel");
    return G2048_OK;
}

}  // namespace g2048

using namespace g2048;

extern "C" {

int64_t g2048_mlp_packed_floats(int32_t hidden, int32_t layers) {
    const int HP = padded_hidden(hidden);
    if (HP < 0 || layers < 0 || layers > 8) return -1;
    return pk_total_all(HP, layers);
}

int g2048_mlp_pack(int32_t hidden, int32_t layers, const float* stem_w, const float* stem_ln_w, const float* stem_ln_b,
                   const float* const* block_w, const float* const* block_ln_w, const float* const* block_ln_b,
                   const float* action_w, const float* action_b, const float* value_w, const float* value_b,
                   float* packed, void* stream) {
    const int HP = padded_hidden(hidden);
    if (HP < 0 || hidden < 1) return fail(G2048_ESHAPE, "g2048_mlp_pack: hidden=%d unsupported (1..208)", hidden);
    if (layers < 0 || layers > 8) return fail(G2048_ESHAPE, "g2048_mlp_pack: layers=%d unsupported (0..8)", layers);
    G2048_REQUIRE(stem_w && stem_ln_w && stem_ln_b && action_w && action_b && value_w && value_b && packed,
                  "g2048_mlp_pack: NULL pointer argument");
    G2048_REQUIRE(layers == 0 || (block_w && block_ln_w && block_ln_b), "g2048_mlp_pack: NULL block pointer array");
    PackSrc s{};
    s.stem_w = stem_w;
    s.stem_g = stem_ln_w;
    s.stem_b = stem_ln_b;
    for (int l = 0; l < layers; ++l) {
        G2048_REQUIRE(block_w[l] && block_ln_w[l] && block_ln_b[l], "g2048_mlp_pack: NULL block pointer");
        s.blk_w[l] = block_w[l];
        s.blk_g[l] = block_ln_w[l];
        s.blk_b[l] = block_ln_b[l];
    }
    s.act_w = action_w;
    s.act_b = action_b;
    s.val_w = value_w;
    s.val_b = value_b;
    pack_mlp_kernel<<<256, 256, 0, cudaStream_t(stream)>>>(s, hidden, HP, layers, packed);
    G2048_CHECK_LAUNCH("pack_mlp_kernel");
    pack_mlp_images_kernel<<<256, 256, 0, cudaStream_t(stream)>>>(
        s, hidden, HP, layers, reinterpret_cast<uint8_t*>(packed + pk_img_base(HP, layers)));
    G2048_CHECK_LAUNCH("pack_mlp_images_kernel");
    pack_mlp_x3_kernel<<<256, 256, 0, cudaStream_t(stream)>>>(s, hidden, HP, layers, reinterpret_cast<uint8_t*>(packed + pk_x3_base(HP, layers)));
    G2048_CHECK_LAUNCH("pack_mlp_x3_kernel");
    return G2048_OK;
}

int g2048_rollout_mlp(const G2048Rollout* r, void* stream) {
    G2048_REQUIRE(r != nullptr, "g2048_rollout_mlp: NULL params");
    G2048_REQUIRE(r->B >= 0 && r->T >= 0, "g2048_rollout_mlp: negative shape");
    if (r->B == 0 || r->T == 0) return G2048_OK;
    const int HP = padded_hidden(r->hidden);
    if (HP < 0 || r->hidden < 1) return fail(G2048_ESHAPE, "g2048_rollout_mlp: hidden=%d unsupported (1..208)", r->hidden);
    if (r->layers < 0 || r->layers > 8) return fail(G2048_ESHAPE, "g2048_rollout_mlp: layers=%d unsupported (0..8)", r->layers);
    G2048_REQUIRE(r->packed_weights && r->lut && r->boards && r->rec_boards && r->rec_actions && r->rec_legal &&
                      r->rec_logp && r->rec_value && r->rec_points && r->rec_shaping && r->rec_flags,
                  "g2048_rollout_mlp: NULL pointer argument");
    G2048_REQUIRE((reinterpret_cast<uintptr_t>(r->packed_weights) & 15u) == 0, "g2048_rollout_mlp: packed_weights must be 16-byte aligned");
    RolloutParams p{};
    p.B = r->B;
    p.T = r->T;
    p.hidden = r->hidden;
    p.layers = r->layers;
    p.auto_reset = r->auto_reset;
    p.seed = r->seed;
    p.env0 = r->env0;
    p.ctr0 = r->ctr0;
    p.packed = r->packed_weights;
    p.lut = static_cast<const uint32_t*>(r->lut);
    p.boards = r->boards;
    p.alive = r->alive;
    p.forced_actions = r->forced_actions;
    p.rec_boards = r->rec_boards;
    p.rec_actions = r->rec_actions;
    p.rec_legal = r->rec_legal;
    p.rec_logp = r->rec_logp;
    p.rec_value = r->rec_value;
    p.rec_points = r->rec_points;
    p.rec_shaping = r->rec_shaping;
    p.rec_flags = r->rec_flags;
    p.rec_entropy = r->rec_entropy;
    p.sched = static_cast<int32_t*>(r->sched_workspace);
    p.segs = 1;
    cudaStream_t st = cudaStream_t(stream);
    if (r->tensor_cores) {
        G2048_REQUIRE((reinterpret_cast<uintptr_t>(r->packed_weights) & 255u) == 0,
                      "g2048_rollout_mlp: packed_weights must be 256-byte aligned for the tensor-core kernels");
        return r->tensor_cores == G2048_ROLLOUT_BF16 ? launch_rollout_tc(p, HP, st) : launch_rollout_x3(p, HP, st);
    }
    switch (HP) {
        case 64: return launch_rollout<64>(p, st);
        case 128: return launch_rollout<128>(p, st);
        case 192: return launch_rollout<192>(p, st);
        case 208: return launch_rollout<208>(p, st);
    }
    return fail(G2048_ESHAPE, "g2048_rollout_mlp: no kernel for padded hidden %d", HP);
}

}  // extern "C"
