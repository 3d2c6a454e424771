// g2048_rollout_urm.cu -- fused rollout for the GameURM policy (BASELINE config #5), tcgen05 path.
//
// Reference (file:line in RobotSail/2048-PPO): GameURM / GameURMBlock / GameURMAttention /
// GameConvSwiGLU / rms_norm  game.py:1223-1458, default GameURMConfig game.py:31-42 (hidden 64,
// 2 layers, 4 heads, expansion 2.67 -> inter 120, conv kernel 2, 4 loops).  Eval semantics (the
// truncated loops differ only in autograd).  Rollout loop / records as in g2048_rollout.cu.
//
// Mapping: a tile is 128 tokens = 8 envs x 16 cells; token m IS TMEM lane m IS thread m (4 warps), so
// RMSNorm, residual adds, SwiGLU and the stem are thread-local.  The four projections of a block run
// as fp16 x fp16 -> fp32 tcgen05.mma against weight images that stay resident in shared memory
// (156 KiB for 2 layers); thread 0 issues them.  Attention (16 tokens, head_dim 16) runs on CUDA
// cores: K and V of the tile pass through a 32 KiB fp16 buffer, q stays in registers.  The depthwise
// conv (k=2, pad 1, trimmed: out[t] = w0*x[t-1] + w1*x[t] + b) is a lane shuffle; the mean-pool +
// heads a 16-lane shuffle reduction.  The fp32 hidden state and the input embedding live in TMEM.
// Precision: GEMM operands and K/V are fp16 (11 mantissa bits: every tensor on this path is O(1) after the RMS norms;
// round 1 used bf16, 8 bits, and sat 0.10 / 0.08 off the fp32 model after the 8 block applications), everything else
// fp32; tests compare against a torch emulation of exactly this arithmetic and against the reference's fp32 outputs.
#include <cuda_fp16.h>
#include "g2048_urm.cuh"
#include "g2048_tc.cuh"

namespace g2048 {
namespace urm {

struct Smem {
    alignas(1024) uint8_t W[MAX_LAYERS * IMG_LAYER];
    alignas(1024) uint8_t A[128 * 128];              // A operand, K block 0
    alignas(1024) uint8_t KV[128 * 256];             // fp16 [token][K(64) | V(64)]; its first 16 KiB double as A block 1
    alignas(16) float stem_w[H * 3];
    alignas(16) float stem_g[H];
    alignas(16) float stem_b[H];
    alignas(16) float init_h[SEQ * H];
    alignas(16) float headw[5 * H + 8];
    alignas(16) float conv[MAX_LAYERS][3][128];
    uint64_t a_ready, mma_done, w_full;
    uint32_t tmem_base;
};

__global__ void pack_kernel(PackSrc s, int L, float* __restrict__ out) {
    const int64_t nf = f_total(L);
    uint8_t* img = reinterpret_cast<uint8_t*>(out + nf);
    const int64_t img_elems = int64_t(L) * (IMG_LAYER / 2);
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nf + img_elems; i += int64_t(gridDim.x) * blockDim.x) {
        if (i < nf) {
            float v = 0.f;
            if (i < F_STEM_G) v = s.stem_w[i];
            else if (i < F_STEM_B) v = s.stem_g[i - F_STEM_G];
            else if (i < F_INIT) v = s.stem_b[i - F_STEM_B];
            else if (i < F_HEADW) v = s.init_hidden[i - F_INIT];
            else if (i < F_HEADB) {
                const int j = int(i - F_HEADW) / H, n = int(i - F_HEADW) % H;
                v = j < 4 ? s.act_w[j * H + n] : s.val_w[n];
            } else if (i < F_CONV) {
                const int j = int(i - F_HEADB);
                v = j < 4 ? s.act_b[j] : (j == 4 ? s.val_b[0] : 0.f);
            } else if (i < F_CONV + L * F_CONV_STRIDE) {
                const int r = int(i - F_CONV), l = r / F_CONV_STRIDE, k = (r % F_CONV_STRIDE) / 128, c = r % 128;
                if (c < INTER) v = k == 0 ? s.conv_w[l][2 * c] : (k == 1 ? s.conv_w[l][2 * c + 1] : s.conv_b[l][c]);
            }
            out[i] = v;
        } else {
            // one fp16 element of a layer image; element index -> (matrix, row n, k)
            int64_t e = i - nf;
            const int l = int(e / (IMG_LAYER / 2));
            e %= IMG_LAYER / 2;
            uint8_t* base = img + int64_t(l) * IMG_LAYER;
            float v = 0.f;
            int rows, n, k;
            if (e < IMG_O / 2) {
                rows = QKV; n = int(e / 64); k = int(e % 64);
                v = s.qkv[l][n * H + k];
                base += IMG_QKV;
            } else if (e < IMG_GU / 2) {
                e -= IMG_O / 2;
                rows = H; n = int(e / 64); k = int(e % 64);
                v = s.o[l][n * H + k];
                base += IMG_O;
            } else if (e < IMG_D / 2) {
                e -= IMG_GU / 2;
                rows = GU; n = int(e / 64); k = int(e % 64);
                v = s.gu[l][n * H + k];
                base += IMG_GU;
            } else {
                e -= IMG_D / 2;
                rows = H; n = int(e % (H * 64)) / 64; k = int(e / (H * 64)) * 64 + int(e % 64);
                v = k < INTER ? s.down[l][n * INTER + k] : 0.f;
                base += IMG_D;
            }
            *reinterpret_cast<__half*>(base + tc::sw128_offset(rows, n, k)) = __float2half_rn(v);
        }
    }
}

// x * sigmoid(x) with the hardware ex2 / rcp approximations (~1e-6 relative): the surrounding GEMM operands are fp16 (~5e-4), and
// the IEEE divide + expf were a fifth of this kernel's instructions (ncu source view: 240 SiLUs per token and block)
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ void sync128() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// write 8 consecutive K-elements [k0, k0+8) of this thread's A row (k0 % 8 == 0) as fp16
__device__ __forceinline__ void store_a8(uint8_t* a_base, int row, int k0, const float* x) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __half2 pr = __floats2half2_rn(x[2 * q], x[2 * q + 1]);
        w[q] = *reinterpret_cast<const uint32_t*>(&pr);
    }
    const uint32_t blk = uint32_t(k0) >> 6, unit = ((uint32_t(k0) & 63u) >> 3) ^ uint32_t(row & 7);
    *reinterpret_cast<uint4*>(a_base + blk * 16384u + uint32_t(row >> 3) * 1024u + uint32_t(row & 7) * 128u + unit * 16u) =
        make_uint4(w[0], w[1], w[2], w[3]);
}

struct Mma {
    uint32_t a0;
    // D[128][N] = A[128][16*ksteps] * B[N][16*ksteps]^T
    __device__ __forceinline__ void issue(Smem& S, uint32_t tmem_base, uint32_t b_addr, int N, int ksteps, uint64_t st) {
        tc::mbar_wait(&S.a_ready, uint32_t(st) & 1u);
        tc::fence_after_sync();
        const uint32_t idesc = tc::make_idesc_f16(128, N);
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t blk = uint32_t(ks) >> 2, j = uint32_t(ks) & 3u;
            tc::mma_bf16_ss(tmem_base, tc::make_desc_sw128(a0 + blk * 16384u + j * 32u),
                            tc::make_desc_sw128(b_addr + blk * (uint32_t(N) * 128u) + j * 32u), idesc, ks > 0);
        }
        tc::mma_commit(&S.mma_done);
    }
};

constexpr uint32_t T_H = 256;   // TMEM columns: [0,240) accumulator, [256,320) hidden state, [320,384) input embedding
constexpr uint32_t T_E = 320;

__global__ void __launch_bounds__(THREADS, 1) rollout_urm_kernel(RolloutParams p, int loops) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // align by pointer arithmetic on the __shared__ array (not through an integer) so that the
    // compiler keeps the shared address space and emits LDS/STS instead of generic loads
    Smem& S = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (tc::smem_addr(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.layers;
    const int cell = tid & 15;                                   // token = cell of env (tid >> 4) of the tile
    const int64_t ntiles = (p.B + 7) / 8;
    const int64_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const float* pk = p.packed;
    const uint8_t* img = reinterpret_cast<const uint8_t*>(pk + f_total(L));

    if (warp == 0) tc::tmem_alloc(&S.tmem_base, 512);
    if (tid == 0) {
        tc::mbar_init(&S.a_ready, THREADS);
        tc::mbar_init(&S.mma_done, 1);
        tc::mbar_init(&S.w_full, 1);
        tc::mbar_fence_init();
        tc::mbar_expect_tx(&S.w_full, uint32_t(L * IMG_LAYER));
        for (int l = 0; l < L; ++l)
            for (int off = 0; off < IMG_LAYER; off += IMG_LAYER / 4)      // 4 bulk copies of 19 968 B per layer
                tc::bulk_g2s(S.W + l * IMG_LAYER + off, img + size_t(l) * IMG_LAYER + off, IMG_LAYER / 4, &S.w_full);
    }
    for (int i = tid; i < H * 3; i += THREADS) S.stem_w[i] = pk[F_STEM_W + i];
    for (int i = tid; i < H; i += THREADS) {
        S.stem_g[i] = pk[F_STEM_G + i];
        S.stem_b[i] = pk[F_STEM_B + i];
    }
    for (int i = tid; i < SEQ * H; i += THREADS) S.init_h[i] = pk[F_INIT + i];
    for (int i = tid; i < 5 * H + 8; i += THREADS) S.headw[i] = pk[F_HEADW + i];
    for (int i = tid; i < L * F_CONV_STRIDE; i += THREADS) (&S.conv[0][0][0])[i] = pk[F_CONV + i];
    for (uint32_t i = tid * 16; i < uint32_t(sizeof(S.A) + sizeof(S.KV)); i += THREADS * 16)
        *reinterpret_cast<uint4*>(S.A + i) = make_uint4(0, 0, 0, 0);    // A and KV are contiguous
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    tc::mbar_wait(&S.w_full, 0);
    const uint32_t tmem_base = S.tmem_base;
    const uint32_t tlane = tmem_base + (uint32_t(warp * 32) << 16);
    const bool issuer = tid == 0;
    Mma mma{tc::smem_addr(S.A)};
    const uint32_t w_addr = tc::smem_addr(S.W);
    const LutGlobal lut{p.lut};
    uint64_t st = 0;

    // one MMA stage: everybody signals "A written / D consumed", thread 0 issues, everybody waits for D
    auto stage = [&](uint32_t b_addr, int N, int ksteps) {
        tc::fence_async_smem();
        tc::fence_before_sync();
        tc::mbar_arrive(&S.a_ready);
        if (issuer) mma.issue(S, tmem_base, b_addr, N, ksteps, st);
        tc::mbar_wait(&S.mma_done, uint32_t(st) & 1u);
        tc::fence_after_sync();
        ++st;
    };

    for (int64_t tl = 0; tl < my_tiles; ++tl) {
        const int64_t env = (int64_t(blockIdx.x) + tl * gridDim.x) * 8 + (tid >> 4);
        const bool in_range = env < p.B;
        const bool owner = in_range && cell == 0;                 // the env's leader thread owns the board
        Board board = {0u, 0u};
        bool alive = false;
        if (owner) {
            board = make_board(p.boards[env]);
            alive = p.alive ? p.alive[env] != 0 : true;
        }
        for (int t = 0; t < p.T; ++t) {
            const uint64_t ctr = p.ctr0 + uint64_t(t);
            uint32_t lm = 0;
            if (owner) lm = begin_step(p, env, ctr, board, alive);
            // every token needs its cell's exponent: broadcast the board from the leader (lane & 16)
            const uint32_t blo = __shfl_sync(0xffffffffu, board.lo, lane & 16), bhi = __shfl_sync(0xffffffffu, board.hi, lane & 16);
            const float ex = float(((cell < 8 ? blo : bhi) >> (4 * (cell & 7))) & 15u);
            // ---- stem: Linear(3 -> 64, no bias) + LayerNorm + SiLU on [exp, row/3, col/3]   game.py:1376-1380
            float hreg[H];
            {
                const float fr = pos_feature(cell >> 2), fc = pos_feature(cell & 3);
                float sum = 0.f;
#pragma unroll
                for (int n = 0; n < H; ++n) {
                    hreg[n] = fmaf(S.stem_w[3 * n], ex, fmaf(S.stem_w[3 * n + 1], fr, S.stem_w[3 * n + 2] * fc));
                    sum += hreg[n];
                }
                const float mean = sum * (1.0f / H);
                float sq = 0.f;
#pragma unroll
                for (int n = 0; n < H; ++n) {
                    hreg[n] -= mean;
                    sq = fmaf(hreg[n], hreg[n], sq);
                }
                const float rstd = rsqrtf(sq * (1.0f / H) + 1e-5f);
#pragma unroll
                for (int n = 0; n < H; ++n) hreg[n] = silu(fmaf(hreg[n] * rstd, S.stem_g[n], S.stem_b[n]));
#pragma unroll
                for (int c = 0; c < H; c += 8) tc::tmem_st8(tlane + T_E + uint32_t(c), &hreg[c]);   // input embedding
#pragma unroll
                for (int n = 0; n < H; ++n) hreg[n] = S.init_h[cell * H + n];                       // game.py:1431
            }
            for (int loop = 0; loop < loops; ++loop) {
                // hidden += input_embeddings   game.py:1441,1447
#pragma unroll
                for (int c = 0; c < H; c += 8) {
                    float e[8];
                    tc::tmem_ld8(tlane + T_E + uint32_t(c), e);
#pragma unroll
                    for (int j = 0; j < 8; ++j) hreg[c + j] += e[j];
                }
                for (int l = 0; l < L; ++l) {
                    const uint32_t wl = w_addr + uint32_t(l) * IMG_LAYER;
                    // ---------------- attention   game.py:1296-1317
#pragma unroll
                    for (int c = 0; c < H; c += 8) {
                        store_a8(S.A, tid, c, &hreg[c]);
                        tc::tmem_st8(tlane + T_H + uint32_t(c), &hreg[c]);      // fp32 residual copy
                    }
                    tc::tmem_st_wait();
                    stage(wl + IMG_QKV, QKV, H / 16);
#pragma unroll
                    for (int c = 0; c < 2 * H; c += 8) {                        // K | V -> fp16 rows (q stays in TMEM)
                        float kv[8];
                        tc::tmem_ld8(tlane + uint32_t(H + c), kv);
                        uint32_t w[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const __half2 pr = __floats2half2_rn(kv[2 * r], kv[2 * r + 1]);
                            w[r] = *reinterpret_cast<const uint32_t*>(&pr);
                        }
                        *reinterpret_cast<uint4*>(S.KV + tid * 256 + c * 2) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    sync128();
                    {
                        const uint8_t* kv_env = S.KV + (tid & ~15) * 256;       // the 16 tokens of this env
#pragma unroll 1
                        for (int hd = 0; hd < NHEAD; ++hd) {
                            float q[HD];
                            tc::tmem_ld16p(tlane + uint32_t(hd * HD), q);       // this head's query, fp32
                            float sc[SEQ], mx = -INFINITY;
#pragma unroll
                            for (int j = 0; j < SEQ; ++j) {
                                const uint4* kp = reinterpret_cast<const uint4*>(kv_env + j * 256 + hd * 32);
                                const uint4 ka = kp[0], kb = kp[1];
                                const uint32_t kw[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
                                float s = 0.f;
#pragma unroll
                                for (int d = 0; d < 8; ++d) {
                                    const float2 kf = __half22float2(*reinterpret_cast<const __half2*>(&kw[d]));
                                    s = fmaf(q[2 * d], kf.x, s);
                                    s = fmaf(q[2 * d + 1], kf.y, s);
                                }
                                sc[j] = s * 0.25f;                              // 1/sqrt(head_dim)
                                mx = fmaxf(mx, sc[j]);
                            }
                            float den = 0.f;
#pragma unroll
                            for (int j = 0; j < SEQ; ++j) {
                                sc[j] = __expf(sc[j] - mx);
                                den += sc[j];
                            }
                            const float inv = __fdividef(1.0f, den);
                            float o[HD];
#pragma unroll
                            for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
                            for (int j = 0; j < SEQ; ++j) {
                                const uint4* vp = reinterpret_cast<const uint4*>(kv_env + j * 256 + 128 + hd * 32);
                                const uint4 va = vp[0], vb = vp[1];
                                const uint32_t vw[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
                                const float pj = sc[j] * inv;
#pragma unroll
                                for (int d = 0; d < 8; ++d) {
                                    const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(&vw[d]));
                                    o[2 * d] = fmaf(pj, vf.x, o[2 * d]);
                                    o[2 * d + 1] = fmaf(pj, vf.y, o[2 * d + 1]);
                                }
                            }
                            store_a8(S.A, tid, hd * HD, &o[0]);
                            store_a8(S.A, tid, hd * HD + 8, &o[8]);
                        }
                    }
                    stage(wl + IMG_O, H, H / 16);
                    // hidden = rms_norm(hidden + attn_out)   game.py:1345-1346
                    {
                        float sq = 0.f;
#pragma unroll
                        for (int c = 0; c < H; c += 8) {
                            float d[8], r[8];
                            tc::tmem_ld8(tlane + uint32_t(c), d);
                            tc::tmem_ld8(tlane + T_H + uint32_t(c), r);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                hreg[c + j] = r[j] + d[j];
                                sq = fmaf(hreg[c + j], hreg[c + j], sq);
                            }
                        }
                        const float rs = rsqrtf(sq * (1.0f / H) + 1e-5f);
#pragma unroll
                        for (int n = 0; n < H; ++n) hreg[n] *= rs;
                    }
                    // ---------------- ConvSwiGLU   game.py:1264-1276
#pragma unroll
                    for (int c = 0; c < H; c += 8) {
                        store_a8(S.A, tid, c, &hreg[c]);
                        tc::tmem_st8(tlane + T_H + uint32_t(c), &hreg[c]);
                    }
                    tc::tmem_st_wait();
                    stage(wl + IMG_GU, GU, H / 16);
#pragma unroll 1
                    for (int c = 0; c < 128; c += 8) {
                        float x[8];
                        if (c < INTER) {
                            float g[8], u[8];
                            tc::tmem_ld8(tlane + uint32_t(c), g);
                            tc::tmem_ld8(tlane + uint32_t(INTER + c), u);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float cur = silu(g[j]) * u[j];
                                float prev = __shfl_up_sync(0xffffffffu, cur, 1);       // token t-1 of the same env
                                prev = cell == 0 ? 0.f : prev;
                                const float y = fmaf(S.conv[l][0][c + j], prev, fmaf(S.conv[l][1][c + j], cur, S.conv[l][2][c + j]));
                                x[j] = silu(y);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) x[j] = 0.f;                    // K padding 120..127
                        }
                        store_a8(S.A, tid, c, x);                                      // block 1 aliases the (dead) KV buffer
                    }
                    stage(wl + IMG_D, H, 128 / 16);
                    // hidden = rms_norm(hidden + mlp_out)   game.py:1349-1350
                    {
                        float sq = 0.f;
#pragma unroll
                        for (int c = 0; c < H; c += 8) {
                            float d[8], r[8];
                            tc::tmem_ld8(tlane + uint32_t(c), d);
                            tc::tmem_ld8(tlane + T_H + uint32_t(c), r);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                hreg[c + j] = r[j] + d[j];
                                sq = fmaf(hreg[c + j], hreg[c + j], sq);
                            }
                        }
                        const float rs = rsqrtf(sq * (1.0f / H) + 1e-5f);
#pragma unroll
                        for (int n = 0; n < H; ++n) hreg[n] *= rs;
                    }
                }
            }
            // ---- mean-pool over the 16 tokens + heads   game.py:1451-1456
            float o[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                float s = 0.f;
#pragma unroll
                for (int n = 0; n < H; ++n) s = fmaf(S.headw[j * H + n], hreg[n], s);
#pragma unroll
                for (int m = 8; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
                o[j] = s * (1.0f / SEQ) + S.headw[5 * H + j];
            }
            if (owner) policy_env_step(p, lut, t, env, ctr, lm, o, board, alive);
        }
        if (owner) {
            p.boards[env] = pack_board(board);
            if (p.alive) p.alive[env] = alive ? 1 : 0;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace urm
}  // namespace g2048

using namespace g2048;

extern "C" {

int64_t g2048_urm_packed_floats(int32_t hidden, int32_t layers, int32_t heads, int32_t inter) {
    if (hidden != urm::H || heads != urm::NHEAD || inter != urm::INTER || layers < 1 || layers > urm::MAX_LAYERS) return -1;
    return urm::total_floats_all(layers);
}

int g2048_urm_pack(int32_t layers, const float* stem_w, const float* stem_ln_w, const float* stem_ln_b,
                   const float* init_hidden, const float* const* qkv_w, const float* const* o_w,
                   const float* const* gate_up_w, const float* const* dwconv_w, const float* const* dwconv_b,
                   const float* const* down_w, const float* action_w, const float* action_b, const float* value_w,
                   const float* value_b, float* packed, void* stream) {
    if (layers < 1 || layers > urm::MAX_LAYERS) return fail(G2048_ESHAPE, "g2048_urm_pack: layers=%d unsupported (1..%d)", layers, urm::MAX_LAYERS);
    G2048_REQUIRE(stem_w && stem_ln_w && stem_ln_b && init_hidden && qkv_w && o_w && gate_up_w && dwconv_w && dwconv_b &&
                      down_w && action_w && action_b && value_w && value_b && packed,
                  "g2048_urm_pack: NULL pointer argument");
    urm::PackSrc s{};
    s.stem_w = stem_w; s.stem_g = stem_ln_w; s.stem_b = stem_ln_b; s.init_hidden = init_hidden;
    s.act_w = action_w; s.act_b = action_b; s.val_w = value_w; s.val_b = value_b;
    for (int l = 0; l < layers; ++l) {
        G2048_REQUIRE(qkv_w[l] && o_w[l] && gate_up_w[l] && dwconv_w[l] && dwconv_b[l] && down_w[l], "g2048_urm_pack: NULL layer pointer");
        s.qkv[l] = qkv_w[l]; s.o[l] = o_w[l]; s.gu[l] = gate_up_w[l];
        s.conv_w[l] = dwconv_w[l]; s.conv_b[l] = dwconv_b[l]; s.down[l] = down_w[l];
    }
    urm::pack_kernel<<<256, 256, 0, cudaStream_t(stream)>>>(s, layers, packed);
    G2048_CHECK_LAUNCH("urm::pack_kernel");
    return urm::launch_urm_x3_pack(s, layers, packed, cudaStream_t(stream));
}

int g2048_rollout_urm(const G2048Rollout* r, int32_t loops, void* stream) {
    G2048_REQUIRE(r != nullptr, "g2048_rollout_urm: NULL params");
    G2048_REQUIRE(r->B >= 0 && r->T >= 0 && loops >= 0, "g2048_rollout_urm: negative shape");
    if (r->B == 0 || r->T == 0) return G2048_OK;
    if (r->hidden != urm::H || r->layers < 1 || r->layers > urm::MAX_LAYERS)
        return fail(G2048_ESHAPE, "g2048_rollout_urm: only hidden=64, 1..%d layers (4 heads, inter 120) are built", urm::MAX_LAYERS);
    G2048_REQUIRE(r->packed_weights && r->lut && r->boards && r->rec_boards && r->rec_actions && r->rec_legal &&
                      r->rec_logp && r->rec_value && r->rec_points && r->rec_shaping && r->rec_flags,
                  "g2048_rollout_urm: NULL pointer argument");
    G2048_REQUIRE((reinterpret_cast<uintptr_t>(r->packed_weights) & 255u) == 0, "g2048_rollout_urm: packed_weights must be 256-byte aligned");
    RolloutParams p{};
    p.B = r->B; p.T = r->T; p.hidden = r->hidden; p.layers = r->layers; p.auto_reset = r->auto_reset;
    p.seed = r->seed; p.env0 = r->env0; p.ctr0 = r->ctr0;
    p.packed = r->packed_weights; p.lut = static_cast<const uint32_t*>(r->lut);
    p.boards = r->boards; p.alive = r->alive; p.forced_actions = r->forced_actions;
    p.rec_boards = r->rec_boards; p.rec_actions = r->rec_actions; p.rec_legal = r->rec_legal; p.rec_logp = r->rec_logp;
    p.rec_value = r->rec_value; p.rec_points = r->rec_points; p.rec_shaping = r->rec_shaping; p.rec_flags = r->rec_flags;
    p.rec_entropy = r->rec_entropy;
    // fp32 grade (split-fp16 operands) unless the caller asks for the single-fp16-operand variant by name
    if (r->tensor_cores != G2048_ROLLOUT_BF16) return urm::launch_rollout_urm_x3(p, loops, cudaStream_t(stream));
    const int smem = int(sizeof(urm::Smem)) + 1024;
    G2048_CHECK_CUDA(ensure_smem(urm::rollout_urm_kernel, smem));
    const int64_t ntiles = (p.B + 7) / 8;
    const int grid = int(ntiles < num_sms() ? ntiles : num_sms());
    urm::rollout_urm_kernel<<<grid, urm::THREADS, smem, cudaStream_t(stream)>>>(p, loops);
    G2048_CHECK_LAUNCH("rollout_urm_kernel");
    return G2048_OK;
}

}  // extern "C"
