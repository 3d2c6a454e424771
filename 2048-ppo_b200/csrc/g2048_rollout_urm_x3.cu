// g2048_rollout_urm_x3.cu -- fused rollout for the GameURM policy (BASELINE config #5) at fp32 grade, tcgen05 path.
//
// Reference (file:line in RobotSail/2048-PPO): GameURM / GameURMBlock / GameURMAttention / GameConvSwiGLU / rms_norm
// game.py:1223-1458, default GameURMConfig game.py:31-42 (hidden 64, 2 layers, 4 heads, expansion 2.67 -> inter 120, conv
// kernel 2, 4 loops), eval semantics; rollout loop / records as in g2048_rollout.cu (train.py:213-345).
//
// Precision.  The reference's forward is fp32.  Every projection operand here is TWO fp16 terms, x = hi + lo (22 mantissa
// bits), and every k-step is three tcgen05.mma products lo*hi + hi*lo + hi*hi accumulated in fp32 in tensor memory (as in
// g2048_rollout_x3.cu); K and V stay fp32 in shared memory and the attention, norms, SiLUs and the depthwise conv are fp32 on
// the CUDA cores.  (g2048_rollout_urm.cu, single fp16 operands and fp16 K / V, is 1e-2 off the reference after the 8 block
// applications: a labelled variant.)
//
// Mapping.  A tile is 128 tokens = 8 envs x 16 cells; token m of a tile IS TMEM lane m IS one thread (4 warps), so RMS norm,
// residual adds, SwiGLU and the conv (a lane shuffle) are thread-local.  A CTA plays TWO tiles side by side (8 row warps) plus
// an MMA-issuer warp and a weight-producer warp: hi | lo weight images of two layers are 312 KiB and cannot stay resident, so
// the producer streams the k-blocks L2 -> SMEM through a ring of bulk copies in the one order every tile consumes them, and
// each block is used by both tiles before its slot is refilled (half the L2 traffic per tile); the issuer alternates between
// the tiles stage by stage, so that one tile's MMAs run under the other's epilogue.
// Per tile: 64 KiB of shared memory X that is, in turn, the hi | lo operand of the next projection (two 32 KiB regions) and the
// fp32 K | V buffer of the attention; 256 TMEM columns = accumulator [0,240), whose last 64 columns hold the fp32 residual
// while the attention runs (its accumulator is 192 wide; through the SwiGLU half the residual stays in registers).
// Stages of one block application (A operand -> accumulator columns):
//   QKV  hn (R0)            -> q | k | v  [0,192)      then K | V -> X (fp32), attention on CUDA cores, o -> R0
//   O    o  (R0)            -> [0,64)                  hidden = rms_norm(residual + .)          game.py:1345-1346
//   G1   hn (R0), N = 128   -> gate | up of channels 0..63 [0,128)      silu(gate) * up, conv, silu -> R1 (k-blocks 0..3 of D)
//   G2   hn (R0), N = 112   -> gate | up of channels 64..119 [128,240)  same -> R0 (k-blocks 4..7 of D)   game.py:1264-1276
//   D    x  (R1, then R0)   -> [0,64)                  hidden = rms_norm(residual + .)          game.py:1349-1350
// G1 and G2 are issued back to back with a completion barrier each, and the down projection starts on k-blocks 0..3 as soon as
// they are written, so that the second half of the SwiGLU work runs under MMAs of the same tile.
#include <cuda_fp16.h>
#include "g2048_urm.cuh"
#include "g2048_tc.cuh"

namespace g2048 {
namespace urm {
namespace x3 {

constexpr int TILES = 2;
constexpr int ROW_THREADS = TILES * 128;
constexpr int XTHREADS = ROW_THREADS + 64;        // + issuer warp, producer warp
constexpr int RING = 6;
constexpr uint32_t SLOT = 15360;                  // the largest unit: one gate | up k-block (G1 128 rows + G2 112 rows, x 64 B)
constexpr uint32_t G2_OFF = 8192;                 // G2's part of a gate | up unit
constexpr uint32_t REGION = 32768;                // one hi | lo operand region of X: [hi: 4 k-blocks x 4096 B | lo: same]
constexpr uint32_t LO_OFF = 16384;
constexpr int UNITS = 14;                         // ring units per layer: QKV 4, O 2, GU 4, D 4
constexpr uint32_t T_RES = 192;                   // TMEM column of the fp32 residual inside a tile's 256 columns
constexpr uint32_t T_G2 = 128;                    // TMEM column of G2's accumulator

struct Smem {
    alignas(1024) uint8_t X[TILES][65536];
    alignas(1024) uint8_t W[RING][SLOT];
    alignas(16) float conv[MAX_LAYERS][3][128];
    alignas(16) float headw[5 * H + 8];
    alignas(16) float zeros[128];
    uint64_t ready[TILES], rdy_d1[TILES], rdy_d2[TILES], done[TILES], done_g1[TILES], done_g2[TILES], w_full[RING], w_empty[RING];
    uint32_t tmem_base;
};
static_assert(sizeof(Smem) + 1024 <= 232448, "URM x3 kernel exceeds the 227 KB shared memory limit");

// size of ring unit u (0..13) of a layer's weight stream
__device__ __forceinline__ uint32_t unit_bytes(uint32_t u) { return u < 4u ? 12288u : (u >= 6u && u < 10u ? 15360u : 8192u); }

// ---------------------------------------------------------------------------------------------------------------- pack
__device__ __forceinline__ void put_split(uint8_t* kblock, int rows, int n, int k, float v) {
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const uint32_t off = tc::sw32_offset(n, k & 15);
    *reinterpret_cast<__half*>(kblock + off) = hi;
    *reinterpret_cast<__half*>(kblock + uint32_t(rows) * 32u + off) = lo;
}

__global__ void pack_x3_kernel(PackSrc s, int L, float* __restrict__ out) {
    float* emb = out + x3_base(L);
    uint8_t* img = reinterpret_cast<uint8_t*>(emb + X3_EMB_FLOATS);
    // elements per layer: QKV 192 x 64, O 64 x 64, GU 240 x 64, D 64 x 128
    constexpr int E_QKV = QKV * H, E_O = H * H, E_GU = GU * H, E_D = H * 128;
    constexpr int E_LAYER = E_QKV + E_O + E_GU + E_D;
    const int64_t total = int64_t(16 * SEQ) + int64_t(L) * E_LAYER;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        if (i < 16 * SEQ) {
            // stem of one (exponent, cell) pair: Linear(3 -> 64, no bias) + LayerNorm + SiLU   game.py:1376-1380, 92-101
            const int e = int(i) / SEQ, cell = int(i) % SEQ;
            const float ex = float(e), fr = pos_feature(cell >> 2), fc = pos_feature(cell & 3);
            float y[H], sum = 0.f;
            for (int n = 0; n < H; ++n) {
                y[n] = fmaf(s.stem_w[3 * n], ex, fmaf(s.stem_w[3 * n + 1], fr, s.stem_w[3 * n + 2] * fc));
                sum += y[n];
            }
            const float mean = sum / float(H);
            float sq = 0.f;
            for (int n = 0; n < H; ++n) {
                y[n] -= mean;
                sq = fmaf(y[n], y[n], sq);
            }
            const float rstd = 1.0f / sqrtf(sq / float(H) + 1e-5f);
            for (int n = 0; n < H; ++n) {
                const float v = fmaf(y[n] * rstd, s.stem_g[n], s.stem_b[n]);
                emb[i * H + n] = v / (1.0f + expf(-v));
            }
            continue;
        }
        int64_t e = i - 16 * SEQ;
        const int l = int(e / E_LAYER);
        e %= E_LAYER;
        uint8_t* base = img + int64_t(l) * X3_LAYER;
        if (e < E_QKV) {
            const int n = int(e) / H, k = int(e) % H;
            put_split(base + X3_QKV + (k >> 4) * (QKV * 64), QKV, n, k, s.qkv[l][n * H + k]);
        } else if ((e -= E_QKV) < E_O) {
            const int n = int(e) / H, k = int(e) % H;
            put_split(base + X3_O + (k >> 4) * (H * 64), H, n, k, s.o[l][n * H + k]);
        } else if ((e -= E_O) < E_GU) {
            // one 15 360-byte unit per k-block: G1 = gate | up of channels 0..63 (128 rows), then G2 = gate | up of channels 64..119 (112)
            const int n = int(e) / H, k = int(e) % H;                      // n: gate 0..119 | up 0..119 (the reference's row order)
            const int ch = n % INTER, up = n / INTER;
            uint8_t* unit = base + X3_GU + (k >> 4) * (GU * 64);
            if (ch < 64) put_split(unit, 128, up * 64 + ch, k, s.gu[l][n * H + k]);
            else put_split(unit + G2_OFF, 112, up * 56 + (ch - 64), k, s.gu[l][n * H + k]);
        } else {
            e -= E_GU;
            const int n = int(e) / 128, k = int(e) % 128;                  // K = 120 -> 128, zero padded
            put_split(base + X3_D + (k >> 4) * (H * 64), H, n, k, k < INTER ? s.down[l][n * INTER + k] : 0.f);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// 16 consecutive K-elements of this thread's operand row -> hi | lo fp16 terms in one k-block (128 rows x 32 B per part, 32-byte
// swizzle); `blk` = shared address of the k-block's hi part + row * 32, `sw` = (row >> 2) & 1
__device__ __forceinline__ void store_kblock(uint32_t blk, uint32_t sw, const float* x) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) tc::split2_f16(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
    const uint32_t a0 = blk + (sw << 4), a1 = blk + ((sw ^ 1u) << 4);
    sts128(a0, hi[0], hi[1], hi[2], hi[3]);
    sts128(a1, hi[4], hi[5], hi[6], hi[7]);
    sts128(a0 + LO_OFF, lo[0], lo[1], lo[2], lo[3]);
    sts128(a1 + LO_OFF, lo[4], lo[5], lo[6], lo[7]);
}

// The row threads of a tile and the issuer meet on two mbarriers per tile: `ready` (4 warp arrivals: the operand of the next
// stage is in shared memory and the accumulator has been read) and `done` (the stage's MMAs have completed).
struct TileSync {
    uint64_t *ready, *done;
    uint32_t done_par;
    int lane;
    __device__ __forceinline__ void signal_on(uint64_t* bar) {
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bar);
    }
    __device__ __forceinline__ void signal() { signal_on(ready); }
    __device__ __forceinline__ void wait() {
        tc::mbar_wait(done, done_par);
        done_par ^= 1u;
        tc::fence_after_sync();
    }
    // a completion barrier with one phase per block application (every waiter waits for every phase)
    __device__ __forceinline__ void wait_on(uint64_t* bar, uint32_t& par) {
        tc::mbar_wait(bar, par);
        par ^= 1u;
        tc::fence_after_sync();
    }
};

// hidden = rms_norm(residual + accumulator[0,64)); the residual comes from TMEM (TRES) or is h itself   game.py:1223-1229, 1345-1350
template <bool TRES>
__device__ __forceinline__ void residual_norm(uint32_t tl, float (&h)[H]) {
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < H; c += 16) {
        float d[16], r[16];
        if (TRES) {
            tc::tmem_ld8x2(tl + uint32_t(c), d, tl + T_RES + uint32_t(c), r);
            tc::tmem_ld8x2(tl + uint32_t(c + 8), d + 8, tl + T_RES + uint32_t(c + 8), r + 8);
        } else {
            tc::tmem_ld16p(tl + uint32_t(c), d);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            h[c + j] = (TRES ? r[j] : h[c + j]) + d[j];
            sq = fmaf(h[c + j], h[c + j], sq);
        }
    }
    const float rs = rsqrtf(sq * (1.0f / H) + 1e-5f);
#pragma unroll
    for (int n = 0; n < H; ++n) h[n] *= rs;
}

// the hidden state as the operand of the next projection (R0) and, RES, as the fp32 residual in TMEM
template <bool RES>
__device__ __forceinline__ void publish_hidden(uint32_t x_row, uint32_t sw, uint32_t tl, const float (&h)[H]) {
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
        store_kblock(x_row + uint32_t(kb) * 4096u, sw, &h[16 * kb]);
        if (RES) {
            tc::tmem_st8(tl + T_RES + uint32_t(16 * kb), &h[16 * kb]);
            tc::tmem_st8(tl + T_RES + uint32_t(16 * kb + 8), &h[16 * kb + 8]);
        }
    }
    if (RES) tc::tmem_st_wait();
}

// x * sigmoid(x), two at once on packed fp32 math, with the hardware ex2 / rcp approximations (~1e-6 relative)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// (one rcp for the pair: 1/a = b / (ab), 1/b = a / (ab) -- the SwiGLU phase is bound by the XU pipe.  The exponents are capped at 63
// so that the product ab <= 2^126 can neither overflow nor push its reciprocal below the normal range (where .ftz would flush it
// and zero BOTH sigmoids): beyond x < -43.7 the sigmoid is then 2^-63 instead of something smaller, an error below 1e-17 in
// x * sigmoid(x).)
__device__ __forceinline__ float2 silu2(float2 x) {
    const float2 t = __fmul2_rn(x, make_float2(-1.4426950408889634f, -1.4426950408889634f));
    const float2 d = __fadd2_rn(make_float2(ex2_approx(fminf(t.x, 63.0f)), ex2_approx(fminf(t.y, 63.0f))), make_float2(1.0f, 1.0f));
    const float r = rcp_approx(d.x * d.y);
    return __fmul2_rn(x, __fmul2_rn(make_float2(r, r), make_float2(d.y, d.x)));
}

// silu(gate) * up -> depthwise conv over the tokens (k = 2, pad 1, trimmed: out[t] = w0 x[t-1] + w1 x[t] + b) -> silu, for the 8
// channels ch0.. (gate / up already in registers); `cw0` = the w0 row, or a row of zeros for the first token of an env (no x[t-1]).
// game.py:1264-1276
__device__ __forceinline__ void swiglu8(const float* g, const float* u, const float* __restrict__ cw0, const float* __restrict__ cw, int ch0,
                                        float* x) {
    const float4* w0 = reinterpret_cast<const float4*>(cw0 + ch0);
    const float4* w1 = reinterpret_cast<const float4*>(cw + 128 + ch0);
    const float4* bb = reinterpret_cast<const float4*>(cw + 256 + ch0);
    const float4 wa = w0[0], wb = w0[1], va = w1[0], vb = w1[1], ba = bb[0], bc = bb[1];
    const float2 k0[4] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y), make_float2(wb.z, wb.w)};
    const float2 k1[4] = {make_float2(va.x, va.y), make_float2(va.z, va.w), make_float2(vb.x, vb.y), make_float2(vb.z, vb.w)};
    const float2 kb[4] = {make_float2(ba.x, ba.y), make_float2(ba.z, ba.w), make_float2(bc.x, bc.y), make_float2(bc.z, bc.w)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 cur = __fmul2_rn(silu2(make_float2(g[2 * j], g[2 * j + 1])), make_float2(u[2 * j], u[2 * j + 1]));
        const float2 prev = make_float2(__shfl_up_sync(0xffffffffu, cur.x, 1), __shfl_up_sync(0xffffffffu, cur.y, 1));   // token t-1
        const float2 y = silu2(__ffma2_rn(k0[j], prev, __ffma2_rn(k1[j], cur, kb[j])));
        x[2 * j] = y.x;
        x[2 * j + 1] = y.y;
    }
}

// ---------------------------------------------------------------------------------------------------------------- control warps
__device__ __forceinline__ void producer(Smem& S, uint32_t stage_sets, int L, const uint8_t* img) {
    // stage_sets = block applications of this CTA's whole launch / L; unit k of the launch-wide sequence (period L * UNITS)
    // goes to slot k % RING once both tiles' MMAs on the slot's previous unit have completed (w_empty, 2 arrivals)
    uint32_t slot = 0, par = 0, u = 0;
    int l = 0;
    bool first_lap = true;
    const uint8_t* src = img;
    for (uint32_t k = stage_sets * uint32_t(L) * UNITS; k > 0; --k) {
        if (!first_lap) tc::mbar_wait(&S.w_empty[slot], par);
        const uint32_t bytes = unit_bytes(u);
        if (elect_one()) {
            tc::mbar_expect_tx(&S.w_full[slot], bytes);
            tc::bulk_g2s(S.W[slot], src, bytes, &S.w_full[slot]);
        }
        src += bytes;
        if (++u == UNITS) {
            u = 0;
            if (++l == L) { l = 0; src = img; }
        }
        if (++slot == uint32_t(RING)) {
            slot = 0;
            if (!first_lap) par ^= 1u;
            first_lap = false;
        }
    }
}

__device__ __forceinline__ void issuer(Smem& S, uint32_t tmem_base, uint32_t stage_sets, int L) {
    // descriptors: only the address field (bits 0..13, 16-byte units) moves
    const uint64_t d0 = tc::make_desc_sw32(0, 16, 256);
    const uint32_t d_hi = uint32_t(d0 >> 32), d_lo0 = uint32_t(d0);
    auto desc = [&](uint32_t saddr) { return uint64_t(d_lo0 | ((saddr >> 4) & 0x3FFFu)) | uint64_t(d_hi) << 32; };
    // the three products of one k-block: A at ah (hi; lo at + LO_OFF), B at bh (hi; lo at + wpart)
    auto mma3 = [&](uint32_t d_tmem, uint32_t ah, uint32_t bh, uint32_t wpart, uint32_t idesc, bool acc) {
        tc::mma_bf16_ss(d_tmem, desc(ah + LO_OFF), desc(bh), idesc, acc);
        tc::mma_bf16_ss(d_tmem, desc(ah), desc(bh + wpart), idesc, true);
        tc::mma_bf16_ss(d_tmem, desc(ah), desc(bh), idesc, true);
    };
    const uint32_t xa[TILES] = {tc::smem_addr(S.X[0]), tc::smem_addr(S.X[1])};
    const uint32_t wa = tc::smem_addr(S.W[0]);
    const uint32_t i_qkv = tc::make_idesc_f16(128, QKV), i_h = tc::make_idesc_f16(128, H), i_g1 = tc::make_idesc_f16(128, 128),
                   i_g2 = tc::make_idesc_f16(128, 112);
    uint32_t slot = 0, full_par = 0;               // first ring slot of the current stage and the parity of its w_full phase
    uint32_t ready_par = 0, d1_par = 0, d2_par = 0;  // both tiles' barriers advance together
    auto next = [&](uint32_t& sl, uint32_t& fp) {
        if (++sl == uint32_t(RING)) { sl = 0; fp ^= 1u; }
    };
    for (uint32_t k = stage_sets * uint32_t(L); k > 0; --k) {
        uint32_t sl = slot, fp = full_par;
        // ---- q | k | v: 4 units of one k-block
#pragma unroll
        for (int tile = 0; tile < TILES; ++tile) {
            tc::mbar_wait(&S.ready[tile], ready_par);
            tc::fence_after_sync();
            sl = slot; fp = full_par;
#pragma unroll 1
            for (int u = 0; u < 4; ++u) {
                if (tile == 0) tc::mbar_wait(&S.w_full[sl], fp);      // tile 1 reuses what tile 0 has waited for
                if (elect_one()) {
                    mma3(tmem_base + uint32_t(tile) * 256u, xa[tile] + uint32_t(u) * 4096u, wa + sl * SLOT, QKV * 32u, i_qkv, u > 0);
                    tc::mma_commit(&S.w_empty[sl]);
                }
                __syncwarp();
                next(sl, fp);
            }
            if (elect_one()) tc::mma_commit(&S.done[tile]);
            __syncwarp();
        }
        slot = sl; full_par = fp;
        ready_par ^= 1u;
        // ---- o projection: 2 units of two k-blocks
#pragma unroll
        for (int tile = 0; tile < TILES; ++tile) {
            tc::mbar_wait(&S.ready[tile], ready_par);
            tc::fence_after_sync();
            sl = slot; fp = full_par;
#pragma unroll 1
            for (int u = 0; u < 2; ++u) {
                if (tile == 0) tc::mbar_wait(&S.w_full[sl], fp);
                if (elect_one()) {
                    mma3(tmem_base + uint32_t(tile) * 256u, xa[tile] + uint32_t(2 * u) * 4096u, wa + sl * SLOT, H * 32u, i_h, u > 0);
                    mma3(tmem_base + uint32_t(tile) * 256u, xa[tile] + uint32_t(2 * u + 1) * 4096u, wa + sl * SLOT + H * 64u, H * 32u, i_h, true);
                    tc::mma_commit(&S.w_empty[sl]);
                }
                __syncwarp();
                next(sl, fp);
            }
            if (elect_one()) tc::mma_commit(&S.done[tile]);
            __syncwarp();
        }
        slot = sl; full_par = fp;
        ready_par ^= 1u;
        // ---- gate | up: 4 units [G1 | G2] of one k-block; all of G1 first (its own completion barrier), then G2
#pragma unroll
        for (int tile = 0; tile < TILES; ++tile) {
            tc::mbar_wait(&S.ready[tile], ready_par);
            tc::fence_after_sync();
            sl = slot; fp = full_par;
#pragma unroll 1
            for (int u = 0; u < 4; ++u) {
                if (tile == 0) tc::mbar_wait(&S.w_full[sl], fp);
                if (elect_one()) mma3(tmem_base + uint32_t(tile) * 256u, xa[tile] + uint32_t(u) * 4096u, wa + sl * SLOT, 128u * 32u, i_g1, u > 0);
                __syncwarp();
                next(sl, fp);
            }
            if (elect_one()) tc::mma_commit(&S.done_g1[tile]);
            __syncwarp();
            sl = slot; fp = full_par;
#pragma unroll 1
            for (int u = 0; u < 4; ++u) {
                if (elect_one()) {
                    mma3(tmem_base + uint32_t(tile) * 256u + T_G2, xa[tile] + uint32_t(u) * 4096u, wa + sl * SLOT + G2_OFF, 112u * 32u, i_g2, u > 0);
                    tc::mma_commit(&S.w_empty[sl]);
                }
                __syncwarp();
                next(sl, fp);
            }
            if (elect_one()) tc::mma_commit(&S.done_g2[tile]);
            __syncwarp();
        }
        slot = sl; full_par = fp;
        ready_par ^= 1u;
        // ---- down projection: 4 units of two k-blocks; k-blocks 0..3 (R1) once half 0 has written them, 4..7 (R0) after half 1
#pragma unroll
        for (int tile = 0; tile < TILES; ++tile) {
            sl = slot; fp = full_par;
#pragma unroll 1
            for (int u = 0; u < 4; ++u) {
                if (u == 0) {
                    tc::mbar_wait(&S.rdy_d1[tile], d1_par);
                    tc::fence_after_sync();
                } else if (u == 2) {
                    tc::mbar_wait(&S.rdy_d2[tile], d2_par);
                    tc::fence_after_sync();
                }
                if (tile == 0) tc::mbar_wait(&S.w_full[sl], fp);
                if (elect_one()) {
                    const uint32_t a0 = xa[tile] + (u < 2 ? REGION + uint32_t(2 * u) * 4096u : uint32_t(2 * u - 4) * 4096u);
                    mma3(tmem_base + uint32_t(tile) * 256u, a0, wa + sl * SLOT, H * 32u, i_h, u > 0);
                    mma3(tmem_base + uint32_t(tile) * 256u, a0 + 4096u, wa + sl * SLOT + H * 64u, H * 32u, i_h, true);
                    tc::mma_commit(&S.w_empty[sl]);
                }
                __syncwarp();
                next(sl, fp);
            }
            if (elect_one()) tc::mma_commit(&S.done[tile]);
            __syncwarp();
        }
        slot = sl; full_par = fp;
        d1_par ^= 1u;
        d2_par ^= 1u;
    }
}

// ---------------------------------------------------------------------------------------------------------------- kernel
__global__ void __launch_bounds__(XTHREADS, 1) rollout_urm_x3_kernel(RolloutParams p, int loops) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (tc::smem_addr(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.layers;
    const int64_t ntiles = (p.B + 7) / 8, npairs = (ntiles + TILES - 1) / TILES;
    const int64_t my_pairs = npairs > blockIdx.x ? (npairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const float* pk = p.packed;
    const float* emb_tab = pk + x3_base(L);
    const uint8_t* img = reinterpret_cast<const uint8_t*>(emb_tab + X3_EMB_FLOATS);
    const uint32_t stage_sets = uint32_t(my_pairs) * uint32_t(p.T) * uint32_t(loops);

    if (warp == 0) tc::tmem_alloc(&S.tmem_base, 512);
    if (tid == 0) {
        for (int t = 0; t < TILES; ++t) {
            tc::mbar_init(&S.ready[t], 4);
            tc::mbar_init(&S.rdy_d1[t], 4);
            tc::mbar_init(&S.rdy_d2[t], 4);
            tc::mbar_init(&S.done[t], 1);
            tc::mbar_init(&S.done_g1[t], 1);
            tc::mbar_init(&S.done_g2[t], 1);
        }
        for (int s = 0; s < RING; ++s) {
            tc::mbar_init(&S.w_full[s], 1);
            tc::mbar_init(&S.w_empty[s], TILES);
        }
        tc::mbar_fence_init();
    }
    for (int i = tid; i < 5 * H + 8; i += XTHREADS) S.headw[i] = pk[F_HEADW + i];
    for (int i = tid; i < L * F_CONV_STRIDE; i += XTHREADS) (&S.conv[0][0][0])[i] = pk[F_CONV + i];
    for (int i = tid; i < 128; i += XTHREADS) S.zeros[i] = 0.f;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = S.tmem_base;

    if (warp == ROW_THREADS / 32) {
        issuer(S, tmem_base, stage_sets, L);
    } else if (warp == ROW_THREADS / 32 + 1) {
        producer(S, stage_sets, L, img);
    } else {
        const int tile = warp >> 2, row = tid & 127, cell = row & 15;
        const uint32_t tl = tmem_base + (uint32_t((warp & 3) * 32) << 16) + uint32_t(tile) * 256u;
        const uint32_t xb = tc::smem_addr(S.X[tile]);
        const uint32_t x_row = xb + uint32_t(row) * 32u, sw = uint32_t(row >> 2) & 1u;
        const uint32_t kv_slot = xb + (uint32_t(row) ^ (uint32_t(row >> 4) & 1u)) * 16u;     // this token's 16-byte slot of a K | V chunk
        // the 16 slots of this env; plain loads (not asm) so that the compiler can request a head's K or V rows well ahead of their use
        const float4* const kv_env = reinterpret_cast<const float4*>(S.X[tile]) + (row & ~15);
        const int bar_id = 1 + tile;
        const LutGlobal lut{p.lut};
        TileSync ts{&S.ready[tile], &S.done[tile], 0u, lane};
        uint32_t g1_par = 0, g2_par = 0;
        for (int64_t pr = 0; pr < my_pairs; ++pr) {
            const int64_t env = ((int64_t(blockIdx.x) + pr * gridDim.x) * TILES + tile) * 8 + (row >> 4);
            const bool owner = env < p.B && cell == 0;              // the env's leader thread owns the board
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
                const uint32_t ex = ((cell < 8 ? blo : bhi) >> (4 * (cell & 7))) & 15u;
                const float4* emb4 = reinterpret_cast<const float4*>(emb_tab + (ex * SEQ + uint32_t(cell)) * H);   // game.py:1376-1380
                float h[H];
                {
                    const float4* i4 = reinterpret_cast<const float4*>(pk + F_INIT + cell * H);                    // game.py:1431
#pragma unroll
                    for (int c = 0; c < H / 4; ++c) {
                        const float4 v = __ldg(i4 + c);
                        h[4 * c] = v.x; h[4 * c + 1] = v.y; h[4 * c + 2] = v.z; h[4 * c + 3] = v.w;
                    }
                }
                for (int loop = 0; loop < loops; ++loop) {
                    // hidden += input_embeddings   game.py:1441,1447
#pragma unroll
                    for (int c = 0; c < H / 4; ++c) {
                        const float4 v = __ldg(emb4 + c);
                        h[4 * c] += v.x; h[4 * c + 1] += v.y; h[4 * c + 2] += v.z; h[4 * c + 3] += v.w;
                    }
                    for (int l = 0; l < L; ++l) {
                        // ---------------- attention   game.py:1296-1317
                        publish_hidden<true>(x_row, sw, tl, h);
                        ts.signal();
                        ts.wait();                                               // q | k | v in the accumulator
                        // K | V -> shared memory, fp32, chunk-major: chunk c (4 floats) of token m at c * 2048 + slot(m) * 16
#pragma unroll
                        for (int c = 0; c < 2 * H; c += 16) {
                            float kv[16];
                            tc::tmem_ld16p(tl + uint32_t(H + c), kv);
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                sts128(kv_slot + uint32_t(c / 4 + q) * 2048u, __float_as_uint(kv[4 * q]), __float_as_uint(kv[4 * q + 1]),
                                       __float_as_uint(kv[4 * q + 2]), __float_as_uint(kv[4 * q + 3]));
                        }
                        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                        // Every 128-bit K / V load serves the two envs of the warp only: the attention is bound by shared-memory
                        // wavefronts.  So a thread works on TWO queries per K / V row -- its own token's and its lane neighbour's
                        // (lane ^ 1, same env) -- for two of the four heads (even lanes heads 0, 1; odd lanes heads 2, 3): half the
                        // loads, the queries of the other two heads and the finished outputs cross by shuffles.
                        const int odd = lane & 1;
                        float om[2 * HD], ox[2 * HD];                             // outputs of my two heads: my token, the neighbour's
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int hd = 2 * odd + j;                           // my head; 2 * (1 - odd) + j goes to the neighbour
                            float qlo[HD], qhi[HD];                               // (TMEM addresses are warp-uniform: both heads, then select)
                            tc::tmem_ld16p(tl + uint32_t(j * HD), qlo);
                            tc::tmem_ld16p(tl + uint32_t((2 + j) * HD), qhi);
                            float2 qa[HD / 2], qb[HD / 2];                        // my token's query, the neighbour's
#pragma unroll
                            for (int d = 0; d < HD / 2; ++d) {    // 1/sqrt(head_dim), and log2(e): the softmax below is in base 2
                                const float sc0 = 0.25f * 1.4426950408889634f;
                                const float m0 = odd ? qhi[2 * d] : qlo[2 * d], m1 = odd ? qhi[2 * d + 1] : qlo[2 * d + 1];
                                const float s0 = odd ? qlo[2 * d] : qhi[2 * d], s1 = odd ? qlo[2 * d + 1] : qhi[2 * d + 1];
                                qa[d] = make_float2(m0 * sc0, m1 * sc0);
                                qb[d] = make_float2(__shfl_xor_sync(0xffffffffu, s0, 1) * sc0, __shfl_xor_sync(0xffffffffu, s1, 1) * sc0);
                            }
                            const float4* const kp = kv_env + hd * 4 * 128;
                            float sa[SEQ], sb[SEQ], mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
                            for (int s2 = 0; s2 < SEQ; ++s2) {
                                float2 aa = make_float2(0.f, 0.f), ab = make_float2(0.f, 0.f);
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    const float4 kk = kp[c * 128 + s2];
                                    aa = __ffma2_rn(qa[2 * c], make_float2(kk.x, kk.y), aa);
                                    aa = __ffma2_rn(qa[2 * c + 1], make_float2(kk.z, kk.w), aa);
                                    ab = __ffma2_rn(qb[2 * c], make_float2(kk.x, kk.y), ab);
                                    ab = __ffma2_rn(qb[2 * c + 1], make_float2(kk.z, kk.w), ab);
                                }
                                sa[s2] = aa.x + aa.y;
                                sb[s2] = ab.x + ab.y;
                                mxa = fmaxf(mxa, sa[s2]);
                                mxb = fmaxf(mxb, sb[s2]);
                            }
                            float dena = 0.f, denb = 0.f;
#pragma unroll
                            for (int s2 = 0; s2 < SEQ; ++s2) {
                                sa[s2] = ex2_approx(sa[s2] - mxa);
                                sb[s2] = ex2_approx(sb[s2] - mxb);
                                dena += sa[s2];
                                denb += sb[s2];
                            }
                            const float inva = __fdividef(1.0f, dena), invb = __fdividef(1.0f, denb);
                            float2 oa[HD / 2], ob[HD / 2];
#pragma unroll
                            for (int d = 0; d < HD / 2; ++d) oa[d] = ob[d] = make_float2(0.f, 0.f);
#pragma unroll
                            for (int s2 = 0; s2 < SEQ; ++s2) {
                                const float2 pa = make_float2(sa[s2], sa[s2]), pb = make_float2(sb[s2], sb[s2]);
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    const float4 vv = kp[(16 * 128) + c * 128 + s2];
                                    oa[2 * c] = __ffma2_rn(pa, make_float2(vv.x, vv.y), oa[2 * c]);
                                    oa[2 * c + 1] = __ffma2_rn(pa, make_float2(vv.z, vv.w), oa[2 * c + 1]);
                                    ob[2 * c] = __ffma2_rn(pb, make_float2(vv.x, vv.y), ob[2 * c]);
                                    ob[2 * c + 1] = __ffma2_rn(pb, make_float2(vv.z, vv.w), ob[2 * c + 1]);
                                }
                            }
#pragma unroll
                            for (int d = 0; d < HD / 2; ++d) {
                                om[j * HD + 2 * d] = oa[d].x * inva;
                                om[j * HD + 2 * d + 1] = oa[d].y * inva;
                                ox[j * HD + 2 * d] = ob[d].x * invb;
                                ox[j * HD + 2 * d + 1] = ob[d].y * invb;
                            }
                        }
#pragma unroll
                        for (int d = 0; d < 2 * HD; ++d) ox[d] = __shfl_xor_sync(0xffffffffu, ox[d], 1);   // now: my token, the other two heads
                        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // every token of the tile has read K | V
#pragma unroll
                        for (int j = 0; j < 2; ++j) {                                     // k-block = head
                            store_kblock(x_row + uint32_t(2 * odd + j) * 4096u, sw, &om[16 * j]);
                            store_kblock(x_row + uint32_t(2 * (1 - odd) + j) * 4096u, sw, &ox[16 * j]);
                        }
                        ts.signal();
                        ts.wait();                                               // o projection in [0,64)
                        residual_norm<true>(tl, h);
                        // ---------------- ConvSwiGLU   game.py:1264-1276   (h stays in registers as the residual)
                        publish_hidden<false>(x_row, sw, tl, h);
                        ts.signal();
                        const float* cw = &S.conv[l][0][0];
                        const float* cw0 = cell == 0 ? S.zeros : cw;
                        // channels 0..63 (G1, [0,128)) -> R1, then channels 64..119 (G2, [128,240)) -> R0; 8 channels at a time, the
                        // gate / up columns of the next group requested before this one is worked on
#pragma unroll 1
                        for (int part = 0; part < 2; ++part) {
                            ts.wait_on(part == 0 ? &S.done_g1[tile] : &S.done_g2[tile], part == 0 ? g1_par : g2_par);
                            const int ngrp = part == 0 ? 8 : 7, ch_base = 64 * part;
                            const uint32_t g_col = part == 0 ? 0u : T_G2, u_col = part == 0 ? 64u : T_G2 + 56u;
                            const uint32_t x_out = x_row + (part == 0 ? REGION : 0u);
                            uint32_t gr[8], ur[8];
                            tc::tmem_ld8_issue(tl + g_col, gr);
                            tc::tmem_ld8_issue(tl + u_col, ur);
#pragma unroll 1
                            for (int kb = 0; kb < 4; ++kb) {
                                float x[16];
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh) {
                                    const int grp = 2 * kb + hh;
                                    if (grp < ngrp) {
                                        float g[8], u[8];
                                        tc::tmem_ld_wait_all();
#pragma unroll
                                        for (int j = 0; j < 8; ++j) {
                                            g[j] = tc::tmem_ld_pin(gr[j]);
                                            u[j] = tc::tmem_ld_pin(ur[j]);
                                        }
                                        if (grp + 1 < ngrp) {
                                            tc::tmem_ld8_issue(tl + g_col + uint32_t(8 * (grp + 1)), gr);
                                            tc::tmem_ld8_issue(tl + u_col + uint32_t(8 * (grp + 1)), ur);
                                        }
                                        swiglu8(g, u, cw0, cw, ch_base + 8 * grp, &x[8 * hh]);
                                    } else {
#pragma unroll
                                        for (int j = 0; j < 8; ++j) x[8 * hh + j] = 0.f;                  // K padding 120..127
                                    }
                                }
                                store_kblock(x_out + uint32_t(kb) * 4096u, sw, x);
                            }
                            ts.signal_on(part == 0 ? &S.rdy_d1[tile] : &S.rdy_d2[tile]);
                        }
                        ts.wait();                                               // down projection in [0,64)
                        residual_norm<false>(tl, h);
                    }
                }
                // ---- mean-pool over the 16 tokens + heads   game.py:1451-1456
                float out[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    float s = 0.f;
#pragma unroll
                    for (int n = 0; n < H; ++n) s = fmaf(S.headw[j * H + n], h[n], s);
#pragma unroll
                    for (int m = 8; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
                    out[j] = s * (1.0f / SEQ) + S.headw[5 * H + j];
                }
                if (owner) policy_env_step(p, lut, t, env, ctr, lm, out, board, alive);
            }
            if (owner) {
                p.boards[env] = pack_board(board);
                if (p.alive) p.alive[env] = alive ? 1 : 0;
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace x3

int launch_urm_x3_pack(const PackSrc& s, int L, float* packed, cudaStream_t st) {
    x3::pack_x3_kernel<<<256, 256, 0, st>>>(s, L, packed);
    G2048_CHECK_LAUNCH("urm::x3::pack_x3_kernel");
    return G2048_OK;
}

int launch_rollout_urm_x3(const RolloutParams& p, int loops, cudaStream_t st) {
    const int smem = int(sizeof(x3::Smem)) + 1024;
    G2048_CHECK_CUDA(ensure_smem(x3::rollout_urm_x3_kernel, smem));
    const int64_t ntiles = (p.B + 7) / 8, npairs = (ntiles + x3::TILES - 1) / x3::TILES;
    const int grid = int(npairs < num_sms() ? npairs : num_sms());
    const int64_t pairs_per_cta = (npairs + grid - 1) / grid;
    if (pairs_per_cta * p.T * loops * p.layers * x3::UNITS >= (int64_t(1) << 31))
        return fail(G2048_ESHAPE, "g2048_rollout_urm: B * T too large for one launch; split the horizon");
    x3::rollout_urm_x3_kernel<<<grid, x3::XTHREADS, smem, st>>>(p, loops);
    G2048_CHECK_LAUNCH("rollout_urm_x3_kernel");
    return G2048_OK;
}

}  // namespace urm
}  // namespace g2048
