// g2048_update_fused.cu -- the policy update's forward, loss and backward-data pass of GameMLP as ONE
// persistent tcgen05 kernel (train.py:491-556: model(x), the loss section, loss.backward()).
//
// A CTA owns tiles of 128 samples for the whole chain; sample m of the tile is TMEM lane m:
//
//   boards -> x (16 exponents)                                             game.py:92-101
//   z0 = x W0e^T + b0 ; h0 = relu(LN0(z0))                                 game.py:1069-1073
//   z_l = h_{l-1} W_l^T ; h_l = h_{l-1} + relu(LN_l(z_l)),  l = 1..L       game.py:1038-1046
//   logits, V = heads(h_L) ; per-sample PPO-clip / critic / entropy terms  train.py:497-554
//   dh_L = heads^T (dlogits, dV) ; for l = L..0:  g = dh_l * [LN_l(z_l) > 0], dz_l = LN_l-backward(g),
//   dh_{l-1} = dh_l + dz_l W_l                                             (what autograd would run)
//
// Every GEMM is a split-bf16 tensor-core product with fp32 accumulation in tensor memory (g2048_linear.cu).
// The FORWARD products, which decide every ReLU branch and the loss, use three bf16 terms per operand
// (x = hi + lo + r, 24 mantissa bits) and six products, hi*hi + hi*lo + lo*hi + lo*lo + r*hi + hi*r ("x6",
// fp32-grade: measured ~1e-6 of scale): a first MMA phase on (hi, lo), then the row threads overwrite the lo
// operand with r (recomputed from the fp32 residual stream in TMEM) and a second phase adds the r terms.
// The backward products (smooth in their inputs) use two terms and three products ("x3", ~1e-5).
// LayerNorm, ReLU, residual stream, heads, loss and LayerNorm-backward are fp32 on the
// CUDA cores, thread-per-row (4 threads share a row, one column quarter each).  The residual stream
// h (forward) / dh (backward) lives in TMEM columns [256, 256+HP), the accumulator in [0, HP).
// Weight k-blocks (hi|lo, HP x 32 B each) stream L2 -> SMEM through a 5-slot ring of bulk async copies in
// the fixed order the chain consumes them; one thread refills it while the MMAs of a stage drain it.
//
// What leaves the SM per sample: h_0..h_L and dz_0..dz_L (fp32, tiled; the operands of the weight-gradient
// GEMMs dW_l = dz_l^T h_{l-1}, run afterwards by x3_wgrad_kernel), the 5 head gradients, and nothing
// else: z_l, LN statistics, logits, per-sample loss terms never touch HBM (z_0..z_{L-1} round-trip
// through a per-CTA scratch that stays in L2).  LayerNorm-parameter gradients are column sums over
// samples: reduced per warp through shared memory, added to per-(CTA, lane-quarter) fp32 partials with
// single-writer red.global (deterministic), summed in fixed order by a second kernel.
#include <cfloat>
#include "g2048_device.cuh"
#include "g2048_host.h"
#include "g2048_loss.cuh"
#include "g2048_tc.cuh"

namespace g2048 {
namespace uf {

constexpr int MAXH = 208, MAXL = 2, MAXKB = MAXH / 16;
constexpr int SPLIT = 4;
constexpr int ROW_THREADS = 128 * SPLIT;        // 16 warps: warp w -> lane quarter (w & 3), column part (w >> 2)
constexpr int THREADS = ROW_THREADS;            // thread 0 issues the MMAs, thread 32 streams the weights while both wait
                                                // for a stage (4 warps per scheduler -> 128 registers per thread)
constexpr int RING = 4;
constexpr uint32_t X_COL = 256;
constexpr uint32_t A_PART = MAXKB * 4096;       // one operand part: 13 k-blocks of 128 rows x 32 B
constexpr uint32_t B_SLOT = MAXH * 96;          // one weight k-block: hi | lo | r (forward) or hi | lo (backward)
constexpr uint32_t ROLL_VALID = 0x80u;

struct Params {
    int64_t n, ntiles;
    int h, HP, L, decouple;
    int backward;                 // 0: forward only (logits / value out), 1: forward + loss + backward-data
    const uint64_t* boards;
    const uint8_t *actions, *legal, *flags;
    const float* old_logp;
    int old_stride;
    const float *adv, *g_norm;
    float clip_eps, c_v, beta_ent, inv_n;
    uint32_t drop_thr;            // dropout (game.py:1042): an element is dropped iff its 16-bit Philox draw < drop_thr; 0 = off
    float drop_scale;             // 1 / (1 - p)
    uint64_t drop_seed, sample0;  // Philox key; index of this call's first sample in the mask's counter space
    const float* pf;              // fp32 section of the pack
    const uint8_t* img;           // weight k-blocks in consumption order
    float* h_out;                 // [L+1][ntiles] bf16 hi|lo operand images of 128 samples x HP (see store_image), 4 B / value
    float* dz_out;                // same layout
    float* dhead;                 // [n][8]: d loss / d (logits, V), 3 zero pads
    float* logits;                // [n][4] or NULL
    float* value;                 // [n] or NULL
    float* zscratch;              // [grid][L][HP/8][128][8]
    float* ln_part;               // [grid][4][L+1][2][HP]  (zeroed by the caller)
    float* head_part;             // [grid][8]: d head biases (5) + pad
    double* loss_part;            // [grid][4]
};

// fp32 section: b0 | gamma[L+1] | beta[L+1] | head weights [5] (rows of HP) | head biases [8]
__host__ __device__ inline int64_t pf_b0(int) { return 0; }
__host__ __device__ inline int64_t pf_gamma(int HP, int l) { return int64_t(1 + l) * HP; }
__host__ __device__ inline int64_t pf_beta(int HP, int L, int l) { return int64_t(2 + L + l) * HP; }
__host__ __device__ inline int64_t pf_headw(int HP, int L) { return int64_t(3 + 2 * L) * HP; }
__host__ __device__ inline int64_t pf_headb(int HP, int L) { return int64_t(8 + 2 * L) * HP; }
__host__ __device__ inline int64_t pf_floats(int HP, int L) { return pf_headb(HP, L) + 8; }
__host__ __device__ inline int64_t img_offset_bytes(int HP, int L) { return (pf_floats(HP, L) * 4 + 1023) / 1024 * 1024; }
// weight image: forward blocks (stem, then W_1..W_L, 3 parts each), then the transposed blocks W_L^T..W_1^T (2 parts).
// A tile consumes: stem | per block l: the KB forward blocks TWICE (the two MMA phases of the x6 product) | the
// transposed blocks in order.
__host__ __device__ inline int fwd_blocks(int HP, int L) { return 1 + L * (HP / 16); }
__host__ __device__ inline int blocks_per_tile(int HP, int L, bool backward) { return 1 + (backward ? 3 : 2) * L * (HP / 16); }
__host__ __device__ inline int64_t pack_bytes(int HP, int L) {
    return img_offset_bytes(HP, L) + int64_t(fwd_blocks(HP, L)) * HP * 96 + int64_t(L) * (HP / 16) * HP * 64;
}

struct Smem {
    alignas(1024) uint8_t A[2][A_PART];          // activations / gradients, bf16 hi | lo
    alignas(1024) uint8_t B[RING][B_SLOT];
    alignas(16) float b0[MAXH];
    alignas(16) float gamma[MAXL + 1][MAXH];
    alignas(16) float beta[MAXL + 1][MAXH];
    alignas(16) float headw[5][MAXH];
    alignas(16) float headb[8];
    float red[2][SPLIT][128];
    float stats[MAXL + 1][2][128];               // mean, rstd of every LayerNorm row
    alignas(16) float dhead[128][8];
    alignas(16) float scratch[ROW_THREADS / 32][8][32];   // per-warp transposition buffer (column sums, head partials)
    double lsum[128][4];
    uint64_t a_ready, mma_done, b_full[RING], b_empty[RING];
    uint32_t tmem_base;
};
static_assert(sizeof(Smem) + 1024 <= 232448, "update kernel exceeds the 227 KB shared memory limit");

__device__ __forceinline__ void row_sync() { asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory"); }

__device__ __forceinline__ void red_add(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// 8 consecutive fp32 values -> their bf16 hi / lo terms (x = hi + lo to 16 mantissa bits), packed 8 x bf16 = 16 bytes each
__device__ __forceinline__ void split8(const float* x, uint4& hi4, uint4& lo4) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(x[2 * q], x[2 * q + 1]);
        const float2 f = __bfloat1622float2(h2);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(x[2 * q] - f.x, x[2 * q + 1] - f.y);
        hi[q] = *reinterpret_cast<const uint32_t*>(&h2);
        lo[q] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    hi4 = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    lo4 = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
// 8 consecutive columns (col % 8 == 0) of row `row` -> bf16 hi/lo operand bytes
__device__ __forceinline__ void store_operand(Smem& S, int row, int col, const uint4& hi4, const uint4& lo4) {
    const uint32_t off = uint32_t(col >> 4) * 4096u + uint32_t(row) * 32u + uint32_t(((((col >> 3) & 1) ^ (row >> 2)) & 1) << 4);
    *reinterpret_cast<uint4*>(S.A[0] + off) = hi4;
    *reinterpret_cast<uint4*>(S.A[1] + off) = lo4;
}
__device__ __forceinline__ void store_operand(Smem& S, int row, int col, const float* x) {
    uint4 hi4, lo4;
    split8(x, hi4, lo4);
    store_operand(S, row, col, hi4, lo4);
}
// The operand tile in shared memory ([hi | lo][16-feature block][sample 0..127][32 B], 13 x 4 KiB per part) doubles as
// the tensor's image in HBM: x3_wgrad_kernel bulk-copies the 32-sample slices of its blocks straight into its ring, so
// the weight-gradient pass neither converts nor touches registers, and this kernel writes h_l / dz_l with two bulk
// copies per tile (full lines) instead of per-thread stores.  Rows past the sample count are zeroed on the way in
// (the weight gradient sums whole stages).  Same 4 bytes per value as fp32.
__device__ __forceinline__ void copy_out_tile(Smem& S, uint8_t* tile_img, int HP) {
    const uint32_t part = uint32_t(HP >> 4) * 4096u;
    tc::bulk_s2g(tile_img, S.A[0], part);
    tc::bulk_s2g(tile_img + part, S.A[1], part);
    tc::bulk_commit();
}

// third split term r = x - hi - lo of 8 consecutive columns -> the lo operand buffer (second MMA phase)
__device__ __forceinline__ void store_operand_r(Smem& S, int row, int col, const float* x) {
    uint32_t rr[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(x[2 * q], x[2 * q + 1]);
        const float2 f = __bfloat1622float2(h2);
        const float e0 = x[2 * q] - f.x, e1 = x[2 * q + 1] - f.y;
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(e0, e1);
        const float2 g = __bfloat1622float2(l2);
        const __nv_bfloat162 r2 = __floats2bfloat162_rn(e0 - g.x, e1 - g.y);
        rr[q] = *reinterpret_cast<const uint32_t*>(&r2);
    }
    const uint32_t off = uint32_t(col >> 4) * 4096u + uint32_t(row) * 32u + uint32_t(((((col >> 3) & 1) ^ (row >> 2)) & 1) << 4);
    *reinterpret_cast<uint4*>(S.A[1] + off) = make_uint4(rr[0], rr[1], rr[2], rr[3]);
}

// 32-byte (one sector) global accesses: the activation / gradient tensors this kernel writes are TILED,
// [tile][column group of 8][row 0..127][8 floats], so that the 32 rows of a warp store 1 KiB contiguously
// (row-major rows are 784 B apart: 32 separate sectors per store instruction, measured 2x sector traffic)
__device__ __forceinline__ void st256(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld256(const float* p, float* v) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p)
                 : "memory");
}

// sum over the 4 column parts of a row (each part contributes one value per slot)
__device__ __forceinline__ float exchange(Smem& S, int slot, int part, int row, float v) {
    S.red[slot][part][row] = v;
    row_sync();
    return (S.red[slot][0][row] + S.red[slot][1][row]) + (S.red[slot][2][row] + S.red[slot][3][row]);
}

// Column sums over the 32 rows of a warp for 4 columns of two tensors; the 8 totals go to `dst_a[0..3]`
// and `dst_b[0..3]` with single-writer red.global (the same lane writes the same address every tile).
__device__ __forceinline__ void colsum4x2(float (*scr)[32], int lane, const float* a, const float* b, float* dst_a, float* dst_b) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        scr[j][lane] = a[j];
        scr[4 + j][lane] = b[j];
    }
    __syncwarp();
    const int v = lane >> 2, seg = lane & 3;
    const float4 p0 = *reinterpret_cast<const float4*>(&scr[v][seg * 8]), p1 = *reinterpret_cast<const float4*>(&scr[v][seg * 8 + 4]);
    float s = ((p0.x + p0.y) + (p0.z + p0.w)) + ((p1.x + p1.y) + (p1.z + p1.w));
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (seg == 0) red_add(v < 4 ? dst_a + v : dst_b + (v - 4), s);
    __syncwarp();
}

// Dropout mask of 8 consecutive columns of one sample in block l (game.py:1038-1046: x + Dropout(ReLU(LN(Linear x)))):
// one Philox4x32-10 call, counter = (sample index, l, column group), key = the call's dropout seed; column j of the
// group is KEPT iff the j-th 16-bit lane of the 128 random bits is >= drop_thr = round(p * 65536).  Bit j of the result.
__device__ __forceinline__ uint32_t dropout_keep8(const Params& p, int64_t sample, int l, int col) {
    const U4 r = philox4x32_10(uint32_t(sample), uint32_t(uint64_t(sample) >> 32), uint32_t(l), uint32_t(col >> 3),
                               uint32_t(p.drop_seed), uint32_t(p.drop_seed >> 32));
    const uint32_t t2 = p.drop_thr * 0x00010001u;
    const uint32_t a = __vsetgeu2(r.x, t2), b = __vsetgeu2(r.y, t2), c = __vsetgeu2(r.z, t2), d = __vsetgeu2(r.w, t2);   // bits 0, 16
    return (a & 1u) | ((a >> 15) & 2u) | ((b & 1u) << 2) | ((b >> 13) & 8u) | ((c & 1u) << 4) | ((c >> 11) & 32u) | ((d & 1u) << 6) |
           ((d >> 9) & 128u);
}

struct RowCtx {
    int row, part, lane, warp, c0, ng;      // this thread's row, column part, first column, number of 8-column groups
    uint32_t tD, tX;                         // TMEM addresses of (lane quarter, column c0) in D and X
    int64_t grow;                            // global sample index
    int64_t tile;                            // global tile index (grow / 128)
    bool valid;                              // grow < n
};

// Forward epilogue of LayerNorm l: h_l = [h_{l-1} +] relu(LN(z_l)); writes X (TMEM), the next A operand,
// h_out[l], the z scratch (l < L) and, for l == L, the 5 head dot products (complete on part 0).
template <bool STEM, bool DROP>
__device__ __forceinline__ void fwd_epilogue(Smem& S, const Params& p, const RowCtx& c, int l, float (&o)[5], uint64_t& keep_bits) {
    const int HP = p.HP, h = p.h, L = p.L;
    const bool last = l == L;
    const float inv_h = 1.0f / float(h);
    const float* gam = S.gamma[l] + c.c0;
    const float* bet = S.beta[l] + c.c0;
    // one pass: sum and sum of squares of the row (padded columns hold z = 0 and add nothing)
    float sum = 0.f, sq = 0.f;
#pragma unroll 1
    for (int g = 0; g < c.ng; g += 2) {
        float v[8], w[8];
        if (g + 1 < c.ng) {
            tc::tmem_ld8x2(c.tD + uint32_t(8 * g), v, c.tD + uint32_t(8 * g + 8), w);
        } else {
            tc::tmem_ld8(c.tD + uint32_t(8 * g), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float a = STEM ? v[j] + S.b0[c.c0 + 8 * g + j] : v[j];
            const float b = (STEM && g + 1 < c.ng) ? w[j] + S.b0[c.c0 + 8 * g + 8 + j] : w[j];
            sum += a + b;
            sq = fmaf(a, a, fmaf(b, b, sq));
        }
    }
    S.red[0][c.part][c.row] = sum;
    S.red[1][c.part][c.row] = sq;
    row_sync();
    const float mean = ((S.red[0][0][c.row] + S.red[0][1][c.row]) + (S.red[0][2][c.row] + S.red[0][3][c.row])) * inv_h;
    const float msq = ((S.red[1][0][c.row] + S.red[1][1][c.row]) + (S.red[1][2][c.row] + S.red[1][3][c.row])) * inv_h;
    const float var = fmaxf(msq - mean * mean, 0.f);
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    const float shift = -mean * rstd;       // xhat = z * rstd + shift, the same expression in the backward pass
    if (c.part == 0) {
        S.stats[l][0][c.row] = mean;
        S.stats[l][1][c.row] = rstd;
    }
    // tiled addresses of (this row, column group 0); + 1024 floats per column group
    float* zrow = last ? nullptr : p.zscratch + ((size_t(blockIdx.x) * L + l) * (HP / 8) * 128 + c.row) * 8;
#pragma unroll
    for (int q = 0; q < 5; ++q) o[q] = 0.f;
#pragma unroll 1
    for (int g = 0; g < c.ng; ++g) {
        const int col = c.c0 + 8 * g;
        float z[8], x[8];
        if (STEM) tc::tmem_ld8(c.tD + uint32_t(8 * g), z);
        else tc::tmem_ld8x2(c.tD + uint32_t(8 * g), z, c.tX + uint32_t(8 * g), x);
        uint32_t keep = 0xFFu;
        if (!STEM && DROP) {
            keep = dropout_keep8(p, p.sample0 + c.grow, l, col);
            keep_bits |= uint64_t(keep) << (8 * g);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (STEM) z[j] += S.b0[col + j];
            const float y = fmaf(fmaf(z[j], rstd, shift), gam[8 * g + j], bet[8 * g + j]);
            float r = fmaxf(y, 0.f);
            if (!STEM && DROP) r = ((keep >> j) & 1u) ? r * p.drop_scale : 0.f;
            x[j] = STEM ? r : x[j] + r;
        }
        tc::tmem_st8(c.tX + uint32_t(8 * g), x);
        uint4 hi4, lo4;
        split8(x, hi4, lo4);
        if (!c.valid) hi4 = lo4 = make_uint4(0u, 0u, 0u, 0u);      // only in the last tile of a launch
        store_operand(S, c.row, col, hi4, lo4);
        if (zrow && p.backward) st256(zrow + size_t(col) * 128, z);
        if (last) {
#pragma unroll
            for (int q = 0; q < 5; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) o[q] = fmaf(S.headw[q][col + j], x[j], o[q]);
        }
    }
    tc::tmem_st_wait();
    if (last) {
        tc::fence_async_smem();      // the operand tile (h_L) is bulk-copied out after the next barrier
        // partial head dots of parts 1..3 -> part 0 (through the dhead buffer, free at this point)
        if (c.part != 0) {
#pragma unroll
            for (int q = 0; q < 5; ++q) S.scratch[c.part * 4 + (c.row >> 5)][q][c.lane] = o[q];
        }
        row_sync();
        if (c.part == 0) {
#pragma unroll
            for (int q = 0; q < 5; ++q) {
#pragma unroll
                for (int part = 1; part < SPLIT; ++part) o[q] += S.scratch[part * 4 + (c.row >> 5)][q][c.lane];
                o[q] += S.headb[q];
            }
        }
        row_sync();      // scratch is reused by the column sums of the backward pass
    }
}

// Backward through LayerNorm l and its ReLU.  first (l == L): z_L is still in D and dh_L comes from the head
// gradients; otherwise z_l comes from the scratch and dh_l = X + D (D = dz_{l+1} W_{l+1}).
// Pass A stores xhat in D and dh_l in X, pass B turns them into dz_l (-> dz_out[l], next A operand).
template <bool DROP>
__device__ __forceinline__ void bwd_epilogue(Smem& S, const Params& p, const RowCtx& c, int l, bool first, uint64_t keep_bits) {
    // keep_bits: the forward's dropout mask of this thread's columns in block l (bit 8 g + j); the gradient passes through
    // Dropout as g * keep / (1 - p)
    const float dscale = (DROP && l > 0) ? p.drop_scale : 1.0f;
    if (!(DROP && l > 0)) keep_bits = ~0ull;
    const int HP = p.HP, h = p.h, L = p.L;
    const float inv_h = 1.0f / float(h);
    const float mean = S.stats[l][0][c.row], rstd = S.stats[l][1][c.row], shift = -mean * rstd;
    const float* gam = S.gamma[l] + c.c0;
    const float* bet = S.beta[l] + c.c0;
    const float* zrow = first ? nullptr : p.zscratch + ((size_t(blockIdx.x) * L + l) * (HP / 8) * 128 + c.row) * 8;
    float dh5[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) dh5[q] = S.dhead[c.row][q];
    if (p.decouple) dh5[4] = 0.f;                      // value head sees x.detach() (game.py:1208)
    float* lnp = p.ln_part + (((size_t(blockIdx.x) * 4 + (c.row >> 5)) * (L + 1) + l) * 2) * HP;   // [dgamma | dbeta]
    float (*scr)[32] = S.scratch[c.warp];
    float s1 = 0.f, s2 = 0.f;
    // z_l of the next column group is fetched (scratch, L2) while this one is processed: the load latency was the
    // single largest stall of the kernel (ncu source view: 5.5 % of all samples on the first use of z)
    float zn[8];
    if (!first) ld256(zrow + size_t(c.c0) * 128, zn);
#pragma unroll 1
    for (int g = 0; g < c.ng; ++g) {
        const int col = c.c0 + 8 * g;
        float z[8], dh[8], gx[8], gg[8];
        if (first) {
            tc::tmem_ld8(c.tD + uint32_t(8 * g), z);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = 0.f;
#pragma unroll
                for (int q = 0; q < 5; ++q) a = fmaf(dh5[q], S.headw[q][col + j], a);
                dh[j] = a;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) z[j] = zn[j];
            if (g + 1 < c.ng) ld256(zrow + size_t(col + 8) * 128, zn);
            float d[8];
            tc::tmem_ld8x2(c.tX + uint32_t(8 * g), dh, c.tD + uint32_t(8 * g), d);
#pragma unroll
            for (int j = 0; j < 8; ++j) dh[j] += d[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xh = fmaf(z[j], rstd, shift);
            const float y = fmaf(xh, gam[8 * g + j], bet[8 * g + j]);
            const float gj = DROP ? ((y > 0.f && ((keep_bits >> (8 * g + j)) & 1ull)) ? dh[j] * dscale : 0.f) : (y > 0.f ? dh[j] : 0.f);
            const float t = gj * gam[8 * g + j];
            s1 += t;
            s2 = fmaf(t, xh, s2);
            z[j] = xh;
            gg[j] = gj;
            gx[j] = gj * xh;
        }
        tc::tmem_st8(c.tD + uint32_t(8 * g), z);       // xhat
        tc::tmem_st8(c.tX + uint32_t(8 * g), dh);      // dh_l
        colsum4x2(scr, c.lane, gx, gg, lnp + col, lnp + HP + col);
        colsum4x2(scr, c.lane, gx + 4, gg + 4, lnp + col + 4, lnp + HP + col + 4);
    }
    tc::tmem_st_wait();
    if (first && threadIdx.x == 0) tc::bulk_wait_read0();   // h_L has left the operand buffers (copy issued before pass A) before pass B writes dz_L there
    const float m1 = exchange(S, 0, c.part, c.row, s1) * inv_h;
    const float m2 = exchange(S, 1, c.part, c.row, s2) * inv_h;
#pragma unroll 1
    for (int g = 0; g < c.ng; ++g) {
        const int col = c.c0 + 8 * g;
        float xh[8], dh[8], dz[8];
        tc::tmem_ld8x2(c.tD + uint32_t(8 * g), xh, c.tX + uint32_t(8 * g), dh);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float y = fmaf(xh[j], gam[8 * g + j], bet[8 * g + j]);
            const float t = (DROP ? ((y > 0.f && ((keep_bits >> (8 * g + j)) & 1ull)) ? dh[j] * dscale : 0.f) : (y > 0.f ? dh[j] : 0.f)) * gam[8 * g + j];
            dz[j] = (col + j < h) ? rstd * (t - m1 - xh[j] * m2) : 0.f;
        }
        uint4 hi4, lo4;
        split8(dz, hi4, lo4);
        if (!c.valid) hi4 = lo4 = make_uint4(0u, 0u, 0u, 0u);
        store_operand(S, c.row, col, hi4, lo4);                    // next MMA operand (l > 0) and the image of dz_l
    }
}

template <bool DROP>
__global__ void __launch_bounds__(THREADS, 1) update_mlp_kernel(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (tc::smem_addr(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int HP = p.HP, L = p.L, KB = HP / 16;
    const int64_t ntiles = p.ntiles;
    const int my_tiles = ntiles > blockIdx.x ? int((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    const int nblk = blocks_per_tile(HP, L, p.backward != 0);
    const uint32_t part_bytes = uint32_t(HP) * 32u;

    if (warp == 0) tc::tmem_alloc(&S.tmem_base, 512);
    if (tid == 0) {
        tc::mbar_init(&S.a_ready, ROW_THREADS);
        tc::mbar_init(&S.mma_done, 1);
        for (int i = 0; i < RING; ++i) {
            tc::mbar_init(&S.b_full[i], 1);
            tc::mbar_init(&S.b_empty[i], 1);
        }
        tc::mbar_fence_init();
    }
    for (int i = tid; i < HP; i += THREADS) {
        S.b0[i] = p.pf[pf_b0(HP) + i];
        for (int l = 0; l <= L; ++l) {
            S.gamma[l][i] = p.pf[pf_gamma(HP, l) + i];
            S.beta[l][i] = p.pf[pf_beta(HP, L, l) + i];
        }
        for (int q = 0; q < 5; ++q) S.headw[q][i] = p.pf[pf_headw(HP, L) + int64_t(q) * HP + i];
    }
    if (tid < 8) S.headb[tid] = p.pf[pf_headb(HP, L) + tid];
    for (uint32_t i = tid * 16; i < 2 * A_PART; i += THREADS * 16) *reinterpret_cast<uint4*>(&S.A[0][0] + i) = make_uint4(0, 0, 0, 0);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = S.tmem_base;

    {
        // ---------------- row threads
        RowCtx c;
        c.warp = warp;
        c.lane = lane;
        c.part = warp >> 2;
        c.row = (warp & 3) * 32 + lane;
        const int G = HP / 8, g0 = (G * c.part) / SPLIT, g1 = (G * (c.part + 1)) / SPLIT;
        c.c0 = 8 * g0;
        c.ng = g1 - g0;
        c.tD = tmem_base + (uint32_t((warp & 3) * 32) << 16) + uint32_t(c.c0);
        c.tX = c.tD + X_COL;
        const bool issuer = tid == 0, producer = tid == 32;
        // weight producer state: block pq of the launch-wide sequence goes to slot pq % RING
        const uint32_t total_blocks = uint32_t(my_tiles) * uint32_t(nblk);
        uint32_t pq = 0, pb = 0;
        auto refill = [&]() {        // non-blocking: issue every copy whose slot is free
            while (pq < total_blocks) {
                const uint32_t slot = pq % RING;
                if (pq >= RING && !tc::mbar_test(&S.b_empty[slot], ((pq / RING) & 1u) ^ 1u)) break;
                // position pb of the tile's sequence -> block of the image
                const uint32_t nfwd = 1u + 2u * uint32_t(L * KB);
                uint32_t off, bytes;
                if (pb < nfwd) {
                    const uint32_t t2 = pb - 1u, id = pb == 0 ? 0u : 1u + (t2 / uint32_t(2 * KB)) * uint32_t(KB) + t2 % uint32_t(KB);
                    off = id * 3u * part_bytes;
                    bytes = 3u * part_bytes;
                } else {
                    off = uint32_t(fwd_blocks(HP, L)) * 3u * part_bytes + (pb - nfwd) * 2u * part_bytes;
                    bytes = 2u * part_bytes;
                }
                tc::mbar_expect_tx(&S.b_full[slot], bytes);
                tc::bulk_g2s(S.B[slot], p.img + off, bytes, &S.b_full[slot]);
                ++pq;
                if (++pb == uint32_t(nblk)) pb = 0;
            }
        };
        if (producer) refill();
        const uint32_t idesc = tc::make_idesc_bf16_major(128, HP, false, false);
        const uint64_t dA0 = tc::make_desc_sw32(tc::smem_addr(S.A[0]), 16, 256), dA1 = tc::make_desc_sw32(tc::smem_addr(S.A[1]), 16, 256),
                       dB0 = tc::make_desc_sw32(tc::smem_addr(S.B[0]), 16, 256);
        uint32_t stage = 0, bq = 0;                 // running MMA-stage and weight-block counters (issuer)
        double lacc[4] = {0.0, 0.0, 0.0, 0.0};
        float hb[5] = {0.f, 0.f, 0.f, 0.f, 0.f};

        // One MMA stage over `kblocks` k-blocks of A against the next weight blocks of the ring.
        //   ST_STEM: A exact in bf16 (hi only)            D  = hi*r + hi*lo + hi*hi
        //   ST_FWD1: A = (hi, lo)                          D  = lo*lo + lo*hi + hi*lo + hi*hi
        //   ST_FWD2: the lo buffer now holds r             D += r*hi + hi*r
        //   ST_BWD : A = (hi, lo), 2-part weight blocks    D  = lo*hi + hi*lo + hi*hi
        enum { ST_STEM, ST_FWD1, ST_FWD2, ST_BWD };
        // `img`: where the operand tile of this stage goes as a tensor image (h_{l-1} / dz_l), or nullptr
        auto run_stage = [&](int kblocks, int kind, uint8_t* img) {
            tc::fence_async_smem();
            tc::fence_before_sync();
            tc::mbar_arrive(&S.a_ready);
            if (issuer) {
                tc::mbar_wait(&S.a_ready, stage & 1u);
                tc::fence_after_sync();
                if (img) copy_out_tile(S, img, HP);
                for (int j = 0; j < kblocks; ++j, ++bq) {
                    const uint32_t slot = bq % RING;
                    tc::mbar_wait(&S.b_full[slot], (bq / RING) & 1u);
                    tc::fence_after_sync();
                    // descriptors = one base per buffer + a byte offset (>> 4) in the address field
                    const uint64_t bh = dB0 + uint64_t((slot * B_SLOT) >> 4), bl = bh + uint64_t(part_bytes >> 4),
                                   br = bh + uint64_t((2u * part_bytes) >> 4);
                    const uint64_t ah = dA0 + uint64_t(uint32_t(j) * (4096u >> 4)), al = dA1 + uint64_t(uint32_t(j) * (4096u >> 4));
                    if (kind == ST_STEM) {
                        tc::mma_bf16_ss(tmem_base, ah, br, idesc, j > 0);
                        tc::mma_bf16_ss(tmem_base, ah, bl, idesc, true);
                        tc::mma_bf16_ss(tmem_base, ah, bh, idesc, true);
                    } else if (kind == ST_FWD1) {
                        tc::mma_bf16_ss(tmem_base, al, bl, idesc, j > 0);
                        tc::mma_bf16_ss(tmem_base, al, bh, idesc, true);
                        tc::mma_bf16_ss(tmem_base, ah, bl, idesc, true);
                        tc::mma_bf16_ss(tmem_base, ah, bh, idesc, true);
                    } else if (kind == ST_FWD2) {
                        tc::mma_bf16_ss(tmem_base, al, bh, idesc, true);
                        tc::mma_bf16_ss(tmem_base, ah, br, idesc, true);
                    } else {
                        tc::mma_bf16_ss(tmem_base, al, bh, idesc, j > 0);
                        tc::mma_bf16_ss(tmem_base, ah, bl, idesc, true);
                        tc::mma_bf16_ss(tmem_base, ah, bh, idesc, true);
                    }
                    tc::mma_commit(&S.b_empty[slot]);
                }
                if (img) tc::bulk_wait_read0();      // mma_done (below) releases the operand buffers to the row threads
                tc::mma_commit(&S.mma_done);
            }
            if (producer) {
                while (!tc::mbar_test(&S.mma_done, stage & 1u)) refill();
                refill();            // every slot of this stage is free now: prefetch the next stage's first blocks
            }
            tc::mbar_wait(&S.mma_done, stage & 1u);
            tc::fence_after_sync();
            ++stage;
        };
        // between the two forward phases: r = x - hi - lo of this thread's columns, x = the residual stream in TMEM
        auto write_residual_terms = [&]() {
#pragma unroll 1
            for (int g = 0; g < c.ng; ++g) {
                float x[8];
                tc::tmem_ld8(c.tX + uint32_t(8 * g), x);
                store_operand_r(S, c.row, c.c0 + 8 * g, x);
            }
        };

        for (int t = 0; t < my_tiles; ++t) {
            c.tile = int64_t(blockIdx.x) + int64_t(t) * gridDim.x;
            c.grow = c.tile * 128 + c.row;
            c.valid = c.grow < p.n;
            // model input: the 16 exponents are exact in bf16 (k-block 0, hi part; row / column features
            // are folded into b0)
            const uint64_t b_in = (c.part == 0 && c.valid) ? p.boards[c.grow] : 0ull;   // in flight during the wait below
            if (p.backward && t > 0) {                 // the previous tile's dz_0 has left the operand buffers
                if (issuer) tc::bulk_wait_read0();
                row_sync();
            }
            if (c.part == 0) {
                const uint64_t b = b_in;
                float e[8];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) e[j] = float((b >> (4 * (8 * u + j))) & 15ull);
                    store_operand(S, c.row, 8 * u, e);
                }
            }
            float o[5];
            auto image = [&](float* base, int l) -> uint8_t* {      // image of tensor l of this tile: 128 samples x HP x 4 B
                return p.backward ? reinterpret_cast<uint8_t*>(base) + (size_t(l) * p.ntiles + c.tile) * size_t(HP) * 512 : nullptr;
            };
            uint64_t keep_bits[MAXL + 1];                           // dropout masks of this thread's columns, per block
#pragma unroll
            for (int l = 0; l <= MAXL; ++l) keep_bits[l] = 0ull;
            run_stage(1, ST_STEM, nullptr);
            fwd_epilogue<true, DROP>(S, p, c, 0, o, keep_bits[0]);
#pragma unroll
            for (int l = 1; l <= MAXL; ++l) {
                if (l > L) break;
                run_stage(KB, ST_FWD1, image(p.h_out, l - 1));     // the operand tile is h_{l-1}
                write_residual_terms();
                run_stage(KB, ST_FWD2, nullptr);
                fwd_epilogue<false, DROP>(S, p, c, l, o, keep_bits[l]);
            }
            if (issuer && p.backward) copy_out_tile(S, image(p.h_out, L), HP);   // h_L: ordered by the barriers of the last epilogue
            // ---- heads -> loss terms and their gradients (one thread per row)
            if (c.part == 0) {
                float gl[4] = {0.f, 0.f, 0.f, 0.f}, dv = 0.f;
                if (c.valid && p.logits) *reinterpret_cast<float4*>(p.logits + c.grow * 4) = make_float4(o[0], o[1], o[2], o[3]);
                if (c.valid && p.value) p.value[c.grow] = o[4];
                if (p.backward && c.valid && (!p.flags || (p.flags[c.grow] & ROLL_VALID))) {
                    const float l4[4] = {o[0], o[1], o[2], o[3]};
                    const uint32_t a = p.actions[c.grow] & 3u;
                    const float lp_old = p.old_logp[c.grow * p.old_stride + (p.old_stride == 4 ? a : 0)];
                    float ppo, vl, H, dvl;
                    ppo_sample(l4, p.legal[c.grow] & 15u, a, lp_old, p.adv[c.grow], o[4], p.g_norm[c.grow], p.clip_eps,
                               p.beta_ent, ppo, vl, H, gl, dvl);
#pragma unroll
                    for (int k = 0; k < 4; ++k) gl[k] *= -p.inv_n;
                    dv = p.inv_n * p.c_v * dvl;
                    lacc[0] += double(ppo);
                    lacc[1] += double(vl);
                    lacc[2] += double(H);
                    lacc[3] += 1.0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) hb[k] += gl[k];
                    hb[4] += dv;
                }
                *reinterpret_cast<float4*>(&S.dhead[c.row][0]) = make_float4(gl[0], gl[1], gl[2], gl[3]);
                *reinterpret_cast<float4*>(&S.dhead[c.row][4]) = make_float4(dv, 0.f, 0.f, 0.f);
                if (p.backward && c.valid) {
                    *reinterpret_cast<float4*>(p.dhead + c.grow * 8) = make_float4(gl[0], gl[1], gl[2], gl[3]);
                    *reinterpret_cast<float4*>(p.dhead + c.grow * 8 + 4) = make_float4(dv, 0.f, 0.f, 0.f);
                }
            }
            if (!p.backward) continue;
            row_sync();
            // ---- backward-data
            bwd_epilogue<DROP>(S, p, c, L, true, L == 2 ? keep_bits[2] : keep_bits[1]);
#pragma unroll
            for (int l = MAXL; l >= 1; --l) {
                if (l > L) continue;
                run_stage(KB, ST_BWD, image(p.dz_out, l));          // D = dz_l W_l; the operand tile is dz_l
                bwd_epilogue<DROP>(S, p, c, l - 1, false, keep_bits[l - 1]);
            }
            // dz_0 has no MMA after it: copy it out between two barriers before the next tile's input overwrites block 0
            tc::fence_async_smem();
            row_sync();
            if (issuer) copy_out_tile(S, image(p.dz_out, 0), HP);   // waited for at the top of the next tile
        }
        if (issuer) tc::bulk_wait0();                  // every image is complete in HBM before the kernel ends
        // ---- per-CTA loss sums and head-bias gradients, fixed order
        if (p.backward) {
            row_sync();
            if (c.part == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) S.lsum[c.row][k] = lacc[k];
#pragma unroll
                for (int k = 0; k < 5; ++k) S.dhead[c.row][k] = hb[k];
            }
            row_sync();
            if (tid < 4) {
                double s = 0.0;
                for (int r = 0; r < 128; ++r) s += S.lsum[r][tid];
                p.loss_part[size_t(blockIdx.x) * 4 + tid] = s;
            } else if (tid >= 32 && tid < 40) {
                const int k = tid - 32;
                float s = 0.f;
                if (k < 5)
                    for (int r = 0; r < 128; ++r) s += S.dhead[r][k];
                p.head_part[size_t(blockIdx.x) * 8 + k] = s;
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

// fixed-order reduction of the per-CTA partials: LN grads [L+1][2][h] (dgamma | dbeta), head biases [5], loss sums [4]
__global__ void update_reduce_kernel(const float* __restrict__ ln_part, const float* __restrict__ head_part,
                                     const double* __restrict__ loss_part, int parts, int L, int HP, int h,
                                     float* __restrict__ ln_grad, float* __restrict__ head_bias_grad, double* __restrict__ stats) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_ln = (L + 1) * 2 * h;
    if (idx < n_ln) {
        const int col = idx % h, lt = idx / h;          // lt = l*2 + tensor
        float s = 0.f;
        for (int c = 0; c < parts * 4; ++c) s += ln_part[(size_t(c) * (L + 1) * 2 + lt) * HP + col];
        ln_grad[idx] = s;
    } else if (idx < n_ln + 5) {
        const int k = idx - n_ln;
        float s = 0.f;
        for (int c = 0; c < parts; ++c) s += head_part[size_t(c) * 8 + k];
        head_bias_grad[k] = s;
    } else if (idx < n_ln + 9) {
        const int k = idx - n_ln - 5;
        double s = 0.0;
        for (int c = 0; c < parts; ++c) s += loss_part[size_t(c) * 4 + k];
        stats[k] = s;
    }
}

struct PackSrc {
    const float* stem_w;          // [h, 48]
    const float* w[MAXL];         // [h, h]
    const float* gamma[MAXL + 1];
    const float* beta[MAXL + 1];
    const float *action_w, *action_b, *value_w, *value_b;
};

__global__ void update_pack_kernel(PackSrc s, int h, int HP, int L, float* __restrict__ pf, uint8_t* __restrict__ img) {
    const int KB = HP / 16;
    const int nfwd = fwd_blocks(HP, L), nblk = nfwd + L * KB;
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    // ---- fp32 section
    if (idx < pf_floats(HP, L)) {
        const int i = int(idx);
        float v = 0.f;
        const int rowi = i / HP, col = i % HP;
        if (i < pf_headb(HP, L)) {
            if (col < h) {
                if (rowi == 0) {
                    // b0 = sum_cell W0[:, 3c+1] * (c/4)/3 + W0[:, 3c+2] * (c%4)/3  (game.py:92-101 features)
                    for (int cell = 0; cell < 16; ++cell)
                        v += s.stem_w[col * 48 + 3 * cell + 1] * pos_feature(cell >> 2) + s.stem_w[col * 48 + 3 * cell + 2] * pos_feature(cell & 3);
                } else if (rowi < 2 + L) v = s.gamma[rowi - 1][col];
                else if (rowi < 3 + 2 * L) v = s.beta[rowi - 2 - L][col];
                else {
                    const int q = rowi - 3 - 2 * L;
                    v = q < 4 ? s.action_w[q * h + col] : s.value_w[col];
                }
            }
        } else {
            const int q = i - int(pf_headb(HP, L));
            v = q < 4 ? s.action_b[q] : q == 4 ? s.value_b[0] : 0.f;
        }
        pf[i] = v;
    }
    // ---- weight k-blocks: block 0 = stem exponent columns, then W_1..W_L (hi | lo | r), then W_L^T..W_1^T (hi | lo)
    if (idx < int64_t(nblk) * HP * 16) {
        const int b = int(idx / (HP * 16)), rem = int(idx % (HP * 16)), n = rem / 16, kk = rem % 16;
        float v = 0.f;
        if (b == 0) {
            if (n < h) v = s.stem_w[n * 48 + 3 * kk];
        } else if (b < nfwd) {
            const int l = (b - 1) / KB, k = ((b - 1) % KB) * 16 + kk;
            if (n < h && k < h) v = s.w[l][size_t(n) * h + k];
        } else {
            const int t = b - nfwd, l = L - 1 - t / KB, k = (t % KB) * 16 + kk;
            if (n < h && k < h) v = s.w[l][size_t(k) * h + n];
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const float e = v - __bfloat162float(hi);
        const __nv_bfloat16 lo = __float2bfloat16_rn(e);
        const size_t part = size_t(HP) * 32;
        uint8_t* blk = b < nfwd ? img + size_t(b) * 3 * part : img + size_t(nfwd) * 3 * part + size_t(b - nfwd) * 2 * part;
        *reinterpret_cast<__nv_bfloat16*>(blk + tc::sw32_offset(n, kk)) = hi;
        *reinterpret_cast<__nv_bfloat16*>(blk + part + tc::sw32_offset(n, kk)) = lo;
        if (b < nfwd) *reinterpret_cast<__nv_bfloat16*>(blk + 2 * part + tc::sw32_offset(n, kk)) = __float2bfloat16_rn(e - __bfloat162float(lo));
    }
}

static int padded(int h) { return (h + 15) / 16 * 16; }
static bool shape_ok(int h, int L) { return h >= 16 && h <= MAXH && h % 4 == 0 && L >= 1 && L <= MAXL; }

}  // namespace uf
}  // namespace g2048

using namespace g2048;
using namespace g2048::uf;

extern "C" {

int64_t g2048_update_mlp_pack_bytes(int32_t hidden, int32_t layers) {
    if (!shape_ok(hidden, layers)) return -1;
    return pack_bytes(padded(hidden), layers);
}

int g2048_update_mlp_pack(int32_t hidden, int32_t layers, const float* stem_w, const float* stem_ln_w, const float* stem_ln_b,
                          const float* const* block_w, const float* const* block_ln_w, const float* const* block_ln_b,
                          const float* action_w, const float* action_b, const float* value_w, const float* value_b,
                          void* packed, void* stream) {
    if (!shape_ok(hidden, layers)) return fail(G2048_ESHAPE, "g2048_update_mlp_pack: hidden=%d (16..208, %%4), layers=%d (1..2) unsupported", hidden, layers);
    G2048_REQUIRE(stem_w && stem_ln_w && stem_ln_b && block_w && block_ln_w && block_ln_b && action_w && action_b && value_w &&
                      value_b && packed, "g2048_update_mlp_pack: NULL pointer argument");
    PackSrc s{};
    s.stem_w = stem_w;
    s.gamma[0] = stem_ln_w;
    s.beta[0] = stem_ln_b;
    for (int l = 0; l < layers; ++l) {
        G2048_REQUIRE(block_w[l] && block_ln_w[l] && block_ln_b[l], "g2048_update_mlp_pack: NULL block pointer");
        s.w[l] = block_w[l];
        s.gamma[l + 1] = block_ln_w[l];
        s.beta[l + 1] = block_ln_b[l];
    }
    s.action_w = action_w; s.action_b = action_b; s.value_w = value_w; s.value_b = value_b;
    const int HP = padded(hidden);
    const int64_t work = int64_t(fwd_blocks(HP, layers) + layers * (HP / 16)) * HP * 16;
    uint8_t* base = static_cast<uint8_t*>(packed);
    update_pack_kernel<<<unsigned((work + 255) / 256), 256, 0, cudaStream_t(stream)>>>(
        s, hidden, HP, layers, reinterpret_cast<float*>(base), base + img_offset_bytes(HP, layers));
    G2048_CHECK_LAUNCH("update_pack_kernel");
    return G2048_OK;
}

int64_t g2048_update_mlp_workspace_bytes(int32_t hidden, int32_t layers) {
    if (!shape_ok(hidden, layers)) return -1;
    const int64_t HP = padded(hidden), g = num_sms();
    return g * layers * 128 * HP * 4 + g * 4 * (layers + 1) * 2 * HP * 4 + g * 8 * 4 + g * 4 * 8 + 1024;
}

int g2048_update_mlp_fwd_bwd(const G2048UpdateMlp* u, void* stream) {
    G2048_REQUIRE(u != nullptr, "g2048_update_mlp_fwd_bwd: params is NULL");
    G2048_REQUIRE(u->n >= 0, "g2048_update_mlp_fwd_bwd: n < 0");
    if (!shape_ok(u->hidden, u->layers)) return fail(G2048_ESHAPE, "g2048_update_mlp_fwd_bwd: hidden=%d, layers=%d unsupported", u->hidden, u->layers);
    G2048_REQUIRE(u->dropout_p >= 0.f && u->dropout_p < 1.f, "g2048_update_mlp_fwd_bwd: dropout_p must be in [0, 1)");
    const int L = u->layers, h = u->hidden, HP = padded(h);
    cudaStream_t st = cudaStream_t(stream);
    const bool bw = u->backward != 0;
    if (bw) G2048_REQUIRE(u->ln_grad && u->head_bias_grad && u->stats, "g2048_update_mlp_fwd_bwd: NULL gradient output");
    if (u->n == 0) {
        if (bw) {
            G2048_CHECK_CUDA(cudaMemsetAsync(u->ln_grad, 0, size_t(L + 1) * 2 * h * 4, st));
            G2048_CHECK_CUDA(cudaMemsetAsync(u->head_bias_grad, 0, 5 * 4, st));
            G2048_CHECK_CUDA(cudaMemsetAsync(u->stats, 0, 4 * 8, st));
        }
        return G2048_OK;
    }
    G2048_REQUIRE(u->boards && u->packed && u->workspace, "g2048_update_mlp_fwd_bwd: NULL pointer argument");
    if (bw) {
        G2048_REQUIRE(u->actions && u->legal && u->old_logp && u->adv && u->g_norm && u->h_out && u->dz_out && u->dhead,
                      "g2048_update_mlp_fwd_bwd: NULL pointer argument (backward)");
        G2048_REQUIRE(u->old_logp_stride == 1 || u->old_logp_stride == 4, "g2048_update_mlp_fwd_bwd: old_logp_stride must be 1 or 4");
    } else {
        G2048_REQUIRE(u->logits || u->value, "g2048_update_mlp_fwd_bwd: forward-only call without outputs");
    }
    const int64_t ntiles = (u->n + 127) / 128;
    const int grid = int(ntiles < num_sms() ? ntiles : num_sms());
    Params p{};
    p.n = u->n; p.ntiles = ntiles; p.h = h; p.HP = HP; p.L = L; p.decouple = u->decouple_critic; p.backward = bw;
    p.boards = u->boards; p.actions = u->actions; p.legal = u->legal; p.flags = u->flags;
    p.old_logp = u->old_logp; p.old_stride = u->old_logp_stride; p.adv = u->adv; p.g_norm = u->g_norm;
    p.clip_eps = u->clip_eps; p.c_v = u->critic_strength; p.beta_ent = u->entropy_strength; p.inv_n = u->inv_n;
    if (u->dropout_p > 0.f) {
        p.drop_thr = uint32_t(u->dropout_p * 65536.0f + 0.5f);
        p.drop_scale = 1.0f / (1.0f - u->dropout_p);
        p.drop_seed = u->dropout_seed;
        p.sample0 = u->dropout_sample0;
    }
    const uint8_t* base = static_cast<const uint8_t*>(u->packed);
    p.pf = reinterpret_cast<const float*>(base);
    p.img = base + img_offset_bytes(HP, L);
    p.h_out = u->h_out; p.dz_out = u->dz_out; p.dhead = u->dhead; p.logits = u->logits; p.value = u->value;
    uint8_t* ws = static_cast<uint8_t*>(u->workspace);
    const size_t z_bytes = size_t(num_sms()) * L * 128 * HP * 4, ln_bytes = size_t(num_sms()) * 4 * (L + 1) * 2 * HP * 4;
    p.zscratch = reinterpret_cast<float*>(ws);
    p.ln_part = reinterpret_cast<float*>(ws + z_bytes);
    p.head_part = reinterpret_cast<float*>(ws + z_bytes + ln_bytes);
    p.loss_part = reinterpret_cast<double*>(ws + z_bytes + ln_bytes + size_t(num_sms()) * 8 * 4);
    if (bw) G2048_CHECK_CUDA(cudaMemsetAsync(p.ln_part, 0, ln_bytes, st));
    const int smem = int(sizeof(Smem)) + 1024;
    if (p.drop_thr) {
        G2048_CHECK_CUDA(ensure_smem(update_mlp_kernel<true>, smem));
        update_mlp_kernel<true><<<grid, THREADS, smem, st>>>(p);
    } else {
        G2048_CHECK_CUDA(ensure_smem(update_mlp_kernel<false>, smem));
        update_mlp_kernel<false><<<grid, THREADS, smem, st>>>(p);
    }
    G2048_CHECK_LAUNCH("update_mlp_kernel");
    if (bw) {
        const int work = (L + 1) * 2 * h + 9;
        update_reduce_kernel<<<(work + 127) / 128, 128, 0, st>>>(p.ln_part, p.head_part, p.loss_part, grid, L, HP, h, u->ln_grad,
                                                                u->head_bias_grad, u->stats);
        G2048_CHECK_LAUNCH("update_reduce_kernel");
    }
    return G2048_OK;
}

}  // extern "C"
