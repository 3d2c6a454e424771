// g2048_update_x3.cu -- the policy update's forward, loss and backward-data pass of GameMLP as ONE persistent, pipelined
// tcgen05 kernel (train.py:491-556: model(x), the loss section, loss.backward()).
//
// A CTA owns tiles of 128 samples for the whole chain; sample m of the tile is TMEM lane m:
//
//   boards -> x (16 exponents)                                             game.py:92-101
//   z0 = x W0e^T + b0 ; h0 = relu(LN0(z0))                                 game.py:1069-1073
//   z_l = h_{l-1} W_l^T ; h_l = h_{l-1} + Dropout(relu(LN_l(z_l))),  l = 1..L   game.py:1038-1046
//   logits, V = heads(h_L) ; per-sample PPO-clip / critic / entropy terms  train.py:497-554
//   dh_L = heads^T (dlogits, dV) ; for l = L..0:  g = dh_l * [LN_l(z_l) > 0] * keep_l / (1-p), dz_l = LN_l-backward(g),
//   dh_{l-1} = dh_l + dz_l W_l                                             (what autograd would run)
//
// Every GEMM operand is TWO fp16 terms, x = hi + lo (22 mantissa bits), and every k-step is three tcgen05.mma products
// lo*hi + hi*lo + hi*hi accumulated in fp32 in tensor memory: fp32-grade (~2e-7 of scale) in the forward, which decides
// every ReLU branch and the loss, AND in the backward.  fp16 terms need O(1) magnitudes: activations are, and the
// gradient chain is kept there by the caller's power-of-two loss scale (inv_n arrives pre-multiplied; every gradient this
// kernel and the weight-gradient GEMMs emit carries the same factor).  Round 1 used bf16 terms: three of them and six
// products in two MMA phases for the forward (24 bits), two and three products for the backward (16 bits).
//
// Pipeline inside a tile (the structure of g2048_rollout_x3.cu): pass 1 of an epilogue pulls the thread's accumulator
// columns into registers, which frees the accumulator D; pass 2 walks the 16-column k-blocks in rounds, writes the fp32
// residual stream (TMEM X) and the hi | lo operand bytes of the next GEMM and signals each round; an issuer warp starts the
// next GEMM's MMAs on the blocks of a round as soon as it is signalled, so the tensor pipe runs under pass 2; a producer
// warp streams the weight k-blocks (hi | lo, HP x 64 B, forward blocks then transposed blocks in the order a tile consumes
// them) through a ring of bulk async copies; a storer warp bulk-copies every finished operand tile to HBM -- the tiles ARE
// the h_l / dz_l tensors the weight-gradient GEMMs read (fp16 hi|lo images, 4 bytes per value).
//
// What leaves the SM per sample: h_0..h_L and dz_0..dz_L (operand images), the 5 head gradients, and nothing else: z_l,
// LN statistics, logits, per-sample loss terms never touch HBM (z_0..z_{L-1} round-trip through a per-CTA scratch that
// stays in L2).  LayerNorm-parameter gradients are column sums over samples: reduced per warp through shared memory, added
// to per-(CTA, lane-quarter) fp32 partials with single-writer red.global (deterministic), summed in fixed order afterwards.
#include <cfloat>
#include <cuda_fp16.h>
#include "g2048_device.cuh"
#include "g2048_host.h"
#include "g2048_loss.cuh"
#include "g2048_tc.cuh"

namespace g2048 {
namespace ux {

constexpr int MAXL = 2;
constexpr int SPLIT = 4;
constexpr int ROW_THREADS = 128 * SPLIT;        // 16 warps: warp w -> lane quarter (w & 3), column part (w >> 2)
constexpr int THREADS = ROW_THREADS + 96;       // + issuer, producer, storer warps (5 warps per scheduler: 96 registers each)
constexpr int ISSUER_WARP = 16, PRODUCER_WARP = 17, STORER_WARP = 18;
constexpr uint32_t X_COL = 256;
constexpr uint32_t ROLL_VALID = 0x80u;

__host__ __device__ inline int padded(int h) { return h <= 64 ? 64 : h <= 128 ? 128 : h <= 192 ? 192 : 208; }

struct Params {
    int64_t n, ntiles;
    int h, L, decouple;
    int backward;                 // 0: forward only (logits / value out), 1: forward + loss + backward-data
    const uint64_t* boards;
    const uint8_t *actions, *legal, *flags;
    const float* old_logp;
    int old_stride;
    const float *adv, *g_norm;
    float clip_eps, c_v, beta_ent, inv_n;
    uint32_t drop_thr;            // dropout (game.py:1042): an element is dropped iff its 16-bit Philox draw < drop_thr; 0 = off
    float drop_scale;             // 1 / (1 - p)
    uint64_t sample0;             // index of this call's first sample in the mask's counter space
    PhiloxKeys drop_keys;         // the ten round keys of the call's dropout seed (constant-bank operands of the rounds)
    const float* pf;              // fp32 section of the pack
    const uint8_t* img;           // weight k-blocks in consumption order
    uint8_t* h_out;               // [L+1][ntiles] fp16 hi|lo operand images of 128 samples x HP, 4 B / value
    uint8_t* dz_out;              // same layout
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
// weight image: k-blocks of HP rows x (32 B hi | 32 B lo): the stem's exponent columns, W_1..W_L, then W_L^T..W_1^T
__host__ __device__ inline int img_blocks(int HP, int L) { return 1 + 2 * L * (HP / 16); }
__host__ __device__ inline int blocks_per_tile(int HP, int L, bool backward) { return 1 + (backward ? 2 : 1) * L * (HP / 16); }
__host__ __device__ inline int64_t pack_bytes(int HP, int L) { return img_offset_bytes(HP, L) + int64_t(img_blocks(HP, L)) * HP * 64; }

constexpr int ring_slots(int HP) {
    const int fixed = 2 * (HP / 16) * 4096 + (1 + 2 * (MAXL + 1) + 5) * HP * 4 + 32 + 2 * SPLIT * 128 * 4 + (MAXL + 1) * 2 * 128 * 4 +
                      128 * 8 * 4 + (ROW_THREADS / 32) * 8 * 32 * 4 + 128 * 4 * 8 + 1024;
    const int n = (232448 - 1024 - fixed) / (HP * 64);
    return n > 8 ? 8 : n;
}

template <int HP>
struct Smem {
    static constexpr int NB = HP / 16;            // k-blocks
    static constexpr int NBF = NB / 4;            // full blocks per column part (block 4i + part in round i)
    static constexpr int NR = NB % 4;             // remainder blocks, split in 4-column units over the parts
    static constexpr int ROUNDS = NBF + NR;
    static constexpr uint32_t PART = NB * 4096u;  // one operand part: NB blocks of 128 rows x 32 B
    static constexpr uint32_t WPART = HP * 32u;   // one part of a weight k-block
    static constexpr uint32_t SLOT = 2u * WPART;
    static constexpr int RING = ring_slots(HP);
    alignas(1024) uint8_t A[2][PART];             // activations / gradients, fp16 hi | lo
    alignas(1024) uint8_t W[RING][SLOT];
    alignas(16) float b0[HP];
    alignas(16) float gamma[MAXL + 1][HP];
    alignas(16) float beta[MAXL + 1][HP];
    alignas(16) float headw[5][HP];
    alignas(16) float headb[8];
    float red[2][SPLIT][128];
    float stats[MAXL + 1][2][128];                // mean, rstd of every LayerNorm row
    alignas(16) float dhead[128][8];
    alignas(16) float scratch[ROW_THREADS / 32][8][32];   // per-warp transposition buffer (column sums, head partials)
    double lsum[128][4];
    uint64_t in_ready, mma_done, img_ready, a_free, rnd_ready[ROUNDS], w_full[RING], w_empty[RING];
    uint32_t tmem_base;
};
static_assert(sizeof(Smem<208>) + 1024 <= 232448, "update kernel exceeds the 227 KB shared memory limit");
static_assert(Smem<208>::RING >= 4, "weight ring too short to cover the L2 latency");

__device__ __forceinline__ void row_sync() { asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory"); }
__device__ __forceinline__ void red_add(float* addr, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// one warp's arrival on a barrier: every lane's operand bytes visible to the async proxy first
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    tc::fence_async_smem();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(bar);
}

// 32-byte (one sector) global accesses to the TILED z scratch, [column group of 8][row 0..127][8 floats]: the 32 rows of a
// warp touch 1 KiB contiguously
__device__ __forceinline__ void st256(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld256(const float* p, float* v) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p)
                 : "memory");
}

// Column sums over the 32 rows of a warp for 4 columns of two tensors; the 8 totals go to `dst_a[0..3]` and `dst_b[0..3]`
// with single-writer red.global (the same lane writes the same address every tile).
// (not inlined, like the Philox call of the dropout mask: the epilogues below are unrolled over a thread's 6-7 column units,
// and a kernel body of 230 KB spent 38 % of its stall samples waiting for instruction fetches -- ncu, profiles/)
__device__ __forceinline__ void colsum4x2(float (*scr)[32], int lane, float a0, float a1, float a2, float a3, float b0, float b1, float b2,
                                       float b3, float* dst_a, float* dst_b, float scale = 1.0f) {
    scr[0][lane] = a0; scr[1][lane] = a1; scr[2][lane] = a2; scr[3][lane] = a3;
    scr[4][lane] = b0; scr[5][lane] = b1; scr[6][lane] = b2; scr[7][lane] = b3;
    __syncwarp();
    const int v = lane >> 2, seg = lane & 3;
    const float4 p0 = *reinterpret_cast<const float4*>(&scr[v][seg * 8]), p1 = *reinterpret_cast<const float4*>(&scr[v][seg * 8 + 4]);
    float s = ((p0.x + p0.y) + (p0.z + p0.w)) + ((p1.x + p1.y) + (p1.z + p1.w));
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (seg == 0) red_add(v < 4 ? dst_a + v : dst_b + (v - 4), s * scale);
    __syncwarp();
}

// v if (bit `mask` of keep is set and y > 0) else 0 -- the ReLU gate and the dropout mask in three instructions (a LOP3 that
// sets a predicate, an FSETP that ANDs with it, an FSEL).  Written in PTX: the C form `(keep & (1u << j)) && y > 0.f` is
// canonicalised into shift / and / compare / select / compare / select, six instructions per element (SASS, profiles/).
__device__ __forceinline__ float gate_keep_pos(float y, float v, uint32_t keep, uint32_t mask) {
    float r;
    asm("{\n\t.reg .pred pk, pg;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %3, %4;\n\t"
        "setp.ne.u32 pk, t, 0;\n\t"
        "setp.gt.and.f32 pg, %1, 0f00000000, pk;\n\t"
        "selp.f32 %0, %2, 0f00000000, pg;\n\t}"
        : "=f"(r)
        : "f"(y), "f"(v), "r"(keep), "r"(mask));
    return r;
}

// Dropout mask of 8 consecutive columns of one sample in block l (game.py:1038-1046: x + Dropout(ReLU(LN(Linear x)))):
// one Philox4x32-10 call, counter = (sample index, l, column group), key = the call's dropout seed; column j of the
// group is KEPT iff the j-th 16-bit lane of the 128 random bits is >= drop_thr = round(p * 65536).  Bit j of the result.
__device__ __forceinline__ uint32_t dropout_keep8(const Params& p, int64_t sample, int l, int col) {
    const U4 r = philox4x32_10(uint32_t(sample), uint32_t(uint64_t(sample) >> 32), uint32_t(l), uint32_t(col >> 3),
                               p.drop_keys);
    const uint32_t t2 = p.drop_thr * 0x00010001u;
    const uint32_t a = __vsetgeu2(r.x, t2), b = __vsetgeu2(r.y, t2), c = __vsetgeu2(r.z, t2), d = __vsetgeu2(r.w, t2);   // bits 0, 16
    const uint32_t t = a + 4u * b + 16u * c + 64u * d;       // lanes 0 / 1 of a, b, c, d at bits 0, 2, 4, 6 / 16, 18, 20, 22 (three IMADs)
    return (t & 0x55u) | ((t >> 15) & 0xAAu);
}

// The masks of one thread's columns of block l (its NBF 16-column k-blocks and its quarter of the 4-column units), packed
// as the epilogues read them: bits 16 i + 8 u + j = column 16 (4 i + part) + 8 u + j, bits 16 NBF + 4 r + j = column
// 16 (4 NBF + r) + 4 part + j.  Called BEFORE the wait for the stage's MMAs, so that the Philox rounds run under them.
template <int NBF, int NR>
__device__ __forceinline__ uint64_t dropout_keep_row(const Params& p, int64_t sample, int l, int part) {
    uint64_t kb = 0ull;                 // (one call site per kernel; unrolled: the calls are independent dependency chains)
#pragma unroll
    for (int i = 0; i < 2 * NBF; ++i) kb |= uint64_t(dropout_keep8(p, sample, l, 16 * (4 * (i >> 1) + part) + 8 * (i & 1))) << (8 * i);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int col = 16 * (4 * NBF + r) + 4 * part;
        kb |= uint64_t((dropout_keep8(p, sample, l, col) >> (col & 7)) & 0xFu) << (16 * NBF + 4 * r);
    }
    return kb;
}

struct RowCtx {
    int row, part, lane, warp;
    uint32_t tD, tX;             // TMEM addresses of this thread's lane quarter: D and X, column 0
    uint32_t a_row;              // shared address of A[0] + row * 32
    uint32_t sw;                 // (row >> 2) & 1: the 16-byte halves of a 32-byte operand row are swapped
    int64_t grow;                // global sample index
    int64_t tile;
    bool valid;                  // grow < n
    uint32_t images;             // operand tiles handed to the storer so far (backward mode)
};

// the storer has copied the previous operand tile out of A: this pass may overwrite it
template <int HP>
__device__ __forceinline__ void wait_a_free(Smem<HP>& S, const Params& p, const RowCtx& c) {
    if (p.backward && c.images > 0) tc::mbar_wait(&S.a_free, (c.images - 1u) & 1u);
}

// 8 columns starting at `col0` (one 16-byte operand unit of block `blk`, unit index u): fp16 hi | lo operand bytes
template <int HP>
__device__ __forceinline__ void store_unit(const RowCtx& c, int blk, int u, const float* x, bool zero) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tc::split2_f16(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
    if (zero) {                    // rows past the sample count (only in the last tile of a launch) are zero in the images
#pragma unroll
        for (int j = 0; j < 4; ++j) hi[j] = lo[j] = 0u;
    }
    const uint32_t a = c.a_row + uint32_t(blk) * 4096u + ((uint32_t(u) ^ c.sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a + Smem<HP>::PART), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
}
// 4 columns 4 part .. 4 part + 3 of remainder block `blk`
template <int HP>
__device__ __forceinline__ void store_quad(const RowCtx& c, int blk, const float* x, bool zero) {
    uint32_t hi[2], lo[2];
    tc::split2_f16(x[0], x[1], hi[0], lo[0]);
    tc::split2_f16(x[2], x[3], hi[1], lo[1]);
    if (zero) hi[0] = hi[1] = lo[0] = lo[1] = 0u;
    const uint32_t a = c.a_row + uint32_t(blk) * 4096u + ((uint32_t(c.part >> 1) ^ c.sw) << 4) + uint32_t(c.part & 1) * 8u;
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(a), "r"(hi[0]), "r"(hi[1]) : "memory");
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(a + Smem<HP>::PART), "r"(lo[0]), "r"(lo[1]) : "memory");
}

// ---------------------------------------------------------------------------------------------------- forward epilogue
// LayerNorm l: h_l = [h_{l-1} +] Dropout(relu(LN(z_l))).  Pass 1: this thread's columns of D -> registers, statistics,
// z -> scratch (backward, l < L).  Pass 2: X (TMEM), the next A operand / the image of h_l, rounds; for l == L the 5 head dot
// products (complete on part 0).  Thread columns: blocks 4i + part (16 each) and 4 columns of every remainder block.
// One copy of the code serves the stem (l == 0: bias b0, no residual, no dropout) and the residual blocks.
template <int HP, bool DROP>
__device__ __forceinline__ void fwd_epilogue(Smem<HP>& S, const Params& p, RowCtx& c, int l, float (&o)[5], const uint64_t keep_bits) {
    using SM = Smem<HP>;
    constexpr int NBF = SM::NBF, NR = SM::NR;
    const int L = p.L;
    const bool last = l == L, STEM = l == 0;
    float z[NBF > 0 ? NBF : 1][16], zr[NR > 0 ? NR : 1][4];
    {
        uint32_t raw[NBF > 0 ? NBF : 1][16], rawr[NR > 0 ? NR : 1][4];
#pragma unroll
        for (int i = 0; i < NBF; ++i) tc::tmem_ld16_issue(c.tD + uint32_t(16 * (4 * i + c.part)), raw[i]);
#pragma unroll
        for (int r = 0; r < NR; ++r) tc::tmem_ld4_issue(c.tD + uint32_t(16 * (4 * NBF + r) + 4 * c.part), rawr[r]);
        tc::tmem_ld_wait_all();
#pragma unroll
        for (int i = 0; i < NBF; ++i)
#pragma unroll
            for (int j = 0; j < 16; ++j) z[i][j] = tc::tmem_ld_pin(raw[i][j]);
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) zr[r][j] = tc::tmem_ld_pin(rawr[r][j]);
    }
    // ---- pass 1: statistics in one pass (padded columns hold z = 0 and add nothing); z -> scratch for the backward pass
    float* zrow = (p.backward && !last) ? p.zscratch + ((size_t(blockIdx.x) * L + l) * (HP / 8) * 128 + c.row) * 8 : nullptr;
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < NBF; ++i) {
        const int col0 = 16 * (4 * i + c.part);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (STEM) z[i][j] += S.b0[col0 + j];
            sum += z[i][j];
            sq = fmaf(z[i][j], z[i][j], sq);
        }
        if (zrow) {
            st256(zrow + size_t(col0) * 128, &z[i][0]);
            st256(zrow + size_t(col0 + 8) * 128, &z[i][8]);
        }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int col0 = 16 * (4 * NBF + r) + 4 * c.part;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (STEM) zr[r][j] += S.b0[col0 + j];
            sum += zr[r][j];
            sq = fmaf(zr[r][j], zr[r][j], sq);
        }
        if (zrow) *reinterpret_cast<float4*>(zrow + size_t(col0 & ~7) * 128 + (col0 & 7)) = make_float4(zr[r][0], zr[r][1], zr[r][2], zr[r][3]);
    }
    S.red[0][c.part][c.row] = sum;
    S.red[1][c.part][c.row] = sq;
    tc::fence_before_sync();          // every thread's reads of D are complete before the next GEMM may overwrite it
    row_sync();
    const float inv_h = 1.0f / float(p.h);
    const float mean = ((S.red[0][0][c.row] + S.red[0][1][c.row]) + (S.red[0][2][c.row] + S.red[0][3][c.row])) * inv_h;
    const float msq = ((S.red[1][0][c.row] + S.red[1][1][c.row]) + (S.red[1][2][c.row] + S.red[1][3][c.row])) * inv_h;
    const float var = fmaxf(msq - mean * mean, 0.f);
    // rsqrt + one Newton step (within an ulp of 1 / sqrt; the IEEE divide and square root are ~25 instructions per row thread)
    const float ve = var + 1e-5f, r0 = rsqrtf(ve);
    const float rstd = r0 * fmaf(-0.5f * ve, r0 * r0, 1.5f);
    const float shift = -mean * rstd;       // xhat = z * rstd + shift, the same expression in the backward pass
    if (c.part == 0) {
        S.stats[l][0][c.row] = mean;
        S.stats[l][1][c.row] = rstd;
    }
    const float* gam = S.gamma[l];
    const float* bet = S.beta[l];
#pragma unroll
    for (int q = 0; q < 5; ++q) o[q] = 0.f;
    const bool write_a = !last || p.backward;        // h_L is an operand only as the image the head weight gradients read
    if (write_a) wait_a_free(S, p, c);

    // ---- pass 2, 8 columns at a time
    const float dscale = STEM ? 1.0f : p.drop_scale;
    auto unit8 = [&](const float* zz, int col, float* x, uint32_t keep) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float y = fmaf(fmaf(zz[j], rstd, shift), gam[col + j], bet[col + j]);
            if (DROP) {                // ReLU and mask in one gate; the 1 / (1 - p) scale rides on the residual add
                x[j] = fmaf(gate_keep_pos(y, y, keep, 1u << j), dscale, x[j]);
            } else {
                x[j] += fmaxf(y, 0.f); // x = h_{l-1} (0 for the stem)
            }
        }
        if (last) {
#pragma unroll
            for (int q = 0; q < 5; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) o[q] = fmaf(S.headw[q][col + j], x[j], o[q]);
        }
    };
#pragma unroll
    for (int i = 0; i < NBF; ++i) {
        const int blk = 4 * i + c.part;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int col = 16 * blk + 8 * u;
            float x[8];
            if (!STEM) tc::tmem_ld8(c.tX + uint32_t(col), x);
            const uint32_t keep = (!STEM && DROP) ? uint32_t(keep_bits >> (16 * i + 8 * u)) & 0xFFu : 0xFFu;
            if (STEM) {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = 0.f;
            }
            unit8(&z[i][8 * u], col, x, keep);
            if (!last) tc::tmem_st8(c.tX + uint32_t(col), x);
            if (write_a) store_unit<HP>(c, blk, u, x, !c.valid);
        }
        if (!last) warp_arrive(&S.rnd_ready[i], c.lane);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int blk = 4 * NBF + r, col = 16 * blk + 4 * c.part;
        float x[4];
        if (!STEM) {
            uint32_t raw[4];
            tc::tmem_ld4_issue(c.tX + uint32_t(col), raw);
            tc::tmem_ld_wait_all();
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = tc::tmem_ld_pin(raw[j]);
        }
        if (STEM) x[0] = x[1] = x[2] = x[3] = 0.f;
        const uint32_t keep = (!STEM && DROP) ? uint32_t(keep_bits >> (16 * NBF + 4 * r)) & 0xFu : 0xFu;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float y = fmaf(fmaf(zr[r][j], rstd, shift), gam[col + j], bet[col + j]);
            if (DROP) {
                x[j] = fmaf(gate_keep_pos(y, y, keep, 1u << j), dscale, x[j]);
            } else {
                x[j] += fmaxf(y, 0.f);
            }
        }
        if (last) {
#pragma unroll
            for (int q = 0; q < 5; ++q)
#pragma unroll
                for (int j = 0; j < 4; ++j) o[q] = fmaf(S.headw[q][col + j], x[j], o[q]);
        } else {
            tc::tmem_st4(c.tX + uint32_t(col), x);
        }
        if (write_a) store_quad<HP>(c, blk, x, !c.valid);
        if (!last) warp_arrive(&S.rnd_ready[NBF + r], c.lane);
    }
    tc::tmem_st_wait();
    if (p.backward) {                  // the operand tile is complete: the image of h_l may leave
        warp_arrive(&S.img_ready, c.lane);
        ++c.images;
    }
    if (last) {
        // partial head dots of parts 1..3 -> part 0 (through the column-sum scratch, free at this point)
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

// --------------------------------------------------------------------------------------------------- backward epilogue
// Backward through LayerNorm l and its ReLU / Dropout.  first (l == L): z_L is still in D and dh_L comes from the head
// gradients; otherwise z_l comes from the scratch and dh_l = X + D (D = dz_{l+1} W_{l+1}).  Pass A keeps xhat in registers
// (which frees D for the next GEMM) and parks dh_l in X; pass B turns them into dz_l: the image of dz_l and, for l > 0, the
// next GEMM's operand, signalled round by round.
template <int HP, bool DROP>
__device__ __forceinline__ void bwd_epilogue(Smem<HP>& S, const Params& p, RowCtx& c, int l, bool first, uint64_t keep_bits) {
    using SM = Smem<HP>;
    constexpr int NBF = SM::NBF, NR = SM::NR;
    const int h = p.h, L = p.L;
    const float inv_h = 1.0f / float(h);
    const float mean = S.stats[l][0][c.row], rstd = S.stats[l][1][c.row], shift = -mean * rstd;
    const float* gam = S.gamma[l];
    const float* bet = S.beta[l];
    const float* zrow = first ? nullptr : p.zscratch + ((size_t(blockIdx.x) * L + l) * (HP / 8) * 128 + c.row) * 8;
    const float dscale = (DROP && l > 0) ? p.drop_scale : 1.0f;
    if (!(DROP && l > 0)) keep_bits = ~0ull;
    float dh5[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) dh5[q] = S.dhead[c.row][q];
    if (p.decouple) dh5[4] = 0.f;                      // value head sees x.detach() (game.py:1208)
    float* lnp = p.ln_part + (((size_t(blockIdx.x) * 4 + (c.row >> 5)) * (L + 1) + l) * 2) * HP;   // [dgamma | dbeta]
    float (*scr)[32] = S.scratch[c.warp];
    float xh[NBF > 0 ? NBF : 1][16], xhr[NR > 0 ? NR : 1][4];
    float s1 = 0.f, s2 = 0.f;

    // ---- pass A, 8 columns at a time (4 for the remainder units).  (Requesting z_l of all 52 columns up front, straight into
    // the xhat registers, was measured: 5 % slower -- the longer live ranges spill.)
    auto unitA = [&](int col, int n, const float* z, const float* dh, float* xhat, uint32_t keep) {
        float gx[8], gg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j >= n) break;
            const float x = fmaf(z[j], rstd, shift);
            const float y = fmaf(x, gam[col + j], bet[col + j]);
            // (without the dropout scale 1 / (1 - p): everything below is linear in it, so it is applied once per column sum and
            // once per row -- rstd_d in pass B)
            const float gj = DROP ? gate_keep_pos(y, dh[j], keep, 1u << j) : (y > 0.f ? dh[j] : 0.f);
            const float t = gj * gam[col + j];
            s1 += t;
            s2 = fmaf(t, x, s2);
            xhat[j] = x;
            gg[j] = gj;
            gx[j] = gj * x;
        }
        colsum4x2(scr, c.lane, gx[0], gx[1], gx[2], gx[3], gg[0], gg[1], gg[2], gg[3], lnp + col, lnp + HP + col, dscale);
        if (n == 8) colsum4x2(scr, c.lane, gx[4], gx[5], gx[6], gx[7], gg[4], gg[5], gg[6], gg[7], lnp + col + 4, lnp + HP + col + 4, dscale);
    };
#pragma unroll
    for (int i = 0; i < NBF; ++i) {
        const int blk = 4 * i + c.part;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int col = 16 * blk + 8 * u;
            float z[8], dh[8];
            if (first) {
                tc::tmem_ld8(c.tD + uint32_t(col), z);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float a = 0.f;
#pragma unroll
                    for (int q = 0; q < 5; ++q) a = fmaf(dh5[q], S.headw[q][col + j], a);
                    dh[j] = a;
                }
            } else {
                ld256(zrow + size_t(col) * 128, z);
                float d[8];
                tc::tmem_ld8x2(c.tX + uint32_t(col), dh, c.tD + uint32_t(col), d);
#pragma unroll
                for (int j = 0; j < 8; ++j) dh[j] += d[j];
            }
            unitA(col, 8, z, dh, &xh[i][8 * u], uint32_t(keep_bits >> (16 * i + 8 * u)) & 0xFFu);
            tc::tmem_st8(c.tX + uint32_t(col), dh);      // dh_l
        }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int col = 16 * (4 * NBF + r) + 4 * c.part;
        float z[8], dh[8];
        if (first) {
            uint32_t raw[4];
            tc::tmem_ld4_issue(c.tD + uint32_t(col), raw);
            tc::tmem_ld_wait_all();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                z[j] = tc::tmem_ld_pin(raw[j]);
                float a = 0.f;
#pragma unroll
                for (int q = 0; q < 5; ++q) a = fmaf(dh5[q], S.headw[q][col + j], a);
                dh[j] = a;
            }
        } else {
            const float4 zz = *reinterpret_cast<const float4*>(zrow + size_t(col & ~7) * 128 + (col & 7));
            z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
            uint32_t ra[4], rb[4];
            tc::tmem_ld4_issue(c.tX + uint32_t(col), ra);
            tc::tmem_ld4_issue(c.tD + uint32_t(col), rb);
            tc::tmem_ld_wait_all();
#pragma unroll
            for (int j = 0; j < 4; ++j) dh[j] = tc::tmem_ld_pin(ra[j]) + tc::tmem_ld_pin(rb[j]);
        }
        unitA(col, 4, z, dh, &xhr[r][0], uint32_t(keep_bits >> (16 * NBF + 4 * r)) & 0xFu);
        tc::tmem_st4(c.tX + uint32_t(col), dh);
    }
    tc::tmem_st_wait();
    S.red[0][c.part][c.row] = s1;
    S.red[1][c.part][c.row] = s2;
    tc::fence_before_sync();          // D has been read by everyone before the next GEMM's first block is signalled
    row_sync();
    const float m1 = ((S.red[0][0][c.row] + S.red[0][1][c.row]) + (S.red[0][2][c.row] + S.red[0][3][c.row])) * inv_h;
    const float m2 = ((S.red[1][0][c.row] + S.red[1][1][c.row]) + (S.red[1][2][c.row] + S.red[1][3][c.row])) * inv_h;
    wait_a_free(S, p, c);

    // ---- pass B: dz_l = rstd * (t - mean(t) - xhat * mean(t * xhat))
    const float rstd_d = rstd * dscale;
    const bool feeds_gemm = l > 0;
#pragma unroll
    for (int i = 0; i < NBF; ++i) {
        const int blk = 4 * i + c.part;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int col = 16 * blk + 8 * u;
            float dh[8], dz[8];
            tc::tmem_ld8(c.tX + uint32_t(col), dh);
            const uint32_t keep = uint32_t(keep_bits >> (16 * i + 8 * u)) & 0xFFu;
            // padded columns (>= h, a multiple of 4) get dz = 0: the factor is chosen per half unit instead of a select per element
            const float rs[2] = {col + 4 <= h ? rstd_d : 0.f, col + 8 <= h ? rstd_d : 0.f};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float x = xh[i][8 * u + j];
                const float y = fmaf(x, gam[col + j], bet[col + j]);
                const float t = (DROP ? gate_keep_pos(y, dh[j], keep, 1u << j) : (y > 0.f ? dh[j] : 0.f)) * gam[col + j];
                // (measured: the factor form is 0.25 ms per chunk faster in the DROP kernel and 0.3 ms slower in the other one)
                dz[j] = DROP ? rs[j >> 2] * (t - m1 - x * m2) : ((col + j < h) ? rstd_d * (t - m1 - x * m2) : 0.f);
            }
            store_unit<HP>(c, blk, u, dz, !c.valid);
        }
        if (feeds_gemm) warp_arrive(&S.rnd_ready[i], c.lane);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int blk = 4 * NBF + r, col = 16 * blk + 4 * c.part;
        float dz[4];
        uint32_t raw[4];
        tc::tmem_ld4_issue(c.tX + uint32_t(col), raw);
        tc::tmem_ld_wait_all();
        const uint32_t keep = uint32_t(keep_bits >> (16 * NBF + 4 * r)) & 0xFu;
        const float rs = col + 4 <= h ? rstd_d : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float x = xhr[r][j], dhj = tc::tmem_ld_pin(raw[j]);
            const float y = fmaf(x, gam[col + j], bet[col + j]);
            const float t = (DROP ? gate_keep_pos(y, dhj, keep, 1u << j) : (y > 0.f ? dhj : 0.f)) * gam[col + j];
            dz[j] = DROP ? rs * (t - m1 - x * m2) : ((col + j < h) ? rstd_d * (t - m1 - x * m2) : 0.f);
        }
        store_quad<HP>(c, blk, dz, !c.valid);
        if (feeds_gemm) warp_arrive(&S.rnd_ready[NBF + r], c.lane);
    }
    warp_arrive(&S.img_ready, c.lane);      // the image of dz_l may leave
    ++c.images;
}

// ------------------------------------------------------------------------------------------------------- control warps
// issuer: the MMAs of every stage of every tile (warp-uniform code, one elected lane issues; see g2048_rollout_x3.cu)
template <int HP>
__device__ __forceinline__ void issuer(Smem<HP>& S, const Params& p, uint32_t tmem_base, uint32_t my_tiles) {
    using SM = Smem<HP>;
    constexpr int NBF = SM::NBF, NR = SM::NR, RING = SM::RING;
    const int L = p.L;
    const uint32_t idesc = tc::make_idesc_f16(128, HP);
    const uint64_t dA = tc::make_desc_sw32(tc::smem_addr(S.A[0]), 16, 256), dW = tc::make_desc_sw32(tc::smem_addr(S.W[0]), 16, 256);
    const uint32_t dA_lo = uint32_t(dA), dW_lo = uint32_t(dW), d_hi = uint32_t(dA >> 32);
    auto desc = [&](uint32_t lo) { return uint64_t(lo) | uint64_t(d_hi) << 32; };
    uint32_t slot = 0, full_par = 0, w_off = 0;
    // kind 0: stem (the 16 exponents are exact in fp16: hi part only, two products); 1: hi | lo operand, three products
    auto issue_block = [&](uint32_t a_off, bool first, bool stem) {
        tc::mbar_wait(&S.w_full[slot], full_par);
        if (elect_one()) {
            const uint32_t ah = dA_lo + a_off, al = ah + (SM::PART >> 4), bh = dW_lo + w_off, bl = bh + (SM::WPART >> 4);
            if (stem) {
                tc::mma_bf16_ss(tmem_base, desc(ah), desc(bl), idesc, false);
            } else {
                tc::mma_bf16_ss(tmem_base, desc(al), desc(bh), idesc, !first);
                tc::mma_bf16_ss(tmem_base, desc(ah), desc(bl), idesc, true);
            }
            tc::mma_bf16_ss(tmem_base, desc(ah), desc(bh), idesc, true);
            tc::mma_commit(&S.w_empty[slot]);
        }
        w_off += SM::SLOT >> 4;
        if (++slot == uint32_t(RING)) { slot = 0; w_off = 0; full_par ^= 1u; }
    };
    uint32_t in_par = 0, rnd_par = 0;
    auto gemm = [&]() {
#pragma unroll
        for (int i = 0; i < NBF; ++i) {
            tc::mbar_wait(&S.rnd_ready[i], rnd_par);
            tc::fence_after_sync();
#pragma unroll
            for (int part = 0; part < 4; ++part) issue_block(uint32_t(4 * i + part) * 256u, i == 0 && part == 0, false);
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            tc::mbar_wait(&S.rnd_ready[NBF + r], rnd_par);
            tc::fence_after_sync();
            issue_block(uint32_t(4 * NBF + r) * 256u, NBF == 0 && r == 0, false);
        }
        if (elect_one()) tc::mma_commit(&S.mma_done);
        rnd_par ^= 1u;
    };
    for (uint32_t t = 0; t < my_tiles; ++t) {
        tc::mbar_wait(&S.in_ready, in_par);
        in_par ^= 1u;
        tc::fence_after_sync();
        issue_block(0u, true, true);
        if (elect_one()) tc::mma_commit(&S.mma_done);
        for (int l = 1; l <= L; ++l) gemm();                 // z_l = h_{l-1} W_l^T
        if (p.backward)
            for (int l = L; l >= 1; --l) gemm();             // D = dz_l W_l
    }
}

// producer: block k of the launch-wide weight sequence (period = blocks of one tile) goes to slot k % RING once the MMAs
// that read the slot's previous block have completed
template <int HP>
__device__ __forceinline__ void producer(Smem<HP>& S, const Params& p, uint32_t my_tiles) {
    using SM = Smem<HP>;
    constexpr int RING = SM::RING;
    const uint32_t nblk = uint32_t(blocks_per_tile(HP, p.L, p.backward != 0));
    uint32_t slot = 0, par = 1, idx = 0;
    const uint8_t* src = p.img;
    bool first_lap = true;
    for (uint32_t k = my_tiles * nblk; k > 0; --k) {
        if (!first_lap) tc::mbar_wait(&S.w_empty[slot], par);
        if (elect_one()) {
            tc::mbar_expect_tx(&S.w_full[slot], SM::SLOT);
            tc::bulk_g2s(S.W[slot], src, SM::SLOT, &S.w_full[slot]);
        }
        src += SM::SLOT;
        if (++idx == nblk) { idx = 0; src = p.img; }
        if (++slot == uint32_t(RING)) { slot = 0; par ^= 1u; first_lap = false; }
    }
}

// storer (backward mode): every finished operand tile leaves as the image of h_l / dz_l, two bulk copies of one part each;
// the row threads overwrite the tile only after its bytes have been read
template <int HP>
__device__ __forceinline__ void storer(Smem<HP>& S, const Params& p, uint32_t my_tiles) {
    using SM = Smem<HP>;
    if (!p.backward) return;
    const int L = p.L;
    const size_t tile_bytes = size_t(HP) * 512;                           // 128 samples x HP x 4 B
    uint32_t par = 0;
    const uint64_t stream_policy = tc::l2_policy_evict_first();
    for (uint32_t t = 0; t < my_tiles; ++t) {
        const size_t tile = size_t(blockIdx.x) + size_t(t) * gridDim.x;
        for (int k = 0; k < 2 * (L + 1); ++k) {                          // h_0 .. h_L, dz_L .. dz_0
            const bool fwd = k <= L;
            const int l = fwd ? k : 2 * L + 1 - k;
            uint8_t* dst = (fwd ? p.h_out : p.dz_out) + (size_t(l) * size_t(p.ntiles) + tile) * tile_bytes;
            tc::mbar_wait(&S.img_ready, par);
            par ^= 1u;
            if (elect_one()) {
                tc::bulk_s2g_hint(dst, S.A[0], SM::PART, stream_policy);       // evict-first: the images are read by the weight-gradient
                tc::bulk_s2g_hint(dst + SM::PART, S.A[1], SM::PART, stream_policy);   // kernels later, the z scratch by this kernel soon
                tc::bulk_commit();
                tc::bulk_wait_read0();
                tc::mbar_arrive(&S.a_free);
            }
            __syncwarp();
        }
    }
    if (elect_one()) tc::bulk_wait0();                                   // every image is complete in HBM before the kernel ends
    __syncwarp();
}

template <int HP, bool DROP>
__global__ void __launch_bounds__(THREADS, 1) update_mlp_x3_kernel(const Params p) {
    using SM = Smem<HP>;
    extern __shared__ uint8_t smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw + ((1024u - (tc::smem_addr(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.L;
    const int64_t ntiles = p.ntiles;
    const uint32_t my_tiles = ntiles > blockIdx.x ? uint32_t((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;

    if (warp == 0) tc::tmem_alloc(&S.tmem_base, 512);
    if (tid == 0) {
        tc::mbar_init(&S.in_ready, 4);
        tc::mbar_init(&S.mma_done, 1);
        tc::mbar_init(&S.img_ready, ROW_THREADS / 32);
        tc::mbar_init(&S.a_free, 1);
        for (int i = 0; i < SM::ROUNDS; ++i) tc::mbar_init(&S.rnd_ready[i], ROW_THREADS / 32);
        for (int i = 0; i < SM::RING; ++i) {
            tc::mbar_init(&S.w_full[i], 1);
            tc::mbar_init(&S.w_empty[i], 1);
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
    for (uint32_t i = tid * 16; i < 2 * SM::PART; i += THREADS * 16) *reinterpret_cast<uint4*>(&S.A[0][0] + i) = make_uint4(0, 0, 0, 0);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = S.tmem_base;

    if (warp == ISSUER_WARP) {
        issuer<HP>(S, p, tmem_base, my_tiles);
    } else if (warp == PRODUCER_WARP) {
        producer<HP>(S, p, my_tiles);
    } else if (warp == STORER_WARP) {
        storer<HP>(S, p, my_tiles);
    } else {
        // ---------------- row threads
        RowCtx c;
        c.warp = warp;
        c.lane = lane;
        c.part = warp >> 2;
        c.row = (warp & 3) * 32 + lane;
        c.tD = tmem_base + (uint32_t((warp & 3) * 32) << 16);
        c.tX = c.tD + X_COL;
        c.a_row = tc::smem_addr(S.A[0]) + uint32_t(c.row) * 32u;
        c.sw = uint32_t(c.row >> 2) & 1u;
        c.images = 0;
        uint32_t mma_par = 0;
        double lacc[4] = {0.0, 0.0, 0.0, 0.0};
        float hb[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (uint32_t t = 0; t < my_tiles; ++t) {
            c.tile = int64_t(blockIdx.x) + int64_t(t) * gridDim.x;
            c.grow = c.tile * 128 + c.row;
            c.valid = c.grow < p.n;
            if (c.part == 0) {
                // model input: the 16 exponents are exact in fp16 (k-block 0, hi part; row / column features are folded into b0,
                // SURVEY A10); the previous tile's dz_0 must have left the operand buffers
                const uint64_t b = c.valid ? p.boards[c.grow] : 0ull;
                wait_a_free(S, p, c);
                uint32_t w[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const __half2 pr = __floats2half2_rn(float((b >> (8 * q)) & 15ull), float((b >> (8 * q + 4)) & 15ull));
                    w[q] = *reinterpret_cast<const uint32_t*>(&pr);
                }
                const uint32_t a0 = c.a_row + (c.sw << 4), a1 = a0 ^ 16u;
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a1), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
                warp_arrive(&S.in_ready, lane);
            }
            float o[5];
            uint64_t keep1 = 0ull, keep2 = 0ull;                    // dropout masks of this thread's columns in blocks 1 and 2
            for (int l = 0; l <= L; ++l) {                          // z_l is in D once the stage's MMAs are done
                uint64_t kb = 0ull;
                if (DROP && l > 0) kb = dropout_keep_row<SM::NBF, SM::NR>(p, p.sample0 + c.grow, l, c.part);   // under the MMAs
                tc::mbar_wait(&S.mma_done, mma_par);
                mma_par ^= 1u;
                tc::fence_after_sync();
                fwd_epilogue<HP, DROP>(S, p, c, l, o, kb);
                if (DROP) {
                    if (l == 1) keep1 = kb;
                    if (l == 2) keep2 = kb;
                }
            }
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
                    ppo_sample(l4, p.legal[c.grow] & 15u, a, lp_old, p.adv[c.grow], o[4], p.g_norm[c.grow], p.clip_eps, p.beta_ent, ppo, vl,
                               H, gl, dvl);
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
            // ---- backward-data: LayerNorm L first (z_L still in D), then l = L-1 .. 0 once D = dz_{l+1} W_{l+1} is done
            for (int l = L; l >= 0; --l) {
                if (l < L) {
                    tc::mbar_wait(&S.mma_done, mma_par);
                    mma_par ^= 1u;
                    tc::fence_after_sync();
                }
                bwd_epilogue<HP, DROP>(S, p, c, l, l == L, l == 2 ? keep2 : keep1);
            }
        }
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
    const int NB = HP / 16;
    const int nfwd = 1 + L * NB, nblk = nfwd + L * NB;
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
    // ---- weight k-blocks (hi | lo in fp16): block 0 = stem exponent columns, then W_1..W_L, then W_L^T..W_1^T
    if (idx < int64_t(nblk) * HP * 16) {
        const int b = int(idx / (HP * 16)), rem = int(idx % (HP * 16)), n = rem / 16, kk = rem % 16;
        float v = 0.f;
        if (b == 0) {
            if (n < h) v = s.stem_w[n * 48 + 3 * kk];
        } else if (b < nfwd) {
            const int l = (b - 1) / NB, k = ((b - 1) % NB) * 16 + kk;
            if (n < h && k < h) v = s.w[l][size_t(n) * h + k];
        } else {
            const int t = b - nfwd, l = L - 1 - t / NB, k = (t % NB) * 16 + kk;
            if (n < h && k < h) v = s.w[l][size_t(k) * h + n];
        }
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const size_t part = size_t(HP) * 32;
        uint8_t* blk = img + size_t(b) * 2 * part;
        *reinterpret_cast<__half*>(blk + tc::sw32_offset(n, kk)) = hi;
        *reinterpret_cast<__half*>(blk + part + tc::sw32_offset(n, kk)) = lo;
    }
}

static bool shape_ok(int h, int L) { return h >= 16 && h <= 208 && h % 4 == 0 && L >= 1 && L <= MAXL; }

template <int HP>
static int launch(const Params& p, int grid, cudaStream_t st) {
    const int smem = int(sizeof(Smem<HP>)) + 1024;
    if (p.drop_thr) {
        G2048_CHECK_CUDA(ensure_smem(update_mlp_x3_kernel<HP, true>, smem));
        update_mlp_x3_kernel<HP, true><<<grid, THREADS, smem, st>>>(p);
    } else {
        G2048_CHECK_CUDA(ensure_smem(update_mlp_x3_kernel<HP, false>, smem));
        update_mlp_x3_kernel<HP, false><<<grid, THREADS, smem, st>>>(p);
    }
    G2048_CHECK_LAUNCH("update_mlp_x3_kernel");
    return G2048_OK;
}

}  // namespace ux
}  // namespace g2048

using namespace g2048;
using namespace g2048::ux;

extern "C" {

int32_t g2048_update_mlp_padded(int32_t hidden) { return hidden >= 1 && hidden <= 208 ? padded(hidden) : -1; }

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
    const int64_t work = int64_t(img_blocks(HP, layers)) * HP * 16;
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
        G2048_REQUIRE((reinterpret_cast<uintptr_t>(u->h_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(u->dz_out) & 15) == 0,
                      "g2048_update_mlp_fwd_bwd: h_out / dz_out must be 16-byte aligned");
    } else {
        G2048_REQUIRE(u->logits || u->value, "g2048_update_mlp_fwd_bwd: forward-only call without outputs");
    }
    const int64_t ntiles = (u->n + 127) / 128;
    const int grid = int(ntiles < num_sms() ? ntiles : num_sms());
    Params p{};
    p.n = u->n; p.ntiles = ntiles; p.h = h; p.L = L; p.decouple = u->decouple_critic; p.backward = bw;
    p.boards = u->boards; p.actions = u->actions; p.legal = u->legal; p.flags = u->flags;
    p.old_logp = u->old_logp; p.old_stride = u->old_logp_stride; p.adv = u->adv; p.g_norm = u->g_norm;
    p.clip_eps = u->clip_eps; p.c_v = u->critic_strength; p.beta_ent = u->entropy_strength; p.inv_n = u->inv_n;
    if (u->dropout_p > 0.f) {
        p.drop_thr = uint32_t(u->dropout_p * 65536.0f + 0.5f);
        p.drop_scale = 1.0f / (1.0f - u->dropout_p);
        p.drop_keys = philox_round_keys(u->dropout_seed);
        p.sample0 = u->dropout_sample0;
    }
    const uint8_t* base = static_cast<const uint8_t*>(u->packed);
    p.pf = reinterpret_cast<const float*>(base);
    p.img = base + img_offset_bytes(HP, L);
    p.h_out = reinterpret_cast<uint8_t*>(u->h_out); p.dz_out = reinterpret_cast<uint8_t*>(u->dz_out);
    p.dhead = u->dhead; p.logits = u->logits; p.value = u->value;
    uint8_t* ws = static_cast<uint8_t*>(u->workspace);
    const size_t z_bytes = size_t(num_sms()) * L * 128 * HP * 4, ln_bytes = size_t(num_sms()) * 4 * (L + 1) * 2 * HP * 4;
    p.zscratch = reinterpret_cast<float*>(ws);
    p.ln_part = reinterpret_cast<float*>(ws + z_bytes);
    p.head_part = reinterpret_cast<float*>(ws + z_bytes + ln_bytes);
    p.loss_part = reinterpret_cast<double*>(ws + z_bytes + ln_bytes + size_t(num_sms()) * 8 * 4);
    if (bw) G2048_CHECK_CUDA(cudaMemsetAsync(p.ln_part, 0, ln_bytes, st));
    int rc;
    switch (HP) {
        case 64: rc = launch<64>(p, grid, st); break;
        case 128: rc = launch<128>(p, grid, st); break;
        case 192: rc = launch<192>(p, grid, st); break;
        default: rc = launch<208>(p, grid, st); break;
    }
    if (rc != G2048_OK) return rc;
    if (bw) {
        const int work = (L + 1) * 2 * h + 9;
        update_reduce_kernel<<<(work + 127) / 128, 128, 0, st>>>(p.ln_part, p.head_part, p.loss_part, grid, L, HP, h, u->ln_grad,
                                                                u->head_bias_grad, u->stats);
        G2048_CHECK_LAUNCH("update_reduce_kernel");
    }
    return G2048_OK;
}

}  // extern "C"
