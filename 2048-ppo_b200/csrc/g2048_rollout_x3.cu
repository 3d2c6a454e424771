// g2048_rollout_x3.cu -- the fused rollout kernel on the tensor cores at fp32 grade (the default at large env batch).
//
// Same records, env semantics and 128-env tiles as g2048_rollout_tc.cu (env m of the tile = TMEM lane m, four threads
// per env row), but every GEMM operand is TWO fp16 terms, x = hi + lo (22 mantissa bits), and every k-step is three
// tcgen05.mma products lo*hi + hi*lo + hi*hi accumulated in fp32 in tensor memory: the recorded log-probs and values
// agree with the reference's fp32 forward (train.py:256-274, game.py:1192-1203) to ~1e-6, where bf16 operands give
// ~1e-2.  (fp16 rather than bf16 terms: 11 + 11 bits instead of 8 + 8; activations and weights of this network sit
// far inside the fp16 range, and what falls below it is < 6e-8 absolute.)
//
// Three times the MMAs no longer hide behind nothing, so the stages are pipelined inside a tile:
//   * pass 1 of an epilogue pulls the thread's 48..52 accumulator columns into REGISTERS (and takes the LayerNorm
//     statistics from them), which frees the accumulator D for the next GEMM at once;
//   * pass 2 works through the 16-column k-blocks in rounds (round i = blocks 4i..4i+3, one per column part), writes
//     each block's hi | lo operand bytes and the fp32 residual stream (TMEM), and signals the round;
//   * a control warp issues the next layer's MMAs for the blocks of a round as soon as it is signalled -- the tensor
//     pipe runs under pass 2 -- and streams the weight k-blocks (hi | lo, HP x 64 B) L2 -> SMEM through a ring of
//     bulk async copies, refilling a slot two blocks behind the issue point.
// TMEM: columns [0,HP) = D, [256,256+HP) = the fp32 residual stream X.  SMEM: A = hi | lo operand tile (32-byte
// swizzle, HP/16 blocks of 128 rows x 32 B per part), the weight ring, the stem block, LayerNorm / head parameters.
#include <cuda_fp16.h>
#include "g2048_rollout_tail.cuh"
#include "g2048_tc.cuh"

// Softmax / log / divide of the sampling tail through the hardware ex2 / lg2 / rcp units (one thread per env walks this chain while
// the other three wait: 8 % of the C3 rollout time with the IEEE library functions).  Measured (tools/x3_tail_precision.py, three
// models): the largest log-prob / value / entropy distances from the torch fp32 and float64 policies are the SAME to three digits
// with either form -- what separates the kernel from torch is the fp32 rounding of the trunk, not these functions.
#ifndef G2048_X3_FAST_TAIL
#define G2048_X3_FAST_TAIL true
#endif
namespace g2048 {
namespace x3 {

constexpr int SPLIT = 4;                          // threads per env row (column parts)
constexpr int ENV_THREADS = 128 * SPLIT;          // 16 warps: warp w owns lane quarter (w & 3) and column part (w >> 2)
constexpr int THREADS = ENV_THREADS + 64;         // + two control warps: MMA issuer, weight producer (5 warps on a scheduler
                                                  // cap the registers at 96 per thread whether there are 17 or 20 of them)
constexpr uint32_t X_COL = 256;
constexpr int MAX_LAYERS = 4;

// The move of one env in one direction, computed ahead of the policy's choice (see precompute_moves).
struct MovePre {
    uint64_t moved;          // board after the move, before the spawn
    uint32_t pa[2];          // packed potentials of `moved`
    int32_t points;          // merge points
    uint32_t meta;           // bit 0 valid (the move changes the board), bit 1 nibble overflow, bits 8.. largest exponent created
};
// Per-env exchange between the four threads of a row.
struct Xch {
    uint64_t board;          // the board the step starts from (part 0, at the start of the step)
    uint32_t pb[2];          // packed potentials of `board` (part 2)
    uint32_t draw[3];        // the step's Philox draws: spawn cell, spawn value, action (part 1)
    uint32_t pad_;
};

constexpr int ring_slots(int HP) {
    // what is left of 227 KB after the operand tile, the stem block, the parameters and the exchange buffers
    const int fixed = 2 * (HP / 16) * 4096 + HP * 64 + (1 + 2 * (MAX_LAYERS + 1)) * HP * 4 + (5 * HP + 8) * 4 +
                      2 * 4 * 128 * 4 + 3 * 128 * 5 * 4 + 128 * int(sizeof(Xch)) + 128 * 4 * int(sizeof(MovePre)) + 1024;
    const int n = (232448 - fixed) / (HP * 64);
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
    alignas(1024) uint8_t A[2][PART];
    alignas(1024) uint8_t W[RING][SLOT];
    alignas(1024) uint8_t Wstem[SLOT];
    alignas(16) float b0[HP];
    alignas(16) float ln_g[MAX_LAYERS + 1][HP];   // [0] = stem LayerNorm
    alignas(16) float ln_b[MAX_LAYERS + 1][HP];
    alignas(16) float headw[5 * HP + 8];
    float red[2][SPLIT][128];                     // [sum | sq][column part][row]
    float headp[SPLIT - 1][128][5];               // partial head dots of parts 1..3
    Xch xch[128];
    alignas(8) MovePre pre[128][4];
    uint64_t in_ready, mma_done, stem_full, rnd_ready[ROUNDS > 0 ? ROUNDS : 1], w_full[RING], w_empty[RING];
    uint32_t tmem_base;
};
static_assert(sizeof(Smem<208>) + 1024 <= 232448, "x3 rollout kernel exceeds the 227 KB shared memory limit");
static_assert(Smem<208>::RING >= 4, "weight ring too short to cover the L2 latency");

// The four threads of an env row sit in the four warps of one lane quarter (warps q, q + 4, q + 8, q + 12), and everything they
// exchange is per row: one named barrier per quarter (128 threads) instead of one over all 512 env threads, so that a slow
// quarter does not hold up the other three.
__device__ __forceinline__ void row_sync(int quarter) { asm volatile("bar.sync %0, 128;" ::"r"(quarter + 1) : "memory"); }

// one warp's arrival on a round barrier: every lane's operand bytes visible to the async proxy first
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    tc::fence_async_smem();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(bar);
}

// While the tensor pipe runs the first residual block, the four threads of a row play the env's move in all four
// directions (thread `dir` = its column part: lookups, merge points, potentials of the moved board) and park the results
// in shared memory; part 1 also draws the step's Philox numbers, part 2 (LEFT: its lookups are the board's own rows) the
// potentials of the current board.  The tail of the step -- after the heads -- then only samples, picks the chosen
// direction's result, spawns and records: ~1000 dependent instructions with three rounds of L2 table reads and two
// 512-thread barriers left the critical path (they sat there with one active warp per scheduler; ncu, profiles/).
template <int HP>
__device__ __forceinline__ void precompute_moves(Smem<HP>& S, const RolloutParams& p, const LutGlobal& lut, int row, uint32_t dir,
                                                 int64_t env, uint64_t ctr) {
    const Board board = make_board(S.xch[row].board);
    const Board bt = transpose(board);
    const Board canon = to_canonical(board, bt, dir);
    const Lines mv = lookup_rows(canon, lut);
    Lines cols = mv;
    if (dir == 2u) cols = lookup_rows(bt, lut);          // in flight with the rows
    if (dir == 1u) {
        const U4 d = env_draws(p.seed, p.env0 + uint64_t(env), ctr);
        S.xch[row].draw[0] = d.x;
        S.xch[row].draw[1] = d.y;
        S.xch[row].draw[2] = d.z;
    }
    const Board moved_c = result_of(mv);
    const bool valid = !same(moved_c, canon);
    int points, max_tile;
    bool ovf;
    merge_stats(mv, points, max_tile, ovf);
    const Board moved = from_canonical(moved_c, dir);
    const uint2 pa = pack_potentials(potentials(moved, lookup_rows(moved_c, lut), lookup_rows(transpose(moved_c), lut)));
    if (dir == 2u) {
        const uint2 pb = pack_potentials(potentials(board, mv, cols));
        S.xch[row].pb[0] = pb.x;
        S.xch[row].pb[1] = pb.y;
    }
    MovePre& m = S.pre[row][dir];
    m.moved = pack_board(moved);
    m.pa[0] = pa.x;
    m.pa[1] = pa.y;
    m.points = points;
    m.meta = uint32_t(valid) | uint32_t(ovf) << 1 | uint32_t(max_tile) << 8;
}

struct RowCtx {
    int row, part, lane;
    uint32_t tD, tX;             // TMEM addresses of this thread's lane quarter: D and X, column 0
    uint32_t a_row;              // shared address of A[0] + row * 32
    uint32_t sw;                 // (row >> 2) & 1: the 16-byte halves of a 32-byte operand row are swapped
};

// LayerNorm (eps 1e-5) + ReLU (+ residual) of one env row over this thread's columns (game.py:1038-1046, 1069-1073):
// blocks 4i + part (16 columns each) and a 4-column unit of every remainder block.  !HEADS: writes X and the next A
// operand and signals the rounds; HEADS (last stage): the 5 head dot products over this thread's columns instead.
template <int HP, bool STEM, bool HEADS>
__device__ __forceinline__ void epilogue(Smem<HP>& S, const RowCtx& c, int h, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float (&o)[5]) {
    using SM = Smem<HP>;
    constexpr int NBF = SM::NBF, NR = SM::NR;
    // Columns >= h are padding: their weights, biases, gamma and beta are zero in the packed buffer, so they produce
    // z = 0 and x = 0 without any masking here.
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
    // ---- pass 1: statistics (packed fp32 math: two columns per instruction)
    float2 sum2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NBF; ++i) {
        if (STEM) {
            const float4* b4 = reinterpret_cast<const float4*>(S.b0 + 16 * (4 * i + c.part));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 b = b4[q];
                z[i][4 * q] += b.x; z[i][4 * q + 1] += b.y; z[i][4 * q + 2] += b.z; z[i][4 * q + 3] += b.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 v = make_float2(z[i][2 * j], z[i][2 * j + 1]);
            sum2 = __fadd2_rn(sum2, v);
            sq2 = __ffma2_rn(v, v, sq2);
        }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        if (STEM) {
            const float4 b = *reinterpret_cast<const float4*>(S.b0 + 16 * (4 * NBF + r) + 4 * c.part);
            zr[r][0] += b.x; zr[r][1] += b.y; zr[r][2] += b.z; zr[r][3] += b.w;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float2 v = make_float2(zr[r][2 * j], zr[r][2 * j + 1]);
            sum2 = __fadd2_rn(sum2, v);
            sq2 = __ffma2_rn(v, v, sq2);
        }
    }
    S.red[0][c.part][c.row] = sum2.x + sum2.y;
    S.red[1][c.part][c.row] = sq2.x + sq2.y;
    tc::fence_before_sync();          // every thread's reads of D are complete before the next GEMM may overwrite it
    row_sync(c.row >> 5);
    const float inv_h = 1.0f / float(h);
    const float tsum = (S.red[0][0][c.row] + S.red[0][1][c.row]) + (S.red[0][2][c.row] + S.red[0][3][c.row]);
    const float tsq = (S.red[1][0][c.row] + S.red[1][1][c.row]) + (S.red[1][2][c.row] + S.red[1][3][c.row]);
    const float mean = tsum * inv_h;
    const float var = fmaxf(tsq * inv_h - mean * mean, 0.f);
    // rsqrt + one Newton step (within an ulp of 1 / sqrt; the IEEE divide and square root are ~25 instructions per row thread)
    const float ve = var + 1e-5f, r0 = rsqrtf(ve);
    const float rstd = r0 * fmaf(-0.5f * ve, r0 * r0, 1.5f);
    const float2 rstd2 = make_float2(rstd, rstd), shift2 = make_float2(-mean * rstd, -mean * rstd);
    float2 o2[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) o2[q] = make_float2(0.f, 0.f);

    // ---- pass 2, full blocks, 8 columns (one 16-byte operand unit per part) at a time; the residual stream of the next
    // unit is requested from TMEM before this one is worked on
    uint32_t xraw[8];
    if (!STEM && NBF > 0) tc::tmem_ld8_issue(c.tX + uint32_t(16 * c.part), xraw);
#pragma unroll
    for (int i = 0; i < NBF; ++i) {
        const int blk = 4 * i + c.part;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int col0 = 16 * blk + 8 * u;
            float x[8];
            if (!STEM) {
                tc::tmem_ld_wait_all();
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = tc::tmem_ld_pin(xraw[j]);
                if (u == 0) tc::tmem_ld8_issue(c.tX + uint32_t(col0 + 8), xraw);
                else if (i + 1 < NBF) tc::tmem_ld8_issue(c.tX + uint32_t(16 * (blk + 4)), xraw);
            }
            const float4* g4 = reinterpret_cast<const float4*>(gamma + col0);
            const float4* e4 = reinterpret_cast<const float4*>(beta + col0);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 g = g4[q], e = e4[q];
                const float* zz = &z[i][8 * u + 4 * q];
                float2 y0 = __ffma2_rn(__ffma2_rn(make_float2(zz[0], zz[1]), rstd2, shift2), make_float2(g.x, g.y), make_float2(e.x, e.y));
                float2 y1 = __ffma2_rn(__ffma2_rn(make_float2(zz[2], zz[3]), rstd2, shift2), make_float2(g.z, g.w), make_float2(e.z, e.w));
                y0.x = fmaxf(y0.x, 0.f); y0.y = fmaxf(y0.y, 0.f); y1.x = fmaxf(y1.x, 0.f); y1.y = fmaxf(y1.y, 0.f);
                if (!STEM) {
                    y0 = __fadd2_rn(make_float2(x[4 * q], x[4 * q + 1]), y0);
                    y1 = __fadd2_rn(make_float2(x[4 * q + 2], x[4 * q + 3]), y1);
                }
                x[4 * q] = y0.x; x[4 * q + 1] = y0.y; x[4 * q + 2] = y1.x; x[4 * q + 3] = y1.y;
            }
            if (HEADS) {
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const float4* hw = reinterpret_cast<const float4*>(S.headw + q * HP + col0);
                    const float4 w0 = hw[0], w1 = hw[1];
                    o2[q] = __ffma2_rn(make_float2(w0.x, w0.y), make_float2(x[0], x[1]), o2[q]);
                    o2[q] = __ffma2_rn(make_float2(w0.z, w0.w), make_float2(x[2], x[3]), o2[q]);
                    o2[q] = __ffma2_rn(make_float2(w1.x, w1.y), make_float2(x[4], x[5]), o2[q]);
                    o2[q] = __ffma2_rn(make_float2(w1.z, w1.w), make_float2(x[6], x[7]), o2[q]);
                }
            } else {
                tc::tmem_st8(c.tX + uint32_t(col0), x);
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) tc::split2_f16(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
                const uint32_t a = c.a_row + uint32_t(blk) * 4096u + ((uint32_t(u) ^ c.sw) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a + SM::PART), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
            }
        }
        if (!HEADS) warp_arrive(&S.rnd_ready[i], c.lane);
    }
    // ---- pass 2, remainder blocks: 4 columns per part
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int blk = 4 * NBF + r, col0 = 16 * blk + 4 * c.part;
        float x[4];
        if (!STEM) {
            uint32_t raw[4];
            tc::tmem_ld4_issue(c.tX + uint32_t(col0), raw);
            tc::tmem_ld_wait_all();
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = tc::tmem_ld_pin(raw[j]);
        }
        const float4 g = *reinterpret_cast<const float4*>(gamma + col0), e = *reinterpret_cast<const float4*>(beta + col0);
        float2 y0 = __ffma2_rn(__ffma2_rn(make_float2(zr[r][0], zr[r][1]), rstd2, shift2), make_float2(g.x, g.y), make_float2(e.x, e.y));
        float2 y1 = __ffma2_rn(__ffma2_rn(make_float2(zr[r][2], zr[r][3]), rstd2, shift2), make_float2(g.z, g.w), make_float2(e.z, e.w));
        y0.x = fmaxf(y0.x, 0.f); y0.y = fmaxf(y0.y, 0.f); y1.x = fmaxf(y1.x, 0.f); y1.y = fmaxf(y1.y, 0.f);
        if (!STEM) {
            y0 = __fadd2_rn(make_float2(x[0], x[1]), y0);
            y1 = __fadd2_rn(make_float2(x[2], x[3]), y1);
        }
        x[0] = y0.x; x[1] = y0.y; x[2] = y1.x; x[3] = y1.y;
        if (HEADS) {
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const float4 w = *reinterpret_cast<const float4*>(S.headw + q * HP + col0);
                o2[q] = __ffma2_rn(make_float2(w.x, w.y), y0, o2[q]);
                o2[q] = __ffma2_rn(make_float2(w.z, w.w), y1, o2[q]);
            }
        } else {
            tc::tmem_st4(c.tX + uint32_t(col0), x);
            uint32_t hi[2], lo[2];
            tc::split2_f16(x[0], x[1], hi[0], lo[0]);
            tc::split2_f16(x[2], x[3], hi[1], lo[1]);
            // columns 4 part .. 4 part + 3 of the block: 16-byte unit (part >> 1) ^ sw, byte 8 (part & 1) inside it
            const uint32_t a = c.a_row + uint32_t(blk) * 4096u + ((uint32_t(c.part >> 1) ^ c.sw) << 4) + uint32_t(c.part & 1) * 8u;
            asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(a), "r"(hi[0]), "r"(hi[1]) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(a + SM::PART), "r"(lo[0]), "r"(lo[1]) : "memory");
            warp_arrive(&S.rnd_ready[NBF + r], c.lane);
        }
    }
    if (HEADS) {
#pragma unroll
        for (int q = 0; q < 5; ++q) o[q] = o2[q].x + o2[q].y;
        if (c.part != 0) {
#pragma unroll
            for (int q = 0; q < 5; ++q) S.headp[c.part - 1][c.row][q] = o[q];
        }
        row_sync(c.row >> 5);
        if (c.part == 0) {
#pragma unroll
            for (int q = 0; q < 5; ++q) {
#pragma unroll
                for (int part = 1; part < SPLIT; ++part) o[q] += S.headp[part - 1][c.row][q];
                o[q] += S.headw[5 * HP + q];
            }
        }
    } else {
        tc::tmem_st_wait();
    }
}

// Two control warps, both running warp-uniform code with one elected lane doing the asynchronous issue (so that the
// descriptors live in uniform registers and a tcgen05.mma is one predicated instruction, not a per-lane waterfall loop):
//   issuer   -- the MMAs of every stage of every step of this CTA's tiles.  One instruction stream paces the tensor pipe
//               (a k-block is 3 MMAs = 312 tensor-pipe clocks), so it carries its ring position, barrier parities and
//               descriptors incrementally and does nothing else.  (History, ncu in profiles/: with the weight refill and
//               64-bit block counters in the same thread a block took ~1000 clocks to issue and the pipe idled 80 %.)
//   producer -- the weight ring: block k of the launch-wide sequence (period L * NB) goes to slot k % RING once the MMAs
//               that read the slot's previous block have completed (w_empty).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

template <int HP>
__device__ __forceinline__ void producer(Smem<HP>& S, const RolloutParams& p, uint32_t steps_total, const uint8_t* img) {
    using SM = Smem<HP>;
    constexpr int RING = SM::RING;
    const uint32_t per_step = uint32_t(p.layers) * SM::NB;
    const uint8_t* const layers = img + SM::SLOT;            // block 0 of the image is the stem
    if (steps_total == 0) return;
    if (elect_one()) {
        tc::mbar_expect_tx(&S.stem_full, SM::SLOT);
        tc::bulk_g2s(S.Wstem, img, SM::SLOT, &S.stem_full);
    }
    uint32_t slot = 0, par = 1, idx = 0;                     // par: parity of the w_empty phase that frees `slot` (none the first time round)
    const uint8_t* src = layers;
    bool first_lap = true;
    for (uint32_t k = steps_total * per_step; k > 0; --k) {
        if (!first_lap) tc::mbar_wait(&S.w_empty[slot], par);
        if (elect_one()) {
            tc::mbar_expect_tx(&S.w_full[slot], SM::SLOT);
            tc::bulk_g2s(S.W[slot], src, SM::SLOT, &S.w_full[slot]);
        }
        src += SM::SLOT;
        if (++idx == per_step) { idx = 0; src = layers; }
        if (++slot == uint32_t(RING)) { slot = 0; par ^= 1u; first_lap = false; }
    }
}

template <int HP>
__device__ __forceinline__ void issuer(Smem<HP>& S, const RolloutParams& p, uint32_t tmem_base, uint32_t steps_total) {
    using SM = Smem<HP>;
    constexpr int NBF = SM::NBF, NR = SM::NR, RING = SM::RING;
    const int L = p.layers;
    const uint32_t idesc = tc::make_idesc_f16(128, HP);
    // descriptors: the address field (bits 0..13, 16-byte units) is the only part that moves, and it never carries out
    const uint64_t dA = tc::make_desc_sw32(tc::smem_addr(S.A[0]), 16, 256), dW = tc::make_desc_sw32(tc::smem_addr(S.W[0]), 16, 256),
                   dS = tc::make_desc_sw32(tc::smem_addr(S.Wstem), 16, 256);
    const uint32_t dA_lo = uint32_t(dA), dW_lo = uint32_t(dW), d_hi = uint32_t(dA >> 32);
    auto desc = [&](uint32_t lo) { return uint64_t(lo) | uint64_t(d_hi) << 32; };
    if (steps_total == 0) return;
    tc::mbar_wait(&S.stem_full, 0);
    uint32_t slot = 0, full_par = 0, w_off = 0;              // slot of the next block, parity of its w_full phase, slot * (SLOT >> 4)
    auto issue_block = [&](uint32_t a_off, bool first) {     // a_off = (block * 4096) >> 4
        tc::mbar_wait(&S.w_full[slot], full_par);
        if (elect_one()) {
            const uint32_t ah = dA_lo + a_off, al = ah + (SM::PART >> 4), bh = dW_lo + w_off, bl = bh + (SM::WPART >> 4);
            tc::mma_bf16_ss(tmem_base, desc(al), desc(bh), idesc, !first);
            tc::mma_bf16_ss(tmem_base, desc(ah), desc(bl), idesc, true);
            tc::mma_bf16_ss(tmem_base, desc(ah), desc(bh), idesc, true);
            tc::mma_commit(&S.w_empty[slot]);
        }
        w_off += SM::SLOT >> 4;
        if (++slot == uint32_t(RING)) { slot = 0; w_off = 0; full_par ^= 1u; }
    };
    uint32_t in_par = 0, rnd_par = 0;
    for (uint32_t step = 0; step < steps_total; ++step) {
        // stem: the 16 exponents are exact in fp16 (hi part of block 0), so two products against hi | lo of the weights
        tc::mbar_wait(&S.in_ready, in_par);
        in_par ^= 1u;
        tc::fence_after_sync();
        if (elect_one()) {
            tc::mma_bf16_ss(tmem_base, dA, dS + uint64_t(SM::WPART >> 4), idesc, false);
            tc::mma_bf16_ss(tmem_base, dA, dS, idesc, true);
            tc::mma_commit(&S.mma_done);
        }
        for (int l = 0; l < L; ++l) {
#pragma unroll
            for (int i = 0; i < NBF; ++i) {
                tc::mbar_wait(&S.rnd_ready[i], rnd_par);
                tc::fence_after_sync();
#pragma unroll
                for (int part = 0; part < 4; ++part) issue_block(uint32_t(4 * i + part) * 256u, i == 0 && part == 0);
            }
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                tc::mbar_wait(&S.rnd_ready[NBF + r], rnd_par);
                tc::fence_after_sync();
                issue_block(uint32_t(4 * NBF + r) * 256u, NBF == 0 && r == 0);
            }
            if (elect_one()) tc::mma_commit(&S.mma_done);
            rnd_par ^= 1u;
        }
    }
}

template <int HP>
__global__ void __launch_bounds__(THREADS, 1) rollout_mlp_x3_kernel(RolloutParams p) {
    using SM = Smem<HP>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw + ((1024u - (tc::smem_addr(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = p.layers, h = p.hidden;
    // Work units: (tile, horizon segment).  With one segment per tile a CTA plays a tile's whole horizon and C3's 512 tiles are
    // 3.46 waves on 148 SMs: every SM waits for the 68 that play a fourth tile.  With H segments the units are dealt round-robin
    // (unit u = segment u / ntiles of tile u % ntiles -> CTA u % grid), a tile's board state passes from one segment to the next
    // through p.boards / p.alive in global memory behind a per-tile flag, and the makespan is ceil(512 H / 148) / H horizons
    // (3.5 at H = 2).  The unit a segment waits for always sits at an earlier position of some CTA's list: no cycle.
    const int64_t ntiles = (p.B + 127) / 128;
    const int H = p.segs > 0 ? p.segs : 1;
    const int64_t units = ntiles * H;
    const int64_t my_units = units > blockIdx.x ? (units - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    uint32_t my_steps = 0;
    for (int64_t k = 0; k < my_units; ++k) {
        const int seg = int((int64_t(blockIdx.x) + k * gridDim.x) / ntiles);
        my_steps += uint32_t(int64_t(p.T) * (seg + 1) / H - int64_t(p.T) * seg / H);
    }
    const uint8_t* img = reinterpret_cast<const uint8_t*>(p.packed + pk_x3_base(HP, L));

    // ---- one-time setup
    if (warp == 0) tc::tmem_alloc(&S.tmem_base, 512);
    if (tid == 0) {
        tc::mbar_init(&S.in_ready, 4);
        tc::mbar_init(&S.mma_done, 1);
        tc::mbar_init(&S.stem_full, 1);
        for (int i = 0; i < SM::ROUNDS; ++i) tc::mbar_init(&S.rnd_ready[i], ENV_THREADS / 32);
        for (int i = 0; i < SM::RING; ++i) {
            tc::mbar_init(&S.w_full[i], 1);
            tc::mbar_init(&S.w_empty[i], 1);
        }
        tc::mbar_fence_init();
    }
    for (int i = tid; i < HP; i += THREADS) {
        S.b0[i] = p.packed[pk_stem_b0(HP) + i];
        S.ln_g[0][i] = p.packed[pk_stem_g(HP) + i];
        S.ln_b[0][i] = p.packed[pk_stem_beta(HP) + i];
        for (int l = 0; l < L; ++l) {
            S.ln_g[l + 1][i] = p.packed[pk_layer(HP, l) + int64_t(HP) * HP + i];
            S.ln_b[l + 1][i] = p.packed[pk_layer(HP, l) + int64_t(HP) * (HP + 1) + i];
        }
    }
    for (int i = tid; i < 5 * HP + 8; i += THREADS) S.headw[i] = p.packed[pk_heads(HP, L) + i];
    for (uint32_t i = tid * 16; i < 2 * SM::PART; i += THREADS * 16) *reinterpret_cast<uint4*>(&S.A[0][0] + i) = make_uint4(0, 0, 0, 0);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = S.tmem_base;

    if (warp == ENV_THREADS / 32) {
        issuer<HP>(S, p, tmem_base, my_steps);
    } else if (warp == ENV_THREADS / 32 + 1) {
        producer<HP>(S, p, my_steps, img);
    } else {
        // ---------------- row = env in tile = TMEM lane; four threads (column parts) per row
        const LutGlobal lut{p.lut};
        const int quarter = warp & 3, half = warp >> 2;
        RowCtx c;
        c.part = half;
        c.lane = lane;
        c.row = quarter * 32 + lane;
        c.tD = tmem_base + (uint32_t(quarter * 32) << 16);
        c.tX = c.tD + X_COL;
        c.a_row = tc::smem_addr(S.A[0]) + uint32_t(c.row) * 32u;
        c.sw = uint32_t(c.row >> 2) & 1u;
        const int row = c.row;
        uint32_t mma_par = 0;                       // parity of the mma_done phase the next stage waits for
        for (int64_t k = 0; k < my_units; ++k) {
            const int64_t u = int64_t(blockIdx.x) + k * gridDim.x, tile = u % ntiles;
            const int seg = int(u / ntiles);
            const int t_begin = int(int64_t(p.T) * seg / H), t_end = int(int64_t(p.T) * (seg + 1) / H);
            const int64_t env = tile * 128 + row;
            const bool owner = half == 0 && env < p.B;
            if (seg > 0) {                                       // the tile's previous segment has left its boards in p.boards
                if (tid == 0) {
                    int v;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p.sched + tile) : "memory");
                    } while (v < seg);
                }
                asm volatile("bar.sync 5, %0;" ::"n"(ENV_THREADS) : "memory");
            }
            Board board = {0u, 0u};
            bool alive = false;
            if (owner) {
                board = make_board(seg > 0 ? __ldcg(p.boards + env) : p.boards[env]);
                alive = p.alive ? (seg > 0 ? __ldcg(p.alive + env) : p.alive[env]) != 0 : true;
            }
            for (int t = t_begin; t < t_end; ++t) {
                const uint64_t ctr = p.ctr0 + uint64_t(t);
                uint32_t lm = 0;
                if (owner) lm = begin_step(p, env, ctr, board, alive);
                if (half == 0) {
                    S.xch[row].board = pack_board(board);            // read by part 1 in the tail of this step
                    // model input: the 16 exponents as fp16 (exact) = the hi part of k-block 0; the row / column
                    // features are folded into the stem bias b0
                    uint32_t w[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t src = q < 4 ? board.lo : board.hi;
                        const __half2 pr = __floats2half2_rn(float((src >> (8 * (q & 3))) & 15u), float((src >> (8 * (q & 3) + 4)) & 15u));
                        w[q] = *reinterpret_cast<const uint32_t*>(&pr);
                    }
                    const uint32_t a0 = c.a_row + (c.sw << 4), a1 = a0 ^ 16u;
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a0), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a1), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
                    warp_arrive(&S.in_ready, lane);
                }
                float o[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
                if (L == 0) {                                        // no MMA shadow to hide in: play the moves up front
                    row_sync(quarter);
                    precompute_moves<HP>(S, p, lut, row, uint32_t(half), env, ctr);
                }
                // ---- stages: s = 0 stem, s = 1..L residual blocks
                for (int s = 0; s <= L; ++s, mma_par ^= 1u) {
                    tc::mbar_wait(&S.mma_done, mma_par);
                    tc::fence_after_sync();
                    if (s == 0) {
                        if (L == 0) epilogue<HP, true, true>(S, c, h, S.ln_g[0], S.ln_b[0], o);
                        else epilogue<HP, true, false>(S, c, h, S.ln_g[0], S.ln_b[0], o);
                        // the board of this step is visible to the whole row since the barrier inside the epilogue; what
                        // is written here is read after the barriers of the later epilogues
                        if (L > 0) precompute_moves<HP>(S, p, lut, row, uint32_t(half), env, ctr);
                    } else if (s == L) {
                        epilogue<HP, false, true>(S, c, h, S.ln_g[s], S.ln_b[s], o);
                    } else {
                        epilogue<HP, false, false>(S, c, h, S.ln_g[s], S.ln_b[s], o);
                    }
                }
                // ---- tail (part 0): sample, take the chosen direction's move, spawn, record (train.py:266-326)
                const int64_t ri = int64_t(t) * p.B + env;
                if (owner && alive) {
                    TailState ts;
                    tail_softmax_sample<G2048_X3_FAST_TAIL>(p, ri, lm, o, S.xch[row].draw[2], ts);
                    const MovePre m = S.pre[row][ts.a];
                    ts.u0 = S.xch[row].draw[0];
                    ts.u1 = S.xch[row].draw[1];
                    ts.moved = make_board(m.moved);
                    ts.points = m.points;
                    ts.valid = (m.meta & 1u) != 0u;
                    ts.ovf = (m.meta & 2u) != 0u;
                    ts.max_tile = int(m.meta >> 8);
                    const Board next = tail_spawn<G2048_X3_FAST_TAIL>(board, ts);
                    tail_record(p, ri, board, ts, make_uint2(S.xch[row].pb[0], S.xch[row].pb[1]), make_uint2(m.pa[0], m.pa[1]));
                    board = next;
                    if (ts.flags & FLAG_DONE) {
                        if (p.auto_reset) board = reset_board(env_draws(p.seed ^ RESET_KEY_TWEAK, p.env0 + uint64_t(env), ctr));
                        else alive = false;
                    }
                } else if (owner) {
                    tail_record_idle(p, ri, board);
                }
            }
            if (owner) {
                p.boards[env] = pack_board(board);
                if (p.alive) p.alive[env] = alive ? 1 : 0;
            }
            if (H > 1) {                                         // publish the segment: boards first, then the flag
                __threadfence();
                asm volatile("bar.sync 5, %0;" ::"n"(ENV_THREADS) : "memory");
                if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.sched + tile), "r"(seg + 1) : "memory");
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

template <int HP>
static int launch(const RolloutParams& p_in, cudaStream_t st) {
    RolloutParams p = p_in;
    const int smem = int(sizeof(Smem<HP>)) + 1024;
    auto kern = rollout_mlp_x3_kernel<HP>;
    G2048_CHECK_CUDA(ensure_smem(kern, smem));
    const int64_t ntiles = (p.B + 127) / 128;
    const int grid = int(ntiles < num_sms() ? ntiles : num_sms());
    // horizon segments (see the kernel): only where they shorten the makespan, where there is a place to hand the env state
    // over (a game that ended must stay ended: without auto-reset that needs the caller's alive array) and a flag array
    p.segs = 1;
    if (p.sched && ntiles > grid && p.T >= 64 && (p.auto_reset || p.alive)) {
        const double one = double((ntiles + grid - 1) / grid);
        double best = one;
        for (int hseg = 2; hseg <= 4; ++hseg) {
            const double m = double((ntiles * hseg + grid - 1) / grid) / hseg;
            if (m < best * 0.97) {
                best = m;
                p.segs = hseg;
            }
        }
    }
    if (p.segs > 1) G2048_CHECK_CUDA(cudaMemsetAsync(p.sched, 0, size_t(ntiles) * sizeof(int32_t), st));
    const int64_t units_per_cta = (ntiles * p.segs + grid - 1) / grid;
    if (units_per_cta * (p.T / p.segs + 1) * (int64_t(p.layers) * (HP / 16) + 1) >= (int64_t(1) << 31))
        return fail(G2048_ESHAPE, "g2048_rollout_mlp: B * T too large for one launch of the tensor-core kernel; split the horizon");
    kern<<<grid, THREADS, smem, st>>>(p);
    G2048_CHECK_LAUNCH("rollout_mlp_x3_kernel");
    return G2048_OK;
}

}  // namespace x3

int launch_rollout_x3(const RolloutParams& p, int HP, cudaStream_t st) {
    if (p.layers > x3::MAX_LAYERS)
        return fail(G2048_ESHAPE, "g2048_rollout_mlp: the split-fp16 tensor-core kernel takes at most %d residual blocks (got %d); use the fp32 kernel",
                    x3::MAX_LAYERS, p.layers);
    switch (HP) {
        case 64: return x3::launch<64>(p, st);
        case 128: return x3::launch<128>(p, st);
        case 192: return x3::launch<192>(p, st);
        case 208: return x3::launch<208>(p, st);
    }
    return fail(G2048_ESHAPE, "g2048_rollout_mlp: no tensor-core kernel for padded hidden %d", HP);
}

}  // namespace g2048
