// g2048_device.cuh -- device-side core of the 2048 environment for sm_100a.
//
// Board packing: one uint64 per board, cell (r,c) = nibble 4*(4r+c) holding the
// tile exponent (0 = empty).  Kept as two 32-bit halves in registers: `lo` =
// rows 0,1 and `hi` = rows 2,3, because every operation below is 32-bit SASS
// (LOP3 / SHF / PRMT / VABSDIFF4) anyway.
//
// Reference semantics (file:line in the reference checkout):
//   row slide+merge          game.py:224-257      -> row table (see lut_entry_for_row)
//   simulate_move            game.py:121-160      -> move_canonical()
//   legality / terminal      game.py:103-119,259-330 -> legal_mask()
//   potentials               game.py:339-399,671-800 -> potentials()
//   spawn                    game.py:923-940      -> spawn_tile()
//   step                     game.py:952-1030     -> env_step()
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace g2048 {

// ------------------------------------------------------------------ pipe balancing
// The env kernels are bound by the integer ALU pipe (LOP3/SHF/PRMT; ncu: 91 % busy) while the
// FMA pipe idles.  A shift by a constant is also a multiply: x << s = x * 2^s (IMAD) and
// x >> s = umulhi(x, 2^(32-s)) (IMAD.HI), both issued on the FMA pipe.  The powers of two come
// from constant memory so that ptxas cannot turn them back into shifts.
#ifndef G2048_SHIFT_ON_FMA
#define G2048_SHIFT_ON_FMA 0   // measured on B200 (r01): 119.1 us vs 115.1 us per 4.19M transitions -> off
#endif
static __constant__ uint32_t kPow2[33] = {
    1u << 0,  1u << 1,  1u << 2,  1u << 3,  1u << 4,  1u << 5,  1u << 6,  1u << 7,  1u << 8,  1u << 9,  1u << 10,
    1u << 11, 1u << 12, 1u << 13, 1u << 14, 1u << 15, 1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20, 1u << 21,
    1u << 22, 1u << 23, 1u << 24, 1u << 25, 1u << 26, 1u << 27, 1u << 28, 1u << 29, 1u << 30, 1u << 31, 0u};
template <int S>
__device__ __forceinline__ uint32_t shl(uint32_t x) {
#if G2048_SHIFT_ON_FMA
    return x * kPow2[S];
#else
    return x << S;
#endif
}
template <int S>
__device__ __forceinline__ uint32_t shr(uint32_t x) {
#if G2048_SHIFT_ON_FMA
    return __umulhi(x, kPow2[32 - S]);
#else
    return x >> S;
#endif
}
// acc + (x << S) for fields that do not overlap acc (bit-field packing)
template <int S>
__device__ __forceinline__ uint32_t put(uint32_t acc, uint32_t x) {
#if G2048_SHIFT_ON_FMA
    return x * kPow2[S] + acc;
#else
    return acc | (x << S);
#endif
}

// ------------------------------------------------------------------ row table
// One u32 per 16-bit row (cell 0 in the low nibble), describing a LEFT move of
// that row and its per-line potentials:
//   [15:0]  result row (a merge that would create exponent 16 saturates at 15)
//   [19:16] c1 = (exponent created by the first merge) - 1, 0 = no merge
//   [23:20] c2 = same for the second merge
//   [25:24] ge = #adjacent pairs, both non-zero, left >= right   (game.py:714-719)
//   [27:26] le = #adjacent pairs, both non-zero, left <= right
//   [31:28] max exponent in the row
// 65536 entries = 256 KiB in global memory (L2 resident).  Kernels that are
// throughput-bound on it stage the first LUT_SMEM_ROWS entries (224 KiB) in shared
// memory; boards holding a 4096+ tile (has_big_tile) read the table through L2.
constexpr int LUT_ROWS = 65536;
constexpr int LUT_SMEM_ROWS = 0xE000;
constexpr int LUT_SMEM_BYTES = LUT_SMEM_ROWS * 4;
// A second table of the same size follows the first one in the caller's buffer: the "move table"
// used by the 4-move expansion, which needs the merge points directly instead of the potentials:
//   [15:0] left-move result, [27:16] merge points / 4, [31:28] largest exponent created.
// Its fields are exact for rows whose cells are all <= 11 (points <= 8192, created exponent <= 12):
// boards with a bigger tile take the general-table path.  Rows below 0xC000 (192 KiB) are staged.
constexpr int MOVE_LUT_OFFSET = LUT_ROWS;            // in entries
constexpr int MOVE_SMEM_ROWS = 0xC000;
constexpr int MOVE_SMEM_BYTES = MOVE_SMEM_ROWS * 4;
constexpr int ROW_TABLES_BYTES = 2 * LUT_ROWS * 4;
// A third region follows: the dense step tables of the fused step kernel (see "dense step tables" below).
constexpr int DENSE_M_STRIDE = 145;                                 // 12*12 + 1: spreads cells 2,3 over the banks
constexpr int DENSE_M_ROWS = 143 * DENSE_M_STRIDE + 144;            // 20 879
constexpr int DENSE_M_BYTES = ((DENSE_M_ROWS * 8 + 15) / 16) * 16;  // 167 040
constexpr int DENSE_S_ROWS = 13 * 13 * 13 * 13;                     // 28 561
constexpr int DENSE_S_BYTES = ((DENSE_S_ROWS * 2 + 15) / 16) * 16;  // 57 136
constexpr int DENSE_BYTES = DENSE_M_BYTES + DENSE_S_BYTES;          // 224 176 (staged whole)
constexpr int DENSE_OFFSET_BYTES = ROW_TABLES_BYTES;
constexpr int LUT_BYTES = ROW_TABLES_BYTES + DENSE_BYTES;

__host__ __device__ inline uint32_t lut_entry_for_row(uint32_t row) {
    int c[4] = {int(row & 15), int((row >> 4) & 15), int((row >> 8) & 15), int((row >> 12) & 15)};
    int nz[4], n = 0;
    for (int i = 0; i < 4; ++i)
        if (c[i]) nz[n++] = c[i];
    int out[4] = {0, 0, 0, 0}, m = 0, code[2] = {0, 0}, nm = 0;
    for (int i = 0; i < n;) {
        if (i + 1 < n && nz[i] == nz[i + 1]) {
            int ne = nz[i] + 1;
            out[m++] = ne > 15 ? 15 : ne;
            code[nm++] = ne - 1;
            i += 2;
        } else {
            out[m++] = nz[i];
            i += 1;
        }
    }
    int ge = 0, le = 0, mx = 0;
    for (int i = 0; i < 3; ++i)
        if (c[i] && c[i + 1]) {
            ge += c[i] >= c[i + 1];
            le += c[i] <= c[i + 1];
        }
    for (int i = 0; i < 4; ++i) mx = c[i] > mx ? c[i] : mx;
    return uint32_t(out[0]) | uint32_t(out[1]) << 4 | uint32_t(out[2]) << 8 | uint32_t(out[3]) << 12 |
           uint32_t(code[0]) << 16 | uint32_t(code[1]) << 20 | uint32_t(ge) << 24 | uint32_t(le) << 26 |
           uint32_t(mx) << 28;
}

// The move table is stored bank-hashed: row i lives in slot i ^ ((i >> 8) & 31), which folds cells 2-3
// into the five shared-memory bank bits.  Real boards are full of empty cells, so the raw low bits
// (cell 0, cell 1) cluster on a few banks: ncu measured 6.4 wavefronts per lookup unhashed.
__host__ __device__ __forceinline__ uint32_t move_slot(uint32_t row) { return row ^ ((row >> 8) & 31u); }

__host__ __device__ inline uint32_t move_entry_for_row(uint32_t row) {
    const uint32_t e = lut_entry_for_row(row);
    const uint32_t c1 = (e >> 16) & 15u, c2 = (e >> 20) & 15u;
    uint32_t pts = (c1 ? (2u << c1) : 0u) + (c2 ? (2u << c2) : 0u);
    uint32_t mt = c1 > c2 ? c1 : c2;
    mt = mt ? mt + 1u : 0u;
    pts >>= 2;
    if (pts > 0xFFFu) pts = 0xFFFu;        // only for rows outside the table's contract (a cell >= 12)
    if (mt > 15u) mt = 15u;
    return (e & 0xFFFFu) | pts << 16 | mt << 28;
}

// Row-table readers.  LutShared: the staged copy, valid only for rows < LUT_SMEM_ROWS --
// callers guarantee that by sending boards with a cell >= 12 (see has_big_tile) down the
// LutGlobal path; LutGlobal: L2 only (also used where shared memory belongs to something
// else, e.g. the rollout kernel).
struct LutShared {
    const uint32_t* s;
    __device__ __forceinline__ uint32_t operator()(uint32_t row) const { return s[row]; }
};
struct MoveLutShared {     // staged move table, bank-hashed slots
    const uint32_t* s;
    __device__ __forceinline__ uint32_t operator()(uint32_t row) const { return s[move_slot(row)]; }
};
struct LutGlobal {
    const uint32_t* g;
    __device__ __forceinline__ uint32_t operator()(uint32_t row) const { return __ldg(g + row); }
};

// ------------------------------------------------------------------ dense step tables
// The fused step kernel (step + all shaping terms) is bound by the integer ALU pipe, not by HBM, so
// it trades bit arithmetic for table lookups.  Two tables, staged whole into shared memory (219 KiB):
//
//  M (u64, dense index of a row with every cell <= 11: digits c0 + 12 c1 + 145 (c2 + 12 c3)),
//    read for the 4 lines along the move axis (canonical frame):
//      lo [15:0]  left-move result      [27:16] merge points / 4      [31:28] largest exponent created
//      hi [3:0]   ge   [7:4]   le   [15:8]  |smoothness| of the row itself          (before the move)
//         [19:16] ge   [23:20] le   [31:24] |smoothness| of the result row          (after the move)
//  S (u16, dense base-13 index of a row with every cell <= 12), read for the 4 lines across the move
//    axis of the board before the move and of the board after it:
//      [3:0] ge   [7:4] le   [15:8] |smoothness|
//
// Per-line values are ge, le <= 3 and |smoothness| <= 33, so the SUM of four entries never carries
// from one field into the next: a board's pair counts and smoothness are one add chain on whole
// entries.  ge / le as in the row table (game.py:714-719), |smoothness| = sum of |a-b| over adjacent
// cells that are both non-zero (game.py:339-357).
__host__ __device__ inline uint32_t line_stats(uint32_t row) {
    int c[4] = {int(row & 15), int((row >> 4) & 15), int((row >> 8) & 15), int((row >> 12) & 15)};
    uint32_t ge = 0, le = 0, sm = 0;
    for (int i = 0; i < 3; ++i)
        if (c[i] && c[i + 1]) {
            ge += c[i] >= c[i + 1];
            le += c[i] <= c[i + 1];
            sm += uint32_t(c[i] > c[i + 1] ? c[i] - c[i + 1] : c[i + 1] - c[i]);
        }
    return ge | le << 4 | sm << 8;
}
// dense M slot -> 16-bit row (0xFFFFFFFF for the unused slots of the stride padding)
__host__ __device__ inline uint32_t dense_m_row(uint32_t slot) {
    uint32_t lo = slot % uint32_t(DENSE_M_STRIDE), hi = slot / uint32_t(DENSE_M_STRIDE);
    if (lo >= 144u || hi >= 144u) return 0xFFFFFFFFu;
    return (lo % 12u) | (lo / 12u) << 4 | (hi % 12u) << 8 | (hi / 12u) << 12;
}
__host__ __device__ inline uint32_t dense_s_row(uint32_t slot) {
    return (slot % 13u) | ((slot / 13u) % 13u) << 4 | ((slot / 169u) % 13u) << 8 | (slot / 2197u) << 12;
}
__host__ __device__ inline uint64_t dense_m_entry(uint32_t slot) {
    const uint32_t row = dense_m_row(slot);
    if (row == 0xFFFFFFFFu) return 0ull;
    const uint32_t mv = move_entry_for_row(row);              // result | points/4 | created (exact: cells <= 11)
    const uint32_t hi = line_stats(row) | line_stats(mv & 0xFFFFu) << 16;
    return uint64_t(mv) | uint64_t(hi) << 32;
}
__host__ __device__ inline uint32_t dense_s_entry(uint32_t slot) { return line_stats(dense_s_row(slot)); }

// The two rows of a board half -> their two dense indices, in the two 16-bit lanes of the result.
// Stage 1 turns every byte lo + 16 hi into lo + BASE hi (no borrow between bytes: the bytes only
// shrink), stage 2 every 16-bit lane d0 + 256 d1 into d0 + STRIDE d1.
template <uint32_t BASE, uint32_t STRIDE>
__device__ __forceinline__ uint32_t dense2(uint32_t x) {
    const uint32_t t = (x >> 4) & 0x0F0F0F0Fu;
    const uint32_t d = x - (16u - BASE) * t;
    const uint32_t h = (d >> 8) & 0x00FF00FFu;
    return d - (256u - STRIDE) * h;
}


// ------------------------------------------------------------------ board ops
struct Board {
    uint32_t lo, hi;
};

__device__ __forceinline__ Board make_board(uint64_t b) { return {uint32_t(b), uint32_t(b >> 32)}; }
__device__ __forceinline__ uint64_t pack_board(Board b) { return uint64_t(b.lo) | uint64_t(b.hi) << 32; }
__device__ __forceinline__ bool same(Board a, Board b) { return ((a.lo ^ b.lo) | (a.hi ^ b.hi)) == 0; }

// 4x4 nibble transpose: 2x2 nibble blocks inside each half, then 2x2 blocks of bytes.
__device__ __forceinline__ Board transpose(Board b) {
    uint32_t a0 = (b.lo & 0xF0F00F0Fu) | shl<12>(b.lo & 0x0000F0F0u) | shr<12>(b.lo & 0x0F0F0000u);
    uint32_t a1 = (b.hi & 0xF0F00F0Fu) | shl<12>(b.hi & 0x0000F0F0u) | shr<12>(b.hi & 0x0F0F0000u);
    return {__byte_perm(a0, a1, 0x6240), __byte_perm(a0, a1, 0x7351)};
}

// reverse the four nibbles of every 16-bit row
__device__ __forceinline__ uint32_t rev_rows32(uint32_t x) {
    uint32_t y = shl<4>(x & 0x0F0F0F0Fu) | (shr<4>(x) & 0x0F0F0F0Fu);
    return __byte_perm(y, 0, 0x2301);
}
__device__ __forceinline__ Board rev_rows(Board b) { return {rev_rows32(b.lo), rev_rows32(b.hi)}; }

// some cell holds exponent >= 12 (tile 4096+).  Boards without one only produce rows below
// 0xD000 before and after a move (a merge raises an exponent by one), which the staged part
// of the row table covers.
__device__ __forceinline__ bool has_big_tile(Board b) {
    return (((b.lo & shr<1>(b.lo)) | (b.hi & shr<1>(b.hi))) & 0x44444444u) != 0u;
}

// bit 4i set <=> nibble i is non-zero
__device__ __forceinline__ uint32_t nz_flags32(uint32_t x) {
    uint32_t y = x | shr<1>(x);
    return (y | shr<2>(y)) & 0x11111111u;
}
// bit 4i set <=> nibble i is zero
__device__ __forceinline__ uint32_t z_flags32(uint32_t x) { return nz_flags32(x) ^ 0x11111111u; }

// The four line entries of a board (its rows, in order).
struct Lines {
    uint32_t e0, e1, e2, e3;
};
template <class Lut>
__device__ __forceinline__ Lines lookup_rows(Board b, const Lut& lut) {
    return {lut(b.lo & 0xFFFFu), lut(b.lo >> 16), lut(b.hi & 0xFFFFu), lut(b.hi >> 16)};
}
__device__ __forceinline__ Board result_of(Lines l) {
    return {__byte_perm(l.e0, l.e1, 0x5410), __byte_perm(l.e2, l.e3, 0x5410)};
}
// byte k of the result = byte `which` of line k
__device__ __forceinline__ uint32_t gather_byte2(Lines l) {
    return __byte_perm(__byte_perm(l.e0, l.e1, 0x0062), __byte_perm(l.e2, l.e3, 0x0062), 0x5410);
}
__device__ __forceinline__ uint32_t gather_byte3(Lines l) {
    return __byte_perm(__byte_perm(l.e0, l.e1, 0x0073), __byte_perm(l.e2, l.e3, 0x0073), 0x5410);
}

// merge points / largest exponent created / overflow of one move, from the 4 moved lines
__device__ __forceinline__ void merge_stats(Lines l, int& points, int& max_tile, bool& overflow) {
    uint32_t w = gather_byte2(l);  // per line: [c2 | c1]
    uint32_t lo4 = w & 0x0F0F0F0Fu, hi4 = shr<4>(w) & 0x0F0F0F0Fu;
    uint32_t sum = 0, any = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t a = 1u << ((lo4 >> (8 * k)) & 0xFFu);
        uint32_t b = 1u << ((hi4 >> (8 * k)) & 0xFFu);
        sum += (a & ~1u) + (b & ~1u);
        any |= a | b;
    }
    points = int(sum << 1);                       // sum of 2^(c+1) over merges (game.py:237)
    max_tile = any > 1u ? 32 - __clz(any) : 0;    // max (c+1) (game.py:238)
    overflow = (any >> 15) & 1u;                  // a merge created exponent 16
}

// merge points / largest exponent created of one move from 4 move-table entries
__device__ __forceinline__ void move_stats(Lines l, int& points, int& max_tile) {
    const uint32_t t = __byte_perm(l.e0, l.e1, 0x7632) & 0x0FFF0FFFu;   // [points/4 of e0 | of e1] in 16-bit lanes
    const uint32_t u = __byte_perm(l.e2, l.e3, 0x7632) & 0x0FFF0FFFu;
    const uint32_t v = t + u;                                            // <= 4096 per lane: no carry across lanes
    points = int(((v & 0xFFFFu) + (v >> 16)) << 2);
    max_tile = int(max(max(l.e0, l.e1), max(l.e2, l.e3)) >> 28);        // the top nibble dominates the compare
}

// ------------------------------------------------------------------ legality
// bit d set <=> direction d (0=UP 1=DOWN 2=LEFT 3=RIGHT) changes the board, i.e.
// can-slide or can-merge (game.py:116-119).  Slide toward a side <=> some empty cell
// has a tile as its immediate neighbour on the far side; merge <=> equal adjacent tiles.
__device__ __forceinline__ uint32_t nz_flags8(uint32_t x) { return (((x & 0x77777777u) + 0x77777777u) | x) & 0x88888888u; }
__device__ __forceinline__ uint32_t z_flags8(uint32_t x) { return ~(((x & 0x77777777u) + 0x77777777u) | x) & 0x88888888u; }
__device__ __forceinline__ uint32_t legal_mask(Board b) {
    // flags live at bit 3 of every nibble (exact non-zero test in three instructions)
    uint32_t nzl = nz_flags8(b.lo), nzh = nz_flags8(b.hi);
    uint32_t el = nzl ^ 0x88888888u, eh = nzh ^ 0x88888888u;
    // horizontal neighbours (c, c+1), c = 0..2
    uint32_t dl = b.lo ^ (b.lo >> 4), dh = b.hi ^ (b.hi >> 4);
    uint32_t mh = ((z_flags8(dl) & nzl) | (z_flags8(dh) & nzh)) & 0x08880888u;
    // vertical neighbours (r, r+1), r = 0..2
    uint32_t lo16 = __funnelshift_r(b.lo, b.hi, 16);
    uint32_t vl = b.lo ^ lo16, vh = b.hi ^ (b.hi >> 16);
    uint32_t mv = (z_flags8(vl) & nzl) | (z_flags8(vh) & nzh & 0x00008888u);
    uint32_t left = ((el & (nzl >> 4)) | (eh & (nzh >> 4))) & 0x08880888u;
    uint32_t right = ((el & (nzl << 4)) | (eh & (nzh << 4))) & 0x88808880u;
    uint32_t nz16 = __funnelshift_r(nzl, nzh, 16);
    uint32_t up = (el & nz16) | (eh & (nzh >> 16));
    uint32_t down = (eh & nz16) | (el & (nzl << 16));
    uint32_t m = 0;
    if (up | mv) m |= 1u;
    if (down | mv) m |= 2u;
    if (left | mh) m |= 4u;
    if (right | mh) m |= 8u;
    return m;
}

// ------------------------------------------------------------------ potentials
struct Potentials {
    int mono;       // game.py:683-800
    int empt;       // game.py:671-680
    int smooth_abs; // -smoothness_score (game.py:339-357), 0..360
    int max_exp;    // game.py:989
    int in_corner;  // 1 if ANY max tile sits in a corner (game.py:386-399)
};

// sum over the 4 lines of the 2-bit ge / le fields (entry bits 24-25 / 26-27)
__device__ __forceinline__ void ge_le_sums(Lines l, uint32_t& ge, uint32_t& le) {
    uint32_t w = gather_byte3(l);
    ge = __vsadu4(w & 0x03030303u, 0u);
    le = __vsadu4(shr<2>(w) & 0x03030303u, 0u);
}

// -smoothness: sum of |a-b| over the 24 neighbour pairs with both cells non-zero.
// Cells are spread into byte lanes (E = even columns, O = odd columns) so that the
// native 4-way byte abs-diff applies; `nz*` are the nibble non-zero flags of the board.
__device__ __forceinline__ int smoothness_abs(Board b, uint32_t nzl, uint32_t nzh) {
    const uint32_t M = 0x0F0F0F0Fu;
    uint32_t El = b.lo & M, Ol = shr<4>(b.lo) & M, Eh = b.hi & M, Oh = shr<4>(b.hi) & M;
    // 0x0F in every byte lane whose cell is non-zero
    uint32_t mEl = (nzl & 0x01010101u) * 15u, mOl = (shr<4>(nzl) & 0x01010101u) * 15u;
    uint32_t mEh = (nzh & 0x01010101u) * 15u, mOh = (shr<4>(nzh) & 0x01010101u) * 15u;
    // horizontal (c0,c1) and (c2,c3)
    uint32_t acc = (__vabsdiffu4(El, Ol) & mEl & mOl) + (__vabsdiffu4(Eh, Oh) & mEh & mOh);
    // horizontal (c1,c2): O byte k against E byte k+1, valid in byte lanes 0 and 2
    acc += (__vabsdiffu4(Ol, shr<8>(El)) & mOl & shr<8>(mEl) & 0x000F000Fu) +
           (__vabsdiffu4(Oh, shr<8>(Eh)) & mOh & shr<8>(mEh) & 0x000F000Fu);
    // vertical: rows (0,1),(1,2) live in lo vs funnel(lo,hi); rows (2,3) in hi vs hi>>16
    uint32_t Em = __funnelshift_r(El, Eh, 16), Om = __funnelshift_r(Ol, Oh, 16);
    uint32_t mEm = __funnelshift_r(mEl, mEh, 16), mOm = __funnelshift_r(mOl, mOh, 16);
    acc += (__vabsdiffu4(El, Em) & mEl & mEm) + (__vabsdiffu4(Ol, Om) & mOl & mOm);
    acc += (__vabsdiffu4(Eh, shr<16>(Eh)) & mEh & shr<16>(mEh)) + (__vabsdiffu4(Oh, shr<16>(Oh)) & mOh & shr<16>(mOh));
    return int(__vsadu4(acc, 0u));   // <= 8 terms of <= 15 per byte lane: no carry between lanes
}

// `rows`/`cols`: line entries of the board's rows and columns in ANY of the 8 board
// symmetries (pair counts, max and smoothness are symmetry invariant); `b` itself must be
// in the real frame because the first-max-in-row-major-order corner rule is not.
__device__ __forceinline__ Potentials potentials(Board b, Lines rows, Lines cols) {
    Potentials p;
    uint32_t hge, hle, vge, vle;
    ge_le_sums(rows, hge, hle);
    ge_le_sums(cols, vge, vle);
    int pairs = int(max(hge, hle) + max(vge, vle));   // == best of the four rotations
    uint32_t mx = shr<28>(max(max(rows.e0, rows.e1), max(rows.e2, rows.e3)));
    uint32_t nzl = nz_flags32(b.lo), nzh = nz_flags32(b.hi);
    p.empt = 16 - __popc(nzl) - __popc(nzh);
    p.max_exp = int(mx);
    uint32_t rep = mx * 0x11111111u;
    uint32_t ql = z_flags32(b.lo ^ rep), qh = z_flags32(b.hi ^ rep);  // cells equal to the max
    // first max in row-major order = lowest set flag; corners are cells 0,3 (lo) and 12,15 (hi)
    bool first_corner = ql ? ((ql & (0u - ql)) & 0x00001001u) != 0u : ((qh & (0u - qh)) & 0x10010000u) != 0u;
    p.mono = first_corner ? pairs * 2 : pairs / 2;    // game.py:755-758
    p.in_corner = ((ql & 0x00001001u) | (qh & 0x10010000u)) != 0u;
    p.smooth_abs = smoothness_abs(b, nzl, nzh);
    return p;
}

// ------------------------------------------------------------------ Philox4x32-10
struct U4 {
    uint32_t x, y, z, w;
};
__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                            uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0;
        c1 = l1;
        c2 = h0 ^ c3 ^ k1;
        c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return {c0, c1, c2, c3};
}
// The ten round keys of a launch are the same for every thread: kernels that are bound by the integer ALU pipe take
// them precomputed as a by-value kernel parameter (constant bank operands of the XORs) instead of re-deriving them
// per thread (20 integer adds per call).
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
};
__host__ __device__ inline PhiloxKeys philox_round_keys(uint64_t seed) {
    PhiloxKeys rk;
    uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        rk.k0[r] = k0;
        rk.k1[r] = k1;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return rk;
}
__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& rk) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ rk.k0[r];
        c1 = l1;
        c2 = h0 ^ c3 ^ rk.k1[r];
        c3 = l0;
    }
    return {c0, c1, c2, c3};
}
__device__ __forceinline__ U4 env_draws(const PhiloxKeys& rk, uint64_t env_id, uint64_t ctr) {
    return philox4x32_10(uint32_t(env_id), uint32_t(env_id >> 32), uint32_t(ctr), uint32_t(ctr >> 32), rk);
}
// counter = (env_id, ctr), key = seed: one call per env step
__device__ __forceinline__ U4 env_draws(uint64_t seed, uint64_t env_id, uint64_t ctr) {
    return philox4x32_10(uint32_t(env_id), uint32_t(env_id >> 32), uint32_t(ctr), uint32_t(ctr >> 32),
                         uint32_t(seed), uint32_t(seed >> 32));
}

// ------------------------------------------------------------------ spawn
// game.py:923-940.  k = mulhi(u0, #empty) picks the k-th empty cell in row-major order,
// exponent 2 iff u1 >= 3865470567 (<=> not u1/2^32 < 0.9).  The board must have an empty
// cell and at most 15 of them... (reset handles the empty board itself).
__device__ __forceinline__ Board spawn_tile(Board b, uint32_t u0, uint32_t u1) {
    uint32_t zl = z_flags32(b.lo), zh = z_flags32(b.hi);
    uint32_t pl = zl * 0x11111111u;            // nibble i = #empty cells among nibbles 0..i (<= 8)
    uint32_t nl = shr<28>(pl), n = nl + __popc(zh);
    uint32_t k = __umulhi(u0, n);
    bool in_lo = k < nl;
    uint32_t z = in_lo ? zl : zh;
    uint32_t p = in_lo ? pl : zh * 0x11111111u;
    uint32_t t = in_lo ? k + 1u : k + 1u - nl;
    uint32_t hit = z_flags32(p ^ (t * 0x11111111u)) & z;   // the empty nibble whose prefix count is t
    uint32_t tile = (u1 >= 3865470567u ? 2u : 1u) << (__ffs(int(hit)) - 1);
    if (n != 0u) {
        if (in_lo) b.lo |= tile;
        else b.hi |= tile;
    }
    return b;
}

// game.py:942-950 reset: two spawns on the empty board, draws (x,y) then (z,w)
__device__ __forceinline__ Board reset_board(U4 d) {
    uint32_t k1 = __umulhi(d.x, 16u);
    uint32_t v1 = d.y >= 3865470567u ? 2u : 1u;
    uint32_t k2 = __umulhi(d.z, 15u);
    uint32_t v2 = d.w >= 3865470567u ? 2u : 1u;
    k2 += k2 >= k1;
    uint64_t b = uint64_t(v1) << (4 * k1) | uint64_t(v2) << (4 * k2);
    return make_board(b);
}

// ------------------------------------------------------------------ move
// Canonical frame: every move becomes a LEFT move of `canon`:
//   LEFT: board, RIGHT: rows reversed, UP: transposed, DOWN: transposed then rows reversed.
__device__ __forceinline__ Board to_canonical(Board b, Board bt, uint32_t action) {
    Board c = (action & 2u) ? b : bt;             // 2,3 horizontal; 0,1 vertical
    return (action & 1u) ? rev_rows(c) : c;       // DOWN / RIGHT reverse
}
__device__ __forceinline__ Board from_canonical(Board c, uint32_t action) {
    if (action & 1u) c = rev_rows(c);
    return (action & 2u) ? c : transpose(c);
}

// ------------------------------------------------------------------ step
// Packed shaping record (one u64 per transition, all zero for an invalid move):
//  lo: [5:0] mono_before [11:6] mono_after [16:12] empt_before [21:17] empt_after
//      [26:22] max_tile_created [30:27] max_exp_before [31] max-in-corner before
//  hi: [3:0] max_exp_after [4] max-in-corner after [13:5] -smooth_before [22:14] -smooth_after
// "after" = after the move, before the spawn (game.py:994-1002).
struct StepOut {
    Board board;       // post-spawn (unchanged if invalid)
    int points;
    uint32_t flags;    // bits 0-3 legal mask of `board`, 4 done, 5 invalid, 6 overflow
    uint32_t shape_lo, shape_hi;
};

constexpr uint32_t FLAG_DONE = 16u, FLAG_INVALID = 32u, FLAG_OVERFLOW = 64u;

template <bool SHAPING, class Lut>
__device__ __forceinline__ StepOut env_step(Board b, uint32_t action, uint32_t u0, uint32_t u1, const Lut& lut) {
    StepOut o;
    Board bt = transpose(b);
    Board canon = to_canonical(b, bt, action);
    Lines mv = lookup_rows(canon, lut);
    Board moved_c = result_of(mv);
    bool valid = !same(moved_c, canon);            // game.py:959 (legal <=> the move changes the board)
    int points, max_tile;
    bool ovf;
    merge_stats(mv, points, max_tile, ovf);
    Board moved = from_canonical(moved_c, action);
    o.shape_lo = 0u;
    o.shape_hi = 0u;
    if (SHAPING) {
        // before: lines along the move axis come from the move lookups themselves
        Board cross = (action & 2u) ? bt : b;
        Potentials pb = potentials(b, mv, lookup_rows(cross, lut));
        Potentials pa = potentials(moved, lookup_rows(moved_c, lut), lookup_rows(transpose(moved_c), lut));
        uint32_t lo = uint32_t(pb.mono);
        lo = put<6>(lo, uint32_t(pa.mono));
        lo = put<12>(lo, uint32_t(pb.empt));
        lo = put<17>(lo, uint32_t(pa.empt));
        lo = put<22>(lo, uint32_t(max_tile));
        lo = put<27>(lo, uint32_t(pb.max_exp));
        lo = put<31>(lo, uint32_t(pb.in_corner));
        uint32_t hi = uint32_t(pa.max_exp);
        hi = put<4>(hi, uint32_t(pa.in_corner));
        hi = put<5>(hi, uint32_t(pb.smooth_abs));
        hi = put<14>(hi, uint32_t(pa.smooth_abs));
        o.shape_lo = valid ? lo : 0u;
        o.shape_hi = valid ? hi : 0u;
    }
    Board spawned = spawn_tile(moved, u0, u1);     // game.py:1005
    o.board = valid ? spawned : b;
    o.points = valid ? points : 0;
    uint32_t lm = legal_mask(o.board);             // game.py:1006 / 963
    o.flags = lm | (lm == 0u ? FLAG_DONE : 0u) | (valid ? 0u : FLAG_INVALID) | ((valid && ovf) ? FLAG_OVERFLOW : 0u);
    return o;
}

// ------------------------------------------------------------------ step on the dense tables

// largest nibble of the board: a 16-bit unsigned max is decided by the top nibble of each lane, so
// four shifted copies put every cell of a row there once (what lies below does not matter).
__device__ __forceinline__ uint32_t max_nibble(Board b) {
    uint32_t m0 = __vmaxu2(b.lo, b.hi);
    uint32_t m1 = __vmaxu2(b.lo << 4, b.hi << 4);
    uint32_t m2 = __vmaxu2(b.lo << 8, b.hi << 8);
    uint32_t m3 = __vmaxu2(b.lo << 12, b.hi << 12);
    uint32_t m = __vmaxu2(__vmaxu2(m0, m1), __vmaxu2(m2, m3));
    return max(m >> 28, (m >> 12) & 15u);
}

// corner rules for a board whose largest exponent `mx` is known:
// mono doubling <=> the FIRST max cell in row-major order is a corner (game.py:755-758),
// corner bonus sign <=> ANY max cell is a corner (game.py:386-399)
__device__ __forceinline__ void corner_rules(Board b, uint32_t mx, bool& first_corner, bool& in_corner) {
    const uint32_t rep = mx * 0x11111111u;
    const uint32_t ql = z_flags8(b.lo ^ rep), qh = z_flags8(b.hi ^ rep);     // cells equal to the max (bit 3 flags)
    first_corner = ql ? ((ql & (0u - ql)) & 0x00008008u) != 0u : ((qh & (0u - qh)) & 0x80080000u) != 0u;
    in_corner = ((ql & 0x00008008u) | (qh & 0x80080000u)) != 0u;
}

// rev_rows(x) when s4 == 4 and sel == 0x2301, x itself when s4 == 0 and sel == 0x3210: the reversal as straight-line
// code (random actions would make a branch diverge in every warp)
__device__ __forceinline__ uint32_t rev_rows_if(uint32_t x, uint32_t s4, uint32_t sel) {
    const uint32_t y = ((x << s4) & 0xF0F0F0F0u) | ((x >> s4) & 0x0F0F0F0Fu);
    return __byte_perm(y, 0u, sel);
}

// shared-space loads with the table base as the instruction's immediate offset
__device__ __forceinline__ uint2 lds_u64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("{\n\t.reg .u16 t;\n\tld.shared.u16 t, [%1];\n\tcvt.u32.u16 %0, t;\n\t}" : "=r"(v) : "r"(addr));
    return v;
}
struct DenseSmem {
    uint32_t m, s;         // shared-space byte addresses of the two tables
    __device__ __forceinline__ uint2 M(uint32_t byte_off) const { return lds_u64(m + byte_off); }
    __device__ __forceinline__ uint32_t Mlo(uint32_t byte_off) const { return lds_u32(m + byte_off); }
    __device__ __forceinline__ uint32_t S(uint32_t byte_off) const { return lds_u16(s + byte_off); }
};
// the four S entries of a board's rows, summed.  The byte offsets (2 x index <= 57 120) fit the 16-bit lanes.
__device__ __forceinline__ uint32_t s_sum(Board x, const DenseSmem& tab) {
    const uint32_t a = dense2<13, 169>(x.lo), b = dense2<13, 169>(x.hi);
    const uint32_t a2 = a + a, b2 = b + b;
    return tab.S(a2 & 0xFFFFu) + tab.S(a2 >> 16) + tab.S(b2 & 0xFFFFu) + tab.S(b2 >> 16);
}

// env_step<true> for a board with every cell <= 11, i.e. mx_b = max_nibble(b) <= 11 (so every line before the move indexes M / S and
// every line after it indexes S).  12 table reads: the 4 lines along the move axis (M: move result,
// points, created tile, and the per-line potentials of the line before AND after the move), the 4
// lines across it before the move (S) and after it (S).  Same results as env_step, bit for bit.
// SHAPING = false (the caller passed no shaping array): only the 4 M reads, no potentials.
template <bool SHAPING>
__device__ __forceinline__ StepOut env_step_dense(Board b, uint32_t mx_b, uint32_t action, uint32_t u0, uint32_t u1, const DenseSmem& tab) {
    StepOut o;
    const Board bt = transpose(b);
    const bool horiz = (action & 2u) != 0u;
    const uint32_t s4 = (action & 1u) << 2, sel = (action & 1u) ? 0x2301u : 0x3210u;
    const Board along = horiz ? b : bt, cross = horiz ? bt : b;
    const Board canon = {rev_rows_if(along.lo, s4, sel), rev_rows_if(along.hi, s4, sel)};
    const uint32_t ia = dense2<12, DENSE_M_STRIDE>(canon.lo), ib = dense2<12, DENSE_M_STRIDE>(canon.hi);
    uint2 m0, m1, m2, m3;
    if constexpr (SHAPING) {
        m0 = tab.M((ia << 3) & 0x7FFF8u), m1 = tab.M((ia >> 13) & 0x7FFF8u);
        m2 = tab.M((ib << 3) & 0x7FFF8u), m3 = tab.M((ib >> 13) & 0x7FFF8u);
    } else {
        m0 = make_uint2(tab.Mlo((ia << 3) & 0x7FFF8u), 0u), m1 = make_uint2(tab.Mlo((ia >> 13) & 0x7FFF8u), 0u);
        m2 = make_uint2(tab.Mlo((ib << 3) & 0x7FFF8u), 0u), m3 = make_uint2(tab.Mlo((ib >> 13) & 0x7FFF8u), 0u);
    }
    uint32_t cb = 0u;
    if constexpr (SHAPING) cb = s_sum(cross, tab);
    const Board moved_c = {__byte_perm(m0.x, m1.x, 0x5410), __byte_perm(m2.x, m3.x, 0x5410)};
    const bool valid = !same(moved_c, canon);                      // game.py:959
    // merge points and the largest exponent created (move table layout)
    const uint32_t pt = (__byte_perm(m0.x, m1.x, 0x7632) & 0x0FFF0FFFu) + (__byte_perm(m2.x, m3.x, 0x7632) & 0x0FFF0FFFu);
    const uint32_t points = ((pt & 0xFFFFu) + (pt >> 16)) << 2;
    const uint32_t created = max(max(m0.x, m1.x), max(m2.x, m3.x)) >> 28;
    // back to the real frame; the transpose doubles as the source of the cross lines after the move
    const Board un = {rev_rows_if(moved_c.lo, s4, sel), rev_rows_if(moved_c.hi, s4, sel)};
    const Board unt = transpose(un);
    const Board moved = horiz ? un : unt;
    uint32_t shape_lo = 0u, shape_hi = 0u;
    if constexpr (SHAPING) {
        const uint32_t ca = s_sum(unt, tab);
        // potentials of both boards at once: 16-bit lane 0 = before the move, lane 1 = after it
        const uint32_t al = m0.y + m1.y + m2.y + m3.y;                 // lines along the move axis
        const uint32_t cr = ca * 65536u + cb;                          // lines across it
        const uint32_t pairs = __vmaxu2(al & 0x000F000Fu, (al >> 4) & 0x000F000Fu) +
                               __vmaxu2(cr & 0x000F000Fu, (cr >> 4) & 0x000F000Fu);     // SURVEY A7, <= 24 per lane
        const uint32_t smooth = ((al >> 8) & 0x00FF00FFu) + ((cr >> 8) & 0x00FF00FFu);    // <= 264 per lane
        const uint32_t mx_a = max(mx_b, created);                      // a merge only ever raises the maximum
        bool fc_b, ic_b, fc_a, ic_a;
        corner_rules(b, mx_b, fc_b, ic_b);
        corner_rules(moved, mx_a, fc_a, ic_a);
        const uint32_t dbl = pairs + pairs, hlf = (pairs >> 1) & 0x000F000Fu;            // game.py:755-758
        const uint32_t mono_b = (fc_b ? dbl : hlf) & 0x3Fu, mono_a = (fc_a ? dbl : hlf) >> 16;
        const uint32_t smooth_b = smooth & 0xFFFFu, smooth_a = smooth >> 16;
        const uint32_t empt_b = 16u - __popc(nz_flags8(b.lo)) - __popc(nz_flags8(b.hi));
        shape_lo = mono_b | mono_a << 6 | empt_b << 12 | created << 22 | mx_b << 27 | uint32_t(ic_b) << 31;   // + empt_a below
        shape_hi = mx_a | uint32_t(ic_a) << 4 | smooth_b << 5 | smooth_a << 14;
    }
    // spawn (game.py:923-940, as spawn_tile) -- its count of empty cells is emptiness_after
    const uint32_t zl = z_flags8(moved.lo) >> 3, zh = z_flags8(moved.hi) >> 3;
    const uint32_t pl = zl * 0x11111111u;
    const uint32_t nl = pl >> 28, empt_a = nl + __popc(zh);
    const uint32_t k = __umulhi(u0, empt_a);
    const bool in_lo = k < nl;
    const uint32_t z = in_lo ? zl : zh;
    const uint32_t p = in_lo ? pl : zh * 0x11111111u;
    const uint32_t t = in_lo ? k + 1u : k + 1u - nl;
    const uint32_t hit = z_flags8(p ^ (t * 0x11111111u)) & (z << 3);   // the empty nibble whose prefix count is t
    const uint32_t tile = (hit >> 3) * (u1 >= 3865470567u ? 2u : 1u);   // hit is one flag (none on a full board)
    const Board spawned = {moved.lo | (in_lo ? tile : 0u), moved.hi | (in_lo ? 0u : tile)};
    o.shape_lo = (SHAPING && valid) ? (shape_lo | empt_a << 17) : 0u;
    o.shape_hi = (SHAPING && valid) ? shape_hi : 0u;
    o.board = valid ? spawned : b;
    o.points = valid ? int(points) : 0;
    const uint32_t lm = legal_mask(o.board);                       // game.py:1006 / 963
    o.flags = lm | (lm == 0u ? FLAG_DONE : 0u) | (valid ? 0u : FLAG_INVALID);   // no cell >= 12: no overflow possible
    return o;
}

// ------------------------------------------------------------------ the four moves of one board on the dense tables (C2 form)
// What the four transitions of a board share: its transpose, its largest exponent, the corner rules and the count of empty cells
// before the move, and the across-the-move line potentials before the move (the rows of the board for UP / DOWN, its columns for
// LEFT / RIGHT).  env_step_dense_m<.., ACTION> is env_step_dense with the direction as a template parameter (no run-time
// selects, no conditional row reversal) on top of these; same results, bit for bit.
struct Step4Shared {
    Board b, bt;
    uint32_t mx_b, empt_b;
    uint32_t s_rows, s_cols;       // s_sum(b), s_sum(bt): potentials of the board's rows / columns (SHAPING only)
    bool fc_b, ic_b;
};
template <bool SHAPING>
__device__ __forceinline__ Step4Shared step4_shared(Board b, uint32_t mx_b, const DenseSmem& tab) {
    Step4Shared sh;
    sh.b = b;
    sh.bt = transpose(b);
    sh.mx_b = mx_b;
    sh.empt_b = sh.s_rows = sh.s_cols = 0u;
    sh.fc_b = sh.ic_b = false;
    if constexpr (SHAPING) {
        sh.s_rows = s_sum(b, tab);
        sh.s_cols = s_sum(sh.bt, tab);
        corner_rules(b, mx_b, sh.fc_b, sh.ic_b);
        sh.empt_b = 16u - __popc(nz_flags8(b.lo)) - __popc(nz_flags8(b.hi));
    }
    return sh;
}
template <bool SHAPING, int ACTION>
__device__ __forceinline__ StepOut env_step_dense_m(const Step4Shared& sh, uint32_t u0, uint32_t u1, const DenseSmem& tab) {
    constexpr bool horiz = (ACTION & 2) != 0, rev = (ACTION & 1) != 0;
    StepOut o;
    const Board along = horiz ? sh.b : sh.bt;
    const Board canon = rev ? rev_rows(along) : along;
    const uint32_t ia = dense2<12, DENSE_M_STRIDE>(canon.lo), ib = dense2<12, DENSE_M_STRIDE>(canon.hi);
    uint2 m0, m1, m2, m3;
    if constexpr (SHAPING) {
        m0 = tab.M((ia << 3) & 0x7FFF8u), m1 = tab.M((ia >> 13) & 0x7FFF8u);
        m2 = tab.M((ib << 3) & 0x7FFF8u), m3 = tab.M((ib >> 13) & 0x7FFF8u);
    } else {
        m0 = make_uint2(tab.Mlo((ia << 3) & 0x7FFF8u), 0u), m1 = make_uint2(tab.Mlo((ia >> 13) & 0x7FFF8u), 0u);
        m2 = make_uint2(tab.Mlo((ib << 3) & 0x7FFF8u), 0u), m3 = make_uint2(tab.Mlo((ib >> 13) & 0x7FFF8u), 0u);
    }
    const Board moved_c = {__byte_perm(m0.x, m1.x, 0x5410), __byte_perm(m2.x, m3.x, 0x5410)};
    const bool valid = !same(moved_c, canon);                      // game.py:959
    const uint32_t pt = (__byte_perm(m0.x, m1.x, 0x7632) & 0x0FFF0FFFu) + (__byte_perm(m2.x, m3.x, 0x7632) & 0x0FFF0FFFu);
    const uint32_t points = ((pt & 0xFFFFu) + (pt >> 16)) << 2;
    const uint32_t created = max(max(m0.x, m1.x), max(m2.x, m3.x)) >> 28;
    const Board un = rev ? rev_rows(moved_c) : moved_c;
    const Board unt = transpose(un);
    const Board moved = horiz ? un : unt;
    uint32_t shape_lo = 0u, shape_hi = 0u;
    if constexpr (SHAPING) {
        const uint32_t ca = s_sum(unt, tab), cb = horiz ? sh.s_cols : sh.s_rows;
        const uint32_t al = m0.y + m1.y + m2.y + m3.y;                 // lines along the move axis: lane 0 before, lane 1 after
        const uint32_t cr = ca * 65536u + cb;                          // lines across it
        const uint32_t pairs = __vmaxu2(al & 0x000F000Fu, (al >> 4) & 0x000F000Fu) +
                               __vmaxu2(cr & 0x000F000Fu, (cr >> 4) & 0x000F000Fu);
        const uint32_t smooth = ((al >> 8) & 0x00FF00FFu) + ((cr >> 8) & 0x00FF00FFu);
        const uint32_t mx_a = max(sh.mx_b, created);
        bool fc_a, ic_a;
        corner_rules(moved, mx_a, fc_a, ic_a);
        const uint32_t dbl = pairs + pairs, hlf = (pairs >> 1) & 0x000F000Fu;            // game.py:755-758
        const uint32_t mono_b = (sh.fc_b ? dbl : hlf) & 0x3Fu, mono_a = (fc_a ? dbl : hlf) >> 16;
        const uint32_t smooth_b = smooth & 0xFFFFu, smooth_a = smooth >> 16;
        shape_lo = mono_b | mono_a << 6 | sh.empt_b << 12 | created << 22 | sh.mx_b << 27 | uint32_t(sh.ic_b) << 31;
        shape_hi = mx_a | uint32_t(ic_a) << 4 | smooth_b << 5 | smooth_a << 14;
    }
    // spawn (game.py:923-940, as spawn_tile) -- its count of empty cells is emptiness_after
    const uint32_t zl = z_flags8(moved.lo) >> 3, zh = z_flags8(moved.hi) >> 3;
    const uint32_t pl = zl * 0x11111111u;
    const uint32_t nl = pl >> 28, empt_a = nl + __popc(zh);
    const uint32_t k = __umulhi(u0, empt_a);
    const bool in_lo = k < nl;
    const uint32_t z = in_lo ? zl : zh;
    const uint32_t p = in_lo ? pl : zh * 0x11111111u;
    const uint32_t t = in_lo ? k + 1u : k + 1u - nl;
    const uint32_t hit = z_flags8(p ^ (t * 0x11111111u)) & (z << 3);
    const uint32_t tile = (hit >> 3) * (u1 >= 3865470567u ? 2u : 1u);
    const Board spawned = {moved.lo | (in_lo ? tile : 0u), moved.hi | (in_lo ? 0u : tile)};
    o.shape_lo = (SHAPING && valid) ? (shape_lo | empt_a << 17) : 0u;
    o.shape_hi = (SHAPING && valid) ? shape_hi : 0u;
    o.board = valid ? spawned : sh.b;
    o.points = valid ? int(points) : 0;
    const uint32_t lm = legal_mask(o.board);                       // game.py:1006 / 963
    o.flags = lm | (lm == 0u ? FLAG_DONE : 0u) | (valid ? 0u : FLAG_INVALID);
    return o;
}

// ------------------------------------------------------------------ model input
// game.py:92-101: 16 x [exponent, row/3, col/3]
__device__ __forceinline__ float pos_feature(int i) { return float(i) / 3.0f; }

}  // namespace g2048
