// g2048_env.cu -- environment kernels (row table, reset, step, expand4, potentials, encode)
// and their C-ABI entry points.  sm_100a only.
//
// Throughput kernels (step / expand4) are persistent: one CTA per SM, the row table staged
// once per CTA into shared memory with bulk async copies (cp.async.bulk + mbarrier), then a
// grid-stride loop with fully coalesced 8-byte board loads/stores.  They are HBM-/issue-bound
// integer kernels; tensor cores are irrelevant here.
#include "g2048_device.cuh"
#include "g2048_host.h"

namespace g2048 {

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Stage the first BYTES of a table into shared memory (<= 32 KiB bulk copies on one mbarrier): issue, then wait.
template <uint32_t BYTES>
__device__ __forceinline__ void stage_lut_issue(uint32_t* slut, const uint32_t* glut, uint64_t* bar) {
    constexpr uint32_t CHUNK = 32768;
    static_assert(BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, BYTES);
#pragma unroll
        for (uint32_t off = 0; off < BYTES; off += CHUNK)
            bulk_copy_g2s(reinterpret_cast<uint8_t*>(slut) + off, reinterpret_cast<const uint8_t*>(glut) + off,
                          BYTES - off < CHUNK ? BYTES - off : CHUNK, bar);
    }
}
__device__ __forceinline__ void stage_lut_wait(uint64_t* bar) {
    __syncthreads();          // barrier init visible to all waiters
    mbar_wait(bar, 0);
}
template <uint32_t BYTES = uint32_t(LUT_SMEM_BYTES)>
__device__ __forceinline__ void stage_lut(uint32_t* slut, const uint32_t* glut, uint64_t* bar) {
    stage_lut_issue<BYTES>(slut, glut, bar);
    stage_lut_wait(bar);
}

// Programmatic dependent launch: the persistent table kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so a CTA may start (and stage its table, which no
// kernel in flight writes: g2048_build_lut returns only once the table is complete) while the previous
// kernel of the stream is still draining.  pdl_wait() returns when that kernel has completed and its writes
// are visible; nothing the caller owns is read or written before it.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------ kernels
__global__ void build_lut_kernel(uint32_t* lut) {
    uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row < uint32_t(LUT_ROWS)) {
        lut[row] = lut_entry_for_row(row);
        lut[MOVE_LUT_OFFSET + move_slot(row)] = move_entry_for_row(row);
    }
    // dense step tables (M: u64 slots, S: u16 slots) behind the two row tables
    uint8_t* dense = reinterpret_cast<uint8_t*>(lut) + DENSE_OFFSET_BYTES;
    if (row < uint32_t(DENSE_M_BYTES / 8)) reinterpret_cast<uint64_t*>(dense)[row] = row < uint32_t(DENSE_M_ROWS) ? dense_m_entry(row) : 0ull;
    if (row < uint32_t(DENSE_S_BYTES / 2))
        reinterpret_cast<uint16_t*>(dense + DENSE_M_BYTES)[row] = row < uint32_t(DENSE_S_ROWS) ? uint16_t(dense_s_entry(row)) : uint16_t(0);
}

__global__ void reset_kernel(uint64_t* boards, int64_t n, const uint32_t* replay, uint64_t seed, uint64_t env0,
                             uint64_t ctr) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    U4 d;
    if (replay) {
        uint4 r = reinterpret_cast<const uint4*>(replay)[i];
        d = {r.x, r.y, r.z, r.w};
    } else {
        d = env_draws(seed, env0 + uint64_t(i), ctr);
    }
    boards[i] = pack_board(reset_board(d));
}

constexpr int STEP_THREADS = 1024;

template <bool SHAPING, bool STAGED>
__device__ __forceinline__ void step_loop(const uint32_t* slut, const uint32_t* __restrict__ glut,
                                          const uint64_t* __restrict__ in,
                                          const uint8_t* __restrict__ actions, uint64_t* __restrict__ out,
                                          int32_t* __restrict__ points, uint8_t* __restrict__ flags,
                                          uint64_t* __restrict__ shaping, int64_t n,
                                          const uint32_t* __restrict__ replay, const PhiloxKeys& seed, uint64_t env0,
                                          uint64_t ctr, uint64_t next_board, uint32_t next_action) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        Board b = make_board(next_board);
        uint32_t a = next_action & 3u;
        if (i + stride < n) {                         // next transition's inputs in flight while this one is computed
            next_board = __ldg(in + i + stride);
            next_action = __ldg(actions + i + stride);
        }
        uint32_t u0, u1;
        if (replay) {
            uint2 r = __ldg(reinterpret_cast<const uint2*>(replay) + i);
            u0 = r.x;
            u1 = r.y;
        } else {
            U4 d = env_draws(seed, env0 + uint64_t(i), ctr);
            u0 = d.x;
            u1 = d.y;
        }
        StepOut o;
        if (STAGED && !has_big_tile(b)) o = env_step<SHAPING>(b, a, u0, u1, LutShared{slut});
        else o = env_step<SHAPING>(b, a, u0, u1, LutGlobal{glut});   // rare: a 4096+ tile on the board
        out[i] = pack_board(o.board);
        points[i] = o.points;
        flags[i] = uint8_t(o.flags);
        if (SHAPING) shaping[i] = uint64_t(o.shape_lo) | uint64_t(o.shape_hi) << 32;
    }
}

// Step (+ every shaping term when SHAPING) on the dense tables (M | S staged whole, 219 KiB; M alone without shaping):
// persistent, one CTA per SM, one transition per thread and iteration.  Boards with a 4096+ tile take the general path through L2.
// REPLAY = the caller passed replayed draws.  A template parameter, not a run-time test: with the (predicated-off) replay load in
// the loop, the first constant load of the Philox round keys reused its destination register and had to wait on the scoreboard that
// the load shares with the NEXT transition's board prefetch -- every iteration then waited out a full HBM latency (ncu source view,
// round 2: 13 % of the kernel's stall samples on that one LDC).
template <bool SHAPING, bool REPLAY>
__global__ void __launch_bounds__(STEP_THREADS, 1)
step_kernel_dense(const uint32_t* __restrict__ glut, const uint64_t* __restrict__ in, const uint8_t* __restrict__ actions,
                  uint64_t* __restrict__ out, int32_t* __restrict__ points, uint8_t* __restrict__ flags,
                  uint64_t* __restrict__ shaping, int64_t n, const uint32_t* __restrict__ replay, const PhiloxKeys seed,
                  uint64_t env0, uint64_t ctr) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    // 32-bit indices (the host splits launches at 2^30 transitions): one IMAD.WIDE per address instead of
    // 64-bit add chains on the integer pipe this kernel is bound by
    const uint32_t stride = gridDim.x * blockDim.x, n32 = uint32_t(n);
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    stage_lut_issue<uint32_t(SHAPING ? DENSE_BYTES : DENSE_M_BYTES)>(      // S is only read for the potentials
        reinterpret_cast<uint32_t*>(smem_raw), reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(glut) + DENSE_OFFSET_BYTES), &bar);
    pdl_wait();
    uint64_t next_board = i < n32 ? __ldg(in + i) : 0ull;                 // in flight during the table staging wait
    uint32_t next_action = i < n32 ? __ldg(actions + i) : 0u;
    stage_lut_wait(&bar);
    const DenseSmem tab{smem_u32(smem_raw), smem_u32(smem_raw) + uint32_t(DENSE_M_BYTES)};
    for (; i < n32; i += stride) {
        const Board b = make_board(next_board);
        const uint32_t a = next_action & 3u;
        const uint32_t nx = i + stride;
        if (nx < n32) {                               // next transition's inputs in flight while this one is computed
            next_board = __ldg(in + nx);
            next_action = __ldg(actions + nx);
        }
        uint32_t u0, u1;
        if (REPLAY) {
            const uint2 r = __ldg(reinterpret_cast<const uint2*>(replay) + i);
            u0 = r.x;
            u1 = r.y;
        } else {
            const U4 d = env_draws(seed, env0 + uint64_t(i), ctr);
            u0 = d.x;
            u1 = d.y;
        }
        StepOut o;
        const uint32_t mx = max_nibble(b);
        if (mx <= 11u) o = env_step_dense<SHAPING>(b, mx, a, u0, u1, tab);
        else o = env_step<SHAPING>(b, a, u0, u1, LutGlobal{glut});        // rare: a 4096+ tile on the board
        out[i] = pack_board(o.board);
        points[i] = o.points;
        flags[i] = uint8_t(o.flags);
        if (SHAPING) shaping[i] = uint64_t(o.shape_lo) | uint64_t(o.shape_hi) << 32;
    }
}

// The C2 form of the step (BASELINE config 2: "1M random boards x 4 moves"): one thread plays all four moves of a board, each with
// its own spawn -- transition (b, m) = move m (0 UP, 1 DOWN, 2 LEFT, 3 RIGHT) on board b with the draws of env id env0 + 4 b + m,
// i.e. exactly what g2048_step returns for the 4 n (board, action) pairs -- and the four transitions share the board's transpose,
// largest exponent, before-move potentials, corner rules and empty count (env_step_dense_m).  Outputs are [n, 4]: one board's four
// records are 32 / 16 / 4 / 32 contiguous bytes, written with 128-bit stores.
#ifndef G2048_STEP4_THREADS
#define G2048_STEP4_THREADS 512
#endif
constexpr int STEP4_THREADS = G2048_STEP4_THREADS;      // 512: 128 registers per thread (four transitions' outputs are live at once)

template <bool SHAPING, bool REPLAY>
__global__ void __launch_bounds__(STEP4_THREADS, 1)
step4_kernel_dense(const uint32_t* __restrict__ glut, const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                   int32_t* __restrict__ points, uint8_t* __restrict__ flags, uint64_t* __restrict__ shaping, int64_t n,
                   const uint32_t* __restrict__ replay, const PhiloxKeys seed, uint64_t env0, uint64_t ctr) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    const uint32_t stride = gridDim.x * blockDim.x, n32 = uint32_t(n);      // (the host splits launches at 2^28 boards)
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    stage_lut_issue<uint32_t(SHAPING ? DENSE_BYTES : DENSE_M_BYTES)>(
        reinterpret_cast<uint32_t*>(smem_raw), reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(glut) + DENSE_OFFSET_BYTES), &bar);
    pdl_wait();
    uint64_t next_board = i < n32 ? __ldg(in + i) : 0ull;
    stage_lut_wait(&bar);
    const DenseSmem tab{smem_u32(smem_raw), smem_u32(smem_raw) + uint32_t(DENSE_M_BYTES)};
    for (; i < n32; i += stride) {
        const Board b = make_board(next_board);
        const uint32_t nx = i + stride;
        if (nx < n32) next_board = __ldg(in + nx);
        uint32_t u0[4], u1[4];
        if (REPLAY) {
            const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(replay) + 2 * size_t(i)), r1 = __ldg(reinterpret_cast<const uint4*>(replay) + 2 * size_t(i) + 1);
            u0[0] = r0.x; u1[0] = r0.y; u0[1] = r0.z; u1[1] = r0.w; u0[2] = r1.x; u1[2] = r1.y; u0[3] = r1.z; u1[3] = r1.w;
        } else {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const U4 d = env_draws(seed, env0 + 4ull * uint64_t(i) + uint64_t(m), ctr);
                u0[m] = d.x;
                u1[m] = d.y;
            }
        }
        StepOut o[4];
        const uint32_t mx = max_nibble(b);
        if (mx <= 11u) {
            const Step4Shared sh = step4_shared<SHAPING>(b, mx, tab);
            o[0] = env_step_dense_m<SHAPING, 0>(sh, u0[0], u1[0], tab);
            o[1] = env_step_dense_m<SHAPING, 1>(sh, u0[1], u1[1], tab);
            o[2] = env_step_dense_m<SHAPING, 2>(sh, u0[2], u1[2], tab);
            o[3] = env_step_dense_m<SHAPING, 3>(sh, u0[3], u1[3], tab);
        } else {                                                       // rare: a 4096+ tile on the board
#pragma unroll                                                         // (unrolled: o[] must stay in registers)
            for (int m = 0; m < 4; ++m) o[m] = env_step<SHAPING>(b, uint32_t(m), u0[m], u1[m], LutGlobal{glut});
        }
        uint4* ob = reinterpret_cast<uint4*>(out + 4 * size_t(i));
        ob[0] = make_uint4(o[0].board.lo, o[0].board.hi, o[1].board.lo, o[1].board.hi);
        ob[1] = make_uint4(o[2].board.lo, o[2].board.hi, o[3].board.lo, o[3].board.hi);
        reinterpret_cast<int4*>(points)[i] = make_int4(o[0].points, o[1].points, o[2].points, o[3].points);
        reinterpret_cast<uint32_t*>(flags)[i] = o[0].flags | o[1].flags << 8 | o[2].flags << 16 | o[3].flags << 24;
        if (SHAPING) {
            uint4* os = reinterpret_cast<uint4*>(shaping + 4 * size_t(i));
            os[0] = make_uint4(o[0].shape_lo, o[0].shape_hi, o[1].shape_lo, o[1].shape_hi);
            os[1] = make_uint4(o[2].shape_lo, o[2].shape_hi, o[3].shape_lo, o[3].shape_hi);
        }
    }
}

template <bool SHAPING>
__global__ void __launch_bounds__(256)
step4_kernel_direct(const uint32_t* __restrict__ glut, const uint64_t* in, uint64_t* out, int32_t* points, uint8_t* flags, uint64_t* shaping,
                    int64_t n, const uint32_t* replay, const PhiloxKeys seed, uint64_t env0, uint64_t ctr) {
    const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;      // one transition per thread
    if (t >= 4 * n) return;
    const Board b = make_board(in[t >> 2]);
    uint32_t u0, u1;
    if (replay) {
        u0 = replay[2 * t];
        u1 = replay[2 * t + 1];
    } else {
        const U4 d = env_draws(seed, env0 + uint64_t(t), ctr);
        u0 = d.x;
        u1 = d.y;
    }
    const StepOut o = env_step<SHAPING>(b, uint32_t(t & 3), u0, u1, LutGlobal{glut});
    out[t] = pack_board(o.board);
    points[t] = o.points;
    flags[t] = uint8_t(o.flags);
    if (SHAPING) shaping[t] = uint64_t(o.shape_lo) | uint64_t(o.shape_hi) << 32;
}

template <bool SHAPING>
__global__ void __launch_bounds__(256)
step_kernel_direct(const uint32_t* __restrict__ glut, const uint64_t* in, const uint8_t* actions, uint64_t* out,
                   int32_t* points, uint8_t* flags, uint64_t* shaping, int64_t n, const uint32_t* replay,
                   const PhiloxKeys seed, uint64_t env0, uint64_t ctr) {
    const int64_t i0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    step_loop<SHAPING, false>(nullptr, glut, in, actions, out, points, flags, shaping, n, replay, seed, env0, ctr,
                              i0 < n ? __ldg(in + i0) : 0ull, i0 < n ? uint32_t(__ldg(actions + i0)) : 0u);
}

struct FourLines {
    Lines up, down, left, right;
};
template <class Lut>
__device__ __forceinline__ FourLines lookup_moves(Board b, const Lut& lut) {
    Board bt = transpose(b);
    FourLines f;
    f.up = lookup_rows(bt, lut);               // UP    = left move of the columns
    f.down = lookup_rows(rev_rows(bt), lut);   // DOWN  = right move of the columns
    f.left = lookup_rows(b, lut);              // LEFT
    f.right = lookup_rows(rev_rows(b), lut);   // RIGHT
    return f;
}

template <bool STAGED>
__device__ __forceinline__ void expand4_loop(const uint32_t* slut, const uint32_t* __restrict__ glut,
                                             const uint64_t* __restrict__ boards,
                                             uint64_t* __restrict__ succ, int32_t* __restrict__ points,
                                             uint8_t* __restrict__ legal, uint8_t* __restrict__ max_tile, int64_t n,
                                             uint64_t prefetched) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        Board b = make_board(prefetched);
        if (i + stride < n) prefetched = __ldg(boards + i + stride);     // next board in flight while this one is expanded
        FourLines f;
        int p[4], mt[4];
        if (STAGED && !has_big_tile(b)) {
            // every cell <= 11: the staged move table gives result, points and created tile directly
            f = lookup_moves(b, MoveLutShared{slut});
            move_stats(f.up, p[0], mt[0]);
            move_stats(f.down, p[1], mt[1]);
            move_stats(f.left, p[2], mt[2]);
            move_stats(f.right, p[3], mt[3]);
        } else {
            f = lookup_moves(b, LutGlobal{glut});
            bool ovf;
            merge_stats(f.up, p[0], mt[0], ovf);
            merge_stats(f.down, p[1], mt[1], ovf);
            merge_stats(f.left, p[2], mt[2], ovf);
            merge_stats(f.right, p[3], mt[3], ovf);
        }
        const Lines &lu = f.up, &ld = f.down, &ll = f.left, &lr = f.right;
        Board s[4] = {transpose(result_of(lu)), transpose(rev_rows(result_of(ld))), result_of(ll),
                      rev_rows(result_of(lr))};
        uint32_t lm = 0;
#pragma unroll
        for (int d = 0; d < 4; ++d) lm |= same(s[d], b) ? 0u : (1u << d);
        // 4 successors = 32 contiguous bytes per board, 16 bytes of points
        // one 256-bit store: the warp writes its 1 KiB of successors with full sectors in one instruction
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(succ + 4 * i), "r"(s[0].lo), "r"(s[0].hi), "r"(s[1].lo),
                     "r"(s[1].hi), "r"(s[2].lo), "r"(s[2].hi), "r"(s[3].lo), "r"(s[3].hi)
                     : "memory");
        *reinterpret_cast<int4*>(points + 4 * i) = make_int4(p[0], p[1], p[2], p[3]);
        legal[i] = uint8_t(lm);
        if (max_tile)
            *reinterpret_cast<uint32_t*>(max_tile + 4 * i) =
                uint32_t(mt[0]) | uint32_t(mt[1]) << 8 | uint32_t(mt[2]) << 16 | uint32_t(mt[3]) << 24;
    }
}

__global__ void __launch_bounds__(STEP_THREADS, 1)
expand4_kernel_staged(const uint32_t* __restrict__ glut, const uint64_t* boards, uint64_t* succ, int32_t* points,
                      uint8_t* legal, uint8_t* max_tile, int64_t n) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    uint32_t* slut = reinterpret_cast<uint32_t*>(smem_raw);
    const int64_t i0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    stage_lut_issue<uint32_t(MOVE_SMEM_BYTES)>(slut, glut + MOVE_LUT_OFFSET, &bar);
    pdl_wait();
    const uint64_t first = i0 < n ? __ldg(boards + i0) : 0ull;            // in flight during the table staging wait
    stage_lut_wait(&bar);
    expand4_loop<true>(slut, glut, boards, succ, points, legal, max_tile, n, first);
}

__global__ void __launch_bounds__(256)
expand4_kernel_direct(const uint32_t* __restrict__ glut, const uint64_t* boards, uint64_t* succ, int32_t* points,
                      uint8_t* legal, uint8_t* max_tile, int64_t n) {
    const int64_t i0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    expand4_loop<false>(nullptr, glut, boards, succ, points, legal, max_tile, n, i0 < n ? __ldg(boards + i0) : 0ull);
}

__global__ void __launch_bounds__(256)
potentials_kernel(const uint32_t* __restrict__ glut, const uint64_t* __restrict__ boards, int32_t* __restrict__ out,
                  int64_t n) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    LutGlobal lut{glut};
    Board b = make_board(boards[i]);
    Potentials p = potentials(b, lookup_rows(b, lut), lookup_rows(transpose(b), lut));
    int32_t* o = out + 6 * i;
    o[0] = p.mono;
    o[1] = p.empt;
    o[2] = -p.smooth_abs;
    o[3] = p.in_corner ? p.max_exp : -p.max_exp;
    o[4] = p.max_exp;
    o[5] = int32_t(legal_mask(b));
}

// one thread per (board, cell): coalesced 12-byte-per-thread stores of [exp, row/3, col/3]
__global__ void __launch_bounds__(256) encode_kernel(const uint64_t* __restrict__ boards, float* __restrict__ out, int64_t n) {
    int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n * 16) return;
    int64_t i = t >> 4;
    int cell = int(t & 15);
    uint64_t b = __ldg(boards + i);
    float* o = out + 3 * t;
    o[0] = float((b >> (4 * cell)) & 15u);
    o[1] = pos_feature(cell >> 2);
    o[2] = pos_feature(cell & 3);
}

// ------------------------------------------------------------------ float potentials (logged only)
// adjacency_bonus game.py:402-442, monotonic_chain_score game.py:445-506, _choose_anchor_corner
// game.py:634-668, topological_score game.py:803-921.  Not on the throughput path (their weights
// never enter the reward, train.py:709-714): one thread per (board, pre-spawn successor) pair, plain
// loops over the 16 cells, doubles with the Python operation order (no FMA contraction) so the
// results are bit-identical to the reference's floats.  The chain score's DFS is replaced by the
// equivalent dynamic programme over exponents (a chain descends by exactly one per step, so
// best[cell] = exp + max best[neighbour with exp-1], filled in increasing exponent order).
struct Cells {
    int v[16];
    __device__ __forceinline__ int at(int r, int c) const { return v[4 * r + c]; }
};
__device__ __forceinline__ Cells unpack_cells(uint64_t b) {
    Cells c;
#pragma unroll
    for (int i = 0; i < 16; ++i) c.v[i] = int((b >> (4 * i)) & 15u);
    return c;
}
__device__ double adjacency_bonus(const Cells& g) {
    int mx = 0, mp = 0;
    for (int i = 0; i < 16; ++i)
        if (g.v[i] > mx) { mx = g.v[i]; mp = i; }
    double bonus = 0.0;
    const int r = mp >> 2, c = mp & 3;
    if (r > 0 && g.at(r - 1, c) > 0) bonus = __dadd_rn(bonus, __dmul_rn(double(g.at(r - 1, c)), 0.5));
    if (r < 3 && g.at(r + 1, c) > 0) bonus = __dadd_rn(bonus, __dmul_rn(double(g.at(r + 1, c)), 0.5));
    if (c > 0 && g.at(r, c - 1) > 0) bonus = __dadd_rn(bonus, __dmul_rn(double(g.at(r, c - 1)), 0.5));
    if (c < 3 && g.at(r, c + 1) > 0) bonus = __dadd_rn(bonus, __dmul_rn(double(g.at(r, c + 1)), 0.5));
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (g.at(i, j) >= 5) {
                if (j < 3 && g.at(i, j + 1) >= 5) bonus = __dadd_rn(bonus, __dmul_rn(double(g.at(i, j) + g.at(i, j + 1)), 0.25));
                if (i < 3 && g.at(i + 1, j) >= 5) bonus = __dadd_rn(bonus, __dmul_rn(double(g.at(i, j) + g.at(i + 1, j)), 0.25));
            }
    return bonus;
}
__device__ double chain_score(const Cells& g) {
    int mx = 0;
    for (int i = 0; i < 16; ++i) mx = max(mx, g.v[i]);
    if (mx == 0) return 0.0;
    int best[16];
    for (int i = 0; i < 16; ++i) best[i] = 0;
    for (int e = 1; e <= mx; ++e)
        for (int i = 0; i < 16; ++i)
            if (g.v[i] == e) {
                const int r = i >> 2, c = i & 3;
                int cont = 0;
                if (r > 0 && g.v[i - 4] == e - 1) cont = max(cont, best[i - 4]);
                if (r < 3 && g.v[i + 4] == e - 1) cont = max(cont, best[i + 4]);
                if (c > 0 && g.v[i - 1] == e - 1) cont = max(cont, best[i - 1]);
                if (c < 3 && g.v[i + 1] == e - 1) cont = max(cont, best[i + 1]);
                best[i] = e + cont;
            }
    int out = 0;
    for (int i = 0; i < 16; ++i)
        if (g.v[i] == mx) out = max(out, best[i]);
    return double(out);
}
__device__ int anchor_corner(const Cells& g) {
    int mx = 0;
    for (int i = 0; i < 16; ++i) mx = max(mx, g.v[i]);
    if (mx == 0) return 0;
    int first = -1;
    for (int i = 0; i < 16; ++i)
        if (g.v[i] == mx) {
            if (first < 0) first = i;
            if (i == 0 || i == 3 || i == 12 || i == 15) return i;       // first max tile that sits in a corner
        }
    const int r = first >> 2, c = first & 3;
    const int corners[4] = {0, 3, 12, 15};
    int bestc = 0, bd = 1 << 30;
    for (int k = 0; k < 4; ++k) {
        const int d = abs((corners[k] >> 2) - r) + abs((corners[k] & 3) - c);
        if (d < bd) { bd = d; bestc = corners[k]; }
    }
    return bestc;
}
__device__ double topological_score(const Cells& g, int anchor) {
    int mx = 0, ntiles = 0;
    for (int i = 0; i < 16; ++i) { mx = max(mx, g.v[i]); ntiles += g.v[i] > 0; }
    if (!ntiles) return 0.0;
    const int cr = anchor >> 2, cc = anchor & 3, rd = cr == 0 ? 1 : -1, cd = cc == 0 ? 1 : -1;
    int order[16], idx_of[16];
    for (int i = 0, n = 0; i < 4; ++i)
        for (int s = 0; s < 4; ++s, ++n) {
            const int row = cr + i * rd, col = (i & 1) ? cc + 3 * cd - s * cd : cc + s * cd;
            order[n] = 4 * row + col;
            idx_of[4 * row + col] = n;
        }
    double score = 0.0;
    for (int i = 0; i < 16; ++i)
        if (g.v[i] > 0) score = __dadd_rn(score, __dmul_rn(double((16 - idx_of[i]) * g.v[i]), 0.1));
    double bonus = 0.0, penalty = 0.0, prev = INFINITY;
    for (int k = 0; k < 16; ++k) {
        const int v = g.v[order[k]];
        if (v == 0) continue;
        if (double(v) <= prev) bonus = __dadd_rn(bonus, __dmul_rn(double(v), 0.2));
        else penalty = __dadd_rn(penalty, __dmul_rn(__dsub_rn(double(v), prev), 0.5));
        prev = double(v);
    }
    score = __dadd_rn(score, __dsub_rn(bonus, penalty));
    if (g.v[anchor] == mx) score = __dadd_rn(score, __dmul_rn(double(mx), 2.0));
    for (int i = 0; i < 16; ++i) {
        const int v = g.v[i];
        if (v < 4) continue;
        const int r = i >> 2, c = i & 3;
        int lower = 0, total = 0;
        if (r > 0 && g.v[i - 4] > 0) { total++; lower += g.v[i - 4] < v - 2; }
        if (r < 3 && g.v[i + 4] > 0) { total++; lower += g.v[i + 4] < v - 2; }
        if (c > 0 && g.v[i - 1] > 0) { total++; lower += g.v[i - 1] < v - 2; }
        if (c < 3 && g.v[i + 1] > 0) { total++; lower += g.v[i + 1] < v - 2; }
        if (total >= 2 && lower >= total - 1 && idx_of[i] > 4) score = __dsub_rn(score, double(v));
    }
    return score;
}
__global__ void __launch_bounds__(128)
potentials_ext_kernel(const uint64_t* __restrict__ before, const uint64_t* __restrict__ after, double* __restrict__ out, int64_t n) {
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Cells a = unpack_cells(before[i]), b = unpack_cells(after[i]);
    const int anchor = anchor_corner(a);
    double* o = out + 7 * i;
    o[0] = adjacency_bonus(a);
    o[1] = adjacency_bonus(b);
    o[2] = chain_score(a);
    o[3] = chain_score(b);
    o[4] = topological_score(a, anchor);
    o[5] = topological_score(b, anchor);
    o[6] = double(anchor);
}

// Symmetry augmentation (train.py:774-881, game.py:508-590) of recorded steps: op 0 = mirror
// horizontal, 1 = mirror vertical, 2/3/4 = rotate 90/180/270 clockwise.  Boards are transformed by
// nibble permutations; the action, the legal mask and the four log-probs move with the direction
// remap new[remap(old)] = old (train.py:784-824).
__device__ __forceinline__ Board mirror_v(Board b) {
    return {__funnelshift_l(b.hi, b.hi, 16), __funnelshift_l(b.lo, b.lo, 16)};
}
__device__ __forceinline__ Board apply_symmetry(Board b, uint32_t op) {
    switch (op) {
        case 0: return rev_rows(b);
        case 1: return mirror_v(b);
        case 2: return rev_rows(transpose(b));      // rotated[j][3-i] = g[i][j]
        case 3: return rev_rows(mirror_v(b));
        default: return mirror_v(transpose(b));     // rotated[3-j][i] = g[i][j]
    }
}
__global__ void __launch_bounds__(256)
augment_kernel(const uint64_t* __restrict__ before, const uint64_t* __restrict__ after, const uint8_t* __restrict__ action,
               const uint8_t* __restrict__ legal, const float4* __restrict__ logp, const uint8_t* __restrict__ op,
               uint64_t* __restrict__ before_out, uint64_t* __restrict__ after_out, uint8_t* __restrict__ action_out,
               uint8_t* __restrict__ legal_out, float4* __restrict__ logp_out, int64_t n) {
    // remap tables, one nibble per old direction (UP, DOWN, LEFT, RIGHT)
    const uint32_t remap_tab[5] = {0x2310u, 0x3201u, 0x1023u, 0x2301u, 0x0132u};
    int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t o = min(uint32_t(op[i]), 4u), tab = remap_tab[o];
    before_out[i] = pack_board(apply_symmetry(make_board(before[i]), o));
    after_out[i] = pack_board(apply_symmetry(make_board(after[i]), o));
    const uint32_t a = action[i] & 3u, lm = legal[i] & 15u;
    action_out[i] = uint8_t((tab >> (4 * a)) & 3u);
    const float4 l = logp[i];
    const float lp[4] = {l.x, l.y, l.z, l.w};
    float out[4];
    uint32_t lm_out = 0;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const uint32_t nd = (tab >> (4 * d)) & 3u;
        lm_out |= ((lm >> d) & 1u) << nd;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (nd == uint32_t(k)) out[k] = lp[d];
    }
    legal_out[i] = uint8_t(lm_out);
    logp_out[i] = make_float4(out[0], out[1], out[2], out[3]);
}

// below this many units the 224 KiB table staging (per CTA) costs more than it saves
constexpr int64_t STAGED_MIN_UNITS = 1 << 17;

// kernel<<<grid, block, smem, stream>>>(args...) with cudaLaunchAttributeProgrammaticStreamSerialization
template <class... P, class... A>
cudaError_t launch_pdl(void (*kernel)(P...), int grid, int block, int smem, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(grid));
    cfg.blockDim = dim3(unsigned(block));
    cfg.dynamicSmemBytes = size_t(smem);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, P(args)...);
}

}  // namespace g2048

using namespace g2048;

extern "C" {

int64_t g2048_lut_bytes(void) { return LUT_BYTES; }

int g2048_build_lut(void* d_lut, void* stream) {
    G2048_REQUIRE(d_lut != nullptr, "g2048_build_lut: d_lut is NULL");
    build_lut_kernel<<<LUT_ROWS / 256, 256, 0, cudaStream_t(stream)>>>(static_cast<uint32_t*>(d_lut));
    G2048_CHECK_LAUNCH("build_lut_kernel");
    // One-time initialisation, the only entry point that blocks: kernels launched with programmatic stream
    // serialisation stage the table before they wait for their predecessor, so it must be complete on return.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    G2048_CHECK_CUDA(cudaStreamIsCapturing(cudaStream_t(stream), &cap));
    G2048_REQUIRE(cap == cudaStreamCaptureStatusNone, "g2048_build_lut: build the table outside stream capture");
    G2048_CHECK_CUDA(cudaStreamSynchronize(cudaStream_t(stream)));
    return G2048_OK;
}

int g2048_reset(uint64_t* boards, int64_t n, const uint32_t* replay, uint64_t seed, uint64_t env0, uint64_t ctr,
                void* stream) {
    G2048_REQUIRE(n >= 0, "g2048_reset: n < 0");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(boards != nullptr, "g2048_reset: boards is NULL");
    reset_kernel<<<unsigned((n + 255) / 256), 256, 0, cudaStream_t(stream)>>>(boards, n, replay, seed, env0, ctr);
    G2048_CHECK_LAUNCH("reset_kernel");
    return G2048_OK;
}

int g2048_step(const void* d_lut, const uint64_t* boards_in, const uint8_t* actions, uint64_t* boards_out,
               int32_t* points, uint8_t* flags, uint64_t* shaping, int64_t n, const uint32_t* replay, uint64_t seed,
               uint64_t env0, uint64_t ctr, void* stream) {
    G2048_REQUIRE(n >= 0, "g2048_step: n < 0");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_lut && boards_in && actions && boards_out && points && flags, "g2048_step: NULL pointer argument");
    const uint32_t* lut = static_cast<const uint32_t*>(d_lut);
    cudaStream_t st = cudaStream_t(stream);
    if (n >= STAGED_MIN_UNITS) {
        auto kern = shaping ? (replay ? step_kernel_dense<true, true> : step_kernel_dense<true, false>)
                            : (replay ? step_kernel_dense<false, true> : step_kernel_dense<false, false>);
        G2048_CHECK_CUDA(ensure_smem(kern, DENSE_BYTES));
        constexpr int64_t SPLIT = int64_t(1) << 30;               // the kernel indexes with 32 bits
        for (int64_t o = 0; o < n; o += SPLIT) {
            const int64_t m = n - o < SPLIT ? n - o : SPLIT;
            G2048_CHECK_CUDA(launch_pdl(kern, num_sms(), STEP_THREADS, DENSE_BYTES, st, lut, boards_in + o, actions + o,
                                        boards_out + o, points + o, flags + o, shaping ? shaping + o : static_cast<uint64_t*>(nullptr), m,
                                        replay ? replay + 2 * o : static_cast<const uint32_t*>(nullptr),
                                        philox_round_keys(seed), env0 + uint64_t(o), ctr));
        }
    } else {
        auto kern = shaping ? step_kernel_direct<true> : step_kernel_direct<false>;
        kern<<<unsigned((n + 255) / 256), 256, 0, st>>>(lut, boards_in, actions, boards_out, points, flags, shaping, n,
                                                        replay, philox_round_keys(seed), env0, ctr);
        G2048_CHECK_LAUNCH("step_kernel_direct");
    }
    return G2048_OK;
}

int g2048_step4(const void* d_lut, const uint64_t* boards_in, uint64_t* boards_out, int32_t* points, uint8_t* flags, uint64_t* shaping,
                int64_t n, const uint32_t* replay, uint64_t seed, uint64_t env0, uint64_t ctr, void* stream) {
    G2048_REQUIRE(n >= 0, "g2048_step4: n < 0");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_lut && boards_in && boards_out && points && flags, "g2048_step4: NULL pointer argument");
    G2048_REQUIRE(((reinterpret_cast<uintptr_t>(boards_out) | reinterpret_cast<uintptr_t>(points) | reinterpret_cast<uintptr_t>(shaping) |
                    reinterpret_cast<uintptr_t>(replay)) & 15) == 0 && (reinterpret_cast<uintptr_t>(flags) & 3) == 0,
                  "g2048_step4: the [n,4] outputs (and replay) must be 16-byte aligned, flags 4-byte aligned");
    const uint32_t* lut = static_cast<const uint32_t*>(d_lut);
    cudaStream_t st = cudaStream_t(stream);
    if (4 * n >= STAGED_MIN_UNITS) {
        auto kern = shaping ? (replay ? step4_kernel_dense<true, true> : step4_kernel_dense<true, false>)
                            : (replay ? step4_kernel_dense<false, true> : step4_kernel_dense<false, false>);
        G2048_CHECK_CUDA(ensure_smem(kern, DENSE_BYTES));
        constexpr int64_t SPLIT = int64_t(1) << 28;               // the kernel indexes with 32 bits
        for (int64_t o = 0; o < n; o += SPLIT) {
            const int64_t m = n - o < SPLIT ? n - o : SPLIT;
            G2048_CHECK_CUDA(launch_pdl(kern, num_sms(), STEP4_THREADS, DENSE_BYTES, st, lut, boards_in + o, boards_out + 4 * o, points + 4 * o,
                                        flags + 4 * o, shaping ? shaping + 4 * o : static_cast<uint64_t*>(nullptr), m,
                                        replay ? replay + 8 * o : static_cast<const uint32_t*>(nullptr), philox_round_keys(seed),
                                        env0 + 4ull * uint64_t(o), ctr));
        }
    } else {
        auto kern = shaping ? step4_kernel_direct<true> : step4_kernel_direct<false>;
        kern<<<unsigned((4 * n + 255) / 256), 256, 0, st>>>(lut, boards_in, boards_out, points, flags, shaping, n, replay,
                                                            philox_round_keys(seed), env0, ctr);
        G2048_CHECK_LAUNCH("step4_kernel_direct");
    }
    return G2048_OK;
}

int g2048_expand4(const void* d_lut, const uint64_t* boards, uint64_t* succ, int32_t* points, uint8_t* legal,
                  uint8_t* max_tile, int64_t n, void* stream) {
    G2048_REQUIRE(n >= 0, "g2048_expand4: n < 0");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_lut && boards && succ && points && legal, "g2048_expand4: NULL pointer argument");
    const uint32_t* lut = static_cast<const uint32_t*>(d_lut);
    cudaStream_t st = cudaStream_t(stream);
    if (n >= STAGED_MIN_UNITS) {
        G2048_CHECK_CUDA(ensure_smem(expand4_kernel_staged, MOVE_SMEM_BYTES));
        G2048_CHECK_CUDA(launch_pdl(expand4_kernel_staged, num_sms(), STEP_THREADS, MOVE_SMEM_BYTES, st, lut, boards, succ, points,
                                    legal, max_tile, n));
    } else {
        expand4_kernel_direct<<<unsigned((n + 255) / 256), 256, 0, st>>>(lut, boards, succ, points, legal, max_tile, n);
        G2048_CHECK_LAUNCH("expand4_kernel_direct");
    }
    return G2048_OK;
}

int g2048_potentials(const void* d_lut, const uint64_t* boards, int32_t* out, int64_t n, void* stream) {
    G2048_REQUIRE(n >= 0, "g2048_potentials: n < 0");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(d_lut && boards && out, "g2048_potentials: NULL pointer argument");
    potentials_kernel<<<unsigned((n + 255) / 256), 256, 0, cudaStream_t(stream)>>>(static_cast<const uint32_t*>(d_lut),
                                                                                  boards, out, n);
    G2048_CHECK_LAUNCH("potentials_kernel");
    return G2048_OK;
}

int g2048_potentials_ext(const uint64_t* before, const uint64_t* after, double* out, int64_t n, void* stream) {
    G2048_REQUIRE(n >= 0, "g2048_potentials_ext: n < 0");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(before && after && out, "g2048_potentials_ext: NULL pointer argument");
    potentials_ext_kernel<<<unsigned((n + 127) / 128), 128, 0, cudaStream_t(stream)>>>(before, after, out, n);
    G2048_CHECK_LAUNCH("potentials_ext_kernel");
    return G2048_OK;
}

int g2048_augment(const uint64_t* before, const uint64_t* after, const uint8_t* action, const uint8_t* legal,
                  const float* logp, const uint8_t* op, uint64_t* before_out, uint64_t* after_out, uint8_t* action_out,
                  uint8_t* legal_out, float* logp_out, int64_t n, void* stream) {
    G2048_REQUIRE(n >= 0, "g2048_augment: n < 0");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(before && after && action && legal && logp && op && before_out && after_out && action_out && legal_out &&
                      logp_out,
                  "g2048_augment: NULL pointer argument");
    augment_kernel<<<unsigned((n + 255) / 256), 256, 0, cudaStream_t(stream)>>>(
        before, after, action, legal, reinterpret_cast<const float4*>(logp), op, before_out, after_out, action_out,
        legal_out, reinterpret_cast<float4*>(logp_out), n);
    G2048_CHECK_LAUNCH("augment_kernel");
    return G2048_OK;
}

int g2048_encode(const uint64_t* boards, float* out, int64_t n, void* stream) {
    G2048_REQUIRE(n >= 0, "g2048_encode: n < 0");
    if (n == 0) return G2048_OK;
    G2048_REQUIRE(boards && out, "g2048_encode: NULL pointer argument");
    int64_t threads = n * 16;
    encode_kernel<<<unsigned((threads + 255) / 256), 256, 0, cudaStream_t(stream)>>>(boards, out, n);
    G2048_CHECK_LAUNCH("encode_kernel");
    return G2048_OK;
}

}  // extern "C"
