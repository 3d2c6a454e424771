// g2048_rollout_tail.cuh -- the policy / env-step tail of a rollout step, split over the four threads that share an
// env row in the tensor-core rollout kernels (g2048_rollout_tc.cu: bf16 operands; g2048_rollout_x3.cu: split-fp16
// operands, fp32-grade).  Same records as policy_env_step (g2048_rollout.cuh), bit for bit on the integer fields.
#pragma once
#include "g2048_rollout.cuh"

namespace g2048 {

// Per-env exchange between the four threads of a row during the policy / env-step tail of a step.
struct TcXch {
    uint64_t board;      // the board the step starts from (written by part 0 at the start of the step)
    uint64_t moved;      // part 0 -> part 2: board after the move, before the spawn; then part 2 -> part 0: its potentials
    uint32_t pb[2];      // part 1 -> part 0: potentials of `board`
};

// ------------------------------------------------------------------ policy / env-step tail, split over a row's threads
// policy_env_step (g2048_rollout.cuh) is ~1 000 dependent instructions with three rounds of L2 table reads; run by
// the owner thread alone it kept 12 of the 16 warps idle for a quarter of the kernel (ncu source view).  Here the
// potentials of the current board (part 1) and of the moved board (part 2) run on the row's other threads while
// part 0 samples, moves, spawns and computes the legal mask.  Same records, bit for bit.
__device__ __forceinline__ uint2 pack_potentials(const Potentials& q) {
    return make_uint2(uint32_t(q.mono) | uint32_t(q.empt) << 6 | uint32_t(q.max_exp) << 11 | uint32_t(q.in_corner) << 15,
                      uint32_t(q.smooth_abs));
}
__device__ __forceinline__ uint2 board_potentials(Board b, const LutGlobal& lut) {
    return pack_potentials(potentials(b, lookup_rows(b, lut), lookup_rows(transpose(b), lut)));
}
struct TailState {
    float lp[4], e[4], mx, se, ent, value;
    uint32_t a, u0, u1, lm;
    Board moved;
    int points, max_tile;
    bool valid, ovf;
    uint32_t flags;
};
// part 0, first third: masked log-softmax, sample, move (train.py:266-294 up to the move of game.py:952-1003)
template <bool FAST = true>
__device__ __forceinline__ void tail_sample_and_move(const RolloutParams& p, const LutGlobal& lut, int64_t ri, int64_t env, uint64_t ctr,
                                                     uint32_t lm, const float (&o)[5], Board board, TailState& ts) {
    // FAST: exp / log / divide as the hardware approximations (ex2 / lg2 / rcp, ~1e-6 relative): the bf16 kernel, whose logits
    // already carry bf16 GEMM error (~1e-2), and since its measurement (tools/x3_tail_precision.py: no difference in the
    // distance from torch) the x3 kernel; the fp32 FFMA kernel keeps expf / logf.  Only what the
    // sample needs comes before the move; log-probs and entropy are finished in tail_spawn, off the critical path.
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if ((lm >> j) & 1u) mx = fmaxf(mx, o[j]);
    float e[4], se = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        e[j] = ((lm >> j) & 1u) ? (FAST ? __expf(o[j] - mx) : expf(o[j] - mx)) : 0.f;
        se += e[j];
        ts.lp[j] = o[j];
        ts.e[j] = e[j];
    }
    ts.mx = mx;
    ts.se = se;
    const U4 d = env_draws(p.seed, p.env0 + uint64_t(env), ctr);
    uint32_t a;
    if (p.forced_actions) {
        a = p.forced_actions[ri] & 3u;
    } else {
        const float thr = float(d.z >> 8) * (1.0f / 16777216.0f) * se;
        float cum = 0.f;
        a = 31u - uint32_t(__clz(int(lm)));
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
    ts.a = a;
    ts.u0 = d.x;
    ts.u1 = d.y;
    ts.lm = lm;
    ts.value = o[4];
    const Board bt = transpose(board);
    const Board canon = to_canonical(board, bt, a);
    const Lines mv = lookup_rows(canon, lut);
    const Board moved_c = result_of(mv);
    ts.valid = !same(moved_c, canon);
    merge_stats(mv, ts.points, ts.max_tile, ts.ovf);
    ts.moved = from_canonical(moved_c, a);
}
// masked softmax + categorical sample only (train.py:266-291), for kernels that computed the Philox draws and the moves
// ahead of time: `dz` = word 2 of the step's draws
template <bool FAST = true>
__device__ __forceinline__ void tail_softmax_sample(const RolloutParams& p, int64_t ri, uint32_t lm, const float (&o)[5], uint32_t dz,
                                                    TailState& ts) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if ((lm >> j) & 1u) mx = fmaxf(mx, o[j]);
    float e[4], se = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        e[j] = ((lm >> j) & 1u) ? (FAST ? __expf(o[j] - mx) : expf(o[j] - mx)) : 0.f;
        se += e[j];
        ts.lp[j] = o[j];
        ts.e[j] = e[j];
    }
    ts.mx = mx;
    ts.se = se;
    uint32_t a;
    if (p.forced_actions) {
        a = p.forced_actions[ri] & 3u;
    } else {
        const float thr = float(dz >> 8) * (1.0f / 16777216.0f) * se;
        float cum = 0.f;
        a = 31u - uint32_t(__clz(int(lm)));
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
    ts.a = a;
    ts.lm = lm;
    ts.value = o[4];
}
// part 0, second third: spawn, legal mask, flags (game.py:1005-1006)
template <bool FAST = true>
__device__ __forceinline__ Board tail_spawn(Board board, TailState& ts) {
    // masked log-softmax and entropy (train.py:271-274, 290-291, 326) from the exponentials of tail_sample_and_move
    const float lse = ts.mx + (FAST ? __logf(ts.se) : logf(ts.se)), inv = FAST ? __fdividef(1.0f, ts.se) : 1.0f / ts.se;
    ts.ent = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool legal = (ts.lm >> j) & 1u;
        const float pj = FAST ? ts.e[j] * inv : ts.e[j] / ts.se;
        ts.lp[j] = legal ? ts.lp[j] - lse : -INFINITY;
        if (pj > 0.f) ts.ent -= pj * (FAST ? __logf(pj) : logf(pj));
    }
    const Board spawned = spawn_tile(ts.moved, ts.u0, ts.u1);
    const Board next = ts.valid ? spawned : board;
    const uint32_t lm = legal_mask(next);
    ts.flags = lm | (lm == 0u ? FLAG_DONE : 0u) | (ts.valid ? 0u : FLAG_INVALID) | ((ts.valid && ts.ovf) ? FLAG_OVERFLOW : 0u);
    return next;
}
// part 0, last third: shaping record from the two potential words, the [t, env] record
__device__ __forceinline__ void tail_record(const RolloutParams& p, int64_t ri, Board board, const TailState& ts, uint2 pb, uint2 pa) {
    uint32_t lo = (pb.x & 63u) | (pa.x & 63u) << 6 | ((pb.x >> 6) & 31u) << 12 | ((pa.x >> 6) & 31u) << 17 | uint32_t(ts.max_tile) << 22 |
                  ((pb.x >> 11) & 15u) << 27 | ((pb.x >> 15) & 1u) << 31;
    uint32_t hi = ((pa.x >> 11) & 15u) | ((pa.x >> 15) & 1u) << 4 | pb.y << 5 | pa.y << 14;
    if (!ts.valid) lo = hi = 0u;
    p.rec_boards[ri] = pack_board(board);
    p.rec_actions[ri] = uint8_t(ts.a);
    p.rec_legal[ri] = uint8_t(ts.lm);
    reinterpret_cast<float4*>(p.rec_logp)[ri] = make_float4(ts.lp[0], ts.lp[1], ts.lp[2], ts.lp[3]);
    p.rec_value[ri] = ts.value;
    p.rec_points[ri] = ts.valid ? ts.points : 0;
    p.rec_shaping[ri] = uint64_t(lo) | uint64_t(hi) << 32;
    p.rec_flags[ri] = uint8_t(ts.flags | 0x80u);
    if (p.rec_entropy) p.rec_entropy[ri] = ts.ent;
}
__device__ __forceinline__ void tail_record_idle(const RolloutParams& p, int64_t ri, Board board) {
    p.rec_flags[ri] = 0;
    p.rec_boards[ri] = pack_board(board);
    p.rec_actions[ri] = 0;
    p.rec_legal[ri] = 0;
    p.rec_value[ri] = 0.f;
    p.rec_points[ri] = 0;
    p.rec_shaping[ri] = 0;
    reinterpret_cast<float4*>(p.rec_logp)[ri] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.rec_entropy) p.rec_entropy[ri] = 0.f;
}

}  // namespace g2048
