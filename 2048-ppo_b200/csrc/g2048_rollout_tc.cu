// g2048_rollout_tc.cu -- tensor-core (tcgen05 / TMEM) variant of the fused rollout kernel, used at
// large env batch.  Same records and env semantics as g2048_rollout.cu; the stem and residual-block
// GEMMs run as bf16 x bf16 -> fp32 tcgen05.mma with the accumulator in tensor memory.
//
// Mapping: a CTA owns tiles of 128 envs; env m of the tile IS TMEM lane m.  Sixteen env warps: warp w
// reads lane quarter (w & 3) -- the hardware restriction of tcgen05.ld -- and column part (w >> 2)
// of TC_SPLIT = 4, walking its 48..56 columns in groups of 8 straight from TMEM, so LayerNorm needs
// one 2-value exchange per row through shared memory and everything else (ReLU, residual, heads,
// masked softmax, sampling, env step) is thread-local.  The epilogue is issue-bound (measured:
// tools/tmem_bw.cu shows TMEM reads at ~900 B/clk/SM are nowhere near the limit), so it uses packed
// fp32 math (FFMA2 / FADD2, two columns per instruction) and shared-space loads throughout.  Env
// thread 0 doubles as the MMA issuer (between its a_ready arrive and its mma_done wait it has
// nothing else to do) and streams the pre-swizzled bf16 weight images L2 -> SMEM with bulk async
// copies.  Hand-off is by mbarriers only:
//   a_ready : 512 env threads -> issuer   (A operand written, TMEM reads of D finished)
//   mma_done: tcgen05.commit  -> env threads + issuer (D complete, B buffer reusable)
//   b_full  : bulk copy tx    -> issuer   (weights of the next block landed)
// TMEM: columns [0,HP) = accumulator D, [256,256+HP) = the fp32 residual stream X.
// Precision: operands are rounded to bf16 (inputs are exact: exponents 0..15), accumulation,
// LayerNorm, residual stream, heads and log-softmax stay fp32; the recorded log-probs differ from
// the fp32 policy by ~1e-2 (stated in tests/test_rollout_gpu.py), which PPO's own ratio absorbs.
#include "g2048_rollout_tail.cuh"
#include "g2048_tc.cuh"

namespace g2048 {

constexpr int TC_SPLIT = 4;              // threads per env row (column quarters)
constexpr int TC_ENV_THREADS = 128 * TC_SPLIT;   // 16 warps: warp w owns lane quarter (w & 3) and column quarter (w >> 2)
constexpr int TC_THREADS = TC_ENV_THREADS;   // env thread 0 doubles as the MMA issuer / weight producer
constexpr uint32_t TC_X_COL = 256;      // TMEM column of the residual stream
constexpr int TC_MAX_LAYERS = 6;        // LayerNorm parameter slots in shared memory (the fp32 kernel takes up to 8 blocks)

template <int HP>
struct TcSmem {
    static constexpr int KB = (HP + 63) / 64;
    alignas(1024) uint8_t A[KB * 128 * 128];        // activations, bf16 swizzled
    alignas(1024) uint8_t Bw[KB * HP * 128];        // current block's weights
    alignas(1024) uint8_t Bstem[HP * 128];          // stem weights (k-step 0 of a 64-wide block)
    alignas(16) float b0[HP];
    alignas(16) float stem_g[HP];
    alignas(16) float stem_b[HP];
    alignas(16) float ln_g[TC_MAX_LAYERS][HP];
    alignas(16) float ln_b[TC_MAX_LAYERS][HP];
    alignas(16) float headw[5 * HP + 8];
    float red[2][TC_SPLIT][128];                    // [sum | sq][column part][row]
    float headp[TC_SPLIT - 1][128][5];             // partial head dots of parts 1..TC_SPLIT-1
    TcXch xch[128];
    uint64_t a_ready, mma_done, b_full, stem_full;
    uint32_t tmem_base;
};
static_assert(sizeof(TcSmem<208>) + 1024 <= 232448, "TcSmem exceeds the 227 KB per-CTA shared memory limit");

__device__ __forceinline__ void env_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_ENV_THREADS) : "memory"); }

// LayerNorm (eps 1e-5) + ReLU (+ residual) of one env row, split over the TC_SPLIT threads that share
// the row (contiguous column ranges, walked in groups of 8 straight from TMEM so the code stays
// compact); result -> residual stream X (TMEM, fp32) and next A operand (SMEM, bf16
// swizzled); optionally the 5 head dot products (partial over this part).  Mean and variance come
// from one pass (sum, sum of squares in fp32): ample for a path whose GEMM operands are bf16.
// game.py:1038-1046, 1069-1073, 1199-1203.
template <int HP, bool STEM, bool HEADS>
__device__ __forceinline__ void epilogue(TcSmem<HP>& S, uint32_t tmem_lane, int row, int half, int h,
                                         const float* __restrict__ bias, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float (&o)[5]) {
    // Columns >= h are padding: their weights, biases, gamma and beta are zero in the packed buffer,
    // so they produce z = 0 and x = 0 without any masking here.
    // this thread's column range: HP/8 groups of 8 columns dealt round-robin-contiguously to the TC_SPLIT parts
    constexpr int G = HP / 8;
    const int g0 = (G * half) / TC_SPLIT, g1 = (G * (half + 1)) / TC_SPLIT, ng = g1 - g0;
    const int c0 = 8 * g0;
    const uint32_t tD = tmem_lane + uint32_t(c0), tX = tmem_lane + TC_X_COL + uint32_t(c0);
    const float inv_h = 1.0f / float(h);
    const float4* bias4 = reinterpret_cast<const float4*>(bias + (STEM ? c0 : 0));
    const float4* gamma4 = reinterpret_cast<const float4*>(gamma + c0);
    const float4* beta4 = reinterpret_cast<const float4*>(beta + c0);
    // packed fp32 math (FADD2 / FFMA2): two columns per instruction
    float2 sum2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int g = 0; g < ng; ++g) {
        float v[8];
        tc::tmem_ld8(tD + uint32_t(8 * g), v);
        if (STEM) {
            const float4 ba = bias4[2 * g], bb = bias4[2 * g + 1];
            v[0] += ba.x; v[1] += ba.y; v[2] += ba.z; v[3] += ba.w;
            v[4] += bb.x; v[5] += bb.y; v[6] += bb.z; v[7] += bb.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 vv = make_float2(v[2 * j], v[2 * j + 1]);
            sum2 = __fadd2_rn(sum2, vv);
            sq2 = __ffma2_rn(vv, vv, sq2);
        }
    }
    S.red[0][half][row] = sum2.x + sum2.y;
    S.red[1][half][row] = sq2.x + sq2.y;
    env_sync();
    float tsum = 0.f, tsq = 0.f;
#pragma unroll
    for (int q = 0; q < TC_SPLIT; ++q) {
        tsum += S.red[0][q][row];
        tsq += S.red[1][q][row];
    }
    const float mean = tsum * inv_h;
    const float var = fmaxf(tsq * inv_h - mean * mean, 0.f);
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    const float2 rstd2 = make_float2(rstd, rstd), shift2 = make_float2(-mean * rstd, -mean * rstd);
    float2 o2[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) o2[q] = make_float2(0.f, 0.f);
    uint8_t* arow = S.A + uint32_t(row >> 3) * 1024u + uint32_t(row & 7) * 128u;
#pragma unroll 1
    for (int g = 0; g < ng; ++g) {
        float v[8], x[8];
        tc::tmem_ld8(tD + uint32_t(8 * g), v);
        if (!STEM) tc::tmem_ld8(tX + uint32_t(8 * g), x);
        if (STEM) {
            const float4 ba = bias4[2 * g], bb = bias4[2 * g + 1];
            v[0] += ba.x; v[1] += ba.y; v[2] += ba.z; v[3] += ba.w;
            v[4] += bb.x; v[5] += bb.y; v[6] += bb.z; v[7] += bb.w;
        }
        const float4 ga = gamma4[2 * g], gb = gamma4[2 * g + 1], ea = beta4[2 * g], eb = beta4[2 * g + 1];
        const float2 gm[4] = {make_float2(ga.x, ga.y), make_float2(ga.z, ga.w), make_float2(gb.x, gb.y), make_float2(gb.z, gb.w)};
        const float2 bt[4] = {make_float2(ea.x, ea.y), make_float2(ea.z, ea.w), make_float2(eb.x, eb.y), make_float2(eb.z, eb.w)};
        float2 xx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 y = __ffma2_rn(__ffma2_rn(make_float2(v[2 * j], v[2 * j + 1]), rstd2, shift2), gm[j], bt[j]);
            y.x = fmaxf(y.x, 0.f);
            y.y = fmaxf(y.y, 0.f);
            xx[j] = STEM ? y : __fadd2_rn(make_float2(x[2 * j], x[2 * j + 1]), y);
            x[2 * j] = xx[j].x;
            x[2 * j + 1] = xx[j].y;
        }
        if (HEADS) {
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const float4* hw = reinterpret_cast<const float4*>(S.headw + q * HP + c0 + 8 * g);
                const float4 ha = hw[0], hb = hw[1];
                o2[q] = __ffma2_rn(make_float2(ha.x, ha.y), xx[0], o2[q]);
                o2[q] = __ffma2_rn(make_float2(ha.z, ha.w), xx[1], o2[q]);
                o2[q] = __ffma2_rn(make_float2(hb.x, hb.y), xx[2], o2[q]);
                o2[q] = __ffma2_rn(make_float2(hb.z, hb.w), xx[3], o2[q]);
            }
        }
        tc::tmem_st8(tX + uint32_t(8 * g), x);
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const __nv_bfloat162 pr = __floats2bfloat162_rn(x[2 * q], x[2 * q + 1]);
            w[q] = *reinterpret_cast<const uint32_t*>(&pr);
        }
        const uint32_t col = uint32_t(c0 + 8 * g), blk = col >> 6, unit = ((col & 63u) >> 3) ^ uint32_t(row & 7);
        *reinterpret_cast<uint4*>(arow + blk * (128u * 128u) + unit * 16u) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (HEADS) {
#pragma unroll
        for (int q = 0; q < 5; ++q) o[q] = o2[q].x + o2[q].y;
    }
    tc::tmem_st_wait();
    if (HEADS) {
        if (half != 0) {
#pragma unroll
            for (int q = 0; q < 5; ++q) S.headp[half - 1][row][q] = o[q];
        }
        env_sync();
        if (half == 0) {
#pragma unroll
            for (int q = 0; q < 5; ++q) {
#pragma unroll
                for (int part = 1; part < TC_SPLIT; ++part) o[q] += S.headp[part - 1][row][q];
                o[q] += S.headw[5 * HP + q];
            }
        }
    }
}


// MMA issue + weight streaming, run by env thread 0 between its a_ready arrive and its mma_done wait.
template <int HP>
struct Issuer {
    uint32_t idesc, a0, bw, bs, layer_bytes;
    const uint8_t* img_layers;
    uint32_t b_loads = 0, b_waits = 0;
    int L;

    __device__ __forceinline__ void load_block(TcSmem<HP>& S, int l) {
        tc::mbar_expect_tx(&S.b_full, layer_bytes);
        tc::bulk_g2s(S.Bw, img_layers + size_t(l) * layer_bytes, layer_bytes, &S.b_full);
        ++b_loads;
    }
    // s = 0: stem (one K=16 step against Bstem), s >= 1: residual block s-1 (HP/16 steps against Bw)
    __device__ __forceinline__ void issue(TcSmem<HP>& S, uint32_t tmem_base, int s, uint64_t st) {
        if (s > 0 && b_waits < b_loads) {          // weights of this block landed?
            tc::mbar_wait(&S.b_full, b_waits & 1u);
            ++b_waits;
        }
        tc::mbar_wait(&S.a_ready, uint32_t(st) & 1u);
        tc::fence_after_sync();
        if (s == 0) {
            tc::mma_bf16_ss(tmem_base, tc::make_desc_sw128(a0), tc::make_desc_sw128(bs), idesc, false);
        } else {
#pragma unroll 1
            for (int ks = 0; ks < HP / tc::UMMA_K; ++ks) {
                const uint32_t blk = uint32_t(ks) >> 2, j = uint32_t(ks) & 3u;
                tc::mma_bf16_ss(tmem_base, tc::make_desc_sw128(a0 + blk * (128u * 128u) + j * 32u),
                                tc::make_desc_sw128(bw + blk * (uint32_t(HP) * 128u) + j * 32u), idesc, ks > 0);
            }
        }
        tc::mma_commit(&S.mma_done);
    }
};

template <int HP>
__global__ void __launch_bounds__(TC_THREADS, 1) rollout_mlp_tc_kernel(RolloutParams p) {
    using SM = TcSmem<HP>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw + ((1024u - (tc::smem_addr(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5;
    const int L = p.layers, h = p.hidden;
    const int64_t ntiles = (p.B + 127) / 128;
    const int64_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint8_t* img = reinterpret_cast<const uint8_t*>(p.packed + pk_img_base(HP, L));
    const uint32_t stem_bytes = uint32_t(img_stem_bytes(HP));

    // ---- one-time setup
    if (warp == 0) tc::tmem_alloc(&S.tmem_base, 512);
    if (tid == 0) {
        tc::mbar_init(&S.a_ready, TC_ENV_THREADS);
        tc::mbar_init(&S.mma_done, 1);
        tc::mbar_init(&S.b_full, 1);
        tc::mbar_init(&S.stem_full, 1);
        tc::mbar_fence_init();
    }
    for (int i = tid; i < HP; i += TC_THREADS) {
        S.b0[i] = p.packed[pk_stem_b0(HP) + i];
        S.stem_g[i] = p.packed[pk_stem_g(HP) + i];
        S.stem_b[i] = p.packed[pk_stem_beta(HP) + i];
        for (int l = 0; l < L; ++l) {
            S.ln_g[l][i] = p.packed[pk_layer(HP, l) + int64_t(HP) * HP + i];
            S.ln_b[l][i] = p.packed[pk_layer(HP, l) + int64_t(HP) * (HP + 1) + i];
        }
    }
    for (int i = tid; i < 5 * HP + 8; i += TC_THREADS) S.headw[i] = p.packed[pk_heads(HP, L) + i];
    for (uint32_t i = tid * 16; i < uint32_t(sizeof(S.A)); i += TC_THREADS * 16)
        *reinterpret_cast<uint4*>(S.A + i) = make_uint4(0, 0, 0, 0);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = S.tmem_base;

    const uint64_t stages_total = uint64_t(my_tiles) * uint64_t(p.T) * uint64_t(1 + L);
    const bool issuer = tid == 0;
    Issuer<HP> iss;
    iss.L = L;
    if (issuer && stages_total > 0) {
        iss.idesc = tc::make_idesc_bf16(128, HP);
        iss.a0 = tc::smem_addr(S.A);
        iss.bw = tc::smem_addr(S.Bw);
        iss.bs = tc::smem_addr(S.Bstem);
        iss.layer_bytes = uint32_t(img_layer_bytes(HP));
        iss.img_layers = img + stem_bytes;
        tc::mbar_expect_tx(&S.stem_full, stem_bytes);
        tc::bulk_g2s(S.Bstem, img, stem_bytes, &S.stem_full);
        if (L > 0) iss.load_block(S, 0);
        tc::mbar_wait(&S.stem_full, 0);
    }

    // ---------------- row = env in tile = TMEM lane; two threads (column halves) per row
    const LutGlobal lut{p.lut};
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + (tid & 31);
    const uint32_t tmem_lane = tmem_base + (uint32_t(quarter * 32) << 16);
    uint8_t* arow = S.A + uint32_t(row >> 3) * 1024u + uint32_t(row & 7) * 128u;
    uint64_t st = 0;
    for (int64_t tl = 0; tl < my_tiles; ++tl) {
        const int64_t env = (int64_t(blockIdx.x) + tl * gridDim.x) * 128 + row;
        const bool owner = half == 0 && env < p.B;
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
            if (half == 0) {
                S.xch[row].board = pack_board(board);            // read by part 1 in the tail of this step
                // model input: 16 exponents as bf16 (exact) = units 0,1 of block 0; row/col features
                // are folded into the stem bias b0
                uint32_t w[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t src = q < 4 ? board.lo : board.hi;
                    const float e0 = float((src >> (8 * (q & 3))) & 15u), e1 = float((src >> (8 * (q & 3) + 4)) & 15u);
                    const __nv_bfloat162 pr = __floats2bfloat162_rn(e0, e1);
                    w[q] = *reinterpret_cast<const uint32_t*>(&pr);
                }
                const uint32_t r7 = uint32_t(row & 7);
                *reinterpret_cast<uint4*>(arow + ((0u ^ r7) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(arow + ((1u ^ r7) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
            }
            float o[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            // ---- stages: s = 0 stem, s = 1..L residual blocks
            for (int s = 0; s <= L; ++s, ++st) {
                tc::fence_async_smem();
                tc::fence_before_sync();
                tc::mbar_arrive(&S.a_ready);
                if (issuer) iss.issue(S, tmem_base, s, st);
                tc::mbar_wait(&S.mma_done, uint32_t(st) & 1u);
                tc::fence_after_sync();
                // the weight buffer is free once this block's MMAs are done: fetch the next block's
                if (issuer && s > 0 && L > 1 && st + 1 < stages_total) iss.load_block(S, s % L);
                if (s == 0) {
                    if (L == 0) epilogue<HP, true, true>(S, tmem_lane, row, half, h, S.b0, S.stem_g, S.stem_b, o);
                    else epilogue<HP, true, false>(S, tmem_lane, row, half, h, S.b0, S.stem_g, S.stem_b, o);
                } else if (s == L) {
                    epilogue<HP, false, true>(S, tmem_lane, row, half, h, nullptr, S.ln_g[s - 1], S.ln_b[s - 1], o);
                } else {
                    epilogue<HP, false, false>(S, tmem_lane, row, half, h, nullptr, S.ln_g[s - 1], S.ln_b[s - 1], o);
                }
            }
            // ---- policy + env-step tail over the row's threads (see tail_* above); the barriers are taken by all 512
            const int64_t ri = int64_t(t) * p.B + env;
            TailState ts;
            const bool act = owner && alive;
            if (half == 1) {
                const uint2 pb = board_potentials(make_board(S.xch[row].board), lut);
                S.xch[row].pb[0] = pb.x;
                S.xch[row].pb[1] = pb.y;
            }
            if (act) {
                tail_sample_and_move(p, lut, ri, env, ctr, lm, o, board, ts);
                S.xch[row].moved = pack_board(ts.moved);
            }
            env_sync();
            if (half == 2) {
                const uint2 pa = board_potentials(make_board(S.xch[row].moved), lut);
                S.xch[row].moved = uint64_t(pa.x) | uint64_t(pa.y) << 32;
            }
            Board next = board;
            if (act) next = tail_spawn(board, ts);
            env_sync();
            if (act) {
                const uint64_t paw = S.xch[row].moved;
                tail_record(p, ri, board, ts, make_uint2(S.xch[row].pb[0], S.xch[row].pb[1]), make_uint2(uint32_t(paw), uint32_t(paw >> 32)));
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
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

template <int HP>
static int launch_tc(const RolloutParams& p, cudaStream_t st) {
    const int smem = int(sizeof(TcSmem<HP>)) + 1024;
    auto kern = rollout_mlp_tc_kernel<HP>;
    G2048_CHECK_CUDA(ensure_smem(kern, smem));
    const int64_t ntiles = (p.B + 127) / 128;
    const int grid = int(ntiles < num_sms() ? ntiles : num_sms());
    kern<<<grid, TC_THREADS, smem, st>>>(p);
    G2048_CHECK_LAUNCH("rollout_mlp_tc_kernel");
    return G2048_OK;
}

int launch_rollout_tc(const RolloutParams& p, int HP, cudaStream_t st) {
    if (p.layers > TC_MAX_LAYERS)
        return fail(G2048_ESHAPE, "g2048_rollout_mlp: the tensor-core kernel takes at most %d residual blocks (got %d); use the fp32 kernel",
                    TC_MAX_LAYERS, p.layers);
    switch (HP) {
        case 64: return launch_tc<64>(p, st);
        case 128: return launch_tc<128>(p, st);
        case 192: return launch_tc<192>(p, st);
        case 208: return launch_tc<208>(p, st);
    }
    return fail(G2048_ESHAPE, "g2048_rollout_mlp: no tensor-core kernel for padded hidden %d", HP);
}

}  // namespace g2048
