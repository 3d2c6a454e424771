// g2048_linear.cu -- the Linear-layer GEMMs of the policy update (train.py:497-561: forward,
// loss.backward()) on tcgen05 with fp32-grade results.
//
// The reference trains in fp32 (torch CPU / cuBLAS SGEMM).  Tensor cores take bf16 operands, so every
// fp32 operand is split in two bf16 terms, x = hi + lo with hi = bf16(x), lo = bf16(x - hi), and a
// product is evaluated as  A*B ~= Alo*Bhi + Ahi*Blo + Ahi*Bhi  (three tcgen05.mma per k-step, fp32
// accumulation in tensor memory).  The dropped terms are O(2^-17) of |a||b| per product, i.e. the
// result is within ~1e-6 relative of an fp32 GEMM in practice (tests/test_linear_gpu.py states the bound).
//
// Shapes: activations are [M, F] row-major fp32 with M = samples (millions) and F <= 208 features
// (GameMLP hidden 196, game.py:24-28), weights are [N, K] (torch Linear layout).  Three products:
//   forward   Y [M,N]  = X [M,K]  * W[N,K]^T          x3_gemm_kernel, image of W
//   dgrad     dX[M,K]  = dY[M,N]  * W[N,K]             x3_gemm_kernel, image of W^T
//   wgrad     dW[N,K]  = dY[M,N]^T * X[M,K]            x3_wgrad_kernel (reduction over samples)
//
// Both kernels are persistent (one CTA per SM) and warp-specialised: 8 loader warps stream fp32 rows
// from HBM, split them and write the bf16 hi/lo operands into a shared-memory ring in the 32-byte
// swizzled layout of g2048_tc.cuh; one thread issues the MMAs; 4 warps drain the accumulators from
// tensor memory.  All hand-offs are mbarriers.  In the GEMM the weight image (hi and lo, <= 169 KiB)
// stays resident in shared memory for the whole launch and two accumulators alternate so the
// epilogue of tile i overlaps the MMAs of tile i+1.  The wgrad kernel uses the SAME operand bytes with
// MN-major descriptors (rows = samples = the MMA's K index), accumulates a CTA's whole sample range in
// tensor memory (flushed to an fp32 partial every 8192 samples to bound the accumulation error) and a
// second kernel adds the per-CTA partials in a fixed order (deterministic).
#include <cuda_fp16.h>
#include "g2048_host.h"
#include "g2048_tc.cuh"

namespace g2048 {
namespace lx {

constexpr int MAXF = 208;                      // largest (padded) feature count
constexpr int MAXB = MAXF / 16;                // 16-feature blocks
constexpr int NUM_LOADERS = 256;               // warps 0-7
constexpr int NUM_THREADS = 416;               // + warps 8-11 (epilogue) + warp 12 (issuer)
constexpr int EPI_WARP0 = 8, ISSUER_WARP = 12;

__host__ __device__ constexpr int round16(int x) { return (x + 15) / 16 * 16; }

__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
    hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}
// the same with fp16 terms (22 mantissa bits; for operands of O(1) magnitude: the fused update's activations and its
// loss-scaled gradients)
__device__ __forceinline__ void split4_f16(const float4 v, uint2& hi, uint2& lo) {
    uint32_t h0, l0, h1, l1;
    tc::split2_f16(v.x, v.y, h0, l0);
    tc::split2_f16(v.z, v.w, h1, l1);
    hi = make_uint2(h0, h1);
    lo = make_uint2(l0, l1);
}
// byte offset of float4 `f` (0..3) of row `row` inside a 32-byte-swizzled block
__device__ __forceinline__ uint32_t f4_offset(int row, int f) {
    return uint32_t(row) * 32u + uint32_t((((f >> 1) ^ (row >> 2)) & 1) << 4) + uint32_t(f & 1) * 8u;
}
__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
    return p + ((1024u - (tc::smem_addr(p) & 1023u)) & 1023u);
}

// ------------------------------------------------------------------ weight image
// image = [part hi|lo][k-block][NP rows][32 B]; element (n, k) = transpose ? W[k][n] : W[n][k], zero padded
__global__ void x3_pack_kernel(const float* __restrict__ W, int R, int Ccols, int transpose, uint8_t* __restrict__ img,
                               int NP, int KB) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= NP * KB * 16) return;
    const int n = idx / (KB * 16), k = idx % (KB * 16);
    const int rows = transpose ? Ccols : R, cols = transpose ? R : Ccols;
    float v = 0.f;
    if (n < rows && k < cols) v = transpose ? W[size_t(k) * Ccols + n] : W[size_t(n) * Ccols + k];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const uint32_t off = uint32_t(k >> 4) * uint32_t(NP) * 32u + tc::sw32_offset(n, k & 15);
    *reinterpret_cast<__nv_bfloat16*>(img + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(img + uint32_t(KB) * uint32_t(NP) * 32u + off) = lo;
}

// ------------------------------------------------------------------ C[M,N] = A[M,K] * B[N,K]^T
constexpr int G_SLOTS = 6;                     // ring of k-blocks of A
constexpr uint32_t G_PART = 128 * 32;          // one operand part of a slot: 128 rows x 32 B
constexpr uint32_t G_SLOT = 2 * G_PART;        // hi + lo
constexpr int G_BATCH = 4;                     // k-blocks whose loads a loader thread keeps in flight

struct GemmBars {
    uint64_t full[G_SLOTS], empty[G_SLOTS], d_full[2], d_empty[2], b_full;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
x3_gemm_kernel(const float* __restrict__ A, const uint8_t* __restrict__ img, float* __restrict__ C, int64_t M, int N,
               int K, int NP, int KB) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = align1024(smem_raw);
    const uint32_t b_bytes = 2u * uint32_t(KB) * uint32_t(NP) * 32u;
    uint8_t* sB = base;
    uint8_t* sA = base + ((b_bytes + 1023u) & ~1023u);
    GemmBars& S = *reinterpret_cast<GemmBars*>(sA + G_SLOTS * G_SLOT);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t ntiles = (M + 127) / 128;
    const int my_tiles = ntiles > blockIdx.x ? int((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    const int total_q = my_tiles * KB;

    if (warp == ISSUER_WARP) tc::tmem_alloc(&S.tmem_base, 512);
    if (tid == 0) {
        for (int i = 0; i < G_SLOTS; ++i) {
            tc::mbar_init(&S.full[i], NUM_LOADERS);
            tc::mbar_init(&S.empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&S.d_full[i], 1);
            tc::mbar_init(&S.d_empty[i], 128);
        }
        tc::mbar_init(&S.b_full, 1);
        tc::mbar_fence_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = S.tmem_base;

    if (warp < EPI_WARP0) {
        // ---------------- loaders: fp32 rows -> bf16 hi/lo k-blocks
        const int lt = tid;
        for (int q0 = 0; q0 < total_q; q0 += G_BATCH) {
            float4 v[G_BATCH][2];
#pragma unroll
            for (int b = 0; b < G_BATCH; ++b) {
                const int q = q0 + b;
                const int t = q / KB, j = q - t * KB;
                const int64_t row0 = (int64_t(blockIdx.x) + int64_t(t) * gridDim.x) * 128;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = lt + NUM_LOADERS * e, row = i >> 2, f = i & 3;
                    const int64_t grow = row0 + row;
                    const int col = j * 16 + f * 4;
                    v[b][e] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (q < total_q && grow < M && col < K) v[b][e] = __ldg(reinterpret_cast<const float4*>(A + grow * K + col));
                }
            }
#pragma unroll
            for (int b = 0; b < G_BATCH; ++b) {
                const int q = q0 + b;
                if (q >= total_q) break;
                const int slot = q % G_SLOTS;
                const uint32_t ph = uint32_t(q / G_SLOTS) & 1u;
                tc::mbar_wait(&S.empty[slot], ph ^ 1u);
                uint8_t* dst = sA + uint32_t(slot) * G_SLOT;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = lt + NUM_LOADERS * e, row = i >> 2, f = i & 3;
                    uint2 hi, lo;
                    split4(v[b][e], hi, lo);
                    const uint32_t off = f4_offset(row, f);
                    *reinterpret_cast<uint2*>(dst + off) = hi;
                    *reinterpret_cast<uint2*>(dst + G_PART + off) = lo;
                }
                tc::fence_async_smem();
                tc::mbar_arrive(&S.full[slot]);
            }
        }
    } else if (warp < ISSUER_WARP) {
        // ---------------- epilogue: accumulator -> C rows
        const int ew = warp - EPI_WARP0, row = ew * 32 + lane;
        const uint32_t tlane = tmem_base + (uint32_t(ew * 32) << 16);
        for (int t = 0; t < my_tiles; ++t) {
            const int buf = t & 1;
            tc::mbar_wait(&S.d_full[buf], uint32_t(t >> 1) & 1u);
            tc::fence_after_sync();
            const int64_t grow = (int64_t(blockIdx.x) + int64_t(t) * gridDim.x) * 128 + row;
            float* crow = C + grow * N;
            for (int c = 0; c < NP / 16; ++c) {
                float v[16];
                tc::tmem_ld16p(tlane + uint32_t(buf * 256 + c * 16), v);
                if (grow < M) {
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        if (c * 16 + g * 4 < N)
                            *reinterpret_cast<float4*>(crow + c * 16 + g * 4) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                }
            }
            tc::fence_before_sync();
            tc::mbar_arrive(&S.d_empty[buf]);
        }
    } else if (lane == 0) {
        // ---------------- MMA issuer (+ one-time weight image load)
        if (my_tiles > 0) {
            tc::mbar_expect_tx(&S.b_full, b_bytes);
            for (uint32_t off = 0; off < b_bytes; off += 32768u)
                tc::bulk_g2s(sB + off, img + off, min(32768u, b_bytes - off), &S.b_full);
            tc::mbar_wait(&S.b_full, 0);
        }
        const uint32_t idesc = tc::make_idesc_bf16_major(128, NP, false, false);
        const uint32_t a_addr = tc::smem_addr(sA), b_addr = tc::smem_addr(sB);
        const uint32_t blk = uint32_t(NP) * 32u;
        int q = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int buf = t & 1;
            tc::mbar_wait(&S.d_empty[buf], (uint32_t(t >> 1) & 1u) ^ 1u);
            tc::fence_after_sync();
            const uint32_t d = tmem_base + uint32_t(buf * 256);
            for (int j = 0; j < KB; ++j, ++q) {
                const int slot = q % G_SLOTS;
                tc::mbar_wait(&S.full[slot], uint32_t(q / G_SLOTS) & 1u);
                tc::fence_after_sync();
                const uint32_t a_hi = a_addr + uint32_t(slot) * G_SLOT, a_lo = a_hi + G_PART;
                const uint32_t b_hi = b_addr + uint32_t(j) * blk, b_lo = b_addr + uint32_t(KB + j) * blk;
                tc::mma_bf16_ss(d, tc::make_desc_sw32(a_lo, 16, 256), tc::make_desc_sw32(b_hi, 16, 256), idesc, j > 0);
                tc::mma_bf16_ss(d, tc::make_desc_sw32(a_hi, 16, 256), tc::make_desc_sw32(b_lo, 16, 256), idesc, true);
                tc::mma_bf16_ss(d, tc::make_desc_sw32(a_hi, 16, 256), tc::make_desc_sw32(b_hi, 16, 256), idesc, true);
                tc::mma_commit(&S.empty[slot]);
            }
            tc::mma_commit(&S.d_full[buf]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == ISSUER_WARP) tc::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ dW[N,K] = dY[M,N]^T * X[M,K]
constexpr int W_STAGES = 4;                       // ring of 32-sample stages
constexpr int W_ROWS = 32;                        // samples per stage
constexpr uint32_t W_BLOCK = W_ROWS * 32;         // one 16-feature block of a stage: 1 KiB
constexpr uint32_t W_PART = MAXB * W_BLOCK;       // one operand part (13 blocks)
constexpr uint32_t W_STAGE = 4 * W_PART;          // dY hi | dY lo | X hi | X lo
constexpr int W_FLUSH = 256;                      // stages between accumulator flushes (8192 samples)
constexpr int W_UNITS = 2 * MAXB * (W_ROWS / 8);  // (matrix, block, 8-row group) load units per stage = 104
// Two groups of 8 loader warps take alternate stages: the loads of stage q+1 are in flight while stage q is split
// and stored (one group alone exposes the HBM latency once per stage: measured 51-64 % of the HBM peak).
constexpr int W_GROUPS = 2;
constexpr int W_EPI_WARP0 = 8 * W_GROUPS, W_ISSUER_WARP = W_EPI_WARP0 + 4;
constexpr int W_THREADS = (W_ISSUER_WARP + 1) * 32;      // 672 -> at most 96 registers per thread

struct WgradBars {
    uint64_t full[W_STAGES], empty[W_STAGES], d_full, d_empty;
    uint32_t tmem_base;
};

// DY_IMG / X_IMG: that operand is a bf16 hi|lo image (dy_hp / x_hp = its padded width) instead of row-major fp32.
// X_BOARDS: X is the model input of packed boards (uint64 per sample, K = 48): the features [exponent, row/3, col/3] per
// cell (game.py:92-101) are formed in the loader's registers instead of being materialised by g2048_encode.
template <bool DY_IMG, bool X_IMG, bool X_BOARDS = false>
__global__ void __launch_bounds__(W_THREADS, 1)
x3_wgrad_kernel(const float* __restrict__ dY, const float* __restrict__ X, float* __restrict__ partial, int64_t M, int N,
                int K, int dy_hp, int x_hp, int f16) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sW = align1024(smem_raw);
    WgradBars& S = *reinterpret_cast<WgradBars*>(sW + W_STAGES * W_STAGE);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NPB = (N + 15) / 16, KPB = (K + 15) / 16;
    const int64_t stages = (M + W_ROWS - 1) / W_ROWS;
    const int64_t s_begin = stages * blockIdx.x / gridDim.x, s_end = stages * (blockIdx.x + 1) / gridDim.x;
    const int my_stages = int(s_end - s_begin);

    if (warp == W_ISSUER_WARP) tc::tmem_alloc(&S.tmem_base, 512);
    if (tid == 0) {
        for (int i = 0; i < W_STAGES; ++i) {
            tc::mbar_init(&S.full[i], NUM_LOADERS);
            tc::mbar_init(&S.empty[i], 1);
        }
        tc::mbar_init(&S.d_full, 1);
        tc::mbar_init(&S.d_empty, 128);
        tc::mbar_fence_init();
    }
    // blocks past the real feature count are read by the second M = 128 half: keep them finite
    for (uint32_t i = tid * 16; i < W_STAGES * W_STAGE; i += W_THREADS * 16) *reinterpret_cast<uint4*>(sW + i) = make_uint4(0, 0, 0, 0);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = S.tmem_base;

    if (warp < W_EPI_WARP0) {
        // ---------------- loaders: group (warp / 8) takes the stages q = group, group + W_GROUPS, ...
        // A matrix with hp > 0 is a hi|lo operand image written by update_mlp_x3_kernel (its storer warp): its two
        // parts of a stage are bulk-copied straight into the ring (no registers, no conversion); hp == 0 is
        // row-major fp32, loaded, split and stored by the loader threads.
        constexpr int PER_WARP = W_UNITS / 8;     // 13
        const int wg = warp & 7;
        for (int q = warp >> 3; q < my_stages; q += W_GROUPS) {
            const int64_t sample0 = (s_begin + q) * W_ROWS;
            float4 v[PER_WARP];
#pragma unroll
            for (int i = 0; i < PER_WARP; ++i) {
                const int u = wg + 8 * i, mat = u / (W_UNITS / 2), rem = u % (W_UNITS / 2), rg = rem & 3, b = rem >> 2;
                const int row = rg * 8 + (lane >> 2), f = lane & 3, col = b * 16 + f * 4;
                const int ld = mat ? K : N;
                const bool img = mat ? X_IMG : DY_IMG;
                const float* src = mat ? X : dY;
                const int64_t s = sample0 + row;
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (X_BOARDS && mat == 1) {
                    if (s < M && col < 48) {
                        const uint64_t board = __ldg(reinterpret_cast<const uint64_t*>(X) + s);
                        float e[4];
                        const uint32_t lo32 = uint32_t(board), hi32 = uint32_t(board >> 32);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t idx = uint32_t(col + j), cell = (idx * 43691u) >> 17, kind = idx - 3u * cell;   // idx / 3, idx < 48
                            const uint32_t half = cell & 8u ? hi32 : lo32;
                            const uint32_t pos = kind == 1u ? (cell >> 2) : (cell & 3u);
                            // float(pos) / 3.0f as in g2048_encode, from constants
                            const float pf = pos == 0u ? 0.0f : pos == 1u ? (1.0f / 3.0f) : pos == 2u ? (2.0f / 3.0f) : 1.0f;
                            e[j] = kind == 0u ? float((half >> (4u * (cell & 7u))) & 15u) : pf;
                        }
                        v[i] = make_float4(e[0], e[1], e[2], e[3]);
                    }
                } else if (!img && s < M && col < ld) v[i] = __ldg(reinterpret_cast<const float4*>(src + s * ld + col));
            }
            const int slot = q % W_STAGES;
            tc::mbar_wait(&S.empty[slot], (uint32_t(q / W_STAGES) & 1u) ^ 1u);
            uint8_t* dst = sW + uint32_t(slot) * W_STAGE;
            if (wg == 0) {
                // image operands: per 128-sample tile [hi | lo][16-feature block][sample 0..127][32 B] (the update kernel's
                // operand tile, copy_out_tile); a stage takes the 1 KiB slice of its 32 samples from every block.  One
                // copy per lane: lanes 0..12 the hi blocks, 13..25 the lo blocks.
                const int64_t gstage = s_begin + q;
#pragma unroll
                for (int mat = 0; mat < 2; ++mat) {
                    if (!(mat ? X_IMG : DY_IMG)) continue;
                    const int kb = (mat ? x_hp : dy_hp) >> 4;
                    if (lane == 0) tc::mbar_add_tx(&S.full[slot], uint32_t(2 * kb) * W_BLOCK);   // this warp's arrivals follow below
                    if (lane < 2 * kb) {
                        const int part = lane >= kb, b = lane - part * kb;
                        const uint8_t* src = reinterpret_cast<const uint8_t*>(mat ? X : dY) +
                                             (size_t(gstage >> 2) * size_t(2 * kb) + size_t(part * kb + b)) * 4096u + size_t(gstage & 3) * 1024u;
                        tc::bulk_g2s(dst + uint32_t(mat) * 2u * W_PART + uint32_t(part) * W_PART + uint32_t(b) * W_BLOCK, src, W_BLOCK,
                                     &S.full[slot]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < PER_WARP; ++i) {
                const int u = wg + 8 * i, mat = u / (W_UNITS / 2), rem = u % (W_UNITS / 2), rg = rem & 3, b = rem >> 2;
                if ((mat ? X_IMG : DY_IMG) || b >= (mat ? KPB : NPB)) continue;
                const int row = rg * 8 + (lane >> 2), f = lane & 3;
                uint2 hi, lo;
                if (f16) split4_f16(v[i], hi, lo);
                else split4(v[i], hi, lo);
                const uint32_t off = uint32_t(mat) * 2u * W_PART + uint32_t(b) * W_BLOCK + f4_offset(row, f);
                *reinterpret_cast<uint2*>(dst + off) = hi;
                *reinterpret_cast<uint2*>(dst + W_PART + off) = lo;
            }
            tc::fence_async_smem();
            tc::mbar_arrive(&S.full[slot]);
        }
    } else if (warp < W_ISSUER_WARP) {
        // ---------------- epilogue: add the accumulators into this CTA's fp32 partial
        const int ew = warp - W_EPI_WARP0, row = ew * 32 + lane;
        const uint32_t tlane = tmem_base + (uint32_t(ew * 32) << 16);
        const int nflush = (my_stages + W_FLUSH - 1) / W_FLUSH;
        float* mine = partial + size_t(blockIdx.x) * MAXF * MAXF;
        for (int fl = 0; fl < nflush; ++fl) {
            tc::mbar_wait(&S.d_full, uint32_t(fl) & 1u);
            tc::fence_after_sync();
            for (int half = 0; half * 128 < N; ++half) {
                const int n = half * 128 + row;
                float* prow = mine + size_t(n) * MAXF;
                for (int c = 0; c < KPB; ++c) {
                    float v[16];
                    tc::tmem_ld16p(tlane + uint32_t(half * 256 + c * 16), v);
                    if (n < N) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float4* p4 = reinterpret_cast<float4*>(prow + c * 16 + g * 4);
                            float4 acc = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                            if (fl > 0) {
                                const float4 old = *p4;
                                acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
                            }
                            *p4 = acc;
                        }
                    }
                }
            }
            tc::fence_before_sync();
            tc::mbar_arrive(&S.d_empty);
        }
    } else if (lane == 0) {
        // ---------------- MMA issuer
        // both operands bf16, or both fp16 (f16: the images of the pipelined fused update; kind::f16 does not mix formats)
        const uint32_t idesc = f16 ? (tc::make_idesc_f16(128, KPB * 16) | (1u << 15) | (1u << 16)) : tc::make_idesc_bf16_major(128, KPB * 16, true, true);
        // one descriptor for the ring base; every operand of every MMA is that plus a byte offset (>> 4) in the
        // address field: the issuing thread's descriptor arithmetic is one 64-bit add per operand
        const uint64_t d0 = tc::make_desc_sw32(tc::smem_addr(sW), W_BLOCK, 256);
        const int nhalf = N > 128 ? 2 : 1;
        bool acc = false;
        int since = 0;
        uint32_t nfl = 0;
        for (int q = 0; q < my_stages; ++q) {
            const int slot = q % W_STAGES;
            tc::mbar_wait(&S.full[slot], uint32_t(q / W_STAGES) & 1u);
            tc::fence_after_sync();
            const uint64_t ds = d0 + uint64_t((uint32_t(slot) * W_STAGE) >> 4);
#pragma unroll
            for (int ks = 0; ks < W_ROWS / 16; ++ks) {
                const uint64_t b_hi = ds + uint64_t((2u * W_PART + uint32_t(ks) * 512u) >> 4), b_lo = b_hi + uint64_t(W_PART >> 4);
                for (int half = 0; half < nhalf; ++half) {
                    const uint64_t a_hi = ds + uint64_t((uint32_t(half) * 8u * W_BLOCK + uint32_t(ks) * 512u) >> 4), a_lo = a_hi + uint64_t(W_PART >> 4);
                    const uint32_t d = tmem_base + uint32_t(half * 256);
                    tc::mma_bf16_ss(d, a_lo, b_hi, idesc, acc || ks > 0);
                    tc::mma_bf16_ss(d, a_hi, b_lo, idesc, true);
                    tc::mma_bf16_ss(d, a_hi, b_hi, idesc, true);
                }
            }
            acc = true;
            tc::mma_commit(&S.empty[slot]);
            if (++since == W_FLUSH || q == my_stages - 1) {
                tc::mma_commit(&S.d_full);
                tc::mbar_wait(&S.d_empty, nfl & 1u);
                tc::fence_after_sync();
                ++nfl;
                acc = false;
                since = 0;
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == W_ISSUER_WARP) tc::tmem_dealloc(tmem_base, 512);
}

// fixed-order sum of the per-CTA partials
__global__ void x3_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dW, int N, int K, int parts) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * K) return;
    const int n = idx / K, k = idx % K;
    float s = 0.f;
    for (int c = 0; c < parts; ++c) s += partial[(size_t(c) * MAXF + n) * MAXF + k];
    dW[idx] = s;
}

static bool feat_ok(int f) { return f >= 4 && f <= MAXF && f % 4 == 0; }

}  // namespace lx
}  // namespace g2048

using namespace g2048;
using namespace g2048::lx;

extern "C" {

int64_t g2048_x3_image_bytes(int32_t rows, int32_t cols) {
    if (rows <= 0 || cols <= 0) return -1;
    return int64_t(2) * ((cols + 15) / 16) * round16(rows) * 32;
}

int g2048_x3_pack(const float* W, int32_t R, int32_t Ccols, int32_t transpose, void* image, void* stream) {
    G2048_REQUIRE(W && image, "g2048_x3_pack: NULL pointer argument");
    G2048_REQUIRE(R >= 1 && R <= MAXF && Ccols >= 1 && Ccols <= MAXF, "g2048_x3_pack: weight shape outside [1,208]^2");
    const int rows = transpose ? Ccols : R, cols = transpose ? R : Ccols;
    const int NP = round16(rows), KB = (cols + 15) / 16, n = NP * KB * 16;
    x3_pack_kernel<<<(n + 255) / 256, 256, 0, cudaStream_t(stream)>>>(W, R, Ccols, transpose, static_cast<uint8_t*>(image), NP, KB);
    G2048_CHECK_LAUNCH("x3_pack_kernel");
    return G2048_OK;
}

int g2048_x3_gemm(const float* A, const void* image, float* C, int64_t M, int32_t N, int32_t K, void* stream) {
    G2048_REQUIRE(M >= 0, "g2048_x3_gemm: M < 0");
    if (M == 0) return G2048_OK;
    G2048_REQUIRE(A && image && C, "g2048_x3_gemm: NULL pointer argument");
    if (!feat_ok(N) || !feat_ok(K)) return fail(G2048_ESHAPE, "g2048_x3_gemm: N=%d, K=%d must be multiples of 4 in [4,208]", N, K);
    G2048_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(image) & 15) == 0, "g2048_x3_gemm: pointers must be 16-byte aligned");
    const int NP = round16(N), KB = (K + 15) / 16;
    const uint32_t b_bytes = 2u * KB * NP * 32u;
    const int smem = int(((b_bytes + 1023u) & ~1023u) + G_SLOTS * G_SLOT + sizeof(GemmBars) + 1024);
    G2048_CHECK_CUDA(ensure_smem(x3_gemm_kernel, smem));
    const int64_t ntiles = (M + 127) / 128;
    const int grid = int(ntiles < num_sms() ? ntiles : num_sms());
    x3_gemm_kernel<<<grid, NUM_THREADS, smem, cudaStream_t(stream)>>>(A, static_cast<const uint8_t*>(image), C, M, N, K, NP, KB);
    G2048_CHECK_LAUNCH("x3_gemm_kernel");
    return G2048_OK;
}

int64_t g2048_x3_wgrad_workspace_bytes(void) { return int64_t(num_sms()) * MAXF * MAXF * 4; }

int g2048_x3_wgrad(const float* dY, const float* X, float* dW, void* workspace, int64_t M, int32_t N, int32_t K, void* stream) {
    return g2048_x3_wgrad_tiled(dY, X, dW, workspace, M, N, K, 0, 0, stream);
}

int g2048_x3_wgrad_tiled(const float* dY, const float* X, float* dW, void* workspace, int64_t M, int32_t N, int32_t K,
                         int32_t dy_hp, int32_t x_hp, void* stream) {
    return g2048_x3_wgrad_images(dY, X, dW, workspace, M, N, K, dy_hp, x_hp, 0, stream);
}

int g2048_x3_wgrad_images(const float* dY, const float* X, float* dW, void* workspace, int64_t M, int32_t N, int32_t K,
                          int32_t dy_hp, int32_t x_hp, int32_t fp16, void* stream) {
    G2048_REQUIRE(M >= 0, "g2048_x3_wgrad: M < 0");
    G2048_REQUIRE(x_hp >= 0 || K == 48, "g2048_x3_wgrad: x_hp < 0 (X = packed boards) needs K == 48");
    G2048_REQUIRE((dy_hp == 0 || (dy_hp % 16 == 0 && dy_hp >= N && dy_hp <= MAXF)) && (x_hp <= 0 || (x_hp % 16 == 0 && x_hp >= K && x_hp <= MAXF)),
                  "g2048_x3_wgrad: tiled operands need a padded width that is a multiple of 16 in [features, 208]");
    G2048_REQUIRE(dW != nullptr, "g2048_x3_wgrad: dW is NULL");
    if (!feat_ok(N) || !feat_ok(K)) return fail(G2048_ESHAPE, "g2048_x3_wgrad: N=%d, K=%d must be multiples of 4 in [4,208]", N, K);
    cudaStream_t st = cudaStream_t(stream);
    if (M == 0) {
        G2048_CHECK_CUDA(cudaMemsetAsync(dW, 0, size_t(N) * K * 4, st));
        return G2048_OK;
    }
    G2048_REQUIRE(dY && X && workspace, "g2048_x3_wgrad: NULL pointer argument");
    G2048_REQUIRE((reinterpret_cast<uintptr_t>(dY) & 15) == 0 && (reinterpret_cast<uintptr_t>(X) & (x_hp < 0 ? 7 : 15)) == 0 &&
                      (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "g2048_x3_wgrad: pointers must be 16-byte aligned");
    const int smem = int(W_STAGES * W_STAGE + sizeof(WgradBars) + 1024);
    const int64_t stages = (M + W_ROWS - 1) / W_ROWS;
    const int grid = int(stages < num_sms() ? stages : num_sms());
    auto kern = x_hp < 0 ? (dy_hp ? x3_wgrad_kernel<true, false, true> : x3_wgrad_kernel<false, false, true>)
                : dy_hp  ? (x_hp ? x3_wgrad_kernel<true, true> : x3_wgrad_kernel<true, false>)
                         : (x_hp ? x3_wgrad_kernel<false, true> : x3_wgrad_kernel<false, false>);
    G2048_CHECK_CUDA(ensure_smem(kern, smem));
    kern<<<grid, W_THREADS, smem, st>>>(dY, X, static_cast<float*>(workspace), M, N, K, dy_hp, x_hp, fp16 ? 1 : 0);
    G2048_CHECK_LAUNCH("x3_wgrad_kernel");
    x3_wgrad_reduce_kernel<<<(N * K + 255) / 256, 256, 0, st>>>(static_cast<const float*>(workspace), dW, N, K, grid);
    G2048_CHECK_LAUNCH("x3_wgrad_reduce_kernel");
    return G2048_OK;
}

}  // extern "C"
