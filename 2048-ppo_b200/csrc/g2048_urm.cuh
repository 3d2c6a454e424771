// g2048_urm.cuh -- what the two GameURM rollout kernels share: model constants, the packed parameter layout, the pack sources.
// (g2048_rollout_urm.cu: single fp16 operands, weights resident in shared memory -- a labelled variant; g2048_rollout_urm_x3.cu:
// split-fp16 operands, fp32 grade, the default.)
#pragma once
#include "g2048_rollout.cuh"

namespace g2048 {
namespace urm {

constexpr int H = 64, SEQ = 16, NHEAD = 4, HD = 16, INTER = 120, QKV = 3 * H, GU = 2 * INTER;
constexpr int MAX_LAYERS = 2;
constexpr int THREADS = 128;

// packed parameter buffer: fp32 section, then (128-byte aligned) fp16 images per layer
constexpr int F_STEM_W = 0;                       // [64][3]
constexpr int F_STEM_G = F_STEM_W + H * 3;        // [64]
constexpr int F_STEM_B = F_STEM_G + H;            // [64]
constexpr int F_INIT = F_STEM_B + H;              // [16][64]
constexpr int F_HEADW = F_INIT + SEQ * H;         // [5][64]
constexpr int F_HEADB = F_HEADW + 5 * H;          // [8]
constexpr int F_CONV = F_HEADB + 8;               // per layer: w0[128], w1[128], b[128]
constexpr int F_CONV_STRIDE = 3 * 128;
__host__ __device__ constexpr int f_total(int L) { return (F_CONV + L * F_CONV_STRIDE + 31) / 32 * 32; }
constexpr int IMG_QKV = 0;                                   // [192 rows][128 B]
constexpr int IMG_O = IMG_QKV + QKV * 128;                   // [64 rows][128 B]
constexpr int IMG_GU = IMG_O + H * 128;                      // [240 rows][128 B]
constexpr int IMG_D = IMG_GU + GU * 128;                     // 2 blocks x [64 rows][128 B] (K = 120 -> 128)
constexpr int IMG_LAYER = IMG_D + 2 * H * 128;               // 79 872 B
__host__ __device__ constexpr int64_t total_floats(int L) { return f_total(L) + int64_t(L) * IMG_LAYER / 4; }


// ---- fp32-grade section (g2048_rollout_urm_x3.cu), behind the fp16 images, 256-byte aligned:
//   emb table   [16 exponents][16 cells][64] fp32: the stem output Linear(3 -> 64) + LayerNorm + SiLU (game.py:1376-1380) of every
//               (exponent, cell) pair -- the stem sees nothing else -- computed once at pack time;
//   weight stream per layer, split-fp16 (w = hi + lo) k-blocks of 16 input features in the 32-byte-swizzled K-major layout of
//               g2048_tc.cuh, each k-block [hi: N rows x 32 B | lo: N rows x 32 B], in the order the kernel consumes them:
//               QKV (N = 192, 4 k-blocks), O (N = 64, 4), GU (4 k-blocks, each [G1 = gate | up of channels 0..63, 128 rows][G2 = gate | up
//               of channels 64..119, 112 rows]), D (N = 64, K = 120 -> 128: 8 k-blocks).
constexpr int X3_EMB_FLOATS = 16 * SEQ * H;
constexpr int X3_QKV = 0;                                    // 4 x 12 288 B
constexpr int X3_O = X3_QKV + 4 * QKV * 64;                  // 4 x 4 096 B
constexpr int X3_GU = X3_O + 4 * H * 64;                     // 4 x 15 360 B
constexpr int X3_D = X3_GU + 4 * GU * 64;                    // 8 x 4 096 B
constexpr int X3_LAYER = X3_D + 8 * H * 64;                  // 159 744 B
__host__ __device__ constexpr int64_t x3_base(int L) { return (total_floats(L) + 63) / 64 * 64; }   // floats
__host__ __device__ constexpr int64_t total_floats_all(int L) { return x3_base(L) + X3_EMB_FLOATS + int64_t(L) * X3_LAYER / 4; }

struct PackSrc {
    const float *stem_w, *stem_g, *stem_b, *init_hidden, *act_w, *act_b, *val_w, *val_b;
    const float* qkv[MAX_LAYERS];
    const float* o[MAX_LAYERS];
    const float* gu[MAX_LAYERS];
    const float* conv_w[MAX_LAYERS];
    const float* conv_b[MAX_LAYERS];
    const float* down[MAX_LAYERS];
};


int launch_urm_x3_pack(const PackSrc& s, int L, float* packed, cudaStream_t st);
int launch_rollout_urm_x3(const RolloutParams& p, int loops, cudaStream_t st);

}  // namespace urm
}  // namespace g2048
