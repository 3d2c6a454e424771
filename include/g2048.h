/*
 * g2048.h -- C ABI of libg2048.so, the B200 (sm_100a) engine for the 2048-PPO hot path.
 *
 * This is the drop-in boundary: the entry points below are what a binding for the
 * reference's env / rollout / loss path would call.  The reference (pure Python) has no
 * FFI of its own; each entry cites the reference interface it replaces (file:line in
 * RobotSail/2048-PPO).  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 (G2048_OK) or a negative G2048_E* code; it never throws,
 *     never exits and never falls back to the CPU.  g2048_last_error() returns the
 *     message of the last failure on the calling thread.
 *   - all data pointers are DEVICE pointers owned by the caller (e.g. torch tensors'
 *     data_ptr()); nothing is allocated, freed or retained by the library.
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*) and never
 *     synchronise the host.
 *   - boards: one uint64 per board, cell (r,c) = nibble 4*(4r+c), value = tile exponent
 *     (0 empty, e -> tile 2^e, e <= 15).  The reference's Grid is list[list[int]] of
 *     exponents (game.py:3,50-51).
 *   - actions / direction ids: 0=UP 1=DOWN 2=LEFT 3=RIGHT (train.py:266, game.py:1092).
 *   - spawn draws: either a replay tensor of u32 pairs (u0,u1) or, when it is NULL, words
 *     0,1 of Philox4x32-10(counter=(env_id, ctr), key=seed), env_id = env0 + index.
 *     cell = k-th empty cell in row-major order, k = mulhi32(u0, #empty); exponent 2 iff
 *     u1 >= 3865470567 (<=> not (u1/2^32 < 0.9), game.py:937-939).
 */
#ifndef G2048_H
#define G2048_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G2048_OK 0
#define G2048_EINVAL (-1)   /* bad argument (null pointer, negative size, bad enum) */
#define G2048_ECUDA (-2)    /* a CUDA runtime call / launch failed */
#define G2048_EARCH (-3)    /* device is not sm_100 */
#define G2048_ESHAPE (-4)   /* unsupported model shape */

/* flags byte written by g2048_step */
#define G2048_FLAG_LEGAL_MASK 0x0F /* bit d: direction d is legal on the returned board */
#define G2048_FLAG_DONE 0x10       /* returned board has no legal move (game.py:1006 / 963) */
#define G2048_FLAG_INVALID 0x20    /* the requested move was illegal: board unchanged (game.py:959-978) */
#define G2048_FLAG_OVERFLOW 0x40   /* a merge created exponent 16, which a nibble cannot hold */

/* packed shaping record written by g2048_step (u64 per transition; 0 for an invalid move).
 * "after" = after the move, BEFORE the spawn (game.py:994-1002). */
#define G2048_SH_MONO_BEFORE(w) ((int)((w) & 63))                 /* game.py:985  */
#define G2048_SH_MONO_AFTER(w) ((int)(((w) >> 6) & 63))           /* game.py:999  */
#define G2048_SH_EMPT_BEFORE(w) ((int)(((w) >> 12) & 31))         /* game.py:988  */
#define G2048_SH_EMPT_AFTER(w) ((int)(((w) >> 17) & 31))          /* game.py:1000 */
#define G2048_SH_MAX_TILE_CREATED(w) ((int)(((w) >> 22) & 31))    /* game.py:158  */
#define G2048_SH_MAX_EXP_BEFORE(w) ((int)(((w) >> 27) & 15))      /* game.py:989  */
#define G2048_SH_CORNER_BEFORE(w) ((((w) >> 31) & 1) ? G2048_SH_MAX_EXP_BEFORE(w) : -G2048_SH_MAX_EXP_BEFORE(w)) /* game.py:982 */
#define G2048_SH_MAX_EXP_AFTER(w) ((int)(((w) >> 32) & 15))       /* game.py:1002 */
#define G2048_SH_CORNER_AFTER(w) ((((w) >> 36) & 1) ? G2048_SH_MAX_EXP_AFTER(w) : -G2048_SH_MAX_EXP_AFTER(w))    /* game.py:996 */
#define G2048_SH_SMOOTH_BEFORE(w) (-(int)(((w) >> 37) & 511))     /* game.py:981  */
#define G2048_SH_SMOOTH_AFTER(w) (-(int)(((w) >> 46) & 511))      /* game.py:995  */

const char* g2048_last_error(void);
const char* g2048_version(void);

/* Checks that `device` is a compute-capability-10.x GPU and makes it current. */
int g2048_init(int device);

/* Tables (g2048_lut_bytes() = 748464 bytes), built on the device: the general row table (move result,
 * merge codes, per-line potentials) and the move table of the 4-move expansion (move result, merge points,
 * created tile), 65536 x u32 each, followed by the dense step tables of the fused step kernel (M: u64 per
 * row with cells <= 11, S: u16 per row with cells <= 12; layout in csrc/g2048_device.cuh).  Replace
 * game.py:224-257 (_merge_and_shift_left/right_with_score) and the per-line parts of game.py:339-357,
 * 683-800 for every row.  g2048_build_lut is one-time initialisation and the only entry point that blocks:
 * it returns once the table is complete (the persistent kernels are launched with programmatic stream
 * serialisation and stage the table before they wait for their predecessor in the stream). */
int64_t g2048_lut_bytes(void);
int g2048_build_lut(void* d_lut, void* stream);

/* game.py:942-950 Game2048.reset for n boards.  replay: u32[n,4] = (u0,u1) first spawn,
 * (u2,u3) second spawn; NULL -> Philox words 0..3 at (env0+i, ctr). */
int g2048_reset(uint64_t* boards, int64_t n, const uint32_t* replay, uint64_t seed, uint64_t env0,
                uint64_t ctr, void* stream);

/* game.py:952-1030 Game2048.step for n (board, action) pairs.
 * boards_out / points / flags are required; shaping may be NULL (then the potentials are
 * not computed).  replay: u32[n,2] or NULL (Philox words 0,1 at (env0+i, ctr)). */
int g2048_step(const void* d_lut, const uint64_t* boards_in, const uint8_t* actions, uint64_t* boards_out,
               int32_t* points, uint8_t* flags, uint64_t* shaping, int64_t n, const uint32_t* replay,
               uint64_t seed, uint64_t env0, uint64_t ctr, void* stream);

/* The same step in the form of BASELINE config 2 ("1M random boards x 4 moves"): all four moves of every board, each with its
 * own spawn.  Transition (b, m) plays move m (0 UP, 1 DOWN, 2 LEFT, 3 RIGHT) on boards_in[b] with the draws of env id
 * env0 + 4 b + m -- bit for bit what g2048_step returns for the 4 n pairs (boards_in[b], m) -- and the outputs are [n,4] arrays
 * (16-byte aligned; flags 4-byte aligned).  One thread plays a board's four moves and shares what they have in common (transpose,
 * largest exponent, before-move potentials and corner rules, empty count): game.py:952-1030 four times per board at ~0.7 of the
 * instructions.  shaping may be NULL; replay: u32[n,4,2] or NULL. */
int g2048_step4(const void* d_lut, const uint64_t* boards_in, uint64_t* boards_out, int32_t* points, uint8_t* flags, uint64_t* shaping,
                int64_t n, const uint32_t* replay, uint64_t seed, uint64_t env0, uint64_t ctr, void* stream);

/* All four moves of every board without spawning: game.py:121-160 simulate_move,
 * game.py:167-184 preview_move_rewards, game.py:295-299 current_valid_directions.
 * succ[n,4] (== board where illegal), points[n,4] (0 where illegal), legal[n] (bit d),
 * max_tile[n,4] optional (NULL to skip). */
int g2048_expand4(const void* d_lut, const uint64_t* boards, uint64_t* succ, int32_t* points, uint8_t* legal,
                  uint8_t* max_tile, int64_t n, void* stream);

/* Board potentials, out int32[n,6] = monotonicity (game.py:683-800), emptiness (671-680),
 * smoothness (339-357), corner bonus (360-399), max exponent, legal mask (295-299). */
int g2048_potentials(const void* d_lut, const uint64_t* boards, int32_t* out, int64_t n, void* stream);

/* The float potentials Game2048.step also reports but the reward never uses (train.py:709-714):
 * out f64[n,7] = adjacency_bonus before/after (game.py:402-442), monotonic_chain_score before/after
 * (game.py:445-506), topological_score before/after with the anchor corner chosen on `before`
 * (game.py:634-668, 803-921, 986-1001), anchor as 4*row+col.  `after` = the pre-spawn successor.
 * Bit-identical to the reference's Python floats. */
int g2048_potentials_ext(const uint64_t* before, const uint64_t* after, double* out, int64_t n, void* stream);

/* game.py:92-101 to_model_format: out f32[n,48] = 16 x [exponent, row/3, col/3]. */
int g2048_encode(const uint64_t* boards, float* out, int64_t n, void* stream);

/* Symmetry augmentation of recorded steps (train.py:774-881; game.py:508-590 mirror_grid /
 * rotate_grid): op[i] = 0 mirror horizontal, 1 mirror vertical, 2 / 3 / 4 rotate 90 / 180 / 270
 * clockwise.  state_before / result_state boards are transformed, the action, the legal-direction
 * bits and the four log-probs (f32[n,4]) follow the direction remap of train.py:784-824. */
int g2048_augment(const uint64_t* before, const uint64_t* after, const uint8_t* action, const uint8_t* legal,
                  const float* logp, const uint8_t* op, uint64_t* before_out, uint64_t* after_out, uint8_t* action_out,
                  uint8_t* legal_out, float* logp_out, int64_t n, void* stream);

/* ---- rollout buffers -------------------------------------------------------------------
 * Time-major [T,B] arrays (index t*B + b).  flags = the g2048_step flags of the move plus
 * G2048_FLAG_VALID when the slot holds a recorded move (a finished game without auto-reset
 * leaves invalid slots behind).  A move with G2048_FLAG_DONE is the last of its episode. */
#define G2048_FLAG_VALID 0x80

/* Size of the scratch buffer the two reductions below need (device memory, caller-owned). */
int64_t g2048_reduce_workspace_bytes(void);

/* train.py:698-772 calculate_advantage (reward, discounted rewards-to-go per episode,
 * normalisation with the bias-corrected EMA moments, advantage), float64 inside, one pass.
 *   reward = w_points*points + w_mono*(gamma*mono_after - mono_before)
 *                            + w_empt*(gamma*empt_after - empt_before)      (train.py:702-719)
 *   with the terminal fix-up mono_after = empt_after = 0 on a DONE move       (train.py:318-322)
 *   g_norm = (G - mu_corrected) / (stddev + 1e-8); adv = g_norm - value       (train.py:760,772)
 * reward_out / g_raw_out may be NULL.  stats_out: f64[3] = {sum G, sum G^2, #valid} for the
 * batch mean / variance of train.py:738-739 (and for the cross-GPU allreduce). */
int g2048_rtg_advantage(const int32_t* points, const uint64_t* shaping, const uint8_t* flags, const float* value,
                        int32_t T, int64_t B, double gamma, double w_points, double w_mono, double w_empt,
                        double mu_corrected, double stddev, float* reward_out, float* g_raw_out, float* g_norm_out,
                        float* adv_out, double* stats_out, void* workspace, void* stream);

/* The same scan for rollouts of a FIXED horizon over games that continue past the buffer (persistent auto-reset envs):
 * bootstrap[b] (may be NULL = 0) is the raw, un-normalised return-to-go expected after the last slot of column b -- e.g. the
 * critic's value of the carried-over board, V * (stddev + 1e-8) + mu_corrected.  It only reaches the slots after the last
 * DONE move of the column; a column whose last slot is DONE or invalid ignores it.  The reference itself truncates with 0
 * (train.py:724-728: an episode cut by max_steps is simply shorter), which is what g2048_rtg_advantage does. */
int g2048_rtg_advantage_bootstrap(const int32_t* points, const uint64_t* shaping, const uint8_t* flags, const float* value,
                                  const float* bootstrap, int32_t T, int64_t B, double gamma, double w_points, double w_mono,
                                  double w_empt, double mu_corrected, double stddev, float* reward_out, float* g_raw_out,
                                  float* g_norm_out, float* adv_out, double* stats_out, void* workspace, void* stream);

/* train.py:497-554: masked log-softmax, PPO-clip surrogate (eps = clip_eps), entropy of the
 * clamped masked logits, smooth-L1 critic loss; loss = -(1/N) sum(ppo - c_v*vl + beta_ent*H).
 * Forward and analytic backward in one pass: dlogits f32[n,4] and dvalue f32[n] hold
 * d loss / d logits and d loss / d value (inv_n = 1/N already applied; N may be a global count).
 * old_logp: f32[n,old_logp_stride], stride 4 = all four rollout log-probs (train.py:326),
 * stride 1 = only the chosen action's.  legal: bit d = direction d legal (the reference's
 * action_mask is the complement, train.py:138,268).  flags may be NULL (= all valid).
 * stats_out: f64[4] = {sum ppo, sum smooth_l1, sum entropy, #valid}. */
int g2048_ppo_loss(const float* logits, const float* value, const float* old_logp, int32_t old_logp_stride,
                   const uint8_t* actions, const uint8_t* legal, const uint8_t* flags, const float* adv,
                   const float* g_norm, int64_t n, float clip_eps, float c_v, float beta_ent, float inv_n,
                   float* dlogits, float* dvalue, double* stats_out, void* workspace, void* stream);

/* train.py:577-597: the KL(old || new) statistic model_optimize_step logs after every optimizer step.  Per sample
 * sum over the legal moves of p_old (log p_old - log p_new), both the masked softmax of their logits (f32[n,4], 16-byte
 * aligned; old = the logits of the loss forward, new = a second forward with the updated weights).  kl_out: f32[n]
 * per-sample values, may be NULL; flags may be NULL (= all valid); slots that are invalid or have no legal move count 0.
 * stats_out: f64[3] = {sum KL, #valid, max KL}; fixed-order reduction (deterministic). */
int g2048_masked_kl(const float* old_logits, const float* new_logits, const uint8_t* legal, const uint8_t* flags, int64_t n,
                    float* kl_out, double* stats_out, void* workspace, void* stream);

/* ---- fused actor-critic rollout ------------------------------------------------------------
 * Replaces the loop of train.py:213-345 (play_game_for_episode) for B environments at once and
 * is what batched_rollout.play_games_batched (the module train.py:30 imports) is built on.
 *
 * Policy = the reference's GameMLP (game.py:1049-1220) in eval mode.  g2048_mlp_pack converts
 * its state_dict tensors (device pointers, row-major as torch stores them) into the kernel's
 * layout: stem.0.weight [h,48], stem.1.{weight,bias} [h], per block mlp.0.weight [h,h] and
 * mlp.1.{weight,bias} [h] (host arrays of `layers` device pointers), action_head.{weight [4,h],
 * bias [4]}, value_head.{weight [1,h], bias [1]}.  hidden <= 208, layers <= 8.
 * g2048_mlp_packed_floats returns the number of floats `packed` must hold (or -1). */
int64_t g2048_mlp_packed_floats(int32_t hidden, int32_t layers);
int g2048_mlp_pack(int32_t hidden, int32_t layers, const float* stem_w, const float* stem_ln_w, const float* stem_ln_b,
                   const float* const* block_w, const float* const* block_ln_w, const float* const* block_ln_b,
                   const float* action_w, const float* action_b, const float* value_w, const float* value_b,
                   float* packed, void* stream);

#define G2048_ROLLOUT_FP32 0
#define G2048_ROLLOUT_BF16 1
#define G2048_ROLLOUT_X3 2

typedef struct G2048Rollout {
    int64_t B;                 /* environments */
    int32_t T;                 /* steps played by this call */
    int32_t hidden, layers;    /* GameMLP shape (MLPConfig.hidden_dim / num_layers) */
    int32_t auto_reset;        /* 1: a finished game restarts at once (fixed [T,B] rollouts);
                                  0: it goes idle and its remaining slots are invalid (play to the end) */
    uint64_t seed, env0, ctr0; /* Philox key, id of env 0, counter of step 0 (step t uses ctr0+t):
                                  words 0,1 = spawn, word 2 = action sample; resets use the key
                                  seed ^ 0x9E3779B97F4A7C15 at the same counter, words 0..3 */
    const float* packed_weights;
    const void* lut;
    uint64_t* boards;              /* [B] in: start boards, out: boards after the last step */
    uint8_t* alive;                /* [B] in/out, may be NULL (= all alive); only used without auto_reset */
    const uint8_t* forced_actions; /* [T,B] or NULL: replay these actions instead of sampling */
    /* records, time-major [T,B] (train.py:299-326 StepData) */
    uint64_t* rec_boards;   /* state_before */
    uint8_t* rec_actions;   /* selected_direction */
    uint8_t* rec_legal;     /* bit d: direction d legal in state_before (complement of action_mask) */
    float* rec_logp;        /* [T,B,4] policy_logprobs: log_softmax of the masked logits, -inf where illegal */
    float* rec_value;       /* predicted_future_value */
    int32_t* rec_points;    /* points_earned */
    uint64_t* rec_shaping;  /* packed G2048_SH_* record of the move */
    uint8_t* rec_flags;     /* g2048_step flags of the move | G2048_FLAG_VALID */
    float* rec_entropy;     /* entropy of the masked action distribution; may be NULL */
    int32_t tensor_cores;   /* G2048_ROLLOUT_FP32 (0): fp32 FFMA GEMMs (log-probs within ~1e-6 of the torch policy);
                               G2048_ROLLOUT_BF16 (1): bf16 tcgen05 GEMMs with fp32 accumulation in tensor memory
                                  (recorded log-probs within ~1e-2 of the fp32 policy: a labelled variant);
                               G2048_ROLLOUT_X3 (2): split-fp16 tcgen05 GEMMs (every operand = hi + lo in fp16, three
                                  products per k-step, fp32 accumulation): fp32-grade, log-probs / values within
                                  2e-5 of the torch fp32 policy (game.py:1192-1203, train.py:256-274); <= 4 blocks */
    int32_t reserved_;
    void* sched_workspace;  /* optional: 4 * ceil(B / 128) bytes of device scratch.  With it the X3 kernel cuts the horizon of a tile
                               into segments dealt round-robin over the SMs (a tile's boards pass from segment to segment through
                               `boards` / `alive`), which fills the last, partial wave of tiles: C3's 512 tiles are 3.46 waves on
                               148 SMs.  Results do not depend on it.  NULL (or a zero-filled tail of the struct): off. */
} G2048Rollout;

int g2048_rollout_mlp(const G2048Rollout* params, void* stream);

/* GameURM policy (game.py:1223-1458, default GameURMConfig game.py:31-42: hidden 64, 4 heads,
 * inter 120, conv kernel 2; 1..2 layers) -- BASELINE config #5.  Same G2048Rollout records.
 * `tensor_cores` = G2048_ROLLOUT_X3 (or _FP32: there is no FFMA variant): the four projections of a block
 * on tcgen05 with split-fp16 operands (x = hi + lo, three products per k-step, fp32 accumulation in tensor
 * memory), K / V, attention, norms, SiLUs and the depthwise conv in fp32 on CUDA cores: fp32 grade against the
 * reference's forward.  G2048_ROLLOUT_BF16: the round-1 kernel with single fp16 operands and fp16 K / V
 * (log-probs 1e-2 off the fp32 model): a labelled variant.  `loops` = GameURMConfig.num_loops.  Pointer
 * arrays hold `layers` device pointers:
 * layers.{l}.attn.qkv_proj.weight [192,64], attn.o_proj.weight [64,64], mlp.gate_up_proj.weight
 * [240,64], mlp.dwconv.weight [120,1,2], mlp.dwconv.bias [120], mlp.down_proj.weight [64,120];
 * stem.0.weight [64,3], stem.1.{weight,bias} [64], init_hidden [1,16,64], heads as for the MLP.
 * `packed` (g2048_urm_packed_floats floats, 256-byte aligned) holds both kernels' images. */
int64_t g2048_urm_packed_floats(int32_t hidden, int32_t layers, int32_t heads, int32_t inter);
int g2048_urm_pack(int32_t layers, const float* stem_w, const float* stem_ln_w, const float* stem_ln_b,
                   const float* init_hidden, const float* const* qkv_w, const float* const* o_w,
                   const float* const* gate_up_w, const float* const* dwconv_w, const float* const* dwconv_b,
                   const float* const* down_w, const float* action_w, const float* action_b, const float* value_w,
                   const float* value_b, float* packed, void* stream);
int g2048_rollout_urm(const G2048Rollout* params, int32_t loops, void* stream);

/* ---- GameURM update: the block ops that are not projections, forward and hand-written backward ----------
 * Replaces, inside a GameURM update step, torch's scaled_dot_product_attention (game.py:1296-1317), the inner chain of
 * GameConvSwiGLU (silu(gate) * up -> depthwise Conv1d(kernel 2, padding 1, trimmed) -> silu, game.py:1264-1276) and
 * rms_norm(hidden + branch) (game.py:1223-1229, 1345-1350) with their autograd backward; the projections run on
 * g2048_x3_gemm / g2048_x3_wgrad.  Default GameURMConfig shapes only: 16 tokens, 4 heads x head_dim 16, hidden 64, inter 120.
 * All tensors row-major fp32 device memory; B = boards (envs), rows = B * 16 tokens.
 *   attention  qkv [B,16,192] = per token [q | k | v], head h at columns 16h..16h+15 of each third; out / dout [B,16,64]
 *   swiglu     gate, up, y, dy, dgate, dup [B,16,120]; conv_w [120,2] (= dwconv.weight [120,1,2]), conv_b [120];
 *              dconv_w / dconv_b are overwritten with the batch sums (fixed-order reduction); workspace of
 *              g2048_urm_swiglu_workspace_floats() floats
 *   norm       y = s * rsqrt(mean(s^2) + eps), s = x + r, rows of 64; rs [rows] keeps the row factor for the backward, which
 *              returns ds (the gradient of both x and r) from y, rs and dy */
int g2048_urm_attn_fwd(const float* qkv, float* out, int64_t B, void* stream);
int g2048_urm_attn_bwd(const float* qkv, const float* dout, float* dqkv, int64_t B, void* stream);
int g2048_urm_swiglu_fwd(const float* gate, const float* up, const float* conv_w, const float* conv_b, float* y, int64_t B, void* stream);
int64_t g2048_urm_swiglu_workspace_floats(void);
int g2048_urm_swiglu_bwd(const float* gate, const float* up, const float* conv_w, const float* conv_b, const float* dy, float* dgate,
                         float* dup, float* dconv_w, float* dconv_b, float* workspace, int64_t B, void* stream);
int g2048_urm_norm_fwd(const float* x, const float* r, float* y, float* rs, int64_t rows, float eps, void* stream);
int g2048_urm_norm_bwd(const float* y, const float* rs, const float* dy, float* ds, int64_t rows, void* stream);

/* ---- policy update: fused y = res + ReLU(LayerNorm(z)) and its backward ---------------------
 * The elementwise chain of a GameMLP block (game.py:1038-1046; stem: game.py:1069-1073, res = NULL)
 * in eval/p=0 dropout form.  Row-major [n,h] fp32, h a multiple of 4 up to 256, eps = 1e-5 in the
 * reference.  forward also returns the per-row mean and rstd the backward needs.  backward
 * returns dz, dgamma, dbeta (d res = gout, unchanged); `workspace` holds
 * g2048_ln_workspace_floats(h) floats.  The GEMMs around it stay in torch/cuBLAS. */
int64_t g2048_ln_workspace_floats(int32_t h);
int g2048_ln_relu_res_fwd(const float* z, const float* gamma, const float* beta, const float* res, float* y,
                          float* mean, float* rstd, int64_t n, int32_t h, float eps, void* stream);
int g2048_ln_relu_res_bwd(const float* z, const float* gamma, const float* beta, const float* mean, const float* rstd,
                          const float* gout, float* dz, float* dgamma, float* dbeta, float* workspace, int64_t n,
                          int32_t h, void* stream);

/* ---- policy update: fp32-grade Linear GEMMs on tcgen05 (csrc/g2048_linear.cu) ----------------
 * The Linear layers of GameMLP (game.py:1038-1046, 1069-1073) inside model_optimize_step's forward
 * and loss.backward() (train.py:497-556), which the reference runs as fp32 torch matmuls.  Operands
 * are split in-kernel into two bf16 terms (x = hi + lo) and every product is Alo*Bhi + Ahi*Blo + Ahi*Bhi
 * with fp32 accumulation in tensor memory: fp32-grade results (about 1e-6 relative) at tensor-core rate.
 * Activations are row-major fp32 [M, features], features a multiple of 4 in [4, 208]; all pointers
 * 16-byte aligned.
 *   g2048_x3_pack   weight W [R, C] (torch Linear layout [out, in]) -> operand image for x3_gemm;
 *                   transpose = 0: image of W   (forward  Y = X W^T,  N = R, K = C)
 *                   transpose = 1: image of W^T (dgrad   dX = dY W,   N = C, K = R)
 *                   `image` holds g2048_x3_image_bytes(N, K) bytes.
 *   g2048_x3_gemm   C[M, N] = A[M, K] * B[N, K]^T, B given as a packed image.
 *   g2048_x3_wgrad  dW[N, K] = dY[M, N]^T * X[M, K] (sum over samples; per-SM partials are added in a
 *                   fixed order, so the result is deterministic); `workspace` holds
 *                   g2048_x3_wgrad_workspace_bytes() bytes. */
int64_t g2048_x3_image_bytes(int32_t rows, int32_t cols);
int g2048_x3_pack(const float* W, int32_t R, int32_t C, int32_t transpose, void* image, void* stream);
int g2048_x3_gemm(const float* A, const void* image, float* C, int64_t M, int32_t N, int32_t K, void* stream);
int64_t g2048_x3_wgrad_workspace_bytes(void);
int g2048_x3_wgrad(const float* dY, const float* X, float* dW, void* workspace, int64_t M, int32_t N, int32_t K,
                   void* stream);
/* same, with either operand as the bf16 hi|lo operand image g2048_update_mlp_fwd_bwd writes (h_out / dz_out, 4 bytes
 * per value like fp32): per tile of 128 samples [hi | lo][16-feature block][sample 0..127][32 B, the two 16-byte halves
 * swapped on (sample >> 2) & 1], rows past the sample count zero -- the update kernel's own MMA operand tile, copied out
 * by two bulk copies; the weight-gradient kernel bulk-copies the 32-sample slices of its blocks into its operand ring.
 * x_hp = -1: X is the array of packed boards (uint64 per sample) and K == 48: the model input of game.py:92-101 is formed
 * in the kernel (the stem's weight gradient without a g2048_encode pass).  `*_hp` = padded column count of that operand (hidden rounded up to 16), 0 = row-major fp32. */
int g2048_x3_wgrad_tiled(const float* dY, const float* X, float* dW, void* workspace, int64_t M, int32_t N, int32_t K,
                         int32_t dy_hp, int32_t x_hp, void* stream);

/* The same with the term format chosen: fp16 != 0 -- both operands are split (or arrive as images) in fp16 terms, x = hi + lo
 * with 22 mantissa bits, for operands of O(1) magnitude such as the fused update's activations and its loss-scaled
 * gradients; 0 = bf16 terms (16 bits, the range of fp32).  kind::f16 does not mix the two formats in one product. */
int g2048_x3_wgrad_images(const float* dY, const float* X, float* dW, void* workspace, int64_t M, int32_t N, int32_t K,
                          int32_t dy_hp, int32_t x_hp, int32_t fp16, void* stream);

/* ---- policy update: fused forward + loss + backward-data of GameMLP (csrc/g2048_update_x3.cu) ----
 * One persistent tcgen05 kernel runs, per 128-sample tile and without leaving the SM, what
 * model_optimize_step does between `model(x)` and the weight gradients (train.py:491-556): the GameMLP
 * forward (game.py:1145-1220) from packed boards, the PPO-clip + critic + entropy terms (train.py:497-554)
 * and autograd's backward down to every pre-LayerNorm gradient dz_l.  GEMMs are split-fp16: every operand is two fp16
 * terms (22 mantissa bits) and every k-step three products, forward and backward (fp32-grade, ~2e-7 of scale).  fp16 terms
 * want O(1) magnitudes, so the caller passes inv_n PRE-MULTIPLIED by a power-of-two loss scale S (for a mean over 3e7
 * samples 1/n alone is below the fp16 range): dhead, dz_out, ln_grad, head_bias_grad and every weight gradient formed from
 * the images then carry the factor S, which the caller divides out (stats do not).
 * It emits what the weight-gradient GEMMs need, h_out[l] (l = 0 stem output .. L) and dz_out[l], as the fp16 hi|lo
 * operand images of g2048_x3_wgrad_images (ceil(n/128) * 128 * HP * 4 bytes per l, HP = g2048_update_mlp_padded(hidden)),
 * plus dhead [n, 8] = d loss / d (4 logits, V, 0, 0, 0), row-major; then (g2048_x3_wgrad_images, fp16 = 1)
 *     d stem.0.weight        = wgrad(dz_out[0], g2048_encode(boards))        [hidden, 48]
 *     d backbone.l.mlp.0.w   = wgrad(dz_out[l+1], h_out[l])                  [hidden, hidden]
 *     d (action|value) head  = wgrad(dhead, h_out[L]) rows 0..3 | 4          [8, hidden]
 * and returns the small gradients itself: ln_grad [L+1][2][hidden] (d LayerNorm weight | bias per layer,
 * stem first), head_bias_grad [5] (action_head.bias, value_head.bias) and stats double[4] = {sum ppo,
 * sum smooth_l1, sum entropy, count} as g2048_ppo_loss.  All per-SM partial sums are combined in a fixed
 * order (deterministic).  hidden: multiple of 4 in [16, 208]; layers 1..2; dropout: see dropout_p.
 * backward = 0: forward only (writes logits [n,4] and/or value [n]; used by the parity tests). */
int32_t g2048_update_mlp_padded(int32_t hidden);   /* HP: column count of the operand tiles / images (64, 128, 192 or 208) */
int64_t g2048_update_mlp_pack_bytes(int32_t hidden, int32_t layers);
int g2048_update_mlp_pack(int32_t hidden, int32_t layers, const float* stem_w, const float* stem_ln_w, const float* stem_ln_b,
                          const float* const* block_w, const float* const* block_ln_w, const float* const* block_ln_b,
                          const float* action_w, const float* action_b, const float* value_w, const float* value_b,
                          void* packed, void* stream);
int64_t g2048_update_mlp_workspace_bytes(int32_t hidden, int32_t layers);

typedef struct G2048UpdateMlp {
    int64_t n;                     /* samples */
    int32_t hidden, layers;
    int32_t decouple_critic;       /* MLPConfig.decouple_critic: the value head's gradient stops at h_L (game.py:1208) */
    int32_t backward;              /* 0 = forward only */
    const uint64_t* boards;        /* [n] state_before of every sample */
    const uint8_t* actions;        /* [n] */
    const uint8_t* legal;          /* [n] legal-move bits */
    const uint8_t* flags;          /* [n] or NULL; samples without G2048_FLAG_VALID contribute nothing */
    const float* old_logp;         /* [n, old_logp_stride] rollout log-probs (stride 4: all actions, 1: the chosen one) */
    int32_t old_logp_stride;
    int32_t reserved_;
    const float* adv;              /* [n] */
    const float* g_norm;           /* [n] normalised return-to-go */
    float clip_eps, critic_strength, entropy_strength;
    float inv_n;                   /* 1 / (global sample count of the minibatch): the loss is a mean (train.py:554) */
    const void* packed;            /* g2048_update_mlp_pack output */
    void* workspace;               /* g2048_update_mlp_workspace_bytes bytes */
    float* h_out;                  /* [layers+1][ceil(n/128)*128*HP*4 bytes] operand images */
    float* dz_out;                 /* same */
    float* dhead;                  /* [n][8] */
    float* logits;                 /* [n][4] or NULL */
    float* value;                  /* [n] or NULL */
    float* ln_grad;                /* [layers+1][2][hidden] */
    float* head_bias_grad;         /* [5] */
    double* stats;                 /* [4] */
    /* Dropout of the residual blocks (game.py:1038-1046; the reference trains with MLPConfig.dropout = 0.1 active,
     * train.py:483): element (sample i, block l, column c) is dropped iff lane (c & 7) of the 8 16-bit lanes of
     * Philox4x32-10(counter = (dropout_sample0 + i [64 bit], l, c >> 3), key = dropout_seed) -- word order x, y, z, w,
     * low half first -- is < round(dropout_p * 65536); kept elements are scaled by 1 / (1 - dropout_p).  The backward
     * pass applies the same mask.  dropout_p = 0 (or the struct zero-filled): off. */
    float dropout_p;
    int32_t reserved2_;
    uint64_t dropout_seed;
    uint64_t dropout_sample0;
} G2048UpdateMlp;

int g2048_update_mlp_fwd_bwd(const G2048UpdateMlp* params, void* stream);

/* tcgen05 building-block self-test (not part of the reference's interface): C[128,N] =
 * A[128,K] * W[N,K]^T with bf16-rounded operands and fp32 accumulation in tensor memory.
 * K, N multiples of 16, <= 256.  Pins the UMMA descriptor / swizzle conventions on hardware. */
int g2048_tc_gemm_selftest(const float* A, const float* W, float* C, int32_t K, int32_t N, void* stream);
/* the same with the term format chosen (1 = fp16, 0 = bf16) for BOTH operands: a kind::f16 MMA with one bf16 and one fp16
 * operand faults on B200 (measured), so a_f16 != w_f16 is rejected with G2048_EINVAL */
int g2048_tc_gemm_selftest_fmt(const float* A, const float* W, float* C, int32_t K, int32_t N, int32_t a_f16, int32_t w_f16,
                               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* G2048_H */
