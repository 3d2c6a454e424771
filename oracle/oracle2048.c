/*
 * oracle2048.c -- CPU restatement of the RobotSail/2048-PPO hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package under
 * 2048-ppo_b200/) may link, load or call this file.  It is used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * as the checker and as the timed CPU baseline ("port").
 *
 * Every function follows the reference's algorithm literally (list scans,
 * four explicit rotations, ...), NOT the bit tricks the CUDA kernels use, so
 * that the two implementations are independent.  Citations are file:line in
 * the reference checkout (/root/reference).
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against fixtures
 * generated from the reference's own game.py / train.py in the build container
 * (oracle/make_golden.py -> tests/golden/ npz files) and against the reference's
 * shipped replay docs/data/best_game.json (1249 transitions).
 *
 * Board packing (ours): one uint64 per board, cell (r,c) holds the tile
 * exponent (0 = empty, e = tile 2^e) in nibble 4*(4r+c).  The reference's
 * Grid is list[list[int]] of exponents (game.py:3,50-51).
 *
 * Directions / action ids follow train.py:266: 0=UP 1=DOWN 2=LEFT 3=RIGHT.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GS 4

enum { ORC_UP = 0, ORC_DOWN = 1, ORC_LEFT = 2, ORC_RIGHT = 3 };

/* per-transition record mirroring the info dict of Game2048.step
 * (game.py:964-977 invalid path, game.py:1012-1029 valid path). */
typedef struct {
    int32_t points;            /* game.py:154,1010 */
    int32_t done;              /* game.py:963,1006 */
    int32_t invalid;           /* game.py:965,1013 */
    int32_t overflow;          /* ours: a merge produced exponent 16 (nibble overflow) */
    int32_t mono_before;       /* game.py:985 */
    int32_t mono_after;        /* game.py:999  (pre-spawn board) */
    int32_t empt_before;       /* game.py:988 */
    int32_t empt_after;        /* game.py:1000 (pre-spawn board) */
    int32_t max_tile_created;  /* game.py:158,1015 */
    int32_t max_exp_before;    /* game.py:989 */
    int32_t max_exp_after;     /* game.py:1002 */
    int32_t corner_before;     /* game.py:982 */
    int32_t corner_after;      /* game.py:996 */
    int32_t smooth_before;     /* game.py:981 */
    int32_t smooth_after;      /* game.py:995 */
    int32_t legal_before;      /* bit d = direction d legal on the input board (game.py:295-299) */
    int32_t legal_after;       /* same for the returned (post-spawn) board */
} orc_info_t;

/* ---------------------------------------------------------------- packing */

void orc_unpack(uint64_t b, int g[GS][GS]) {
    for (int r = 0; r < GS; ++r)
        for (int c = 0; c < GS; ++c)
            g[r][c] = (int)((b >> (4 * (4 * r + c))) & 0xF);
}

/* returns 1 if any exponent does not fit a nibble */
int orc_pack(int g[GS][GS], uint64_t* out) {
    uint64_t b = 0;
    int ovf = 0;
    for (int r = 0; r < GS; ++r)
        for (int c = 0; c < GS; ++c) {
            int e = g[r][c];
            if (e > 15) { ovf = 1; e = 15; }
            b |= (uint64_t)e << (4 * (4 * r + c));
        }
    *out = b;
    return ovf;
}

/* ------------------------------------------------------------- A1: rows  */

/* game.py:224-244 _merge_and_shift_left_with_score */
void orc_merge_left(const int row[GS], int out[GS], int* score, int* max_tile) {
    int nz[GS], n = 0;
    for (int i = 0; i < GS; ++i)
        if (row[i] != 0) nz[n++] = row[i];
    int m = 0, s = 0, mt = 0, i = 0;
    int merged[GS];
    while (i < n) {
        if (i + 1 < n && nz[i] == nz[i + 1]) {
            int ne = nz[i] + 1;
            merged[m++] = ne;
            s += 1 << ne;
            if (ne > mt) mt = ne;
            i += 2;
        } else {
            merged[m++] = nz[i];
            i += 1;
        }
    }
    for (int k = 0; k < GS; ++k) out[k] = k < m ? merged[k] : 0;
    *score = s;
    *max_tile = mt;
}

/* game.py:252-257 _merge_and_shift_right_with_score */
void orc_merge_right(const int row[GS], int out[GS], int* score, int* max_tile) {
    int rev[GS], tmp[GS];
    for (int i = 0; i < GS; ++i) rev[i] = row[GS - 1 - i];
    orc_merge_left(rev, tmp, score, max_tile);
    for (int i = 0; i < GS; ++i) out[i] = tmp[GS - 1 - i];
}

/* ------------------------------------------------------- A2: simulate_move */

/* game.py:121-160 simulate_move */
void orc_simulate_move(int g[GS][GS], int dir, int out[GS][GS], int* score, int* max_tile) {
    int total = 0, mt = 0;
    if (dir == ORC_UP || dir == ORC_DOWN) {
        int work[GS][GS], res[GS][GS];
        for (int i = 0; i < GS; ++i)
            for (int j = 0; j < GS; ++j) work[i][j] = g[j][i];
        for (int i = 0; i < GS; ++i) {
            int s, m;
            if (dir == ORC_UP) orc_merge_left(work[i], res[i], &s, &m);
            else               orc_merge_right(work[i], res[i], &s, &m);
            total += s;
            if (m > mt) mt = m;
        }
        for (int i = 0; i < GS; ++i)
            for (int j = 0; j < GS; ++j) out[i][j] = res[j][i];
    } else {
        for (int i = 0; i < GS; ++i) {
            int s, m;
            if (dir == ORC_LEFT) orc_merge_left(g[i], out[i], &s, &m);
            else                 orc_merge_right(g[i], out[i], &s, &m);
            total += s;
            if (m > mt) mt = m;
        }
    }
    *score = total;
    *max_tile = mt;
}

/* ------------------------------------------------------- A3: legality     */

/* shared prologue of game.py:259-281 / 301-322: transpose for UP/DOWN, then
 * UP is treated as LEFT, and LEFT reverses every row so the scan below always
 * runs "toward the right". */
static void orient(int g[GS][GS], int dir, int s[GS][GS]) {
    int t[GS][GS];
    if (dir == ORC_UP || dir == ORC_DOWN) {
        for (int i = 0; i < GS; ++i)
            for (int j = 0; j < GS; ++j) t[i][j] = g[j][i];
        if (dir == ORC_UP) dir = ORC_LEFT;
    } else {
        memcpy(t, g, sizeof(t));
    }
    if (dir == ORC_LEFT) {
        for (int i = 0; i < GS; ++i)
            for (int j = 0; j < GS; ++j) s[i][j] = t[i][GS - 1 - j];
    } else {
        memcpy(s, t, sizeof(t));
    }
}

/* game.py:259-293 can_move_in_direction */
int orc_can_move(int g[GS][GS], int dir) {
    int s[GS][GS];
    orient(g, dir, s);
    int can = 0;
    for (int r = 0; r < GS; ++r) {
        int found = 0;
        for (int c = 0; c < GS; ++c) {
            if (s[r][c] > 0) found = 1;
            if (found && s[r][c] == 0) { can = 1; break; }
        }
    }
    return can;
}

/* game.py:301-330 can_merge_in_direction */
int orc_can_merge(int g[GS][GS], int dir) {
    int s[GS][GS];
    orient(g, dir, s);
    int can = 0;
    for (int r = 0; r < GS; ++r)
        for (int c = 0; c + 1 < GS; ++c)
            if (s[r][c] == s[r][c + 1] && s[r][c] != 0) can = 1;
    return can;
}

/* game.py:116-119 direction_has_step */
int orc_direction_has_step(int g[GS][GS], int dir) {
    return orc_can_move(g, dir) || orc_can_merge(g, dir);
}

/* game.py:295-299 current_valid_directions, as a bit mask in action-id order */
int orc_legal_mask(int g[GS][GS]) {
    int m = 0;
    for (int d = 0; d < 4; ++d)
        if (orc_direction_has_step(g, d)) m |= 1 << d;
    return m;
}

/* game.py:103-114 has_next_step */
int orc_has_next_step(int g[GS][GS]) { return orc_legal_mask(g) != 0; }

/* ------------------------------------------------------- A6-A8: potentials */

/* game.py:671-680 emptiness */
int orc_emptiness(int g[GS][GS]) {
    int n = 0;
    for (int r = 0; r < GS; ++r)
        for (int c = 0; c < GS; ++c) n += g[r][c] == 0;
    return n;
}

/* game.py:339-357 smoothness_score (integral; returned as int) */
int orc_smoothness(int g[GS][GS]) {
    int score = 0;
    for (int i = 0; i < GS; ++i)
        for (int j = 0; j < GS; ++j) {
            if (g[i][j] == 0) continue;
            if (j < GS - 1 && g[i][j + 1] != 0) score -= abs(g[i][j] - g[i][j + 1]);
            if (i < GS - 1 && g[i + 1][j] != 0) score -= abs(g[i][j] - g[i + 1][j]);
        }
    return score;
}

static int is_corner(int r, int c) { return (r == 0 || r == GS - 1) && (c == 0 || c == GS - 1); }

/* game.py:360-399 corner_bonus */
int orc_corner_bonus(int g[GS][GS]) {
    int mx = 0;
    for (int i = 0; i < GS; ++i)
        for (int j = 0; j < GS; ++j)
            if (g[i][j] > mx) mx = g[i][j];
    if (mx == 0) return 0;
    int in_corner = 0;
    for (int i = 0; i < GS && !in_corner; ++i)
        for (int j = 0; j < GS; ++j)
            if (g[i][j] == mx && is_corner(i, j)) { in_corner = 1; break; }
    return in_corner ? mx : -mx;
}

/* game.py:683-800 monotonicity: best of the four rotations, then the
 * first-row-major-max corner rule (x2 in a corner, //2 otherwise). */
int orc_monotonicity(int g[GS][GS]) {
    int best = -1;
    int cur[GS][GS], nxt[GS][GS];
    memcpy(cur, g, sizeof(cur));
    for (int rot = 0; rot < 4; ++rot) {
        int count = 0;
        for (int r = 0; r < GS; ++r)
            for (int c = 0; c < GS - 1; ++c) {
                int l = cur[r][c], rr = cur[r][c + 1];
                if (l > 0 && rr > 0 && l >= rr) count++;
            }
        for (int c = 0; c < GS; ++c)
            for (int r = 0; r < GS - 1; ++r) {
                int t = cur[r][c], b = cur[r + 1][c];
                if (t > 0 && b > 0 && t >= b) count++;
            }
        if (count > best) best = count;
        /* game.py:733 rotate 90 degrees clockwise */
        for (int i = 0; i < GS; ++i)
            for (int j = 0; j < GS; ++j) nxt[i][j] = cur[GS - 1 - j][i];
        memcpy(cur, nxt, sizeof(cur));
    }
    int mx = 0;
    for (int r = 0; r < GS; ++r)
        for (int c = 0; c < GS; ++c)
            if (g[r][c] > mx) mx = g[r][c];
    int pr = -1, pc = -1;
    for (int r = 0; r < GS && pr < 0; ++r)
        for (int c = 0; c < GS; ++c)
            if (g[r][c] == mx) { pr = r; pc = c; break; }
    if (is_corner(pr, pc)) best *= 2;
    else best = best / 2; /* best >= 0 so C division == Python // */
    return best;
}

int orc_max_exponent(int g[GS][GS]) {
    int mx = 0;
    for (int r = 0; r < GS; ++r)
        for (int c = 0; c < GS; ++c)
            if (g[r][c] > mx) mx = g[r][c];
    return mx;
}

/* ------------------------------------------------------- A4: spawn / reset */

/* The reference draws through Python's `random` (game.py:937,939).  Parity is
 * by replaying one u32 pair per spawn into both sides:
 *   k  = mulhi32(u0, n_empty)  indexes the row-major empty list (random.choice)
 *   v  = 2 iff u1 >= 3865470567, i.e. iff NOT (u1 / 2^32 < 0.9)   (game.py:939)
 * game.py:923-940 _add_tile.  Returns 0 if the board is full. */
int orc_add_tile(int g[GS][GS], uint32_t u0, uint32_t u1) {
    int er[16], ec[16], n = 0;
    for (int i = 0; i < GS; ++i)
        for (int j = 0; j < GS; ++j)
            if (g[i][j] == 0) { er[n] = i; ec[n] = j; n++; }
    if (n == 0) return 0;
    uint32_t k = (uint32_t)(((uint64_t)u0 * (uint64_t)n) >> 32);
    double x = (double)u1 / 4294967296.0;
    g[er[k]][ec[k]] = x < 0.9 ? 1 : 2;
    return 1;
}

/* ------------------------------------------------------- Philox4x32-10     */
/* Counter-based generator (Salmon et al., SC'11).  Not in the reference; both
 * the CUDA kernels and this oracle derive the spawn/sampling words from it so
 * that the same (seed, env_id, step) yields the same draws on both sides.
 * counter = (env_lo, env_hi, ctr_lo, ctr_hi), key = (seed_lo, seed_hi). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_draws(uint64_t seed, uint64_t env_id, uint64_t ctr, uint32_t out[4]) {
    uint32_t c[4] = {(uint32_t)env_id, (uint32_t)(env_id >> 32), (uint32_t)ctr, (uint32_t)(ctr >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32_10(c, k, out);
}

/* game.py:942-950 reset: empty grid + two spawns.  draws = {u0,u1,u2,u3}:
 * first spawn uses (u0,u1), second (u2,u3). */
uint64_t orc_reset_one(const uint32_t d[4]) {
    int g[GS][GS];
    memset(g, 0, sizeof(g));
    orc_add_tile(g, d[0], d[1]);
    orc_add_tile(g, d[2], d[3]);
    uint64_t b;
    orc_pack(g, &b);
    return b;
}

/* ------------------------------------------------------- A5: step          */

/* game.py:952-1030 step.  `u0,u1` is the spawn draw pair. */
uint64_t orc_step_one(uint64_t board, int dir, uint32_t u0, uint32_t u1, orc_info_t* info) {
    int g[GS][GS];
    orc_unpack(board, g);
    memset(info, 0, sizeof(*info));
    info->legal_before = orc_legal_mask(g);
    if (!orc_direction_has_step(g, dir)) { /* game.py:959-978 */
        info->invalid = 1;
        info->done = !orc_has_next_step(g);
        info->legal_after = info->legal_before;
        return board;
    }
    info->smooth_before = orc_smoothness(g);
    info->corner_before = orc_corner_bonus(g);
    info->mono_before = orc_monotonicity(g);
    info->empt_before = orc_emptiness(g);
    info->max_exp_before = orc_max_exponent(g);

    int ng[GS][GS], pts, mt;
    orc_simulate_move(g, dir, ng, &pts, &mt);
    info->points = pts;
    info->max_tile_created = mt;

    /* potentials after the move, BEFORE the spawn (game.py:994-1002) */
    info->smooth_after = orc_smoothness(ng);
    info->corner_after = orc_corner_bonus(ng);
    info->mono_after = orc_monotonicity(ng);
    info->empt_after = orc_emptiness(ng);
    info->max_exp_after = orc_max_exponent(ng);

    orc_add_tile(ng, u0, u1);                 /* game.py:1005 */
    info->done = !orc_has_next_step(ng);      /* game.py:1006 */
    info->legal_after = orc_legal_mask(ng);
    uint64_t out;
    info->overflow = orc_pack(ng, &out);
    return out;
}

/* ------------------------------------------------------- batch entry points
 * (ctypes-facing; OpenMP over boards so bench.py can use every host core) */

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* draws: replay[n][2] or NULL -> Philox(seed, env0+i, ctr) words 0,1 */
void orc_step_batch(const uint64_t* boards, const uint8_t* actions, const uint32_t* replay,
                    uint64_t seed, uint64_t env0, uint64_t ctr, int64_t n,
                    uint64_t* boards_out, orc_info_t* infos) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint32_t d[4];
        if (replay) { d[0] = replay[2 * i]; d[1] = replay[2 * i + 1]; }
        else orc_draws(seed, env0 + (uint64_t)i, ctr, d);
        boards_out[i] = orc_step_one(boards[i], actions[i], d[0], d[1], &infos[i]);
    }
}

/* 4-move expansion of every board (preview_move_rewards game.py:167-184 +
 * current_valid_directions game.py:295-299): pre-spawn successors, merge
 * points (0 for illegal directions, successor == board there), max tile
 * created and the legal mask. */
void orc_expand4_batch(const uint64_t* boards, int64_t n, uint64_t* succ, int32_t* points,
                       uint8_t* max_tile, uint8_t* legal) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int g[GS][GS];
        orc_unpack(boards[i], g);
        int lm = orc_legal_mask(g);
        legal[i] = (uint8_t)lm;
        for (int d = 0; d < 4; ++d) {
            if (!((lm >> d) & 1)) {
                succ[4 * i + d] = boards[i];
                points[4 * i + d] = 0;
                max_tile[4 * i + d] = 0;
                continue;
            }
            int ng[GS][GS], pts, mt;
            orc_simulate_move(g, d, ng, &pts, &mt);
            uint64_t b;
            orc_pack(ng, &b);
            succ[4 * i + d] = b;
            points[4 * i + d] = pts;
            max_tile[4 * i + d] = (uint8_t)mt;
        }
    }
}

void orc_reset_batch(uint64_t* boards, int64_t n, const uint32_t* replay, uint64_t seed,
                     uint64_t env0, uint64_t ctr) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint32_t d[4];
        if (replay) memcpy(d, replay + 4 * i, sizeof(d));
        else orc_draws(seed, env0 + (uint64_t)i, ctr, d);
        boards[i] = orc_reset_one(d);
    }
}

/* potentials of a batch of boards: out[i] = {mono, empt, smooth, corner, maxexp, legal} */
void orc_potentials_batch(const uint64_t* boards, int64_t n, int32_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int g[GS][GS];
        orc_unpack(boards[i], g);
        out[6 * i + 0] = orc_monotonicity(g);
        out[6 * i + 1] = orc_emptiness(g);
        out[6 * i + 2] = orc_smoothness(g);
        out[6 * i + 3] = orc_corner_bonus(g);
        out[6 * i + 4] = orc_max_exponent(g);
        out[6 * i + 5] = orc_legal_mask(g);
    }
}

/* Row table over all 16^4 rows (index: cell 0 in the low nibble), for the
 * known-answer test in SURVEY.md section 4: out4[i][0..3], score[i], max_tile[i]. */
void orc_row_table(uint8_t* out4, uint32_t* score, uint8_t* max_tile) {
    for (int i = 0; i < 65536; ++i) {
        int row[GS] = {i & 15, (i >> 4) & 15, (i >> 8) & 15, (i >> 12) & 15};
        int o[GS], s, m;
        orc_merge_left(row, o, &s, &m);
        for (int k = 0; k < GS; ++k) out4[4 * i + k] = (uint8_t)o[k];
        score[i] = (uint32_t)s;
        max_tile[i] = (uint8_t)m;
    }
}

/* ------------------------------------------------------- A10: model input  */

/* game.py:92-101 to_model_format: 16 x [exp, row/3, col/3] in float32; the
 * position features are float32(r)/float32(3) as torch computes them. */
void orc_encode_batch(const uint64_t* boards, int64_t n, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        for (int cell = 0; cell < 16; ++cell) {
            int e = (int)((boards[i] >> (4 * cell)) & 0xF);
            out[48 * i + 3 * cell + 0] = (float)e;
            out[48 * i + 3 * cell + 1] = (float)(cell / 4) / 3.0f;
            out[48 * i + 3 * cell + 2] = (float)(cell % 4) / 3.0f;
        }
    }
}

/* ------------------------------------------------------- A14: RTG/advantage */

/* train.py:698-772 + 898-901 calculate_advantage (without the augmentation
 * half), in double like the reference's Python floats.
 *
 * Layout: time-major [T][B] arrays (index t*B + b); `valid[t*B+b]` marks a
 * recorded move, `done[t*B+b]` marks the last move of an episode (the game
 * ended on it), so that several episodes may follow each other in one column
 * (auto-reset).  An episode that is cut by the end of the buffer ends with
 * G = 0 beyond it (train.py:240: max_steps just stops the game).
 *
 * Terminal fix-up of train.py:318-322 (mono_after = empt_after = 0 when the
 * move ended the game) is applied here.
 *
 * moments_io = {rtg_mu, rtg_m2}; rtg_step is 1-indexed (train.py:1705).
 * stats_out  = {sum G, sum G^2, N, batch_mean, batch_var}. */
void orc_rtg_adv(const int32_t* points, const uint8_t* mono_b, const uint8_t* mono_a,
                 const uint8_t* empt_b, const uint8_t* empt_a, const uint8_t* done,
                 const uint8_t* valid, const float* value, int64_t T, int64_t B, double gamma,
                 double w_points, double w_mono, double w_empt, double rtg_beta, int64_t rtg_step,
                 double* moments_io, float* reward_out, float* g_raw_out, float* g_norm_out,
                 float* adv_out, double* stats_out) {
    const double eps = 1e-8;
    double rtg_mu = moments_io[0], rtg_m2 = moments_io[1];
    double bias = 1.0 - pow(rtg_beta, (double)(rtg_step > 1 ? rtg_step : 1));
    if (bias < eps) bias = eps;                               /* train.py:746 */
    double mu_c = rtg_mu / bias;                              /* train.py:749 */
    double m2_c = rtg_m2 / bias;                              /* train.py:752 */
    double var = m2_c - mu_c * mu_c;
    if (var < eps) var = eps;                                 /* train.py:753 */
    double sd = sqrt(var);                                    /* train.py:754 */

    double* G = (double*)__builtin_malloc(sizeof(double) * (size_t)(T * B));
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        double g = 0.0;
        for (int64_t t = T - 1; t >= 0; --t) {
            int64_t i = t * B + b;
            if (!valid[i]) { g = 0.0; G[i] = 0.0; continue; }
            if (done[i]) g = 0.0;                             /* new episode starts after this move */
            double ma = done[i] ? 0.0 : (double)mono_a[i];   /* train.py:318-322 */
            double ea = done[i] ? 0.0 : (double)empt_a[i];
            double pr = (double)points[i] * w_points;          /* train.py:702 */
            double shaped = w_mono * (gamma * ma - (double)mono_b[i]) +
                            w_empt * (gamma * ea - (double)empt_b[i]);   /* train.py:709-714 */
            double r = pr + shaped;                            /* train.py:719 */
            g = r + gamma * g;                                 /* train.py:727 */
            G[i] = g;
            reward_out[i] = (float)r;
            g_raw_out[i] = (float)g;
            double gn = (g - mu_c) / (sd + eps);               /* train.py:760 */
            g_norm_out[i] = (float)gn;
            adv_out[i] = (float)(gn - (double)value[i]);       /* train.py:772 */
        }
    }
    /* batch statistics train.py:732-739 (sequential, like Python's sum) */
    double s1 = 0.0; int64_t N = 0;
    for (int64_t i = 0; i < T * B; ++i)
        if (valid[i]) { s1 += G[i]; N++; }
    double mean = N ? s1 / (double)N : 0.0;
    double v = 0.0, s2 = 0.0;
    for (int64_t i = 0; i < T * B; ++i)
        if (valid[i]) { v += (G[i] - mean) * (G[i] - mean); s2 += G[i] * G[i]; }
    double bvar = N > 1 ? v / (double)N : 0.0;
    stats_out[0] = s1; stats_out[1] = s2; stats_out[2] = (double)N;
    stats_out[3] = mean; stats_out[4] = bvar;
    if (N) {                                                   /* train.py:898-901 */
        moments_io[0] = rtg_beta * rtg_mu + (1.0 - rtg_beta) * mean;
        moments_io[1] = rtg_beta * rtg_m2 + (1.0 - rtg_beta) * (bvar + mean * mean);
    }
    __builtin_free(G);
}

/* ------------------------------------------------------- N3: float potentials (logged only)
 * adjacency_bonus game.py:402-442, monotonic_chain_score game.py:445-506,
 * _choose_anchor_corner game.py:634-668, _get_snake_order game.py:611-632,
 * topological_score game.py:803-921.  Doubles, same operation order as the Python floats. */

double orc_adjacency_bonus(int g[GS][GS]) {
    int mx = 0, mr = 0, mc = 0;
    for (int i = 0; i < GS; ++i)
        for (int j = 0; j < GS; ++j)
            if (g[i][j] > mx) { mx = g[i][j]; mr = i; mc = j; }
    double bonus = 0.0;
    const int di[4] = {-1, 1, 0, 0}, dj[4] = {0, 0, -1, 1};
    for (int d = 0; d < 4; ++d) {
        int ni = mr + di[d], nj = mc + dj[d];
        if (ni >= 0 && ni < GS && nj >= 0 && nj < GS && g[ni][nj] > 0) bonus += g[ni][nj] * 0.5;
    }
    for (int i = 0; i < GS; ++i)
        for (int j = 0; j < GS; ++j)
            if (g[i][j] >= 5) {
                if (j < GS - 1 && g[i][j + 1] >= 5) bonus += (g[i][j] + g[i][j + 1]) * 0.25;
                if (i < GS - 1 && g[i + 1][j] >= 5) bonus += (g[i][j] + g[i + 1][j]) * 0.25;
            }
    return bonus;
}

static double chain_dfs(int g[GS][GS], int i, int j, int expected, int visited[GS][GS]) {
    if (i < 0 || i >= GS || j < 0 || j >= GS) return 0.0;
    if (visited[i][j]) return 0.0;
    if (g[i][j] != expected) return 0.0;
    visited[i][j] = 1;
    double best = 0.0;
    const int di[4] = {-1, 1, 0, 0}, dj[4] = {0, 0, -1, 1};
    for (int d = 0; d < 4; ++d) {
        double c = chain_dfs(g, i + di[d], j + dj[d], expected - 1, visited);
        if (c > best) best = c;
    }
    visited[i][j] = 0;
    return (double)expected + best;
}

double orc_chain_score(int g[GS][GS]) {
    int mx = orc_max_exponent(g);
    if (mx == 0) return 0.0;
    double best = 0.0;
    for (int i = 0; i < GS; ++i)
        for (int j = 0; j < GS; ++j)
            if (g[i][j] == mx) {
                int visited[GS][GS];
                memset(visited, 0, sizeof(visited));
                double s = chain_dfs(g, i, j, mx, visited);
                if (s > best) best = s;
            }
    return best;
}

/* returns 4*row + col of the anchor corner */
int orc_anchor_corner(int g[GS][GS]) {
    const int cr[4] = {0, 0, GS - 1, GS - 1}, cc[4] = {0, GS - 1, 0, GS - 1};
    int mx = 0, pr[16], pc[16], n = 0;
    for (int i = 0; i < GS; ++i)
        for (int j = 0; j < GS; ++j) {
            if (g[i][j] > mx) { mx = g[i][j]; n = 0; pr[n] = i; pc[n] = j; n++; }
            else if (g[i][j] == mx && mx > 0) { pr[n] = i; pc[n] = j; n++; }
        }
    if (n == 0) return 0;
    for (int k = 0; k < n; ++k)
        if (is_corner(pr[k], pc[k])) return 4 * pr[k] + pc[k];
    int best = 0, bd = 1 << 30;
    for (int k = 0; k < 4; ++k) {
        int d = abs(cr[k] - pr[0]) + abs(cc[k] - pc[0]);
        if (d < bd) { bd = d; best = k; }
    }
    return 4 * cr[best] + cc[best];
}

double orc_topological(int g[GS][GS], int anchor) {
    int ntiles = 0, mx = 0;
    for (int i = 0; i < GS; ++i)
        for (int j = 0; j < GS; ++j)
            if (g[i][j] > 0) { ntiles++; if (g[i][j] > mx) mx = g[i][j]; }
    if (!ntiles) return 0.0;
    const int cr = anchor / 4, cc = anchor % 4;
    const int rd = cr == 0 ? 1 : -1, cd = cc == 0 ? 1 : -1;
    int order_r[16], order_c[16], idx_of[GS][GS], n = 0;
    for (int i = 0; i < GS; ++i) {           /* game.py:621-631 snake from the corner */
        int row = cr + i * rd;
        for (int s = 0; s < GS; ++s) {
            int col = (i % 2 == 0) ? cc + s * cd : cc + (GS - 1) * cd - s * cd;
            order_r[n] = row; order_c[n] = col; idx_of[row][col] = n; n++;
        }
    }
    double score = 0.0;
    for (int i = 0; i < GS; ++i)
        for (int j = 0; j < GS; ++j)
            if (g[i][j] > 0) score += (double)((16 - idx_of[i][j]) * g[i][j]) * 0.1;
    double bonus = 0.0, penalty = 0.0, prev = INFINITY;
    for (int k = 0; k < 16; ++k) {
        int v = g[order_r[k]][order_c[k]];
        if (v == 0) continue;
        if ((double)v <= prev) bonus += v * 0.2;
        else penalty += ((double)v - prev) * 0.5;
        prev = (double)v;
    }
    score += bonus - penalty;
    if (g[cr][cc] == mx) score += mx * 2.0;
    const int di[4] = {-1, 1, 0, 0}, dj[4] = {0, 0, -1, 1};
    for (int i = 0; i < GS; ++i)
        for (int j = 0; j < GS; ++j) {
            int v = g[i][j];
            if (v < 4) continue;
            int lower = 0, total = 0;
            for (int d = 0; d < 4; ++d) {
                int ni = i + di[d], nj = j + dj[d];
                if (ni >= 0 && ni < GS && nj >= 0 && nj < GS && g[ni][nj] > 0) {
                    total++;
                    if (g[ni][nj] < v - 2) lower++;
                }
            }
            if (total >= 2 && lower >= total - 1 && idx_of[i][j] > 4) score -= v * 1.0;
        }
    return score;
}

/* out[i] = {adjacency_before, adjacency_after, chain_before, chain_after, topological_before,
 * topological_after (same anchor, game.py:986-1001), anchor (4*row+col)} for (board, pre-spawn successor) */
void orc_potentials_ext_batch(const uint64_t* before, const uint64_t* after, int64_t n, double* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int a[GS][GS], b[GS][GS];
        orc_unpack(before[i], a);
        orc_unpack(after[i], b);
        int anchor = orc_anchor_corner(a);
        double* o = out + 7 * i;
        o[0] = orc_adjacency_bonus(a); o[1] = orc_adjacency_bonus(b);
        o[2] = orc_chain_score(a);     o[3] = orc_chain_score(b);
        o[4] = orc_topological(a, anchor); o[5] = orc_topological(b, anchor);
        o[6] = (double)anchor;
    }
}
