/*
 * snake_oracle.c — CPU restatement of the reference Snake environment.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this file.  The
 * product path (libsnake_b200.so) never links, loads or calls it.
 *
 * What it restates (file:line are into the reference repository,
 * lucagiorgetti/Laplace-DQN-Snake-game):
 *   structs.jl:33-99    SnakeGame constructor            -> or_game_init
 *   utils.jl:7-10       available_actions                -> or_available_actions
 *   utils.jl:13-40      sample_food!                     -> sample_food
 *   utils.jl:43-52      update_board!                    -> update_board
 *   utils.jl:55-58      check_collision                  -> check_collision
 *   utils.jl:61-81      remove_tail! / grow_maybe!       -> grow_maybe
 *   utils.jl:85-96      move_wrapper!                    -> move_wrapper
 *   utils.jl:100-109    step!                            -> or_game_step
 *   utils.jl:112-132    virtual_step                     -> or_game_virtual_step
 *   utils.jl:135-139    assemble_state!                  -> or_game_assemble_state
 *   utils.jl:141-149    assemble_states_vector           -> or_game_next_state
 *   utils.jl:153-172    epsilon_greedy                   -> or_epsilon_greedy
 *   utils.jl:448-451    masked max-Q target              -> or_masked_target
 *   compute_D.jl:21-31,76-81  Welford fit! + centring    -> or_center_columns
 *
 * Parity pins (see tests/test_oracle_golden.py): G1 food list from the
 * reference's BSON checkpoint, G2 the 237-step score-33 GIF trajectory.
 *
 * Data shapes deliberately follow the reference (dense 10x10 Int64 board in
 * column-major order, head-first snake vector, erase-from-list food, a
 * board_history that grows by one board copy per step, whole-game deep copies
 * in virtual_step) so that the "reference-shaped" mode (keep_history = 1) does
 * the same per-step work the Julia code does.  keep_history = 0 keeps only the
 * last three boards plus the history length; results are identical.
 *
 * Coordinates are Julia's: 1-based (row, col); board[(r-1) + 10*(c-1)].
 * Direction codes follow utils.jl:8: 0=U(-1,0) 1=D(1,0) 2=L(0,-1) 3=R(0,1).
 *
 * Deliberate deviation (documented in DESIGN.md): where the reference throws a
 * BoundsError (sample_food! finds empty cells but no usable list entry,
 * utils.jl:23,37) the oracle sets game->error and leaves the board without
 * food instead of aborting.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OR_BS 10
#define OR_CELLS 100
#define OR_MAX_FOOD 64
#define OR_MAX_SNAKE 128
#define OR_ERR_FOOD 1u      /* R9: BoundsError in the reference */
#define OR_ERR_ACTION 2u    /* action index outside 0..2 / dir outside 0..3 */

typedef struct { int8_t r, c; } or_cell;

typedef struct or_game {
    int64_t board[OR_CELLS];
    int64_t *history;            /* board_history (structs.jl:10) */
    int64_t n_hist, cap_hist;
    int keep_history;
    int64_t ring[3][OR_CELLS];   /* last three boards when !keep_history */
    or_cell snake[OR_MAX_SNAKE]; /* head first (structs.jl:18) */
    int snake_len;
    or_cell direction, prev_dir;
    int64_t score;
    float reward;
    int lost;
    or_cell food_list[OR_MAX_FOOD];
    int n_food;
    uint32_t error;
    int64_t state[2 * OR_CELLS]; /* (10,10,2,1) column-major */
} or_game;

static const or_cell OR_DIRS[4] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};

static inline int64_t *cell_ptr(int64_t *board, or_cell p) { return &board[(p.r - 1) + OR_BS * (p.c - 1)]; }

/* ---- history container ---------------------------------------------------------------- */
static void hist_push(or_game *g, const int64_t *board) {
    if (g->keep_history) {
        if (g->n_hist == g->cap_hist) {
            g->cap_hist = g->cap_hist ? 2 * g->cap_hist : 8;
            g->history = (int64_t *)realloc(g->history, (size_t)g->cap_hist * OR_CELLS * sizeof(int64_t));
        }
        memcpy(g->history + g->n_hist * OR_CELLS, board, OR_CELLS * sizeof(int64_t));
    } else {
        memcpy(g->ring[g->n_hist % 3], board, OR_CELLS * sizeof(int64_t));
    }
    g->n_hist++;
}
/* board_history[end - back] */
static const int64_t *hist_back(const or_game *g, int back) {
    int64_t i = g->n_hist - 1 - back;
    return g->keep_history ? g->history + i * OR_CELLS : g->ring[i % 3];
}

/* ---- structs.jl:33-99 ----------------------------------------------------------------- */
void or_game_init(or_game *g, const uint8_t *food_rc, int n_food, int keep_history) {
    int64_t *keep_buf = g->history;
    int64_t keep_cap = g->cap_hist;
    memset(g, 0, sizeof(*g));
    g->history = keep_buf; g->cap_hist = keep_cap; g->keep_history = keep_history;
    for (int k = 0; k < OR_BS; k++) {                       /* walls, structs.jl:37-40 */
        g->board[0 + OR_BS * k] = -1; g->board[(OR_BS - 1) + OR_BS * k] = -1;
        g->board[k + OR_BS * 0] = -1; g->board[k + OR_BS * (OR_BS - 1)] = -1;
    }
    *cell_ptr(g->board, (or_cell){4, 5}) = 2;               /* structs.jl:43 */
    g->snake[0] = (or_cell){OR_BS - 2, 2};                  /* structs.jl:47 */
    g->snake[1] = (or_cell){OR_BS - 1, 2};
    g->snake_len = 2;
    for (int i = 0; i < g->snake_len; i++) *cell_ptr(g->board, g->snake[i]) = 1;
    hist_push(g, g->board); hist_push(g, g->board);         /* structs.jl:53 */
    memcpy(g->state, g->board, sizeof(g->board));
    memcpy(g->state + OR_CELLS, g->board, sizeof(g->board));
    g->direction = (or_cell){0, 0};                         /* structs.jl:65 */
    g->prev_dir = (or_cell){-1, 0};                         /* structs.jl:66 */
    g->n_food = n_food;                                     /* structs.jl:70 (list injected) */
    for (int i = 0; i < n_food; i++) g->food_list[i] = (or_cell){(int8_t)food_rc[2 * i], (int8_t)food_rc[2 * i + 1]};
}

void or_game_free(or_game *g) { free(g->history); g->history = NULL; g->cap_hist = 0; }

/* deepcopy(game) (utils.jl:122): copies everything, including the whole board_history */
static void game_deepcopy(or_game *dst, const or_game *src) {
    int64_t *buf = dst->history; int64_t cap = dst->cap_hist;
    memcpy(dst, src, sizeof(*dst));
    dst->history = buf; dst->cap_hist = cap;
    if (src->keep_history) {
        if (dst->cap_hist < src->n_hist + 1) {
            dst->cap_hist = src->n_hist + 8;
            dst->history = (int64_t *)realloc(dst->history, (size_t)dst->cap_hist * OR_CELLS * sizeof(int64_t));
        }
        memcpy(dst->history, src->history, (size_t)src->n_hist * OR_CELLS * sizeof(int64_t));
    }
}

/* ---- utils.jl:7-10 -------------------------------------------------------------------- */
int or_available_actions(const or_game *g, uint8_t out_dirs[3]) {
    int n = 0;
    for (int d = 0; d < 4; d++)
        if (!(OR_DIRS[d].r + g->prev_dir.r == 0 && OR_DIRS[d].c + g->prev_dir.c == 0)) {
            if (n < 3) out_dirs[n] = (uint8_t)d;
            n++;
        }
    return n;   /* 3 for any unit prev_dir */
}

/* ---- utils.jl:13-40 ------------------------------------------------------------------- */
static void sample_food(or_game *g) {
    int n_empty = 0;
    for (int k = 0; k < OR_CELLS; k++) n_empty += (g->board[k] == 0);     /* findall(==(0), board) */
    if (n_empty == 0) return;                                           /* utils.jl:18-21 */
    int found = -1;
    for (int i = 0; i < g->n_food; i++) {                               /* for f in food_list */
        if (*cell_ptr(g->board, g->food_list[i]) == 0) { found = i; break; }
    }
    if (found < 0) { g->error |= OR_ERR_FOOD; return; }                 /* board[0] = 2 -> BoundsError */
    or_cell f = g->food_list[found];
    /* findfirst(==(f), food_list) is `found` itself: an earlier equal entry would have matched first */
    memmove(&g->food_list[found], &g->food_list[found + 1], (size_t)(g->n_food - found - 1) * sizeof(or_cell));
    g->n_food--;
    *cell_ptr(g->board, f) = 2;                                         /* utils.jl:37 */
}

/* ---- utils.jl:43-52 ------------------------------------------------------------------- */
static void update_board(or_game *g) {
    for (int k = 0; k < OR_CELLS; k++) if (g->board[k] == 1) g->board[k] = 0;
    for (int i = 0; i < g->snake_len; i++) *cell_ptr(g->board, g->snake[i]) = 1;
}

/* ---- utils.jl:55-58 ------------------------------------------------------------------- */
static int check_collision(const or_game *g) {
    or_cell h = g->snake[0];
    int cnt = 0;
    for (int i = 0; i < g->snake_len; i++) cnt += (g->snake[i].r == h.r && g->snake[i].c == h.c);
    return g->board[(h.r - 1) + OR_BS * (h.c - 1)] == -1 || cnt > 1 ||
           (g->prev_dir.r + g->direction.r == 0 && g->prev_dir.c + g->direction.c == 0);
}

/* ---- utils.jl:61-81 ------------------------------------------------------------------- */
static void grow_maybe(or_game *g) {
    or_cell nh = {(int8_t)(g->snake[0].r + g->direction.r), (int8_t)(g->snake[0].c + g->direction.c)};
    memmove(&g->snake[1], &g->snake[0], (size_t)g->snake_len * sizeof(or_cell));   /* pushfirst! */
    g->snake[0] = nh; g->snake_len++;
    if (*cell_ptr(g->board, nh) == 2) {
        g->score += 1;
        g->reward = 1.0f;                    /* eating_reward, structs.jl:89 */
        sample_food(g);
    } else {
        g->snake_len--;                      /* remove_tail! */
        g->reward = -0.01f;                  /* male_di_vivere, structs.jl:91 */
    }
}

/* ---- utils.jl:85-96 ------------------------------------------------------------------- */
static void move_wrapper(or_game *g) {
    grow_maybe(g);
    if (check_collision(g) || g->n_hist > 500) {
        g->lost = 1;
        g->reward = -1.0f;                   /* suicide_penalty, structs.jl:90 */
    }
    update_board(g);
    g->prev_dir = g->direction;
}

/* ---- utils.jl:100-109 ----------------------------------------------------------------- */
void or_game_step(or_game *g, int dir) {
    g->direction = OR_DIRS[dir & 3];
    move_wrapper(g);
    hist_push(g, g->board);                  /* push!(board_history, deepcopy(board)) */
    /* action/reward/done histories (utils.jl:106-108) are write-only logs; not modelled */
}

/* ---- utils.jl:135-139 ----------------------------------------------------------------- */
void or_game_assemble_state(or_game *g) {
    int off = g->lost ? 1 : 0;               /* lost: board_history[end-2:end-1] */
    memcpy(g->state, hist_back(g, 1 + off), OR_CELLS * sizeof(int64_t));
    memcpy(g->state + OR_CELLS, hist_back(g, off), OR_CELLS * sizeof(int64_t));
}

/* next_states_vec[t] of utils.jl:141-149 = (board_{t-1}, board_t): the last two boards */
void or_game_next_state(const or_game *g, int64_t out[2 * OR_CELLS]) {
    memcpy(out, hist_back(g, 1), OR_CELLS * sizeof(int64_t));
    memcpy(out + OR_CELLS, hist_back(g, 0), OR_CELLS * sizeof(int64_t));
}

/* ---- utils.jl:112-132 ----------------------------------------------------------------- */
/* scratch: three game copies reused across calls (their history buffers are kept) */
typedef struct { or_game v[3]; } or_scratch;

void or_game_virtual_step(or_game *g, or_scratch *sc, uint8_t av_out[3], uint8_t lost_out[3]) {
    if (g->lost) {                                          /* utils.jl:113-117 */
        for (int k = 0; k < 3; k++) { av_out[k] = 255; lost_out[k] = 1; }
        return;
    }
    or_game_assemble_state(g);
    uint8_t av[3];
    or_available_actions(g, av);
    for (int k = 0; k < 3; k++) game_deepcopy(&sc->v[k], g);   /* [deepcopy(game) for _ in 1:3] */
    for (int k = 0; k < 3; k++) {
        or_game_step(&sc->v[k], av[k]);
        lost_out[k] = (uint8_t)sc->v[k].lost;
        av_out[k] = av[k];
        g->error |= sc->v[k].error;                         /* a BoundsError in a copy aborts the caller too */
    }
}

/* ---- utils.jl:153-172 (draws injected: u = Float32(rand()), ridx = index rand() picks) -- */
/* Julia argmax == findmax with isless ordering: first maximal element, NaN > everything,
 * -0.0 < +0.0 (Base reduce.jl, Julia 1.10). */
static int jl_isless_f32(float a, float b) {
    if (isnan(a)) return 0;
    if (isnan(b)) return 1;
    if (a < b) return 1;
    if (a == b) return signbit(a) && !signbit(b);
    return 0;
}
int or_epsilon_greedy_idx(const float q[3], float eps, float u, int ridx) {
    if (u < eps) return ridx;
    int best = 0;
    for (int k = 1; k < 3; k++) if (jl_isless_f32(q[best], q[k])) best = k;
    return best;
}

/* ---- utils.jl:448-451 ----------------------------------------------------------------- */
/* q_next[mask] .= -100 ; maximum over dims=1 ; @. rewards + 0.97 * max * (1 - dones)
 * 0.97 is a Float64 literal, so every element is promoted: y is Float64. */
static float jl_max_f32(float a, float b) {           /* Base.max: NaN-propagating, +0 > -0 */
    if (isnan(a) || isnan(b)) return a + b;
    if (a > b) return a;
    if (b > a) return b;
    return signbit(a) ? b : a;
}
void or_masked_target(const float *q_next /*3xB col-major*/, const uint8_t *mask /*3xB*/, const float *r,
                      const uint8_t *done, double gamma, float fill, double *y, int64_t B) {
    for (int64_t i = 0; i < B; i++) {
        float q0 = mask[3 * i + 0] ? fill : q_next[3 * i + 0];
        float q1 = mask[3 * i + 1] ? fill : q_next[3 * i + 1];
        float q2 = mask[3 * i + 2] ? fill : q_next[3 * i + 2];
        float m = jl_max_f32(jl_max_f32(q0, q1), q2);
        volatile double t = gamma * (double)m;          /* (0.97 * max) * (1 - done), left to right */
        t = t * (double)(1 - (int)(done[i] != 0));
        y[i] = (double)r[i] + t;
    }
}

/* ---- compute_D.jl:21-31, 76-81 (la_utils.jl:26-36, 163-169) ---------------------------- */
/* D is P x K column-major Float64.  Welford over columns, then D .-= mean.
 * var = m2 ./ max(n-1, 1).  Compiled with -ffp-contract=off: Julia does not fuse. */
void or_center_columns(double *D, int64_t P, int64_t K, double *mean, double *var) {
    double *m2 = (double *)calloc((size_t)P, sizeof(double));
    for (int64_t p = 0; p < P; p++) mean[p] = 0.0;
    for (int64_t k = 0; k < K; k++) {
        const double *x = D + k * P;
        double n = (double)(k + 1);
        for (int64_t p = 0; p < P; p++) {
            double d = x[p] - mean[p];
            mean[p] = mean[p] + d / n;
            m2[p] = m2[p] + d * (x[p] - mean[p]);
        }
    }
    double den = (double)(K - 1 > 1 ? K - 1 : 1);
    for (int64_t p = 0; p < P; p++) var[p] = m2[p] / den;
    for (int64_t k = 0; k < K; k++)
        for (int64_t p = 0; p < P; p++) D[k * P + p] = D[k * P + p] - mean[p];
    free(m2);
}

/* ======================================================================================= */
/* Batched driver with the same call semantics as include/snake_b200.h, so the parity tests */
/* can compare whole output arrays.  One or_game per env; auto-reset = a fresh SnakeGame()   */
/* (utils.jl:199) after the terminal outputs have been emitted.                              */
/* ======================================================================================= */
typedef struct or_batch {
    int64_t n;
    int auto_reset, keep_history;
    or_game *g;
    or_scratch sc;
    uint8_t food_rc[2 * OR_MAX_FOOD];
    int n_food;
    float *ep_return;          /* Float32 running sum, utils.jl:200,207 */
    uint8_t *mask;             /* last virtual_step result per env, 3 per env */
    uint8_t *has_mask;
} or_batch;

or_batch *or_batch_create(int64_t n, const uint8_t *food_rc, int n_food, int auto_reset, int keep_history) {
    or_batch *b = (or_batch *)calloc(1, sizeof(or_batch));
    b->n = n; b->auto_reset = auto_reset; b->keep_history = keep_history; b->n_food = n_food;
    memcpy(b->food_rc, food_rc, (size_t)(2 * n_food));
    b->g = (or_game *)calloc((size_t)n, sizeof(or_game));
    b->ep_return = (float *)calloc((size_t)n, sizeof(float));
    b->mask = (uint8_t *)calloc((size_t)(3 * n), 1);
    b->has_mask = (uint8_t *)calloc((size_t)n, 1);
    for (int64_t i = 0; i < n; i++) or_game_init(&b->g[i], food_rc, n_food, keep_history);
    return b;
}
void or_batch_destroy(or_batch *b) {
    for (int64_t i = 0; i < b->n; i++) or_game_free(&b->g[i]);
    for (int k = 0; k < 3; k++) or_game_free(&b->sc.v[k]);
    free(b->g); free(b->ep_return); free(b->mask); free(b->has_mask); free(b);
}
void or_batch_reset(or_batch *b) {
    for (int64_t i = 0; i < b->n; i++) {
        or_game_init(&b->g[i], b->food_rc, b->n_food, b->keep_history);
        b->ep_return[i] = 0.0f; b->has_mask[i] = 0;
    }
}

static float cell_to_f32(int64_t v) { return (float)v; }

/* One batched step.  action: idx 0..2 into available_actions (is_abs = 0) or absolute dir
 * 0..3 (is_abs = 1).  All outputs optional.  obs_* receive next_state (board_{t-1}, board_t)
 * laid out (10,10,2,N) column-major.  ep_ret / ep_score receive the episode's Float32 return
 * and score after this step.  An env that is already lost (auto_reset off) is left untouched
 * and reports reward 0, done 1, mask 1 1 1. */
void or_batch_step(or_batch *b, const uint8_t *action, int is_abs, float *reward, uint8_t *done,
                   float *obs_f32, int8_t *obs_i8, int64_t *obs_i64, uint8_t *mask,
                   float *ep_ret, int32_t *ep_score, uint8_t *av_next) {
    int64_t tmp[2 * OR_CELLS];
    for (int64_t i = 0; i < b->n; i++) {
        or_game *g = &b->g[i];
        uint8_t lost3[3] = {1, 1, 1}, av3[3] = {255, 255, 255};
        float r = 0.0f;
        if (!g->lost) {
            int dir;
            if (is_abs) {
                dir = action[i];
                if (dir > 3) { g->error |= OR_ERR_ACTION; dir = 0; }
            } else {
                uint8_t av[3];
                or_available_actions(g, av);
                int idx = action[i];
                if (idx > 2) { g->error |= OR_ERR_ACTION; idx = 0; }
                dir = av[idx];
            }
            or_game_step(g, dir);
            or_game_virtual_step(g, &b->sc, av3, lost3);
            r = g->reward;
            b->ep_return[i] += r;
        }
        if (reward) reward[i] = r;
        if (done) done[i] = (uint8_t)g->lost;
        if (mask) memcpy(mask + 3 * i, lost3, 3);
        if (av_next) memcpy(av_next + 3 * i, av3, 3);
        memcpy(b->mask + 3 * i, lost3, 3); b->has_mask[i] = 1;
        if (ep_ret) ep_ret[i] = b->ep_return[i];
        if (ep_score) ep_score[i] = (int32_t)g->score;
        if (obs_f32 || obs_i8 || obs_i64) {
            or_game_next_state(g, tmp);
            for (int k = 0; k < 2 * OR_CELLS; k++) {
                if (obs_f32) obs_f32[i * 200 + k] = cell_to_f32(tmp[k]);
                if (obs_i8) obs_i8[i * 200 + k] = (int8_t)tmp[k];
                if (obs_i64) obs_i64[i * 200 + k] = tmp[k];
            }
        }
        if (g->lost && b->auto_reset) {
            uint32_t err = g->error;
            or_game_init(g, b->food_rc, b->n_food, b->keep_history);
            g->error = err;                 /* error bits are sticky across episodes */
            b->ep_return[i] = 0.0f;
        }
    }
}

/* current two-frame state (board_{t-1}, board_t) of every env */
void or_batch_state(or_batch *b, float *obs_f32, int8_t *obs_i8, int64_t *obs_i64) {
    int64_t tmp[2 * OR_CELLS];
    for (int64_t i = 0; i < b->n; i++) {
        or_game_next_state(&b->g[i], tmp);
        for (int k = 0; k < 2 * OR_CELLS; k++) {
            if (obs_f32) obs_f32[i * 200 + k] = cell_to_f32(tmp[k]);
            if (obs_i8) obs_i8[i * 200 + k] = (int8_t)tmp[k];
            if (obs_i64) obs_i64[i * 200 + k] = tmp[k];
        }
    }
}

/* losing mask of the CURRENT state of every env (virtual_step on demand) */
void or_batch_losing_mask(or_batch *b, uint8_t *mask, uint8_t *av_next) {
    for (int64_t i = 0; i < b->n; i++) {
        uint8_t lost3[3], av3[3];
        or_game_virtual_step(&b->g[i], &b->sc, av3, lost3);
        memcpy(mask + 3 * i, lost3, 3);
        if (av_next) memcpy(av_next + 3 * i, av3, 3);
    }
}

void or_batch_available_actions(or_batch *b, uint8_t *out) {
    for (int64_t i = 0; i < b->n; i++) or_available_actions(&b->g[i], out + 3 * i);
}

void or_batch_select(or_batch *b, const float *q, float eps, const float *u, const uint8_t *ridx, uint8_t *out) {
    for (int64_t i = 0; i < b->n; i++) out[i] = (uint8_t)or_epsilon_greedy_idx(q + 3 * i, eps, u[i], ridx[i]);
}

/* per-env scalars for cross-checking */
void or_batch_scalars(or_batch *b, int32_t *score, uint8_t *lost, uint32_t *error, int32_t *snake_len,
                      uint8_t *head_rc, uint8_t *food_rc, int32_t *n_food_left, int32_t *n_hist) {
    for (int64_t i = 0; i < b->n; i++) {
        or_game *g = &b->g[i];
        if (score) score[i] = (int32_t)g->score;
        if (lost) lost[i] = (uint8_t)g->lost;
        if (error) error[i] = g->error;
        if (snake_len) snake_len[i] = g->snake_len;
        if (head_rc) { head_rc[2 * i] = (uint8_t)g->snake[0].r; head_rc[2 * i + 1] = (uint8_t)g->snake[0].c; }
        if (food_rc) {
            food_rc[2 * i] = food_rc[2 * i + 1] = 0;
            for (int k = 0; k < OR_CELLS; k++) if (g->board[k] == 2) { food_rc[2 * i] = (uint8_t)(k % 10 + 1); food_rc[2 * i + 1] = (uint8_t)(k / 10 + 1); }
        }
        if (n_food_left) n_food_left[i] = g->n_food;
        if (n_hist) n_hist[i] = (int32_t)g->n_hist;
    }
}

/* single-game accessors for the golden-trajectory tests */
or_game *or_game_new(const uint8_t *food_rc, int n_food, int keep_history) {
    or_game *g = (or_game *)calloc(1, sizeof(or_game));
    or_game_init(g, food_rc, n_food, keep_history);
    return g;
}
void or_game_delete(or_game *g) { or_game_free(g); free(g); }
void or_game_get_board(const or_game *g, int64_t out[OR_CELLS]) { memcpy(out, g->board, sizeof(g->board)); }
int64_t or_game_score(const or_game *g) { return g->score; }
int or_game_lost(const or_game *g) { return g->lost; }
float or_game_reward(const or_game *g) { return g->reward; }
uint32_t or_game_error(const or_game *g) { return g->error; }
int or_game_n_food(const or_game *g) { return g->n_food; }
int64_t or_game_n_hist(const or_game *g) { return g->n_hist; }
void or_game_get_state(const or_game *g, int64_t out[2 * OR_CELLS]) { memcpy(out, g->state, sizeof(g->state)); }
size_t or_scratch_size(void) { return sizeof(or_scratch); }

/* ---- CPU baseline loop (bench.py cpu_baseline / --impl reference) ---------------------- */
/* Steps envs [lo,hi) for n_steps with actions idx = splitmix64(seed, env, t) % 3, producing
 * the same per-step outputs as config 2 (f32 obs, mask, reward, done) into per-thread
 * scratch.  Returns a checksum so the work cannot be optimised away. */
static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
uint64_t or_batch_run_random(or_batch *b, int64_t lo, int64_t hi, int64_t n_steps, uint64_t seed) {
    or_scratch *sc = (or_scratch *)calloc(1, sizeof(or_scratch));
    float obs[200];
    int64_t tmp[200];
    uint64_t sum = 0;
    for (int64_t t = 0; t < n_steps; t++) {
        for (int64_t i = lo; i < hi; i++) {
            or_game *g = &b->g[i];
            uint8_t av[3], av3[3], lost3[3];
            or_available_actions(g, av);
            uint64_t x = splitmix64(seed ^ splitmix64((uint64_t)i * 0x100000001B3ull + (uint64_t)t));
            or_game_step(g, av[x % 3]);
            or_game_virtual_step(g, sc, av3, lost3);
            or_game_next_state(g, tmp);
            for (int k = 0; k < 200; k++) obs[k] = (float)tmp[k];      /* Float32.(s), utils.jl:361-362 */
            sum += (uint64_t)(obs[37] + obs[163] + 2.0f) + lost3[0] + 2 * lost3[1] + 4 * lost3[2] + (uint64_t)g->lost;
            if (g->lost) or_game_init(g, b->food_rc, b->n_food, b->keep_history);
        }
    }
    for (int k = 0; k < 3; k++) or_game_free(&sc->v[k]);
    free(sc);
    return sum;
}
