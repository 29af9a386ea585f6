/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement (plain C) of the reference's Craft hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker / the CPU baseline.  The product
 * (psketch_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against
 *   (1) the reference's own golden vectors: all ref_actions sequences of
 *       data/craft_medium_{dev,test}.json plus the regenerated train split (make_data.py,
 *       seed 123), converted to tests/golden/craft_medium_splits.npz, and
 *   (2) outputs of the unmodified reference run in the build container
 *       (oracle/gen_golden.py -> tests/golden/craft_*_states.npz): features, step results for
 *       all six actions, satisfies for all tasks, expert action and closest-resource paths.
 *
 * Each function cites the reference lines it follows.  The grid is held as kind ids
 * (u8, 0 = free, index x*H + y) instead of the reference's one-hot float64 [W,H,K]; the
 * reference asserts exactly one kind per occupied cell (worlds/craft.py:371), so the two
 * forms carry the same information.  Inventory counts are int32.
 *
 * Deliberate, documented deviations from the reference (all outside its tested domain):
 *   - cells outside the grid are treated as blocked / empty instead of numpy's negative-index
 *     wrap-around or IndexError (worlds/craft.py:293,420; teachers/base.py:78);
 *   - the BFS queue is sized to the state space instead of 1000 slots (teachers/base.py:42),
 *     so enlarged grids work;
 *   - find_closest: a reachable goal followed (in scan order) by an unreachable one makes the
 *     reference raise TypeError (teachers/base.py:31); we report status 2 and still return
 *     the closest reachable goal.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_KINDS 32
#define ORC_MAX_RECIPES 16
#define ORC_MAX_TASKS 32
#define ORC_MAX_NODES 16

enum { KC_FREE = 0, KC_INERT = 1, KC_WORKSHOP = 2, KC_WATER = 3, KC_STONE = 4, KC_GRAB = 5 };
enum { SAT_NEVER = 0, SAT_INV = 1, SAT_FACING = 2 };
enum { LEAF_NONE = 0, LEAF_USE = 1, LEAF_GO = 2, LEAF_BAD = 3 };
enum { A_DOWN = 0, A_UP = 1, A_LEFT = 2, A_RIGHT = 3, A_USE = 4, A_STOP = 5 };

typedef struct {
    int32_t W, H, K, win_w, win_h;
    int32_t n_recipes;
    int32_t water, stone, bridge, axe;
    uint8_t kind_class[ORC_MAX_KINDS];
    uint8_t recipes[ORC_MAX_RECIPES][8]; /* out, ws, n_in, in0, cnt0, in1, cnt1, yield */
    uint8_t task_nodes[ORC_MAX_TASKS][ORC_MAX_NODES][4]; /* sat, arg, leaf, skip_to */
    uint8_t task_len[ORC_MAX_TASKS];
} orc_tables;

/* worlds/craft.py:77-91 — coord_change of DOWN, UP, LEFT, RIGHT */
static const int DX[4] = {0, 0, -1, 1};
static const int DY[4] = {-1, 1, 0, 0};

static inline int in_bounds(const orc_tables *t, int x, int y) {
    return x >= 0 && y >= 0 && x < t->W && y < t->H;
}
static inline int cell(const orc_tables *t, const uint8_t *grid, int x, int y) {
    return in_bounds(t, x, y) ? grid[x * t->H + y] : 0;
}
/* movement / navigation: anything off-grid blocks */
static inline int blocked(const orc_tables *t, const uint8_t *grid, int x, int y) {
    return in_bounds(t, x, y) ? grid[x * t->H + y] != 0 : 1;
}

int orc_tables_size(void) { return (int)sizeof(orc_tables); }

/* worlds/craft.py:285-294.  returns 1 True, 0 False, 2 None */
int orc_satisfies_node(const orc_tables *t, const uint8_t *grid, const int32_t *inv, int x,
                       int y, int dir, int sat, int arg) {
    if (sat == SAT_INV) return inv[arg] > 0;
    if (sat == SAT_FACING) return cell(t, grid, x + DX[dir], y + DY[dir]) == arg;
    return 2;
}

int orc_satisfies(const orc_tables *t, const uint8_t *grid, const int32_t *inv, int x, int y,
                  int dir, int task) {
    const uint8_t *n = t->task_nodes[task][0];
    return orc_satisfies_node(t, grid, inv, x, y, dir, n[0], n[1]);
}

/* worlds/craft.py:332-424.  In place.  Returns reward (always 0), or -1 for a bad action
 * (the reference raises Exception, craft.py:415-416). */
int orc_step(const orc_tables *t, uint8_t *grid, int32_t *inv, int *px, int *py, int *pdir,
             int action) {
    int x = *px, y = *py, dx = 0, dy = 0;
    if (action >= A_DOWN && action <= A_RIGHT) { /* craft.py:341-352 */
        dx = DX[action];
        dy = DY[action];
        *pdir = action;
    } else if (action == A_STOP) {
        /* craft.py:353-354 */
    } else if (action == A_USE) { /* craft.py:356-412 */
        int dir = *pdir;
        int nx = x + DX[dir], ny = y + DY[dir];
        /* neighbors() yields the in-bounds cell in front, or nothing (craft.py:426-437) */
        if (in_bounds(t, nx, ny)) {
            int thing = grid[nx * t->H + ny];
            int cls = thing ? t->kind_class[thing] : KC_FREE;
            if (cls == KC_GRAB) { /* craft.py:383-386 */
                inv[thing] += 1;
                grid[nx * t->H + ny] = 0;
            } else if (cls == KC_WORKSHOP) { /* craft.py:388-401: every recipe, in order */
                for (int r = 0; r < t->n_recipes; r++) {
                    const uint8_t *rc = t->recipes[r];
                    if (rc[1] != thing) continue;
                    int ok = 1;
                    for (int j = 0; j < rc[2]; j++)
                        if (inv[rc[3 + 2 * j]] < rc[4 + 2 * j]) ok = 0;
                    if (!ok) continue;
                    inv[rc[0]] += rc[7];
                    for (int j = 0; j < rc[2]; j++) inv[rc[3 + 2 * j]] -= rc[4 + 2 * j];
                }
            } else if (cls == KC_WATER) { /* craft.py:403-406 */
                if (t->bridge && inv[t->bridge] > 0) {
                    grid[nx * t->H + ny] = 0;
                    inv[t->bridge] -= 1;
                }
            } else if (cls == KC_STONE) { /* craft.py:408-410 */
                if (t->axe && inv[t->axe] > 0) grid[nx * t->H + ny] = 0;
            }
        }
    } else {
        return -1;
    }
    /* craft.py:418-421 — tested against the OLD grid; for USE/STOP dx=dy=0 and the agent's own
     * cell is free, so the position is unchanged either way. */
    if (dx || dy) {
        if (!blocked(t, grid, x + dx, y + dy)) {
            *px = x + dx;
            *py = y + dy;
        }
    }
    return 0;
}

/* worlds/craft.py:296-330 with misc/array.py:3-25 (pad_slice) and block_reduce(max).
 * out has n_features = 2*win_w*win_h*K + K + 4 + 1 floats. */
void orc_features(const orc_tables *t, const uint8_t *grid, const int32_t *inv, int x, int y,
                  int dir, float *out) {
    const int K = t->K, ww = t->win_w, wh = t->win_h;
    const int hw = ww / 2, hh = wh / 2;           /* craft.py:299-300 */
    const int bhw = (ww * ww) / 2, bhh = (wh * wh) / 2; /* craft.py:301-302 */
    const int nf = 2 * ww * wh * K + K + 4 + 1;
    memset(out, 0, sizeof(float) * (size_t)nf);
    /* local window, ravel order (dx, dy, k) (craft.py:304-305,324) */
    for (int i = 0; i < ww; i++)
        for (int j = 0; j < wh; j++) {
            int k = cell(t, grid, x - hw + i, y - hh + j);
            if (k) out[(i * wh + j) * K + k] = 1.0f;
        }
    /* big window (2*bhw+1 x 2*bhh+1), max-pooled in (ww x wh) blocks (craft.py:306-310) */
    float *big = out + ww * wh * K;
    for (int i = 0; i < 2 * bhw + 1; i++)
        for (int j = 0; j < 2 * bhh + 1; j++) {
            int k = cell(t, grid, x - bhw + i, y - bhh + j);
            if (k) big[((i / ww) * wh + (j / wh)) * K + k] = 1.0f;
        }
    float *tail = out + 2 * ww * wh * K;
    for (int k = 0; k < K; k++) tail[k] = (float)inv[k]; /* craft.py:325 */
    tail[K + dir] = 1.0f;                                 /* craft.py:321-322 */
    tail[K + 4] = 0.0f;                                   /* craft.py:326 */
}

/* teachers/base.py:36-87 — FIFO BFS over (pos, dir); goal test on dequeue; actions expanded
 * in order DOWN, UP, LEFT, RIGHT; returns path length or -1 (None).  seq (may be NULL)
 * receives the action sequence. */
typedef struct { int16_t x, y, dir; int32_t parent; int8_t act; } bfs_item;

static int shortest_path(const orc_tables *t, const uint8_t *grid, int sx, int sy, int sdir,
                         int gx, int gy, uint8_t *seq, int seq_cap, bfs_item *queue,
                         uint8_t *seen) {
    const int W = t->W, H = t->H;
    memset(seen, 0, (size_t)(W * H * 4));
    int start = 0, end = 0;
    queue[end++] = (bfs_item){(int16_t)sx, (int16_t)sy, (int16_t)sdir, -1, -1};
    seen[(sx * H + sy) * 4 + sdir] = 1;
    while (start < end) {
        int cur = start++;
        bfs_item it = queue[cur];
        if (it.x + DX[it.dir] == gx && it.y + DY[it.dir] == gy) { /* base.py:57-66 */
            int len = 0;
            for (int k = cur; queue[k].parent != -1; k = queue[k].parent) len++;
            if (seq) {
                int p = len;
                for (int k = cur; queue[k].parent != -1; k = queue[k].parent) {
                    p--;
                    if (p < seq_cap) seq[p] = (uint8_t)queue[k].act;
                }
            }
            return len;
        }
        for (int a = 0; a < 4; a++) { /* base.py:68-85 */
            int nx = it.x + DX[a], ny = it.y + DY[a];
            if (blocked(t, grid, nx, ny)) { nx = it.x; ny = it.y; }
            int key = (nx * H + ny) * 4 + a;
            if (!seen[key]) {
                seen[key] = 1;
                queue[end++] = (bfs_item){(int16_t)nx, (int16_t)ny, (int16_t)a, cur, (int8_t)a};
            }
        }
    }
    return -1;
}

static void copy_seq(uint8_t *dst, const uint8_t *src, int len, int cap) {
    int n = len < cap ? len : cap;
    memcpy(dst, src, (size_t)n);
    memset(dst + n, 255, (size_t)(cap - n));
}

/* teachers/base.py:27-34 + craft.py:453-455 (np.nonzero order: x-major, then y).
 * status: 0 = found, 1 = nothing reachable (best_action_seq None), 2 = found, but the
 * reference would have raised TypeError (reachable goal followed by an unreachable one). */
int orc_find_closest(const orc_tables *t, const uint8_t *grid, int x, int y, int dir, int kind,
                     int *gx_out, int *gy_out, int *len_out, uint8_t *seq, int seq_cap) {
    const int W = t->W, H = t->H;
    bfs_item *queue = (bfs_item *)malloc(sizeof(bfs_item) * (size_t)(W * H * 4 + 4));
    uint8_t *seen = (uint8_t *)malloc((size_t)(W * H * 4));
    uint8_t *tmp = seq ? (uint8_t *)malloc((size_t)(seq_cap > 0 ? seq_cap : 1)) : NULL;
    int best = -1, bx = -1, by = -1, raised = 0;
    for (int cx = 0; cx < W; cx++)
        for (int cy = 0; cy < H; cy++) {
            if (grid[cx * H + cy] != kind) continue;
            int len = shortest_path(t, grid, x, y, dir, cx, cy, tmp, seq_cap, queue, seen);
            if (best < 0) {
                /* best_goal[1] is None -> replaced unconditionally, even by None (base.py:31) */
                best = len; bx = cx; by = cy;
                if (seq && len >= 0) copy_seq(seq, tmp, len, seq_cap);
            } else if (len < 0) {
                raised = 1; /* len(None) */
            } else if (len < best) {
                best = len; bx = cx; by = cy;
                if (seq) copy_seq(seq, tmp, len, seq_cap);
            }
        }
    free(queue); free(seen); free(tmp);
    *gx_out = bx; *gy_out = by; *len_out = best;
    if (best < 0) return 1;
    return raised ? 2 : 0;
}

/* teachers/base.py:10-25 on the pre-order flattening (psketch_b200/tables.py): returns the
 * node index of the first incomplete leaf, or -1 (None). */
static int find_incomplete(const orc_tables *t, const uint8_t *grid, const int32_t *inv, int x,
                           int y, int dir, int task) {
    int n = t->task_len[task], i = 0;
    while (i < n) {
        const uint8_t *nd = t->task_nodes[task][i];
        if (orc_satisfies_node(t, grid, inv, x, y, dir, nd[0], nd[1]) == 1) {
            if (nd[2] & 0x40) return -2; /* last subtask of an unsatisfied task: base.py:23-24 asserts */
            i = nd[3];
        } else if ((nd[2] & 0x0F) != LEAF_NONE) return i;
        else i++;
    }
    return -1;
}

/* teachers/demonstration.py:9-30.  Returns the action; 255 where the reference asserts
 * (leaf that is neither 'use' nor 'go'); *dist_out = BFS path length or -1. *status as
 * orc_find_closest (0 when no BFS ran). */
int orc_expert(const orc_tables *t, const uint8_t *grid, const int32_t *inv, int x, int y,
               int dir, int task, int *dist_out, int *status_out) {
    if (dist_out) *dist_out = -1;
    if (status_out) *status_out = 0;
    int leaf = find_incomplete(t, grid, inv, x, y, dir, task);
    if (leaf == -2) return 255;
    if (leaf < 0) return A_STOP;
    const uint8_t *nd = t->task_nodes[task][leaf];
    if ((nd[2] & 0x0F) == LEAF_USE) return A_USE;
    if ((nd[2] & 0x0F) != LEAF_GO) return 255;
    int gx, gy, len;
    uint8_t seq[1];
    int st = orc_find_closest(t, grid, x, y, dir, nd[1], &gx, &gy, &len, seq, 1);
    if (status_out) *status_out = st;
    if (dist_out) *dist_out = len;
    if (st == 1) return A_STOP;
    /* len == 0 cannot occur here: facing the goal means go[X] was satisfied */
    return len > 0 ? seq[0] : 255;
}

/* ---------------------------------------------------------------------------------------
 * Batched forms over SoA arrays (grid u8[N, W*H], inv i32[N, K], pos i32[N,2], dir i32[N]).
 * OpenMP across envs when compiled with -fopenmp; used by the parity tests at large N and by
 * bench.py's native CPU baseline. */
void orc_batch_features(const orc_tables *t, int64_t N, const uint8_t *grid,
                        const int32_t *inv, const int32_t *pos, const int32_t *dir, float *out) {
    const int C = t->W * t->H, K = t->K;
    const int nf = 2 * t->win_w * t->win_h * K + K + 4 + 1;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < N; e++)
        orc_features(t, grid + e * C, inv + e * K, pos[2 * e], pos[2 * e + 1], dir[e],
                     out + e * (int64_t)nf);
}

void orc_batch_step(const orc_tables *t, int64_t N, uint8_t *grid, int32_t *inv, int32_t *pos,
                    int32_t *dir, const int32_t *action, int32_t *status) {
    const int C = t->W * t->H, K = t->K;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < N; e++) {
        int x = pos[2 * e], y = pos[2 * e + 1], d = dir[e];
        status[e] = orc_step(t, grid + e * C, inv + e * K, &x, &y, &d, action[e]);
        pos[2 * e] = x; pos[2 * e + 1] = y; dir[e] = d;
    }
}

void orc_batch_expert(const orc_tables *t, int64_t N, const uint8_t *grid, const int32_t *inv,
                      const int32_t *pos, const int32_t *dir, const int32_t *task,
                      int32_t *action, int32_t *dist, int32_t *status) {
    const int C = t->W * t->H, K = t->K;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t e = 0; e < N; e++) {
        int d, s;
        action[e] = orc_expert(t, grid + e * C, inv + e * K, pos[2 * e], pos[2 * e + 1], dir[e],
                               task[e], &d, &s);
        dist[e] = d; status[e] = s;
    }
}

void orc_batch_satisfies(const orc_tables *t, int64_t N, const uint8_t *grid,
                         const int32_t *inv, const int32_t *pos, const int32_t *dir,
                         const int32_t *task, int32_t *out) {
    const int C = t->W * t->H, K = t->K;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < N; e++)
        out[e] = orc_satisfies(t, grid + e * C, inv + e * K, pos[2 * e], pos[2 * e + 1], dir[e],
                               task[e]);
}

void orc_batch_find_closest(const orc_tables *t, int64_t N, const uint8_t *grid,
                            const int32_t *pos, const int32_t *dir, const int32_t *kind,
                            int32_t *goal, int32_t *len, int32_t *status, uint8_t *seq,
                            int seq_cap) {
    const int C = t->W * t->H;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t e = 0; e < N; e++) {
        int gx, gy, l;
        status[e] = orc_find_closest(t, grid + e * C, pos[2 * e], pos[2 * e + 1], dir[e],
                                     kind[e], &gx, &gy, &l, seq ? seq + e * seq_cap : NULL,
                                     seq_cap);
        goal[2 * e] = gx; goal[2 * e + 1] = gy; len[e] = l;
    }
}

/* One rollout tick per env, the loop of trainers/imitation.py:42-73 with the teacher's action
 * applied (make_data.py:146-152) and features computed for the student
 * (students/imitation.py:72):
 *     a = expert(s); f = features(s); timer -= 1; done = (a == STOP) or timer <= 0
 *     done  -> success = satisfies(task); s = init state; timer = max_timesteps
 *     !done -> s = step(s, a)
 * `ticks` ticks are run per env.  stats[0..3] += episodes, successes, env-steps, sum of feature
 * checksums (keeps the feature computation live). */
void orc_rollout(const orc_tables *t, int64_t N, int ticks, int max_timesteps,
                 const uint8_t *init_grid, const int32_t *init_pos, const int32_t *task,
                 uint8_t *grid, int32_t *inv, int32_t *pos, int32_t *dir, int32_t *timer,
                 float *feat_out /* [N, nf] or NULL */, int32_t *action_out /* [N] last */,
                 int64_t *stats) {
    const int C = t->W * t->H, K = t->K;
    const int nf = 2 * t->win_w * t->win_h * K + K + 4 + 1;
    int64_t episodes = 0, successes = 0, steps = 0;
    double checksum = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : episodes, successes, steps, checksum)
    for (int64_t e = 0; e < N; e++) {
        float *fbuf = feat_out ? feat_out + e * (int64_t)nf : (float *)malloc(sizeof(float) * nf);
        uint8_t *g = grid + e * C;
        int32_t *iv = inv + e * K;
        int x = pos[2 * e], y = pos[2 * e + 1], d = dir[e], tm = timer[e];
        for (int k = 0; k < ticks; k++) {
            int dist, st;
            int a = orc_expert(t, g, iv, x, y, d, task[e], &dist, &st);
            orc_features(t, g, iv, x, y, d, fbuf);
            double s = 0;
            for (int i = 0; i < nf; i++) s += fbuf[i];
            checksum += s;
            steps++;
            tm -= 1;
            int done = (a == A_STOP) || tm <= 0;
            if (done) {
                episodes++;
                successes += orc_satisfies(t, g, iv, x, y, d, task[e]) == 1;
                memcpy(g, init_grid + e * C, (size_t)C);
                memset(iv, 0, sizeof(int32_t) * (size_t)K);
                x = init_pos[2 * e]; y = init_pos[2 * e + 1]; d = 0; tm = max_timesteps;
            } else {
                orc_step(t, g, iv, &x, &y, &d, a);
            }
            if (action_out) action_out[e] = a;
        }
        pos[2 * e] = x; pos[2 * e + 1] = y; dir[e] = d; timer[e] = tm;
        if (!feat_out) free(fbuf);
    }
    stats[0] += episodes; stats[1] += successes; stats[2] += steps;
    stats[3] += (int64_t)checksum;
}
