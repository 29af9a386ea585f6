/*
 * TEST INFRASTRUCTURE ONLY — CPU restatement (plain C) of the reference's Light world
 * (worlds/light.py) plus a brute-force shortest-plan search used to validate the GPU teacher.
 *
 * Parity status: step / features / satisfies PINNED against states exported from the unmodified
 * reference (tests/golden/light_states.npz, oracle/gen_golden.py:export_light).  The teacher has
 * no reference implementation ("parity unpinned" for psk_light_expert): orc_light_expert is the
 * specification's brute force — breadth-first search over (x, y, key mask) with the oracle's own
 * step function, first action = the smallest action index that starts some shortest plan.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ROOM 6
#define MAXB 32
#define MAXD 8
#define MAXK 8

typedef struct {
    int32_t bw, bh, n_doors, n_keys, goal_rx, goal_ry;
    uint8_t walls[MAXB][MAXB]; /* [x][y], cells beyond the board = 1 */
    int32_t doors[MAXD][2];
    int32_t keys[MAXK][4]; /* kx, ky, door x, door y */
} orc_light_scen;

int orc_light_scen_size(void) { return (int)sizeof(orc_light_scen); }

static const int DX[4] = {0, 0, -1, 1};
static const int DY[4] = {-1, 1, 0, 0};

static int door_locked(const orc_light_scen *s, int x, int y, unsigned alive) {
    int is_door = 0;
    for (int d = 0; d < s->n_doors; d++) is_door |= s->doors[d][0] == x && s->doors[d][1] == y;
    if (!is_door) return 0;
    for (int k = 0; k < s->n_keys; k++)
        if (((alive >> k) & 1) && s->keys[k][2] == x && s->keys[k][3] == y) return 1;
    return 0;
}

/* worlds/light.py:212-235.  returns -1 for an action the reference cannot handle. */
int orc_light_step(const orc_light_scen *s, int *px, int *py, unsigned *palive, int action) {
    int x = *px, y = *py, dx = 0, dy = 0;
    unsigned alive = *palive, n_alive = alive;
    if (action >= 0 && action < 4) {
        dx = DX[action];
        dy = DY[action];
    } else if (action == 4) {
        for (int k = 0; k < s->n_keys; k++)
            if (((alive >> k) & 1) && s->keys[k][0] == x && s->keys[k][1] == y) n_alive &= ~(1u << k);
    } else {
        return -1;
    }
    int nx = x + dx, ny = y + dy;
    if (nx < 0 || ny < 0 || nx >= MAXB || ny >= MAXB || s->walls[nx][ny]) { nx = x; ny = y; }
    if (door_locked(s, nx, ny, alive)) { nx = x; ny = y; } /* old key set, light.py:233 */
    *px = nx; *py = ny; *palive = n_alive;
    return 0;
}

/* strength of a door/key map at distance (ax, ay): 10 - sqrt(.), clipped at 0, floor-divided by
 * 10 (worlds/light.py:120-122) */
static double strength(int ax, int ay) {
    double v = 10.0 - sqrt((double)(ax * ax + ay * ay));
    if (v < 0) v = 0;
    return floor(v / 10.0);
}

static int fdiv(int a, int b) { /* python floor division */
    int q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) q--;
    return q;
}

/* worlds/light.py:191-204 evaluating the maps of :105-146 at the agent's cell */
void orc_light_features(const orc_light_scen *s, int x, int y, unsigned alive, float *out) {
    for (int i = 0; i < 12; i++) out[i] = 0.f;
    const int rx = x / ROOM, ry = y / ROOM;
    for (int d = 0; d < s->n_doors; d++) {
        const int dx = s->doors[d][0], dy = s->doors[d][1];
        double f[4] = {0, 0, 0, 0};
        do {
            if (rx != fdiv(dx + 1, ROOM) && rx != fdiv(dx - 1, ROOM)) break;
            if (ry != fdiv(dy + 1, ROOM) && ry != fdiv(dy - 1, ROOM)) break;
            if (!(x == dx && y == dy) && (x % ROOM == 0 || y % ROOM == 0)) break;
            const double st = strength(x - dx, y - dy);
            if (dx <= x) f[0] += st;
            if (dx >= x) f[1] += st;
            if (dy <= y) f[2] += st;
            if (dy >= y) f[3] += st;
        } while (0);
        const int base = door_locked(s, dx, dy, alive) ? 0 : 4;
        for (int i = 0; i < 4; i++) out[base + i] += (float)f[i];
    }
    for (int k = 0; k < s->n_keys; k++) {
        if (!((alive >> k) & 1)) continue;
        const int kx = s->keys[k][0], ky = s->keys[k][1];
        if (kx / ROOM != rx || ky / ROOM != ry) continue;
        if (x % ROOM == 0 || y % ROOM == 0) continue;
        const double st = strength(x - kx, y - ky);
        if (kx <= x) out[8] += (float)st;
        if (kx >= x) out[9] += (float)st;
        if (ky <= y) out[10] += (float)st;
        if (ky >= y) out[11] += (float)st;
    }
}

int orc_light_satisfies(const orc_light_scen *s, int x, int y) {
    return x / ROOM == s->goal_rx && y / ROOM == s->goal_ry; /* light.py:208-210 */
}

/* Brute force: BFS over (x, y, mask) from the state; *dist = fewest actions to the goal room
 * (0 if already there, -1 if unreachable); returns the smallest first action of a shortest plan
 * (254 already there, 255 unreachable). */
int orc_light_expert(const orc_light_scen *s, int x0, int y0, unsigned alive0, int *dist) {
    if (orc_light_satisfies(s, x0, y0)) { *dist = 0; return 254; }
    const int NS = MAXB * MAXB * 256;
    int32_t *d = (int32_t *)malloc(sizeof(int32_t) * NS);
    int32_t *queue = (int32_t *)malloc(sizeof(int32_t) * NS);
    int best = -1, best_a = 255;
    for (int a = 0; a < 5; a++) {
        int x = x0, y = y0;
        unsigned m = alive0;
        orc_light_step(s, &x, &y, &m, a);
        if (x == x0 && y == y0 && m == alive0) continue; /* no progress */
        /* distance from the successor to the goal room */
        for (int i = 0; i < NS; i++) d[i] = -1;
        int head = 0, tail = 0, found = -1;
        int id = (x * MAXB + y) * 256 + (int)m;
        d[id] = 0;
        queue[tail++] = id;
        while (head < tail) {
            int cur = queue[head++];
            int cx = cur / 256 / MAXB, cy = cur / 256 % MAXB;
            unsigned cm = (unsigned)(cur % 256);
            if (orc_light_satisfies(s, cx, cy)) { found = d[cur]; break; }
            for (int b = 0; b < 5; b++) {
                int nx = cx, ny = cy;
                unsigned nm = cm;
                orc_light_step(s, &nx, &ny, &nm, b);
                int nid = (nx * MAXB + ny) * 256 + (int)nm;
                if (d[nid] < 0) { d[nid] = d[cur] + 1; queue[tail++] = nid; }
            }
        }
        if (found >= 0 && (best < 0 || found + 1 < best)) { best = found + 1; best_a = a; }
    }
    free(d); free(queue);
    *dist = best;
    return best_a;
}

/* batched wrappers: state i32[n][3] = x, y, alive mask */
void orc_light_batch(const orc_light_scen *scen, const int32_t *scen_idx, int64_t n,
                     const int32_t *state, const int32_t *action, int32_t *state_out,
                     float *feat_out, int32_t *sat_out) {
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; e++) {
        const orc_light_scen *s = scen + scen_idx[e];
        int x = state[3 * e], y = state[3 * e + 1];
        unsigned m = (unsigned)state[3 * e + 2];
        if (feat_out) orc_light_features(s, x, y, m, feat_out + 12 * e);
        if (sat_out) sat_out[e] = orc_light_satisfies(s, x, y);
        if (action && state_out) {
            int st = orc_light_step(s, &x, &y, &m, action[e]);
            state_out[3 * e] = st < 0 ? -1 : x;
            state_out[3 * e + 1] = y;
            state_out[3 * e + 2] = (int32_t)m;
        }
    }
}

void orc_light_batch_expert(const orc_light_scen *scen, const int32_t *scen_idx, int64_t n,
                            const int32_t *state, int32_t *action, int32_t *dist) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t e = 0; e < n; e++) {
        int d;
        action[e] = orc_light_expert(scen + scen_idx[e], state[3 * e], state[3 * e + 1],
                                     (unsigned)state[3 * e + 2], &d);
        dist[e] = d;
    }
}
