// psk_craft.cu — sm_100a kernels of the batched Craft environment + BFS teacher and their C ABI.
//
// Kernels (DESIGN.md has the roofline of each):
//   craft_step_kernel      one env per thread; agent record moved as one 256-bit LDG/STG, one
//                          byte gathered from the grid row, recipe table in shared memory.
//   craft_features_kernel  one CTA per tile of E envs; the f32 feature tile is built in shared
//                          memory (zero-fill + scatter of the ~15 non-zeros per env) and leaves
//                          with ONE TMA bulk store (cp.async.bulk, UBLKCP) per tile,
//                          double-buffered so the store of tile i overlaps the build of i+1.
//   craft_expert_kernel    one env per thread; hint-tree walk, then a level-synchronous BFS over
//                          (pos, dir) on 64/128-bit bitboards carrying 4 "first action" colours.
//   craft_advance_kernel   rollout bookkeeping (timer / done / success / auto-reset) + step.
//   craft_tick_kernel      expert + features + advance fused: state read once per tick.
#include <cstdio>
#include <cstring>

#include "psk_common.cuh"

namespace psk {

// =============================================================================================
// step  (worlds/craft.py:332-424)
// =============================================================================================
// Applies `act` to the agent record `a` and the env's grid row (global or shared memory).
// Returns true when the grid row was modified.  flags collects PSK_FLAG_* bits.
template <int W, int H>
__device__ __forceinline__ void step_env(const SharedTables &st, Agent &a, uint8_t *row, int act,
                                         uint32_t &flags) {
    int x = a.x(), y = a.y(), dir = a.dir();
    if (act < 4) {  // craft.py:341-352 + 418-421: turn always, move iff the target cell is free
        dir = act;
        const int tx = x + dx_of(act), ty = y + dy_of(act);
        if (tx >= 0 && ty >= 0 && tx < W && ty < H && row[tx * H + ty] == 0) {
            x = tx;
            y = ty;
        }
        a.set_pose(x, y, dir);
    } else if (act == PSK_ACT_USE) {  // craft.py:356-412
        const int fx = x + dx_of(dir), fy = y + dy_of(dir);
        if (fx >= 0 && fy >= 0 && fx < W && fy < H) {  // neighbors(), craft.py:426-437
            const int thing = row[fx * H + fy];
            const int cls = st.kind_class(thing);
            const psk_craft_tables &T = st.t();
            if (cls == KC_GRAB) {  // craft.py:383-386
                if (thing < PSK_MAX_INV) {
                    if (a.inv(thing) == 255) flags |= PSK_FLAG_INV_OVERFLOW;
                    else a.inv_add(thing, 1);
                }
                row[fx * H + fy] = 0;
            } else if (cls == KC_WORKSHOP) {  // craft.py:388-401: all recipes, in file order
                for (int r = 0; r < T.n_recipes; r++) {
                    const uint2 rc = st.recipe(r);
                    const int out = rc.x & 0xFF, ws = (rc.x >> 8) & 0xFF, n_in = (rc.x >> 16) & 0xFF;
                    if (ws != thing) continue;
                    const int in0 = rc.x >> 24, c0 = rc.y & 0xFF;
                    const int in1 = (rc.y >> 8) & 0xFF, c1 = (rc.y >> 16) & 0xFF, yld = rc.y >> 24;
                    if (n_in > 0 && a.inv(in0) < c0) continue;
                    if (n_in > 1 && a.inv(in1) < c1) continue;
                    if (a.inv(out) + yld > 255) { flags |= PSK_FLAG_INV_OVERFLOW; continue; }
                    a.inv_add(out, yld);
                    if (n_in > 0) a.inv_add(in0, -c0);
                    if (n_in > 1) a.inv_add(in1, -c1);
                }
            } else if (cls == KC_WATER) {  // craft.py:403-406
                if (T.bridge_kind && a.inv(T.bridge_kind) > 0) {
                    row[fx * H + fy] = 0;
                    a.inv_add(T.bridge_kind, -1);
                }
            } else if (cls == KC_STONE) {  // craft.py:408-410 (the axe is kept)
                if (T.axe_kind && a.inv(T.axe_kind) > 0) row[fx * H + fy] = 0;
            }
        }
    } else if (act != PSK_ACT_STOP) {
        flags |= PSK_FLAG_BAD_ACTION;  // craft.py:415-416
    }
}

template <int W, int H>
__global__ void __launch_bounds__(256)
craft_step_kernel(const __grid_constant__ psk_craft_tables T, uint8_t *__restrict__ grid,
                  uint8_t *__restrict__ agent, const uint8_t *__restrict__ action,
                  const uint8_t *__restrict__ active, float *__restrict__ reward,
                  int32_t *err_flags, int64_t n, int cell_stride) {
    __shared__ SharedTables st;
    stage_tables(st, T);
    uint32_t flags = 0;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
         e += (int64_t)gridDim.x * blockDim.x) {
        if (reward) reward[e] = 0.0f;  // craft.py:338,424
        if (active && !active[e]) continue;
        Agent a = load_agent(agent, e);
        const Agent before = a;
        step_env<W, H>(st, a, grid + e * cell_stride, action[e], flags);
        bool changed = false;
#pragma unroll
        for (int i = 0; i < 8; i++) changed |= a.w[i] != before.w[i];
        if (changed) store_agent(agent, e, a);
    }
    if (flags && err_flags) atomicOr(err_flags, (int)flags);
}

// =============================================================================================
// satisfies  (worlds/craft.py:285-294)  and the hint-tree walk (teachers/base.py:10-25)
// =============================================================================================
template <int W, int H>
__device__ __forceinline__ int facing_kind(const Agent &a, const uint8_t *row) {
    const int fx = a.x() + dx_of(a.dir()), fy = a.y() + dy_of(a.dir());
    if (fx < 0 || fy < 0 || fx >= W || fy >= H) return 0;
    return row[fx * H + fy];
}

__device__ __forceinline__ int node_satisfied(uint32_t nd, const Agent &a, int facing) {
    const int sat = nd & 0xFF, arg = (nd >> 8) & 0xFF;
    if (sat == SAT_INV) return arg < PSK_MAX_INV ? (a.inv(arg) > 0) : 0;
    if (sat == SAT_FACING) return facing == arg;
    return 2;  // None
}

// first incomplete leaf of `task` (node word), or 0 when there is none (-> STOP)
__device__ __forceinline__ uint32_t find_incomplete(const SharedTables &st, int task,
                                                    const Agent &a, int facing) {
    const int n = st.task_len(task);
    int i = 0;
    while (i < n) {
        const uint32_t nd = st.node(task, i);
        if (node_satisfied(nd, a, facing) == 1) i = nd >> 24;
        else if ((nd >> 16) & 0xFF) return nd | 0x80000000u;  // leaf (bit 31 marks "found")
        else i++;
    }
    return 0;
}

template <int W, int H>
__global__ void __launch_bounds__(256)
craft_satisfies_kernel(const __grid_constant__ psk_craft_tables T,
                       const uint8_t *__restrict__ grid, const uint8_t *__restrict__ agent,
                       const uint8_t *__restrict__ task, uint8_t *__restrict__ out, int64_t n,
                       int cell_stride) {
    __shared__ SharedTables st;
    stage_tables(st, T);
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
         e += (int64_t)gridDim.x * blockDim.x) {
        const Agent a = load_agent_ro(agent, e);
        const int tk = task ? task[e] : a.task();
        const int facing = facing_kind<W, H>(a, grid + e * cell_stride);
        out[e] = st.task_len(tk) ? node_satisfied(st.node(tk, 0), a, facing) : 2;
    }
}

// =============================================================================================
// expert  (teachers/demonstration.py:9-30, teachers/base.py:27-87)
// =============================================================================================
// The reference runs one FIFO BFS per goal cell over states (pos, dir), expanding DOWN, UP,
// LEFT, RIGHT, keeps the strictly shortest path (first goal cell in x-major order wins ties) and
// returns its first action.  Because the successor of (p, d) under action a is (p', a) with p'
// independent of d, and because the FIFO tree path of a node is the lexicographically smallest
// shortest path to it, the answer has an order-independent form that maps onto bitboards:
//   * one level-synchronous BFS; V[a] = visited positions with dir a;
//   * P[f] = frontier positions that some shortest path starting with action f reaches;
//   * the first level L at which a new state faces a goal cell gives the distance; the goal is
//     the lowest-index goal cell hit at L; the action is the smallest f whose colour faces it.
// Returns the path length (-1: unreachable); first (0..3) is valid when the length is > 0.
template <int W, int H>
__device__ __forceinline__ int bfs_first_action(typename Board<W, H>::BT occ,
                                                typename Board<W, H>::BT goal, int x, int y,
                                                int d0, int &first, int &goal_idx) {
    using B = Board<W, H>;
    using BT = typename B::BT;
    const BT freeb = ~occ & B::all();
    // cm[a]: positions whose a-neighbour is on the grid and free (the agent moves);
    // T[a]:  positions whose a-neighbour is a goal cell (facing the goal with dir a).
    const BT cm[4] = {B::template unshift<0>(freeb), B::template unshift<1>(freeb),
                      B::template unshift<2>(freeb), B::template unshift<3>(freeb)};
    const BT Tg[4] = {B::template unshift<0>(goal), B::template unshift<1>(goal),
                      B::template unshift<2>(goal), B::template unshift<3>(goal)};
    const BT root = B::bit(x * H + y);
    BT V[4];
#pragma unroll
    for (int a = 0; a < 4; a++) V[a] = (a == d0) ? root : BT(0);
    first = -1;
    goal_idx = -1;
    {  // level 0: already facing a goal cell (teachers/base.py:57-66 on the first dequeue)
        BT t = 0;
#pragma unroll
        for (int a = 0; a < 4; a++) t |= (a == d0) ? (Tg[a] & root) : BT(0);
        if (t) {
            goal_idx = (x + dx_of(d0)) * H + (y + dy_of(d0));
            return 0;
        }
    }
    BT P[4] = {root, root, root, root};
    bool first_level = true;
    for (int level = 1; level <= 4 * W * H + 1; level++) {
        BT nP[4] = {0, 0, 0, 0}, nV[4] = {0, 0, 0, 0};
#pragma unroll
        for (int f = 0; f < 4; f++) {
            const BT p = P[f];
#pragma unroll
            for (int a = 0; a < 4; a++) {
                BT c;
                if (a == 0) c = B::template shift<0>(p & cm[0]) | (p & ~cm[0]);
                else if (a == 1) c = B::template shift<1>(p & cm[1]) | (p & ~cm[1]);
                else if (a == 2) c = B::template shift<2>(p & cm[2]) | (p & ~cm[2]);
                else c = B::template shift<3>(p & cm[3]) | (p & ~cm[3]);
                BT nw = c & ~V[a];
                if (first_level && a != f) nw = 0;  // at depth 1 colour f is exactly action f
                nP[f] |= nw;
                nV[a] |= nw;
            }
        }
        // goal cells faced by a new state of this level
        const BT hit = B::template shift<0>(nV[0] & Tg[0]) | B::template shift<1>(nV[1] & Tg[1]) |
                       B::template shift<2>(nV[2] & Tg[2]) | B::template shift<3>(nV[3] & Tg[3]);
        if (hit) {
            goal_idx = B::lowest(hit);
            const BT g = B::bit(goal_idx);
            const BT tg[4] = {B::template unshift<0>(g), B::template unshift<1>(g),
                              B::template unshift<2>(g), B::template unshift<3>(g)};
#pragma unroll
            for (int f = 3; f >= 0; f--) {  // descending, so the smallest f is written last
                const BT p = P[f];
                BT faced = 0;
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    BT c;
                    if (a == 0) c = B::template shift<0>(p & cm[0]) | (p & ~cm[0]);
                    else if (a == 1) c = B::template shift<1>(p & cm[1]) | (p & ~cm[1]);
                    else if (a == 2) c = B::template shift<2>(p & cm[2]) | (p & ~cm[2]);
                    else c = B::template shift<3>(p & cm[3]) | (p & ~cm[3]);
                    BT nw = c & ~V[a];
                    if (first_level && a != f) nw = 0;
                    faced |= nw & tg[a];
                }
                if (faced) first = f;
            }
            return level;
        }
        if (!(nV[0] | nV[1] | nV[2] | nV[3])) return -1;  // queue drained (base.py:87)
#pragma unroll
        for (int a = 0; a < 4; a++) {
            V[a] |= nV[a];
            P[a] = nP[a];
        }
        first_level = false;
    }
    return -1;
}

// occupancy / goal bitboards from one env's grid row (any address space), CP = padded cells
template <int W, int H>
__device__ __forceinline__ void build_boards(const uint32_t *row_words, int goal_kind,
                                             typename Board<W, H>::BT &occ,
                                             typename Board<W, H>::BT &goal) {
    using BT = typename Board<W, H>::BT;
    constexpr int NW = (W * H + 3) / 4;
    const uint32_t gk = uint32_t(goal_kind) * 0x01010101u;
    occ = 0;
    goal = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        const uint32_t w = row_words[i];
        occ |= BT(mask_nibble(__vcmpne4(w, 0u))) << (4 * i);
        goal |= BT(mask_nibble(__vcmpeq4(w, gk))) << (4 * i);
    }
    occ &= Board<W, H>::all();
    goal &= Board<W, H>::all();
    if (goal_kind == 0) goal = 0;
}

// Loads one env's grid row into registers with 128-bit loads (row is 16-byte aligned).
template <int NW4>
__device__ __forceinline__ void load_row(const uint8_t *row, uint32_t *words) {
#pragma unroll
    for (int i = 0; i < NW4; i++) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(row) + i);
        words[4 * i + 0] = v.x;
        words[4 * i + 1] = v.y;
        words[4 * i + 2] = v.z;
        words[4 * i + 3] = v.w;
    }
}

// Teacher action for one env.  row_words: the env's grid row in registers.
template <int W, int H>
__device__ __forceinline__ int expert_env(const SharedTables &st, const Agent &a, int task,
                                          const uint32_t *row_words, int facing, int &dist,
                                          uint32_t &flags) {
    dist = -1;
    const uint32_t leaf = find_incomplete(st, task, a, facing);
    if (!leaf) return PSK_ACT_STOP;                         // demonstration.py:15-16
    const int kind = (leaf >> 16) & 0x7F;
    if (kind == LEAF_USE) return PSK_ACT_USE;               // demonstration.py:20-21
    if (kind != LEAF_GO) {                                  // demonstration.py:18 (assert)
        flags |= PSK_FLAG_BAD_LEAF;
        return PSK_ACT_INVALID;
    }
    typename Board<W, H>::BT occ, goal;
    build_boards<W, H>(row_words, (leaf >> 8) & 0xFF, occ, goal);
    int first, gidx;
    dist = bfs_first_action<W, H>(occ, goal, a.x(), a.y(), a.dir(), first, gidx);
    if (dist < 0) return PSK_ACT_STOP;                      // demonstration.py:25-26
    return first >= 0 ? first : PSK_ACT_INVALID;
}

template <int W, int H>
__global__ void __launch_bounds__(128)
craft_expert_kernel(const __grid_constant__ psk_craft_tables T, const uint8_t *__restrict__ grid,
                    const uint8_t *__restrict__ agent, const uint8_t *__restrict__ task,
                    uint8_t *__restrict__ action, int16_t *__restrict__ dist_out,
                    int32_t *err_flags, int64_t n, int cell_stride) {
    __shared__ SharedTables st;
    stage_tables(st, T);
    constexpr int NW4 = (W * H + 15) / 16;
    uint32_t flags = 0;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
         e += (int64_t)gridDim.x * blockDim.x) {
        const Agent a = load_agent_ro(agent, e);
        const uint8_t *row = grid + e * cell_stride;
        uint32_t words[NW4 * 4];
        load_row<NW4>(row, words);
        const int tk = task ? task[e] : a.task();
        const int facing = facing_kind<W, H>(a, row);
        int dist;
        const int act = expert_env<W, H>(st, a, tk, words, facing, dist, flags);
        action[e] = (uint8_t)act;
        if (dist_out) dist_out[e] = (int16_t)dist;
    }
    if (flags && err_flags) atomicOr(err_flags, (int)flags);
}

// find_closest_resources (teachers/base.py:27-34) for an explicit kind, with the full
// lexicographically-smallest action sequence reconstructed by re-running the BFS towards the
// chosen goal cell after every move (only the length is used on the training path).
template <int W, int H>
__global__ void __launch_bounds__(128)
craft_find_closest_kernel(const __grid_constant__ psk_craft_tables T,
                          const uint8_t *__restrict__ grid, const uint8_t *__restrict__ agent,
                          const uint8_t *__restrict__ kind, uint8_t *__restrict__ goal_out,
                          int16_t *__restrict__ len_out, uint8_t *__restrict__ seq, int seq_cap,
                          int64_t n, int cell_stride) {
    using B = Board<W, H>;
    using BT = typename B::BT;
    constexpr int NW4 = (W * H + 15) / 16;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
         e += (int64_t)gridDim.x * blockDim.x) {
        const Agent a = load_agent_ro(agent, e);
        uint32_t words[NW4 * 4];
        load_row<NW4>(grid + e * cell_stride, words);
        BT occ, goal;
        build_boards<W, H>(words, kind[e], occ, goal);
        int first, gidx;
        int x = a.x(), y = a.y(), d = a.dir();
        const int len = bfs_first_action<W, H>(occ, goal, x, y, d, first, gidx);
        len_out[e] = (int16_t)len;
        if (len < 0) {
            // best_goal keeps the LAST goal cell scanned when none is reachable (base.py:31)
            int last = -1;
            if (goal) {
                BT g = goal;
                while (g) { last = B::lowest(g); g &= g - 1; }
            }
            goal_out[2 * e] = last < 0 ? 255 : last / H;
            goal_out[2 * e + 1] = last < 0 ? 255 : last % H;
        } else {
            goal_out[2 * e] = gidx / H;
            goal_out[2 * e + 1] = gidx % H;
        }
        if (seq) {
            uint8_t *sq = seq + e * (int64_t)seq_cap;
            int k = 0;
            if (len > 0) {
                const BT g1 = B::bit(gidx);
                int f = first;
                for (; k < len && k < seq_cap; k++) {
                    sq[k] = (uint8_t)f;
                    const int tx = x + dx_of(f), ty = y + dy_of(f);
                    const bool can = tx >= 0 && ty >= 0 && tx < W && ty < H &&
                                     !((occ >> (tx * H + ty)) & 1);
                    if (can) { x = tx; y = ty; }
                    d = f;
                    if (k + 1 < len) {
                        int g2;
                        bfs_first_action<W, H>(occ, g1, x, y, d, f, g2);
                    }
                }
            }
            for (; k < seq_cap; k++) sq[k] = 255;
        }
    }
}

// =============================================================================================
// features  (worlds/craft.py:296-330)
// =============================================================================================
// Scatter of one env's non-zero features into its f32 row `frow` (shared memory), done by the
// TPE threads of the env; thread `j` owns cells [j*CPT, (j+1)*CPT) of the padded grid row.
template <int W, int H, int WIN, int TPE>
__device__ __forceinline__ void scatter_features(float *frow, const uint8_t *row, int cell_stride,
                                                 const Agent &a, int K, int j) {
    constexpr int HW = WIN / 2, BHW = (WIN * WIN) / 2;  // craft.py:299-302
    const int px = a.x(), py = a.y();
    const int cpt = cell_stride / TPE;                  // multiple of 8 (cell_stride % 64 == 0)
    const int c0 = j * cpt;
    for (int cb = 0; cb < cpt; cb += 8) {
        const uint2 v = *reinterpret_cast<const uint2 *>(row + c0 + cb);
        if ((v.x | v.y) == 0) continue;
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const int k = ((b < 4 ? v.x : v.y) >> ((b & 3) * 8)) & 0xFF;
            const int c = c0 + cb + b;
            if (k == 0 || c >= W * H) continue;
            const int dx = c / H - px, dy = c % H - py;
            // local window, ravel order (dx, dy, kind)  (craft.py:304-305)
            if (dx >= -HW && dx <= HW && dy >= -HW && dy <= HW)
                frow[((dx + HW) * WIN + (dy + HW)) * K + k] = 1.0f;
            // WIN^2 x WIN^2 window max-pooled in WIN x WIN blocks  (craft.py:306-310)
            const int bx = dx + BHW, by = dy + BHW;
            if (bx >= 0 && bx < WIN * WIN && by >= 0 && by < WIN * WIN)
                frow[WIN * WIN * K + ((bx / WIN) * WIN + (by / WIN)) * K + k] = 1.0f;
        }
    }
    // inventory counts (craft.py:325): thread j converts inventory word j (4 kinds)
    float *tail = frow + 2 * WIN * WIN * K;
    for (int wi = j; wi < 6; wi += TPE) {
        const uint32_t w = a.inv_word(wi);
        if (w == 0) continue;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int cnt = (w >> (8 * b)) & 0xFF;
            if (cnt && wi * 4 + b < K) tail[wi * 4 + b] = (float)cnt;
        }
    }
    if (j == TPE - 1) tail[K + a.dir()] = 1.0f;  // craft.py:321-322; the final element stays 0
}

// TMA helpers (cp.async.bulk, shared::cta -> global)
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// E envs per tile, TPE threads per env, CTA = E*TPE threads, two f32 tile buffers in dynamic smem.
template <int W, int H, int WIN, int E, int TPE, bool USE_TMA>
__global__ void __launch_bounds__(E *TPE)
craft_features_kernel(const uint8_t *__restrict__ grid, const uint8_t *__restrict__ agent,
                      float *__restrict__ out, int64_t n, int cell_stride, int K, int nf) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tile[2] = {reinterpret_cast<float *>(smem_raw),
                      reinterpret_cast<float *>(smem_raw) + (size_t)E * nf};
    const int tid = threadIdx.x;
    const int le = tid / TPE, j = tid % TPE;
    const int64_t n_tiles = (n + E - 1) / E;
    const int tile_f4 = E * nf / 4;  // E % 4 == 0, so the tile is a whole number of float4
    int it = 0;
    for (int64_t tile_id = blockIdx.x; tile_id < n_tiles; tile_id += gridDim.x, it++) {
        float *buf = tile[it & 1];
        const int64_t e = tile_id * E + le;
        const bool live = e < n;
        // inputs first, so the loads are in flight while the buffer is recycled
        Agent a;
        if (live) a = load_agent_ro(agent, e);
        if (USE_TMA) {
            // buffer `it&1` was handed to the TMA two iterations ago: wait until it was read
            if (tid == 0) bulk_wait_read<1>();
            __syncthreads();
        }
        float4 *b4 = reinterpret_cast<float4 *>(buf);
        for (int i = tid; i < tile_f4; i += E * TPE) b4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        if (live)
            scatter_features<W, H, WIN, TPE>(buf + (size_t)le * nf, grid + e * cell_stride,
                                             cell_stride, a, K, j);
        const int64_t e0 = tile_id * E;
        const int ne = (int)((n - e0) < E ? (n - e0) : E);
        const uint32_t bytes = (uint32_t)ne * (uint32_t)nf * 4u;
        float *gdst = out + e0 * nf;
        if (USE_TMA && (bytes & 15u) == 0) {
            fence_async_smem();  // generic-proxy writes -> visible to the async proxy
            __syncthreads();
            if (tid == 0) bulk_store(gdst, buf, bytes);
        } else {
            __syncthreads();
            if ((bytes & 15u) == 0) {
                float4 *g4 = reinterpret_cast<float4 *>(gdst);
                for (int i = tid; i < (int)(bytes / 16); i += E * TPE) __stcs(g4 + i, b4[i]);
            } else {
                for (int i = tid; i < ne * nf; i += E * TPE) gdst[i] = buf[i];
            }
            __syncthreads();
        }
    }
    if (USE_TMA) {
        if (tid == 0) bulk_wait_read<0>();  // smem must outlive the last bulk store's read
    }
}

// =============================================================================================
// reset / advance  (worlds/craft.py:258-273; trainers/imitation.py:63-73)
// =============================================================================================
__global__ void __launch_bounds__(256)
craft_reset_kernel(uint8_t *__restrict__ grid, uint8_t *__restrict__ agent,
                   const uint8_t *__restrict__ scen_grid, const int32_t *__restrict__ scen_idx,
                   const uint8_t *__restrict__ init_agent, const uint8_t *__restrict__ mask,
                   int64_t n, int cell_stride) {
    // one 16-byte chunk per thread: chunks [0, cs/16) are the grid row, the last two the agent
    const int chunks = cell_stride / 16 + 2;
    const int64_t total = n * chunks;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i / chunks;
        const int c = (int)(i % chunks);
        if (mask && !mask[e]) continue;
        if (c < cell_stride / 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(
                                      scen_grid + (int64_t)scen_idx[e] * cell_stride) + c);
            reinterpret_cast<uint4 *>(grid + e * cell_stride)[c] = v;
        } else {
            const int h = c - cell_stride / 16;
            reinterpret_cast<uint4 *>(agent + e * PSK_AGENT_BYTES)[h] =
                __ldg(reinterpret_cast<const uint4 *>(init_agent + e * PSK_AGENT_BYTES) + h);
        }
    }
}

// Per-env end of a rollout tick.  Returns true when the episode ended (state was reset).
template <int W, int H>
__device__ __forceinline__ bool advance_env(const SharedTables &st, Agent &a, uint8_t *row,
                                            int act, const uint8_t *scen_row,
                                            const uint8_t *init_agent_row, int cell_stride,
                                            bool &success, uint32_t &flags) {
    const int timer = a.timer() - 1;                          // imitation.py:63
    const bool done = (act == PSK_ACT_STOP) || timer <= 0;    // imitation.py:64-65
    if (done) {
        const int facing = facing_kind<W, H>(a, row);
        success = st.task_len(a.task())
                      ? node_satisfied(st.node(a.task(), 0), a, facing) == 1
                      : false;                                // imitation.py:69
        // auto-reset: CraftScenario.init (craft.py:270-273)
        for (int c = 0; c < cell_stride / 16; c++)
            reinterpret_cast<uint4 *>(row)[c] = __ldg(reinterpret_cast<const uint4 *>(scen_row) + c);
        const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(init_agent_row));
        const uint4 hi = __ldg(reinterpret_cast<const uint4 *>(init_agent_row) + 1);
        a.w[0] = lo.x; a.w[1] = lo.y; a.w[2] = lo.z; a.w[3] = lo.w;
        a.w[4] = hi.x; a.w[5] = hi.y; a.w[6] = hi.z; a.w[7] = hi.w;
    } else {
        success = false;
        step_env<W, H>(st, a, row, act, flags);               // imitation.py:72
        a.set_timer(timer);
    }
    return done;
}

__device__ __forceinline__ void add_stats(unsigned long long *stats, bool done, bool success,
                                          bool counted) {
    const unsigned full = __activemask();
    const int n_done = __popc(__ballot_sync(full, done));
    const int n_succ = __popc(__ballot_sync(full, success));
    const int n_step = __popc(__ballot_sync(full, counted));
    const int leader = __ffs(full) - 1;
    if (stats && (int)(threadIdx.x & 31) == leader) {
        if (n_done) atomicAdd(stats + 0, (unsigned long long)n_done);
        if (n_succ) atomicAdd(stats + 1, (unsigned long long)n_succ);
        if (n_step) atomicAdd(stats + 2, (unsigned long long)n_step);
    }
}

template <int W, int H>
__global__ void __launch_bounds__(256)
craft_advance_kernel(const __grid_constant__ psk_craft_tables T, uint8_t *__restrict__ grid,
                     uint8_t *__restrict__ agent, const uint8_t *__restrict__ action,
                     const uint8_t *__restrict__ scen_grid, const int32_t *__restrict__ scen_idx,
                     const uint8_t *__restrict__ init_agent, uint8_t *__restrict__ done_out,
                     uint8_t *__restrict__ success_out, unsigned long long *stats,
                     int32_t *err_flags, int64_t n, int cell_stride) {
    __shared__ SharedTables st;
    stage_tables(st, T);
    uint32_t flags = 0;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
         e += (int64_t)gridDim.x * blockDim.x) {
        Agent a = load_agent(agent, e);
        bool success;
        const bool done = advance_env<W, H>(st, a, grid + e * cell_stride, action[e],
                                            scen_grid + (int64_t)scen_idx[e] * cell_stride,
                                            init_agent + e * PSK_AGENT_BYTES, cell_stride,
                                            success, flags);
        store_agent(agent, e, a);
        if (done_out) done_out[e] = done;
        if (success_out) success_out[e] = success;
        add_stats(stats, done, success, true);
    }
    if (flags && err_flags) atomicOr(err_flags, (int)flags);
}

// =============================================================================================
// fused tick: expert + features + advance, state read once
// =============================================================================================
// CTA = 128 threads = one super-tile of 128 envs.
//   phase A (thread per env): agent -> registers, grid row -> registers + shared copy,
//                             teacher action (bitboard BFS).
//   phase B (all threads):    features of the 128 envs in 128/E sub-tiles, each built in a
//                             double-buffered f32 smem tile and TMA-bulk-stored.
//   phase C (thread per env): advance (step / done / success / auto-reset) on the shared row,
//                             write-back of what changed.
template <int W, int H, int WIN, int E>
__global__ void __launch_bounds__(128)
craft_tick_kernel(const __grid_constant__ psk_craft_tables T, uint8_t *__restrict__ grid,
                  uint8_t *__restrict__ agent, const uint8_t *__restrict__ action_in,
                  const uint8_t *__restrict__ scen_grid, const int32_t *__restrict__ scen_idx,
                  const uint8_t *__restrict__ init_agent, float *__restrict__ features_out,
                  uint8_t *__restrict__ expert_out, uint8_t *__restrict__ done_out,
                  uint8_t *__restrict__ success_out, unsigned long long *stats,
                  int32_t *err_flags, int64_t n, int cell_stride, int K, int nf) {
    constexpr int NT = 128, TPE = NT / E;
    constexpr int CP = ((W * H + 63) / 64) * 64;
    constexpr int NW4 = (W * H + 15) / 16;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ SharedTables st;
    __shared__ __align__(16) uint8_t s_rows[NT * CP];
    __shared__ __align__(16) uint32_t s_agent[NT * 8];
    float *tile[2] = {reinterpret_cast<float *>(smem_raw),
                      reinterpret_cast<float *>(smem_raw) + (size_t)E * nf};
    stage_tables(st, T);
    const int tid = threadIdx.x;
    uint32_t flags = 0;
    int it = 0;  // running count of feature sub-tiles handed to the TMA (buffer parity)
    const int64_t n_super = (n + NT - 1) / NT;
    for (int64_t sp = blockIdx.x; sp < n_super; sp += gridDim.x) {
        const int64_t e = sp * NT + tid;
        const bool live = e < n;
        // ---- phase A
        Agent a;
        int act = PSK_ACT_STOP;
        if (live) {
            a = load_agent(agent, e);
            uint32_t words[NW4 * 4];
            const uint8_t *row = grid + e * cell_stride;
            load_row<NW4>(row, words);
#pragma unroll
            for (int i = 0; i < NW4; i++)
                reinterpret_cast<uint4 *>(s_rows + tid * CP)[i] =
                    make_uint4(words[4 * i], words[4 * i + 1], words[4 * i + 2], words[4 * i + 3]);
#pragma unroll
            for (int i = 0; i < 8; i++) s_agent[tid * 8 + i] = a.w[i];
            const int facing = facing_kind<W, H>(a, s_rows + tid * CP);
            int dist;
            act = expert_env<W, H>(st, a, a.task(), words, facing, dist, flags);
            expert_out[e] = (uint8_t)act;
            if (action_in) act = action_in[e];
        }
        __syncthreads();
        // ---- phase B
        if (features_out) {
            const int tile_f4 = E * nf / 4;
            for (int sub = 0; sub < NT / E; sub++, it++) {
                const int64_t e0 = sp * NT + sub * E;
                if (e0 >= n) break;
                float *buf = tile[it & 1];
                if (tid == 0) bulk_wait_read<1>();
                __syncthreads();
                float4 *b4 = reinterpret_cast<float4 *>(buf);
                for (int i = tid; i < tile_f4; i += NT) b4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                __syncthreads();
                const int le = tid / TPE, j = tid % TPE;
                const int se = sub * E + le;  // env slot inside the super-tile
                if (e0 + le < n) {
                    Agent b;
#pragma unroll
                    for (int i = 0; i < 8; i++) b.w[i] = s_agent[se * 8 + i];
                    scatter_features<W, H, WIN, TPE>(buf + (size_t)le * nf, s_rows + se * CP, CP,
                                                     b, K, j);
                }
                const int ne = (int)((n - e0) < E ? (n - e0) : E);
                const uint32_t bytes = (uint32_t)ne * (uint32_t)nf * 4u;
                float *gdst = features_out + e0 * nf;
                if ((bytes & 15u) == 0) {
                    fence_async_smem();
                    __syncthreads();
                    if (tid == 0) bulk_store(gdst, buf, bytes);
                } else {
                    __syncthreads();
                    for (int i = tid; i < ne * nf; i += NT) gdst[i] = buf[i];
                    __syncthreads();
                }
            }
        }
        // ---- phase C
        bool done = false, success = false;
        if (live) {
            const Agent before = a;
            uint8_t *srow = s_rows + tid * CP;
            done = advance_env<W, H>(st, a, srow, act,
                                     scen_grid + (int64_t)scen_idx[e] * cell_stride,
                                     init_agent + e * PSK_AGENT_BYTES, CP, success, flags);
            bool changed = false;
#pragma unroll
            for (int i = 0; i < 8; i++) changed |= a.w[i] != before.w[i];
            if (changed) store_agent(agent, e, a);
            // the grid row changes on reset, pick-up, bridge and axe; write 16-byte chunks that differ
            uint8_t *row = grid + e * cell_stride;
            if (done || act == PSK_ACT_USE) {
#pragma unroll
                for (int i = 0; i < NW4; i++) {
                    const uint4 nv = reinterpret_cast<const uint4 *>(srow)[i];
                    reinterpret_cast<uint4 *>(row)[i] = nv;
                }
            }
            if (done_out) done_out[e] = done;
            if (success_out) success_out[e] = success;
        }
        add_stats(stats, done, success, live);
        __syncthreads();  // s_rows / s_agent are recycled by the next super-tile
    }
    if (tid == 0) bulk_wait_read<0>();
    if (flags && err_flags) atomicOr(err_flags, (int)flags);
}

// =============================================================================================
// host side: dispatch on (width, height, window)
// =============================================================================================
static int g_num_sms = 0;
static int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

static inline int grid_for(int64_t n, int block, int ctas_per_sm) {
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)num_sms() * ctas_per_sm;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

static inline int check(cudaError_t e) { return e == cudaSuccess ? PSK_OK : PSK_ERR_CUDA; }

template <int W, int H, int WIN> struct Config {
    static constexpr int CP = ((W * H + 63) / 64) * 64;
    // envs per feature tile: two f32 tiles of E rows must fit next to 3 more CTAs on the SM
    static constexpr int E = (WIN == 3) ? 16 : 8;
    static constexpr int TPE = 8;

    static bool matches(const psk_craft_tables *t) {
        return t->width == W && t->height == H && t->window_w == WIN && t->window_h == WIN &&
               t->n_kinds <= PSK_MAX_INV;
    }
    static int nf(const psk_craft_tables *t) { return 2 * WIN * WIN * t->n_kinds + t->n_kinds + 5; }

    static int step(const psk_craft_tables *t, psk_craft_state s, const uint8_t *action,
                    const uint8_t *active, float *reward, int32_t *err, cudaStream_t st) {
        craft_step_kernel<W, H><<<grid_for(s.n, 256, 8), 256, 0, st>>>(
            *t, s.grid, s.agent, action, active, reward, err, s.n, s.cell_stride);
        return check(cudaGetLastError());
    }
    static int satisfies(const psk_craft_tables *t, psk_craft_state s, const uint8_t *task,
                         uint8_t *out, cudaStream_t st) {
        craft_satisfies_kernel<W, H><<<grid_for(s.n, 256, 8), 256, 0, st>>>(
            *t, s.grid, s.agent, task, out, s.n, s.cell_stride);
        return check(cudaGetLastError());
    }
    static int expert(const psk_craft_tables *t, psk_craft_state s, const uint8_t *task,
                      uint8_t *action, int16_t *dist, int32_t *err, cudaStream_t st) {
        craft_expert_kernel<W, H><<<grid_for(s.n, 128, 8), 128, 0, st>>>(
            *t, s.grid, s.agent, task, action, dist, err, s.n, s.cell_stride);
        return check(cudaGetLastError());
    }
    static int find_closest(const psk_craft_tables *t, psk_craft_state s, const uint8_t *kind,
                            uint8_t *goal, int16_t *len, uint8_t *seq, int seq_cap,
                            cudaStream_t st) {
        craft_find_closest_kernel<W, H><<<grid_for(s.n, 128, 8), 128, 0, st>>>(
            *t, s.grid, s.agent, kind, goal, len, seq, seq_cap, s.n, s.cell_stride);
        return check(cudaGetLastError());
    }
    template <bool TMA>
    static int features_impl(const psk_craft_tables *t, psk_craft_state s, float *out,
                             cudaStream_t st) {
        const int f = nf(t);
        const size_t smem = (size_t)2 * E * f * sizeof(float);
        auto kern = craft_features_kernel<W, H, WIN, E, TPE, TMA>;
        static size_t configured = 0;  // opt in to > 48 KB dynamic smem once per size
        if (configured != smem) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem) != cudaSuccess)
                return PSK_ERR_CUDA;
            configured = smem;
        }
        const int per_sm = (int)((220 * 1024) / (smem + 1024));
        const int64_t tiles = (s.n + E - 1) / E;
        int64_t g = (int64_t)num_sms() * (per_sm > 0 ? per_sm : 1);
        if (g > tiles) g = tiles > 0 ? tiles : 1;
        kern<<<(int)g, E * TPE, smem, st>>>(s.grid, s.agent, out, s.n, s.cell_stride, t->n_kinds, f);
        return check(cudaGetLastError());
    }
    static int features(const psk_craft_tables *t, psk_craft_state s, float *out, int impl,
                        cudaStream_t st) {
        if ((reinterpret_cast<uintptr_t>(out) & 15) != 0) impl = 1;
        return impl == 1 ? features_impl<false>(t, s, out, st) : features_impl<true>(t, s, out, st);
    }
    static int advance(const psk_craft_tables *t, psk_craft_state s, psk_craft_episodes ep,
                       const uint8_t *action, uint8_t *done, uint8_t *success,
                       unsigned long long *stats, int32_t *err, cudaStream_t st) {
        craft_advance_kernel<W, H><<<grid_for(s.n, 256, 8), 256, 0, st>>>(
            *t, s.grid, s.agent, action, ep.scen_grid, ep.scen_idx, ep.init_agent, done, success,
            stats, err, s.n, s.cell_stride);
        return check(cudaGetLastError());
    }
    static int tick_fused(const psk_craft_tables *t, psk_craft_state s, psk_craft_episodes ep,
                          const uint8_t *action_in, float *features_out, uint8_t *expert_out,
                          uint8_t *done, uint8_t *success, unsigned long long *stats,
                          int32_t *err, cudaStream_t st) {
        const int f = nf(t);
        const size_t smem = (size_t)2 * E * f * sizeof(float);
        auto kern = craft_tick_kernel<W, H, WIN, E>;
        static size_t configured = 0;
        if (configured != smem) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem) != cudaSuccess)
                return PSK_ERR_CUDA;
            configured = smem;
        }
        const size_t static_smem = sizeof(SharedTables) + (size_t)128 * CP + 128 * 32;
        const int per_sm = (int)((220 * 1024) / (smem + static_smem + 1024));
        const int64_t tiles = (s.n + 127) / 128;
        int64_t g = (int64_t)num_sms() * (per_sm > 0 ? per_sm : 1);
        if (g > tiles) g = tiles > 0 ? tiles : 1;
        kern<<<(int)g, 128, smem, st>>>(*t, s.grid, s.agent, action_in, ep.scen_grid, ep.scen_idx,
                                        ep.init_agent, features_out, expert_out, done, success,
                                        stats, err, s.n, s.cell_stride, t->n_kinds, f);
        return check(cudaGetLastError());
    }
};

using Medium = Config<8, 8, 3>;    // configs/worlds/craft_medium.yaml
using Large = Config<10, 10, 5>;   // configs/worlds/craft_large.yaml

#define PSK_DISPATCH(t, CALL)                         \
    do {                                              \
        if (Medium::matches(t)) return Medium::CALL;  \
        if (Large::matches(t)) return Large::CALL;    \
        return PSK_ERR_UNSUPPORTED;                   \
    } while (0)

static bool state_ok(const psk_craft_tables *t, const psk_craft_state &s) {
    if (!t || s.n < 0) return false;
    if (s.n == 0) return true;
    if (!s.grid || !s.agent) return false;
    const int need = ((t->width * t->height + 63) / 64) * 64;
    if (s.cell_stride != need) return false;
    if ((reinterpret_cast<uintptr_t>(s.grid) & 15) || (reinterpret_cast<uintptr_t>(s.agent) & 31))
        return false;
    return true;
}

}  // namespace psk

using namespace psk;

extern "C" {

const char *psk_version(void) { return "psketch_b200 0.1 sm_100a"; }

int psk_craft_supported(const psk_craft_tables *t) {
    if (!t) return 0;
    return Medium::matches(t) || Large::matches(t);
}

int psk_craft_n_features(const psk_craft_tables *t) {
    if (!psk_craft_supported(t)) return -1;
    return 2 * t->window_w * t->window_h * t->n_kinds + t->n_kinds + 5;
}

int psk_craft_step(const psk_craft_tables *t, psk_craft_state s, const uint8_t *action,
                   const uint8_t *active, float *reward, int32_t *err_flags, void *stream) {
    if (!state_ok(t, s) || (!action && s.n)) return PSK_ERR_BADARG;
    if (s.n == 0) return PSK_OK;
    PSK_DISPATCH(t, step(t, s, action, active, reward, err_flags, (cudaStream_t)stream));
}

int psk_craft_features(const psk_craft_tables *t, psk_craft_state s, float *out, int impl,
                       void *stream) {
    if (!state_ok(t, s) || (!out && s.n)) return PSK_ERR_BADARG;
    if (s.n == 0) return PSK_OK;
    PSK_DISPATCH(t, features(t, s, out, impl, (cudaStream_t)stream));
}

int psk_craft_satisfies(const psk_craft_tables *t, psk_craft_state s, const uint8_t *task,
                        uint8_t *out, void *stream) {
    if (!state_ok(t, s) || (!out && s.n)) return PSK_ERR_BADARG;
    if (s.n == 0) return PSK_OK;
    PSK_DISPATCH(t, satisfies(t, s, task, out, (cudaStream_t)stream));
}

int psk_craft_expert(const psk_craft_tables *t, psk_craft_state s, const uint8_t *task,
                     uint8_t *action, int16_t *dist, int32_t *err_flags, void *stream) {
    if (!state_ok(t, s) || (!action && s.n)) return PSK_ERR_BADARG;
    if (s.n == 0) return PSK_OK;
    PSK_DISPATCH(t, expert(t, s, task, action, dist, err_flags, (cudaStream_t)stream));
}

int psk_craft_find_closest(const psk_craft_tables *t, psk_craft_state s, const uint8_t *kind,
                           uint8_t *goal, int16_t *length, uint8_t *seq, int32_t seq_cap,
                           void *stream) {
    if (!state_ok(t, s) || ((!kind || !goal || !length) && s.n) || seq_cap < 0)
        return PSK_ERR_BADARG;
    if (s.n == 0) return PSK_OK;
    PSK_DISPATCH(t, find_closest(t, s, kind, goal, length, seq, seq_cap, (cudaStream_t)stream));
}

int psk_craft_reset(psk_craft_state s, psk_craft_episodes ep, const uint8_t *mask, void *stream) {
    if (s.n < 0 || s.cell_stride <= 0 || s.cell_stride % 64) return PSK_ERR_BADARG;
    if (s.n == 0) return PSK_OK;
    if (!s.grid || !s.agent || !ep.scen_grid || !ep.scen_idx || !ep.init_agent)
        return PSK_ERR_BADARG;
    const int64_t total = s.n * (s.cell_stride / 16 + 2);
    craft_reset_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        s.grid, s.agent, ep.scen_grid, ep.scen_idx, ep.init_agent, mask, s.n, s.cell_stride);
    return check(cudaGetLastError());
}

int psk_craft_tick(const psk_craft_tables *t, psk_craft_state s, psk_craft_episodes ep,
                   const uint8_t *action_in, float *features_out, uint8_t *expert_out,
                   uint8_t *done_out, uint8_t *success_out, unsigned long long *stats,
                   int32_t *err_flags, int fused, void *stream) {
    if (!state_ok(t, s)) return PSK_ERR_BADARG;
    if (s.n == 0) return PSK_OK;
    if (!expert_out || !ep.scen_grid || !ep.scen_idx || !ep.init_agent) return PSK_ERR_BADARG;
    if (features_out && (reinterpret_cast<uintptr_t>(features_out) & 15)) return PSK_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (fused) {
        PSK_DISPATCH(t, tick_fused(t, s, ep, action_in, features_out, expert_out, done_out,
                                   success_out, stats, err_flags, st));
    }
    int rc = psk_craft_expert(t, s, nullptr, expert_out, nullptr, err_flags, stream);
    if (rc) return rc;
    if (features_out) {
        rc = psk_craft_features(t, s, features_out, 0, stream);
        if (rc) return rc;
    }
    const uint8_t *act = action_in ? action_in : expert_out;
    PSK_DISPATCH(t, advance(t, s, ep, act, done_out, success_out, stats, err_flags, st));
}

}  // extern "C"
