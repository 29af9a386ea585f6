"""Pure-Python/numpy restatement of the reference's per-env hot path — TEST INFRASTRUCTURE and
the CPU BASELINE ("the reference's Python CPU loop" of BASELINE.json), never the product.

It keeps the reference's data representation (one-hot float64 grid [W,H,K], float64 inventory,
persistent states, python tuples for positions) and performs the same numpy work per call, so
that timing it on the GPU box's host cores stands in for the reference itself, which cannot
travel to that box.  Pinned in tests/test_ref_port.py against the golden trajectories and the
reference-exported states; in the build container it is also timed against the real reference
(DESIGN.md records the ratio).

Restated from: worlds/craft.py:275-455 (state, satisfies, features, step, neighbours,
navigation grid, resource positions), misc/array.py:3-25 (zero-padded window),
skimage.measure.block_reduce (block max), teachers/base.py:10-87 and
teachers/demonstration.py:9-30 (hint-tree walk, per-goal FIFO BFS, closest resource).
"""
import numpy as np

MOVES = ((0, -1), (0, 1), (-1, 0), (1, 0))      # DOWN, UP, LEFT, RIGHT
USE, STOP = 4, 5


def window(grid, x0, x1, y0, y1):
    """Zero-padded slice grid[x0:x1, y0:y1, :] (misc/array.py:3-25)."""
    cx0, cy0 = max(x0, 0), max(y0, 0)
    part = grid[cx0:x1, cy0:y1, :]
    out = np.zeros((x1 - x0, y1 - y0) + grid.shape[2:])
    ox, oy = cx0 - x0, cy0 - y0
    out[ox:ox + part.shape[0], oy:oy + part.shape[1], :] = part
    return out


def block_max(a, bw, bh):
    """block_reduce(a, (bw, bh, 1), np.max) for shapes that divide evenly."""
    nx, ny, k = a.shape[0] // bw, a.shape[1] // bh, a.shape[2]
    return a.reshape(nx, bw, ny, bh, k, 1).max(axis=(1, 3, 5))


class PortWorld(object):
    def __init__(self, tables):
        self.tables = tables
        self.cookbook = tables.cookbook
        self.W, self.H = tables.W, tables.H
        self.win_w, self.win_h = tables.win_w, tables.win_h
        self.n_kinds = tables.K
        self.n_features = tables.n_features
        env = self.cookbook.environment
        self.grabbable = [i for i in range(self.n_kinds) if i not in env]
        self.workshops = list(tables.workshop_kinds)
        self.water, self.stone = tables.water_kind, tables.stone_kind

    def init_state(self, grid, pos, dir=0):
        return PortState(self, grid, tuple(pos), dir, np.zeros(self.n_kinds))

    def onehot(self, ids):
        ids = np.asarray(ids).reshape(self.W, self.H)
        g = np.zeros((self.W, self.H, self.n_kinds))
        xs, ys = np.nonzero(ids)
        g[xs, ys, ids[xs, ys]] = 1
        return g


class PortState(object):
    def __init__(self, world, grid, pos, dir, inventory):
        self.world, self.grid, self.pos, self.dir, self.inventory = world, grid, pos, dir, inventory
        self._features = None

    def satisfies(self, task):
        thing = self.world.cookbook.index[task.goal_arg]
        if task.goal_name in ("make", "get"):
            return self.inventory[thing] > 0
        if task.goal_name == "go":
            d = MOVES[self.dir]
            return self.grid[self.pos[0] + d[0], self.pos[1] + d[1], thing] > 0
        return None

    def features(self):
        if self._features is None:
            w = self.world
            x, y = self.pos
            hw, hh = w.win_w // 2, w.win_h // 2
            bw, bh = (w.win_w ** 2) // 2, (w.win_h ** 2) // 2
            near = window(self.grid, x - hw, x + hw + 1, y - hh, y + hh + 1)
            far = block_max(window(self.grid, x - bw, x + bw + 1, y - bh, y + bh + 1),
                            w.win_w, w.win_h)
            facing = np.zeros(4)
            facing[self.dir] = 1
            self._features = np.concatenate((near.ravel(), far.ravel(), self.inventory, facing, [0]))
            assert len(self._features) == w.n_features
        return self._features

    def front(self):
        x, y = self.pos
        d = self.dir
        if d == 2 and x > 0:
            return [(x - 1, y)]
        if d == 0 and y > 0:
            return [(x, y - 1)]
        if d == 3 and x < self.world.W - 1:
            return [(x + 1, y)]
        if d == 1 and y < self.world.H - 1:
            return [(x, y + 1)]
        return []

    def step(self, action):
        w = self.world
        x, y = self.pos
        ndir, ninv, ngrid = self.dir, self.inventory, self.grid
        dx = dy = 0
        if 0 <= action < 4:
            dx, dy = MOVES[action]
            ndir = action
        elif action == STOP:
            pass
        elif action == USE:
            cb = w.cookbook
            for nx, ny in self.front():
                here = self.grid[nx, ny, :]
                if not here.any():
                    continue
                assert here.sum() == 1
                thing = int(here.argmax())
                if not (thing in w.grabbable or thing in w.workshops or thing == w.water
                        or thing == w.stone):
                    continue
                ninv = self.inventory.copy()
                ngrid = self.grid.copy()
                if thing in w.grabbable:
                    ninv[thing] += 1
                    ngrid[nx, ny, thing] = 0
                elif thing in w.workshops:
                    shop = cb.index.get(thing)
                    for out, recipe in list(cb.recipes.items()):
                        if recipe["_at"] != shop:
                            continue
                        made = recipe.get("_yield", 1)
                        needs = [i for i in recipe if isinstance(i, int)]
                        if any(ninv[i] < recipe[i] for i in needs):
                            continue
                        ninv[out] += made
                        for i in needs:
                            ninv[i] -= recipe[i]
                elif thing == w.water:
                    if ninv[cb.index["bridge"]] > 0:
                        ngrid[nx, ny, w.water] = 0
                        ninv[cb.index["bridge"]] -= 1
                elif thing == w.stone:
                    if ninv[cb.index["axe"]] > 0:
                        ngrid[nx, ny, w.stone] = 0
                break
        else:
            raise Exception("Unexpected action: %s" % action)
        tx, ty = x + dx, y + dy
        if self.grid[tx, ty, :].any():
            tx, ty = x, y
        return 0, PortState(w, ngrid, (tx, ty), ndir, ninv)

    def nav(self):
        return self.grid.max(axis=2)

    def positions_of(self, goal_arg):
        thing = self.world.cookbook.index[goal_arg]
        return list(zip(*self.grid[:, :, thing].nonzero()))


class PortTeacher(object):
    def incomplete(self, task, state):
        if state.satisfies(task):
            return None
        if task.subtasks is None:
            return task
        for sub in task.subtasks[:-1]:
            found = self.incomplete(sub, state)
            if found is not None:
                return found
        found = self.incomplete(task.subtasks[-1], state)
        assert found is not None
        return found

    def path_to(self, state, goal):
        nav = state.nav()
        came = {}
        queue = [None] * 1000
        head, tail = 0, 0
        first = (state.pos, state.dir)
        queue[tail] = first
        tail += 1
        came[first] = -1
        while head < tail:
            item = queue[head]
            head += 1
            pos, d = item
            m = MOVES[d]
            if (pos[0] + m[0], pos[1] + m[1]) == goal:
                seq = []
                while came[item] != -1:
                    act, item = came[item]
                    seq.append(act)
                seq.reverse()
                return seq
            for a, m in enumerate(MOVES):
                npos = (pos[0] + m[0], pos[1] + m[1])
                if nav[npos[0], npos[1]]:
                    npos = pos
                nxt = (npos, a)
                if nxt not in came:
                    queue[tail] = nxt
                    tail += 1
                    came[nxt] = (a, item)
        return None

    def closest(self, task, state):
        best = (None, None)
        for goal in state.positions_of(task.goal_arg):
            seq = self.path_to(state, goal)
            if best[1] is None or len(seq) < len(best[1]):
                best = (goal, seq)
        return best

    def __call__(self, task, state):
        sub = self.incomplete(task, state)
        if sub is None:
            return STOP
        assert sub.goal_name in ("use", "go")
        if sub.goal_name == "use":
            return USE
        _, seq = self.closest(sub, state)
        if seq is None:
            return STOP
        return seq[0]


def run_instances(tables, grids_ids, inst_env, inst_pos, inst_task, lo, hi, want_actions=False):
    """The reference loop of BASELINE config 1 (make_data.py:146-152 + students/imitation.py:72)
    over instances [lo, hi): a = teacher(task, s); f = s.features(); stop or s = s.step(a).
    Returns (env_steps, feature checksum[, action lists])."""
    world = PortWorld(tables)
    teacher = PortTeacher()
    tm = tables.task_manager
    onehots = {}
    steps, checksum, all_actions = 0, 0.0, []
    for i in range(lo, hi):
        e = int(inst_env[i])
        if e not in onehots:
            onehots[e] = world.onehot(grids_ids[e])
        task = tm.by_id(int(inst_task[i]))
        s = world.init_state(onehots[e], (int(inst_pos[i][0]), int(inst_pos[i][1])))
        acts = []
        while True:
            a = teacher(task, s)
            f = s.features()
            checksum += float(f.sum())
            steps += 1
            acts.append(a)
            if a == STOP:
                break
            _, s = s.step(a)
        if want_actions:
            all_actions.append(acts)
    return (steps, checksum, all_actions) if want_actions else (steps, checksum)


def _worker(args):
    import time
    tables_cfg, grids_ids, inst_env, inst_pos, inst_task, lo, hi = args
    from psketch_b200.tables import CraftTables
    tables = CraftTables(world_config=tables_cfg)
    t0 = time.perf_counter()
    steps, checksum = run_instances(tables, grids_ids, inst_env, inst_pos, inst_task, lo, hi)
    return steps, checksum, time.perf_counter() - t0


class ParallelRunner(object):
    """One Python process per core (fork pool created once).  ``run`` stripes ``n_instances``
    instances over the processes in contiguous blocks and returns (total env-steps, seconds),
    where seconds is the slowest worker's own compute time (pool start-up is not charged)."""

    def __init__(self, world_config, grids_ids, procs):
        import multiprocessing as mp
        self.world_config, self.grids_ids, self.procs = world_config, grids_ids, procs
        self.pool = mp.get_context("fork").Pool(procs)

    def run(self, inst_env, inst_pos, inst_task, n_instances):
        bounds = np.linspace(0, n_instances, self.procs + 1).astype(int)
        jobs = [(self.world_config, self.grids_ids, inst_env, inst_pos, inst_task,
                 int(bounds[p]), int(bounds[p + 1]))
                for p in range(self.procs) if bounds[p + 1] > bounds[p]]
        res = self.pool.map(_worker, jobs, chunksize=1)
        return sum(r[0] for r in res), max(r[2] for r in res)

    def close(self):
        self.pool.close()
        self.pool.join()
