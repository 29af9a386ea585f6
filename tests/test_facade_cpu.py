"""CPU-tier test of the object API's host logic (psketch_b200/worlds/craft.py): recording of
transitions, flushing by generations, all-task evaluation of states whose task is unknown, hint
inheritance, caches and invalidation — with the device backend replaced by a TEST DOUBLE that
computes the same quantities with the CPU oracle.  (The product backend needs a GPU and is covered
by tests/test_facade_gpu.py; nothing outside tests/ routes through the oracle.)  The expected
outputs are those of the reference's own ImitationTrainer.do_rollout (tests/golden/trainer_rollouts.npz)."""
import numpy as np
import pytest

from psketch_b200 import _lib
from psketch_b200.worlds import craft as facade


class OracleBackend(facade._Backend):
    """Same contract as _Backend._run, computed by oracle/craft_oracle.c; counts its launches."""

    def __init__(self, world, oracle):
        self.torch = None
        self.world = world
        self.o = oracle
        self.C = world.tables.W * world.tables.H
        self.K = world.tables.K
        self.T = world.tables.n_tasks - 1
        self.runs = []

    def find_closest(self, cells, agent, kind, seq_cap=96):
        agent = np.asarray(agent, np.uint8)
        goal, length, status, seq = self.o.find_closest(
            np.asarray(cells, np.uint8)[None], agent[None, [_lib.AG_X, _lib.AG_Y]].astype(np.int32),
            agent[None, _lib.AG_DIR].astype(np.int32), np.asarray([kind]), seq_cap=seq_cap)
        g = goal[0].astype(np.int64)
        g = g.astype(np.uint8) if (g >= 0).all() else np.array([255, 255], np.uint8)
        return g, int(length[0]), seq[0]

    def _run(self, states, step):
        n = len(states)
        self.runs.append((n, step))
        src = [s._parent for s in states] if step else states
        grid = np.stack([np.asarray(q._cells, np.uint8) for q in src])
        agent = np.stack([np.asarray(q._agent, np.uint8) for q in src])
        inv = agent[:, :self.K].astype(np.int32)
        pos = agent[:, [_lib.AG_X, _lib.AG_Y]].astype(np.int32)
        dirs = agent[:, _lib.AG_DIR].astype(np.int32)
        if step:
            acts = np.asarray([s._action for s in states], np.int32)
            grid2, inv, pos, dirs, status = self.o.step(grid, inv, pos, dirs, acts)
            assert (status == 0).all()
            agent = agent.copy()
            agent[:, :self.K] = inv
            agent[:, _lib.AG_X], agent[:, _lib.AG_Y], agent[:, _lib.AG_DIR] = pos[:, 0], pos[:, 1], dirs
        else:
            grid2 = grid
        feats = self.o.features(grid2, inv, pos, dirs).astype(np.float64)
        hints = np.asarray([s._task_hint for s in states], np.int32)
        expert = self.o.expert(grid2, inv, pos, dirs, hints)[0]
        sat = self.o.satisfies(grid2, inv, pos, dirs, hints)
        hintless = [i for i, s in enumerate(states) if not s._task_hint]
        if hintless:
            sub = np.asarray(hintless)
            all_e = np.zeros((self.T, len(sub)), np.uint8)
            all_s = np.zeros((self.T, len(sub)), np.uint8)
            for t in range(1, self.T + 1):
                tk = np.full(len(sub), t, np.int32)
                all_e[t - 1] = self.o.expert(grid2[sub], inv[sub], pos[sub], dirs[sub], tk)[0]
                all_s[t - 1] = self.o.satisfies(grid2[sub], inv[sub], pos[sub], dirs[sub], tk)
            for col, i in enumerate(hintless):
                states[i]._all = (all_e, all_s, col)
        for i, s in enumerate(states):
            if step:
                parent = s._parent
                if np.array_equal(grid2[i], parent._cells):
                    s._cells, s._grid = parent._cells, parent._grid
                else:
                    s._cells = grid2[i]
                s._agent = agent[i]
                s._parent = None
            s._cached_features = feats[i]
            if s._task_hint:
                s._expert[s._task_hint] = int(expert[i])
                s._sat[s._task_hint] = int(sat[i])
            s._evaluated = True


class _Cfg(object):
    pass


def _world(medium_tables, medium_oracle):
    cfg = _Cfg()
    cfg.recipes = None
    cfg.world = _Cfg(); cfg.world.name = "CraftWorld"; cfg.world.config = "craft_medium"
    cfg.student = _Cfg(); cfg.student.model = _Cfg()
    cfg.trainer = _Cfg(); cfg.trainer.hints = None
    cfg.random = np.random.RandomState(1)
    world = facade.CraftWorld(cfg, tables=medium_tables)
    world._backend = OracleBackend(world, medium_oracle)
    return world


def _Teacher(oracle):
    """The product teacher: its closest-resource query goes through backend.find_closest, which the
    test double above answers with the oracle."""
    from psketch_b200.teachers import DemonstrationTeacher
    return DemonstrationTeacher(None)


def _batch(world, splits, inst):
    K = world.cookbook.n_kinds
    out = []
    for i in inst:
        ids = splits["dev_grids"][splits["dev_inst_env"][i]].reshape(8, 8)
        onehot = np.zeros((8, 8, K))
        xs, ys = np.nonzero(ids)
        onehot[xs, ys, ids[xs, ys]] = 1
        out.append(dict(grid=onehot, init_pos=tuple(int(v) for v in splits["dev_inst_pos"][i]),
                        task=world.task_manager.by_id(int(splits["dev_inst_task"][i]))))
    return out


def test_trainer_protocol_on_the_facade_logic(splits, medium_tables, medium_oracle, trainer_rollouts):
    from trainer_loop import check_against_fixture
    world = _world(medium_tables, medium_oracle)
    check_against_fixture(trainer_rollouts, lambda inst: _batch(world, splits, inst), world,
                          _Teacher(medium_oracle))
    # one evaluation per timestep: 40 timesteps in each mode, plus the distance probes at the end
    # (the first version of the facade needed two launches per ENV in the first timestep)
    runs = world._backend.runs
    assert len(runs) <= 2 * 41 + 2 * 8, len(runs)
    # (distance probes that were created but never read are evaluated with the next batch)
    assert 32 <= max(n for n, _ in runs) <= 40


def test_states_are_persistent_and_caches_are_per_state(splits, medium_tables, medium_oracle):
    world = _world(medium_tables, medium_oracle)
    item = _batch(world, splits, [0])[0]
    s0 = world.init_state(item["grid"], item["init_pos"])
    f0 = s0.features().copy()
    r, s1 = s0.step(3)
    r2, s1b = s0.step(1)                      # branching from the same state: s0 is not mutated
    assert r == 0 and r2 == 0
    assert s1.dir == 3 and s1b.dir == 1 and s0.dir == 0
    assert np.array_equal(s0.features(), f0)
    assert s1.features() is s1.features()     # cached on the state (worlds/craft.py:297,328)
    # chains of pending transitions resolve generation by generation in one flush
    s = s0
    for a in (0, 0, 3, 3, 4):
        _, s = s.step(a)
    before = len(world._backend.runs)
    assert s.pos is not None
    assert len(world._backend.runs) - before == 5
    # setting the inventory (teachers / tests do) invalidates what was cached
    inv = np.zeros(world.cookbook.n_kinds)
    inv[world.cookbook.index["wood"]] = 2
    s.inventory = inv
    assert s.features()[378 + world.cookbook.index["wood"]] == 2
    task = world.task_manager["get[wood]"]
    assert s.satisfies(task) is True
    with pytest.raises(Exception, match="Unexpected action"):
        s.step(6)
