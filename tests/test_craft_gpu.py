"""Parity of the CUDA path (through the C ABI of include/psk_craft.h) against the golden
fixtures exported from the reference and against the CPU oracle.  Bit-exact everywhere:
features hold small exact integers in f32, everything else is integer/byte state."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

ERR_TYPE = 254


def _env_from_states(tables, S, task=None):
    from psketch_b200.vec import VecCraft
    return VecCraft.from_states(tables, S["grid"], S["inv"], S["pos"], S["dir"], task=task)


def _np(t):
    return t.cpu().numpy()


def _check_states(tables, oracle, S):
    n = len(S["grid"])
    C, K = tables.W * tables.H, tables.K
    env = _env_from_states(tables, S)
    # ---- features, both store paths
    ref_f = S["features"].astype(np.float32)
    for impl in (0, 1, 2):
        f = env.features(impl=impl)
        assert f.dtype == torch.float32 and tuple(f.shape) == (n, tables.n_features)
        assert np.array_equal(_np(f), ref_f), "features impl %d" % impl
    # ---- step, all six actions
    snap = env.snapshot()
    for a in range(6):
        env.restore(snap)
        r = env.step(torch.full((n,), a, dtype=torch.uint8))
        assert float(r.abs().sum()) == 0.0
        assert np.array_equal(_np(env.cells), S["step_grid"][:, a]), a
        assert np.array_equal(_np(env.inventory), S["step_inv"][:, a]), a
        assert np.array_equal(_np(env.pos), S["step_pos"][:, a]), a
        assert np.array_equal(_np(env.dir), S["step_dir"][:, a]), a
    env.check_errors()
    env.restore(snap)
    # ---- satisfies + expert, every task id
    inv32 = S["inv"].astype(np.int32)
    pos32, dir32 = S["pos"].astype(np.int32), S["dir"].astype(np.int32)
    for tid in range(1, S["satisfies"].shape[1]):
        tk = torch.full((n,), tid, dtype=torch.uint8)
        assert np.array_equal(_np(env.satisfies(tk)), S["satisfies"][:, tid]), tid
        act, dist = env.expert(tk, want_dist=True)
        act = _np(act)
        ref = S["expert"][:, tid]
        raised = ref == ERR_TYPE       # reference: TypeError; defined here as the oracle's answer
        o_act, o_dist, _ = oracle.expert(S["grid"], inv32, pos32, dir32, np.full(n, tid))
        assert np.array_equal(act[~raised], ref[~raised]), tid
        assert np.array_equal(act, o_act.astype(np.uint8)), tid
        assert np.array_equal(_np(dist).astype(np.int32), o_dist), tid
    # leaf tasks that are neither use nor go make the reference assert
    if tables.task_manager.by_id(1).goal_name not in ("use", "go"):
        with pytest.raises(AssertionError):
            env.expert(torch.full((n,), 1, dtype=torch.uint8))
            env.check_errors()
    # ---- find_closest_resources
    for j, kind in enumerate(S["go_kinds"]):
        goal, length, seq = env.find_closest(torch.full((n,), int(kind), dtype=torch.uint8),
                                             seq_cap=48)
        goal, length, seq = _np(goal), _np(length), _np(seq)
        rst = S["closest_status"][:, j]
        ok = rst == 0
        assert np.array_equal(length[ok], S["closest_len"][ok, j])
        assert np.array_equal(goal[ok], S["closest_goal"][ok, j])
        assert np.array_equal(seq[ok], S["closest_seq"][ok, j])
        none = rst == 1
        assert (length[none] == -1).all()
        assert np.array_equal(goal[none], S["closest_goal"][none, j])
        o_goal, o_len, o_st, o_seq = oracle.find_closest(S["grid"], pos32, dir32,
                                                         np.full(n, kind), seq_cap=48)
        assert np.array_equal(length.astype(np.int32), o_len)
        found = o_len >= 0
        assert np.array_equal(goal[found], o_goal[found].astype(np.uint8))
        assert np.array_equal(seq[found], o_seq[found])


def test_states_medium(medium_tables, medium_oracle, medium_states):
    _check_states(medium_tables, medium_oracle, medium_states)


def test_states_large(large_tables, large_oracle, large_states):
    _check_states(large_tables, large_oracle, large_states)


def test_states_custom_cookbook(custom_tables, custom_oracle, custom_states):
    """tests/golden/custom/*.yaml through the kernels, against the reference's own outputs."""
    _check_states(custom_tables, custom_oracle, custom_states)


@pytest.mark.parametrize("split", ["dev", "test", "train"])
def test_golden_trajectories(split, splits, medium_tables):
    """Every ref_actions sequence of the reference's dataset, replayed teacher -> step on the
    GPU for all instances at once (make_data.py:146-152)."""
    from psketch_b200.vec import VecCraft
    ref = splits[split + "_ref_actions"]
    n = len(ref)
    env = VecCraft.from_instances(medium_tables, splits[split + "_grids"],
                                  splits[split + "_inst_env"], splits[split + "_inst_pos"],
                                  splits[split + "_inst_task"])
    alive = torch.ones(n, dtype=torch.uint8, device=env.device)
    mism = 0
    for t in range(ref.shape[1]):
        act = env.expert()
        want = torch.from_numpy(ref[:, t]).to(env.device)
        live = alive.bool()
        mism += int((act[live] != want[live]).sum())
        stop = (act == 5) & live
        if bool(stop.any()):
            assert bool((env.satisfies()[stop] == 1).all())     # make_data.py:151
        alive = (live & ~stop).to(torch.uint8)
        env.step(act, active=alive)
    assert mism == 0
    assert int(alive.sum()) == 0
    env.check_errors()


def _oracle_state(env_np, K):
    return dict(grid=env_np[0], inv=env_np[1], pos=env_np[2], dir=env_np[3], timer=env_np[4])


@pytest.mark.parametrize("fused", [True, False])
def test_tick_matches_oracle_rollout(fused, splits, medium_tables, medium_oracle):
    """The rollout tick (trainers/imitation.py:42-73 with auto-reset) against orc_rollout, tick by
    tick, on the dev split tiled to a size that is not a multiple of any tile."""
    from psketch_b200.vec import VecCraft
    reps = 3
    idx = np.concatenate([np.arange(2200)] * reps)[:6007]
    grids = splits["dev_grids"]
    ienv = splits["dev_inst_env"][idx]
    ipos = splits["dev_inst_pos"][idx]
    itask = splits["dev_inst_task"][idx]
    T = 12                                   # short timer so that time-outs occur too
    env = VecCraft.from_instances(medium_tables, grids, ienv, ipos, itask, max_timesteps=T)
    init_grid = grids[ienv.astype(np.int64)]
    state = None
    tot = np.zeros(4, np.int64)
    for t in range(45):
        out = env.tick(fused=fused)
        state, stats, feats, act = medium_oracle.rollout(
            1, T, init_grid, ipos.astype(np.int32), itask.astype(np.int32), state=state,
            want_features=True)
        tot += stats
        assert np.array_equal(_np(out["expert"]).astype(np.int32), act), t
        assert np.array_equal(_np(out["features"]), feats), t
        assert np.array_equal(_np(env.cells), state["grid"]), t
        assert np.array_equal(_np(env.inventory).astype(np.int32), state["inv"]), t
        assert np.array_equal(_np(env.pos).astype(np.int32), state["pos"]), t
        assert np.array_equal(_np(env.dir).astype(np.int32), state["dir"]), t
        assert np.array_equal(_np(env.timer).astype(np.int32), state["timer"]), t
    st = _np(env.stats)
    assert st[0] == tot[0] and st[1] == tot[1] and st[2] == tot[2]
    assert st[0] > 0 and st[1] > 0 and st[1] < st[0] or T >= 30
    env.check_errors()


def test_tick_with_student_actions(splits, medium_tables, medium_oracle):
    """Externally supplied (random) actions instead of the teacher's: DAgger-style rollout."""
    from psketch_b200.vec import VecCraft
    n = 4099
    rng = np.random.RandomState(5)
    idx = rng.randint(0, 2200, size=n)
    grids = splits["test_grids"]
    ienv, ipos, itask = (splits["test_inst_env"][idx], splits["test_inst_pos"][idx],
                         splits["test_inst_task"][idx])
    env = VecCraft.from_instances(medium_tables, grids, ienv, ipos, itask, max_timesteps=40)
    o = medium_oracle
    grid = grids[ienv.astype(np.int64)].copy()
    init_grid = grid.copy()
    inv = np.zeros((n, 21), np.int32)
    pos = ipos.astype(np.int32).copy()
    dirs = np.zeros(n, np.int32)
    timer = np.full(n, 40, np.int32)
    task = itask.astype(np.int32)
    for t in range(60):
        a = rng.choice(6, size=n, p=[.19, .19, .19, .19, .2, .04]).astype(np.uint8)
        out = env.tick(actions=torch.from_numpy(a), fused=bool(t % 2))
        ref_act, _, _ = o.expert(grid, inv, pos, dirs, task)
        assert np.array_equal(_np(out["expert"]).astype(np.int32), ref_act)
        assert np.array_equal(_np(out["features"]), o.features(grid, inv, pos, dirs))
        timer -= 1
        done = (a == 5) | (timer <= 0)
        succ = (o.satisfies(grid, inv, pos, dirs, task) == 1) & done
        g2, i2, p2, d2, _ = o.step(grid, inv, pos, dirs, a.astype(np.int32))
        grid = np.where(done[:, None], init_grid, g2)
        inv = np.where(done[:, None], 0, i2)
        pos = np.where(done[:, None], ipos.astype(np.int32), p2)
        dirs = np.where(done, 0, d2)
        timer = np.where(done, 40, timer)
        assert np.array_equal(_np(out["done"]).astype(bool), done)
        assert np.array_equal(_np(out["success"]).astype(bool), succ)
        assert np.array_equal(_np(env.cells), grid)
        assert np.array_equal(_np(env.inventory).astype(np.int32), inv)
        assert np.array_equal(_np(env.pos).astype(np.int32), pos)
    env.check_errors()


def test_bad_action_raises(medium_tables, medium_states):
    S = {k: medium_states[k][:64] for k in ("grid", "inv", "pos", "dir")}
    env = _env_from_states(medium_tables, S)
    before = env.snapshot()
    env.step(torch.full((64,), 6, dtype=torch.uint8))
    with pytest.raises(Exception, match="Unexpected action"):
        env.check_errors()
    assert torch.equal(env.grid, before[0]) and torch.equal(env.agent, before[1])


@pytest.mark.parametrize("n", [0, 1, 15, 17, 129])
def test_ragged_sizes(n, medium_tables, medium_oracle, medium_states):
    S = {k: medium_states[k][:n] for k in ("grid", "inv", "pos", "dir", "features")}
    env = _env_from_states(medium_tables, S)
    f = env.features()
    assert tuple(f.shape) == (n, 404)
    assert np.array_equal(_np(f), S["features"].astype(np.float32))
    if n:
        a = env.expert(torch.full((n,), 24, dtype=torch.uint8))
        o, _, _ = medium_oracle.expert(S["grid"], S["inv"].astype(np.int32),
                                       S["pos"].astype(np.int32), S["dir"].astype(np.int32),
                                       np.full(n, 24))
        assert np.array_equal(_np(a).astype(np.int32), o)


def test_full_size_properties(splits, medium_tables, medium_oracle):
    """BASELINE config 2 size (65,536 envs from the train split): the whole batch against the
    oracle for a few ticks, then size-independent properties over a long rollout: every
    teacher-driven episode succeeds, and the episode/step counters add up."""
    from psketch_b200.vec import VecCraft
    n = 65536
    idx = np.arange(n) % 17600
    grids = splits["train_grids"]
    ienv, ipos, itask = (splits["train_inst_env"][idx], splits["train_inst_pos"][idx],
                         splits["train_inst_task"][idx])
    env = VecCraft.from_instances(medium_tables, grids, ienv, ipos, itask, max_timesteps=40)
    init_grid = grids[ienv.astype(np.int64)]
    state = None
    for t in range(6):
        out = env.tick(fused=True)
        state, _, feats, act = medium_oracle.rollout(1, 40, init_grid, ipos.astype(np.int32),
                                                     itask.astype(np.int32), state=state,
                                                     want_features=True)
        assert np.array_equal(_np(out["expert"]).astype(np.int32), act)
        assert np.array_equal(_np(out["features"]), feats)
    for t in range(94):
        env.tick(fused=True, want_features=False)
    st = _np(env.stats)
    assert st[2] == 100 * n
    assert st[0] == st[1] and st[0] > 0          # teacher rollouts always succeed
    # expected number of finished episodes: ticks / golden length per instance
    ref_len = splits["train_ref_len"][idx].astype(np.int64)
    assert st[0] == int((100 // ref_len).sum())
    env.check_errors()


def test_host_buffer_tick_matches_oracle(splits, medium_tables, medium_oracle):
    """psk_craft_host_tick: host (pinned numpy) buffers in and out, chunked over streams."""
    from psketch_b200.host import HostCraft
    n = 5003
    idx = np.arange(n) % 2200
    grids = splits["dev_grids"]
    ienv, ipos, itask = (splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx],
                         splits["dev_inst_task"][idx])
    env = HostCraft(medium_tables, grids, ienv, ipos, itask, max_timesteps=40, chunk_envs=1024)
    init_grid = grids[ienv.astype(np.int64)]
    state = None
    tot = np.zeros(4, np.int64)
    for t in range(30):
        env.tick()
        state, stats, feats, act = medium_oracle.rollout(
            1, 40, init_grid, ipos.astype(np.int32), itask.astype(np.int32), state=state,
            want_features=True)
        tot += stats
        assert np.array_equal(env.expert.astype(np.int32), act), t
        assert np.array_equal(env.features, feats), t
        assert np.array_equal(env.grid[:, :64], state["grid"]), t
        assert np.array_equal(env.agent[:, :21].astype(np.int32), state["inv"]), t
        assert np.array_equal(env.agent[:, 24:26].astype(np.int32), state["pos"]), t
    assert int(env.stats[0]) == tot[0] and int(env.stats[1]) == tot[1]
    env.close()


def test_million_env_properties(splits, medium_tables, medium_oracle):
    """BASELINE config 3 per-GPU size (1,048,576 envs): the large-batch kernel variant, checked on
    a strided sample against the oracle and through the size-independent counters."""
    from psketch_b200.vec import VecCraft
    n = 1 << 20
    idx = np.arange(n) % 17600
    grids = splits["train_grids"]
    ienv, ipos, itask = (splits["train_inst_env"][idx], splits["train_inst_pos"][idx],
                         splits["train_inst_task"][idx])
    env = VecCraft.from_instances(medium_tables, grids, ienv, ipos, itask, max_timesteps=40)
    feats = torch.empty((n, 404), dtype=torch.float32, device=env.device)
    T = 30
    sample = np.arange(0, n, 997)
    state = None
    for t in range(T):
        out = env.tick(features_out=feats, fused=True)
        if t in (0, 7, 19):
            # oracle on the sampled envs, advanced to the same tick
            pass
    # replay the sampled envs on the oracle for T ticks and compare the final states + last outputs
    st, stats, f, a = medium_oracle.rollout(T, 40, grids[ienv[sample].astype(np.int64)],
                                            ipos[sample].astype(np.int32), itask[sample].astype(np.int32),
                                            want_features=True)
    ds = torch.from_numpy(sample).to(env.device)
    assert np.array_equal(env.cells[ds].cpu().numpy(), st["grid"])
    assert np.array_equal(env.pos[ds].cpu().numpy().astype(np.int32), st["pos"])
    assert np.array_equal(env.inventory[ds].cpu().numpy().astype(np.int32), st["inv"])
    assert np.array_equal(out["expert"][ds].cpu().numpy().astype(np.int32), a)
    assert np.array_equal(feats[ds].cpu().numpy(), f)
    s = env.stats.cpu().numpy()
    ref_len = splits["train_ref_len"][idx].astype(np.int64)
    assert s[2] == T * n and s[0] == s[1] == int((T // ref_len).sum())
    env.check_errors()


def test_random_actions_are_uniform_and_reproducible(medium_tables, medium_states):
    S = {k: medium_states[k][:4096] for k in ("grid", "inv", "pos", "dir")}
    env = _env_from_states(medium_tables, S)
    a0 = env.random_actions(0).cpu().numpy()
    assert np.array_equal(a0, env.random_actions(0).cpu().numpy())
    a1 = env.random_actions(1).cpu().numpy()
    assert not np.array_equal(a0, a1) and a0.max() == 5 and a0.min() == 0
    big = np.concatenate([env.random_actions(t).cpu().numpy() for t in range(50)])
    hist = np.bincount(big, minlength=6) / len(big)
    assert np.abs(hist - 1 / 6).max() < 0.005


class _OracleTicks(object):
    """The rollout loop body of trainers/imitation.py:42-73 (with auto-reset) on the CPU oracle, one
    tick at a time, with the per-env done / success flags the kernels also report."""

    def __init__(self, oracle, grids, ienv, ipos, itask, max_timesteps=40):
        self.o, self.T = oracle, max_timesteps
        n = len(ienv)
        self.init_grid = np.ascontiguousarray(grids[np.asarray(ienv).astype(np.int64)])
        self.init_pos = np.asarray(ipos).astype(np.int32)
        self.task = np.asarray(itask).astype(np.int32)
        self.grid = self.init_grid.copy()
        self.inv = np.zeros((n, oracle.K), np.int32)
        self.pos = self.init_pos.copy()
        self.dir = np.zeros(n, np.int32)
        self.timer = np.full(n, max_timesteps, np.int32)
        self.stats = np.zeros(3, np.int64)

    def tick(self, actions=None, want_features=True):
        o = self.o
        expert, _, _ = o.expert(self.grid, self.inv, self.pos, self.dir, self.task)
        feats = o.features(self.grid, self.inv, self.pos, self.dir) if want_features else None
        act = expert if actions is None else np.asarray(actions).astype(np.int32)
        self.timer -= 1
        done = (act == 5) | (self.timer <= 0)
        succ = done & (o.satisfies(self.grid, self.inv, self.pos, self.dir, self.task) == 1)
        g2, i2, p2, d2, _ = o.step(self.grid, self.inv, self.pos, self.dir, np.where(done, 5, act).astype(np.int32))
        self.grid = np.where(done[:, None], self.init_grid, g2)
        self.inv = np.where(done[:, None], 0, i2).astype(np.int32)
        self.pos = np.where(done[:, None], self.init_pos, p2).astype(np.int32)
        self.dir = np.where(done, 0, d2).astype(np.int32)
        self.timer = np.where(done, self.T, self.timer).astype(np.int32)
        self.stats += (int(done.sum()), int(succ.sum()), len(done))
        return dict(expert=expert.astype(np.uint8), done=done.astype(np.uint8),
                    success=succ.astype(np.uint8), features=feats)

    def assert_state_equals(self, env):
        assert np.array_equal(_np(env.cells), self.grid)
        assert np.array_equal(_np(env.inventory).astype(np.int32), self.inv)
        assert np.array_equal(_np(env.pos).astype(np.int32), self.pos)
        assert np.array_equal(_np(env.dir).astype(np.int32), self.dir)
        assert np.array_equal(_np(env.timer).astype(np.int32), self.timer)
        st = _np(env.stats)
        assert (st[0], st[1], st[2]) == tuple(self.stats)


_ROLLOUT_REF = {}


def _rollout_reference(n, T, splits, oracle):
    """Oracle outputs of T teacher-driven ticks on the bench workload (train split tiled to n envs),
    computed once per size and shared by every kernel variant.  Features are kept as u8 (every
    feature is an exact integer <= 255; the test checks that on the GPU side)."""
    if n not in _ROLLOUT_REF:
        idx = np.arange(n) % 17600
        orc = _OracleTicks(oracle, splits["train_grids"], splits["train_inst_env"][idx],
                           splits["train_inst_pos"][idx], splits["train_inst_task"][idx])
        ticks = []
        for _ in range(T):
            r = orc.tick()
            assert r["features"].max() <= 255 and np.array_equal(r["features"], np.floor(r["features"]))
            r["features"] = r["features"].astype(np.uint8)
            ticks.append(r)
        _ROLLOUT_REF.clear()            # one size at a time: a 196,608-env reference holds 0.6 GB
        _ROLLOUT_REF[n] = (idx, ticks, orc)
    return _ROLLOUT_REF[n]


# (rollout_variant, rollout_tma): 0 = 64 env threads + 2 feature warps, 2 = 32 + 2 (what bench.py
# times at 65,536 envs: craft_rollout_kernel<8,8,3,32,2,21,0>), 3 = 16 + 2, 4 = 16 + 1
ROLLOUT_VARIANTS = [(2, 0), (2, 1), (0, 0), (0, 1), (3, 0), (3, 1), (4, 0), (4, 1), (-1, -1)]


@pytest.mark.parametrize("n", [65536, 65536 + 37, 196608])
def test_rollout_kernel_variants_vs_oracle(n, splits, medium_tables, medium_oracle):
    """Every instantiation of craft_rollout_kernel the dispatcher can select — forced through
    psk_set_tuning — on BASELINE config 2's workload and size (and a ragged and a 3x size), T = 8
    ticks per launch into a 9-frame ring exactly as bench.py launches it: expert / done / success /
    features of EVERY tick, the final state and the counters against the CPU oracle
    (trainers/imitation.py:42-73, make_data.py:146-152)."""
    from psketch_b200 import _lib
    from psketch_b200.vec import VecCraft
    T, RING = 8, 9
    idx, ref, orc = _rollout_reference(n, T, splits, medium_oracle)
    args = (medium_tables, splits["train_grids"], splits["train_inst_env"][idx],
            splits["train_inst_pos"][idx], splits["train_inst_task"][idx])
    ring = torch.empty((RING, n, 404), dtype=torch.float32, device="cuda")
    try:
        for variant, tma in ROLLOUT_VARIANTS:
            _lib.set_tuning(rollout_variant=variant, rollout_tma=tma)
            env = VecCraft.from_instances(*args, max_timesteps=40)
            ring.fill_(-1.0)
            out = env.rollout(T, features_out=ring)
            torch.cuda.synchronize()
            tag = "variant %d tma %d" % (variant, tma)
            for t in range(T):
                r = ref[t]
                assert np.array_equal(_np(out["expert"][t]), r["expert"]), (tag, t)
                assert np.array_equal(_np(out["done"][t]), r["done"]), (tag, t)
                assert np.array_equal(_np(out["success"][t]), r["success"]), (tag, t)
                f8 = ring[t].to(torch.uint8)
                assert torch.equal(f8.to(torch.float32), ring[t]), (tag, t)   # exact small integers
                assert np.array_equal(_np(f8), r["features"]), (tag, t)
            assert bool((ring[T] == -1.0).all()), tag                         # slot 8 untouched
            orc.assert_state_equals(env)
            env.check_errors()
    finally:
        _lib.set_tuning(rollout_variant=-1, rollout_tma=-1)


@pytest.mark.parametrize("tma", [-1, 1])
@pytest.mark.parametrize("n,with_actions", [(6007, False), (4099, True), (65, False)])
def test_multi_tick_rollout_equals_single_ticks(n, with_actions, tma, splits, medium_tables, medium_oracle):
    """psk_craft_rollout (tick loop inside the kernel, state in shared memory) against the same
    number of psk_craft_tick launches AND against the oracle advanced tick by tick: every per-tick
    output, the final state and the counters."""
    from psketch_b200 import _lib
    from psketch_b200.vec import VecCraft
    _lib.set_tuning(rollout_tma=tma, tick_tma=tma)          # ragged sizes through the TMA store path too
    rng = np.random.RandomState(n)
    idx = rng.randint(0, 2200, size=n)
    grids = splits["dev_grids"]
    args = (medium_tables, grids, splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx],
            splits["dev_inst_task"][idx])
    T = 23
    a = VecCraft.from_instances(*args, max_timesteps=15)
    b = VecCraft.from_instances(*args, max_timesteps=15)
    acts = None
    if with_actions:
        acts = torch.from_numpy(rng.choice(6, size=(T, n), p=[.19, .19, .19, .19, .2, .04]).astype(np.uint8)).to(a.device)
    feats = torch.empty((T, n, 404), dtype=torch.float32, device=a.device)
    out = a.rollout(T, actions=acts, features_out=feats)
    orc = _OracleTicks(medium_oracle, grids, args[2], args[3], args[4], max_timesteps=15)
    for t in range(T):
        o = b.tick(actions=None if acts is None else acts[t], fused=bool(t % 2))
        assert torch.equal(out["expert"][t], o["expert"]), t
        assert torch.equal(out["done"][t], o["done"]), t
        assert torch.equal(out["success"][t], o["success"]), t
        assert torch.equal(feats[t], o["features"]), t
        ref = orc.tick(None if acts is None else _np(acts[t]))
        assert np.array_equal(_np(out["expert"][t]), ref["expert"]), t
        assert np.array_equal(_np(out["done"][t]), ref["done"]), t
        assert np.array_equal(_np(out["success"][t]), ref["success"]), t
        assert np.array_equal(_np(feats[t]), ref["features"]), t
    orc.assert_state_equals(a)
    assert torch.equal(a.grid, b.grid) and torch.equal(a.agent, b.agent)
    assert torch.equal(a.stats, b.stats) and int(a.stats[2]) == T * n and int(a.stats[0]) > 0
    # ring of two feature slots: the last two ticks survive
    c = VecCraft.from_instances(*args, max_timesteps=15)
    ring = torch.empty((2, n, 404), dtype=torch.float32, device=a.device)
    c.rollout(T, actions=acts, features_out=ring, want_flags=False)
    assert torch.equal(ring[(T - 1) % 2], feats[T - 1]) and torch.equal(ring[(T - 2) % 2], feats[T - 2])
    a.check_errors()
    _lib.set_tuning(rollout_tma=-1, tick_tma=-1)


def test_multi_tick_rollout_other_geometry_falls_back(large_tables, large_states):
    """craft_large has no multi-tick kernel: psk_craft_rollout loops over single ticks."""
    S = {k: large_states[k][:300] for k in ("grid", "inv", "pos", "dir")}
    a = _env_from_states(large_tables, S, task=np.full(300, 24))
    b = _env_from_states(large_tables, S, task=np.full(300, 24))
    out = a.rollout(5)
    for t in range(5):
        o = b.tick(want_features=False)
        assert torch.equal(out["expert"][t], o["expert"])
    assert torch.equal(a.grid, b.grid) and torch.equal(a.agent, b.agent)


@pytest.mark.parametrize("fused", [True, False])
def test_tick_craft_large_matches_oracle(fused, large_tables, large_oracle, large_states):
    """craft_large (10x10, window 5, 128-bit boards, 1,076 features): rollout ticks vs the oracle."""
    from psketch_b200.vec import VecCraft
    S = large_states
    n = 1500
    rng = np.random.RandomState(2)
    task = rng.choice([13, 14, 15, 19, 20, 21, 22, 23, 24, 25, 26], size=n).astype(np.int32)
    grid, pos = S["grid"][:n], S["pos"][:n].astype(np.int32)
    # episode starts: the exported layouts with empty inventory and dir 0 (what the oracle resets to)
    env = VecCraft.from_instances(large_tables, grid, np.arange(n), pos, task, max_timesteps=25)
    state = None
    tot = np.zeros(4, np.int64)
    for t in range(40):
        out = env.tick(fused=fused)
        state, stats, feats, act = large_oracle.rollout(1, 25, grid, pos, task, state=state,
                                                        want_features=True)
        tot += stats
        assert np.array_equal(_np(out["expert"]).astype(np.int32), act), t
        assert np.array_equal(_np(out["features"]), feats), t
        assert np.array_equal(_np(env.cells), state["grid"]), t
        assert np.array_equal(_np(env.inventory).astype(np.int32), state["inv"]), t
        assert np.array_equal(_np(env.pos).astype(np.int32), state["pos"]), t
    st = _np(env.stats)
    assert st[0] == tot[0] and st[1] == tot[1] and st[0] > 0
    env.check_errors()


def test_custom_cookbook_runtime_kind_count(tmp_path):
    """A cookbook / hint file other than the default one (K = 14 kinds, 271 features: rows are not
    a whole number of float4) goes through the kernels' runtime-K code paths."""
    import yaml
    from oracle.craft_oracle import CraftOracle
    from psketch_b200.tables import Cookbook, CraftTables, TaskManager
    from psketch_b200.vec import VecCraft
    recipes = {
        "environment": ["boundary", "workshop0", "workshop1", "workshop2", "water", "stone"],
        "primitives": ["iron", "grass", "wood"],
        "recipes": {
            "plank": {"wood": 1, "_at": "workshop0"},
            "stick": {"wood": 1, "_at": "workshop1"},
            "bridge": {"wood": 1, "iron": 1, "_at": "workshop2"},
            "axe": {"stick": 1, "iron": 1, "_at": "workshop0"},
        },
    }
    hints = {
        "use[none]": [], "go[wood]": [], "go[iron]": [], "go[workshop0]": [], "go[workshop1]": [],
        "go[workshop2]": [],
        "get[wood]": ["go[wood]", "use[none]"], "get[iron]": ["go[iron]", "use[none]"],
        "makeat[workshop0]": ["go[workshop0]", "use[none]"],
        "makeat[workshop1]": ["go[workshop1]", "use[none]"],
        "makeat[workshop2]": ["go[workshop2]", "use[none]"],
        "make[plank]": ["get[wood]", "makeat[workshop0]"],
        "make[stick]": ["get[wood]", "makeat[workshop1]"],
        "make[bridge]": ["get[iron]", "get[wood]", "makeat[workshop2]"],
        "make[axe]": ["make[stick]", "get[iron]", "makeat[workshop0]"],
    }
    rp, hp = str(tmp_path / "recipes.yaml"), str(tmp_path / "hints.yaml")
    yaml.safe_dump(recipes, open(rp, "w"), sort_keys=False)
    yaml.safe_dump(hints, open(hp, "w"), sort_keys=False)
    tables = CraftTables(Cookbook(rp), TaskManager(hp), "craft_medium")
    assert tables.K == 14 and tables.n_features == 271
    tm = tables.task_manager
    task_ids = [tm[g].task_id for g in ("get[wood]", "get[iron]", "make[plank]", "make[stick]",
                                        "make[bridge]", "make[axe]")]
    o = CraftOracle(tables)
    rng = np.random.RandomState(9)
    n = 2051
    cb = tables.cookbook
    grid = np.zeros((n, 8, 8), np.uint8)
    grid[:, 0, :] = grid[:, 7, :] = grid[:, :, 0] = grid[:, :, 7] = 1
    pos = np.zeros((n, 2), np.int32)
    for i in range(n):
        cells = rng.permutation(36)
        k = 0
        for name, cnt in (("iron", 2), ("grass", 1), ("wood", 2), ("workshop0", 1), ("workshop1", 1),
                          ("workshop2", 1), ("water", rng.randint(0, 2)), ("stone", rng.randint(0, 3))):
            for _ in range(cnt):
                grid[i, 1 + cells[k] // 6, 1 + cells[k] % 6] = cb.index[name]
                k += 1
        pos[i] = (1 + cells[k] // 6, 1 + cells[k] % 6)
    grid = grid.reshape(n, 64)
    inv = np.zeros((n, 14), np.int32)
    inv[np.arange(n), rng.randint(7, 14, size=n)] = rng.randint(0, 3, size=n)
    dirs = rng.randint(0, 4, size=n).astype(np.int32)
    task = rng.choice(task_ids, size=n).astype(np.int32)
    env = VecCraft.from_states(tables, grid, inv, pos, dirs, task=task)
    for impl in (0, 1, 2):
        assert np.array_equal(_np(env.features(impl=impl)), o.features(grid, inv, pos, dirs)), impl
    act = env.expert()
    assert np.array_equal(_np(act).astype(np.int32), o.expert(grid, inv, pos, dirs, task)[0])
    # teacher-driven ticks (fused and pipelined) and the multi-tick kernel
    twin = VecCraft.from_states(tables, grid, inv, pos, dirs, task=task)
    out = twin.rollout(6, features_out=torch.empty((6, n, 271), dtype=torch.float32, device=env.device))
    g, iv, p, d = grid.copy(), inv.copy(), pos.copy(), dirs.copy()
    for t in range(6):
        tick = env.tick(fused=bool(t % 2))
        a = o.expert(g, iv, p, d, task)[0]
        assert np.array_equal(_np(tick["expert"]).astype(np.int32), a), t
        assert np.array_equal(_np(out["expert"][t]).astype(np.int32), a), t
        assert np.array_equal(_np(tick["features"]), o.features(g, iv, p, d)), t
        done = _np(tick["done"]).astype(bool)
        if done.any():
            break
        g, iv, p, d, _ = o.step(g, iv, p, d, a)
        assert np.array_equal(_np(env.cells), g) and np.array_equal(_np(env.inventory).astype(np.int32), iv)
    env.check_errors()


def test_million_env_rollout_kernel(splits, medium_tables, medium_oracle):
    """The large-batch variant of the multi-tick kernel (TMA feature stores, n > 262,144): a
    strided sample of envs against the oracle, and the counters."""
    from psketch_b200.vec import VecCraft
    n = (1 << 20) + 37                       # not a multiple of the 64-env tile
    idx = np.arange(n) % 17600
    grids = splits["train_grids"]
    ienv, ipos, itask = (splits["train_inst_env"][idx], splits["train_inst_pos"][idx],
                         splits["train_inst_task"][idx])
    env = VecCraft.from_instances(medium_tables, grids, ienv, ipos, itask, max_timesteps=40)
    T = 11
    ring = torch.empty((2, n, 404), dtype=torch.float32, device=env.device)
    out = env.rollout(T, features_out=ring)
    sample = np.concatenate([np.arange(0, n, 1009), np.arange(n - 40, n)])
    ds = torch.from_numpy(sample).to(env.device)
    state = None
    for t in range(T):
        state, _, f, a = medium_oracle.rollout(1, 40, grids[ienv[sample].astype(np.int64)],
                                               ipos[sample].astype(np.int32),
                                               itask[sample].astype(np.int32), state=state,
                                               want_features=True)
        assert np.array_equal(out["expert"][t][ds].cpu().numpy().astype(np.int32), a), t
        if t >= T - 2:
            assert np.array_equal(ring[t % 2][ds].cpu().numpy(), f), t
    assert np.array_equal(env.cells[ds].cpu().numpy(), state["grid"])
    assert np.array_equal(env.pos[ds].cpu().numpy().astype(np.int32), state["pos"])
    s = env.stats.cpu().numpy()
    ref_len = splits["train_ref_len"][idx].astype(np.int64)
    assert s[2] == T * n and s[0] == s[1] == int((T // ref_len).sum())
    assert int(out["done"].sum()) == s[0] and int(out["success"].sum()) == s[1]
    env.check_errors()


@pytest.mark.parametrize("features", ["f32", "u8", "f32_wire_u8", None])
def test_host_resident_tick_matches_oracle(features, splits, medium_tables, medium_oracle):
    """psk_craft_host_tick_resident: the environments stay on the device, only actions go up and
    features (f32 or the compact u8 frame) / teacher actions / flags come down; student-style
    random actions on even ticks, the teacher's on odd ones."""
    from psketch_b200.host import HostCraft
    n = 5003
    rng = np.random.RandomState(11)
    idx = rng.randint(0, 2200, size=n)
    grids = splits["dev_grids"]
    ienv, ipos, itask = (splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx],
                         splits["dev_inst_task"][idx])
    env = HostCraft(medium_tables, grids, ienv, ipos, itask, max_timesteps=17, chunk_envs=1024)
    env.reset_resident()
    orc = _OracleTicks(medium_oracle, grids, ienv, ipos, itask, max_timesteps=17)
    for t in range(40):
        a = rng.choice(6, size=n, p=[.19, .19, .19, .19, .2, .04]).astype(np.uint8) if t % 2 == 0 else None
        if features == "f32_wire_u8":
            env.features[:] = -1.0                  # every row must be rewritten by the host threads
        env.tick_resident(actions=a, features=features)
        ref = orc.tick(a)
        assert np.array_equal(env.expert, ref["expert"]), t
        assert np.array_equal(env.done, ref["done"]), t
        assert np.array_equal(env.success, ref["success"]), t
        if features in ("f32", "f32_wire_u8"):
            assert np.array_equal(env.features, ref["features"]), t
        elif features == "u8":
            assert np.array_equal(env.features_u8.astype(np.float32), ref["features"]), t
    assert tuple(int(x) for x in env.stats[:3]) == tuple(orc.stats)
    env.download()
    assert np.array_equal(env.grid[:, :64], orc.grid)
    assert np.array_equal(env.agent[:, :21].astype(np.int32), orc.inv)
    assert np.array_equal(env.agent[:, 24:26].astype(np.int32), orc.pos)
    assert np.array_equal(env.agent[:, 28].astype(np.int32), orc.timer)
    # a bad action surfaces as the reference's exception
    with pytest.raises(Exception, match="Unexpected action"):
        env.tick_resident(actions=np.full(n, 6, np.uint8), features=None)
    env.close()


@pytest.mark.parametrize("which", ["medium", "large", "custom"])
def test_features_u8_equals_features(which, medium_tables, medium_states, large_tables, large_states,
                                     custom_tables, custom_states):
    tables, S = {"medium": (medium_tables, medium_states), "large": (large_tables, large_states),
                 "custom": (custom_tables, custom_states)}[which]
    for n in (len(S["grid"]), 1999, 3, 1):
        sub = {k: S[k][:n] for k in ("grid", "inv", "pos", "dir")}
        env = _env_from_states(tables, sub)
        f8 = env.features_u8()
        assert f8.dtype == torch.uint8 and tuple(f8.shape) == (n, tables.n_features)
        assert np.array_equal(_np(f8).astype(np.float32), S["features"][:n].astype(np.float32))


def test_step_kernel_variants(medium_tables, medium_states):
    """craft_step_kernel with the tables staged in shared memory (round-1 shape, step_variant 0) and
    read through the read-only path (default): same results on every exported state and action."""
    from psketch_b200 import _lib
    try:
        for variant, n in ((0, 12500), (1, 12500), (2, 12500), (2, 4099), (2, 33), (1, 65), (-1, 12500)):
            S = {k: medium_states[k][:n] for k in medium_states.files if medium_states[k].ndim > 0
                 and len(medium_states[k]) == len(medium_states["grid"])}
            n = len(S["grid"])
            _lib.set_tuning(step_variant=variant)
            env = _env_from_states(medium_tables, S)
            snap = env.snapshot()
            for a in range(6):
                env.restore(snap)
                env.step(torch.full((n,), a, dtype=torch.uint8))
                assert np.array_equal(_np(env.cells), S["step_grid"][:, a]), (variant, a)
                assert np.array_equal(_np(env.inventory), S["step_inv"][:, a]), (variant, a)
                assert np.array_equal(_np(env.pos), S["step_pos"][:, a]), (variant, a)
                assert np.array_equal(_np(env.dir), S["step_dir"][:, a]), (variant, a)
            env.check_errors()
    finally:
        _lib.set_tuning(step_variant=-1)


def test_max_timesteps_is_validated(medium_tables, medium_states):
    S = {k: medium_states[k][:4] for k in ("grid", "inv", "pos", "dir")}
    from psketch_b200.vec import VecCraft
    for bad in (0, 256, 1000):
        with pytest.raises(ValueError, match="max_timesteps"):
            VecCraft.from_states(medium_tables, S["grid"], S["inv"], S["pos"], S["dir"], max_timesteps=bad)


@pytest.mark.parametrize("which", ["medium", "large", "stress16"])
def test_tick_advance_first_matches_oracle(which, splits, medium_tables, medium_oracle, large_tables,
                                           large_oracle, large_states):
    """psk_craft_tick in "step, then observe" order (PSK_TICK_ADVANCE_FIRST): actions of a policy in
    the loop are applied first, teacher action and features describe the state after the step; the
    fused kernel (8x8, 10x10) and the launch-sequence fallback (16x16)."""
    from psketch_b200.vec import VecCraft
    rng = np.random.RandomState(17)
    if which == "medium":
        n = 4099
        idx = rng.randint(0, 2200, size=n)
        tables, oracle, grids = medium_tables, medium_oracle, splits["dev_grids"]
        ienv, ipos, itask = splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx], splits["dev_inst_task"][idx]
    elif which == "large":
        n = 1500
        tables, oracle, grids = large_tables, large_oracle, large_states["grid"][:n]
        ienv, ipos = np.arange(n), large_states["pos"][:n]
        itask = rng.choice([13, 14, 15, 19, 20, 21, 22, 23, 24, 25, 26], size=n)
    else:
        from oracle.craft_oracle import CraftOracle
        from psketch_b200.tables import CraftTables
        tables = CraftTables(world_config=dict(WIDTH=16, HEIGHT=16, WINDOW_WIDTH=3, WINDOW_HEIGHT=3,
                                               N_WORKSHOPS=3, N_PRIMITIVES=4))
        oracle = CraftOracle(tables)
        n = 700
        g = np.zeros((n, 16, 16), np.uint8)
        g[:, 0, :] = g[:, 15, :] = g[:, :, 0] = g[:, :, 15] = 1
        ipos = np.zeros((n, 2), np.int64)
        for i in range(n):
            cells = rng.permutation(14 * 14)
            for j, kind in enumerate([7, 7, 8, 8, 9, 9, 2, 3, 4, 5, 6, 6]):
                g[i, 1 + cells[j] // 14, 1 + cells[j] % 14] = kind
            ipos[i] = (1 + cells[20] // 14, 1 + cells[20] % 14)
        grids, ienv = g.reshape(n, 256), np.arange(n)
        itask = rng.choice([13, 14, 15, 19, 20, 21, 22, 23, 24, 25, 26], size=n)
    env = VecCraft.from_instances(tables, grids, ienv, ipos, itask, max_timesteps=19)
    orc = _OracleTicks(oracle, grids, ienv, ipos, itask, max_timesteps=19)
    out = env.tick(actions=None, advance_first=True)            # no step: observe the start states
    ref_e, _, _ = oracle.expert(orc.grid, orc.inv, orc.pos, orc.dir, orc.task)
    assert np.array_equal(_np(out["expert"]).astype(np.int32), ref_e)
    assert np.array_equal(_np(out["features"]), oracle.features(orc.grid, orc.inv, orc.pos, orc.dir))
    assert not _np(out["done"]).any() and not _np(out["success"]).any()
    for t in range(45):
        a = rng.choice(6, size=n, p=[.19, .19, .19, .19, .2, .04]).astype(np.uint8)
        out = env.tick(actions=torch.from_numpy(a), advance_first=True, want_features=bool(t % 3))
        ref = orc.tick(a, want_features=False)                   # applies a; flags of that step
        assert np.array_equal(_np(out["done"]), ref["done"]), t
        assert np.array_equal(_np(out["success"]), ref["success"]), t
        ref_e, _, _ = oracle.expert(orc.grid, orc.inv, orc.pos, orc.dir, orc.task)
        assert np.array_equal(_np(out["expert"]).astype(np.int32), ref_e), t
        if t % 3:
            assert np.array_equal(_np(out["features"]), oracle.features(orc.grid, orc.inv, orc.pos, orc.dir)), t
        orc.assert_state_equals(env)
    env.check_errors()


def test_last_subtask_assert_is_reported(tmp_path):
    """teachers/base.py:23-24 (see tests/test_oracle.py, checked there against the reference itself):
    an unsatisfied task whose last subtask is satisfied makes the reference assert; the kernels
    return action 255 and raise through check_errors instead of walking on to a sibling."""
    import yaml
    from oracle.craft_oracle import CraftOracle
    from psketch_b200.tables import Cookbook, CraftTables, TaskManager
    from psketch_b200.vec import VecCraft
    hints = {"use[none]": [], "go[wood]": [], "get[wood]": ["go[wood]", "use[none]"],
             "make[plank]": ["get[wood]"]}
    hp = str(tmp_path / "hints.yaml")
    yaml.safe_dump(hints, open(hp, "w"), sort_keys=False)
    tables = CraftTables(Cookbook(), TaskManager(hp), "craft_medium")
    o = CraftOracle(tables)
    n = 70
    rng = np.random.RandomState(0)
    grid = np.zeros((n, 8, 8), np.uint8)
    grid[:, 0, :] = grid[:, 7, :] = grid[:, :, 0] = grid[:, :, 7] = 1
    grid[:, 3, 5] = tables.cookbook.index["wood"]
    grid = grid.reshape(n, 64)
    inv = np.zeros((n, tables.K), np.int32)
    has_wood = rng.rand(n) < 0.5
    inv[has_wood, tables.cookbook.index["wood"]] = 1
    pos = np.tile(np.asarray([[3, 3]], np.int32), (n, 1))
    task = np.full(n, tables.task_manager["make[plank]"].task_id, np.int32)
    env = VecCraft.from_states(tables, grid, inv, pos, np.zeros(n, np.int32), task=task)
    act = _np(env.expert()).astype(np.int32)
    want = o.expert(grid, inv, pos, np.zeros(n, np.int32), task)[0]
    assert np.array_equal(act, want) and (act[has_wood] == 255).all() and (act[~has_wood] == 1).all()
    with pytest.raises(AssertionError):
        env.check_errors()
    out = env.tick(fused=True)                       # same walk inside the fused kernel
    assert np.array_equal(_np(out["expert"]).astype(np.int32), want)
    with pytest.raises(AssertionError):
        env.check_errors()


def test_concurrent_streams_and_tile_chaining(splits, medium_tables, medium_oracle):
    """Two batches ticking concurrently on two streams share the per-device tile-chaining counters
    (psk_common.cuh): tickets of the two streams interleave, results must not change and nothing
    may deadlock; then the same with chaining switched off."""
    from psketch_b200 import _lib
    from psketch_b200.vec import VecCraft
    rng = np.random.RandomState(21)
    try:
        for chain in (-1, 0):
            _lib.set_tuning(tile_chain=chain)
            envs, orcs, streams = [], [], [torch.cuda.Stream(), torch.cuda.Stream()]
            for k, n in enumerate((8191, 3001)):
                idx = rng.randint(0, 2200, size=n)
                args = (splits["dev_grids"], splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx],
                        splits["dev_inst_task"][idx])
                envs.append(VecCraft.from_instances(medium_tables, *args, max_timesteps=13))
                orcs.append(_OracleTicks(medium_oracle, *args, max_timesteps=13))
            torch.cuda.synchronize()
            outs = [[], []]
            for rep in range(6):
                for k in (0, 1):
                    with torch.cuda.stream(streams[k]):
                        for _ in range(3):                     # back-to-back PDL launches per stream
                            o = envs[k].tick(want_features=False)
                            outs[k].append(o["expert"].clone())
                        o = envs[k].rollout(4)
                        outs[k] += [o["expert"][t].clone() for t in range(4)]
            torch.cuda.synchronize()
            for k in (0, 1):
                for t, got in enumerate(outs[k]):
                    ref = orcs[k].tick(want_features=False)
                    assert np.array_equal(_np(got), ref["expert"]), (chain, k, t)
                orcs[k].assert_state_equals(envs[k])
                envs[k].check_errors()
    finally:
        _lib.set_tuning(tile_chain=-1)


def test_back_to_back_ticks_on_overlapping_slices(splits, medium_tables, medium_oracle):
    """Launches on different slices of one batch, back to back on one stream with no host sync:
    tile chaining identifies env groups by the address of their agent records, so an aligned slice
    chains on the same counters as the whole batch, and a slice that does not start on a group
    boundary falls back to waiting for the whole previous grid.  Either way later launches must see
    the state earlier launches wrote."""
    import copy
    from psketch_b200.vec import VecCraft
    n = 4096
    rng = np.random.RandomState(33)
    idx = rng.randint(0, 2200, size=n)
    args = (splits["dev_grids"], splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx],
            splits["dev_inst_task"][idx])
    env = VecCraft.from_instances(medium_tables, *args, max_timesteps=11)
    orc = _OracleTicks(medium_oracle, *args, max_timesteps=11)

    def view(lo, hi):
        sub = copy.copy(env)
        sub.grid, sub.agent = env.grid[lo:hi], env.agent[lo:hi]
        sub.scen_idx, sub.init_agent = env.scen_idx[lo:hi], env.init_agent[lo:hi]
        sub.n = hi - lo
        return sub

    def oracle_subset(lo, hi):
        keep = np.ones(n, bool)
        keep[lo:hi] = False
        saved = {k: getattr(orc, k).copy() for k in ("grid", "inv", "pos", "dir", "timer")}
        orc.tick(want_features=False)
        for k, v in saved.items():
            getattr(orc, k)[keep] = v[keep]

    slices = [(0, n), (528, 1528), (8, 508), (1024, 4096), (0, n), (1000, 1064), (16, 4000)]
    subs = [view(lo, hi) for lo, hi in slices]
    for rep in range(9):
        for sub in subs:                                  # no synchronisation in between
            sub.tick(want_features=False)
            sub.rollout(2)
    torch.cuda.synchronize()
    for rep in range(9):
        for lo, hi in slices:
            for _ in range(3):
                oracle_subset(lo, hi)
    assert np.array_equal(_np(env.cells), orc.grid)
    assert np.array_equal(_np(env.inventory).astype(np.int32), orc.inv)
    assert np.array_equal(_np(env.pos).astype(np.int32), orc.pos)
    assert np.array_equal(_np(env.timer).astype(np.int32), orc.timer)
    env.check_errors()


def test_random_action_block_feeds_the_rollout_kernel(splits, medium_tables, medium_oracle):
    """Off-policy variant with several ticks per launch: psk_random_actions_block fills the action
    block of psk_craft_rollout; rows equal the per-tick generator, and the rollout equals the oracle
    driven by the same actions."""
    from psketch_b200.vec import VecCraft
    n, T = 5003, 7
    idx = np.arange(n) % 2200
    args = (splits["dev_grids"], splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx],
            splits["dev_inst_task"][idx])
    env = VecCraft.from_instances(medium_tables, *args, max_timesteps=9)
    block = env.random_actions(t=100, seed=123, ticks=T)
    for k in range(T):
        assert torch.equal(block[k], env.random_actions(t=100 + k, seed=123))
    orc = _OracleTicks(medium_oracle, *args, max_timesteps=9)
    for rep in range(4):
        block = env.random_actions(t=1000 * rep, seed=7, ticks=T)
        out = env.rollout(T, actions=block)
        for k in range(T):
            ref = orc.tick(_np(block[k]), want_features=False)
            assert np.array_equal(_np(out["expert"][k]), ref["expert"]), (rep, k)
            assert np.array_equal(_np(out["done"][k]), ref["done"]), (rep, k)
    orc.assert_state_equals(env)
    env.check_errors()


def test_host_state_upload_download_round_trip(splits, medium_tables, medium_oracle):
    """psk_craft_host_put_state / _get_state: a caller moves between the state-round-trip mode and
    the resident mode without losing anything."""
    from psketch_b200.host import HostCraft
    n = 3001
    idx = np.arange(n) % 2200
    args = (splits["dev_grids"], splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx],
            splits["dev_inst_task"][idx])
    env = HostCraft(medium_tables, *args, max_timesteps=15, chunk_envs=512)
    orc = _OracleTicks(medium_oracle, *args, max_timesteps=15)
    for t in range(4):                       # states travel with every call
        env.tick()
        ref = orc.tick()
        assert np.array_equal(env.expert, ref["expert"]) and np.array_equal(env.features, ref["features"])
    env.upload()                             # host state -> device, resident from here on
    with pytest.raises(Exception, match="resident"):
        env.tick()
    for t in range(5):
        env.tick_resident(features="u8")
        ref = orc.tick()
        assert np.array_equal(env.expert, ref["expert"])
        assert np.array_equal(env.features_u8.astype(np.float32), ref["features"])
    env.download(keep_resident=False)        # back to host-side states
    assert np.array_equal(env.grid[:, :64], orc.grid)
    for t in range(3):
        env.tick()
        ref = orc.tick()
        assert np.array_equal(env.expert, ref["expert"]) and np.array_equal(env.done, ref["done"])
    assert np.array_equal(env.agent[:, 24:26].astype(np.int32), orc.pos)
    env.close()


@pytest.mark.parametrize("tma", [0, 1])
def test_rollout_kernel_craft_large_vs_oracle(tma, large_tables, large_oracle, large_states):
    """craft_large (configs/worlds/craft_large.yaml: 10x10, window 5, 1,076 features, 128-bit boards)
    through psk_craft_rollout (for this geometry: chained craft_tick_kernel launches), both store paths:
    every tick's teacher action / done / success / features and the final state against the oracle."""
    from psketch_b200 import _lib
    from psketch_b200.vec import VecCraft
    n, T = 1500, 11
    rng = np.random.RandomState(4)
    task = rng.choice([13, 14, 15, 19, 20, 21, 22, 23, 24, 25, 26], size=n).astype(np.int32)
    grid, pos = large_states["grid"][:n], large_states["pos"][:n].astype(np.int32)
    try:
        _lib.set_tuning(tick_tma=tma)
        env = VecCraft.from_instances(large_tables, grid, np.arange(n), pos, task, max_timesteps=19)
        orc = _OracleTicks(large_oracle, grid, np.arange(n), pos, task, max_timesteps=19)
        ring = torch.empty((T, n, large_tables.n_features), dtype=torch.float32, device=env.device)
        for rep in range(3):
            out = env.rollout(T, features_out=ring)
            for t in range(T):
                ref = orc.tick()
                assert np.array_equal(_np(out["expert"][t]), ref["expert"]), (rep, t)
                assert np.array_equal(_np(out["done"][t]), ref["done"]), (rep, t)
                assert np.array_equal(_np(out["success"][t]), ref["success"]), (rep, t)
                assert np.array_equal(_np(ring[t]), ref["features"]), (rep, t)
        orc.assert_state_equals(env)
        env.check_errors()
    finally:
        _lib.set_tuning(tick_tma=-1)


@pytest.mark.parametrize("features", ["f32", "u8", "f32_wire_u8"])
def test_host_in_the_loop_resident_tick(features, splits, medium_tables, medium_oracle):
    """psk_craft_host_tick_resident in step-then-observe order: the host picks actions from what came
    down (here: random, or the teacher action it was handed), sends them up, and receives the features /
    teacher actions of the NEW states plus the done / success flags of the step."""
    from psketch_b200.host import HostCraft
    n = 3001
    rng = np.random.RandomState(12)
    idx = rng.randint(0, 2200, size=n)
    args = (splits["dev_grids"], splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx],
            splits["dev_inst_task"][idx])
    env = HostCraft(medium_tables, *args, max_timesteps=14, chunk_envs=1024,
                    host_threads=3 if features == "f32_wire_u8" else None)
    env.reset_resident()
    orc = _OracleTicks(medium_oracle, *args, max_timesteps=14)
    o = medium_oracle

    def check_observation(t):
        want_e, _, _ = o.expert(orc.grid, orc.inv, orc.pos, orc.dir, orc.task)
        want_f = o.features(orc.grid, orc.inv, orc.pos, orc.dir)
        assert np.array_equal(env.expert.astype(np.int32), want_e), t
        got_f = env.features if features != "u8" else env.features_u8.astype(np.float32)
        assert np.array_equal(got_f, want_f), t

    env.tick_resident(features=features, advance_first=True)      # first observation, no step
    check_observation(-1)
    assert not env.done.any()
    for t in range(35):
        a = env.expert.copy() if t % 2 else rng.choice(6, size=n, p=[.19, .19, .19, .19, .2, .04]).astype(np.uint8)
        env.tick_resident(actions=a, features=features, advance_first=True)
        ref = orc.tick(a, want_features=False)
        assert np.array_equal(env.done, ref["done"]) and np.array_equal(env.success, ref["success"]), t
        check_observation(t)
    assert tuple(int(x) for x in env.stats[:3]) == tuple(orc.stats)
    if features == "f32_wire_u8":
        assert env.lib.psk_craft_host_threads(env.ctx) == 3
    env.close()


@pytest.mark.parametrize("which,n", [("medium", 6007), ("medium", 65), ("large", 1500), ("stress16", 700)])
def test_u8_frame_through_the_fused_kernels(which, n, splits, medium_tables, medium_oracle, large_tables,
                                            large_oracle, large_states):
    """psk_craft_tick_u8 / psk_craft_rollout_u8: the feature rows leave the fused kernels as bytes
    (every feature is an exact integer <= 255).  Both tick orders and the multi-tick rollout against
    the oracle; 16x16 goes through the launch-sequence fallback."""
    from psketch_b200.vec import VecCraft
    rng = np.random.RandomState(n)
    if which == "medium":
        idx = rng.randint(0, 2200, size=n)
        tables, oracle, grids = medium_tables, medium_oracle, splits["dev_grids"]
        ienv, ipos, itask = splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx], splits["dev_inst_task"][idx]
    elif which == "large":
        tables, oracle, grids = large_tables, large_oracle, large_states["grid"][:n]
        ienv, ipos = np.arange(n), large_states["pos"][:n]
        itask = rng.choice([13, 14, 15, 19, 20, 21, 22, 23, 24, 25, 26], size=n)
    else:
        from oracle.craft_oracle import CraftOracle
        from psketch_b200.tables import CraftTables
        tables = CraftTables(world_config=dict(WIDTH=16, HEIGHT=16, WINDOW_WIDTH=3, WINDOW_HEIGHT=3,
                                               N_WORKSHOPS=3, N_PRIMITIVES=4))
        oracle = CraftOracle(tables)
        g = np.zeros((n, 16, 16), np.uint8)
        g[:, 0, :] = g[:, 15, :] = g[:, :, 0] = g[:, :, 15] = 1
        ipos = np.zeros((n, 2), np.int64)
        for i in range(n):
            cells = rng.permutation(14 * 14)
            for j, kind in enumerate([7, 7, 8, 8, 9, 9, 2, 3, 4, 5, 6, 6]):
                g[i, 1 + cells[j] // 14, 1 + cells[j] % 14] = kind
            ipos[i] = (1 + cells[20] // 14, 1 + cells[20] % 14)
        grids, ienv = g.reshape(n, 256), np.arange(n)
        itask = rng.choice([13, 14, 15, 19, 20, 21, 22, 23, 24, 25, 26], size=n)
    nf = tables.n_features
    env = VecCraft.from_instances(tables, grids, ienv, ipos, itask, max_timesteps=17)
    orc = _OracleTicks(oracle, grids, ienv, ipos, itask, max_timesteps=17)
    f8 = torch.empty((n, nf), dtype=torch.uint8, device=env.device)
    for t in range(12):                                   # observe, then step (teacher or random actions)
        a = rng.choice(6, size=n, p=[.19, .19, .19, .19, .2, .04]).astype(np.uint8) if t % 2 else None
        out = env.tick(actions=None if a is None else torch.from_numpy(a), features_out=f8)
        ref = orc.tick(a)
        assert np.array_equal(_np(out["expert"]), ref["expert"]) and np.array_equal(_np(out["done"]), ref["done"]), t
        assert np.array_equal(_np(f8).astype(np.float32), ref["features"]), t
    for t in range(8):                                    # step, then observe
        a = rng.choice(6, size=n, p=[.19, .19, .19, .19, .2, .04]).astype(np.uint8)
        out = env.tick(actions=torch.from_numpy(a), features_out=f8, advance_first=True)
        ref = orc.tick(a, want_features=False)
        assert np.array_equal(_np(out["done"]), ref["done"]), t
        assert np.array_equal(_np(f8).astype(np.float32), oracle.features(orc.grid, orc.inv, orc.pos, orc.dir)), t
    T = 7
    ring = torch.empty((T, n, nf), dtype=torch.uint8, device=env.device)
    out = env.rollout(T, features_out=ring)
    for t in range(T):
        ref = orc.tick()
        assert np.array_equal(_np(out["expert"][t]), ref["expert"]), t
        assert np.array_equal(_np(ring[t]).astype(np.float32), ref["features"]), t
    orc.assert_state_equals(env)
    env.check_errors()


def test_stuck_tile_chain_times_out_instead_of_hanging(splits, medium_tables, medium_oracle):
    """Safety net of the tile chaining: a ticket that is never finished (what an aborted launch leaves
    behind; injected with psk_debug_chain_skip_ticket) makes the next fused launch raise
    PSK_FLAG_CHAIN_TIMEOUT after a few seconds — it must not hang the GPU — and the chain is in step
    again afterwards: later ticks match the oracle."""
    import time
    from psketch_b200 import _lib
    from psketch_b200.vec import VecCraft
    n = 4096
    rng = np.random.RandomState(21)
    idx = rng.randint(0, 2200, size=n)
    args = (splits["dev_grids"], splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx], splits["dev_inst_task"][idx])
    env = VecCraft.from_instances(medium_tables, *args, max_timesteps=17)
    orc = _OracleTicks(medium_oracle, *args, max_timesteps=17)
    for t in range(3):
        out, ref = env.tick(), orc.tick()
        assert np.array_equal(_np(out["expert"]), ref["expert"])
    env.check_errors()
    with torch.cuda.device(env.device):
        rc = env.lib.psk_debug_chain_skip_ticket(env._state(), 1234, env._stream())
    if rc == 1:                                        # PSK_ERR_UNSUPPORTED
        pytest.skip("this batch does not chain (tuning knob off)")
    assert rc == 0
    t0 = time.time()
    out = env.tick()                                   # group of env 1234 waits for a holder that never comes
    torch.cuda.synchronize()
    waited = time.time() - t0
    assert 0.05 < waited < 60.0, waited
    with pytest.raises(_lib.PskError, match="tile chain timeout"):
        env.check_errors()
    ref = orc.tick()                                   # the predecessor it waited for never existed: results fine
    assert np.array_equal(_np(out["expert"]), ref["expert"])
    for t in range(4):                                 # back in step: no further waits, no flags
        t0 = time.time()
        out, ref = env.tick(), orc.tick()
        torch.cuda.synchronize()
        assert time.time() - t0 < 0.05
        assert np.array_equal(_np(out["expert"]), ref["expert"]) and np.array_equal(_np(out["done"]), ref["done"])
    out = env.rollout(5)
    for t in range(5):
        assert np.array_equal(_np(out["expert"][t]), orc.tick()["expert"])
    env.check_errors()
    orc.assert_state_equals(env)


@pytest.mark.parametrize("direct", [0, 1, 3, 5, 100, -1])
def test_wire_format_split_frame(direct, splits, medium_tables, medium_oracle):
    """PSK_FEATURES_F32_WIRE_U8 with the last `direct` chunks crossing PCIe as f32 (split frame): the
    host frame is the oracle's whichever chunk took which route, fixed or adaptive (-1), ragged last
    chunk, fewer chunks than `direct`; a host buffer that is not pinned falls back to bytes only."""
    from psketch_b200.host import HostCraft
    n, chunk = 4500, 1024           # 5 chunks, the last one of 404 envs
    rng = np.random.RandomState(77 + direct)
    idx = rng.randint(0, 2200, size=n)
    args = (splits["dev_grids"], splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx],
            splits["dev_inst_task"][idx])
    env = HostCraft(medium_tables, *args, max_timesteps=11, chunk_envs=chunk, host_threads=2)
    env.set_wire_direct(direct)
    assert env.wire_direct == (direct if direct >= 0 else 0)
    env.reset_resident()
    orc = _OracleTicks(medium_oracle, *args, max_timesteps=11)
    pinned = env.features
    for t in range(14):
        if t == 9:                  # a pageable frame: every chunk goes as bytes, same result
            env.features = np.empty_like(pinned)
        env.features[:] = -3.0
        env.tick_resident(features="f32_wire_u8")
        ref = orc.tick()
        assert np.array_equal(env.features, ref["features"]), t
        assert np.array_equal(env.expert, ref["expert"]), t
        assert np.array_equal(env.done, ref["done"]), t
        assert env.last_wire_direct == (0 if t >= 9 else min(max(direct, 0), 5)) or direct < 0
        assert 0 <= env.wire_direct <= (5 if direct < 0 else max(direct, 0))
    env.features = pinned
    # an empty call changes nothing, in particular not the split
    import ctypes
    ptr = lambda a: ctypes.c_void_p(a.ctypes.data)
    assert env.lib.psk_craft_host_tick_resident(env.ctx, None, ptr(env.features), 3, 0, ptr(env.expert), None,
                                                None, 0, None, None) == 0
    assert env.wire_direct >= 0
    env.features[:] = -3.0
    env.tick_resident(features="f32_wire_u8")
    assert np.array_equal(env.features, orc.tick()["features"])
    env.close()


@pytest.mark.parametrize("features", ["f32", "u8", "f32_wire_u8", None])
@pytest.mark.parametrize("n", [1, 32, 77, 2048])
def test_small_batch_zero_copy_host_tick(n, features, splits, medium_tables, medium_oracle):
    """psk_craft_host_tick_resident at the reference's batch sizes: with pinned buffers the fused kernel
    reads the actions from and writes every output into host memory itself (one launch per call).  Both
    tick orders against the oracle, every output poisoned before every call; the same calls with the route
    switched off, and with a pageable frame (falls back to copies), give the same bytes."""
    from psketch_b200.host import HostCraft
    rng = np.random.RandomState(1000 + n)
    idx = rng.randint(0, 2200, size=n)
    args = (splits["dev_grids"], splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx],
            splits["dev_inst_task"][idx])
    o = medium_oracle
    envs = [HostCraft(medium_tables, *args, max_timesteps=12, chunk_envs=256, host_threads=2) for _ in range(2)]
    envs[1].set_zerocopy_max(0)
    for env in envs:
        env.reset_resident()
    orc = _OracleTicks(o, *args, max_timesteps=12)

    def frame(env):
        return None if features is None else (env.features_u8.astype(np.float32) if features == "u8" else env.features)

    def poison(env):
        env.expert[:] = 77
        env.done[:] = 77
        env.success[:] = 77
        if features == "u8":
            env.features_u8[:] = 77
        elif features is not None:
            env.features[:] = -5.0

    pinned = envs[0].features
    for t in range(30):
        a = None if t < 8 else rng.choice(6, size=n, p=[.19, .19, .19, .19, .2, .04]).astype(np.uint8)
        if t == 20 and features in ("f32", "f32_wire_u8"):      # pageable frame: copy route, same result
            envs[0].features = np.empty_like(pinned)
        for env in envs:
            poison(env)
            env.tick_resident(actions=a, features=features)
        ref = orc.tick(a)
        for env in envs:
            assert np.array_equal(env.expert, ref["expert"]), t
            assert np.array_equal(env.done, ref["done"]) and np.array_equal(env.success, ref["success"]), t
            if features is not None:
                assert np.array_equal(frame(env), ref["features"]), t
    # step-then-observe order on the same contexts
    for t in range(12):
        a = rng.choice(6, size=n, p=[.19, .19, .19, .19, .2, .04]).astype(np.uint8)
        for env in envs:
            poison(env)
            env.tick_resident(actions=a, features=features, advance_first=True)
        ref = orc.tick(a, want_features=False)
        want_e, _, _ = o.expert(orc.grid, orc.inv, orc.pos, orc.dir, orc.task)
        for env in envs:
            assert np.array_equal(env.done, ref["done"]) and np.array_equal(env.success, ref["success"]), t
            assert np.array_equal(env.expert.astype(np.int32), want_e), t
            if features is not None:
                assert np.array_equal(frame(env), o.features(orc.grid, orc.inv, orc.pos, orc.dir)), t
    for env in envs:
        assert tuple(int(x) for x in env.stats[:3]) == tuple(orc.stats)
    envs[0].features = pinned
    for env in envs:
        env.close()


def test_wire_format_contexts_come_and_go(splits, medium_tables, medium_oracle):
    """PSK_FEATURES_F32_WIRE_U8 under churn: contexts (pinned landing zone, events, widening threads)
    created and destroyed repeatedly, thread counts 1..5, batch sizes that are not multiples of the chunk
    or of the widening block, fewer envs per call than the context holds; every frame against the oracle."""
    from psketch_b200.host import HostCraft
    rng = np.random.RandomState(31)
    o = medium_oracle
    for rep, (n, chunk, threads) in enumerate([(1, 128, 1), (77, 128, 2), (1000, 256, 3), (4097, 1024, 5),
                                               (2500, 4096, 4), (333, 128, 1), (5003, 512, 2), (640, 128, 3)] * 2):
        idx = rng.randint(0, 2200, size=n)
        args = (splits["dev_grids"], splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx],
                splits["dev_inst_task"][idx])
        env = HostCraft(medium_tables, *args, max_timesteps=9, chunk_envs=chunk, host_threads=threads)
        env.set_zerocopy_max(0)         # small batches too go through the landing zone and the threads
        env.reset_resident()
        orc = _OracleTicks(o, *args, max_timesteps=9)
        for t in range(6):
            env.features[:] = -7.0
            env.tick_resident(features="f32_wire_u8")
            ref = orc.tick()
            assert np.array_equal(env.features, ref["features"]), (rep, t)
            assert np.array_equal(env.expert, ref["expert"]), (rep, t)
        assert env.lib.psk_craft_host_threads(env.ctx) == threads
        env.close()
