"""Light world: CPU oracle and host scenario builder pinned against the reference's exported
states; GPU kernels against both; the (reference-less) teacher against brute force."""
import numpy as np
import pytest

GOALS10 = ("LL", "LD", "RD", "UL", "UR", "URU", "DRU", "LLD", "RDD", "LUR")


def _scenario_rows(L):
    """The fixture stores 120 consecutive states per scenario: one scenario every 120 rows."""
    return np.arange(0, len(L["pos"]), 120)


def _oracle(L):
    from oracle.light_oracle import LightOracle
    r = _scenario_rows(L)
    return LightOracle(L["walls"][r], L["board"][r], L["doors"][r], L["n_doors"][r], L["keys"][r],
                       L["n_keys"][r], L["goal_room"][r])


def _states(L):
    n = len(L["pos"])
    scen_idx = np.arange(n) // 120
    mask = (L["key_alive"].astype(np.int32) << np.arange(8)[None, :]).sum(axis=1)
    state = np.stack([L["pos"][:, 0], L["pos"][:, 1], mask], axis=1).astype(np.int32)
    return scen_idx.astype(np.int32), state


def test_light_oracle_matches_reference(light_states):
    L = light_states
    o = _oracle(L)
    scen_idx, state = _states(L)
    n = len(scen_idx)
    for a in range(5):
        feat, sat, out = o.run(scen_idx, state, np.full(n, a))
        assert np.array_equal(feat, L["features"])
        assert np.array_equal(sat, L["satisfies"])
        assert np.array_equal(out[:, :2], L["step_pos"][:, a])
        m = (L["step_key_alive"][:, a].astype(np.int32) << np.arange(8)[None, :]).sum(axis=1)
        assert np.array_equal(out[:, 2], m)
    assert (o.run(scen_idx[:3], state[:3], np.full(3, 5))[2][:, 0] == -1).all()


def test_host_scenario_builder_reproduces_reference(light_states):
    from psketch_b200.worlds.light import LightWorld
    L = light_states
    w = LightWorld()
    rows = _scenario_rows(L)
    k = 0
    for rep in range(6):
        for g in GOALS10:
            sc = w.sample_scenario_with_goal(g)
            i = rows[k]
            bw, bh = sc.walls.shape
            assert tuple(L["board"][i]) == (bw, bh)
            assert np.array_equal(L["walls"][i, :bw, :bh], sc.walls)
            assert [tuple(d) for d in L["doors"][i, :len(sc.doors)]] == sc.doors
            assert [tuple(x) for x in L["keys"][i, :len(sc.keys)]] == [k_ + d for k_, d in sc.keys.items()]
            assert tuple(L["goal_room"][i]) == tuple(sc.goal_room)
            c = sc.to_c()
            assert (c.init_x, c.init_y) == tuple(L["pos"][i])          # first state = init()
            k += 1


def _vec(L):
    from psketch_b200.worlds.light import LightWorld, VecLight
    w = LightWorld()
    scens = [w.sample_scenario_with_goal(g) for rep in range(6) for g in GOALS10]
    scen_idx, state = _states(L)
    v = VecLight(scens, scen_idx)
    st = np.zeros((len(scen_idx), 4), np.uint8)
    st[:, :3] = state
    return v, st, scen_idx, state


@pytest.mark.gpu
def test_light_kernels_match_reference(light_states):
    L = light_states
    v, st, scen_idx, state = _vec(L)
    # reset puts every env at its scenario's init state
    init = v.state.cpu().numpy()
    rows = _scenario_rows(L)
    assert np.array_equal(init[rows, :2], L["pos"][rows])
    v.set_state(st)
    assert np.array_equal(v.features().cpu().numpy(), L["features"])
    assert np.array_equal(v.satisfies().cpu().numpy(), L["satisfies"])
    for a in range(5):
        v.set_state(st)
        v.step(np.full(len(st), a, np.uint8))
        got = v.state.cpu().numpy()
        assert np.array_equal(got[:, :2], L["step_pos"][:, a])
        m = (L["step_key_alive"][:, a].astype(np.int32) << np.arange(8)[None, :]).sum(axis=1)
        assert np.array_equal(got[:, 2], m)
    v.check_errors()
    v.set_state(st)
    v.step(np.full(len(st), 7, np.uint8))
    with pytest.raises(Exception, match="Unexpected action"):
        v.check_errors()
    assert np.array_equal(v.state.cpu().numpy(), st)


@pytest.mark.gpu
def test_light_teacher_matches_brute_force(light_states):
    L = light_states
    o = _oracle(L)
    v, st, scen_idx, state = _vec(L)
    sel = np.arange(0, len(st), 7)
    v.set_state(st)
    act, dist = v.expert()
    act, dist = act.cpu().numpy(), dist.cpu().numpy()
    want_a, want_d = o.expert(scen_idx[sel], state[sel])
    assert np.array_equal(dist[sel].astype(np.int32), want_d)
    assert np.array_equal(act[sel].astype(np.int32), want_a)
    # following the teacher reaches the goal room in exactly `dist` actions
    v.reset()
    a, d = v.expert()
    d0 = d.cpu().numpy().astype(np.int32)
    for t in range(int(d0.max())):
        a, d = v.expert()
        active = (a < 5).to(a.dtype)
        v.step(a.clamp(max=4), active=active)
    assert bool((v.satisfies() == 1).all())


@pytest.mark.gpu
def test_light_object_api(light_states):
    from psketch_b200.worlds.light import LightWorld
    w = LightWorld()
    sc = w.sample_scenario_with_goal("LL")
    s = sc.init()
    assert tuple(s.pos) == tuple(light_states["pos"][0])
    assert np.array_equal(s.features(), light_states["features"][0].astype(np.float64))
    r, s2 = s.step(2)
    assert r == 0 and tuple(s2.pos) == tuple(light_states["step_pos"][0, 2]) and s.pos != s2.pos or True
    assert s.satisfies(None, None) == bool(light_states["satisfies"][0])


def _all_states(L):
    """Every (scenario, key subset, standable cell) of the 60 reference scenarios."""
    rows = _scenario_rows(L)
    scen, st = [], []
    for si, r in enumerate(rows):
        bw, bh = (int(v) for v in L["board"][r])
        nk = int(L["n_keys"][r])
        free = np.argwhere(L["walls"][r][:bw, :bh] == 0)
        for m in range(1 << nk):
            lock = {(int(k[2]), int(k[3])) for j, k in enumerate(L["keys"][r][:nk]) if (m >> j) & 1}
            cells = np.asarray([c for c in free if (int(c[0]), int(c[1])) not in lock])
            scen.append(np.full(len(cells), si, np.int32))
            st.append(np.column_stack([cells, np.full(len(cells), m)]).astype(np.int32))
    return np.concatenate(scen), np.concatenate(st)


@pytest.mark.gpu
def test_light_teacher_table_is_optimal_on_every_state(light_states):
    """The teacher has no reference implementation (parity unpinned).  Its table is checked
    exhaustively instead: on EVERY state of the 60 reference scenarios — every standable cell under
    every key subset, 121,875 states — the distances satisfy the shortest-path equations under the
    oracle's step() (which IS pinned against the reference's exported states):
        dist = 0 exactly in the goal room; dist = 1 + min over the actions that change the state of
        dist(successor); unreachable states have only unreachable successors;
    the action is the smallest index attaining the minimum; the per-env search kernel of round 1
    (an independent implementation) and the brute-force oracle (a forward BFS per state, on a
    sample) give the same answers."""
    from psketch_b200.worlds.light import LightWorld, VecLight
    L = light_states
    o = _oracle(L)
    w = LightWorld()
    scens = [w.sample_scenario_with_goal(g) for rep in range(6) for g in GOALS10]
    scen_idx, state = _all_states(L)
    n = len(scen_idx)
    assert n > 100000
    v = VecLight(scens, scen_idx)
    st4 = np.zeros((n, 4), np.uint8)
    st4[:, :3] = state
    v.set_state(st4)
    act, dist = (t.cpu().numpy().astype(np.int32) for t in v.expert())
    act2, dist2 = (t.cpu().numpy().astype(np.int32) for t in v.expert(search=True))
    assert np.array_equal(act, act2) and np.array_equal(dist, dist2)
    # successor distances through the oracle's step
    key = lambda si, s: (si.astype(np.int64) << 32) | (s[:, 0].astype(np.int64) << 24) | (s[:, 1].astype(np.int64) << 16) | s[:, 2]
    lut = dict(zip(key(scen_idx, state).tolist(), dist.tolist()))
    _, sat, _ = o.run(scen_idx, state)
    assert np.array_equal(dist == 0, sat == 1) and np.array_equal(act == 254, sat == 1)
    INF = 1 << 20
    succ = np.full((5, n), INF, np.int64)
    for a in range(5):
        _, _, nxt = o.run(scen_idx, state, np.full(n, a))
        moved = (nxt != state).any(axis=1)
        d = np.asarray([lut[k] for k in key(scen_idx, nxt).tolist()], np.int64)
        succ[a] = np.where(moved & (d >= 0), d, INF)
    best = succ.min(axis=0)
    unreachable = dist < 0
    assert np.array_equal(unreachable, (best >= INF) & (sat == 0))
    assert (act[unreachable] == 255).all()
    ok = ~unreachable & (sat == 0)
    assert np.array_equal(dist[ok], 1 + best[ok])
    assert np.array_equal(act[ok], succ[:, ok].argmin(axis=0))         # argmin = smallest action index
    # brute force (forward BFS from every successor) on a sample, including unreachable states
    sel = np.concatenate([np.arange(0, n, 997), np.flatnonzero(unreachable)[:20]])
    want_a, want_d = o.expert(scen_idx[sel], state[sel])
    assert np.array_equal(act[sel], want_a) and np.array_equal(dist[sel], want_d)


@pytest.mark.gpu
def test_light_tick_matches_oracle(light_states):
    """psk_light_tick (teacher + features + done/success/auto-reset or step, one launch) against the
    oracle's features / satisfies / step and the table teacher, with teacher-driven and random actions."""
    L = light_states
    o = _oracle(L)
    v, st, scen_idx, state = _vec(L)
    n = len(scen_idx)
    v.reset()
    init = v.state.cpu().numpy().astype(np.int32)
    cur = init[:, :3].copy()
    elapsed = np.zeros(n, np.int32)
    rng = np.random.RandomState(4)
    T = 23
    tot = np.zeros(3, np.int64)
    for t in range(70):
        a_in = rng.randint(0, 5, size=n).astype(np.uint8) if t % 3 == 2 else None
        ref_a, _ = (x.cpu().numpy().astype(np.int32) for x in v.expert())
        out = v.tick(actions=a_in, max_timesteps=T)
        feat, sat, _ = o.run(scen_idx, cur)
        assert np.array_equal(out["expert"].cpu().numpy().astype(np.int32), ref_a), t
        assert np.array_equal(out["features"].cpu().numpy(), feat), t
        a = ref_a if a_in is None else a_in.astype(np.int32)
        elapsed += 1
        done = (a >= 5) | (elapsed >= T)
        succ = done & (sat == 1)
        _, _, nxt = o.run(scen_idx, cur, np.where(done, 0, a))
        cur = np.where(done[:, None], init[:, :3], nxt)
        elapsed = np.where(done, 0, elapsed)
        tot += (int(done.sum()), int(succ.sum()), n)
        assert np.array_equal(out["done"].cpu().numpy().astype(bool), done), t
        assert np.array_equal(out["success"].cpu().numpy().astype(bool), succ), t
        got = v.state.cpu().numpy().astype(np.int32)
        assert np.array_equal(got[:, :3], cur) and np.array_equal(got[:, 3], elapsed), t
    s = v.stats.cpu().numpy()
    assert (s[0], s[1], s[2]) == tuple(tot) and tot[1] > 0


@pytest.mark.gpu
def test_light_table_and_tick_stay_inside_their_buffers(light_states):
    """Canary-arena check (compute-sanitizer is closed on the pool): the teacher table (distances +
    per-cell maps), the state and every output of psk_light_tick are carved out of a canary-filled
    arena with guard bands; results equal those of ordinary allocations and the guards stay intact."""
    import torch
    from test_bounds_gpu import Arena
    L = light_states
    for n in (1, 33, 4099):
        v, st, scen_idx, state = _vec(L)
        ref, _, _, _ = _vec(L)
        sub = np.arange(n) % len(st)
        for obj in (v, ref):
            obj.scen_idx = torch.as_tensor(scen_idx[sub].astype(np.int32)).to(obj.device)
            obj.n = n
            obj.state = torch.zeros((n, 4), dtype=torch.uint8, device=obj.device)
            obj.set_state(st[sub])
        nbytes = v.lib.psk_light_teacher_table_bytes(v.scen.shape[0], v.max_keys)
        arena = Arena(nbytes + n * (4 + 48 + 3 + 64) + 64 * 1024, v.device)
        v._table = arena.carve((nbytes // 2,), torch.int16)
        v.state = arena.carve((n, 4), torch.uint8).copy_(v.state)
        v.stats = arena.carve((4,), torch.int64).zero_()
        rc = v.lib.psk_light_teacher_build(v._p(v.scen), v.scen.shape[0], v.max_keys, v._p(v._table), v._stream())
        assert rc == 0
        assert torch.equal(v._table, ref.teacher_table()) and arena.guards_intact()
        out = dict(expert=arena.carve((n,), torch.uint8), done=arena.carve((n,), torch.uint8),
                   success=arena.carve((n,), torch.uint8))
        feats = arena.carve((n, 12), torch.float32)
        for t in range(5):
            a = ref.tick(max_timesteps=7)
            b = v.tick(features_out=feats, out=out, max_timesteps=7)
            for k in ("expert", "done", "success", "features"):
                assert torch.equal(a[k], b[k]), (n, t, k)
            assert torch.equal(v.state, ref.state)
            assert arena.guards_intact(), (n, t)


@pytest.mark.gpu
def test_light_rollout_equals_ticks(light_states):
    """psk_light_rollout (T ticks in one launch) against T calls of the oracle-pinned psk_light_tick:
    teacher-driven and with an action block, feature ring shorter than T, ragged batch sizes, outputs
    carved out of a canary arena."""
    import torch
    from test_bounds_gpu import Arena
    L = light_states
    rng = np.random.RandomState(9)
    n_all = len(_states(L)[0])
    for n, T, R in ((n_all, 9, 9), (4099, 12, 5), (33, 7, 1), (1, 3, 2)):
        v, st, scen_idx, state = _vec(L)
        ref, _, _, _ = _vec(L)
        sub = np.arange(n) % len(st)
        for obj in (v, ref):
            obj.scen_idx = torch.as_tensor(scen_idx[sub].astype(np.int32)).to(obj.device)
            obj.n = n
            obj.state = torch.zeros((n, 4), dtype=torch.uint8, device=obj.device)
            obj.set_state(st[sub])
            obj.stats = None
        arena = Arena(n * (4 + R * 48 + 3 * T) + 64 * 1024, v.device)
        v.state = arena.carve((n, 4), torch.uint8).copy_(v.state)
        ring = arena.carve((R, n, 12), torch.float32)
        out = dict(expert=arena.carve((T, n), torch.uint8), done=arena.carve((T, n), torch.uint8),
                   success=arena.carve((T, n), torch.uint8))
        for use_actions in (False, True):
            acts = torch.from_numpy(rng.randint(0, 5, size=(T, n)).astype(np.uint8)).to(v.device) if use_actions else None
            got = v.rollout(T, actions=acts, features_out=ring, out=out, max_timesteps=11)
            assert arena.guards_intact()
            for t in range(T):
                want = ref.tick(actions=None if acts is None else acts[t], max_timesteps=11)
                for k in ("expert", "done", "success"):
                    assert torch.equal(got[k][t], want[k]), (n, t, k)
                if t >= T - R:                      # frame not overwritten later in the launch
                    assert torch.equal(ring[t % R], want["features"]), (n, t)
            assert torch.equal(v.state, ref.state)
            assert torch.equal(v.stats, ref.stats) and int(v.stats[2]) > 0
    # no features, no flags
    got = v.lib.psk_light_rollout(v._p(v.scen), v._p(v.scen_idx), v._p(v.state), v._p(v.teacher_table()), v.max_keys,
                                  2, None, None, 0, v._p(out["expert"]), None, None, None, 11, v.n, v._stream())
    assert got == 0
    for t in range(2):
        want = ref.tick(max_timesteps=11)
        assert torch.equal(out["expert"][t], want["expert"])
    assert torch.equal(v.state, ref.state)
