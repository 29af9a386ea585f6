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
