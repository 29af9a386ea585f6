"""Enlarged-grid Craft (16x16, 32x32 and 64x64, BASELINE configs[4]): the warp-cooperative row-per-lane
teacher, features, step and the (unfused) tick against the CPU oracle, whose BFS has the
reference's 1000-slot queue cap lifted (it would overflow beyond ~250 free cells)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _tables(size):
    from psketch_b200.tables import CraftTables
    return CraftTables(world_config=dict(WIDTH=size, HEIGHT=size, WINDOW_WIDTH=3, WINDOW_HEIGHT=3,
                                         N_WORKSHOPS=3, N_PRIMITIVES=size // 4, N_WORLDS=1))


def _random_states(tables, n, seed, wall_frac):
    rng = np.random.RandomState(seed)
    W, H, K = tables.W, tables.H, tables.K
    cb = tables.cookbook
    grid = np.zeros((n, W, H), np.uint8)
    grid[:, 0, :] = grid[:, W - 1, :] = grid[:, :, 0] = grid[:, :, H - 1] = 1
    interior = (W - 2) * (H - 2)
    pos = np.zeros((n, 2), np.int32)
    for i in range(n):
        cells = rng.permutation(interior)
        k = 0

        def put(kind, count):
            nonlocal k
            for _ in range(count):
                c = cells[k]
                k += 1
                grid[i, 1 + c // (H - 2), 1 + c % (H - 2)] = kind
        for name in ("iron", "grass", "wood"):
            put(cb.index[name], rng.randint(1, W // 2))
        for w in range(3):
            put(cb.index["workshop%d" % w], 1)
        put(cb.index["water"], rng.randint(0, 4))
        put(cb.index["stone"], int(wall_frac * interior))       # obstacles: mazes and unreachable goals
        c = cells[k]
        pos[i] = (1 + c // (H - 2), 1 + c % (H - 2))
    inv = np.zeros((n, K), np.int32)
    for i in range(n):
        for _ in range(rng.randint(0, 4)):
            inv[i, rng.randint(7, K)] += 1
    dirs = rng.randint(0, 4, size=n).astype(np.int32)
    task = rng.choice([13, 14, 15, 19, 20, 21, 22, 23, 24, 25, 26], size=n).astype(np.int32)
    return grid.reshape(n, W * H), inv, pos, dirs, task


@pytest.mark.parametrize("size,n,wall", [(16, 3000, 0.15), (16, 1500, 0.35), (32, 1500, 0.2), (32, 800, 0.4),
                                         (64, 500, 0.2), (64, 300, 0.4)])
def test_enlarged_grid_parity(size, n, wall):
    from oracle.craft_oracle import CraftOracle
    from psketch_b200.vec import VecCraft
    tables = _tables(size)
    assert tables.n_features == 404
    o = CraftOracle(tables)
    grid, inv, pos, dirs, task = _random_states(tables, n, seed=size * 7 + int(wall * 100), wall_frac=wall)
    env = VecCraft.from_states(tables, grid, inv, pos, dirs, task=task)
    # features
    assert np.array_equal(env.features().cpu().numpy(), o.features(grid, inv, pos, dirs))
    assert np.array_equal(env.features(impl=2).cpu().numpy(), o.features(grid, inv, pos, dirs))
    # teacher
    act, dist = env.expert(want_dist=True)
    o_act, o_dist, _ = o.expert(grid, inv, pos, dirs, task)
    assert np.array_equal(act.cpu().numpy().astype(np.int32), o_act)
    assert np.array_equal(dist.cpu().numpy().astype(np.int32), o_dist)
    assert (o_dist > 16).sum() > 0 and (o_act == 5).sum() > 0       # long paths and unreachable goals occur
    # closest resource with the full action sequence
    kinds = np.random.RandomState(1).choice([2, 3, 4, 7, 8, 9], size=n)
    goal, length, seq = env.find_closest(torch.from_numpy(kinds.astype(np.uint8)), seq_cap=160)
    o_goal, o_len, o_st, o_seq = o.find_closest(grid, pos, dirs, kinds, seq_cap=160)
    assert np.array_equal(length.cpu().numpy().astype(np.int32), o_len)
    found = o_len >= 0
    assert np.array_equal(goal.cpu().numpy()[found].astype(np.int32), o_goal[found])
    assert np.array_equal(seq.cpu().numpy()[found], o_seq[found])
    # a few ticks of teacher-driven rollout (three-kernel pipeline on these sizes)
    state = dict(grid=grid.copy(), inv=inv.copy(), pos=pos.copy(), dir=dirs.copy(),
                 timer=np.full(n, 40, np.int32))
    env.agent[:, 28] = 40
    env.init_agent.copy_(env.agent)
    for t in range(6):
        out = env.tick(fused=True)
        state, _, feats, a = o.rollout(1, 40, grid, pos, task, state=state, want_features=True)
        # the oracle resets to dir 0 / empty inventory; episodes that end are excluded afterwards
        live = out["done"].cpu().numpy() == 0
        assert np.array_equal(out["expert"].cpu().numpy().astype(np.int32), a)
        assert np.array_equal(out["features"].cpu().numpy(), feats)
        if not live.all():
            break
    env.check_errors()
