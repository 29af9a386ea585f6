"""Scenario sampler (Philox) and dataset generator on the GPU: structural invariants, solvability
of every task, and cell statistics against scenarios drawn by the reference's own sampler."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _flood(free, start):
    seen = {start}
    stack = [start]
    while stack:
        x, y = stack.pop()
        for dx, dy in ((0, 1), (0, -1), (1, 0), (-1, 0)):
            q = (x + dx, y + dy)
            if q not in seen and free[q]:
                seen.add(q)
                stack.append(q)
    return seen


def test_sampled_scenarios_invariants_and_statistics(medium_tables):
    from psketch_b200 import data
    n = 20000
    g, pos, fails = data.sample_scenarios(medium_tables, n, seed=123)
    assert fails == 0
    g = g.cpu().numpy()[:, :64].reshape(n, 8, 8)
    pos = pos.cpu().numpy()
    # kind histogram: exactly the reference's 37 occupied cells
    counts = np.stack([(g == k).sum(axis=(1, 2)) for k in range(21)], axis=1)
    want = np.zeros(21, np.int64)
    want[0], want[1] = 27, 28
    want[[2, 3, 4]] = 1
    want[[7, 8, 9]] = 2
    assert (counts == want).all()
    ring = np.ones((8, 8), bool)
    ring[1:7, 1:7] = False
    assert (g[:, ring] == 1).all() and (g[:, ~ring] != 1).all()
    assert (g[np.arange(n), pos[:, 0], pos[:, 1]] == 0).all()
    # connectivity invariants of random_free(keep_connected=True) on a sample of grids
    for i in range(0, n, 97):
        free = g[i] == 0
        cells = list(zip(*np.nonzero(free)))
        assert len(_flood(free, cells[0])) == len(cells)
        for x, y in zip(*np.nonzero(g[i, 1:7, 1:7])):
            x, y = x + 1, y + 1
            assert free[x + 1, y] or free[x - 1, y] or free[x, y + 1] or free[x, y - 1]
    # determinism: a scenario depends only on (seed, index)
    g2, _, _ = data.sample_scenarios(medium_tables, 1000, seed=123, offset=500)
    assert np.array_equal(g2.cpu().numpy()[:, :64].reshape(-1, 8, 8), g[500:1500])
    g3, _, _ = data.sample_scenarios(medium_tables, 1000, seed=124)
    assert not np.array_equal(g3.cpu().numpy()[:, :64].reshape(-1, 8, 8), g[:1000])
    # cell statistics vs 3,000 scenarios from the reference's sampler
    ref = np.load(os.path.join(GOLDEN, "sampler_stats.npz"))
    m = int(ref["n_scen"])
    for k in (2, 3, 4, 7, 8, 9):
        p_ref = ref["kind_cell"][k][1:7, 1:7] / m
        p_gpu = (g == k).mean(axis=0)[1:7, 1:7]
        sigma = np.sqrt(np.maximum(p_gpu * (1 - p_gpu), 1e-4) * (1.0 / m + 1.0 / n))
        assert np.abs(p_ref - p_gpu).max() < 5 * sigma.max(), k
    p_ref = ref["pos_cell"][1:7, 1:7] / m
    p_gpu = np.zeros((8, 8))
    np.add.at(p_gpu, (pos[:, 0], pos[:, 1]), 1.0 / n)
    assert np.abs(p_ref - p_gpu[1:7, 1:7]).max() < 5 * np.sqrt(0.03 * (1.0 / m + 1.0 / n))
    # free-neighbour profile of the placed items
    nb = np.zeros(5)
    free = g == 0
    fn = np.zeros_like(g, dtype=np.int64)
    fn[:, 1:7, 1:7] = (free[:, 2:8, 1:7].astype(int) + free[:, 0:6, 1:7] + free[:, 1:7, 2:8] + free[:, 1:7, 0:6])
    items = g > 1
    for v in range(5):
        nb[v] = (fn[items] == v).sum()
    assert np.abs(nb / nb.sum() - ref["n_free_nbr"] / ref["n_free_nbr"].sum()).max() < 0.01


def test_generated_dataset_is_solvable_and_round_trips(medium_tables, tmp_path):
    from psketch_b200 import data
    packed = data.generate_dataset(medium_tables, n_worlds=40, n_pos=20, seed=7)
    n_inst = 40 * 11 * 20
    assert len(packed["inst_env"]) == n_inst
    ref, ln = packed["ref_actions"], packed["ref_len"]
    assert (ref[np.arange(n_inst), ln - 1] == 5).all()            # every trajectory ends with STOP
    assert ln.min() >= 2 and ln.max() <= 40
    # distinct grids, distinct start cells per (env, task)
    assert len({r.tobytes() for r in packed["grids"]}) == 40
    key = packed["inst_env"].astype(np.int64) * 1000 + packed["inst_task"]
    for k in np.unique(key)[:50]:
        p = packed["inst_pos"][key == k]
        assert len({tuple(q) for q in p}) == 20
    # the recorded ref_actions are what the (reference-pinned) CPU oracle's teacher does on the
    # generated scenarios, step by step, and every trajectory ends satisfied (make_data.py:146-152)
    from oracle.craft_oracle import CraftOracle
    from test_oracle import _replay
    pad = np.where(ref == 255, 5, ref).astype(np.int32)
    assert _replay(CraftOracle(medium_tables), packed["grids"], packed["inst_env"], packed["inst_task"],
                   packed["inst_pos"], pad, ln) == 0
    # wire format round trip (data/dataset.py:39-67)
    path = str(tmp_path / "craft_medium_gen.json")
    data.save_json(path, packed, medium_tables)
    back = data.load_json(path, medium_tables)
    for name in ("grids", "inst_task", "inst_pos", "ref_len"):
        assert np.array_equal(np.asarray(back[name]), np.asarray(packed[name])), name
    assert np.array_equal(back["ref_actions"], packed["ref_actions"][:, :back["ref_actions"].shape[1]])
    parts = data.split_envs(packed)
    assert [len(parts[s]["grids"]) for s in ("train", "dev", "test")] == [32, 4, 4]
    assert sum(len(parts[s]["inst_env"]) for s in parts) == n_inst


def test_policy_rollouts_match_facade_semantics(splits, medium_tables, medium_oracle):
    """Batched rollout driver with a scripted student vs the same loop on the CPU oracle."""
    from psketch_b200.rollout import policy_rollouts
    from psketch_b200.vec import VecCraft
    n = 3000
    rng = np.random.RandomState(3)
    idx = rng.randint(0, 2200, size=n)
    grids = splits["test_grids"]
    ienv, ipos, itask = (splits["test_inst_env"][idx], splits["test_inst_pos"][idx],
                         splits["test_inst_task"][idx])
    env = VecCraft.from_instances(medium_tables, grids, ienv, ipos, itask)
    script = rng.choice(6, size=(40, n), p=[.2, .2, .2, .2, .18, .02]).astype(np.uint8)
    dscript = torch.from_numpy(script).to(env.device)
    mix = torch.from_numpy(rng.rand(n) < 0.5).to(env.device)
    out = policy_rollouts(env, lambda f, t: dscript[t], max_timesteps=40, is_eval=False, mix=mix)
    # oracle replay
    o = medium_oracle
    grid = grids[ienv.astype(np.int64)].copy()
    grid0 = grid.copy()
    inv = np.zeros((n, 21), np.int32)
    pos = ipos.astype(np.int32).copy()
    dirs = np.zeros(n, np.int32)
    task = itask.astype(np.int32)
    timer = np.full(n, 40)
    done = np.zeros(n, bool)
    success = np.zeros(n, bool)
    mixn = mix.cpu().numpy()
    for t in range(out["timesteps"]):
        ref, _, _ = o.expert(grid, inv, pos, dirs, task)
        a = np.where(mixn & ~done, ref, script[t]).astype(np.int32)
        assert np.array_equal(out["ref_seqs"][~done, t].astype(np.int32), ref[~done])
        assert np.array_equal(out["action_seqs"][~done, t].astype(np.int32), a[~done])
        timer -= 1
        newly = ~done & ((a == 5) | (timer <= 0))
        success |= newly & (o.satisfies(grid, inv, pos, dirs, task) == 1)
        done |= newly
        g2, i2, p2, d2, _ = o.step(grid, inv, pos, dirs, a)
        act = ~done
        grid[act], inv[act], pos[act], dirs[act] = g2[act], i2[act], p2[act], d2[act]
    assert done.all()
    assert np.array_equal(out["success"], success)
    is_get = medium_tables.task_is_get[task].astype(bool)
    tm = medium_tables.task_manager
    kinds = np.asarray([medium_tables.cookbook.index[tm.by_id(int(t)).goal_arg] for t in task])
    _, length, _, _ = o.find_closest(grid0, pos, dirs, kinds)
    want = np.where(is_get, np.where(success, 0, length), -1)
    assert np.array_equal(out["distances"].astype(np.int64), want)


def test_describe_batch_on_device_states(splits, medium_tables):
    """PrimitiveLanguageTeacher.describe_batch on CUDA tensors recorded from VecCraft rollouts (the
    student uses private action ids) against describe() called rollout by rollout on host state
    objects — words, the learned action map and the position of the shared random stream; pinned to
    the reference's teachers/primitive_language.py:35-90 through tests/test_language_teacher.py and
    the config-4 replay (tests/test_config4.py)."""
    import types
    from psketch_b200.teachers import PrimitiveLanguageTeacher
    from psketch_b200.teachers.primitive_language import ACTION_WORDS
    from psketch_b200.vec import VecCraft
    n, L = 700, 12
    rng = np.random.RandomState(8)
    idx = rng.randint(0, 2200, size=n)
    env = VecCraft.from_instances(medium_tables, splits["dev_grids"], splits["dev_inst_env"][idx],
                                  splits["dev_inst_pos"][idx], splits["dev_inst_task"][idx])
    perm = rng.permutation(6)                      # real action a is known to the student as perm[a]
    a = PrimitiveLanguageTeacher(types.SimpleNamespace(random=np.random.RandomState(3)))
    b = PrimitiveLanguageTeacher(types.SimpleNamespace(random=np.random.RandomState(3)))
    world = types.SimpleNamespace(action_space=list(range(6)))
    for call in range(3):
        env.reset()
        real = rng.choice(6, size=(L, n), p=[.2, .2, .2, .2, .18, .02]).astype(np.uint8)
        lengths = rng.randint(1, L + 1, size=n)
        snaps = [env.agent.clone()]
        for t in range(L):
            env.step(torch.from_numpy(real[t]), active=torch.from_numpy((t < lengths).astype(np.uint8)))
            snaps.append(env.agent.clone())
        agent_seq = torch.stack(snaps)                                   # [L + 1, N, 32] on the device
        ids = torch.from_numpy(perm[real.T.astype(np.int64)]).to(env.device)   # [N, L] student ids
        got = b.describe_batch(ids, agent_seq, torch.from_numpy(lengths).to(env.device), n_kinds=21)
        assert got.is_cuda
        got = got.cpu().numpy()
        host = agent_seq.cpu().numpy()
        for i in range(n):
            states = [types.SimpleNamespace(pos=(int(host[t, i, 24]), int(host[t, i, 25])),
                                            inventory=host[t, i, :21].astype(float))
                      for t in range(lengths[i] + 1)]
            want = a.describe(world, [int(v) for v in perm[real[:lengths[i], i]]], states)
            assert [ACTION_WORDS[w] for w in got[i, :lengths[i]]] == want, (call, i)
            assert (got[i, lengths[i]:] == -1).all()
        assert a.student_action_map == b.student_action_map
    assert a.random.randint(1 << 30) == b.random.randint(1 << 30)
    assert len(b.student_action_map) == 6
    env.check_errors()
