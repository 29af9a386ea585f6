"""The reference-shaped object API (CraftWorld / CraftState / DemonstrationTeacher) driven by a
restatement of the trainers' rollout loop (trainers/imitation.py:18-101), checked against the
pure-Python port of the reference driven by the same loop with the same scripted student."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


class _Cfg(object):
    pass


def _config():
    c = _Cfg()
    c.recipes = None
    c.world = _Cfg(); c.world.name = "CraftWorld"; c.world.config = "craft_medium"
    c.teacher = _Cfg(); c.teacher.name = "DemonstrationTeacher"
    c.student = _Cfg(); c.student.model = _Cfg()
    c.trainer = _Cfg(); c.trainer.hints = None; c.trainer.max_timesteps = 40
    c.random = np.random.RandomState(1)
    return c


def _rollout(world, teacher, batch, student_actions, is_eval, max_timesteps, stop_index):
    """trainers/imitation.py:18-101 with the student replaced by a script."""
    states = [world.init_state(item["grid"], item["init_pos"]) for item in batch]
    tasks = [item["task"] for item in batch]
    n = len(batch)
    timer = [max_timesteps] * n
    done = [False] * n
    success = [False] * n
    action_seqs = [[] for _ in range(n)]
    ref_seqs = [[] for _ in range(n)]
    feats = []
    t = 0
    while not all(done):
        feats.append(np.stack([s.features() for s in states]))       # students/imitation.py:72
        actions = list(student_actions[t])
        for i in range(n):
            if not is_eval:
                ref = -1 if done[i] else teacher(tasks[i], states[i])
                ref_seqs[i].append(ref)
                if t % 3 == 0:                                        # behaviour cloning mix
                    actions[i] = ref if ref >= 0 else actions[i]
            if not done[i]:
                action_seqs[i].append(actions[i])
            timer[i] -= 1
            done[i] |= actions[i] == stop_index or timer[i] <= 0
            if done[i]:
                success[i] = states[i].satisfies(tasks[i])
            else:
                _, states[i] = states[i].step(actions[i])
        t += 1
    distances = []
    for i in range(n):
        if tasks[i].goal_name == "get":
            if not success[i]:
                st = world.init_state(batch[i]["grid"], states[i].pos, states[i].dir)
                _, seq = teacher.find_closest_resources(tasks[i], st)
                distances.append(len(seq))
            else:
                distances.append(0)
    final = [(s.pos, s.dir, np.asarray(s.inventory).tolist()) for s in states]
    return dict(action_seqs=action_seqs, ref_seqs=ref_seqs, success=success, feats=feats,
                distances=distances, final=final)


def test_rollout_loop_matches_reference_port(splits, medium_tables):
    import psketch_b200.worlds as worlds
    import psketch_b200.teachers as teachers
    from oracle import craft_ref_port as port

    cfg = _config()
    world = worlds.load(cfg)
    teacher = teachers.load(cfg)
    assert cfg.student.model.input_size == 404 and cfg.student.model.n_actions == 6
    assert world.actions.STOP.index == 5 and world.action_space[2].coord_change == (-1, 0)
    pworld = port.PortWorld(medium_tables)
    pteacher = port.PortTeacher()

    rng = np.random.RandomState(11)
    idx = rng.choice(2200, size=32, replace=False)
    tm = world.task_manager
    batch, pbatch = [], []
    for i in idx:
        ids = splits["dev_grids"][splits["dev_inst_env"][i]]
        onehot = pworld.onehot(ids)
        task = tm.by_id(int(splits["dev_inst_task"][i]))
        pos = tuple(int(v) for v in splits["dev_inst_pos"][i])
        batch.append(dict(grid=onehot, init_pos=pos, task=task))
        pbatch.append(dict(grid=onehot, init_pos=pos, task=medium_tables.task_manager.by_id(task.task_id)))
    script = rng.choice(6, size=(64, 32), p=[.2, .2, .2, .2, .17, .03]).tolist()
    for is_eval in (False, True):
        got = _rollout(world, teacher, batch, script, is_eval, 40, 5)

        class _PT(object):                      # port teacher with the same two entry points
            def __call__(self, task, state):
                return pteacher(task, state)

            def find_closest_resources(self, task, state):
                return pteacher.closest(task, state)
        want = _rollout(pworld, _PT(), pbatch, script, is_eval, 40, 5)
        assert got["action_seqs"] == want["action_seqs"]
        assert got["ref_seqs"] == want["ref_seqs"]
        assert [bool(a) for a in got["success"]] == [bool(a) for a in want["success"]]
        assert got["distances"] == want["distances"]
        assert got["final"] == want["final"]
        assert len(got["feats"]) == len(want["feats"])
        for a, b in zip(got["feats"], want["feats"]):
            assert a.dtype == np.float64 and np.array_equal(a, b)


def test_state_api_details(splits, medium_tables):
    import psketch_b200.worlds as worlds
    import psketch_b200.teachers as teachers
    cfg = _config()
    world = worlds.load(cfg)
    teacher = teachers.load(cfg)
    ids = splits["dev_grids"][0]
    s0 = world.init_state(ids.reshape(8, 8), (int(splits["dev_inst_pos"][0][0]), int(splits["dev_inst_pos"][0][1])))
    task = world.task_manager["make[bed]"]
    # persistence: stepping never mutates the receiver; old states stay usable
    r, s1 = s0.step(teacher(task, s0))
    assert r == 0 and s1 is not s0
    f0 = s0.features().copy()
    _, s2 = s1.step(4)
    assert np.array_equal(s0.features(), f0) and s0.features() is s0.features()
    assert s0.grid.shape == (8, 8, 21) and s0.grid.sum() == 37
    assert s2.inventory.shape == (21,)
    with pytest.raises(Exception, match="Unexpected action"):
        s0.step(6)
    with pytest.raises(AssertionError):
        teacher(world.task_manager["left[none]"], s0)
    assert s0.satisfies(world.task_manager["use[none]"]) is None
    # teacher rollout to completion reproduces the golden sequence
    ref = splits["dev_ref_actions"][0]
    task = world.task_manager.by_id(int(splits["dev_inst_task"][0]))
    s, acts = s0, []
    while True:
        a = teacher(task, s)
        acts.append(a)
        if a == 5:
            break
        _, s = s.step(a)
    assert acts == ref[:len(acts)].tolist() and s.satisfies(task)
    with pytest.raises(Exception, match="No such world"):
        cfg.world.name = "NoWorld"
        worlds.load(cfg)


def test_facade_under_the_reference_trainer_protocol(splits, trainer_rollouts):
    """The CUDA-backed CraftWorld / DemonstrationTeacher driven by the trainer's rollout protocol
    (training mode with a behaviour-cloning mix, and evaluation mode): action sequences, teacher
    labels, success, distances, counters and every feature vector the student saw are those the
    reference's own ImitationTrainer.do_rollout produced with the reference world and teacher
    (tests/golden/trainer_rollouts.npz, oracle/gen_golden.py --trainer)."""
    import psketch_b200.teachers as teachers
    import psketch_b200.worlds as worlds
    from trainer_loop import check_against_fixture
    cfg = _config()
    world = worlds.load(cfg)
    teacher = teachers.load(cfg)
    K = world.cookbook.n_kinds

    def make_batch(inst):
        out = []
        for i in inst:
            ids = splits["dev_grids"][splits["dev_inst_env"][i]].reshape(8, 8)
            onehot = np.zeros((8, 8, K))
            xs, ys = np.nonzero(ids)
            onehot[xs, ys, ids[xs, ys]] = 1
            out.append(dict(grid=onehot, init_pos=tuple(int(v) for v in splits["dev_inst_pos"][i]),
                            task=world.task_manager.by_id(int(splits["dev_inst_task"][i]))))
        return out
    check_against_fixture(trainer_rollouts, make_batch, world, teacher)


def test_world_sample_scenario():
    """CraftWorld.sample_scenario — make_data.sample_scenario(world, ...) of the reference: every
    scenario has the boundary ring, 2 of each primitive, the 3 workshops, a free start cell, and all
    free cells are connected; consecutive calls give different layouts; states start from them."""
    from psketch_b200.worlds import CraftWorld
    world = CraftWorld()
    cb = world.cookbook
    seen = set()
    for k in range(300):                               # crosses a pool refill (256 per launch)
        grid, pos = world.sample_scenario()
        assert grid.shape == (8, 8, cb.n_kinds) and (grid.sum(axis=2) <= 1).all()
        ids = grid.argmax(axis=2) * (grid.sum(axis=2) > 0)
        assert (ids[0, :] == 1).all() and (ids[7, :] == 1).all() and (ids[:, 0] == 1).all() and (ids[:, 7] == 1).all()
        counts = np.bincount(ids.ravel(), minlength=cb.n_kinds)
        assert counts[1] == 28 and counts[2] == counts[3] == counts[4] == 1
        assert counts[cb.index["iron"]] == counts[cb.index["grass"]] == counts[cb.index["wood"]] == 2
        assert counts.sum() - counts[0] == 37 and ids[pos] == 0
        free = ids == 0                                # flood fill from the start cell
        reach = np.zeros_like(free)
        reach[pos] = True
        for _ in range(40):
            grow = reach.copy()
            grow[1:, :] |= reach[:-1, :]; grow[:-1, :] |= reach[1:, :]
            grow[:, 1:] |= reach[:, :-1]; grow[:, :-1] |= reach[:, 1:]
            reach = grow & free
        assert (reach == free).all()
        seen.add(ids.tobytes())
        if k < 3:
            s = world.init_state(grid, pos)
            assert s.pos == pos and s.features().shape == (404,)
    assert len(seen) > 290
    with pytest.raises(NotImplementedError):
        world.sample_scenario(make_island=True)
