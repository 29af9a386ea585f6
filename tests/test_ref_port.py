"""Pins the pure-Python port (oracle/craft_ref_port.py — the CPU baseline that stands in for the
reference's Python loop on the GPU box) against the reference's golden vectors."""
import numpy as np

from oracle import craft_ref_port as port


def test_port_reproduces_dev_trajectories(splits, medium_tables):
    n = 2200
    steps, checksum, acts = port.run_instances(
        medium_tables, splits["dev_grids"], splits["dev_inst_env"], splits["dev_inst_pos"],
        splits["dev_inst_task"], 0, n, want_actions=True)
    assert steps == 21981
    ref, ref_len = splits["dev_ref_actions"], splits["dev_ref_len"]
    for i in range(n):
        assert acts[i] == ref[i, :ref_len[i]].tolist()


def test_port_matches_reference_states(medium_states, medium_tables):
    S = medium_states
    world = port.PortWorld(medium_tables)
    teacher = port.PortTeacher()
    tm = medium_tables.task_manager
    rng = np.random.RandomState(0)
    for i in rng.choice(len(S["grid"]), 600, replace=False):
        s = port.PortState(world, world.onehot(S["grid"][i]), tuple(int(v) for v in S["pos"][i]),
                           int(S["dir"][i]), S["inv"][i].astype(np.float64))
        assert np.array_equal(s.features(), S["features"][i].astype(np.float64))
        for a in range(6):
            r, s2 = s.step(a)
            assert r == 0
            assert np.array_equal(s2.grid.argmax(2) * (s2.grid.sum(2) > 0),
                                  S["step_grid"][i, a].reshape(8, 8))
            assert np.array_equal(s2.inventory, S["step_inv"][i, a])
            assert s2.pos == tuple(S["step_pos"][i, a]) and s2.dir == S["step_dir"][i, a]
        for tid in (13, 15, 20, 24, 26):
            task = tm.by_id(tid)
            sat = s.satisfies(task)
            assert (2 if sat is None else int(bool(sat))) == S["satisfies"][i, tid]
            try:
                a = teacher(task, s)
            except TypeError:
                a = 254
            assert a == S["expert"][i, tid]


def test_port_under_the_trainer_protocol(splits, medium_tables, trainer_rollouts):
    """The Python port (the CPU baseline) driven by the trainer's rollout protocol equals what the
    reference's own ImitationTrainer.do_rollout produced with the reference world and teacher."""
    from oracle import craft_ref_port as port
    from trainer_loop import check_against_fixture
    world = port.PortWorld(medium_tables)
    pteacher = port.PortTeacher()

    class Teacher(object):
        def __call__(self, task, state):
            return pteacher(task, state)

        def find_closest_resources(self, task, state):
            return pteacher.closest(task, state)

    def make_batch(inst):
        tm = medium_tables.task_manager
        return [dict(grid=world.onehot(splits["dev_grids"][splits["dev_inst_env"][i]]),
                     init_pos=tuple(int(v) for v in splits["dev_inst_pos"][i]),
                     task=tm.by_id(int(splits["dev_inst_task"][i]))) for i in inst]
    check_against_fixture(trainer_rollouts, make_batch, world, Teacher())
