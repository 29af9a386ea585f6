"""BASELINE configs[3]: the DAgger loop of configs/experiments/imitation.yaml with env, teacher and
student on the device (examples/train_dagger.py).  Short run: the imitation loss must fall and the
student must start solving tasks."""
import os
import sys
import types

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dagger_learns(tmp_path, splits, medium_oracle):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import train_dagger
    from psketch_b200 import data
    traj = str(tmp_path / "dev.traj")
    args = types.SimpleNamespace(batch=1024, iters=450, hidden=256, lr=1e-3, seed=1, log_every=25,
                                 eval_every=450, no_graph=False, save=None, traj=traj)
    log, policy, summary = train_dagger.train(args)
    # the evaluation record in the reference's .traj format (trainers/imitation.py:204-207,228-231):
    # one entry per dev instance, and replaying its actions on the oracle gives its success flag
    info = data.load_eval_info(traj)
    ids = ["instance_%d" % i for i in splits["dev_inst_id"]]
    assert sorted(info) == sorted(ids)
    assert abs(np.mean([info[i]["success"] for i in ids]) - log[-1]["dev_success"]) < 1e-6
    o = medium_oracle
    for k in range(0, 2200, 7):
        seq = info[ids[k]]["actions"]
        assert 1 <= len(seq) <= 40 and (seq[-1] == 5 or len(seq) == 40) and 5 not in seq[:-1]
        grid = splits["dev_grids"][splits["dev_inst_env"][k]][None].copy()
        inv = np.zeros((1, 21), np.int32)
        pos = splits["dev_inst_pos"][k][None].astype(np.int32)
        dirs = np.zeros(1, np.int32)
        for a in seq[:-1]:
            grid, inv, pos, dirs, _ = o.step(grid, inv, pos, dirs, np.asarray([a], np.int32))
        if len(seq) == 40 and seq[-1] != 5:
            pass                                  # timed out: the 40th action is recorded but not executed
        task = np.asarray([splits["dev_inst_task"][k]], np.int32)
        assert int(o.satisfies(grid, inv, pos, dirs, task)[0] == 1) == info[ids[k]]["success"], k
    # (same seed family as profiles/bench_runs/r2_dagger_b1024_600it.json: train success 0.70 and
    # dev success 0.50 after 400 iterations)
    assert log[-1]["loss"] < 0.75 * log[0]["loss"], (log[0], log[-1])
    assert log[-1]["train_success"] > 0.4 > log[0]["train_success"], (log[0], log[-1])
    assert log[-1]["dev_success"] > 0.3, log[-1]
    assert summary["rollout_env_steps_per_s"] > 2e5


def test_language_loop_learns():
    """examples/train_language.py — the primitive_language.yaml loop (instruct, explore, describe,
    hindsight-relabel, follow, imitate) on the device.  Short run: the teacher learns the student's six
    action ids from what it observes, the instructed model learns to follow the teacher's words, and
    the task-conditioned model starts imitating it.
    (profiles/bench_runs/r2_language_b1024.json: follows 0.14 / 0.35 after 50 / 100 iterations.)"""
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import train_language
    args = types.SimpleNamespace(batch=1024, iters=160, hidden=256, lr=1e-3, seed=123, log_every=20,
                                 eval_every=160)
    log, _, summary = train_language.train(args)
    assert summary["teacher_action_map"] == {0: "down", 1: "up", 2: "left", 3: "right", 4: "use", 5: "stop"}
    assert log[-1]["instructed_loss"] < 0.1 * log[0]["instructed_loss"], (log[0], log[-1])
    assert log[-1]["follows_reference"] > 0.15 > log[0]["follows_reference"], (log[0], log[-1])
    # following the reference actions solves the task, so success is at least the follow rate
    assert log[-1]["instructed_success"] >= log[-1]["follows_reference"]
    assert log[-1]["main_loss"] < 2.0 and 0.0 <= log[-1]["dev_success"] <= 1.0
