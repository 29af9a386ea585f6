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


def test_dagger_learns():
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import train_dagger
    args = types.SimpleNamespace(batch=1024, iters=450, hidden=256, lr=1e-3, seed=1, log_every=25,
                                 eval_every=450, no_graph=False, save=None)
    log, policy, summary = train_dagger.train(args)
    # (same seed family as profiles/bench_runs/r2_dagger_b1024_600it.json: train success 0.70 and
    # dev success 0.50 after 400 iterations)
    assert log[-1]["loss"] < 0.75 * log[0]["loss"], (log[0], log[-1])
    assert log[-1]["train_success"] > 0.4 > log[0]["train_success"], (log[0], log[-1])
    assert log[-1]["dev_success"] > 0.3, log[-1]
    assert summary["rollout_env_steps_per_s"] > 2e5
