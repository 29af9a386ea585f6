"""BASELINE configs[3]: DAgger training loop with the GPU env + teacher in the loop and a PyTorch
LSTM student reading the device feature tensor directly.  Short run: the imitation loss must fall
and the student must start solving dev tasks."""
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
    args = types.SimpleNamespace(envs=2048, iters=300, horizon=20, lr=2e-3, seed=1, bc=False,
                                 log_every=40, eval_every=1000)
    log, model = train_dagger.train(args)
    first = sum(r["loss"] for r in log[:5]) / 5
    last = sum(r["loss"] for r in log[-5:]) / 5
    assert last < 0.7 * first, (first, last)
    assert log[-1]["dev_success"] > 0.15, log[-1]
    assert log[-1]["train_success"] > 0.4 > log[5]["train_success"]
