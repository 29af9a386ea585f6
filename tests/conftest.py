import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def splits():
    return np.load(os.path.join(GOLDEN, "craft_medium_splits.npz"))


@pytest.fixture(scope="session")
def medium_states():
    return np.load(os.path.join(GOLDEN, "craft_medium_states.npz"))


@pytest.fixture(scope="session")
def large_states():
    return np.load(os.path.join(GOLDEN, "craft_large_states.npz"))


@pytest.fixture(scope="session")
def custom_states():
    """2,000 states exported from the unmodified reference running tests/golden/custom/*.yaml."""
    return np.load(os.path.join(GOLDEN, "craft_custom_states.npz"))


@pytest.fixture(scope="session")
def custom_tables():
    from psketch_b200.tables import Cookbook, CraftTables, TaskManager
    cdir = os.path.join(GOLDEN, "custom")
    return CraftTables(Cookbook(os.path.join(cdir, "recipes.yaml")),
                       TaskManager(os.path.join(cdir, "hints.yaml")), "craft_medium")


@pytest.fixture(scope="session")
def custom_oracle(custom_tables):
    from oracle.craft_oracle import CraftOracle
    return CraftOracle(custom_tables)


@pytest.fixture(scope="session")
def trainer_rollouts():
    """Outputs of the reference's own ImitationTrainer.do_rollout driven by a scripted student."""
    return np.load(os.path.join(GOLDEN, "trainer_rollouts.npz"))


@pytest.fixture(scope="session")
def light_states():
    return np.load(os.path.join(GOLDEN, "light_states.npz"))


@pytest.fixture(scope="session")
def medium_tables():
    from psketch_b200.tables import CraftTables
    return CraftTables(world_config="craft_medium")


@pytest.fixture(scope="session")
def large_tables():
    from psketch_b200.tables import CraftTables
    return CraftTables(world_config="craft_large")


@pytest.fixture(scope="session")
def medium_oracle(medium_tables):
    from oracle.craft_oracle import CraftOracle
    return CraftOracle(medium_tables)


@pytest.fixture(scope="session")
def large_oracle(large_tables):
    from oracle.craft_oracle import CraftOracle
    return CraftOracle(large_tables)
