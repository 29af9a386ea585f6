"""psketch_b200 — B200-native batched Craft / Light grid worlds and the BFS demonstration teacher
of khanhptnk/psketch, behind the reference's own object API.

    from psketch_b200 import CraftTables, VecCraft          # batched API (N envs per launch)
    from psketch_b200 import worlds, teachers               # drop-in mirrors of the reference API

All arithmetic runs in ``libpsk_b200.so`` (hand-written sm_100a CUDA, C ABI in ``include/``);
there is no CPU fallback.  See DESIGN.md and INTEGRATION.md.
"""
__version__ = "0.1"

from .tables import CraftTables, Cookbook, TaskManager, Task  # noqa: F401


def __getattr__(name):
    # torch-dependent parts are imported on first use so that the tables work without a GPU stack
    if name == "VecCraft":
        from .vec import VecCraft
        return VecCraft
    if name == "HostCraft":
        from .host import HostCraft
        return HostCraft
    if name in ("worlds", "teachers", "data", "rollout", "dist", "students"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
