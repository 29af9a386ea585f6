"""Timings (and ncu targets) for the kernels outside the headline config: enlarged-grid teacher
(row-per-lane warps), craft_large, Light world, scenario sampler."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from profiles.prof_driver import graph_time  # noqa: E402
from psketch_b200.tables import CraftTables  # noqa: E402
from psketch_b200.vec import VecCraft  # noqa: E402


def stress_env(size, n):
    from test_stress_gpu import _random_states, _tables
    tables = _tables(size)
    m = min(n, 4096 if size < 64 else 512)
    grid, inv, pos, dirs, task = _random_states(tables, m, seed=size, wall_frac=0.2)
    rep = (n + m - 1) // m
    tile = lambda a: np.concatenate([a] * rep)[:n]
    return VecCraft.from_states(tables, tile(grid), tile(inv), tile(pos), tile(dirs), task=tile(task))


def main():
    time_it = "--time" in sys.argv
    n = 65536
    rows = []
    for size in (16, 32, 64):
        env = stress_env(size, n)
        act = env.expert()
        feats = env.features()
        if time_it:
            rows.append(("expert_rows %dx%d" % (size, size), graph_time(lambda: env.expert(out=act), 10), n))
            rows.append(("features %dx%d" % (size, size), graph_time(lambda: env.features(out=feats), 10), n))
            rows.append(("step %dx%d" % (size, size), graph_time(lambda: env.step(act), 10), n))
        del env
    # craft_large (10x10, window 5, 1076 features)
    S = np.load(os.path.join(ROOT, "tests", "golden", "craft_large_states.npz"))
    rep = n // len(S["grid"]) + 1
    tile = lambda a: np.concatenate([a] * rep)[:n]
    tl = CraftTables(world_config="craft_large")
    env = VecCraft.from_states(tl, tile(S["grid"]), tile(S["inv"]), tile(S["pos"]), tile(S["dir"]),
                               task=np.full(n, 24))
    act = env.expert()
    feats = env.features()
    if time_it:
        rows.append(("expert craft_large", graph_time(lambda: env.expert(out=act), 10), n))
        rows.append(("features craft_large (4304 B/env)", graph_time(lambda: env.features(out=feats), 10), n))
    del env, feats
    # Light world
    from psketch_b200.worlds.light import LightWorld, VecLight
    w = LightWorld()
    scens = [w.sample_scenario_with_goal(g) for _ in range(6) for g in ("LL", "LD", "RD", "UL", "UR", "URU", "DRU", "LLD", "RDD", "LUR")]
    v = VecLight(scens, np.arange(n) % len(scens))
    a = torch.randint(0, 5, (n,), dtype=torch.uint8, device=v.device)
    for _ in range(10):
        v.step(a)
    v.expert()
    if time_it:
        f = v.features()
        rows.append(("light step", graph_time(lambda: v.step(a), 10), n))
        rows.append(("light features", graph_time(lambda: v.features(out=f), 10), n))
        rows.append(("light expert (warp per env)", graph_time(lambda: v.expert(), 3, inner=2), n))
    # scenario sampler
    from psketch_b200 import data
    data.sample_scenarios(CraftTables(), n, 1)
    if time_it:
        import time
        t = CraftTables()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(5):
            data.sample_scenarios(t, n, 1 + i)
        torch.cuda.synchronize()
        rows.append(("sample_scenarios (wall, incl. host)", (time.perf_counter() - t0) / 5, n))
    torch.cuda.synchronize()
    for name, dt, m in rows:
        print("%-36s n=%d %9.1f us  %.3e env/s" % (name, m, dt * 1e6, m / dt))


if __name__ == "__main__":
    main()
