#!/usr/bin/env python
"""Same-box timing of the launch-bound kernels (step, single fused tick) under their tuning knobs.

    python profiles/time_small.py [--n 65536] > out.json
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_workload, measured_peaks, time_kernel  # noqa: E402
from psketch_b200 import _lib  # noqa: E402
from psketch_b200.tables import CraftTables  # noqa: E402
from psketch_b200.vec import VecCraft  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--tick-variants", default="-1:-1,4:0,0:0,1:0,2:0,3:0,4:1")
    ap.add_argument("--step-variants", default="0,1,2")
    args = ap.parse_args()
    n = args.n
    peak, _ = measured_peaks()
    tables = CraftTables()
    wl = load_workload(n)
    env = VecCraft.from_instances(tables, wl["grids"], wl["env"], wl["pos"], wl["task"], max_timesteps=40)
    for _ in range(7):
        env.tick(want_features=False)          # a state in the middle of the episodes
    snap = env.snapshot()
    act = env.expert()
    nbuf = max(3, int(np.ceil(200e6 / (n * 1616))))
    big = [torch.empty((n, 404), dtype=torch.float32, device=env.device) for _ in range(nbuf)]
    res = {"n": n, "step": {}, "tick": {}}
    # launch floor: what a back-to-back chain of trivial kernels costs per launch inside a CUDA graph
    one = torch.zeros(32, device=env.device)
    res["launch_floor_us"] = time_kernel(lambda: one.add_(1.0), torch, inner=20, reps=50) * 1e6
    print("launch floor (32-element add_ in a graph chain): %.2f us per launch" % res["launch_floor_us"], file=sys.stderr)
    for r in range(args.rounds):
        for v in [int(x) for x in args.step_variants.split(",")]:
            _lib.set_tuning(step_variant=v)
            env.restore(snap)
            dt = time_kernel(lambda: env.step(act), torch, inner=20, reps=50)
            res["step"].setdefault(v, []).append(dt * 1e6)
        for spec in args.tick_variants.split(","):
            v, tma = (int(x) for x in spec.split(":"))
            _lib.set_tuning(tick_variant=v, tick_tma=tma)
            env.restore(snap)
            cnt, out = [0], {}

            def f():
                env.tick(features_out=big[cnt[0] % nbuf], fused=True, out=out)
                cnt[0] += 1
            dt = time_kernel(f, torch, inner=nbuf, reps=30)
            res["tick"].setdefault(spec, []).append(dt * 1e6)
    _lib.set_tuning(step_variant=-1, tick_variant=-1, tick_tma=-1)
    for k, d in res["step"].items():
        print("step variant %s: %s us  (frac %.3f)" % (k, ["%.2f" % x for x in d], 198 * n / min(d) * 1e6 / 1e9 / peak), file=sys.stderr)
    for k, d in res["tick"].items():
        print("tick variant %s: %s us  (frac %.3f)" % (k, ["%.2f" % x for x in d], 1815 * n / min(d) * 1e6 / 1e9 / peak), file=sys.stderr)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
