#!/usr/bin/env python
"""Same-box A/B sweep of craft_rollout_kernel's variants (psk_set_tuning), interleaved rounds.

    python profiles/sweep_rollout.py [--sizes 65536,1048576] [--rounds 3] [--ticks 8] > out.json

Per (variant, store path, size): µs per tick of T-tick launches into a ring of T + 1 frames (CUDA graph
of 4 launches replayed for >= 60 ms), best and median over the rounds."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_workload, measured_peaks  # noqa: E402
from psketch_b200 import _lib  # noqa: E402
from psketch_b200.tables import CraftTables  # noqa: E402
from psketch_b200.vec import VecCraft  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="65536,131072,1048576")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--ticks", type=int, default=8)
    ap.add_argument("--variants", default="0:0,0:1,2:0,2:1,3:0,3:1,4:0,4:1")
    args = ap.parse_args()
    tables = CraftTables()
    T = args.ticks
    peak, _ = measured_peaks()
    variants = [tuple(int(x) for x in v.split(":")) for v in args.variants.split(",")]
    res = []
    for n in [int(x) for x in args.sizes.split(",")]:
        wl = load_workload(n)
        env = VecCraft.from_instances(tables, wl["grids"], wl["env"], wl["pos"], wl["task"], max_timesteps=40)
        ring = torch.empty((T + 1, n, env.n_features), dtype=torch.float32, device=env.device)
        out = {}
        graphs = {}
        for v, tma in variants:
            _lib.set_tuning(rollout_variant=v, rollout_tma=tma)
            env.rollout(T, features_out=ring, out=out)
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for _ in range(4):
                        env.rollout(T, features_out=ring, out=out)
            torch.cuda.current_stream().wait_stream(side)
            graphs[(v, tma)] = g
        times = {k: [] for k in graphs}
        for _ in range(args.rounds):
            for k, g in graphs.items():
                g.replay()
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                g.replay()
                e.record()
                torch.cuda.synchronize()
                reps = max(1, int(np.ceil(60.0 / max(s.elapsed_time(e), 1e-3))))
                s.record()
                for _ in range(reps):
                    g.replay()
                e.record()
                torch.cuda.synchronize()
                times[k].append(s.elapsed_time(e) * 1e3 / (reps * 4 * T))
        env.check_errors()
        for (v, tma), ts in times.items():
            us = float(np.median(ts))
            byt = (1619 + 196.0 / T) * n
            res.append({"n": n, "variant": v, "tma": tma, "us_per_tick_median": us, "us_per_tick_best": float(min(ts)),
                        "env_steps_per_s": n / us * 1e6, "frac": byt / us * 1e6 / 1e9 / peak})
            print("n=%8d variant=%d tma=%d  %8.2f us/tick (best %.2f)  frac %.3f" %
                  (n, v, tma, us, min(ts), res[-1]["frac"]), file=sys.stderr)
        del env, ring, graphs
        torch.cuda.empty_cache()
    _lib.set_tuning(rollout_variant=-1, rollout_tma=-1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
