#!/usr/bin/env python
"""Upper bound for a decoupled rollout from the kernels that exist: what a
teacher/advance producer and a feature-writing consumer deliver when they run CONCURRENTLY on two
streams, against the fused craft_rollout_kernel on the same box.

    producer  P = psk_craft_rollout(T ticks, features_out = NULL): teacher + advance only
    consumer  F = psk_craft_features over T x n state rows (what T ticks of snapshots would be),
                  one-warp-per-4-envs CTAs, the fastest store geometry measured in round 1
    fused         psk_craft_rollout(T ticks) with features

A CUDA graph forks P and F onto two streams and joins them; replayed back to back.  P || F is the
best case of the decoupled design (no snapshot traffic, no pipeline fill); if it does not beat the
fused kernel, neither will the real thing.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_workload, measured_peaks  # noqa: E402
from psketch_b200.tables import CraftTables  # noqa: E402
from psketch_b200.vec import VecCraft  # noqa: E402


def timed(graph, ticks):
    graph.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(60):
        graph.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / (60 * ticks)


def main():
    n, T = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 8
    tables = CraftTables()
    wl = load_workload(n)
    mk = lambda m, w: VecCraft.from_instances(tables, w["grids"], w["env"], w["pos"], w["task"], max_timesteps=40)
    env = mk(n, wl)                                           # fused + producer
    wide = mk(n * T, load_workload(n * T))                    # T x n rows for the consumer
    ring = torch.empty((T + 1, n, 404), dtype=torch.float32, device=env.device)
    frames = torch.empty((2, n * T, 404), dtype=torch.float32, device=env.device)   # alternate: > L2 apart
    out, outp = {}, {}
    env.rollout(T, features_out=ring, out=out)
    env.rollout(T, out=outp)
    wide.features(out=frames[0])
    torch.cuda.synchronize()
    main_s, side = torch.cuda.Stream(), torch.cuda.Stream()
    res = {"n": n, "ticks": T}

    def capture(body):
        g = torch.cuda.CUDAGraph()
        main_s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(main_s):
            with torch.cuda.graph(g, stream=main_s):
                body()
        torch.cuda.current_stream().wait_stream(main_s)
        return g

    def fused():
        for _ in range(4):
            env.rollout(T, features_out=ring, out=out)

    def producer():
        for _ in range(4):
            env.rollout(T, out=outp)

    def consumer():
        for k in range(4):
            wide.features(out=frames[k % 2])

    def both():
        for k in range(4):
            side.wait_stream(main_s)
            with torch.cuda.stream(side):
                wide.features(out=frames[k % 2])
            env.rollout(T, out=outp)
            main_s.wait_stream(side)

    for name, body in (("fused", fused), ("producer_alone", producer), ("consumer_alone", consumer),
                       ("producer_and_consumer_concurrent", both)):
        g = capture(body)
        res[name + "_us_per_tick"] = float(np.median([timed(g, 4 * T) for _ in range(3)]))
    peak, _ = measured_peaks()
    res["fused_frac"] = (1619 + 196 / T) * n / res["fused_us_per_tick"] * 1e6 / 1e9 / peak
    print(json.dumps(res))


if __name__ == "__main__":
    main()
