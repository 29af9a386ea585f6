"""Runs each kernel of the Craft path a few times at a given batch size — the command that is
timed with CUDA events (plain run) and then captured under ncu (see profiles/README.md)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_workload  # noqa: E402
from psketch_b200.tables import CraftTables  # noqa: E402
from psketch_b200.vec import VecCraft  # noqa: E402


def graph_time(fn, reps, inner=10):
    fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(inner):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e-3 / (reps * inner)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--time", action="store_true", help="graph-timed numbers instead of bare launches")
    ap.add_argument("--only", default="", help="comma-separated kernel names to run")
    args = ap.parse_args()
    n = args.n
    tables = CraftTables()
    wl = load_workload(n)
    env = VecCraft.from_instances(tables, wl["grids"], wl["env"], wl["pos"], wl["task"])
    # advance the batch a few ticks so that the states are mid-episode (realistic BFS depths)
    for _ in range(5):
        env.tick(want_features=False)
    nf = env.n_features
    ring = max(2, int(np.ceil(200e6 / (n * nf * 4))))
    feats = [torch.empty((n, nf), dtype=torch.float32, device=env.device) for _ in range(ring)]
    act = env.expert()
    out = {}
    c = [0]

    def k_features_tma():
        env.features(out=feats[c[0] % ring], impl=2); c[0] += 1

    def k_features_plain():
        env.features(out=feats[c[0] % ring], impl=1); c[0] += 1

    def k_expert():
        env.expert(out=act)

    def k_step():
        env.step(act)

    def k_tick():
        env.tick(features_out=feats[c[0] % ring], fused=True, out=out); c[0] += 1

    def k_tick_nofeat():
        env.tick(want_features=False, fused=True, out=out)

    T = 8
    rfeat = torch.empty((max(ring, 2), n, nf), dtype=torch.float32, device=env.device)
    rout = {}

    def k_rollout8():
        env.rollout(T, features_out=rfeat, out=rout)

    kernels = [("rollout8_per_tick", k_rollout8, 1815 * T),
               ("features_tma", k_features_tma, 1712), ("features_plain", k_features_plain, 1712),
               ("expert", k_expert, 97), ("step", k_step, 198), ("tick_fused", k_tick, 1815),
               ("tick_fused_nofeat", k_tick_nofeat, 199)]
    if args.only:
        kernels = [k for k in kernels if k[0] in args.only.split(",")]
    snap = env.snapshot()
    if args.time:
        # reference points: a pure write (memset) and a copy of the same number of bytes
        buf = torch.empty(n * nf, dtype=torch.float32, device=env.device)
        buf2 = torch.empty_like(buf)
        dt = graph_time(lambda: buf.zero_(), reps=10, inner=4)
        print("%-18s n=%d  %8.2f us  %7.1f GB/s written" % ("memset(features)", n, dt * 1e6, buf.numel() * 4 / dt / 1e9))
        dt = graph_time(lambda: buf2.copy_(buf), reps=10, inner=4)
        print("%-18s n=%d  %8.2f us  %7.1f GB/s read+written" % ("copy(features)", n, dt * 1e6, 2 * buf.numel() * 4 / dt / 1e9))
        del buf, buf2
        for name, fn, b in kernels:
            env.restore(snap)
            dt = graph_time(fn, reps=20, inner=max(ring, 10) if "feat" in name or name == "tick_fused" else 10)
            if name.startswith("rollout"):
                dt /= T
            print("%-18s n=%d  %8.2f us  %7.1f GB/s (algorithmic %d B/env)  %.3e env/s  frac_of_6542.7=%.3f"
                  % (name, n, dt * 1e6, (b / (T if name.startswith("rollout") else 1)) * n / dt / 1e9,
                     b // (T if name.startswith("rollout") else 1), n / dt,
                     (b / (T if name.startswith("rollout") else 1)) * n / dt / 1e9 / 6542.7))
    else:
        for name, fn, b in kernels:
            env.restore(snap)
            for _ in range(args.iters):
                fn()
    torch.cuda.synchronize()
    env.check_errors()


if __name__ == "__main__":
    main()
