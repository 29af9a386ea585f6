"""Split frame of PSK_FEATURES_F32_WIRE_U8 (include/psk_craft.h): host-in-the-loop steps per second
with the last d chunks crossing PCIe as f32, d fixed 0..8 and adaptive, interleaved on one box.

    python profiles/wire_split_probe.py [--n 65536] [--chunk 4096] [--threads 8] [--steps 30] [--rounds 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--chunk", type=int, default=4096)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import bench
    from psketch_b200.host import HostCraft
    from psketch_b200.tables import CraftTables
    wl = bench.load_workload(args.n)
    threads = args.threads or max(1, min(8, os.cpu_count() or 2))
    env = HostCraft(CraftTables(), wl["grids"], wl["env"], wl["pos"], wl["task"], max_timesteps=40,
                    chunk_envs=args.chunk, host_threads=threads)
    env.reset_resident()
    env.tick_resident(features="f32_wire_u8", advance_first=True)
    acts = env.expert.copy()

    def step():
        env.tick_resident(actions=acts, features="f32_wire_u8", advance_first=True)
        acts[:] = env.expert

    chunks = -(-args.n // env.chunk_envs)
    res = {}
    for rnd in range(args.rounds):
        for d in list(range(0, min(chunks, 8) + 1)) + [-1]:
            env.set_wire_direct(d)
            for _ in range(12 if d < 0 else 3):
                step()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step()
            dt = (time.perf_counter() - t0) / args.steps
            key = "adaptive" if d < 0 else "d=%d" % d
            res.setdefault(key, []).append(round(dt * 1e3, 4))
            if d < 0:
                res.setdefault("adaptive_choice", []).append(env.last_wire_direct)
    out = {"n": args.n, "chunk_envs": env.chunk_envs, "chunks": chunks, "host_threads": threads,
           "cores": os.cpu_count(), "ms_per_step": res,
           "env_steps_per_s_best_of_rounds": {k: args.n / (min(v) * 1e-3) for k, v in res.items()
                                              if k != "adaptive_choice"}}
    line = json.dumps(out)
    print(line)
    if args.out:
        with open(args.out, "w") as f:
            f.write(line + "\n")
    env.close()


if __name__ == "__main__":
    main()
