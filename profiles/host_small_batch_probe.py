"""Latency of one psk_craft_host_tick_resident call (host in the loop, f32 frame) at small batch sizes:
zero-copy route (the kernel reads / writes the pinned host buffers itself) against the copy-engine route.

    python profiles/host_small_batch_probe.py [--out FILE]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--calls", type=int, default=600)
    ap.add_argument("--small", action="store_true")
    args = ap.parse_args()
    import bench
    from psketch_b200.host import HostCraft
    from psketch_b200.tables import CraftTables
    tables = CraftTables()
    res = {}
    for n in ((32, 128, 512, 1024, 2048, 4096) if not args.small else (32, 512)):
        wl = bench.load_workload(n)
        row = {}
        for rnd in range(2):
            for name, zmax in (("zero_copy", 1 << 30), ("copies", 0)):
                env = HostCraft(tables, wl["grids"], wl["env"], wl["pos"], wl["task"], max_timesteps=40,
                                chunk_envs=max(n, 128))
                env.set_zerocopy_max(zmax)
                env.reset_resident()
                env.tick_resident(features="f32", advance_first=True)
                acts = env.expert.copy()
                for i in range(30 + args.calls):
                    if i == 30:
                        t0 = time.perf_counter()
                    env.tick_resident(actions=acts, features="f32", advance_first=True)
                    acts[:] = env.expert
                us = (time.perf_counter() - t0) / args.calls * 1e6
                row[name] = min(row.get(name, 1e9), round(us, 2))
                # the C entry point alone: arguments built once, no numpy work between the calls
                p = env._p
                cargs = (env.ctx, p(env.action), p(env.features), 1, 1, p(env.expert), p(env.done), p(env.success),
                         env.n, p(env.stats), p(env.err))
                fn = env.lib.psk_craft_host_tick_resident
                for i in range(30 + args.calls):
                    if i == 30:
                        t0 = time.perf_counter()
                    fn(*cargs)
                us = (time.perf_counter() - t0) / args.calls * 1e6
                row[name + "_c_call_only"] = min(row.get(name + "_c_call_only", 1e9), round(us, 2))
                nomail = cargs[:9] + (None, None)           # no statistics / error-flag copy
                for i in range(30 + args.calls):
                    if i == 30:
                        t0 = time.perf_counter()
                    fn(*nomail)
                us = (time.perf_counter() - t0) / args.calls * 1e6
                row[name + "_c_call_no_mail"] = min(row.get(name + "_c_call_no_mail", 1e9), round(us, 2))
                env.close()
        res[str(n)] = row
    line = json.dumps({"us_per_call": res, "calls": args.calls, "what": "host in the loop, f32 frame, best of 2"})
    print(line)
    if args.out:
        open(args.out, "w").write(line + "\n")


if __name__ == "__main__":
    main()
