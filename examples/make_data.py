#!/usr/bin/env python
"""The reference's make_data.py (make_data.py:154-238) on the GPU: samples distinct solvable Craft
scenarios, 20 start cells for every get/make task, rolls the teacher out to STOP for the
`ref_actions`, splits 80/10/10 by environment and writes `<world>_{train,dev,test}.json` in the
reference's dataset format (data/dataset.py:39-67), so that the reference's `Dataset` — or
psketch_b200.data.Dataset — reads them unchanged.

    python examples/make_data.py --out /tmp/psk_data --worlds 100          # the reference's size
    python examples/make_data.py --out /tmp/psk_data --worlds 20000        # 4.4 M instances

The random streams are Philox (device), not numpy's RandomState: the files are statistically
equivalent to the reference's, not byte-identical."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from psketch_b200 import data  # noqa: E402
from psketch_b200.tables import CraftTables  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True, help="directory for the three JSON files")
    ap.add_argument("--world", default="craft_medium", help="world config name (craft_medium, craft_large)")
    ap.add_argument("--worlds", type=int, default=None, help="number of scenarios (default: N_WORLDS of the config)")
    ap.add_argument("--positions", type=int, default=20, help="start cells per (scenario, task)")
    ap.add_argument("--seed", type=int, default=123)
    args = ap.parse_args()
    tables = CraftTables(world_config=args.world)
    t0 = time.perf_counter()
    packed = data.generate_dataset(tables, n_worlds=args.worlds, n_pos=args.positions, seed=args.seed)
    t1 = time.perf_counter()
    os.makedirs(args.out, exist_ok=True)
    for name, part in data.split_envs(packed, seed=args.seed).items():
        path = os.path.join(args.out, "%s_%s.json" % (args.world, name))
        data.save_json(path, part, tables)
        print("%s: %d scenarios, %d instances" % (path, len(part["grids"]), len(part["inst_env"])))
    print("generated in %.3f s, written in %.1f s" % (t1 - t0, time.perf_counter() - t1))


if __name__ == "__main__":
    main()
