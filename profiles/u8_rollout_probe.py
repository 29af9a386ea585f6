#!/usr/bin/env python
"""Drives psk_craft_rollout_u8 (8 ticks per launch, byte frames into a 16-frame ring) for ncu:

    ncu --set full --clock-control none --import-source on -k regex:craft_rollout -s 6 -c 2 \
        -o gpurun_out/r2/u8_rollout python profiles/u8_rollout_probe.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_workload, time_kernel  # noqa: E402
from psketch_b200.tables import CraftTables  # noqa: E402
from psketch_b200.vec import VecCraft  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
tables = CraftTables()
wl = load_workload(n)
env = VecCraft.from_instances(tables, wl["grids"], wl["env"], wl["pos"], wl["task"], max_timesteps=40)
ring = torch.empty((16, n, 404), dtype=torch.uint8, device=env.device)
out = {}
for i in range(12):
    env.rollout(8, features_out=ring[(i % 2) * 8:(i % 2) * 8 + 8], out=out)
torch.cuda.synchronize()
if "--time" in sys.argv:
    c = [0]

    def f():
        env.rollout(8, features_out=ring[(c[0] % 2) * 8:(c[0] % 2) * 8 + 8], out=out)
        c[0] += 1
    dt = time_kernel(f, torch, inner=4, reps=20)
    print("u8 rollout: %.2f us per tick" % (dt * 1e6 / 8))
env.check_errors()
