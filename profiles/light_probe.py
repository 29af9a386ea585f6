#!/usr/bin/env python
"""Drives the Light-world fused tick and the multi-tick rollout at 1,048,576 envs for ncu:

    ncu --set full --clock-control none --import-source on -k regex:light_ -s 40 -c 4 \
        -o gpurun_out/r2/light python profiles/light_probe.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psketch_b200.worlds.light import LightWorld, VecLight  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
goals = ("LL", "LD", "RD", "UL", "UR", "URU", "DRU", "LLD", "RDD", "LUR")
w = LightWorld()
scens = [w.sample_scenario_with_goal(g) for rep in range(6) for g in goals]
v = VecLight(scens, np.arange(n) % len(scens))
feats = [torch.empty((n, 12), dtype=torch.float32, device=v.device) for _ in range(4)]
ring = torch.empty((16, n, 12), dtype=torch.float32, device=v.device)
out, rout = {}, {}
for i in range(40):                       # envs spread over their episodes
    v.tick(features_out=feats[i % 4], out=out, max_timesteps=100)
for i in range(4):
    v.tick(features_out=feats[i % 4], out=out, max_timesteps=100)
    v.rollout(8, features_out=ring[(i % 2) * 8:(i % 2) * 8 + 8], out=rout, max_timesteps=100)
torch.cuda.synchronize()
