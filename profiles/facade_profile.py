"""cProfile of the object facade under the trainers' loop (see facade_timing.py)."""
import cProfile
import os
import pstats
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import facade_timing as ft  # noqa: E402
import psketch_b200.teachers as teachers  # noqa: E402
import psketch_b200.worlds as worlds  # noqa: E402

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = ft.config()
world = worlds.load(cfg)
teacher = teachers.load(cfg)
splits = np.load(os.path.join(ft.ROOT, "tests", "golden", "craft_medium_splits.npz"))
tm = world.task_manager
K = world.cookbook.n_kinds
batch = []
for i in list(range(0, 17600, 7))[:bs]:
    ids = splits["train_grids"][splits["train_inst_env"][i]].reshape(8, 8)
    onehot = np.zeros((8, 8, K))
    xs, ys = np.nonzero(ids)
    onehot[xs, ys, ids[xs, ys]] = 1
    batch.append(dict(grid=onehot, init_pos=tuple(int(v) for v in splits["train_inst_pos"][i]),
                      task=tm.by_id(int(splits["train_inst_task"][i]))))
ft.rollout(world, teacher, batch)
pr = cProfile.Profile()
pr.enable()
for _ in range(max(1, 2048 // bs)):
    ft.rollout(world, teacher, batch)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
