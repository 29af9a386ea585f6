"""Throughput of the reference-shaped object API (CraftWorld / CraftState / DemonstrationTeacher)
under the trainers' rollout loop (trainers/imitation.py:18-101 shape: per timestep
``features()`` for the whole batch, then per env ``teacher(task, state)``, ``satisfies``,
``state.step(a)``), teacher-forced, for several batch sizes.  Prints one JSON line per batch size.
Compare with the CPU port's per-core rate in bench.py --impl reference."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import psketch_b200.teachers as teachers  # noqa: E402
import psketch_b200.worlds as worlds  # noqa: E402


class _Cfg(object):
    pass


def config():
    c = _Cfg()
    c.recipes = None
    c.world = _Cfg(); c.world.name = "CraftWorld"; c.world.config = "craft_medium"
    c.teacher = _Cfg(); c.teacher.name = "DemonstrationTeacher"
    c.student = _Cfg(); c.student.model = _Cfg()
    c.trainer = _Cfg(); c.trainer.hints = None; c.trainer.max_timesteps = 40
    c.random = np.random.RandomState(1)
    return c


def rollout(world, teacher, batch, max_timesteps=40, stop=5):
    states = [world.init_state(item["grid"], item["init_pos"]) for item in batch]
    tasks = [item["task"] for item in batch]
    n = len(batch)
    timer = [max_timesteps] * n
    done = [False] * n
    steps = 0
    while not all(done):
        feats = np.stack([s.features() for s in states])
        for i in range(n):
            if done[i]:
                continue
            a = teacher(tasks[i], states[i])
            timer[i] -= 1
            done[i] = a == stop or timer[i] <= 0
            if done[i]:
                states[i].satisfies(tasks[i])
            else:
                _, states[i] = states[i].step(a)
                steps += 1
    return steps, feats


def main():
    splits = np.load(os.path.join(ROOT, "tests", "golden", "craft_medium_splits.npz"))
    cfg = config()
    world = worlds.load(cfg)
    teacher = teachers.load(cfg)
    tm = world.task_manager
    K = world.cookbook.n_kinds
    rng = np.random.RandomState(0)
    for bs in (32, 256, 2048):
        idx = rng.choice(len(splits["train_inst_env"]), size=bs, replace=False)
        batch = []
        for i in idx:
            ids = splits["train_grids"][splits["train_inst_env"][i]].reshape(8, 8)
            onehot = np.zeros((8, 8, K))
            xs, ys = np.nonzero(ids)
            onehot[xs, ys, ids[xs, ys]] = 1
            batch.append(dict(grid=onehot, init_pos=tuple(int(v) for v in splits["train_inst_pos"][i]),
                              task=tm.by_id(int(splits["train_inst_task"][i]))))
        rollout(world, teacher, batch)                       # warm-up
        reps = max(1, 4096 // bs)
        t0 = time.perf_counter()
        steps = 0
        for _ in range(reps):
            steps += rollout(world, teacher, batch)[0]
        dt = time.perf_counter() - t0
        print(json.dumps({"api": "CraftWorld/CraftState/DemonstrationTeacher (object facade)",
                          "batch": bs, "rollouts": reps, "env_steps": steps,
                          "env_steps_per_s": steps / dt}))


if __name__ == "__main__":
    main()
