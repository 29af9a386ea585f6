"""Where a timestep of the student-in-the-loop rollout goes (torch profiler, eager launches).

    python profiles/student_loop_profile.py [--n 16384]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    args = ap.parse_args()
    import bench
    from psketch_b200.students import GraphedRollout, Seq2SeqPolicy, task_tokens
    from psketch_b200.tables import CraftTables
    from psketch_b200.vec import VecCraft
    dev = torch.device("cuda:0")
    tables = CraftTables()
    wl = bench.load_workload(args.n)
    env = VecCraft.from_instances(tables, wl["grids"], wl["env"], wl["pos"], wl["task"], max_timesteps=255, device=dev)
    torch.manual_seed(0)
    pol = Seq2SeqPolicy(env.n_features, 6, len(tables.task_manager.vocab) + 1, tables.task_manager.vocab["<PAD>"]).to(dev)
    roll = GraphedRollout(env, pol, max_timesteps=40, greedy=False, use_graph=False)
    with torch.no_grad():
        mem = pol.encode(task_tokens(tables, env.task))
    for _ in range(2):
        roll.run(mem)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        roll.run(mem)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))


if __name__ == "__main__":
    main()
