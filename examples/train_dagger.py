#!/usr/bin/env python
"""BASELINE configs[3] at GPU batch sizes: the reference's DAgger loop
(configs/experiments/imitation.yaml, trainers/imitation.py:103-180, policy_mix.init_rate 0) with the
Craft env, the BFS teacher AND the student on the device.

    python examples/train_dagger.py --batch 1024 --iters 1500 --eval-every 250

Per iteration, exactly the reference's algorithm on a batch of training instances:
  rollout   the student acts (sampled) until STOP or 40 steps, the teacher labels every state the
            student visits — one CUDA-graph replay (psketch_b200.students.GraphedRollout);
  learn     loss = sum over timesteps of the mean cross-entropy over the envs still running
            (students/imitation.py:86-98), one Adam step — the decoder re-run over the recorded
            feature window with a single cuDNN LSTM call (Seq2SeqPolicy.decode_sequence).
Evaluation = greedy rollouts of the whole dev split (2,200 instances, one batch).  The student has
the reference's architecture (models/lstm_seq2seq.py), so `--save` writes a checkpoint the
reference's ImitationStudent.load reads (students/imitation.py:106-111).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from psketch_b200.students import GraphedRollout, Seq2SeqPolicy, imitation_loss, task_tokens  # noqa: E402
from psketch_b200.tables import CraftTables  # noqa: E402
from psketch_b200.vec import VecCraft  # noqa: E402


def load_split(name):
    sp = np.load(os.path.join(ROOT, "tests", "golden", "craft_medium_splits.npz"))
    return {k[len(name) + 1:]: sp[k] for k in sp.files if k.startswith(name + "_")}


class Batches(object):
    """data/dataset.py:69-92: shuffled passes over the split, `batch` instances at a time.  The env of
    a batch is re-pointed at the new instances by rewriting its episode-start arrays on the device."""

    def __init__(self, tables, split, batch, device, seed):
        self.split, self.batch, self.rng = split, batch, np.random.RandomState(seed)
        self.n = len(split["inst_env"])
        self.order, self.cursor = None, 0
        idx = np.arange(batch) % self.n
        self.env = VecCraft.from_instances(tables, split["grids"], split["inst_env"][idx], split["inst_pos"][idx],
                                           split["inst_task"][idx], max_timesteps=255, device=device)
        self.d_env = torch.from_numpy(split["inst_env"].astype(np.int32)).to(device)
        self.d_pos = torch.from_numpy(split["inst_pos"].astype(np.uint8)).to(device)
        self.d_task = torch.from_numpy(split["inst_task"].astype(np.uint8)).to(device)

    def next(self):
        if self.order is None or self.cursor + self.batch > len(self.order):
            passes = max(1, -(-self.batch // self.n))          # batches larger than the split: several passes
            self.order = np.concatenate([self.rng.permutation(self.n) for _ in range(passes)])
            self.cursor = 0
        rows = torch.from_numpy(self.order[self.cursor:self.cursor + self.batch]).to(self.env.device)
        self.cursor += self.batch
        self.rows = rows                                    # instance indices of this batch (device)
        env = self.env
        env.scen_idx.copy_(self.d_env[rows])
        env.init_agent[:, 24:26] = self.d_pos[rows]
        env.init_agent[:, 27] = self.d_task[rows]
        return env, self.d_task[rows]


def evaluate(policy, tables, split, device, cache={}, traj=None):
    key = id(split)
    if key not in cache:
        env = VecCraft.from_instances(tables, split["grids"], split["inst_env"], split["inst_pos"],
                                      split["inst_task"], max_timesteps=255, device=device)
        cache[key] = (env, GraphedRollout(env, policy, greedy=True, with_teacher=False))
    env, roll = cache[key]
    policy.eval()
    with torch.no_grad():
        mem = policy.encode(task_tokens(tables, env.task))
    roll.run(mem)
    policy.train()
    if traj:                             # the reference's <split>.traj (trainers/imitation.py:204-207,228-231)
        from psketch_b200 import data
        data.save_eval_info(traj, data.eval_info(split["inst_id"], roll.acts.t().cpu().numpy(),
                                                 roll.success.cpu().numpy()))
    return float(roll.success.float().mean())


def train(args):
    device = torch.device("cuda:0")
    if getattr(args, "tf32", False):        # opt-in: TF32 tensor cores for the student's GEMMs (the
        torch.backends.cuda.matmul.allow_tf32 = True   # reference's student is plain fp32)
        torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(args.seed)
    tables = CraftTables()
    train_split, dev_split = load_split("train"), load_split("dev")
    data = Batches(tables, train_split, args.batch, device, args.seed)
    vocab = len(tables.task_manager.vocab) + 1
    policy = Seq2SeqPolicy(404, 6, vocab, tables.task_manager.vocab["<PAD>"], hidden=args.hidden).to(device)
    opt = torch.optim.Adam(policy.parameters(), lr=args.lr)
    roll = GraphedRollout(data.env, policy, max_timesteps=40, greedy=False, use_graph=not args.no_graph)
    zero_time = torch.zeros((40, args.batch), dtype=torch.long, device=device)
    log, t0 = [], time.time()
    rollout_s = learn_s = 0.0
    env_steps = episodes = 0
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for it in range(args.iters):
        env, tasks = data.next()
        tokens = task_tokens(tables, tasks)
        marks[0].record()
        with torch.no_grad():
            mem = policy.encode(tokens)
        roll.run(mem)
        marks[1].record()
        mem = policy.encode(tokens)
        logits = policy.decode_sequence(roll.feats, zero_time, mem)
        loss, shown = imitation_loss(logits, roll.refs)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        marks[2].record()
        rec = dict(iter=it + 1)
        if (it + 1) % args.log_every == 0 or (it + 1) % args.eval_every == 0 or it == args.iters - 1:
            torch.cuda.synchronize()
            rec.update(loss=float(shown.detach()), train_success=float(roll.success.float().mean()),
                       steps=int(roll.steps), interactions=int(roll.interactions))
        torch.cuda.synchronize() if it < 3 else None
        if it >= 3:                                        # timings after the graph capture / warm-up
            marks[2].synchronize()
            rollout_s += marks[0].elapsed_time(marks[1]) * 1e-3
            learn_s += marks[1].elapsed_time(marks[2]) * 1e-3
            env_steps += int(roll.interactions)
            episodes += args.batch
        if (it + 1) % args.eval_every == 0 or it == args.iters - 1:
            rec["dev_success"] = evaluate(policy, tables, dev_split, device)
        if len(rec) > 1:
            log.append(rec)
            print("iter %5d  loss %.4f  train success %.3f%s  | %.2e env-steps/s in rollouts, %.2e incl. learning"
                  % (it + 1, rec["loss"], rec["train_success"],
                     "  dev success %.3f" % rec["dev_success"] if "dev_success" in rec else "",
                     env_steps / max(rollout_s, 1e-9), env_steps / max(rollout_s + learn_s, 1e-9)), flush=True)
    data.env.check_errors()
    summary = dict(batch=args.batch, iters=args.iters, episodes=episodes, env_steps=env_steps,
                   rollout_env_steps_per_s=env_steps / max(rollout_s, 1e-9),
                   train_env_steps_per_s=env_steps / max(rollout_s + learn_s, 1e-9),
                   rollout_s=rollout_s, learn_s=learn_s, wall_s=time.time() - t0,
                   final=log[-1], best_dev=max((r.get("dev_success", 0.0) for r in log), default=0.0),
                   reference="experiments/dagger_no_mix/run.log: ~1.5e3 interactions/s; train success 79.2 % at "
                             "20 k iterations of batch 32 (640 k episodes, :280); best dev 84.5 % at ~109 k (:928)")
    if args.save:
        torch.save({"model_state_dict": policy.to_reference_state_dict(), "optim_state_dict": {}}, args.save)
    if getattr(args, "traj", None):
        evaluate(policy, tables, dev_split, device, traj=args.traj)
    return log, policy, summary


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=1500)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=123)
    ap.add_argument("--log-every", type=int, default=50)
    ap.add_argument("--eval-every", type=int, default=250)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--tf32", action="store_true", help="TF32 tensor cores for the student's matmuls")
    ap.add_argument("--save", default=None, help="write a reference-format checkpoint (students/imitation.py:100-104)")
    ap.add_argument("--traj", default=None, help="write the final dev evaluation as a reference-format .traj file")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    log, _, summary = train(args)
    print(json.dumps(summary))
    if args.json:
        json.dump(dict(summary=summary, log=log), open(args.json, "w"))


if __name__ == "__main__":
    main()
