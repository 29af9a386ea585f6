#!/usr/bin/env python
"""DAgger / behaviour-cloning training of an LSTM student with the Craft env AND the teacher on
the GPU (BASELINE configs[3]; the reference's configs/experiments/imitation.yaml loop,
trainers/imitation.py:103-180, at batch sizes the Python loop cannot reach).

    python examples/train_dagger.py --envs 4096 --iters 300

Streaming DAgger: every env runs episode after episode (auto-reset inside the tick kernel); each
iteration unrolls `--horizon` timesteps: features (device tensor) -> student logits -> sampled
action -> tick(actions) which returns the teacher's label for the same state; the loss is the
cross-entropy against the teacher over all on-policy states of the window."""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from psketch_b200.rollout import policy_rollouts  # noqa: E402
from psketch_b200.students import LSTMStudent, task_tokens  # noqa: E402
from psketch_b200.tables import CraftTables  # noqa: E402
from psketch_b200.vec import VecCraft  # noqa: E402


def load_split(name):
    sp = np.load(os.path.join(ROOT, "tests", "golden", "craft_medium_splits.npz"))
    return {k[len(name) + 1:]: sp[k] for k in sp.files if k.startswith(name + "_")}


def make_env(tables, split, n, seed, device):
    rng = np.random.RandomState(seed)
    idx = rng.randint(0, len(split["inst_env"]), size=n) if n else np.arange(len(split["inst_env"]))
    return VecCraft.from_instances(tables, split["grids"], split["inst_env"][idx], split["inst_pos"][idx],
                                   split["inst_task"][idx], max_timesteps=40, device=device)


@torch.no_grad()
def evaluate(model, tables, split, device, limit=None):
    env = make_env(tables, split, 0, 0, device)
    tok = task_tokens(tables, env.task)
    state = {"s": model.initial_state(tok)}

    def policy(feats, t):
        logits, state["s"] = model.step(state["s"], feats)
        return logits.argmax(dim=1)
    out = policy_rollouts(env, policy, max_timesteps=40, is_eval=True)
    return float(out["success"].mean())


def train(args):
    device = torch.device("cuda:0")
    torch.manual_seed(args.seed)
    tables = CraftTables()
    train_split, dev_split = load_split("train"), load_split("dev")
    env = make_env(tables, train_split, args.envs, args.seed, device)
    model = LSTMStudent(env.n_features, 6, len(tables.task_manager.vocab) + 1).to(device)
    opt = torch.optim.Adam(model.parameters(), lr=args.lr)
    tok = task_tokens(tables, env.task)
    feats = torch.empty((env.n, env.n_features), dtype=torch.float32, device=device)
    out = {}
    reset = torch.ones(env.n, dtype=torch.bool, device=device)
    carry = None
    log = []
    t0 = time.time()
    for it in range(args.iters):
        state = model.initial_state(tok)
        if carry is not None:
            state["h"], state["c"] = carry
        loss = 0.0
        env.stats.zero_()
        for t in range(args.horizon):
            env.features(out=feats)
            logits, state = model.step(state, feats, reset=reset)
            with torch.no_grad():
                if args.bc:
                    actions = None                      # behaviour cloning: follow the teacher
                else:
                    actions = torch.distributions.Categorical(logits=logits).sample().to(torch.uint8)
            env.tick(actions=actions, want_features=False, out=out)
            loss = loss + F.cross_entropy(logits, out["expert"].long())
            reset = out["done"].bool()
        opt.zero_grad(set_to_none=True)
        (loss / args.horizon).backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
        opt.step()
        carry = (state["h"].detach(), state["c"].detach())
        st = env.stats.cpu().numpy()
        rec = dict(iter=it, loss=float(loss.detach()) / args.horizon, episodes=int(st[0]),
                   train_success=float(st[1]) / max(1, int(st[0])))
        if (it + 1) % args.eval_every == 0 or it == args.iters - 1:
            rec["dev_success"] = evaluate(model, tables, dev_split, device)
        log.append(rec)
        if (it + 1) % args.log_every == 0 or "dev_success" in rec:
            el = time.time() - t0
            print("iter %4d loss %.4f train_success %.3f%s  (%.0f env-steps/s incl. student fwd/bwd)"
                  % (it, rec["loss"], rec["train_success"],
                     " dev_success %.3f" % rec["dev_success"] if "dev_success" in rec else "",
                     (it + 1) * args.horizon * env.n / el), flush=True)
    env.check_errors()
    return log, model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--horizon", type=int, default=20)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=123)
    ap.add_argument("--bc", action="store_true", help="behaviour cloning instead of DAgger")
    ap.add_argument("--log-every", type=int, default=20)
    ap.add_argument("--eval-every", type=int, default=100)
    args = ap.parse_args()
    train(args)


if __name__ == "__main__":
    main()
