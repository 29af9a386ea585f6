"""Batched rollout drivers: the per-env Python loops of the trainers
(trainers/imitation.py:18-101, make_data.py:146-152) as a handful of device launches per
timestep, with the bookkeeping (timer / done / success / action sequences / distances) kept in
tensors."""
import numpy as np
import torch

from .tables import STOP


def teacher_rollouts(env, max_len=64):
    """make_data.py:146-152 for every env at once: follow the teacher until it says STOP.
    One psk_craft_rollout launch of ``max_len`` ticks (the envs auto-reset after their first
    episode; only that first episode is kept).
    Returns (ref_actions u8[N, L] padded with 255, ref_len i32[N], satisfied bool[N])."""
    n = env.n
    env.reset()
    out = env.rollout(max_len)
    env.check_errors()
    expert, done, success = out["expert"], out["done"].bool(), out["success"].bool()
    ticks = torch.arange(max_len, device=env.device).unsqueeze(1)
    first = torch.where(done, ticks, torch.full_like(ticks, max_len)).min(dim=0).values      # [N]
    finished = first < max_len
    length = torch.where(finished, first + 1, torch.full_like(first, max_len)).to(torch.int32)
    idx = first.clamp(max=max_len - 1).unsqueeze(0)
    ok = finished & success.gather(0, idx)[0] & (expert.gather(0, idx)[0] == STOP)
    acts = torch.where(ticks < length.unsqueeze(0), expert, torch.full_like(expert, 255)).t().contiguous()
    env.reset()
    L = int(length.max().item())
    return acts[:, :L].cpu().numpy(), length.cpu().numpy(), ok.cpu().numpy()


def policy_rollouts(env, policy, max_timesteps=40, is_eval=True, mix=None, poll_every=8):
    """trainers/imitation.py:18-101 for every env at once.

    ``policy(features f32[N, n_features], t) -> actions (uint8 tensor [N])`` plays the student;
    when ``is_eval`` is False the teacher's action for every visited state is recorded
    (``ref_seqs``) and ``mix`` (bool tensor [N] or None) marks the envs that follow the teacher
    (behaviour cloning).  ONE kernel launch per timestep: ``psk_craft_tick`` in "step, then observe"
    order applies the previous actions (timer / done / success inside the kernel) and returns the
    features and the teacher's label of the new states; ``done.all()`` is polled on the host every
    ``poll_every`` timesteps only.
    Returns dict(action_seqs, ref_seqs, success, distances, num_steps, num_interactions) with the
    reference's meanings; distances: for failed get-tasks the teacher's path length from the final
    pose on the ORIGINAL grid (imitation.py:83-91), 0 for successful ones, -1 for other tasks."""
    from . import _lib
    n = env.n
    dev = env.device
    max_timesteps = _lib.check_max_timesteps(max_timesteps)
    env.reset()
    env.timer.fill_(max_timesteps)            # the kernel's timer is the trainer's (imitation.py:30,63)
    done = torch.zeros(n, dtype=torch.bool, device=dev)
    success = torch.zeros(n, dtype=torch.bool, device=dev)
    acts = torch.full((n, max_timesteps), 255, dtype=torch.uint8, device=dev)
    refs = torch.full((n, max_timesteps), 255, dtype=torch.uint8, device=dev)
    feats = torch.empty((n, env.n_features), dtype=torch.float32, device=dev)
    counts = torch.zeros(2, dtype=torch.int64, device=dev)      # interactions, env steps
    final_agent = env.agent.clone()                             # pose at the moment an env finished
    out = {}
    a = None
    t = 0
    while True:
        # step with the previous actions (none at t = 0), then observe
        env.tick(actions=a, features_out=feats, out=out, advance_first=True)
        if a is not None:
            ended = ~done & out["done"].bool()
            success |= ended & out["success"].bool()
            done |= ended
            counts[1] += (~done).sum()
            if t >= max_timesteps or (t % poll_every == 0 and bool(done.all())):
                break
        # envs still running: remember the pose they are in (finished envs were reset by the kernel)
        final_agent = torch.where(done.unsqueeze(1), final_agent, env.agent)
        a = policy(feats, t).to(device=dev, dtype=torch.uint8)
        if not is_eval:
            ref = out["expert"]
            refs[:, t] = torch.where(done, refs[:, t], ref)
            counts[0] += (~done).sum()
            if mix is not None:
                a = torch.where(mix & ~done, ref, a)
        acts[:, t] = torch.where(done, acts[:, t], a)
        a = torch.where(done, torch.full_like(a, STOP), a)     # finished envs idle at their start state
        t += 1
    env.check_errors()
    num_inter, num_steps = (int(v) for v in counts.tolist())
    if is_eval:
        num_steps = 0
    # distances for get-tasks that failed: closest resource from the final pose, original grid
    tm = env.tables.task_manager
    is_get = torch.from_numpy(env.tables.task_is_get.astype(np.bool_)).to(dev)[env.task.long()]
    goal_kind = torch.from_numpy(np.asarray(
        [0] + [env.tables.cookbook.index[tm.by_id(i).goal_arg] or 0 for i in range(1, len(tm.tasks))],
        np.uint8)).to(dev)[env.task.long()]
    # the pose each env was in when it finished, on the original grid
    env.reset()
    env.agent[:, 24:27] = final_agent[:, 24:27]
    _, length, _ = env.find_closest(goal_kind)
    env.reset()
    dist = torch.where(is_get, torch.where(success, torch.zeros_like(length), length),
                       torch.full_like(length, -1))
    return dict(action_seqs=acts.cpu().numpy(), ref_seqs=refs.cpu().numpy(),
                success=success.cpu().numpy(), distances=dist.cpu().numpy(),
                num_steps=num_steps, num_interactions=num_inter, timesteps=t)
