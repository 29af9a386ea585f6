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


def _distances(env, success, final_agent):
    """The trainers' ``distances`` metric (trainers/imitation.py:83-91): for get-tasks that failed, the
    teacher's path length to the closest resource from the FINAL pose on the ORIGINAL grid (empty
    inventory), 0 for successful ones, -1 for other tasks.  Leaves the env at its episode starts."""
    dev = env.device
    tm = env.tables.task_manager
    env.reset()                                   # original grids, empty inventories, task ids
    task = env.task.long()
    is_get = torch.from_numpy(env.tables.task_is_get.astype(np.bool_)).to(dev)[task]
    goal_kind = torch.from_numpy(np.asarray(
        [0] + [env.tables.cookbook.index[tm.by_id(i).goal_arg] or 0 for i in range(1, len(tm.tasks))],
        np.uint8)).to(dev)[task]
    env.agent[:, 24:27] = final_agent[:, 24:27]   # x, y, dir of the state the rollout ended in
    _, length, _ = env.find_closest(goal_kind)
    env.reset()
    return torch.where(is_get, torch.where(success, torch.zeros_like(length), length),
                       torch.full_like(length, -1))


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
    dist = _distances(env, success, final_agent)
    return dict(action_seqs=acts.cpu().numpy(), ref_seqs=refs.cpu().numpy(),
                success=success.cpu().numpy(), distances=dist.cpu().numpy(),
                num_steps=num_steps, num_interactions=num_inter, timesteps=t)


def language_decode(env, policy, max_timesteps=40, record=True, poll_every=8):
    """One decoding pass of trainers/primitive_language.py:47-70 (and :93-113) for the whole batch,
    from the envs' episode starts.  Unlike the imitation trainer, EVERY action of a running env is
    executed — the terminating one too (STOP is a no-op step, the 40th action still moves,
    :56-60) — and a finished env keeps its state.  ``policy(features, t) -> u8[N]``.
    Returns dict(actions u8[N, T] padded with 255, lengths i64[N], agents u8[L + 1, N, 32] (the agent
    records before every action and after the last one; None unless ``record``), steps int)."""
    from . import _lib
    T = _lib.check_max_timesteps(max_timesteps)
    n, dev = env.n, env.device
    env.reset()
    done = torch.zeros(n, dtype=torch.bool, device=dev)
    acts = torch.full((n, T), 255, dtype=torch.uint8, device=dev)
    lengths = torch.zeros(n, dtype=torch.int64, device=dev)
    feats = torch.empty((n, env.n_features), dtype=torch.float32, device=dev)
    agents = [env.agent.clone()] if record else None
    steps = torch.zeros((), dtype=torch.int64, device=dev)
    t = 0
    while t < T:
        env.features(out=feats)
        a = policy(feats, t).to(device=dev, dtype=torch.uint8)
        live = ~done
        acts[:, t] = torch.where(live, a, acts[:, t])
        a_exec = torch.where(live, a, torch.full_like(a, STOP))
        env.step(a_exec, active=live.to(torch.uint8))
        steps += live.sum()
        lengths += live
        if record:
            agents.append(env.agent.clone())
        done |= (a == STOP) | (t == T - 1)
        t += 1
        if t % poll_every == 0 and bool(done.all()):
            break
    env.check_errors()
    return dict(actions=acts, lengths=lengths, agents=torch.stack(agents) if record else None,
                steps=int(steps.item()), timesteps=t)


def language_rollouts(env, teacher, policy, ref_actions, max_timesteps=40, is_eval=False,
                      greedy_policy=None, on_descriptions=None):
    """trainers/primitive_language.py:16-143 for the whole batch.

    Training: instructions = the teacher's words for the instances' ``ref_actions`` (u8[N, L] padded
    with 255, ``instruct_batch``); pass 1: ``policy`` (the instructed, exploring student) decodes from
    the episode starts; the teacher describes what the executed actions did (``describe_batch`` over
    the recorded agent records — tensor operations, the same words and random-stream consumption as
    the per-rollout ``describe``); ``on_descriptions(words)`` lets the student receive them; pass 2:
    ``greedy_policy`` decodes again from the SAME episode starts (:78-113).  Evaluation: one pass of
    ``policy``.  success / distances come from the states the last pass ended in (:120-133)."""
    from .teachers.primitive_language import instruct_batch
    ref = torch.as_tensor(np.asarray(ref_actions, np.uint8))
    instructions = instruct_batch(ref)                      # word ids, 0 = padding
    first = language_decode(env, policy, max_timesteps, record=not is_eval)
    out = dict(instructions=instructions, num_interactions=0 if is_eval else int((ref != 255).sum()),
               num_steps=0 if is_eval else first["steps"], descriptions=None)
    last = first
    if not is_eval:
        L = int(first["lengths"].max().item())
        out["descriptions"] = teacher.describe_batch(first["actions"][:, :L].long(), first["agents"][:L + 1],
                                                     first["lengths"], n_kinds=env.K)
        if on_descriptions is not None:
            on_descriptions(out["descriptions"])
        last = language_decode(env, greedy_policy if greedy_policy is not None else policy,
                               max_timesteps, record=False)
    success = env.satisfies() == 1
    final_agent = env.agent.clone()
    dist = _distances(env, success, final_agent)
    out.update(first_pass_actions=first["actions"].cpu().numpy(), action_seqs=last["actions"].cpu().numpy(),
               success=success.cpu().numpy(), distances=dist.cpu().numpy())
    return out


def interactive_rollouts(env, teacher, policy, max_timesteps=40, is_eval=False, on_step=None, poll_every=8):
    """trainers/interactive_primitive_language.py:16-106 for the whole batch.  Per timestep: the
    teacher's one-word instruction for EVERY env — also the finished ones, :49-51 — (the CUDA teacher's
    action; training only), ``policy(features, t, instruction_actions) -> u8[N]``, the step of every
    running env (the terminating action is executed too, :58-61), and the teacher's one-word description
    of what each executed action did (``describe_batch`` with rollouts of length 1, in env order, so the
    action map and the random stream evolve exactly as in the reference — in evaluation as well, :64-66).
    ``on_step(t, instructions, descriptions, live)`` hands the words to the student.
    Returns dict(action_seqs, instructions u8[T, N] (action index = word), descriptions i64[T, N]
    (-1 where the env was finished), success, distances, num_interactions, num_steps)."""
    from . import _lib
    T = _lib.check_max_timesteps(max_timesteps)
    n, dev = env.n, env.device
    env.reset()
    done = torch.zeros(n, dtype=torch.bool, device=dev)
    acts = torch.full((n, T), 255, dtype=torch.uint8, device=dev)
    instr = torch.full((T, n), 255, dtype=torch.uint8, device=dev)
    desc = torch.full((T, n), -1, dtype=torch.int64, device=dev)
    feats = torch.empty((n, env.n_features), dtype=torch.float32, device=dev)
    counts = torch.zeros(2, dtype=torch.int64, device=dev)
    t = 0
    while t < T:
        live = ~done
        words = None
        if not is_eval:
            words = env.expert()                       # asked for finished envs too
            instr[t] = words
            counts[0] += live.sum()
        env.features(out=feats)
        a = policy(feats, t, words).to(device=dev, dtype=torch.uint8)
        before = env.agent.clone()
        acts[:, t] = torch.where(live, a, acts[:, t])
        env.step(torch.where(live, a, torch.full_like(a, STOP)), active=live.to(torch.uint8))
        if not is_eval:
            counts[1] += live.sum()
        d = teacher.describe_batch(a.long().unsqueeze(1), torch.stack([before, env.agent]), live.long(),
                                   n_kinds=env.K)
        desc[t] = d[:, 0]
        if on_step is not None:
            on_step(t, words, desc[t], live)
        done |= (a == STOP) | (t == T - 1)
        t += 1
        if t % poll_every == 0 and bool(done.all()):
            break
    env.check_errors()
    success = env.satisfies() == 1
    final_agent = env.agent.clone()
    dist = _distances(env, success, final_agent)
    ni, ns = (int(v) for v in counts.tolist())
    return dict(action_seqs=acts.cpu().numpy(), instructions=instr[:t].cpu().numpy(),
                descriptions=desc[:t].cpu().numpy(), success=success.cpu().numpy(),
                distances=dist.cpu().numpy(), num_interactions=ni, num_steps=ns, timesteps=t)
