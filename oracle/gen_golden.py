"""Generates the committed golden fixtures under tests/golden/ by running the UNMODIFIED
reference (imported from /root/reference through oracle/ref_shim.py).  TEST INFRASTRUCTURE.

    python -m oracle.gen_golden            # ~2-3 minutes, build container only

Outputs (all numpy .npz, compressed):

* craft_medium_splits.npz — the reference's own golden vectors: data/craft_medium_{dev,test}.json
  as shipped and the train split regenerated with make_data.py (seed 123, make_data.py:155):
  per split ``grids u8[E,W*H]`` (kind ids, index x*H+y), ``inst_env``, ``inst_task``
  (task id = position in the hint file, 1-based), ``inst_pos``, ``ref_actions u8[I,L]`` padded
  with 255, ``ref_len``.
* craft_medium_states.npz / craft_large_states.npz — states (on-policy, random-action and
  perturbed: water/stone cells, injected bridge/axe/other inventory, removed resources) with
  the reference's outputs: ``features`` (stored u8, exact: asserted integral and < 256),
  ``step_*`` for all six actions, ``satisfies`` for every task id (0 False, 1 True, 2 None),
  ``expert`` for every task id (0..5, 255 AssertionError, 254 TypeError), ``closest_*`` from
  find_closest_resources for every kind that has a go[...] task.
* sampler_stats.npz — cell-occupancy statistics of 3,000 scenarios drawn by the reference's own
  sampler (make_data.py:105-144), the yardstick for the Philox sampler kernel.
* light_states.npz — Light world scenarios for the 10 goals of resources/light/hints.yaml with
  random-action rollouts: walls, doors, keys, pos -> features f32[12], step results, satisfies.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ERR_ASSERT, ERR_TYPE, ERR_OTHER = 255, 254, 253


def to_ids(grid_onehot):
    g = np.asarray(grid_onehot)
    assert (g.sum(axis=2) <= 1).all()
    return g.argmax(axis=2).astype(np.uint8) * (g.sum(axis=2) > 0)


def to_onehot(ids, K):
    W, H = ids.shape
    g = np.zeros((W, H, K))
    xs, ys = np.nonzero(ids)
    g[xs, ys, ids[xs, ys]] = 1
    return g


def task_goal(name):
    return name if "[" in name else "%s[%s]" % tuple(name.split(" "))


def export_splits(R, regen_dir):
    out = {}
    tm = R.task_manager
    import json
    for split in ("train", "dev", "test"):
        shipped = os.path.join(ref_shim.REF_ROOT, "data", "craft_medium_%s.json" % split)
        regen = os.path.join(regen_dir, "craft_medium_%s.json" % split)
        data_regen = json.load(open(regen))
        if os.path.exists(shipped):
            data = json.load(open(shipped))
            # the regenerated split must reproduce the shipped one
            assert len(data) == len(data_regen)
            for a, b in zip(data, data_regen):
                assert a["grid"] == b["grid"]
                for ta, tb in zip(a["task_instances"], b["task_instances"]):
                    assert ta["init_pos"] == [list(p) for p in tb["init_pos"]]
                    assert ta["ref_actions"] == tb["ref_actions"] and ta["ids"] == tb["ids"]
        else:
            data = data_regen
        grids, ienv, itask, ipos, iid, acts = [], [], [], [], [], []
        for e, env in enumerate(data):
            grids.append(to_ids(env["grid"]).reshape(-1))
            for ti in env["task_instances"]:
                tid = tm.tasks[tm[task_goal(ti["task"])]]
                for pos, id_, ra in zip(ti["init_pos"], ti["ids"], ti["ref_actions"]):
                    ienv.append(e)
                    itask.append(tid)
                    ipos.append(pos)
                    iid.append(int(id_.split("_")[1]))
                    acts.append(ra)
        L = max(len(a) for a in acts)
        ra = np.full((len(acts), L), 255, np.uint8)
        for i, a in enumerate(acts):
            ra[i, :len(a)] = a
        out[split + "_grids"] = np.stack(grids)
        out[split + "_inst_env"] = np.asarray(ienv, np.int16)
        out[split + "_inst_task"] = np.asarray(itask, np.uint8)
        out[split + "_inst_pos"] = np.asarray(ipos, np.uint8)
        out[split + "_inst_id"] = np.asarray(iid, np.int32)
        out[split + "_ref_actions"] = ra
        out[split + "_ref_len"] = np.asarray([len(a) for a in acts], np.uint8)
        print(split, len(data), "envs", len(acts), "instances", int(out[split + "_ref_len"].sum()),
              "state/action pairs")
    np.savez_compressed(os.path.join(OUT, "craft_medium_splits.npz"), **out)
    return out


def random_layout(rng, world, n_water, n_stone, n_extra):
    """Boundary ring + the usual resources at random free cells (no connectivity constraint, so
    unreachable goals occur) + optional water/stone/crafted items lying around."""
    cb = world.cookbook
    W, H, K = world.WIDTH, world.HEIGHT, cb.n_kinds
    ids = np.zeros((W, H), np.uint8)
    ids[0, :] = ids[W - 1, :] = ids[:, 0] = ids[:, H - 1] = cb.index["boundary"]

    def put(kind):
        for _ in range(200):
            x, y = rng.randint(1, W - 1), rng.randint(1, H - 1)
            if ids[x, y] == 0:
                ids[x, y] = kind
                return

    for name in ("iron", "grass", "wood"):
        for _ in range(rng.randint(0, world.N_PRIMITIVES + 1)):
            put(cb.index[name])
    for i in range(world.N_WORKSHOPS):
        if rng.rand() < 0.9:
            put(cb.index["workshop%d" % i])
    for _ in range(n_water):
        put(cb.index["water"])
    for _ in range(n_stone):
        put(cb.index["stone"])
    for _ in range(n_extra):
        put(rng.randint(7, K))
    return ids


def collect_states(R, splits, n_on, n_off, n_pert, seed):
    """Returns a list of reference CraftState objects (+ nothing else: all outputs are computed
    later by the reference itself)."""
    rng = np.random.RandomState(seed)
    world = R.world
    K = world.cookbook.n_kinds
    W, H = world.WIDTH, world.HEIGHT
    states = []
    if splits is not None:
        grids = np.concatenate([splits["dev_grids"], splits["test_grids"]])
        n_dev = len(splits["dev_grids"])
        ienv = np.concatenate([splits["dev_inst_env"], splits["test_inst_env"] + n_dev])
        ipos = np.concatenate([splits["dev_inst_pos"], splits["test_inst_pos"]])
        ract = np.concatenate([
            np.pad(splits["dev_ref_actions"], ((0, 0), (0, 32 - splits["dev_ref_actions"].shape[1])),
                   constant_values=255),
            np.pad(splits["test_ref_actions"], ((0, 0), (0, 32 - splits["test_ref_actions"].shape[1])),
                   constant_values=255)])
        # on-policy: a random prefix of a golden trajectory
        for _ in range(n_on):
            i = rng.randint(len(ienv))
            g = to_onehot(grids[ienv[i]].reshape(W, H), K)
            s = world.init_state(g, tuple(int(v) for v in ipos[i]))
            L = int((ract[i] != 255).sum())
            for a in ract[i][:rng.randint(0, L)]:
                _, s = s.step(int(a))
            states.append(s)
        # off-policy: random actions (USE-heavy so that inventories fill up)
        for _ in range(n_off):
            i = rng.randint(len(ienv))
            g = to_onehot(grids[ienv[i]].reshape(W, H), K)
            s = world.init_state(g, tuple(int(v) for v in ipos[i]), int(rng.randint(4)))
            for _ in range(rng.randint(0, 60)):
                a = int(rng.choice(6, p=[.17, .17, .17, .17, .27, .05]))
                _, s = s.step(a)
            states.append(s)
    # perturbed layouts and inventories
    for _ in range(n_pert):
        ids = random_layout(rng, world, rng.randint(0, 4), rng.randint(0, 4), rng.randint(0, 4))
        free = np.argwhere(ids == 0)
        if len(free) == 0:
            continue
        x, y = free[rng.randint(len(free))]
        s = world.init_state(to_onehot(ids, K), (int(x), int(y)), int(rng.randint(4)))
        inv = np.zeros(K)
        for _ in range(rng.randint(0, 6)):
            inv[rng.randint(7, K)] += rng.randint(1, 4)
        if rng.rand() < 0.4:
            inv[world.cookbook.index["bridge"]] += 1
        if rng.rand() < 0.4:
            inv[world.cookbook.index["axe"]] += 1
        s.inventory = inv
        for _ in range(rng.randint(0, 12)):
            a = int(rng.choice(6, p=[.17, .17, .17, .17, .27, .05]))
            _, s = s.step(a)
        states.append(s)
    return states


def export_states(R, states, path):
    world, teacher, tm = R.world, R.teacher, R.task_manager
    K = world.cookbook.n_kinds
    W, H = world.WIDTH, world.HEIGHT
    n = len(states)
    n_tasks = len(tm.tasks)          # ids 1..n_tasks-1
    go_kinds = sorted(set(world.cookbook.index[t.goal_arg] for t in tm.tasks
                          if t.goal_name == "go"))
    SEQ = 48
    o = dict(
        W=np.int32(W), H=np.int32(H), K=np.int32(K), win=np.int32(world.WINDOW_WIDTH),
        grid=np.zeros((n, W * H), np.uint8), inv=np.zeros((n, K), np.uint8),
        pos=np.zeros((n, 2), np.uint8), dir=np.zeros(n, np.uint8),
        features=np.zeros((n, world.n_features), np.uint8),
        step_grid=np.zeros((n, 6, W * H), np.uint8), step_inv=np.zeros((n, 6, K), np.uint8),
        step_pos=np.zeros((n, 6, 2), np.uint8), step_dir=np.zeros((n, 6), np.uint8),
        step_reward=np.zeros((n, 6), np.float32),
        satisfies=np.zeros((n, n_tasks), np.uint8), expert=np.zeros((n, n_tasks), np.uint8),
        go_kinds=np.asarray(go_kinds, np.uint8),
        closest_goal=np.full((n, len(go_kinds), 2), 255, np.uint8),
        closest_len=np.full((n, len(go_kinds)), -1, np.int16),
        closest_status=np.zeros((n, len(go_kinds)), np.uint8),
        closest_seq=np.full((n, len(go_kinds), SEQ), 255, np.uint8),
    )
    go_task = {world.cookbook.index[t.goal_arg]: t for t in tm.tasks if t.goal_name == "go"}
    for i, s in enumerate(states):
        o["grid"][i] = to_ids(s.grid).reshape(-1)
        assert (s.inventory == np.round(s.inventory)).all() and s.inventory.max() < 256
        o["inv"][i] = s.inventory
        o["pos"][i] = s.pos
        o["dir"][i] = s.dir
        f = s.features()
        assert (f == np.round(f)).all() and f.min() >= 0 and f.max() < 256
        o["features"][i] = f
        for a in range(6):
            r, s2 = s.step(a)
            o["step_reward"][i, a] = r
            o["step_grid"][i, a] = to_ids(s2.grid).reshape(-1)
            o["step_inv"][i, a] = s2.inventory
            o["step_pos"][i, a] = s2.pos
            o["step_dir"][i, a] = s2.dir
        for task in tm.tasks:
            tid = tm.tasks[task]
            sat = s.satisfies(task)
            o["satisfies"][i, tid] = 2 if sat is None else int(bool(sat))
            try:
                o["expert"][i, tid] = teacher(task, s)
            except AssertionError:
                o["expert"][i, tid] = ERR_ASSERT
            except TypeError:
                o["expert"][i, tid] = ERR_TYPE
        for j, kind in enumerate(go_kinds):
            try:
                goal, seq = teacher.find_closest_resources(go_task[kind], s)
            except TypeError:
                o["closest_status"][i, j] = 2
                continue
            if goal is not None:
                o["closest_goal"][i, j] = goal
            if seq is None:
                o["closest_status"][i, j] = 1
            else:
                o["closest_len"][i, j] = len(seq)
                o["closest_seq"][i, j, :len(seq)] = seq
        if i % 2000 == 0:
            print("  state", i, "/", n, flush=True)
    # bad actions raise (worlds/craft.py:415-416)
    for bad in (-1, 6, 7):
        try:
            states[0].step(bad)
            raise SystemExit("reference accepted action %d" % bad)
        except Exception as e:  # noqa: BLE001
            assert "Unexpected action" in str(e)
    np.savez_compressed(path, **o)
    print("wrote", path, n, "states; expert histogram",
          np.bincount(o["expert"][:, 13:].reshape(-1), minlength=256)[[0, 1, 2, 3, 4, 5, 254, 255]])


def export_light(R, path, seed=7):
    rng = np.random.RandomState(seed)
    lw = R.light_world()
    goals = [g for g in lw.cookbook.index.ordered_contents
             if g in ("LL", "LD", "RD", "UL", "UR", "URU", "DRU", "LLD", "RDD", "LUR")]
    MAXB, MAXD, MAXK = 31, 8, 8
    recs = []
    for rep in range(6):
        for goal in goals:
            scen = lw.sample_scenario_with_goal(lw.cookbook.index[goal])
            s = scen.init()
            for t in range(120):
                recs.append((scen, s))
                a = int(rng.choice(5, p=[.21, .21, .21, .21, .16]))
                _, s = s.step(a)
    n = len(recs)
    o = dict(
        walls=np.ones((n, MAXB, MAXB), np.uint8), board=np.zeros((n, 2), np.uint8),
        doors=np.full((n, MAXD, 2), 255, np.uint8), n_doors=np.zeros(n, np.uint8),
        keys=np.full((n, MAXK, 4), 255, np.uint8), n_keys=np.zeros(n, np.uint8),
        key_alive=np.zeros((n, MAXK), np.uint8),
        goal_room=np.zeros((n, 2), np.uint8), pos=np.zeros((n, 2), np.uint8),
        features=np.zeros((n, 12), np.float32), satisfies=np.zeros(n, np.uint8),
        step_pos=np.zeros((n, 5, 2), np.uint8), step_key_alive=np.zeros((n, 5, MAXK), np.uint8),
    )
    for i, (scen, s) in enumerate(recs):
        bw, bh = scen.walls.shape
        o["board"][i] = (bw, bh)
        o["walls"][i, :bw, :bh] = scen.walls
        o["n_doors"][i] = len(scen.doors)
        for j, d in enumerate(scen.doors):
            o["doors"][i, j] = d
        all_keys = list(scen.keys.items())          # scenario's full key set, insertion order
        o["n_keys"][i] = len(all_keys)
        for j, (k, d) in enumerate(all_keys):
            o["keys"][i, j] = (k[0], k[1], d[0], d[1])
            o["key_alive"][i, j] = k in s.keys
        o["goal_room"][i] = scen.goal_room
        o["pos"][i] = s.pos
        o["features"][i] = s.features()
        o["satisfies"][i] = bool(s.satisfies(None, None))
        for a in range(5):
            _, s2 = s.step(a)
            o["step_pos"][i, a] = s2.pos
            for j, (k, d) in enumerate(all_keys):
                o["step_key_alive"][i, a, j] = k in s2.keys
    np.savez_compressed(path, **o)
    print("wrote", path, n, "light states; boards", sorted(set(map(tuple, o["board"].tolist())))[:6],
          "max doors", o["n_doors"].max(), "max keys", o["n_keys"].max())


def export_sampler_stats(R, path, n_scen=3000, seed=4242):
    """Runs the reference's own sampler (make_data.py:27-144: all_free_cells_reachable,
    random_free, sample_scenario — the function definitions are exec'ed from the source text, the
    module-level dataset script after them is not) and stores the cell-occupancy statistics."""
    src = open(os.path.join(ref_shim.REF_ROOT, "make_data.py")).read()
    head = src[:src.index("config = flags.make_config()")]
    head = head.replace("import flags", "").replace("import models", "")
    ns = {"__name__": "__make_data_defs__"}
    with ref_shim.reference_cwd():
        exec(compile(head, "make_data_defs", "exec"), ns)
    world = R.world
    ns["world"] = world                       # all_free_cells_reachable reads the global
    cfg = R.config
    cfg.random = np.random.RandomState(seed)
    W, H, K = world.WIDTH, world.HEIGHT, world.cookbook.n_kinds
    kind_cell = np.zeros((K, W, H), np.int64)
    pos_cell = np.zeros((W, H), np.int64)
    n_free_nbr = np.zeros(5, np.int64)        # free 4-neighbours of placed items
    grids = []
    for i in range(n_scen):
        grid, init_pos = ns["sample_scenario"](world, None, cfg)
        ids = to_ids(grid)
        for k in range(2, K):
            kind_cell[k] += ids == k
        pos_cell[init_pos] += 1
        for x, y in np.argwhere((ids > 1)):
            nb = sum(ids[x + dx, y + dy] == 0 for dx, dy in ((0, 1), (0, -1), (1, 0), (-1, 0)))
            n_free_nbr[nb] += 1
        if i < 64:
            grids.append(ids.reshape(-1))
        if i % 500 == 0:
            print("  sampled", i, flush=True)
    np.savez_compressed(path, n_scen=np.int64(n_scen), kind_cell=kind_cell, pos_cell=pos_cell,
                        n_free_nbr=n_free_nbr, example_grids=np.stack(grids))
    print("wrote", path)


class ScriptedStudent(object):
    """Stands in for students/imitation.py in the reference's own ImitationTrainer.do_rollout:
    reads the features of every state like the real student (students/imitation.py:72), plays a
    fixed action script and records what the trainer hands back."""

    def __init__(self, script):
        self.script = script
        self.features, self.received = [], []

    def init(self, tasks, states, is_eval):
        self.t = 0

    def act(self, states):
        self.features.append(np.stack([s.features() for s in states]))
        actions = [int(a) for a in self.script[self.t]]
        self.t += 1
        return actions

    def receive(self, ref_actions):
        self.received.append(list(ref_actions))


def export_trainer_rollouts(R, splits, path, batch_size=32, seed=77):
    """Runs the UNMODIFIED trainers/imitation.py:ImitationTrainer.do_rollout (training mode with a
    behaviour-cloning mix of 0.3, and evaluation mode) on the reference world and teacher and
    stores everything it produced."""
    with ref_shim.reference_cwd():
        from trainers.imitation import ImitationTrainer
    rng = np.random.RandomState(seed)
    world, teacher, tm = R.world, R.teacher, R.task_manager
    K = world.cookbook.n_kinds
    idx = rng.choice(len(splits["dev_inst_env"]), size=batch_size, replace=False)
    script = rng.choice(6, size=(64, batch_size), p=[.2, .2, .2, .2, .17, .03]).astype(np.uint8)
    batch = []
    for i in idx:
        ids = splits["dev_grids"][splits["dev_inst_env"][i]].reshape(world.WIDTH, world.HEIGHT)
        batch.append(dict(grid=to_onehot(ids, K), init_pos=tuple(int(v) for v in splits["dev_inst_pos"][i]),
                          task=tm.tasks.get(int(splits["dev_inst_task"][i]))))
    out = dict(inst=idx.astype(np.int32), script=script, mix_rate=np.float64(0.3), mix_seed=np.int64(5))
    trainer = ImitationTrainer(R.config)
    trainer.policy_mix_rate = 0.3
    for mode, is_eval in (("train", False), ("eval", True)):
        R.config.random = np.random.RandomState(5)
        student = ScriptedStudent(script)
        info = trainer.do_rollout(batch, world, student, teacher, is_eval)
        T = len(student.features)
        acts = np.full((batch_size, T), 255, np.uint8)
        for i, seq in enumerate(info["action_seqs"]):
            acts[i, :len(seq)] = seq
        out[mode + "_action_seqs"] = acts
        out[mode + "_success"] = np.asarray([bool(v) for v in info["success"]])
        out[mode + "_distances"] = np.asarray(info["distances"], np.int32)
        out[mode + "_num_interactions"] = np.int64(info["num_interactions"])
        out[mode + "_num_steps"] = np.int64(info["num_steps"])
        out[mode + "_features"] = np.stack(student.features).astype(np.uint8)
        if student.received:
            out[mode + "_ref_actions"] = np.asarray(student.received, np.int16)
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--trainer" in sys.argv:
        R = ref_shim.Reference()
        export_trainer_rollouts(R, np.load(os.path.join(OUT, "craft_medium_splits.npz")),
                                os.path.join(OUT, "trainer_rollouts.npz"))
        return
    if "--custom" in sys.argv:
        # a non-default cookbook / hint file through the unmodified reference
        cdir = os.path.join(OUT, "custom")
        RC = ref_shim.Reference(recipes=os.path.join(cdir, "recipes.yaml"),
                                hints=os.path.join(cdir, "hints.yaml"))
        states = collect_states(RC, None, 0, 0, 2000, seed=31337)
        export_states(RC, states, os.path.join(OUT, "craft_custom_states.npz"))
        return
    if "--sampler-only" in sys.argv:
        export_sampler_stats(ref_shim.Reference(), os.path.join(OUT, "sampler_stats.npz"))
        return
    regen_dir = os.environ.get("PSK_REGEN_DIR", "/tmp/psk_data")
    if not os.path.exists(os.path.join(regen_dir, "craft_medium_train.json")):
        print("regenerating the dataset with the reference's make_data.py (~30 s)")
        ref_shim.regenerate_dataset(regen_dir)
    R = ref_shim.Reference()
    splits = export_splits(R, regen_dir)
    states = collect_states(R, splits, n_on=3000, n_off=3000, n_pert=4000, seed=20261018)
    export_states(R, states, os.path.join(OUT, "craft_medium_states.npz"))
    RL = ref_shim.Reference(world_config="craft_large")
    states = collect_states(RL, None, 0, 0, 2500, seed=99)
    export_states(RL, states, os.path.join(OUT, "craft_large_states.npz"))
    export_light(R, os.path.join(OUT, "light_states.npz"))
    export_sampler_stats(ref_shim.Reference(), os.path.join(OUT, "sampler_stats.npz"))


if __name__ == "__main__":
    main()
