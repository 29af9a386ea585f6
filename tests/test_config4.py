"""BASELINE config 4: the record of the reference's own ImitationTrainer.train + ImitationStudent +
LSTMSeq2SeqModel run (tests/golden/config4_imitation.npz, written by oracle/gen_config4.py from the
UNMODIFIED reference: 30 DAgger iterations of batch 32 and two evaluations of the whole dev split)
replayed on this repo's drop-in world and teacher.

gen_config4.py itself runs the reference's trainer/student code twice — on the reference world and on
the psketch_b200 façade — and asserts bit-identical batches, features, actions, teacher labels and
losses.  Here the same record drives the façade through the trainer's protocol
(trainers/imitation.py:18-101, restated in tests/trainer_loop.py) with the student's recorded
actions, because the reference's student code cannot travel to the GPU box: everything the
environment and the teacher produced must come out identical — every feature block the student read
(hashed), every teacher label, success flags, distances, counters."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN
from trainer_loop import ScriptedStudent, run_protocol


def _hash(feats):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(feats, np.float32).tobytes()).digest()[:8],
                         np.uint64)[0]


def _items(world, splits, split, rows):
    K = world.cookbook.n_kinds
    out = []
    for i in rows:
        ids = splits[split + "_grids"][splits[split + "_inst_env"][i]].reshape(world.WIDTH, world.HEIGHT)
        onehot = np.zeros((world.WIDTH, world.HEIGHT, K))
        xs, ys = np.nonzero(ids)
        onehot[xs, ys, ids[xs, ys]] = 1
        out.append(dict(grid=onehot, init_pos=tuple(int(v) for v in splits[split + "_inst_pos"][i]),
                        task=world.task_manager.by_id(int(splits[split + "_inst_task"][i]))))
    return out


def replay(world, teacher, splits, fx, rollouts):
    n_train = int(fx["n_train_iters"])
    checked = 0
    for r in rollouts:
        B, T, is_eval = int(fx["n_env"][r]), int(fx["n_t"][r]), bool(fx["is_eval"][r])
        assert is_eval == (r >= n_train)
        batch = _items(world, splits, "dev" if is_eval else "train", fx["batch"][r, :B])
        student = ScriptedStudent(fx["acts"][r, :T, :B])
        got = run_protocol(batch, world, teacher, student, is_eval, 40, 0.0, np.random.RandomState(0))
        assert student.t == T, r                                   # same number of timesteps
        for i, seq in enumerate(got["action_seqs"]):
            L = int(fx["seq_len"][r, i])
            assert seq == fx["acts"][r, :L, i].tolist(), (r, i)
        assert [bool(v) for v in got["success"]] == fx["success"][r, :B].astype(bool).tolist(), r
        assert got["distances"] == fx["distances"][r, :int(fx["n_dist"][r])].tolist(), r
        assert got["num_interactions"] == int(fx["num_interactions"][r]), r
        assert got["num_steps"] == int(fx["num_steps"][r]), r
        assert [int(_hash(f)) for f in student.features] == [int(h) for h in fx["feat_hash"][r, :T]], r
        if not is_eval:
            assert np.array_equal(np.asarray(student.received, np.int8), fx["refs"][r, :T, :B]), r
        checked += 1
    return checked


@pytest.fixture(scope="module")
def config4():
    return np.load(os.path.join(GOLDEN, "config4_imitation.npz"))


def test_config4_record_is_a_real_training_run(config4):
    fx = config4
    n_train = int(fx["n_train_iters"])
    assert n_train >= 30 and len(fx["eval_sizes"]) == 2 and int(fx["eval_sizes"][0]) == 69   # 2,200 / 32
    loss = fx["loss"][:n_train]
    assert np.isfinite(loss).all() and loss[-5:].mean() < loss[:5].mean()     # it learns
    assert (fx["refs"][:n_train][fx["acts"][:n_train] != 255] >= -1).all()


def test_config4_replay_on_the_facade_logic(config4, splits, medium_tables, medium_oracle):
    """CPU tier: the façade's host logic with the oracle-backed test double as its device backend,
    on a slice of the record (first / last training iterations, a few evaluation batches)."""
    from test_facade_cpu import _world
    from psketch_b200.teachers import DemonstrationTeacher
    world = _world(medium_tables, medium_oracle)
    n_train = int(config4["n_train_iters"])
    picks = [0, 1, n_train - 1, n_train, n_train + 68, n_train + 69]
    assert replay(world, DemonstrationTeacher(None), splits, config4, picks) == len(picks)


@pytest.mark.gpu
def test_config4_replay_on_the_gpu(config4, splits):
    """GPU tier: the whole record (30 training rollouts + 2 x 69 evaluation batches, ~3,900 batched
    launch sequences) through psketch_b200.worlds.CraftWorld / DemonstrationTeacher with the CUDA
    backend, constructed from the experiment config exactly as worlds.load(config) would."""
    import yaml
    from psketch_b200 import teachers, worlds
    from psketch_b200.worlds.craft import _Struct
    cfg = _Struct(**yaml.safe_load("""
recipes: "resources/craft/recipes.yaml"
world: {name: CraftWorld, config: craft_medium}
student: {name: ImitationStudent, model: {name: LSTMSeq2SeqModel, hidden_size: 256}}
teacher: {name: DemonstrationTeacher}
trainer: {name: ImitationTrainer, hints: "resources/craft/hints.hierarchy.yaml", max_timesteps: 40, batch_size: 32}
"""))
    cfg.random = np.random.RandomState(123)
    world, teacher = worlds.load(cfg), teachers.load(cfg)
    assert cfg.student.model.input_size == 404 and cfg.student.model.n_actions == 6
    total = int(config4["n_train_iters"]) + int(config4["eval_sizes"].sum())
    assert replay(world, teacher, splits, config4, range(total)) == total


# ------------------------------------------------------------------------------------------------
# primitive_language.yaml (trainers/primitive_language.py:16-143 + teachers/primitive_language.py)
WORDS = ("down", "up", "left", "right", "use", "stop")


def _word_ids(seqs):
    return [[WORDS.index(w) for w in ws] for ws in seqs]


def replay_language(world, teacher, splits, fx, rollouts, rng):
    """``rng`` is the experiment's shared numpy stream (config.random): the dataset shuffles and the
    teacher's describe() draw from it in the order the reference's run did, so the batches
    themselves are re-derived here and must match the record."""
    from trainer_loop import ScriptedLanguageStudent, run_language_protocol
    n_train = int(fx["n_train_iters"])
    sizes = [int(v) for v in fx["eval_sizes"]]
    log_every = n_train // len(sizes)
    order = {"train": None, "dev": None}
    cursor = {"train": 0, "dev": 0}
    checked = 0
    for r in rollouts:
        B, T, is_eval = int(fx["n_env"][r]), int(fx["n_t"][r]), bool(fx["is_eval"][r])
        split = "dev" if is_eval else "train"
        if cursor[split] == 0:                                     # data/dataset.py:70-72
            order[split] = list(range(len(splits[split + "_inst_env"])))
            rng.shuffle(order[split])
        rows = order[split][cursor[split]:cursor[split] + 32]
        cursor[split] = cursor[split] + 32 if cursor[split] + 32 < len(order[split]) else 0
        assert rows == fx["batch"][r, :B].tolist(), r
        batch = _items(world, splits, split, rows)
        for item, i in zip(batch, rows):
            item["ref_actions"] = tuple(int(a) for a in splits[split + "_ref_actions"][i][:splits[split + "_ref_len"][i]])
        T2 = int(fx["phase2_n_t"][r])
        student = ScriptedLanguageStudent(fx["acts"][r, :T, :B], fx["phase2_acts"][r, :T2, :B])
        got = run_language_protocol(batch, world, teacher, student, is_eval, 40)
        want_instr = [[int(w) for w in row if w != 255] for row in fx["instructions"][r, :B]]
        assert _word_ids(got["instructions"]) == want_instr, r
        last = fx["acts"] if is_eval else fx["phase2_acts"]
        for i, seq in enumerate(got["action_seqs"]):
            L = int(fx["seq_len"][r, i])
            assert seq == last[r, :L, i].tolist(), (r, i)
        if not is_eval:
            want = [[int(w) for w in row if w != 255] for row in fx["descriptions"][r, :B]]
            assert _word_ids(got["descriptions"]) == want, r
            assert student.t == T2 and len(student.features[0]) == T, r
            assert [int(_hash(f)) for f in student.features[1]] == [int(h) for h in fx["phase2_feat_hash"][r, :T2]], r
        assert [int(_hash(f)) for f in student.features[0]] == [int(h) for h in fx["feat_hash"][r, :T]], r
        assert [bool(v) for v in got["success"]] == fx["success"][r, :B].astype(bool).tolist(), r
        assert got["distances"] == fx["distances"][r, :int(fx["n_dist"][r])].tolist(), r
        assert got["num_interactions"] == int(fx["num_interactions"][r]), r
        assert got["num_steps"] == int(fx["num_steps"][r]), r
        checked += 1
    learned = [WORDS.index(teacher.student_action_map[a]) if a in teacher.student_action_map else 255
               for a in range(6)]
    return checked, learned


def _language_order(fx):
    """Rollout indices of the record in the order the run produced them: log_every training
    iterations, one evaluation of the dev split, and so on."""
    n_train, sizes = int(fx["n_train_iters"]), [int(v) for v in fx["eval_sizes"]]
    per = n_train // len(sizes)
    out, ev = [], n_train
    for k, sz in enumerate(sizes):
        out += list(range(k * per, (k + 1) * per)) + list(range(ev, ev + sz))
        ev += sz
    return out


@pytest.fixture(scope="module")
def config4_language():
    return np.load(os.path.join(GOLDEN, "config4_primitive_language.npz"))


def test_config4_language_replay_on_the_facade_logic(config4_language, splits, medium_tables, medium_oracle):
    """CPU tier: the first training iterations of the primitive_language.yaml record (describe()
    learns the student's action ids there and draws from the shared random stream)."""
    from test_facade_cpu import _world
    from psketch_b200.teachers import PrimitiveLanguageTeacher
    fx = config4_language
    world = _world(medium_tables, medium_oracle)
    rng = np.random.RandomState(123)
    cfg = type("Cfg", (), {"random": rng})()
    checked, _ = replay_language(world, PrimitiveLanguageTeacher(cfg), splits, fx, range(6), rng)
    assert checked == 6


@pytest.mark.gpu
def test_config4_language_replay_on_the_gpu(config4_language, splits):
    """GPU tier: the whole primitive_language.yaml record (30 training rollouts of two decoding passes
    each + 2 evaluations of the dev split) on the CUDA-backed world and language teacher."""
    from psketch_b200 import teachers, worlds
    from psketch_b200.worlds.craft import _Struct
    fx = config4_language
    cfg = _Struct(recipes="resources/craft/recipes.yaml",
                  world={"name": "CraftWorld", "config": "craft_medium"},
                  teacher={"name": "PrimitiveLanguageTeacher"},
                  trainer={"name": "PrimitiveLanguageTrainer", "hints": "resources/craft/hints.hierarchy.yaml",
                           "max_timesteps": 40, "batch_size": 32})
    rng = np.random.RandomState(123)
    cfg.random = rng
    world, teacher = worlds.load(cfg), teachers.load(cfg)
    order = _language_order(fx)
    checked, learned = replay_language(world, teacher, splits, fx, order, rng)
    assert checked == len(order)
    assert learned == fx["final_action_map"].tolist()


# ------------------------------------------------------------------------------------------------
# interactive_primitive_language.yaml (trainers/interactive_primitive_language.py:16-106 +
# teachers/interactive_primitive_language.py): the ONLINE teacher is asked every timestep for every
# env (finished ones too) and describes every executed action from a (previous, new) state pair.
def replay_interactive(world, teacher, splits, fx, rollouts, rng):
    from trainer_loop import ScriptedInteractiveStudent, run_interactive_protocol
    order = {"train": None, "dev": None}
    cursor = {"train": 0, "dev": 0}
    checked = 0
    for r in rollouts:
        B, T, is_eval = int(fx["n_env"][r]), int(fx["n_t"][r]), bool(fx["is_eval"][r])
        split = "dev" if is_eval else "train"
        if cursor[split] == 0:
            order[split] = list(range(len(splits[split + "_inst_env"])))
            rng.shuffle(order[split])
        rows = order[split][cursor[split]:cursor[split] + 32]
        cursor[split] = cursor[split] + 32 if cursor[split] + 32 < len(order[split]) else 0
        assert rows == fx["batch"][r, :B].tolist(), r
        batch = _items(world, splits, split, rows)
        student = ScriptedInteractiveStudent(fx["acts"][r, :T, :B])
        got = run_interactive_protocol(batch, world, teacher, student, is_eval, 40)
        assert student.t == T, r
        for i, seq in enumerate(got["action_seqs"]):
            L = int(fx["seq_len"][r, i])
            assert seq == fx["acts"][r, :L, i].tolist(), (r, i)
        ids = lambda rows_: [[255 if w is None else WORDS.index(w[0]) for w in row] for row in rows_]
        if not is_eval:
            assert ids(student.instructions) == fx["instr_steps"][r, :T, :B].tolist(), r
        # the student is handed the descriptions only in training; eval still computes them
        if not is_eval:
            assert ids(student.descriptions) == fx["desc_steps"][r, :T, :B].tolist(), r
        assert [int(_hash(f)) for f in student.features] == [int(h) for h in fx["feat_hash"][r, :T]], r
        assert [bool(v) for v in got["success"]] == fx["success"][r, :B].astype(bool).tolist(), r
        assert got["distances"] == fx["distances"][r, :int(fx["n_dist"][r])].tolist(), r
        assert got["num_interactions"] == int(fx["num_interactions"][r]), r
        assert got["num_steps"] == int(fx["num_steps"][r]), r
        checked += 1
    learned = [WORDS.index(teacher.student_action_map[a]) if a in teacher.student_action_map else 255
               for a in range(6)]
    return checked, learned


@pytest.fixture(scope="module")
def config4_interactive():
    return np.load(os.path.join(GOLDEN, "config4_interactive_primitive_language.npz"))


def test_config4_interactive_replay_on_the_facade_logic(config4_interactive, splits, medium_tables, medium_oracle):
    from test_facade_cpu import _world
    from psketch_b200.teachers import InteractivePrimitiveLanguageTeacher
    world = _world(medium_tables, medium_oracle)
    rng = np.random.RandomState(123)
    cfg = type("Cfg", (), {"random": rng})()
    checked, _ = replay_interactive(world, InteractivePrimitiveLanguageTeacher(cfg), splits,
                                    config4_interactive, range(5), rng)
    assert checked == 5


@pytest.mark.gpu
def test_config4_interactive_replay_on_the_gpu(config4_interactive, splits):
    from psketch_b200 import teachers, worlds
    from psketch_b200.worlds.craft import _Struct
    fx = config4_interactive
    cfg = _Struct(recipes="resources/craft/recipes.yaml",
                  world={"name": "CraftWorld", "config": "craft_medium"},
                  teacher={"name": "InteractivePrimitiveLanguageTeacher"},
                  trainer={"name": "InteractivePrimitiveLanguageTrainer",
                           "hints": "resources/craft/hints.hierarchy.yaml", "max_timesteps": 40, "batch_size": 32})
    rng = np.random.RandomState(123)
    cfg.random = rng
    world, teacher = worlds.load(cfg), teachers.load(cfg)
    order = _language_order(fx)
    checked, learned = replay_interactive(world, teacher, splits, fx, order, rng)
    assert checked == len(order)
    assert learned == fx["final_action_map"].tolist()


@pytest.mark.gpu
def test_batched_language_rollouts_reproduce_the_reference_trainer(config4_language, splits, medium_tables):
    """psketch_b200.rollout.language_rollouts — trainers/primitive_language.py:16-143 for a whole batch
    as device ops (instruct_batch, two decoding passes with every action executed, describe_batch on
    the recorded agent records, distances on the original grids) — driven by the recorded actions of
    the reference's own PrimitiveLanguageTrainer run: instructions, descriptions (incl. the random
    words drawn while the teacher still learns the student's action ids), final action sequences,
    success, distances and counters must equal the record."""
    import torch
    from psketch_b200.rollout import language_rollouts
    from psketch_b200.teachers import PrimitiveLanguageTeacher
    from psketch_b200.vec import VecCraft
    fx = config4_language
    rng = np.random.RandomState(123)
    teacher = PrimitiveLanguageTeacher(type("Cfg", (), {"random": rng})())
    order = list(range(len(splits["train_inst_env"])))
    rng.shuffle(order)                                             # data/dataset.py:70-72
    for r in range(12):                                            # the first training iterations
        B, T, T2 = int(fx["n_env"][r]), int(fx["n_t"][r]), int(fx["phase2_n_t"][r])
        rows = order[32 * r:32 * r + 32]
        assert rows == fx["batch"][r, :B].tolist()
        env = VecCraft.from_instances(medium_tables, splits["train_grids"], splits["train_inst_env"][rows],
                                      splits["train_inst_pos"][rows], splits["train_inst_task"][rows],
                                      max_timesteps=40)
        p1 = torch.from_numpy(fx["acts"][r]).to(env.device)        # [40, 32], 255 = terminated
        p2 = torch.from_numpy(fx["phase2_acts"][r]).to(env.device)
        got = language_rollouts(env, teacher, lambda f, t: p1[t, :B], splits["train_ref_actions"][rows],
                                greedy_policy=lambda f, t: p2[t, :B])
        instr = got["instructions"].numpy() - 1                    # instruct_batch: 1 + action index, 0 = padding
        for i in range(B):
            want = [int(w) for w in fx["instructions"][r, i] if w != 255]
            assert instr[i][instr[i] >= 0].tolist() == want, (r, i)
            wd = [int(w) for w in fx["descriptions"][r, i] if w != 255]
            gd = got["descriptions"][i].cpu().numpy()
            assert gd[gd >= 0].tolist() == wd, (r, i)
            L = int(fx["seq_len"][r, i])
            assert got["action_seqs"][i, :L].tolist() == fx["phase2_acts"][r, :L, i].tolist(), (r, i)
            assert (got["action_seqs"][i, L:] == 255).all()
        assert got["success"].tolist() == fx["success"][r, :B].astype(bool).tolist(), r
        d = got["distances"]
        assert d[d >= 0].tolist() == fx["distances"][r, :int(fx["n_dist"][r])].tolist(), r
        assert got["num_steps"] == int(fx["num_steps"][r]) and got["num_interactions"] == int(fx["num_interactions"][r]), r
    assert len(teacher.student_action_map) >= 5


@pytest.mark.gpu
def test_batched_policy_rollouts_reproduce_the_reference_trainer(config4, splits, medium_tables):
    """psketch_b200.rollout.policy_rollouts — ImitationTrainer.do_rollout for a whole batch, one
    step-then-observe tick per timestep — driven by the student's recorded actions from the reference's
    own ImitationTrainer run: teacher labels, executed action sequences, success, distances and
    counters of all 30 training rollouts and of the first evaluation of the dev split must equal the
    record."""
    import torch
    from psketch_b200.rollout import policy_rollouts
    from psketch_b200.vec import VecCraft
    fx = config4
    n_train = int(fx["n_train_iters"])
    for r in list(range(n_train)) + list(range(n_train, n_train + int(fx["eval_sizes"][0]))):
        B, T, is_eval = int(fx["n_env"][r]), int(fx["n_t"][r]), bool(fx["is_eval"][r])
        split = "dev" if is_eval else "train"
        rows = fx["batch"][r, :B]
        env = VecCraft.from_instances(medium_tables, splits[split + "_grids"], splits[split + "_inst_env"][rows],
                                      splits[split + "_inst_pos"][rows], splits[split + "_inst_task"][rows],
                                      max_timesteps=40)
        script = torch.from_numpy(fx["acts"][r]).to(env.device)             # [40, 32]
        got = policy_rollouts(env, lambda f, t: script[t, :B], max_timesteps=40, is_eval=is_eval, poll_every=4)
        for i in range(B):
            L = int(fx["seq_len"][r, i])
            assert got["action_seqs"][i, :L].tolist() == fx["acts"][r, :L, i].tolist(), (r, i)
            assert (got["action_seqs"][i, L:] == 255).all(), (r, i)
            if not is_eval:
                want = fx["refs"][r, :T, i]
                live = want >= 0
                assert got["ref_seqs"][i, :T][live].tolist() == want[live].tolist(), (r, i)
                assert (got["ref_seqs"][i, :T][~live] == 255).all(), (r, i)
        assert got["success"].tolist() == fx["success"][r, :B].astype(bool).tolist(), r
        d = got["distances"]
        assert d[d >= 0].tolist() == fx["distances"][r, :int(fx["n_dist"][r])].tolist(), r
        assert got["num_steps"] == int(fx["num_steps"][r]), r
        assert got["num_interactions"] == int(fx["num_interactions"][r]), r


@pytest.mark.gpu
def test_batched_interactive_rollouts_reproduce_the_reference_trainer(config4_interactive, splits, medium_tables):
    """psketch_b200.rollout.interactive_rollouts against the record of the reference's own
    InteractivePrimitiveLanguageTrainer run (first training iterations): per-timestep instructions
    for every env, per-step descriptions (with the random draws of the learning phase), action
    sequences, success, distances, counters."""
    import torch
    from psketch_b200.rollout import interactive_rollouts
    from psketch_b200.teachers import InteractivePrimitiveLanguageTeacher
    from psketch_b200.vec import VecCraft
    fx = config4_interactive
    rng = np.random.RandomState(123)
    teacher = InteractivePrimitiveLanguageTeacher(type("Cfg", (), {"random": rng})())
    order = list(range(len(splits["train_inst_env"])))
    rng.shuffle(order)
    for r in range(12):
        B, T = int(fx["n_env"][r]), int(fx["n_t"][r])
        rows = order[32 * r:32 * r + 32]
        assert rows == fx["batch"][r, :B].tolist()
        env = VecCraft.from_instances(medium_tables, splits["train_grids"], splits["train_inst_env"][rows],
                                      splits["train_inst_pos"][rows], splits["train_inst_task"][rows],
                                      max_timesteps=40)
        script = torch.from_numpy(fx["acts"][r]).to(env.device)
        got = interactive_rollouts(env, teacher, lambda f, t, w: script[t, :B], max_timesteps=40, poll_every=1)
        assert got["timesteps"] == T, r
        assert np.array_equal(got["instructions"], fx["instr_steps"][r, :T, :B]), r
        for i in range(B):
            L = int(fx["seq_len"][r, i])
            assert got["action_seqs"][i, :L].tolist() == fx["acts"][r, :L, i].tolist(), (r, i)
            assert got["descriptions"][:L, i].tolist() == fx["desc_steps"][r, :L, i].tolist(), (r, i)
            assert (got["descriptions"][L:, i] == -1).all(), (r, i)
        assert got["success"].tolist() == fx["success"][r, :B].astype(bool).tolist(), r
        d = got["distances"]
        assert d[d >= 0].tolist() == fx["distances"][r, :int(fx["n_dist"][r])].tolist(), r
        assert got["num_steps"] == int(fx["num_steps"][r]) and got["num_interactions"] == int(fx["num_interactions"][r]), r
