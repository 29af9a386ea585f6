"""BASELINE config 4 — the reference's OWN trainer, student and model, unchanged, on this repo's
drop-in world and teacher.  TEST INFRASTRUCTURE, build container only (imports /root/reference).

    python -m oracle.gen_config4 [--iters 30] [--log-every 15]        # ~2-4 minutes

Runs ``trainers.imitation.ImitationTrainer.train`` (trainers/imitation.py:103-180) with
``students.imitation.ImitationStudent`` + ``models.lstm_seq2seq.LSTMSeq2SeqModel`` on the CPU for a
few DAgger iterations (configs/experiments/imitation.yaml, seed 123, batch 32, evaluation on the whole
dev split every ``log_every`` iterations), twice with identical seeds:

  (a) on the reference's world and teacher  (worlds.load / teachers.load of the reference);
  (b) with ONLY those two factories swapped for psketch_b200.worlds.load / psketch_b200.teachers.load
       — in this container the façade's device backend is the oracle-backed test
      double of tests/test_facade_cpu.py (no GPU here); tests/test_config4_gpu.py replays (b) with the
      CUDA backend on the GPU box.

and asserts that (a) and (b) agree on everything the trainer and the student ever saw or produced:
the batches, every feature vector (hashed), every sampled / greedy action, every teacher label, the
loss of every iteration (bit-equal floats), success flags, distances, counters and the evaluation
trajectories.  The record of (a) is committed as tests/golden/config4_imitation.npz.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def feature_hash(feats_f32):
    """64 bits of SHA-1 over the float32 feature block the student consumed at one timestep."""
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(feats_f32, np.float32).tobytes()).digest()[:8],
                         np.uint64)[0]


def prepare_data(regen_dir, data_dir):
    """The three split files where data.load expects them (data/dataset.py:17-18).  The train split is
    missing from the reference checkout and is regenerated with the reference's make_data.py; its
    Task.__str__ prints 'get wood' where the loader wants 'get[wood]' (SURVEY §0), so the task
    strings are rewritten — data preparation, the loader itself is untouched."""
    if not os.path.exists(os.path.join(regen_dir, "craft_medium_train.json")):
        print("regenerating the dataset with the reference's make_data.py (~30 s)")
        ref_shim.regenerate_dataset(regen_dir)
    os.makedirs(data_dir, exist_ok=True)
    for split in ("train", "dev", "test"):
        data = json.load(open(os.path.join(regen_dir, "craft_medium_%s.json" % split)))
        for env in data:
            for ti in env["task_instances"]:
                if "[" not in ti["task"]:
                    ti["task"] = "%s[%s]" % tuple(ti["task"].split(" "))
        json.dump(data, open(os.path.join(data_dir, "craft_medium_%s.json" % split), "w"))


def make_config(data_dir, exp_dir, iters, log_every, seed=123, experiment="imitation"):
    import torch
    cfg = ref_shim.make_config(experiment, seed)
    cfg.data_dir = data_dir
    cfg.experiment_dir = exp_dir
    cfg.trainer.max_iters = iters
    cfg.trainer.log_every = log_every
    cfg.device = torch.device("cpu")
    torch.manual_seed(seed)
    return cfg


def run(which, data_dir, iters, log_every, medium_oracle=None, experiment="imitation"):
    """One training run; returns the record.  ``which``: 'reference' or 'facade'."""
    import torch
    ref_shim._install_shims()
    if ref_shim.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_shim.REF_ROOT)
    exp_dir = tempfile.mkdtemp(prefix="psk_config4_")
    with ref_shim.reference_cwd():
        import data as ref_data
        import students as ref_students
        import teachers as ref_teachers
        import trainers as ref_trainers
        import worlds as ref_worlds
        from students.imitation import ImitationStudent
        from students.primitive_language import PrimitiveLanguageStudent
        from trainers.imitation import ImitationTrainer
        from trainers.primitive_language import PrimitiveLanguageTrainer

        rec = dict(iters=[], evals=[])

        class RecordingLanguageStudent(PrimitiveLanguageStudent):     # student code runs unchanged
            def init(self, tasks, instructions, states, is_eval):
                super().init(tasks, instructions, states, is_eval)
                self.cur = dict(is_eval=is_eval, acts=[], refs=[], fh=[], phase2_acts=[], phase2_fh=[],
                                instructions=[list(w) for w in instructions], descriptions=None)

            def _log(self, states, actions):
                feats = np.stack([np.asarray(s.features()) for s in states]).astype(np.float32)
                second = self.cur["descriptions"] is not None
                self.cur["phase2_fh" if second else "fh"].append(feature_hash(feats))
                self.cur["phase2_acts" if second else "acts"].append(list(actions))

            def act(self, states, t):
                actions = super().act(states, t)
                self._log(states, actions)
                return actions

            def instructed_act(self, states, t):
                actions = super().instructed_act(states, t)
                self._log(states, actions)
                return actions

            def receive(self, descriptions):
                self.cur["descriptions"] = [list(d) for d in descriptions]
                super().receive(descriptions)

            def learn(self):
                loss = super().learn()
                rec["iters"][-1]["loss"] = loss
                return loss

        class RecordingStudent(ImitationStudent):          # the student's code runs unchanged
            def init(self, tasks, states, is_eval):
                super().init(tasks, states, is_eval)
                self.cur = dict(is_eval=is_eval, acts=[], refs=[], fh=[])

            def act(self, states):
                feats = np.stack([np.asarray(s.features()) for s in states]).astype(np.float32)
                self.cur["fh"].append(feature_hash(feats))
                actions = super().act(states)
                self.cur["acts"].append(list(actions))
                return actions

            def receive(self, ref_actions):
                self.cur["refs"].append(list(ref_actions))
                super().receive(ref_actions)

            def learn(self):
                loss = super().learn()
                rec["iters"][-1]["loss"] = loss
                return loss

        from students.interactive_primitive_language import InteractivePrimitiveLanguageStudent
        from trainers.interactive_primitive_language import InteractivePrimitiveLanguageTrainer

        class RecordingInteractiveStudent(InteractivePrimitiveLanguageStudent):   # runs unchanged
            def init(self, states):
                super().init(states)
                self.cur = dict(is_eval=None, acts=[], refs=[], fh=[], instr_steps=[], desc_steps=[])

            def set_tasks(self, tasks, is_eval):
                self.cur["is_eval"] = is_eval
                super().set_tasks(tasks, is_eval)

            def _log(self, states, actions):
                feats = np.stack([np.asarray(s.features()) for s in states]).astype(np.float32)
                self.cur["fh"].append(feature_hash(feats))
                self.cur["acts"].append(list(actions))

            def act(self, states):
                actions = super().act(states)
                self._log(states, actions)
                return actions

            def instructed_act(self, states):
                actions = super().instructed_act(states)
                self._log(states, actions)
                return actions

            def set_instructions(self, instructions, is_eval):
                # called with the teacher's instructions before instructed_act, and again (from
                # receive) with the descriptions of the same timestep
                key = "desc_steps" if len(self.cur["instr_steps"]) > len(self.cur["desc_steps"]) else "instr_steps"
                self.cur[key].append([None if w is None else list(w) for w in instructions])
                super().set_instructions(instructions, is_eval)

            def learn(self):
                loss = super().learn()
                rec["iters"][-1]["loss"] = loss
                return loss

        from students.active_primitive_language import ActivePrimitiveLanguageStudent
        from trainers.active_primitive_language import ActivePrimitiveLanguageTrainer

        class RecordingActiveStudent(ActivePrimitiveLanguageStudent):           # runs unchanged
            # active_primitive_language.yaml: A/B only (no replay fixture) — rollout infos, losses,
            # the learned action map and the random-stream position must agree
            def init(self, states):
                super().init(states)
                self.cur = dict(is_eval=None, acts=[], refs=[], fh=[])

            def set_tasks(self, tasks, is_eval):
                self.cur["is_eval"] = is_eval
                super().set_tasks(tasks, is_eval)

            def learn(self):
                loss = super().learn()
                rec["iters"][-1]["loss"] = loss
                return loss

        base_trainer = {"primitive_language": PrimitiveLanguageTrainer,
                        "active_primitive_language": ActivePrimitiveLanguageTrainer,
                        "interactive_primitive_language": InteractivePrimitiveLanguageTrainer}.get(
                            experiment, ImitationTrainer)

        class RecordingTrainer(base_trainer):              # the trainer's code runs unchanged
            def do_rollout(self, batch, world, student, teacher, is_eval):
                info = super().do_rollout(batch, world, student, teacher, is_eval)
                entry = dict(ids=[item["id"] for item in batch], info=info, **student.cur)
                if is_eval:
                    rec["evals"][-1].append(entry)
                else:
                    rec["iters"].append(entry)
                return info

            def evaluate(self, dataset, world, student, teacher, save_traj=False):
                rec["evals"].append([])
                return super().evaluate(dataset, world, student, teacher, save_traj)

        config = make_config(data_dir, exp_dir, iters, log_every, experiment=experiment)
        datasets = ref_data.load(config)                   # also sets config.vocab (data/task.py:63)
        if which == "reference":
            world = ref_worlds.load(config)
            teacher = ref_teachers.load(config)
        else:
            import psketch_b200.teachers as my_teachers
            import psketch_b200.worlds as my_worlds
            world = my_worlds.load(config)                 # the swap of INTEGRATION.md §1 ...
            teacher = my_teachers.load(config)             # ... and nothing else
            if medium_oracle is not None:                  # no GPU in this container
                from test_facade_cpu import OracleBackend
                world._backend = OracleBackend(world, medium_oracle)
        assert config.student.model.input_size == 404 and config.student.model.n_actions == 6
        student = {"primitive_language": RecordingLanguageStudent,
                   "interactive_primitive_language": RecordingInteractiveStudent,
                   "active_primitive_language": RecordingActiveStudent}.get(experiment, RecordingStudent)(config)
        trainer = RecordingTrainer(config)
        torch.manual_seed(config.seed)
        config.random.seed(config.seed)
        for d in datasets.values():
            d.item_idx = 0
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):    # the language trainer print()s renders
            trainer.train(datasets, world, student, teacher)
        rec["action_map"] = dict(getattr(teacher, "student_action_map", {}))
        rec["random_state"] = config.random.get_state()[1][:8].tolist()
    shutil.rmtree(exp_dir, ignore_errors=True)
    return rec


def assert_same(a, b):
    assert len(a["iters"]) == len(b["iters"]) and len(a["evals"]) == len(b["evals"])
    assert a["action_map"] == b["action_map"] and a["random_state"] == b["random_state"]
    rollouts_a = a["iters"] + [e for ev in a["evals"] for e in ev]
    rollouts_b = b["iters"] + [e for ev in b["evals"] for e in ev]
    assert len(rollouts_a) == len(rollouts_b)
    for k, (x, y) in enumerate(zip(rollouts_a, rollouts_b)):
        assert x["ids"] == y["ids"], k
        assert x["acts"] == y["acts"], k
        assert x["refs"] == y["refs"], k
        assert [int(h) for h in x["fh"]] == [int(h) for h in y["fh"]], k
        assert x.get("loss") == y.get("loss"), (k, x.get("loss"), y.get("loss"))   # bit-equal floats
        for extra in ("phase2_acts", "instructions", "descriptions", "instr_steps", "desc_steps"):
            assert x.get(extra) == y.get(extra), (k, extra)
        assert [int(h) for h in x.get("phase2_fh", [])] == [int(h) for h in y.get("phase2_fh", [])], k
        ix, iy = x["info"], y["info"]
        assert ix["action_seqs"] == iy["action_seqs"], k
        assert [bool(v) for v in ix["success"]] == [bool(v) for v in iy["success"]], k
        assert ix["distances"] == iy["distances"], k
        assert ix["num_interactions"] == iy["num_interactions"] and ix["num_steps"] == iy["num_steps"], k


def pack(rec, splits):
    """Flat arrays.  Rollout r has B_r envs and T_r timesteps; 2-D blocks are padded to [R, 40, 32]."""
    id_to_idx = {}
    for split in ("train", "dev"):
        for i, v in enumerate(splits[split + "_inst_id"]):
            id_to_idx[(split, int(v))] = i
    rollouts = [("train", e) for e in rec["iters"]] + [("dev", e) for ev in rec["evals"] for e in ev]
    R, TM, BM = len(rollouts), 40, 32
    out = dict(
        n_train_iters=np.int32(len(rec["iters"])),
        eval_sizes=np.asarray([len(ev) for ev in rec["evals"]], np.int32),
        batch=np.full((R, BM), -1, np.int32), n_env=np.zeros(R, np.int32), n_t=np.zeros(R, np.int32),
        is_eval=np.zeros(R, np.uint8),
        acts=np.full((R, TM, BM), 255, np.uint8), refs=np.full((R, TM, BM), -2, np.int8),
        feat_hash=np.zeros((R, TM), np.uint64), loss=np.full(R, np.nan, np.float64),
        success=np.zeros((R, BM), np.uint8), seq_len=np.zeros((R, BM), np.uint8),
        distances=np.full((R, BM), -1, np.int16), n_dist=np.zeros(R, np.int32),
        num_interactions=np.zeros(R, np.int64), num_steps=np.zeros(R, np.int64))
    for r, (split, e) in enumerate(rollouts):
        B, T = len(e["ids"]), len(e["acts"])
        out["n_env"][r], out["n_t"][r], out["is_eval"][r] = B, T, int(e["is_eval"])
        for i, id_ in enumerate(e["ids"]):                 # 'instance_10561' -> row of the split
            out["batch"][r, i] = id_to_idx[(split, int(id_.split("_")[1]))]
        out["acts"][r, :T, :B] = np.asarray(e["acts"], np.int64).astype(np.uint8)     # -1 (terminated) -> 255
        if e["refs"]:
            out["refs"][r, :T, :B] = np.asarray(e["refs"], np.int8)
        out["feat_hash"][r, :T] = e["fh"]
        if "loss" in e:
            out["loss"][r] = e["loss"]
        info = e["info"]
        out["success"][r, :B] = [bool(v) for v in info["success"]]
        out["seq_len"][r, :B] = [len(s) for s in info["action_seqs"]]
        out["n_dist"][r] = len(info["distances"])
        out["distances"][r, :len(info["distances"])] = info["distances"]
        out["num_interactions"][r], out["num_steps"][r] = info["num_interactions"], info["num_steps"]
    if any("instr_steps" in e for _, e in rollouts):
        # interactive_primitive_language.yaml: per timestep the teacher's one-word instruction for
        # EVERY env (finished ones too, trainers/interactive_primitive_language.py:49-51) and its
        # one-word description of what each running env's action did (:61-66); 255 = none
        words = {"down": 0, "up": 1, "left": 2, "right": 3, "use": 4, "stop": 5}
        out["instr_steps"] = np.full((R, TM, BM), 255, np.uint8)
        out["desc_steps"] = np.full((R, TM, BM), 255, np.uint8)
        for r, (split, e) in enumerate(rollouts):
            for key in ("instr_steps", "desc_steps"):
                for t, row in enumerate(e[key]):
                    for i, w in enumerate(row):
                        if w is not None:
                            out[key][r, t, i] = words[w[0]]
        am = rec["action_map"]
        out["final_action_map"] = np.asarray([words.get(am.get(a), 255) for a in range(6)], np.uint8)
    if any("phase2_acts" in e for _, e in rollouts):
        # primitive_language.yaml: the second (greedy, instructed) decoding pass of training rollouts,
        # the instruction words handed to the student and the teacher's descriptions (word = action
        # index of teachers/primitive_language.py:20-32: down 0, up 1, left 2, right 3, use 4, stop 5)
        words = {"down": 0, "up": 1, "left": 2, "right": 3, "use": 4, "stop": 5}
        out["phase2_acts"] = np.full((R, TM, BM), 255, np.uint8)
        out["phase2_n_t"] = np.zeros(R, np.int32)
        out["phase2_feat_hash"] = np.zeros((R, TM), np.uint64)
        out["descriptions"] = np.full((R, BM, TM), 255, np.uint8)
        out["instructions"] = np.full((R, BM, TM), 255, np.uint8)
        for r, (split, e) in enumerate(rollouts):
            B = len(e["ids"])
            T2 = len(e["phase2_acts"])
            out["phase2_n_t"][r] = T2
            if T2:
                out["phase2_acts"][r, :T2, :B] = np.asarray(e["phase2_acts"], np.int64).astype(np.uint8)
                out["phase2_feat_hash"][r, :T2] = e["phase2_fh"]
            for i, ws in enumerate(e["instructions"]):
                out["instructions"][r, i, :len(ws)] = [words[w] for w in ws]
            for i, ws in enumerate(e["descriptions"] or []):
                out["descriptions"][r, i, :len(ws)] = [words[w] for w in ws]
        am = rec["action_map"]
        out["final_action_map"] = np.asarray([words.get(am.get(a), 255) for a in range(6)], np.uint8)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--log-every", type=int, default=15)
    ap.add_argument("--experiment", default="imitation",
                    choices=["imitation", "primitive_language", "interactive_primitive_language",
                             "active_primitive_language"])
    ap.add_argument("--no-golden", action="store_true", help="A/B comparison only, write no fixture")
    args = ap.parse_args()
    regen_dir = os.environ.get("PSK_REGEN_DIR", "/tmp/psk_data")
    data_dir = os.path.join(regen_dir, "config4_data")
    prepare_data(regen_dir, data_dir)
    splits = np.load(os.path.join(OUT, "craft_medium_splits.npz"))
    import time
    t0 = time.time()
    a = run("reference", data_dir, args.iters, args.log_every, experiment=args.experiment)
    t1 = time.time()
    print("reference world + teacher: %d train rollouts, %d evaluations, %.1f s; losses %s ..." %
          (len(a["iters"]), len(a["evals"]), t1 - t0, [round(e["loss"], 4) for e in a["iters"][:4]]))
    from oracle.craft_oracle import CraftOracle
    from psketch_b200.tables import CraftTables
    b = run("facade", data_dir, args.iters, args.log_every, medium_oracle=CraftOracle(CraftTables()),
            experiment=args.experiment)
    print("psketch_b200 world + teacher (oracle-backed backend): %.1f s" % (time.time() - t1))
    assert_same(a, b)
    print("IDENTICAL: batches, features, actions, teacher labels, losses, success, distances, eval trajectories")
    if args.no_golden or args.experiment == "active_primitive_language":
        print("A/B only: no fixture written; final losses", [round(e["loss"], 4) for e in a["iters"][-3:]])
        return
    out = pack(a, splits)
    path = os.path.join(OUT, "config4_%s.npz" % args.experiment)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;",
          "train success %.3f" % np.mean([np.mean([bool(v) for v in e["info"]["success"]]) for e in a["iters"]]),
          "dev success per evaluation", [float(np.mean([np.mean([bool(v) for v in e["info"]["success"]]) for e in ev])) for ev in a["evals"]])


if __name__ == "__main__":
    main()
