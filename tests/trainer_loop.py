"""Test helper: the rollout protocol of the reference's trainer (trainers/imitation.py:18-101),
restated so that it can drive any (world, teacher) pair that has the reference's object API —
the CUDA-backed facade, or the pure-Python port — with a scripted student.  The expected outputs
in tests/golden/trainer_rollouts.npz come from the reference's own, unmodified
``ImitationTrainer.do_rollout`` (oracle/gen_golden.py --trainer)."""
import numpy as np

STOP = 5


class ScriptedStudent(object):
    def __init__(self, script):
        self.script = script
        self.features, self.received = [], []
        self.t = 0

    def act(self, states):
        self.features.append(np.stack([np.asarray(s.features()) for s in states]))
        row = [int(a) for a in self.script[self.t]]
        self.t += 1
        return row


def run_protocol(batch, world, teacher, student, is_eval, max_timesteps, mix_rate, rng):
    n = len(batch)
    tasks = [item["task"] for item in batch]
    states = [world.init_state(item["grid"], item["init_pos"]) for item in batch]
    clock = [max_timesteps] * n
    finished = [False] * n
    solved = [False] * n
    taken = [[] for _ in range(n)]
    interactions = steps = 0
    clone = None if is_eval else rng.binomial(1, mix_rate, size=n)
    while not all(finished):
        chosen = student.act(states)
        labels = [None] * n
        for i in range(n):
            if not is_eval:
                if finished[i]:
                    labels[i] = -1
                else:
                    labels[i] = teacher(tasks[i], states[i])
                    interactions += 1
                if clone[i]:
                    chosen[i] = labels[i]
            if not finished[i]:
                taken[i].append(chosen[i])
            clock[i] -= 1
            if chosen[i] == STOP or clock[i] <= 0:
                finished[i] = True
            if finished[i]:
                solved[i] = states[i].satisfies(tasks[i])
            else:
                states[i] = states[i].step(chosen[i])[1]
                steps += 0 if is_eval else 1
        if not is_eval:
            student.received.append(labels)
    gaps = []
    for i in range(n):
        if tasks[i].goal_name != "get":
            continue
        if solved[i]:
            gaps.append(0)
        else:
            probe = world.init_state(batch[i]["grid"], states[i].pos, states[i].dir)
            gaps.append(len(teacher.find_closest_resources(tasks[i], probe)[1]))
    return dict(action_seqs=taken, success=solved, distances=gaps, num_interactions=interactions,
                num_steps=steps)


def check_against_fixture(fx, make_batch, world, teacher):
    """``make_batch(instance_ids) -> list of dict(grid, init_pos, task)`` for the given world."""
    batch = make_batch(fx["inst"])
    for mode, is_eval in (("train", False), ("eval", True)):
        student = ScriptedStudent(fx["script"])
        rng = np.random.RandomState(int(fx["mix_seed"]))
        got = run_protocol(batch, world, teacher, student, is_eval, 40, float(fx["mix_rate"]), rng)
        want = fx[mode + "_action_seqs"]
        for i, seq in enumerate(got["action_seqs"]):
            L = int((want[i] != 255).sum())
            assert seq == want[i, :L].tolist(), (mode, i)
        assert [bool(v) for v in got["success"]] == fx[mode + "_success"].tolist(), mode
        assert got["distances"] == fx[mode + "_distances"].tolist(), mode
        assert got["num_interactions"] == int(fx[mode + "_num_interactions"]), mode
        assert got["num_steps"] == int(fx[mode + "_num_steps"]), mode
        feats = np.stack(student.features)
        assert feats.dtype == np.float64
        assert np.array_equal(feats, fx[mode + "_features"].astype(np.float64)), mode
        if not is_eval:
            assert np.array_equal(np.asarray(student.received, np.int16), fx["train_ref_actions"])


# ------------------------------------------------------------------------------------------------
# trainers/primitive_language.py:16-143, restated the same way (scripted student): two decoding
# passes from the SAME initial states in training (the states are persistent: the second pass
# restarts from init_states[:], :32,82), describe() over the state sequences of the first pass,
# success / distances from the states the last pass ended in.
class ScriptedLanguageStudent(object):
    def __init__(self, pass1, pass2):
        self.passes = [pass1, pass2]
        self.which = 0
        self.t = 0
        self.features = [[], []]
        self.descriptions = None

    def next_actions(self, states):
        self.features[self.which].append(np.stack([np.asarray(s.features()) for s in states]))
        row = [int(a) if a != 255 else -1 for a in self.passes[self.which][self.t]]
        self.t += 1
        return row

    def second_pass(self, descriptions):
        self.descriptions = descriptions
        self.which, self.t = 1, 0


def run_language_protocol(batch, world, teacher, student, is_eval, max_timesteps):
    n = len(batch)
    tasks = [item["task"] for item in batch]
    starts = [world.init_state(item["grid"], item["init_pos"]) for item in batch]
    hints = [teacher.instruct(world, item["ref_actions"]) for item in batch]
    starts[0].render()                                             # :28 (prints; exercises render)

    def decode(record_states):
        states = starts[:]
        clock = [max_timesteps] * n
        finished = [False] * n
        taken = [[] for _ in range(n)]
        visited = [[s] for s in states]
        steps = 0
        while not all(finished):
            chosen = student.next_actions(states)
            for i in range(n):
                if not finished[i]:
                    states[i] = states[i].step(chosen[i])[1]
                    taken[i].append(chosen[i])
                    if record_states:
                        visited[i].append(states[i])
                    steps += 1
                clock[i] -= 1
                if chosen[i] == STOP or clock[i] <= 0:
                    finished[i] = True
        return states, taken, visited, steps

    states, taken, visited, steps = decode(True)
    descriptions = None
    if not is_eval:
        descriptions = [teacher.describe(world, taken[i], visited[i]) for i in range(n)]
        student.second_pass(descriptions)
        states, taken, _, _ = decode(False)
    solved, gaps = [], []
    for i in range(n):
        solved.append(states[i].satisfies(tasks[i]))
        if tasks[i].goal_name == "get":
            if solved[i]:
                gaps.append(0)
            else:
                probe = world.init_state(batch[i]["grid"], states[i].pos, states[i].dir)
                gaps.append(len(teacher.find_closest_resources(tasks[i], probe)[1]))
    return dict(action_seqs=taken, success=solved, distances=gaps, instructions=hints,
                descriptions=descriptions,
                num_interactions=0 if is_eval else sum(len(h) for h in hints),
                num_steps=0 if is_eval else steps)


# ------------------------------------------------------------------------------------------------
# trainers/interactive_primitive_language.py:16-106, restated: per timestep the teacher gives a
# one-word instruction for EVERY env (also the finished ones: the teacher must tolerate
# post-terminal states), the student acts, and the teacher describes what each running env's
# action did from the (previous state, new state) pair.
class ScriptedInteractiveStudent(object):
    def __init__(self, script):
        self.script, self.t = script, 0
        self.features, self.instructions, self.descriptions = [], [], []

    def next_actions(self, states):
        self.features.append(np.stack([np.asarray(s.features()) for s in states]))
        row = [int(a) if a != 255 else -1 for a in self.script[self.t]]
        self.t += 1
        return row


def run_interactive_protocol(batch, world, teacher, student, is_eval, max_timesteps):
    n = len(batch)
    tasks = [item["task"] for item in batch]
    states = [world.init_state(item["grid"], item["init_pos"]) for item in batch]
    states[0].render()
    clock = [max_timesteps] * n
    finished = [False] * n
    taken = [[] for _ in range(n)]
    interactions = steps = 0
    told = [None] * n
    said = [None] * n
    while not all(finished):
        if not is_eval:
            for i in range(n):
                told[i] = teacher(tasks[i], states[i])
                interactions += 0 if finished[i] else 1
            student.instructions.append([None if w is None else list(w) for w in told])
        chosen = student.next_actions(states)
        before = states[:]
        for i in range(n):
            if not finished[i]:
                states[i] = states[i].step(chosen[i])[1]
                taken[i].append(chosen[i])
                steps += 0 if is_eval else 1
                said[i] = teacher.describe(world, [chosen[i]], [before[i], states[i]])
        student.descriptions.append([None if w is None else list(w) for w in said])
        for i in range(n):
            clock[i] -= 1
            if chosen[i] == STOP or clock[i] <= 0:
                finished[i] = True
    solved, gaps = [], []
    for i in range(n):
        solved.append(states[i].satisfies(tasks[i]))
        if tasks[i].goal_name == "get":
            if solved[i]:
                gaps.append(0)
            else:
                probe = world.init_state(batch[i]["grid"], states[i].pos, states[i].dir)
                gaps.append(len(teacher.find_closest_resources(tasks[i], probe)[1]))
    return dict(action_seqs=taken, success=solved, distances=gaps, num_interactions=interactions,
                num_steps=steps)
