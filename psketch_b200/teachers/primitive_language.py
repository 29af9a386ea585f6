"""Language teachers: host-side word logic on top of the CUDA demonstration
teacher, with the reference's interfaces.

    PrimitiveLanguageTeacher.instruct(world, action_seq) -> words      primitive_language.py:17-33
    PrimitiveLanguageTeacher.describe(world, action_seq, state_seq)    primitive_language.py:35-90
    InteractivePrimitiveLanguageTeacher(task, state) -> [word]         interactive_...py:19-37

``describe`` keeps the reference's stateful ``student_action_map`` and draws from
``config.random`` exactly when the reference does (one ``choice`` per undecidable step), so a
training run consumes the same random stream.  ``instruct_batch`` is the tensor form of
``instruct`` for whole rollouts.
"""
import numpy as np

from .demonstration import DemonstrationTeacher

# word of each action index: DOWN, UP, LEFT, RIGHT, USE, STOP (worlds/craft.py:25-30)
ACTION_WORDS = ("down", "up", "left", "right", "use", "stop")
_MOVE_WORD = {(0, -1): "down", (0, 1): "up", (-1, 0): "left", (1, 0): "right"}


class PrimitiveLanguageTeacher(DemonstrationTeacher):
    def __init__(self, config=None):
        super(PrimitiveLanguageTeacher, self).__init__(config)
        self.student_action_map = {}
        self.random = getattr(config, "random", None) or np.random.RandomState(0)

    def action_to_word(self, action, world=None):
        if not 0 <= int(action) < len(ACTION_WORDS):
            raise AssertionError("unknown action %r" % (action,))
        return ACTION_WORDS[int(action)]

    def instruct(self, world, action_seq):
        return [self.action_to_word(a, world) for a in action_seq]

    def describe(self, world, action_seq, state_seq):
        """Names what the student's actions did, learning the student's private action ids from
        the observed position / inventory changes."""
        words = []
        known = self.student_action_map
        for i, action in enumerate(action_seq):
            word = known.get(action)
            if word is None and len(known) == len(world.action_space) - 1:
                # five of six ids are known: the remaining word belongs to this id
                used = list(known.values())
                for w in ("up", "down", "left", "right", "use", "stop"):
                    if w not in used:
                        known[action] = word = w
                        break
            if word is None:
                here, before = state_seq[i + 1], state_seq[i]
                move = (here.pos[0] - before.pos[0], here.pos[1] - before.pos[1])
                if move == (0, 0):
                    if (np.asarray(here.inventory) != np.asarray(before.inventory)).any():
                        known[action] = word = "use"
                    else:
                        options = ["down", "up", "left", "right", "use"]
                        if i + 1 == len(state_seq) - 1:
                            options.append("stop")
                        word = self.random.choice(options)
                else:
                    if move in _MOVE_WORD:
                        known[action] = _MOVE_WORD[move]
                    word = known[action]
            assert word is not None
            words.append(word)
        return words


    # ---- describe() for whole batches of rollouts, as tensor operations ----------------------------
    def describe_batch(self, action_seqs, agent_seq, lengths, n_kinds=None):
        """``describe`` for N rollouts at once, on whatever device the tensors live on.

        action_seqs  int tensor [N, L]: the student's action ids, rollout i uses the first lengths[i]
        agent_seq    u8 tensor [L + 1, N, 32]: the env's agent records (VecCraft.agent) before every
                     action and after the last one — position bytes 24/25, inventory bytes 0..K-1
        lengths      int tensor [N]
        Returns i64 [N, L]: index into ACTION_WORDS per step, -1 beyond a rollout's length.

        Same words, same ``student_action_map`` and the same draws from ``self.random`` as calling
        ``describe`` rollout by rollout (teachers/primitive_language.py:35-90).  The map only ever
        grows and holds at most 6 entries, so the sequential semantics are resolved by a host loop
        that runs once per NEW map entry (at most 6 times in a teacher's life): each pass finds, with
        tensor operations over all N x L steps, the first step in rollout-major order that teaches a
        new action id; the undecidable steps before it draw their random words in one vectorised
        ``randint`` (which consumes the legacy stream exactly like the per-step ``choice`` calls).
        Once the map is complete a call is a single table lookup."""
        import torch
        acts = torch.as_tensor(action_seqs).long()
        dev = acts.device
        N, L = acts.shape
        lengths = torch.as_tensor(lengths).to(dev).long()
        ag = torch.as_tensor(agent_seq).to(dev)
        K = n_kinds if n_kinds is not None else 24
        pos = ag[:, :, 24:26].to(torch.int16)
        d = (pos[1:] - pos[:-1]).permute(1, 0, 2)                           # [N, L, 2]
        inv_changed = (ag[1:, :, :K] != ag[:-1, :, :K]).any(dim=2).t()      # [N, L]
        dx, dy = d[..., 0], d[..., 1]
        move = torch.full((N, L), -1, dtype=torch.long, device=dev)         # ACTION_WORDS index of a move
        move = torch.where((dx == 0) & (dy == -1), torch.zeros_like(move), move)
        move = torch.where((dx == 0) & (dy == 1), torch.ones_like(move), move)
        move = torch.where((dx == -1) & (dy == 0), torch.full_like(move, 2), move)
        move = torch.where((dx == 1) & (dy == 0), torch.full_like(move, 3), move)
        still = (dx == 0) & (dy == 0)
        teaches = torch.where(still, torch.where(inv_changed, torch.full_like(move, 4), torch.full_like(move, -1)),
                              move)                                         # word a step reveals, or -1
        t_idx = torch.arange(L, device=dev).unsqueeze(0)
        valid = t_idx < lengths.unsqueeze(1)
        is_last = t_idx == (lengths.unsqueeze(1) - 1)
        flat = (torch.arange(N, device=dev).unsqueeze(1) * L + t_idx)      # rollout-major step order
        BIG = N * L
        words = torch.full((N, L), -1, dtype=torch.long, device=dev)
        start = 0                                                           # steps before `start` are final
        known = self.student_action_map
        n_actions = len(ACTION_WORDS)
        while True:
            lut = torch.full((256,), -1, dtype=torch.long)
            for a, w in known.items():
                lut[int(a) & 255] = ACTION_WORDS.index(w)
            w_known = lut.to(dev)[acts & 255]
            todo = valid & (flat >= start)
            words = torch.where(todo & (w_known >= 0), w_known, words)
            unknown = todo & (w_known < 0)
            if len(known) >= n_actions or not bool(unknown.any()):
                break
            if len(known) == n_actions - 1:
                # the one missing word belongs to the first unknown id (primitive_language.py:47-54)
                k = int(torch.where(unknown, flat, torch.full_like(flat, BIG)).min())
                missing = [w for w in ("up", "down", "left", "right", "use", "stop") if w not in known.values()][0]
                known[int(acts.reshape(-1)[k])] = missing
                start = k
                continue
            first = torch.where(unknown & (teaches >= 0), flat, torch.full_like(flat, BIG)).min()
            k = int(first)
            guess = unknown & (teaches < 0) & (flat < k)                    # undecidable: random word
            if bool(guess.any()):
                n_cand = (5 + is_last.long())[guess].cpu().numpy()          # 'stop' only for the last action
                draws = self.random.randint(0, n_cand)                      # == one choice() per step, in order
                words[guess] = torch.from_numpy(np.asarray(draws, np.int64)).to(dev)
            if k >= BIG:
                break
            known[int(acts.reshape(-1)[k])] = ACTION_WORDS[int(teaches.reshape(-1)[k])]
            start = k
        return words


class InteractivePrimitiveLanguageTeacher(PrimitiveLanguageTeacher):
    def __init__(self, config=None):
        super(InteractivePrimitiveLanguageTeacher, self).__init__(config)
        self.demonstration_teacher = DemonstrationTeacher(config)

    def __call__(self, task, state):
        return [self.action_to_word(self.demonstration_teacher(task, state), state.world)]


def instruct_batch(action_seqs, vocab=None):
    """u8[N, L] action ids (255 = padding) -> i64[N, L] word ids (0 = padding): the batched LUT
    form of ``instruct``.  ``vocab`` maps a word to its id (TaskManager.vocab); defaults to
    1 + action index."""
    import torch
    lut = torch.zeros(256, dtype=torch.long)
    for a, w in enumerate(ACTION_WORDS):
        lut[a] = (vocab[w] if vocab is not None else a + 1)
    acts = torch.as_tensor(action_seqs)
    return lut.to(acts.device)[acts.long()]
