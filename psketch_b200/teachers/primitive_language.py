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
