"""Language teachers (host word logic): same outputs and same random-stream consumption as the
reference's teachers/primitive_language.py when that checkout is present, fixed expectations
otherwise.  CPU only (fake states: ``describe`` reads just .pos and .inventory)."""
import os
import sys
import types

import numpy as np
import pytest


class _S(object):
    def __init__(self, pos, inv):
        self.pos, self.inventory = pos, np.asarray(inv, float)


def _world():
    names = ("DOWN", "UP", "LEFT", "RIGHT", "USE", "STOP")
    acts = types.SimpleNamespace(**{n: types.SimpleNamespace(index=i) for i, n in enumerate(names)})
    return types.SimpleNamespace(actions=acts, action_space=[getattr(acts, n) for n in names])


def _episodes(rng, n=40):
    eps = []
    perm = rng.permutation(6)            # the student's private action ids
    delta = {0: (0, -1), 1: (0, 1), 2: (-1, 0), 3: (1, 0)}
    for _ in range(n):
        L = rng.randint(1, 9)
        pos, inv = (3, 3), [0, 0, 0]
        states, actions = [_S(pos, inv)], []
        for t in range(L):
            real = rng.randint(0, 6)
            actions.append(int(perm[real]))
            if real < 4 and rng.rand() < 0.7:
                pos = (pos[0] + delta[real][0], pos[1] + delta[real][1])
            elif real == 4 and rng.rand() < 0.5:
                inv = [inv[0] + 1, inv[1], inv[2]]
            states.append(_S(pos, list(inv)))
        eps.append((actions, states))
    return eps


def test_instruct_and_describe():
    from psketch_b200.teachers import PrimitiveLanguageTeacher
    from psketch_b200.teachers.primitive_language import instruct_batch
    world = _world()
    cfg = types.SimpleNamespace(random=np.random.RandomState(5))
    t = PrimitiveLanguageTeacher(cfg)
    assert t.instruct(world, [1, 0, 2, 3, 4, 5]) == ["up", "down", "left", "right", "use", "stop"]
    with pytest.raises(AssertionError):
        t.instruct(world, [6])
    ids = instruct_batch(np.asarray([[1, 4, 5, 255]], np.uint8))
    assert ids.tolist() == [[2, 5, 6, 0]]
    eps = _episodes(np.random.RandomState(1))
    mine = [t.describe(world, a, s) for a, s in eps]
    ref_root = "/root/reference"
    if not os.path.isdir(os.path.join(ref_root, "teachers")):
        assert all(len(m) == len(a) for m, (a, _) in zip(mine, eps))
        assert len(t.student_action_map) >= 4
        return
    sys.path.insert(0, ref_root)
    try:
        import importlib
        ref_mod = importlib.import_module("teachers.primitive_language")
    finally:
        sys.path.remove(ref_root)
    cfg2 = types.SimpleNamespace(random=np.random.RandomState(5))
    r = ref_mod.PrimitiveLanguageTeacher(cfg2)
    theirs = [r.describe(world, a, s) for a, s in eps]
    assert mine == theirs
    assert t.student_action_map == r.student_action_map
    assert cfg.random.randint(1 << 30) == cfg2.random.randint(1 << 30)     # same stream position
