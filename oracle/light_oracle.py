"""ctypes front-end of light_oracle.c — TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborc_light.so")


class _Scen(ctypes.Structure):
    _fields_ = [("bw", ctypes.c_int32), ("bh", ctypes.c_int32), ("n_doors", ctypes.c_int32),
                ("n_keys", ctypes.c_int32), ("goal_rx", ctypes.c_int32), ("goal_ry", ctypes.c_int32),
                ("walls", ctypes.c_uint8 * (32 * 32)), ("doors", ctypes.c_int32 * 16),
                ("keys", ctypes.c_int32 * 32)]


def build(force=False):
    src = os.path.join(_HERE, "light_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


class LightOracle(object):
    def __init__(self, walls, board, doors, n_doors, keys, n_keys, goal_room):
        """Arrays of S scenarios in the layout of tests/golden/light_states.npz."""
        self.lib = ctypes.CDLL(build())
        assert self.lib.orc_light_scen_size() == ctypes.sizeof(_Scen)
        S = len(board)
        self.scen = (_Scen * S)()
        for i in range(S):
            c = self.scen[i]
            c.bw, c.bh = int(board[i][0]), int(board[i][1])
            c.n_doors, c.n_keys = int(n_doors[i]), int(n_keys[i])
            c.goal_rx, c.goal_ry = int(goal_room[i][0]), int(goal_room[i][1])
            w = np.ones((32, 32), np.uint8)
            w[:31, :31] = walls[i]
            ctypes.memmove(c.walls, np.ascontiguousarray(w).ctypes.data, 1024)
            for j in range(c.n_doors):
                c.doors[2 * j], c.doors[2 * j + 1] = int(doors[i][j][0]), int(doors[i][j][1])
            for j in range(c.n_keys):
                for t in range(4):
                    c.keys[4 * j + t] = int(keys[i][j][t])
        for name in ("orc_light_batch", "orc_light_batch_expert"):
            getattr(self.lib, name).restype = None

    def run(self, scen_idx, state, action=None):
        scen_idx = np.ascontiguousarray(scen_idx, np.int32)
        state = np.ascontiguousarray(state, np.int32)
        n = len(scen_idx)
        feat = np.empty((n, 12), np.float32)
        sat = np.empty(n, np.int32)
        out = np.empty((n, 3), np.int32) if action is not None else None
        act = np.ascontiguousarray(action, np.int32) if action is not None else None
        P = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None
        self.lib.orc_light_batch(self.scen, P(scen_idx, ctypes.c_int32), ctypes.c_int64(n),
                                 P(state, ctypes.c_int32), P(act, ctypes.c_int32),
                                 P(out, ctypes.c_int32), P(feat, ctypes.c_float), P(sat, ctypes.c_int32))
        return feat, sat, out

    def expert(self, scen_idx, state):
        scen_idx = np.ascontiguousarray(scen_idx, np.int32)
        state = np.ascontiguousarray(state, np.int32)
        n = len(scen_idx)
        act = np.empty(n, np.int32)
        dist = np.empty(n, np.int32)
        P = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))
        self.lib.orc_light_batch_expert(self.scen, P(scen_idx, ctypes.c_int32), ctypes.c_int64(n),
                                        P(state, ctypes.c_int32), P(act, ctypes.c_int32),
                                        P(dist, ctypes.c_int32))
        return act, dist
