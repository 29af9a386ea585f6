"""ctypes front-end of ``craft_oracle.c`` — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every method is a thin batched call into the C restatement; the semantics and the reference
citations live in ``craft_oracle.c``.  Arrays are numpy; the grid is ``u8[N, W*H]`` kind ids
(index ``x*H + y``), inventory ``i32[N, K]``, pos ``i32[N, 2]``, dir ``i32[N]``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborc_craft.so")

MAX_KINDS, MAX_RECIPES, MAX_TASKS, MAX_NODES = 32, 16, 32, 16


class _Tables(ctypes.Structure):
    _fields_ = [
        ("W", ctypes.c_int32), ("H", ctypes.c_int32), ("K", ctypes.c_int32),
        ("win_w", ctypes.c_int32), ("win_h", ctypes.c_int32),
        ("n_recipes", ctypes.c_int32),
        ("water", ctypes.c_int32), ("stone", ctypes.c_int32),
        ("bridge", ctypes.c_int32), ("axe", ctypes.c_int32),
        ("kind_class", ctypes.c_uint8 * MAX_KINDS),
        ("recipes", ctypes.c_uint8 * (MAX_RECIPES * 8)),
        ("task_nodes", ctypes.c_uint8 * (MAX_TASKS * MAX_NODES * 4)),
        ("task_len", ctypes.c_uint8 * MAX_TASKS),
    ]


def build(force=False):
    src = os.path.join(_HERE, "craft_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", _SO, src])
    return _SO


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


class CraftOracle(object):
    def __init__(self, tables):
        """``tables``: a psketch_b200.tables.CraftTables (host-side table compilation only)."""
        self.lib = ctypes.CDLL(build())
        assert self.lib.orc_tables_size() == ctypes.sizeof(_Tables)
        t = _Tables()
        t.W, t.H, t.K, t.win_w, t.win_h = tables.W, tables.H, tables.K, tables.win_w, tables.win_h
        t.n_recipes = tables.n_recipes
        t.water, t.stone = tables.water_kind, tables.stone_kind
        t.bridge, t.axe = tables.bridge_kind, tables.axe_kind
        ctypes.memmove(t.kind_class, _u8(tables.kind_class).ctypes.data, MAX_KINDS)
        ctypes.memmove(t.recipes, _u8(tables.recipes).ctypes.data, MAX_RECIPES * 8)
        ctypes.memmove(t.task_nodes, _u8(tables.task_nodes).ctypes.data, MAX_TASKS * MAX_NODES * 4)
        ctypes.memmove(t.task_len, _u8(tables.task_len).ctypes.data, MAX_TASKS)
        self.t = t
        self.tables = tables
        self.W, self.H, self.K = tables.W, tables.H, tables.K
        self.C = self.W * self.H
        self.n_features = tables.n_features
        for name in ("orc_batch_features", "orc_batch_step", "orc_batch_expert",
                     "orc_batch_satisfies", "orc_batch_find_closest", "orc_rollout"):
            getattr(self.lib, name).restype = None

    def set_threads(self, n):
        """OpenMP thread count of the batched calls (torchrun exports OMP_NUM_THREADS=1)."""
        try:
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
        except OSError:
            os.environ["OMP_NUM_THREADS"] = str(n)

    def features(self, grid, inv, pos, dirs):
        grid, inv, pos, dirs = _u8(grid), _i32(inv), _i32(pos), _i32(dirs)
        n = grid.shape[0]
        out = np.empty((n, self.n_features), np.float32)
        self.lib.orc_batch_features(ctypes.byref(self.t), ctypes.c_int64(n),
                                    _p(grid, ctypes.c_uint8), _p(inv, ctypes.c_int32),
                                    _p(pos, ctypes.c_int32), _p(dirs, ctypes.c_int32),
                                    _p(out, ctypes.c_float))
        return out

    def step(self, grid, inv, pos, dirs, action):
        """Returns new (grid, inv, pos, dir, status); status -1 = the reference raises."""
        grid, inv, pos, dirs = _u8(grid).copy(), _i32(inv).copy(), _i32(pos).copy(), _i32(dirs).copy()
        action = _i32(action)
        n = grid.shape[0]
        status = np.empty(n, np.int32)
        self.lib.orc_batch_step(ctypes.byref(self.t), ctypes.c_int64(n),
                                _p(grid, ctypes.c_uint8), _p(inv, ctypes.c_int32),
                                _p(pos, ctypes.c_int32), _p(dirs, ctypes.c_int32),
                                _p(action, ctypes.c_int32), _p(status, ctypes.c_int32))
        return grid, inv, pos, dirs, status

    def expert(self, grid, inv, pos, dirs, task):
        """Returns (action, dist, status) — see orc_expert."""
        grid, inv, pos, dirs, task = _u8(grid), _i32(inv), _i32(pos), _i32(dirs), _i32(task)
        n = grid.shape[0]
        action = np.empty(n, np.int32)
        dist = np.empty(n, np.int32)
        status = np.empty(n, np.int32)
        self.lib.orc_batch_expert(ctypes.byref(self.t), ctypes.c_int64(n),
                                  _p(grid, ctypes.c_uint8), _p(inv, ctypes.c_int32),
                                  _p(pos, ctypes.c_int32), _p(dirs, ctypes.c_int32),
                                  _p(task, ctypes.c_int32), _p(action, ctypes.c_int32),
                                  _p(dist, ctypes.c_int32), _p(status, ctypes.c_int32))
        return action, dist, status

    def satisfies(self, grid, inv, pos, dirs, task):
        """1 True, 0 False, 2 None."""
        grid, inv, pos, dirs, task = _u8(grid), _i32(inv), _i32(pos), _i32(dirs), _i32(task)
        n = grid.shape[0]
        out = np.empty(n, np.int32)
        self.lib.orc_batch_satisfies(ctypes.byref(self.t), ctypes.c_int64(n),
                                     _p(grid, ctypes.c_uint8), _p(inv, ctypes.c_int32),
                                     _p(pos, ctypes.c_int32), _p(dirs, ctypes.c_int32),
                                     _p(task, ctypes.c_int32), _p(out, ctypes.c_int32))
        return out

    def find_closest(self, grid, pos, dirs, kind, seq_cap=64):
        """Returns (goal i32[N,2], length i32[N] (-1 = None), status, seq u8[N, seq_cap])."""
        grid, pos, dirs, kind = _u8(grid), _i32(pos), _i32(dirs), _i32(kind)
        n = grid.shape[0]
        goal = np.empty((n, 2), np.int32)
        length = np.empty(n, np.int32)
        status = np.empty(n, np.int32)
        seq = np.full((n, seq_cap), 255, np.uint8)
        self.lib.orc_batch_find_closest(ctypes.byref(self.t), ctypes.c_int64(n),
                                        _p(grid, ctypes.c_uint8), _p(pos, ctypes.c_int32),
                                        _p(dirs, ctypes.c_int32), _p(kind, ctypes.c_int32),
                                        _p(goal, ctypes.c_int32), _p(length, ctypes.c_int32),
                                        _p(status, ctypes.c_int32), _p(seq, ctypes.c_uint8),
                                        ctypes.c_int(seq_cap))
        return goal, length, status, seq

    def rollout(self, ticks, max_timesteps, init_grid, init_pos, task, state=None,
                want_features=False):
        """Runs ``ticks`` rollout ticks per env (see orc_rollout).  ``state`` is a dict of the
        working arrays (created from the init state when None) and is updated in place.
        Returns (state, stats[4], features or None, last_action)."""
        init_grid, init_pos, task = _u8(init_grid), _i32(init_pos), _i32(task)
        n = init_grid.shape[0]
        if state is None:
            state = dict(grid=init_grid.copy(), inv=np.zeros((n, self.K), np.int32),
                         pos=init_pos.copy(), dir=np.zeros(n, np.int32),
                         timer=np.full(n, max_timesteps, np.int32))
        feats = np.empty((n, self.n_features), np.float32) if want_features else None
        action = np.empty(n, np.int32)
        stats = np.zeros(4, np.int64)
        self.lib.orc_rollout(ctypes.byref(self.t), ctypes.c_int64(n), ctypes.c_int(ticks),
                             ctypes.c_int(max_timesteps), _p(init_grid, ctypes.c_uint8),
                             _p(init_pos, ctypes.c_int32), _p(task, ctypes.c_int32),
                             _p(state["grid"], ctypes.c_uint8), _p(state["inv"], ctypes.c_int32),
                             _p(state["pos"], ctypes.c_int32), _p(state["dir"], ctypes.c_int32),
                             _p(state["timer"], ctypes.c_int32),
                             _p(feats, ctypes.c_float) if feats is not None else None,
                             _p(action, ctypes.c_int32), _p(stats, ctypes.c_int64))
        return state, stats, feats, action
