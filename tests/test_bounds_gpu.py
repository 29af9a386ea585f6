"""Out-of-bounds write check of our own (compute-sanitizer is closed on the GPU pool): every
state / output buffer handed to the C ABI is carved out of one arena pre-filled with a canary
byte, with guard bands on both sides; after each entry point has run on ragged batch sizes the
guard bands must be untouched and the results must equal those of an ordinary (torch-allocated)
batch.  Feature buffers are carved both 16-byte aligned and at a 4-byte offset, which forces the
scalar store path of the feature pipeline."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CANARY = 0xA5
GUARD = 512


class Arena(object):
    def __init__(self, nbytes, device):
        self.buf = torch.full((nbytes,), CANARY, dtype=torch.uint8, device=device)
        self.base = self.buf.data_ptr()
        self.used = torch.zeros(nbytes, dtype=torch.bool, device=device)
        self.top = GUARD

    def carve(self, shape, dtype, align=256, skew=0):
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        start = self.top
        start += (-(self.base + start)) % align
        start += skew
        assert start + nbytes + GUARD <= self.buf.numel(), "arena too small"
        self.top = start + nbytes + GUARD
        self.used[start:start + nbytes] = True
        view = self.buf[start:start + nbytes].view(dtype).view(*shape)
        assert view.data_ptr() == self.base + start
        return view

    def guards_intact(self):
        return bool((self.buf[~self.used] == CANARY).all())


def _carved_env(tables, arena, ref):
    """A VecCraft whose state and episode tables live inside the arena, copied from ``ref``."""
    from psketch_b200.vec import VecCraft
    env = VecCraft(tables, ref.n, max_timesteps=ref.max_timesteps)
    for name in ("grid", "agent", "scen_grid", "scen_idx", "init_agent"):
        src = getattr(ref, name)
        dst = arena.carve(tuple(src.shape), src.dtype)
        dst.copy_(src)
        setattr(env, name, dst)
    env.stats = arena.carve((4,), torch.int64)
    env.stats.zero_()
    env.err_flags = arena.carve((1,), torch.int32)
    env.err_flags.zero_()
    return env


def _instances(splits, n, seed):
    rng = np.random.RandomState(seed)
    pick = rng.randint(0, len(splits["train_inst_env"]), size=n)
    return (splits["train_grids"], splits["train_inst_env"][pick], splits["train_inst_pos"][pick],
            splits["train_inst_task"][pick])


def _same(a, b):
    return torch.equal(a.cpu(), b.cpu())


@pytest.mark.parametrize("n", [1, 65, 4099])
@pytest.mark.parametrize("skew", [0, 4])
def test_medium_entry_points_stay_inside_their_buffers(n, skew, splits, medium_tables):
    from psketch_b200.vec import VecCraft
    tables = medium_tables
    grids, env_idx, pos, task = _instances(splits, n, seed=n + skew)
    ref = VecCraft.from_instances(tables, grids, env_idx, pos, task)
    nf = ref.n_features
    T, R = 5, 2
    arena = Arena(320 * 1024 + n * (2 * 96 + 64 + (R + 11) * nf * 4 + 64 * 24), ref.device)
    env = _carved_env(tables, arena, ref)

    # a few teacher ticks first so that inventories / cleared cells occur
    for _ in range(6):
        a = ref.tick(want_features=False)
        b = env.tick(want_features=False, out=dict(
            expert=arena.carve((n,), torch.uint8), done=arena.carve((n,), torch.uint8),
            success=arena.carve((n,), torch.uint8)))
        assert _same(a["expert"], b["expert"])
    assert arena.guards_intact()

    # single-op entry points
    for impl in (0, 1, 2):
        f = env.features(out=arena.carve((n, nf), torch.float32, skew=skew), impl=impl)
        assert _same(f, ref.features()), impl
        assert arena.guards_intact(), "features impl %d" % impl
    acts = ref.expert()
    out = arena.carve((n,), torch.uint8, skew=skew % 3)
    assert _same(env.expert(out=out), acts)
    assert _same(env.satisfies(), ref.satisfies())
    kind = torch.full((n,), 9, dtype=torch.uint8)
    g0, l0, s0 = ref.find_closest(kind, seq_cap=40)
    g1, l1, s1 = env.find_closest(kind, seq_cap=40)
    found = (l0 >= 0)
    assert _same(l0, l1) and _same(g0[found], g1[found]) and _same(s0[found], s1[found])
    ra = env.random_actions(t=3, out=arena.carve((n,), torch.uint8, skew=skew % 3))
    assert _same(ra, ref.random_actions(t=3))
    ref.step(acts)
    env.step(acts)
    assert _same(env.grid, ref.grid) and _same(env.agent, ref.agent)
    assert arena.guards_intact()

    # fused and unfused ticks with features, student-supplied actions
    # (tick / rollout require 16-byte aligned feature buffers and say so)
    from psketch_b200._lib import PskError
    if skew:
        with pytest.raises(PskError):
            env.tick(features_out=arena.carve((n, nf), torch.float32, skew=skew))
        with pytest.raises(PskError):
            env.rollout(2, features_out=arena.carve((1, n, nf), torch.float32, skew=skew))
        assert arena.guards_intact()
    for fused in (True, False):
        fo = arena.carve((n, nf), torch.float32)
        o = dict(expert=arena.carve((n,), torch.uint8, skew=skew % 3),
                 done=arena.carve((n,), torch.uint8, skew=skew % 3),
                 success=arena.carve((n,), torch.uint8, skew=skew % 3))
        b = env.tick(actions=ra, features_out=fo, fused=fused, out=o)
        a = ref.tick(actions=ra, fused=fused)
        for k in ("expert", "done", "success", "features"):
            assert _same(a[k], b[k]), (fused, k)
        assert arena.guards_intact(), "tick fused=%s" % fused

    # round-2 entry points: u8 feature frame, step-then-observe tick, block of random actions
    f8 = env.features_u8(out=arena.carve((n, nf), torch.uint8, skew=skew % 3))
    assert _same(f8, ref.features_u8()) and _same(f8.float(), ref.features())
    assert arena.guards_intact(), "features_u8"
    fo = arena.carve((n, nf), torch.float32)
    o = dict(expert=arena.carve((n,), torch.uint8, skew=skew % 3),
             done=arena.carve((n,), torch.uint8, skew=skew % 3),
             success=arena.carve((n,), torch.uint8, skew=skew % 3))
    for acts_in in (None, ra):
        b = env.tick(actions=acts_in, features_out=fo, out=o, advance_first=True)
        a = ref.tick(actions=acts_in, advance_first=True)
        for k in ("expert", "done", "success", "features"):
            assert _same(a[k], b[k]), ("advance_first", k)
        assert _same(env.grid, ref.grid) and _same(env.agent, ref.agent)
        assert arena.guards_intact(), "tick advance_first"
    blk = env.random_actions(t=11, ticks=3, out=arena.carve((3, n), torch.uint8, skew=skew % 3))
    assert _same(blk, ref.random_actions(t=11, ticks=3))
    assert arena.guards_intact(), "random_actions block"

    # multi-tick rollout with a feature ring (frames are n*nf*4 bytes apart: not 16-byte
    # multiples for odd n, which is the alignment case the vector path has to refuse)
    ring = arena.carve((R, n, nf), torch.float32)
    o = dict(expert=arena.carve((T, n), torch.uint8, skew=skew % 3),
             done=arena.carve((T, n), torch.uint8, skew=skew % 3),
             success=arena.carve((T, n), torch.uint8, skew=skew % 3))
    b = env.rollout(T, features_out=ring, out=o)
    ring_ref = torch.empty((R, n, nf), dtype=torch.float32, device=ref.device)
    a = ref.rollout(T, features_out=ring_ref)
    for k in ("expert", "done", "success"):
        assert _same(a[k], b[k]), k
    assert _same(ring, ring_ref)
    assert _same(env.grid, ref.grid) and _same(env.agent, ref.agent)
    assert _same(env.stats, ref.stats)
    assert arena.guards_intact(), "rollout"
    env.reset()
    ref.reset()
    assert _same(env.grid, ref.grid) and _same(env.agent, ref.agent)
    assert arena.guards_intact(), "reset"
    env.check_errors()


@pytest.mark.parametrize("n", [3, 1031])
def test_large_and_rows_entry_points_stay_inside_their_buffers(n, large_tables, large_states):
    """craft_large (128-bit boards, 1,076 features) and a 16x16 grid (row-per-lane teacher)."""
    from psketch_b200.tables import CraftTables
    from psketch_b200.vec import VecCraft
    S = large_states
    cases = [(large_tables, S["grid"][:n], S["inv"][:n], S["pos"][:n], S["dir"][:n])]
    t16 = CraftTables(world_config=dict(WIDTH=16, HEIGHT=16, WINDOW_WIDTH=3, WINDOW_HEIGHT=3,
                                        N_WORKSHOPS=3, N_PRIMITIVES=4, N_WORLDS=1))
    rng = np.random.RandomState(n)
    g16 = np.zeros((n, 16, 16), np.uint8)
    g16[:, 0, :] = g16[:, 15, :] = g16[:, :, 0] = g16[:, :, 15] = 1
    for i in range(n):
        for kind in (2, 3, 4, 7, 7, 8, 8, 9, 9, 6, 6, 6, 5):
            x, y = rng.randint(1, 15, size=2)
            g16[i, x, y] = kind
    p16 = np.zeros((n, 2), np.int32)
    for i in range(n):
        free = np.argwhere(g16[i] == 0)
        p16[i] = free[rng.randint(len(free))]
    cases.append((t16, g16.reshape(n, 256), np.zeros((n, t16.K), np.int32), p16,
                  rng.randint(0, 4, size=n)))
    for tables, grid, inv, pos, dirs in cases:
        task = np.random.RandomState(7).choice([13, 14, 15, 19, 20, 21], size=n)
        ref = VecCraft.from_states(tables, grid, inv, pos, dirs, task=task)
        nf = ref.n_features
        arena = Arena(256 * 1024 + n * (3 * (ref.cell_stride + 32) + 6 * nf * 4 + 64 * 16), ref.device)
        env = _carved_env(tables, arena, ref)
        for impl in (0, 2):
            f = env.features(out=arena.carve((n, nf), torch.float32), impl=impl)
            assert _same(f, ref.features())
        a0, d0 = ref.expert(want_dist=True)
        a1, d1 = env.expert(want_dist=True, out=arena.carve((n,), torch.uint8, skew=1))
        assert _same(a0, a1) and _same(d0, d1)
        kind = torch.full((n,), 8, dtype=torch.uint8)
        g0, l0, s0 = ref.find_closest(kind, seq_cap=64)
        g1, l1, s1 = env.find_closest(kind, seq_cap=64)
        assert _same(l0, l1) and _same(s0[l0 >= 0], s1[l1 >= 0])
        for fused in (True, False):
            fo = arena.carve((n, nf), torch.float32)
            o = dict(expert=arena.carve((n,), torch.uint8), done=arena.carve((n,), torch.uint8),
                     success=arena.carve((n,), torch.uint8))
            b = env.tick(features_out=fo, fused=fused, out=o)
            a = ref.tick(fused=fused)
            for k in ("expert", "done", "success", "features"):
                assert _same(a[k], b[k]), (fused, k)
        o = dict(expert=arena.carve((3, n), torch.uint8), done=arena.carve((3, n), torch.uint8),
                 success=arena.carve((3, n), torch.uint8))
        ring = arena.carve((2, n, nf), torch.float32)
        b = env.rollout(3, features_out=ring, out=o)
        ring_ref = torch.empty((2, n, nf), dtype=torch.float32, device=ref.device)
        a = ref.rollout(3, features_out=ring_ref)
        assert _same(a["expert"], b["expert"]) and _same(ring, ring_ref)
        assert _same(env.grid, ref.grid) and _same(env.agent, ref.agent)
        assert arena.guards_intact(), "%dx%d" % (tables.W, tables.H)
        env.check_errors()


def test_batches_with_different_tables_on_two_streams(splits, medium_tables, large_tables, large_states):
    """Two batches with different tables interleaved on two CUDA streams: each keeps its own device
    copy of the tables (a single shared slot would be rewritten under a running kernel)."""
    from psketch_b200.vec import VecCraft
    n = 20000
    grids, env_idx, pos, task = _instances(splits, n, seed=5)
    S = large_states
    m = len(S["grid"])
    task_l = np.random.RandomState(3).choice([13, 14, 15, 19, 20, 21], size=m)

    def run(two_streams):
        a = VecCraft.from_instances(medium_tables, grids, env_idx, pos, task)
        b = VecCraft.from_states(large_tables, S["grid"], S["inv"], S["pos"], S["dir"], task=task_l)
        torch.cuda.synchronize()
        sa = torch.cuda.Stream() if two_streams else torch.cuda.current_stream()
        sb = torch.cuda.Stream() if two_streams else torch.cuda.current_stream()
        ea, eb = [], []
        for _ in range(30):
            with torch.cuda.stream(sa):
                ea.append(a.tick(want_features=False)["expert"].clone())
            with torch.cuda.stream(sb):
                eb.append(b.tick(want_features=False)["expert"].clone())
        torch.cuda.synchronize()
        a.check_errors()
        b.check_errors()
        return torch.stack(ea).cpu(), torch.stack(eb).cpu(), a.agent.cpu(), b.agent.cpu()

    one = run(False)
    two = run(True)
    for x, y in zip(one, two):
        assert torch.equal(x, y)
