"""world_size-2 gloo test of the sharding + statistics all-reduce (the only collective of the
path), with each rank's shard rolled out by the CPU oracle in place of the GPU kernels."""
import os
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, ticks, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from psketch_b200 import dist as pdist
    from psketch_b200.tables import CraftTables
    from oracle.craft_oracle import CraftOracle
    r, w, _ = pdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    sp = np.load(os.path.join(ROOT, "tests", "golden", "craft_medium_splits.npz"))
    lo, hi = pdist.shard_range(n_total, rank, world)
    idx = np.arange(lo, hi) % 2200
    o = CraftOracle(CraftTables())
    ienv = sp["dev_inst_env"][idx]
    _, stats, _, _ = o.rollout(ticks, 40, sp["dev_grids"][ienv.astype(np.int64)],
                               sp["dev_inst_pos"][idx].astype(np.int32),
                               sp["dev_inst_task"][idx].astype(np.int32))
    t = torch.from_numpy(stats.copy())
    pdist.allreduce_stats(t)
    slow = pdist.allreduce_max(1.0 + rank)
    if rank == 0:
        np.save(out_path, np.concatenate([t.numpy(), [slow, lo, hi]]))
    torch.distributed.destroy_process_group()


def test_shard_ranges_cover_everything():
    from psketch_b200.dist import shard_range
    for n, w in ((10, 3), (8388608, 8), (7, 8), (65536, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_stats_allreduce(tmp_path):
    n_total, ticks = 3001, 25
    out = str(tmp_path / "res.npy")
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, n_total, ticks, out), nprocs=2, join=True)
    got = np.load(out)
    # single-process answer over the whole batch
    sys.path.insert(0, ROOT)
    from psketch_b200.tables import CraftTables
    from oracle.craft_oracle import CraftOracle
    sp = np.load(os.path.join(ROOT, "tests", "golden", "craft_medium_splits.npz"))
    idx = np.arange(n_total) % 2200
    o = CraftOracle(CraftTables())
    ienv = sp["dev_inst_env"][idx]
    _, stats, _, _ = o.rollout(ticks, 40, sp["dev_grids"][ienv.astype(np.int64)],
                               sp["dev_inst_pos"][idx].astype(np.int32),
                               sp["dev_inst_task"][idx].astype(np.int32))
    assert got[:3].astype(np.int64).tolist() == stats[:3].tolist()
    assert got[2] == n_total * ticks
    assert got[4] == 2.0 and got[5] == 0 and got[6] == 1501
