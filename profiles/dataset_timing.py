"""Wall-clock of psketch_b200.data.generate_dataset (make_data.py:164-238 on the GPU: distinct
scenarios, 20 start cells for each of the 11 get/make tasks, teacher rollouts to STOP) for the
reference's size (100 worlds -> 22,000 instances; the reference needs ~27 s on one core) and
larger ones.  One JSON line per size."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from psketch_b200.data import generate_dataset  # noqa: E402
from psketch_b200.tables import CraftTables  # noqa: E402

tables = CraftTables()
generate_dataset(tables, n_worlds=10, seed=1)          # warm-up (library load, allocator)
for n_worlds in (100, 2000, 20000):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d = generate_dataset(tables, n_worlds=n_worlds, seed=123)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"n_worlds": n_worlds, "instances": int(len(d["inst_env"])),
                      "teacher_steps": int(d["ref_len"].sum()), "seconds": dt,
                      "instances_per_s": len(d["inst_env"]) / dt}))
