"""Dataset wire format (data/dataset.py:39-67) <-> packed arrays; CPU only."""
import numpy as np


def test_wire_format_reads_reference_goldens(splits, medium_tables):
    """The shipped dev split, re-serialised to the reference's JSON layout and read back."""
    from psketch_b200 import data
    packed = {k[4:]: splits[k] for k in splits.files if k.startswith("dev_")}
    packed["inst_env"] = packed["inst_env"].astype(np.int32)
    packed["ref_len"] = packed["ref_len"].astype(np.int32)
    wire = data.to_wire(packed, medium_tables)
    assert len(wire) == 10 and len(wire[0]["task_instances"]) == 11
    assert np.asarray(wire[0]["grid"]).shape == (8, 8, 21)
    assert wire[0]["task_instances"][0]["task"] == "get[wood]"
    back = data.from_wire(wire, medium_tables)
    assert np.array_equal(back["grids"], packed["grids"])
    assert np.array_equal(back["ref_actions"], packed["ref_actions"])


def test_wire_format_accepts_both_task_spellings(medium_tables):
    """Shipped files say "get[wood]"; a current make_data.py writes "get wood" (data/task.py:28-29)."""
    from psketch_b200 import data
    grid = np.zeros((8, 8, 21))
    grid[0, :, 1] = grid[7, :, 1] = grid[:, 0, 1] = grid[:, 7, 1] = 1
    grid[3, 3, 9] = 1
    env = {"grid": grid.tolist(), "task_instances": [
        {"task": "get[wood]", "init_pos": [[1, 1]], "ids": ["instance_1"], "ref_actions": [[1, 5]]},
        {"task": "make plank", "init_pos": [[2, 2]], "ids": ["instance_2"], "ref_actions": [[3, 4, 5]]}]}
    p = data.from_wire([env], medium_tables)
    assert p["inst_task"].tolist() == [13, 19] and p["ref_len"].tolist() == [2, 3]
    assert p["grids"][0].reshape(8, 8)[3, 3] == 9 and p["ref_actions"][0].tolist() == [1, 5, 255]
