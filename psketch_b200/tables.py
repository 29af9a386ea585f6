"""Host-side compilation of the Craft domain tables into the flat arrays the kernels read.

Reference semantics restated here (nothing is imported from the reference):

* ``Index`` — name <-> id map whose ids start at 1; id 0 is "invalid", unknown names map to
  ``None`` (misc/util.py:46-76).
* ``Cookbook`` — ``environment`` / ``primitives`` / ``recipes`` parsed from a recipes YAML in
  the reference's format; ids are handed out in file order: environment, primitives, then
  for every recipe its ingredients followed by its output (worlds/cookbook.py:8-26).
* kind classes — every id that is not an environment id is grabbable
  (worlds/craft.py:101-107); workshops craft, water needs a bridge, stone needs an axe
  (worlds/craft.py:383-410).
* hint tree — ``goal[arg] -> [sub-goal, ...]`` (data/task.py:34-59,
  resources/craft/hints.hierarchy.yaml); task ids and the vocab are handed out in file order.

The default Craft domain (the one ``resources/craft/recipes.yaml`` and
``resources/craft/hints.hierarchy.yaml`` describe) is embedded below as plain Python data so
that the package works on a box where the reference checkout does not exist; a YAML file in
the reference's format can be passed instead.
"""
import re

import numpy as np

# ----------------------------------------------------------------------------------------
# kind classes (device encoding, see csrc/psk_common.cuh)
KC_FREE = 0       # id 0 / unknown: nothing there
KC_INERT = 1      # boundary: blocks, USE does nothing
KC_WORKSHOP = 2
KC_WATER = 3
KC_STONE = 4
KC_GRAB = 5

# satisfies() classes (worlds/craft.py:285-294)
SAT_NEVER = 0     # goal_name not in {get, make, go}  -> None
SAT_INV = 1       # get / make : inventory[arg] > 0
SAT_FACING = 2    # go         : cell in front holds kind arg

# leaf kinds for the expert's hint-tree walk (teachers/demonstration.py:9-30)
LEAF_NONE = 0     # inner node
LEAF_USE = 1      # 'use'  -> USE
LEAF_GO = 2       # 'go'   -> BFS to nearest arg
LEAF_BAD = 3      # any other leaf: the reference asserts (demonstration.py:18)

MAX_RECIPES = 16
MAX_RECIPE_INPUTS = 2
MAX_TASKS = 32
MAX_TASK_NODES = 16
MAX_KINDS = 32

DOWN, UP, LEFT, RIGHT, USE, STOP = 0, 1, 2, 3, 4, 5
N_ACTIONS = 6
ACTION_NAMES = ("DOWN", "UP", "LEFT", "RIGHT", "USE", "STOP")
COORD_CHANGE = ((0, -1), (0, 1), (-1, 0), (1, 0), (0, 0), (0, 0))   # worlds/craft.py:77-91

_FEXP = re.compile(r"(.*)\[(.*)\]")


def parse_fexp(text):
    """'make[plank]' -> ('make', 'plank')   (misc/util.py:142-145)."""
    m = _FEXP.match(text)
    return m.group(1), m.group(2)


class Index(object):
    """Ids start at 1; 0 is reserved; unknown names read as None (misc/util.py:46-76)."""

    def __init__(self):
        self.contents = {}
        self.ordered_contents = []
        self.reverse_contents = {}

    def __getitem__(self, item):
        return self.contents.get(item)

    def index(self, item):
        if item not in self.contents:
            idx = len(self.contents) + 1
            self.ordered_contents.append(item)
            self.contents[item] = idx
            self.reverse_contents[idx] = item
        return self.contents[item]

    def get(self, idx):
        if idx == 0:
            return "*invalid*"
        return self.reverse_contents[idx]

    def __len__(self):
        return len(self.contents) + 1

    def __iter__(self):
        return iter(self.ordered_contents)


# The default Craft domain as data: (output, {ingredient: count}, workshop), in firing order.
DEFAULT_ENVIRONMENT = ("boundary", "workshop0", "workshop1", "workshop2", "water", "stone")
DEFAULT_PRIMITIVES = ("iron", "grass", "wood", "gold", "gem")
DEFAULT_RECIPES = (
    ("plank", (("wood", 1),), "workshop0"),
    ("axe", (("stick", 1), ("iron", 1)), "workshop0"),
    ("rope", (("grass", 1),), "workshop0"),
    ("stick", (("wood", 1),), "workshop1"),
    ("bed", (("plank", 1), ("grass", 1)), "workshop1"),
    ("shears", (("stick", 1), ("iron", 1)), "workshop1"),
    ("cloth", (("grass", 1),), "workshop2"),
    ("bridge", (("wood", 1), ("iron", 1)), "workshop2"),
    ("ladder", (("plank", 1), ("stick", 1)), "workshop2"),
)
_PRIMS = ("wood", "iron", "grass")
_SHOPS = ("workshop0", "workshop1", "workshop2")
DEFAULT_HINTS = tuple(
    [("%s[none]" % a, ()) for a in ("left", "right", "up", "down", "use", "stop")]
    + [("go[%s]" % p, ()) for p in _PRIMS]
    + [("go[%s]" % s, ()) for s in _SHOPS]
    + [("get[%s]" % p, ("go[%s]" % p, "use[none]")) for p in ("wood", "grass", "iron")]
    + [("makeat[%s]" % s, ("go[%s]" % s, "use[none]")) for s in _SHOPS]
    + [
        ("make[plank]", ("get[wood]", "makeat[workshop0]")),
        ("make[stick]", ("get[wood]", "makeat[workshop1]")),
        ("make[cloth]", ("get[grass]", "makeat[workshop2]")),
        ("make[rope]", ("get[grass]", "makeat[workshop0]")),
        ("make[bridge]", ("get[iron]", "get[wood]", "makeat[workshop2]")),
        ("make[bed]", ("make[plank]", "get[grass]", "makeat[workshop1]")),
        ("make[axe]", ("make[stick]", "get[iron]", "makeat[workshop0]")),
        ("make[shears]", ("make[stick]", "get[iron]", "makeat[workshop1]")),
    ]
)

WORLD_CONFIGS = {
    # configs/worlds/craft_medium.yaml, configs/worlds/craft_large.yaml
    "craft_medium": dict(WIDTH=8, HEIGHT=8, WINDOW_WIDTH=3, WINDOW_HEIGHT=3,
                         N_WORKSHOPS=3, N_PRIMITIVES=2, N_WORLDS=100),
    "craft_large": dict(WIDTH=10, HEIGHT=10, WINDOW_WIDTH=5, WINDOW_HEIGHT=5,
                        N_WORKSHOPS=3, N_PRIMITIVES=4, N_WORLDS=100),
}


class Cookbook(object):
    """Same attributes as the reference's Cookbook (worlds/cookbook.py:7-26):
    ``index``, ``environment``, ``primitives``, ``recipes`` ({out: {in: n, '_at': ws}}),
    ``n_kinds``.  ``recipes_path=None`` selects the embedded default domain."""

    def __init__(self, recipes_path=None):
        if recipes_path is None:
            raw = {
                "environment": list(DEFAULT_ENVIRONMENT),
                "primitives": list(DEFAULT_PRIMITIVES),
                "recipes": {out: dict(list(ins) + [("_at", ws)])
                            for out, ins, ws in DEFAULT_RECIPES},
            }
        else:
            import yaml
            with open(recipes_path) as f:
                raw = yaml.safe_load(f)
        self.index = Index()
        self.environment = set(self.index.index(e) for e in (raw.get("environment") or []))
        self.primitives = set(self.index.index(p) for p in (raw.get("primitives") or []))
        self.recipes = {}
        for output, inputs in (raw.get("recipes") or {}).items():
            entry = {}
            for name, count in inputs.items():
                if "_" in name:
                    entry[name] = count
                else:
                    entry[self.index.index(name)] = count
            self.recipes[self.index.index(output)] = entry
        self.n_kinds = len(self.index)

    def primitives_for(self, goal):
        """Total primitive cost of ``goal`` (worlds/cookbook.py:28-52)."""
        total = {}
        for ingredient, count in self.recipes[goal].items():
            if not isinstance(ingredient, int):
                continue
            if ingredient in self.primitives:
                total[ingredient] = total.get(ingredient, 0) + count
            else:
                sub = self.recipes[ingredient]
                n_made = sub.get("_yield", 1)
                n_needed = -(-count // n_made)
                for k, v in self.primitives_for(ingredient).items():
                    total[k] = total.get(k, 0) + v * n_needed
        return total


class Task(object):
    """Node of the hint tree (data/task.py:9-29)."""

    def __init__(self, goal, subtasks=None):
        self.goal_name, self.goal_arg = parse_fexp(goal)
        self.subtasks = subtasks if subtasks else None
        self.encoding = None
        self.task_id = 0

    def __repr__(self):
        return "Task(%s[%s])" % (self.goal_name, self.goal_arg)

    def __hash__(self):
        return hash(repr(self))

    def __eq__(self, other):
        return self.goal_name == other.goal_name and self.goal_arg == other.goal_arg

    def __str__(self):
        return self.goal_name + " " + self.goal_arg

    @property
    def goal(self):
        return "%s[%s]" % (self.goal_name, self.goal_arg)


class TaskManager(object):
    """``tasks_by_goal``, ``tasks`` (Index of Task, ids from 1), ``vocab`` and 2-token
    encodings exactly as data/task.py:32-75.  ``hints`` may be a path to a YAML file in the
    reference's format, an ordered mapping, or None for the embedded default."""

    def __init__(self, hints=None):
        if hints is None:
            items = [(g, list(s)) for g, s in DEFAULT_HINTS]
        elif isinstance(hints, str):
            import yaml
            with open(hints) as f:
                items = list(yaml.safe_load(f).items())
        else:
            items = list(hints.items())
        self.tasks_by_goal = {}
        self.tasks = Index()
        for goal, subgoals in items:
            subtasks = [self.tasks_by_goal[s] for s in (subgoals or [])]
            task = Task(goal, subtasks)
            self.tasks_by_goal[goal] = task
            task.task_id = self.tasks.index(task)
        self.vocab = Index()
        self.vocab.index("<EOS>")
        self.vocab.index("<PAD>")
        for task in self.tasks:
            self.vocab.index(task.goal_name)
            if task.goal_arg:
                self.vocab.index(task.goal_arg)
        for task in self.tasks:
            task.encoding = [self.vocab[task.goal_name], self.vocab[task.goal_arg]]

    def __getitem__(self, goal):
        return self.tasks_by_goal[goal]

    def by_id(self, task_id):
        return self.tasks.get(task_id)

    def __len__(self):
        return len(self.tasks) - 1


LAST_CHILD = 0x40      # task_nodes[..][2] bit: a satisfied node here = the reference's AssertionError


class CraftTables(object):
    """Everything the kernels need, as small numpy arrays (uploaded once per device by
    ``psk_craft_tables_upload``; layout documented in include/psk_craft.h):

    * ``kind_class u8[MAX_KINDS]``
    * ``recipes u8[MAX_RECIPES, 8]`` rows ``(out, workshop_kind, n_in, in0, cnt0, in1, cnt1,
      yield)`` in firing order (YAML insertion order, worlds/craft.py:391)
    * ``task_nodes u8[MAX_TASKS, MAX_TASK_NODES, 4]`` — per root task the pre-order
      flattening of its hint tree, rows ``(sat_class, arg_kind, leaf_kind, skip_to)``.  The
      recursive walk of teachers/base.py:10-25 becomes: ``i = 0; while i < n: if sat(i):
      i = skip_to[i] elif leaf(i): return i else: i += 1``.
    * ``task_len u8[MAX_TASKS]``
    """

    def __init__(self, cookbook=None, task_manager=None, world_config="craft_medium"):
        self.cookbook = cookbook if cookbook is not None else Cookbook()
        self.task_manager = task_manager if task_manager is not None else TaskManager()
        cfg = WORLD_CONFIGS[world_config] if isinstance(world_config, str) else dict(world_config)
        self.world_config = cfg
        self.W = int(cfg["WIDTH"])
        self.H = int(cfg["HEIGHT"])
        self.win_w = int(cfg["WINDOW_WIDTH"])
        self.win_h = int(cfg["WINDOW_HEIGHT"])
        cb = self.cookbook
        self.K = cb.n_kinds
        if self.K > MAX_KINDS:
            raise ValueError("too many kinds: %d > %d" % (self.K, MAX_KINDS))
        self.n_features = 2 * self.win_w * self.win_h * self.K + self.K + 4 + 1
        n_ws = int(cfg.get("N_WORKSHOPS", 3))
        self.workshop_kinds = [cb.index["workshop%d" % i] for i in range(n_ws)]
        self.water_kind = cb.index["water"] or 0
        self.stone_kind = cb.index["stone"] or 0
        self.bridge_kind = cb.index["bridge"] or 0
        self.axe_kind = cb.index["axe"] or 0

        kc = np.zeros(MAX_KINDS, np.uint8)
        for k in range(1, self.K):
            if k in cb.environment:
                kc[k] = KC_INERT
            else:
                kc[k] = KC_GRAB
        # order of the reference's elif chain: grabbable, workshop, water, stone (craft.py:383-410)
        for k in self.workshop_kinds:
            if k is not None and kc[k] != KC_GRAB:
                kc[k] = KC_WORKSHOP
        if self.water_kind and kc[self.water_kind] == KC_INERT:
            kc[self.water_kind] = KC_WATER
        if self.stone_kind and kc[self.stone_kind] == KC_INERT:
            kc[self.stone_kind] = KC_STONE
        self.kind_class = kc

        rec = np.zeros((MAX_RECIPES, 8), np.uint8)
        n_rec = 0
        for out, entry in cb.recipes.items():
            ws = cb.index[entry["_at"]]
            ins = [(k, v) for k, v in entry.items() if isinstance(k, int)]
            if len(ins) > MAX_RECIPE_INPUTS:
                raise ValueError("recipe with more than %d inputs" % MAX_RECIPE_INPUTS)
            if n_rec >= MAX_RECIPES:
                raise ValueError("too many recipes")
            row = [out, ws or 0, len(ins), 0, 0, 0, 0, entry.get("_yield", 1)]
            for j, (k, v) in enumerate(ins):
                row[3 + 2 * j] = k
                row[4 + 2 * j] = v
            rec[n_rec] = row
            n_rec += 1
        self.recipes = rec
        self.n_recipes = n_rec

        tm = self.task_manager
        n_tasks = len(tm.tasks)            # ids 1..n_tasks-1
        if n_tasks > MAX_TASKS:
            raise ValueError("too many tasks")
        self.n_tasks = n_tasks
        nodes = np.zeros((MAX_TASKS, MAX_TASK_NODES, 4), np.uint8)
        lens = np.zeros(MAX_TASKS, np.uint8)
        self.may_assert = []            # (task, last subtask) pairs where the reference can assert
        for task in tm.tasks:
            flat = []
            self._flatten(task, flat)
            if len(flat) > MAX_TASK_NODES:
                raise ValueError("hint tree of %r too deep (%d nodes)" % (task, len(flat)))
            nodes[task.task_id, :len(flat)] = flat
            lens[task.task_id] = len(flat)
        self.task_nodes = nodes
        self.task_len = lens
        self.task_is_get = np.zeros(MAX_TASKS, np.uint8)
        for task in tm.tasks:
            self.task_is_get[task.task_id] = task.goal_name == "get"

    def sat_class(self, task):
        if task.goal_name in ("make", "get"):
            return SAT_INV
        if task.goal_name == "go":
            return SAT_FACING
        return SAT_NEVER

    def _flatten(self, task, flat):
        here = len(flat)
        arg = self.cookbook.index[task.goal_arg] or 0
        sat = self.sat_class(task)
        if sat != SAT_NEVER and arg == 0:
            # reference: inventory[None] / grid[..., None] — numpy newaxis, result array > 0
            # is ambiguous/garbage; we treat "unknown kind" as never satisfied.
            sat = SAT_NEVER
        if task.subtasks is None:
            if task.goal_name == "use":
                leaf = LEAF_USE
            elif task.goal_name == "go":
                leaf = LEAF_GO
            else:
                leaf = LEAF_BAD
        else:
            leaf = LEAF_NONE
        flat.append([sat, arg, leaf, 0])
        if task.subtasks is not None:
            for sub in task.subtasks:
                child = len(flat)
                self._flatten(sub, flat)
            # The reference asserts that the LAST subtask of an unsatisfied task is itself incomplete
            # (teachers/base.py:23-24).  A child is only ever visited below an unsatisfied parent, so
            # the walk raises exactly when it finds this node satisfied: mark it.  (Never the case
            # with the stock hint file: its last subtasks are use[...] / makeat[...], which are never
            # "satisfied".)
            if flat[child][0] != SAT_NEVER:
                flat[child][2] |= LAST_CHILD
                self.may_assert.append((repr(task), repr(sub)))
        flat[here][3] = len(flat)

    # --------------------------------------------------------------------------------
    def action_table(self):
        return COORD_CHANGE
