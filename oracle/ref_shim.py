"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference from /root/reference.

This module exists to (a) validate the CPU restatement in ``oracle/`` against the real
reference implementation and (b) generate the committed golden fixtures under
``tests/golden/`` (see ``oracle/gen_golden.py``).  It only works in the build container,
where ``/root/reference`` is mounted; nothing in ``tests -m gpu``, ``bench.py`` or
``__graft_entry__.smoke()`` may import it (``/root/reference`` does not exist on the GPU box).

Two shims are needed because two third-party packages are absent from this image
(SURVEY.md §8(c)):

* ``skimage.measure.block_reduce`` (used only at worlds/craft.py:308-310).  For the window
  sizes that occur (9 = 3*3, 25 = 5*5) the image divides evenly into blocks, so the
  published algorithm (pad to a multiple of the block with ``cval``, then reduce every block
  with ``func``) needs no padding and the shim is exact.
* ``jsonargparse`` (flags.py:1).  We skip ``flags.make_config`` and build the ``Struct``
  straight from the experiment YAML, which is what ``make_config`` does after parsing argv
  (flags.py:56-63).
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import yaml

REF_ROOT = os.environ.get("PSKETCH_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "worlds", "craft.py"))


def _block_reduce(image, block_size, func=np.sum, cval=0, func_kwargs=None):
    image = np.asarray(image)
    pads = []
    for dim, b in zip(image.shape, block_size):
        rem = (-dim) % b
        pads.append((0, rem))
    if any(p[1] for p in pads):
        image = np.pad(image, pads, mode="constant", constant_values=cval)
    shape = []
    for dim, b in zip(image.shape, block_size):
        shape += [dim // b, b]
    blocked = image.reshape(shape)
    axes = tuple(range(1, 2 * image.ndim, 2))
    return func(blocked, axis=axes)


def _install_shims():
    if "skimage.measure" not in sys.modules:
        try:
            import skimage.measure  # noqa: F401
        except Exception:
            skimage = types.ModuleType("skimage")
            measure = types.ModuleType("skimage.measure")
            measure.block_reduce = _block_reduce
            skimage.measure = measure
            sys.modules["skimage"] = skimage
            sys.modules["skimage.measure"] = measure
    if "curses" not in sys.modules:
        try:
            import curses  # noqa: F401
        except Exception:
            sys.modules["curses"] = types.ModuleType("curses")


@contextlib.contextmanager
def reference_cwd():
    """The reference opens configs/resources/data by relative path (worlds/craft.py:62-63)."""
    old = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        yield
    finally:
        os.chdir(old)


def make_config(experiment="imitation", seed=123, world_config=None, recipes=None, hints=None):
    from misc.util import Struct  # reference module
    with open(os.path.join(REF_ROOT, "configs", "experiments", experiment + ".yaml")) as f:
        raw = yaml.safe_load(f)
    if world_config is not None:
        raw["world"]["config"] = world_config
    if recipes is not None:                 # another cookbook (absolute path), worlds/craft.py:61
        raw["recipes"] = os.path.abspath(recipes)
    if hints is not None:                   # another hint file, data/task.py:36
        raw["trainer"]["hints"] = os.path.abspath(hints)
    config = Struct(**raw)
    config.random = np.random.RandomState(seed)
    return config


class Reference(object):
    """Handle on the live reference objects (world, teacher, task manager)."""

    def __init__(self, experiment="imitation", seed=123, world_config=None, recipes=None, hints=None):
        if not reference_available():
            raise RuntimeError("reference tree not found at %s" % REF_ROOT)
        _install_shims()
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        with reference_cwd():
            import worlds.craft as ref_craft
            import worlds.light as ref_light
            import teachers.demonstration as ref_demo
            import data.task as ref_task
            self.config = make_config(experiment, seed, world_config, recipes, hints)
            self.task_manager = ref_task.TaskManager(self.config)
            self.world = ref_craft.CraftWorld(self.config)
            self.teacher = ref_demo.DemonstrationTeacher(self.config)
            self.craft_module = ref_craft
            self.light_module = ref_light

    def light_world(self):
        from misc.util import Struct
        cfg = Struct(recipes=os.path.join(REF_ROOT, "resources/light/recipes.yaml"))
        return self.light_module.LightWorld(cfg)

    def load_split(self, split):
        import json
        path = os.path.join(REF_ROOT, "data", "%s_%s.json" % (self.config.world.config, split))
        with open(path) as f:
            return json.load(f)


def regenerate_dataset(out_dir):
    """Run the reference's make_data.py (seed 123, make_data.py:155) unmodified, with a fake
    ``flags`` module, writing the three split JSONs into ``out_dir``.  Recovers the train
    split that is missing from the reference checkout (.MISSING_LARGE_BLOBS:1)."""
    _install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    os.makedirs(out_dir, exist_ok=True)
    fake_flags = types.ModuleType("flags")

    def _make_config():
        cfg = make_config("imitation", 123)
        cfg.data_dir = out_dir
        return cfg

    fake_flags.make_config = _make_config
    saved = {k: sys.modules.get(k) for k in ("flags", "models")}
    sys.modules["flags"] = fake_flags
    sys.modules["models"] = types.ModuleType("models")  # make_data imports but never uses it
    try:
        with reference_cwd(), contextlib.redirect_stdout(io.StringIO()):
            src = open(os.path.join(REF_ROOT, "make_data.py")).read()
            exec(compile(src, os.path.join(REF_ROOT, "make_data.py"), "exec"),
                 {"__name__": "__make_data__"})
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return out_dir
