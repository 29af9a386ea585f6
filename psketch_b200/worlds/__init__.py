"""World registry.  ``load(config)`` resolves ``config.world.name`` like the reference's factory
(worlds/__init__.py:5-11) and raises the same ``Exception("No such world: ...")`` for unknown names."""
from .craft import CraftScenario, CraftState, CraftWorld  # noqa: F401
from .light import LightScenario, LightState, LightWorld, VecLight  # noqa: F401

REGISTRY = {"CraftWorld": CraftWorld, "LightWorld": LightWorld}


def load(config):
    name = config.world.name
    if name not in REGISTRY:
        raise Exception("No such world: {}".format(name))
    return REGISTRY[name](config)
