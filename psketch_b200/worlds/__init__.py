"""World factory with the reference's lookup-by-class-name contract (worlds/__init__.py:5-11)."""
from .craft import CraftWorld, CraftScenario, CraftState  # noqa: F401
from .light import LightWorld, LightScenario, LightState, VecLight  # noqa: F401


def load(config):
    cls_name = config.world.name
    try:
        cls = globals()[cls_name]
    except KeyError:
        raise Exception("No such world: {}".format(cls_name))
    return cls(config)
