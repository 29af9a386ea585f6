"""Teacher registry.  ``load(config)`` resolves ``config.teacher.name`` like the reference's
factory (teachers/__init__.py:6-12), same ``Exception("No such teacher: ...")`` otherwise."""
from .demonstration import BaseTeacher, DemonstrationTeacher  # noqa: F401
from .primitive_language import (InteractivePrimitiveLanguageTeacher,  # noqa: F401
                                 PrimitiveLanguageTeacher)

REGISTRY = {
    "DemonstrationTeacher": DemonstrationTeacher,
    "PrimitiveLanguageTeacher": PrimitiveLanguageTeacher,
    "InteractivePrimitiveLanguageTeacher": InteractivePrimitiveLanguageTeacher,
}


def load(config):
    name = config.teacher.name
    if name not in REGISTRY:
        raise Exception("No such teacher: {}".format(name))
    return REGISTRY[name](config)
