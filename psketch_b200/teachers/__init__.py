"""Teacher factory with the reference's lookup-by-class-name contract (teachers/__init__.py:6-12)."""
from .demonstration import BaseTeacher, DemonstrationTeacher  # noqa: F401
from .primitive_language import (InteractivePrimitiveLanguageTeacher,  # noqa: F401
                                 PrimitiveLanguageTeacher)


def load(config):
    cls_name = config.teacher.name
    try:
        cls = globals()[cls_name]
    except KeyError:
        raise Exception("No such teacher: {}".format(cls_name))
    return cls(config)
