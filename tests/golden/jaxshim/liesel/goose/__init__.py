"""Import-time placeholder for liesel.goose (type names only)."""
from . import types  # noqa: F401


def __getattr__(name):  # kernels, GibbsTransition, ...: inert classes, never instantiated here
    return type(name, (), {"__init__": lambda self, *a, **k: None})
