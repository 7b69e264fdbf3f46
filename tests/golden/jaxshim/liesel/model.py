"""Import-time placeholder for liesel.model: base classes the reference derives from.
Nothing of the graph machinery is implemented; only static methods of the derived classes
are called by the fixture generator."""


class _Node:
    def __init__(self, *a, **k):
        pass

    def update(self):
        return self


class Var(_Node):
    pass


class Value(_Node):
    pass


class Calc(_Node):
    pass


class Dist(_Node):
    pass


class Model(_Node):
    pass


def __getattr__(name):  # any other name: an inert class
    return type(name, (_Node,), {})
