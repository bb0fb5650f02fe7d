"""Import alias: the package directory is ``proto-clip_b200/`` (not a valid Python identifier), so this
shim makes it importable as ``proto_clip_b200`` by pointing ``__path__`` at it and running its __init__."""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_real = _os.path.join(_os.path.dirname(_here), "proto-clip_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
