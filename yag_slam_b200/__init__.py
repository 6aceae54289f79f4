"""Importable alias for the `yag-slam_b200/` package directory (a hyphen is not a
valid Python identifier, so `import yag_slam_b200` is routed to that directory)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "yag-slam_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
