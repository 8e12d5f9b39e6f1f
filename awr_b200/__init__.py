"""Importable alias of the package directory `awr-adaptive-weighting-regression_b200/`
(a hyphenated directory cannot be imported by name)."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "awr-adaptive-weighting-regression_b200")
__path__ = [_impl]
with open(_os.path.join(_impl, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, "__init__.py"), "exec"))
