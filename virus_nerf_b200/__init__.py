"""Importable alias of the package directory ``virus-nerf_b200/`` (a hyphen is not a valid
Python identifier).  ``import virus_nerf_b200`` executes ``virus-nerf_b200/__init__.py`` with
this module's ``__path__`` pointing at the real directory, so sub-modules resolve there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "virus-nerf_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
