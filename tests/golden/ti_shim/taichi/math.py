"""taichi.math subset used by the reference's kernels (see taichi/__init__.py)."""
import numpy as np

from . import Vec, _is_vec, _f, _wrap, pow  # noqa: F401


def _vecn(n, dt):
    def make(*vals):
        if len(vals) == 1 and (_is_vec(vals[0]) or isinstance(vals[0], (list, tuple))):
            vals = list(vals[0])
        elif len(vals) == 1:
            vals = [vals[0]] * n
        return Vec(list(vals), dt)
    make.n, make.dt = n, dt
    return make


class _VecAnn:
    """vec3 / vec2 / uvec3 are both constructors and ndarray element types"""

    def __init__(self, n, dt):
        self.n, self.dt = n, dt
        self._make = _vecn(n, dt)

    def __call__(self, *vals):
        return self._make(*vals)


from . import _VecType  # noqa: E402


class _VT(_VecType):
    def __call__(self, *vals):
        if len(vals) == 1 and (_is_vec(vals[0]) or isinstance(vals[0], (list, tuple))):
            vals = list(vals[0])
        elif len(vals) == 1:
            vals = [vals[0]] * self.n
        return Vec(list(vals), self.dt)


vec2 = _VT(2, np.float32)
vec3 = _VT(3, np.float32)
uvec3 = _VT(3, np.uint32)
ivec3 = _VT(3, np.int32)       # imported (not executed) by modules/triplane.py
ivec2 = _VT(2, np.int32)
uvec2 = _VT(2, np.uint32)
vec4 = _VT(4, np.float32)


def clamp(x, xmin, xmax):
    """taichi.math.clamp = min(xmax, max(xmin, x)) in f32"""
    x = _f(x)
    lo, hi = _f(xmin), _f(xmax)
    return _wrap(np.fmin(hi, np.fmax(lo, x)), x)


def sign(x):
    x = _f(x)
    return _wrap(np.sign(x), x)
