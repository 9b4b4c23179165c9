"""Minimal pure-Python stand-in for the `taichi` package, used ONLY by
tests/golden/make_golden.py to EXECUTE THE REFERENCE'S OWN @ti.kernel SOURCE (modules/*.py)
on tiny inputs and record golden vectors.  Taichi itself (taichi_nightly==1.7.0.post20230921,
requirements.txt:19) cannot be installed in this image.

Semantics implemented (Taichi defaults: default_fp = f32, default_ip = i32):
  * float values are numpy float32 (Python float literals are NEP-50 "weak" so they adopt
    f32; captured Python float constants are converted to f32 when a kernel is built);
  * integer loop indices / i32 values are Python ints, u32 arithmetic is numpy uint32 (wraps);
  * struct-for over an ndarray iterates its indices; ti.static / ti.ndrange / ti.grouped are
    plain Python iteration in index order (= a sequential run of the parallel loop, which is
    the canonical order fixed in DESIGN.md for the atomic counters);
  * exp is the correctly rounded f32 exp; vector min/max ignore NaN (LLVM minnum/maxnum);
  * out-of-bounds ndarray writes are dropped (Taichi release builds do not bounds-check; the
    reference writes T[s+1] one past the end in volume_train.py:47);
  * ti.atomic_add(x[i], v) is rewritten to a read-modify-write returning the old value;
  * kernel.grad (autodiff) is NOT available.
"""
import ast
import inspect
import itertools
import textwrap
import types as _pytypes

import numpy as np

np.seterr(over="ignore", divide="ignore", invalid="ignore")

f32 = np.float32
f16 = np.float16
f64 = np.float64
cuda, cpu, gpu = "cuda", "cpu", "cuda"


def init(*a, **k):
    return None


def sync():
    return None


def loop_config(**k):
    return None


def static(x):
    return x


class Vec(np.ndarray):
    """small fixed-size vector (ti.Vector / vec3 / uvec3)"""
    __array_priority__ = 100

    def __new__(cls, data, dtype=None):
        a = np.array(data, dtype=dtype)
        if dtype is None:
            if a.dtype.kind == "f":
                a = a.astype(np.float32)
        return a.view(cls)

    def max(self):
        return np.fmax.reduce(np.asarray(self))

    def min(self):
        return np.fmin.reduce(np.asarray(self))

    def any(self):
        return bool(np.any(np.asarray(self) != 0))

    def __getitem__(self, i):
        return np.asarray(self)[i]

    def __bool__(self):
        raise TypeError("vector truth value")


def Vector(vals, dt=None):
    return Vec(list(vals), dt)


def _is_vec(x):
    return isinstance(x, np.ndarray) and x.ndim >= 1


def _scalar_cast(v, dt):
    if _is_vec(v):
        if np.dtype(dt).kind in "ui":
            return Vec(np.trunc(np.asarray(v, dtype=np.float64)).astype(np.int64).astype(dt) if np.asarray(v).dtype.kind == "f"
                       else np.asarray(v).astype(dt), dt)
        return Vec(np.asarray(v).astype(dt), dt)
    if np.dtype(dt).kind == "u":
        if isinstance(v, (float, np.floating)):
            v = int(v)          # truncation toward zero
        return dt(int(v) & ((1 << (8 * np.dtype(dt).itemsize)) - 1))
    if np.dtype(dt).kind == "i":
        return int(v)           # i32 values are Python ints (weak)
    return dt(v)


class _DType:
    def __init__(self, np_dt, name):
        self.np_dt, self.name = np_dt, name

    def __call__(self, v):
        return _scalar_cast(v, self.np_dt)


u32 = uint32 = _DType(np.uint32, "u32")
u8 = uint8 = _DType(np.uint8, "u8")
i32 = int32 = _DType(np.int32, "i32")
i64 = int64 = _DType(np.int64, "i64")


def _np_dt(dt):
    if isinstance(dt, _DType):
        return dt.np_dt
    return dt


def cast(v, dt):
    return _scalar_cast(v, _np_dt(dt))


def bit_cast(v, dt):
    dt = _np_dt(dt)
    if dt == np.uint32:
        return np.float32(v).view(np.uint32)
    if dt == np.float32:
        return np.uint32(v).view(np.float32)
    raise NotImplementedError(dt)


def _f(x):
    if _is_vec(x):
        return Vec(np.asarray(x, dtype=np.float32), np.float32)
    return np.float32(x)


def _wrap(r, x):
    return Vec(r, np.float32) if _is_vec(x) else np.float32(r)


def exp(x):
    x = _f(x)
    return _wrap(np.exp(np.asarray(x, dtype=np.float64)).astype(np.float32), x)   # correctly rounded f32 exp


def floor(x):
    x = _f(x)
    return _wrap(np.floor(x), x)


def ceil(x):
    x = _f(x)
    return _wrap(np.ceil(x), x)


def sqrt(x):
    x = _f(x)
    return _wrap(np.sqrt(x), x)


def abs(x):  # noqa: A001
    if _is_vec(x):
        return Vec(np.abs(np.asarray(x)), np.asarray(x).dtype)
    return np.abs(x)


def pow(a, b):  # noqa: A001
    return np.float32(np.float64(a) ** np.float64(b))


def _minmax(a, b, fn, pyfn):
    if _is_vec(a) or _is_vec(b):
        r = fn(np.asarray(a), np.asarray(b))
        return Vec(r, r.dtype)
    if isinstance(a, (float, np.floating)) or isinstance(b, (float, np.floating)):
        return np.float32(fn(np.float32(a), np.float32(b)))
    return pyfn(a, b)


_builtin_min, _builtin_max = min, max


def min(a, b):  # noqa: A001
    return _minmax(a, b, np.fmin, _builtin_min)


def max(a, b):  # noqa: A001
    return _minmax(a, b, np.fmax, _builtin_max)


def random(dt=None):
    raise NotImplementedError("ti.random is not used by the hot-path kernels")


def ndrange(*ns):
    if len(ns) == 1:
        return range(int(ns[0]))
    return itertools.product(*[range(int(n)) for n in ns])


def grouped(arr):
    return itertools.product(*[range(int(n)) for n in arr.shape])


# ---- ndarray proxies ------------------------------------------------------------------
class _NdAnn:
    def __init__(self, vec_n=0, vec_dt=None):
        self.vec_n, self.vec_dt = vec_n, vec_dt


class NdArr:
    """torch tensor (CPU) seen from inside a kernel"""

    def __init__(self, t, vec_n=0):
        self.a = t.numpy() if hasattr(t, "numpy") else np.asarray(t)
        self.vec_n = vec_n
        self.shape = self.a.shape[:-1] if vec_n else self.a.shape

    def __iter__(self):
        if len(self.shape) == 1:
            return iter(range(self.shape[0]))
        return itertools.product(*[range(n) for n in self.shape])

    def _idx(self, i):
        if not isinstance(i, tuple):
            i = (i,)
        return tuple(int(k) for k in i)

    def _inb(self, i):
        return all(0 <= k < n for k, n in zip(i, self.a.shape))

    def __getitem__(self, i):
        i = self._idx(i)
        v = self.a[i]
        if self.vec_n:
            return Vec(v.copy(), v.dtype)
        if self.a.dtype.kind in "iub":
            return int(v)
        return v.dtype.type(v)

    def __setitem__(self, i, v):
        i = self._idx(i)
        if not self._inb(i):
            return      # unchecked OOB write in the reference (volume_train.py:47)
        self.a[i] = np.asarray(v).astype(self.a.dtype) if _is_vec(v) else v


def _atomic_add(arr, idx, v):
    old = arr[idx]
    arr[idx] = old + v
    return old


def atomic_add(*a):
    raise RuntimeError("ti.atomic_add must be rewritten by the kernel decorator")


class _Types:
    @staticmethod
    def ndarray(dtype=None, ndim=None, **k):
        if isinstance(dtype, _VecType):
            return _NdAnn(dtype.n, dtype.dt)
        return _NdAnn()

    @staticmethod
    def vector(n, dtype):
        return _VecType(n, _np_dt(dtype))


class _VecType:
    def __init__(self, n, dt):
        self.n, self.dt = n, dt

    def __call__(self, *vals):
        if len(vals) == 1 and not _is_vec(vals[0]) and not isinstance(vals[0], (list, tuple)):
            return Vec([vals[0]] * self.n, self.dt)
        if len(vals) == 1:
            return Vec(list(vals[0]), self.dt)
        return Vec(list(vals), self.dt)


types = _Types()


def template():
    return None


def func(fn):
    return fn


class _AtomicRewrite(ast.NodeTransformer):
    def visit_Call(self, node):
        self.generic_visit(node)
        f = node.func
        if isinstance(f, ast.Attribute) and f.attr == "atomic_add" and isinstance(node.args[0], ast.Subscript):
            sub = node.args[0]
            return ast.Call(func=ast.Name(id="__ti_atomic_add", ctx=ast.Load()),
                            args=[sub.value, sub.slice, node.args[1]], keywords=[])
        return node


def _convert_const(v):
    if isinstance(v, float):
        return np.float32(v)
    return v


class Kernel:
    def __init__(self, fn):
        self.fn = fn
        src = textwrap.dedent(inspect.getsource(fn))
        tree = ast.parse(src)
        fdef = tree.body[0]
        fdef.decorator_list = []
        for a in fdef.args.args:
            a.annotation = None
        tree = ast.fix_missing_locations(_AtomicRewrite().visit(tree))
        ns = dict(fn.__globals__)
        if fn.__closure__:
            for name, cell in zip(fn.__code__.co_freevars, fn.__closure__):
                ns[name] = _convert_const(cell.cell_contents)
        ns["__ti_atomic_add"] = _atomic_add
        exec(compile(tree, inspect.getsourcefile(fn) or "<kernel>", "exec"), ns)
        self.body = ns[fn.__name__]
        self.sig = inspect.signature(fn)

    def __call__(self, *args, **kwargs):
        bound = self.sig.bind(*args, **kwargs)
        conv = []
        for name, val in bound.arguments.items():
            ann = self.sig.parameters[name].annotation
            if isinstance(ann, _NdAnn):
                conv.append(NdArr(val, ann.vec_n))
            elif ann is float or ann is f32:
                conv.append(np.float32(val))
            elif ann is int or isinstance(ann, _DType):
                conv.append(int(val))
            else:
                conv.append(val)
        return self.body(*conv)

    def grad(self, *a, **k):
        raise NotImplementedError("Taichi autodiff is not emulated by the shim")


def kernel(fn):
    return Kernel(fn)
