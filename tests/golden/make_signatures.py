"""Generates tests/golden/signatures.json -- the call signatures (inspect.signature) of every public function, class
and method of the reference modules on the hot path, taken from THE REFERENCE'S OWN CODE in /root/reference (build
container only; Taichi runs under tests/golden/ti_shim).  tests/test_abi_cpu.py compares the host-side mirrors in
virus_nerf_b200 against this fixture (same names, same parameter names in the same order, same defaults where they are
plain literals).

Run:  python tests/golden/make_signatures.py     (needs /root/reference)
"""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import setup_reference_imports  # noqa: E402

# reference module -> mirror module of the product package
MODULES = {
    "modules.hash_encoder": "modules.hash_encoder",
    "modules.hash_encoder_half": "modules.hash_encoder_half",
    "modules.intersection": "modules.intersection",
    "modules.ray_march": "modules.ray_march",
    "modules.volume_train": "modules.volume_train",
    "modules.volume_render_test": "modules.volume_render_test",
    "modules.spherical_harmonics": "modules.spherical_harmonics",
    "modules.rendering": "modules.rendering",
    "modules.networks": "modules.networks",
    "modules.utils": "modules.utils",
    "modules.grid": "modules.grid",
    "modules.occupancy_grid": "modules.occupancy_grid",
    "modules.ngp_grid": "modules.ngp_grid",
    "datasets.dataset_base": "datasets.dataset_base",
    "training.sampler": "training.sampler",
}


def _default(p):
    d = p.default
    if d is inspect.Parameter.empty:
        return "<required>"
    if d is None or isinstance(d, (bool, int, float, str)):
        return repr(d)
    return "<object>"


def describe(fn):
    """[(name, kind, default), ...] of a callable; None when it has no Python signature"""
    try:
        sig = inspect.signature(fn)
    except (TypeError, ValueError):
        return None
    return [[p.name, p.kind.name, _default(p)] for p in sig.parameters.values()]


def unwrap(obj):
    """the Python function behind a @ti.kernel / @ti.func / torch custom-op wrapper"""
    for attr in ("__wrapped__", "_primal", "func", "fn"):
        inner = getattr(obj, attr, None)
        if inspect.isfunction(inner):
            return inner
    return obj


def collect(mod):
    out = {}
    for name, obj in vars(mod).items():
        if name.startswith("__"):
            continue
        inner = unwrap(obj)       # a @ti.kernel is a shim Kernel object around the reference's function
        if getattr(inner, "__module__", None) != mod.__name__ and getattr(obj, "__module__", None) != mod.__name__:
            continue
        if inspect.isclass(obj):
            methods = {}
            for mname, m in vars(obj).items():
                if mname.startswith("__") and mname != "__init__":
                    continue
                f = m.__func__ if isinstance(m, (staticmethod, classmethod)) else m
                f = unwrap(f)
                if inspect.isfunction(f):
                    methods[mname] = describe(f)
            out[name] = {"kind": "class", "methods": methods}
        elif callable(obj):
            f = unwrap(obj)
            d = describe(f)
            if d is not None:
                out[name] = {"kind": "function", "params": d}
    return out


def main():
    setup_reference_imports()
    import importlib
    sigs = {}
    for ref_name, mirror in MODULES.items():
        mod = importlib.import_module(ref_name)
        sigs[ref_name] = {"mirror": mirror, "members": collect(mod)}
    path = os.path.join(HERE, "signatures.json")
    with open(path, "w") as f:
        json.dump(sigs, f, indent=1, sort_keys=True)
    n = sum(len(v["members"]) for v in sigs.values())
    print(f"wrote {path}: {len(sigs)} modules, {n} members")


if __name__ == "__main__":
    main()
