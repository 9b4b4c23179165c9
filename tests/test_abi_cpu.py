"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/virusnerf.h declares, its host-only entry points work, argument errors follow
the conventions (negative code + vn_last_error), and the Python modules refuse CPU tensors
(there is no fallback path)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vn():
    from virus_nerf_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    _lib.lib()
    return _lib


def _declared_symbols():
    h = open(os.path.join(ROOT, "include", "virusnerf.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(vn_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol(vn):
    L = vn.lib()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/virusnerf.h but not exported"
    assert set(vn.exported_symbols()) == set(names)          # the ctypes table covers the whole header
    assert L.vn_abi_version() == 1


def test_ctypes_call_specs_match_the_header_prototypes(vn):
    """every entry of _lib._SPECS (one character per argument) against the parameter list of the
    prototype in include/virusnerf.h: pointer / int64 / int / float / double / stream, in order"""
    h = open(os.path.join(ROOT, "include", "virusnerf.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    protos = dict(re.findall(r"\bint\s+(vn_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", h, flags=re.S))

    def kind(param):
        param = " ".join(param.split())
        name = param.split()[-1].lstrip("*")
        if "*" in param:
            if name == "stream":
                return "s"
            return "h" if name.startswith("h_") else "p"
        t = param.rsplit(" ", 1)[0].replace("const ", "")
        return {"int64_t": "l", "int": "i", "int32_t": "i", "unsigned": "i", "float": "f", "double": "d"}[t]

    checked = 0
    for name, spec in vn._SPECS.items():
        assert name in protos, name
        params = [q for q in protos[name].split(",") if q.strip() and q.strip() != "void"]
        kinds = "".join(kind(q) for q in params)
        # host-struct pointers are declared with an h_ prefix or as const vn_*_t* (passed as 'h')
        norm = lambda k: k.replace("h", "p")
        assert norm(kinds) == norm(spec), (name, kinds, spec)
        checked += 1
    assert checked >= 45


def test_adam_constants_follow_torch_rounding(vn):
    """vn_adam_config (host): python-double arithmetic on the hyper-parameters, one rounding to f32 -- what
    torch.optim.Adam's single-tensor path feeds its kernels (1 - beta2 is 0.001f, not 1.0f - 0.999f)"""
    import math
    lr, b1, b2, eps = 5e-3, 0.9, 0.999, 1e-15
    for step in (1, 2, 7, 100, 5000):
        got = np.array(vn.adam_config(lr, b1, b2, eps, step), np.float32)
        want = np.array([b2, 1 - b1, 1 - b2, eps, lr / (1 - b1 ** step), math.sqrt(1 - b2 ** step)]).astype(np.float32)
        np.testing.assert_array_equal(got, want)
    assert np.float32(1 - b2) != np.float32(1) - np.float32(b2)        # the distinction this entry point exists for


def test_next_row_mirrors_expose_the_reference_api():
    from virus_nerf_b200.modules import ngp_grid, networks
    from virus_nerf_b200.datasets import dataset_base
    from virus_nerf_b200.training import sampler, evaluation
    for name in ("sample_uniform_and_occupied_cells", "mark_invisible_cells", "update", "getBitfield", "getAllCells",
                 "updateBitfield", "morton2bitfield"):
        assert hasattr(ngp_grid.NGPGrid, name), name
    assert hasattr(networks.NGP, "updateNeRFGrid") and hasattr(networks.NGP, "updateOccGrid")
    for name in ("__call__", "_calcRayPoses", "to", "getMeanHeight", "__len__"):
        assert hasattr(dataset_base.DatasetBase, name), name
    for name in ("__call__", "getValidImgIdxs", "_imgIdxs", "_pixIdxs", "_pixStrategyRandom", "_pixStrategyEntireImg",
                 "_pixStrategyClosest", "_pixStrategyValidDepth"):
        assert hasattr(sampler.Sampler, name), name
    for name in ("createScanRays", "createScanPos", "batchify_render", "batchify_density", "interfere_density_map",
                 "evaluation_depth_nerf", "save_checkpoint", "load_checkpoint"):
        assert callable(getattr(evaluation, name)), name


def test_only_sm100a_code_in_the_library(vn):
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", vn.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_hash_levels_host_entry_matches_oracle(vn, oracle_mod):
    for max_res, log2_T in ((1024, 19), (1024, 22), (2048, 19), (512, 14)):
        lv = vn.hash_levels(16, max_res, 16, 2 ** log2_T)
        o = oracle_mod.HashLevels(16, max_res, 16, 2 ** log2_T)
        assert lv.total_entries == o.total and lv.begin_fast_hash_level == o.begin_fast_hash_level
        np.testing.assert_array_equal(np.array(lv.offsets[:16]), o.offsets)
        np.testing.assert_array_equal(np.array(lv.sizes[:16]), o.sizes)
        np.testing.assert_array_equal(np.array(lv.scales[:16], np.float32), o.scales)
        np.testing.assert_array_equal(np.array(lv.res[:16], np.uint32), o.res)
        assert lv.log_b == o.log_b


def test_error_conventions_without_gpu(vn):
    L = vn.lib()
    lv = vn.HashLevels()
    rc = L.vn_hash_levels_init(ctypes.c_double(16), ctypes.c_double(1024), 1, ctypes.c_int64(2 ** 19),
                               ctypes.cast(ctypes.pointer(lv), ctypes.c_void_p))
    assert rc == -1 and "levels" in vn.last_error()
    with pytest.raises(RuntimeError, match="levels"):
        vn.hash_levels(16, 1024, 99, 2 ** 19)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vn.call("vn_sh_encode", torch.zeros(4, 3), 4, torch.zeros(4, 16))
    with pytest.raises(TypeError):
        vn.call("vn_packbits")


def test_modules_mirror_reference_api_and_refuse_cpu(vn):
    from virus_nerf_b200.modules import (hash_encoder, hash_encoder_half, intersection, networks, occupancy_grid,
                                         ray_march, rendering, spherical_harmonics, utils, volume_render_test,
                                         volume_train)
    assert rendering.MAX_SAMPLES == utils.MAX_SAMPLES == 1024 and utils.NEAR_DISTANCE == 0.01
    enc = hash_encoder.HashEncoder(max_params=2 ** 19, levels=16, base_res=16, max_res=1024)
    assert (enc.out_dim, enc.total_param_size, enc.begin_fast_hash_level) == (32, 11420064, 6)
    assert list(enc.state_dict().keys()) == ["hash_table"]                      # offsets / sizes are non-persistent
    assert 0.0 <= float(enc.hash_table.detach().min()) and float(enc.hash_table.detach().max()) <= 1.0   # U(0,1) init
    half = hash_encoder_half.HashEncoder(max_params=2 ** 14, levels=16, base_res=16, max_res=512)
    assert half.hash_table.dim() == 2 and set(half.state_dict().keys()) == {"hash_table", "hash_grad"}
    assert float(half.hash_table.abs().max()) <= 1e-4
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(torch.rand(8, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        intersection.ray_aabb_intersection(torch.rand(8, 3), torch.rand(8, 3), 0.5)
    for fn in (ray_march.raymarching_train, ray_march.raymarching_test, volume_render_test.composite_test,
               rendering.render, utils.morton3D, utils.morton3D_invert, utils.packbits):
        assert callable(fn)
    assert spherical_harmonics.DirEncoder().out_dim == 16
    assert hasattr(volume_train.VolumeRenderer(), "forward")
    m = networks.MLP(input_dim=32, output_dim=16, net_depth=1, net_width=64, bias_enabled=False)
    assert [tuple(p.shape) for p in m.parameters()] == [(64, 32), (16, 64)]
    for name in ("update", "getBitfield", "_rayUpdate", "_nerfUpdate", "_calcPos", "_rayProb", "_nerfProb",
                 "_updateGrid", "_c2idx", "_idx2c", "getOccupancyCartesianGrid", "getBinaryCartesianGrid",
                 "bitfield2morton", "morton2cartesian", "c2oCoordinates", "cartesian2morton", "morton2bitfield"):
        assert hasattr(occupancy_grid.OccupancyGrid, name), name


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "virus-nerf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
                assert "liboracle" not in src


def test_ctypes_struct_mirrors_match_the_header(vn, tmp_path):
    """vn_hash_levels_t / vn_step_t as laid out by the C compiler vs the ctypes mirrors in _lib.py"""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "virusnerf.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(vn_hash_levels_t), sizeof(vn_step_t),'
                   ' offsetof(vn_step_t, levels), offsetof(vn_step_t, loss_acc), offsetof(vn_step_t, adam_step),'
                   ' offsetof(vn_step_t, w_off), offsetof(vn_step_t, ts_rows)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    c = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    py = [ctypes.sizeof(vn.HashLevels), ctypes.sizeof(vn.Step), vn.Step.levels.offset, vn.Step.loss_acc.offset,
          vn.Step.adam_step.offset, vn.Step.w_off.offset, vn.Step.ts_rows.offset]
    assert c == py, (c, py)


# ---------------------------------------------------------------------------------------------------------------------
# names of the reference that have no host-side mirror, with the reason (everything else in signatures.json must exist)
NOT_MIRRORED = {
    # Taichi field <-> torch copies: the product keeps every tensor in torch device memory, there is no second runtime
    "modules.utils": {"ti2torch", "ti2torch_grad", "ti2torch_vec", "ti2torch_grad_vec", "torch2ti", "torch2ti_grad",
                      "torch2ti_vec", "torch2ti_grad_vec", "random_initialize",
                      # @ti.func device helpers: they live inside the CUDA kernels (csrc/common.cuh)
                      "calc_dt", "mip_from_dt", "mip_from_pos", "frexp_bit", "scalbn",
                      # @ti.kernel bodies behind morton3D / morton3D_invert: replaced by C-ABI entries
                      "morton3D_kernel", "morton3D_invert_kernel",
                      # deployment export / plotting helpers outside the hot path (SURVEY section 8 "out of scope")
                      "save_deployment_model", "depth2img"},
    # kernel factories / raw kernels replaced by C-ABI entries (include/virusnerf.h cites each)
    "modules.hash_encoder": {"build_hash_encoder_kernel"},
    "modules.hash_encoder_half": {"build_hash_encoder_kernel"},
    "modules.intersection": {"ray_aabb_intersect"},
    "modules.ray_march": {"raymarching_train_kernel", "raymarching_test_kernel"},
    "modules.spherical_harmonics": {"dir_encoder"},
    "modules.volume_train": {"volume_rendering_kernel"},
}
# methods that exist in the reference class but are plotting / dataset-file plumbing outside the path
METHODS_NOT_MIRRORED = {
    ("datasets.dataset_base", "DatasetBase"): {"getMeanHeight", "getSyncIdxs", "reduceImgHeight"},
    ("modules.networks", "TruncExp"): {"forward", "backward"},       # fused into the MLP epilogue (kernel side)
}


def _params_of(fn):
    import inspect
    fn = getattr(fn, "__func__", fn)
    return [(p.name, p.kind.name, p.default) for p in inspect.signature(fn).parameters.values()]


def _check_params(where, ref_params, mine, errors):
    """same parameter names in the same order and the same literal defaults; the mirror may append optional parameters
    and may give a default to a parameter the reference requires (both keep every reference call valid)"""
    import inspect
    names = [p[0] for p in mine]
    ref_names = [p[0] for p in ref_params]
    if any(p[1] in ("VAR_POSITIONAL", "VAR_KEYWORD") for p in ref_params):
        return
    if names[:len(ref_names)] != ref_names:
        errors.append(f"{where}: reference {ref_names} vs mirror {names}")
        return
    for p in mine[len(ref_names):]:
        if p[2] is inspect.Parameter.empty and p[1] not in ("VAR_POSITIONAL", "VAR_KEYWORD"):
            errors.append(f"{where}: extra parameter {p[0]} of the mirror has no default")
    for (rn, _, rdef), (mn, _, mdef) in zip(ref_params, mine):
        if rdef not in ("<required>", "<object>") and repr(mdef) != rdef:
            errors.append(f"{where}: default of {rn}: reference {rdef} vs mirror {mdef!r}")


def test_mirrors_match_the_reference_signatures():
    """every public function / class / method the reference defines on the path exists in the mirror package with the
    same parameter list (tests/golden/signatures.json = inspect.signature over /root/reference, made by
    tests/golden/make_signatures.py)"""
    import importlib
    import inspect
    import json
    sigs = json.load(open(os.path.join(ROOT, "tests", "golden", "signatures.json")))
    checked, errors = 0, []
    for ref_mod, entry in sigs.items():
        mod = importlib.import_module("virus_nerf_b200." + entry["mirror"])
        skip = NOT_MIRRORED.get(ref_mod, set())
        for name, d in entry["members"].items():
            if name in skip:
                continue
            if not hasattr(mod, name):
                errors.append(f"{ref_mod}.{name} has no mirror in virus_nerf_b200.{entry['mirror']}")
                continue
            obj = getattr(mod, name)
            if d["kind"] == "function":
                _check_params(f"{ref_mod}.{name}", d["params"], _params_of(obj), errors)
                checked += 1
                continue
            assert inspect.isclass(obj), f"{ref_mod}.{name} must be a class"
            mskip = METHODS_NOT_MIRRORED.get((ref_mod, name), set())
            for mname, mparams in d["methods"].items():
                if mname in mskip or mparams is None:
                    continue
                if not hasattr(obj, mname):
                    errors.append(f"{ref_mod}.{name}.{mname} is missing from the mirror")
                    continue
                raw = inspect.getattr_static(obj, mname)
                fn = raw.__func__ if isinstance(raw, (staticmethod, classmethod)) else raw
                _check_params(f"{ref_mod}.{name}.{mname}", mparams, _params_of(fn), errors)
                checked += 1
    assert not errors, "\n".join(errors)
    assert checked >= 70, checked


def test_pdl_rule_no_noncoherent_loads_of_predecessor_data():
    """csrc/common.cuh, RULE at vn_launch_pdl: arrays that the previous kernels of the stream produce (encoding and its
    gradient, MLP outputs and their gradients, the compositor's per-ray outputs, the flat gradient read by the optimiser)
    must not be read with __ldg (ld.global.nc): an invariant load may be scheduled above griddepcontrol.wait.  A static
    check of the kernels' sources (found as a real failure in round 2: profiles/r2_kbench.md)."""
    csrc = os.path.join(ROOT, "virus-nerf_b200", "csrc")
    banned = re.compile(r"__ldg\(\s*(?:\([^()]*\)\s*)?\(?\s*(a\.enc|a\.dsigmas|a\.drgbs|sigmas\b|rgbs\b|dout\b|dp\b|dL_d\w+|g4\b|g \+|"
                        r"opacity\b|depth\b|rgb \+|src \+ \(int64_t\)q|src \+ q|src\b\))")
    offenders = []
    for name in ("mlp_fused.cu", "mlp_bwd_pipe.cu", "hash_encoder.cu", "optim.cu", "loss_common.cuh"):
        for i, line in enumerate(open(os.path.join(csrc, name)), 1):
            if banned.search(line):
                offenders.append(f"{name}:{i}: {line.strip()[:100]}")
    # composite.cu: the training kernels (everything above the test-time compositor, which is not PDL-launched)
    comp = open(os.path.join(csrc, "composite.cu")).read()
    train = comp[:comp.index("// ---- a10 ----")]
    for i, line in enumerate(train.split("\n"), 1):
        if banned.search(line):
            offenders.append(f"composite.cu:{i}: {line.strip()[:100]}")
    assert not offenders, "\n".join(offenders)
    # and the rule is written down where the launches are defined
    assert "RULE" in open(os.path.join(csrc, "common.cuh")).read()


def test_roofline_traffic_comes_from_the_committed_ncu_capture(tmp_path):
    """bench.py's roofline.traffic = DRAM bytes per point of this round's `ncu --set full` capture x points per launch:
    profiles/r2_ncu_traffic.json must be what tools/ncu_traffic.py derives from the committed raw CSVs"""
    import json
    import subprocess
    import sys
    out = tmp_path / "t.json"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"),
                           os.path.join(ROOT, "profiles", "r2_ncu_step_raw.csv"), "--json", str(out), "--points", "853812"],
                          stdout=subprocess.DEVNULL)
    fresh = json.load(open(out))
    committed = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
    for k in ("mlp_bwd_hash_scatter", "hash_encode_fwd", "mlp_fwd", "adam"):
        assert abs(fresh[k]["dram_bytes_per_point"] - committed[k]["dram_bytes_per_point"]) < 1e-6, k
    # the fused backward moves far fewer DRAM bytes than its 1148 algorithmic bytes per point (table gradient in L2)
    assert 100 < committed["mlp_bwd_hash_scatter"]["dram_bytes_per_point"] < 400
