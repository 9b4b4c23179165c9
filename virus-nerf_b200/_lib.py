"""ctypes binding of libvirusnerf_sm100.so (C ABI declared in include/virusnerf.h).

PyTorch is plumbing here: it owns device memory and streams; every computation on the hot
path is a kernel of the shared library.  Loading fails loudly when the library is missing --
there is no fallback implementation.
"""
import ctypes
import os
import subprocess

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# VN_LIB_PATH: development switch for A/B runs of an experimental build (make BUILD=build_exp OUT=... EXTRA=-D...)
LIB_PATH = os.environ.get("VN_LIB_PATH") or os.path.join(_PKG, "lib", "libvirusnerf_sm100.so")
CSRC = os.path.join(_PKG, "csrc")

VN_MAX_LEVELS = 32
VN_HASH_NO_WARP_AGG = 1
VN_HASH_LEVEL_GROUPS_1 = 16
VN_HASH_LEVEL_GROUPS_4 = 32
VN_HASH_LEVEL_GROUPS_8 = 64
VN_HASH_LEVEL_GROUPS_16 = 128
VN_HASH_LEVEL_GROUPS_2 = 256
VN_HASH_PLANAR = 512
VN_HASH_TIGHT_REGS = 1024
VN_HASH_PAIR_LOADS = 2048
VN_HASH_SKIP_ZERO_GRADS = 4096
VN_HASH_F16_CHUNKS = 8192
VN_HASH_FUSED_SCATTER = 16384
VN_STEP_SKIP_EXPAND = 8


class HashLevels(ctypes.Structure):
    """mirror of vn_hash_levels_t"""
    _fields_ = [
        ("levels", ctypes.c_int32),
        ("begin_fast_hash_level", ctypes.c_int32),
        ("offsets", ctypes.c_int32 * VN_MAX_LEVELS),
        ("sizes", ctypes.c_int32 * VN_MAX_LEVELS),
        ("scales", ctypes.c_float * VN_MAX_LEVELS),
        ("res", ctypes.c_uint32 * VN_MAX_LEVELS),
        ("total_entries", ctypes.c_int64),
        ("log_b", ctypes.c_double),
    ]


_P, _F, _I32, _I64, _D = ctypes.c_void_p, ctypes.c_float, ctypes.c_int32, ctypes.c_int64, ctypes.c_double


class Step(ctypes.Structure):
    """mirror of vn_step_t (include/virusnerf.h): every buffer of one train step"""
    _fields_ = (
        [("N", _I64)] +
        [(n, _P) for n in ("rays_o", "rays_d", "noise", "gt_rgb", "uss", "tof", "rgbd")] +
        [(n, _P) for n in ("hits_t", "counts", "rays_a", "counter", "scan_tmp", "bitfield")] +
        [("cascades", _I32), ("grid_size", _I32), ("max_samples", _I32), ("scale", _F), ("exp_step_factor", _F),
         ("T_threshold", _F), ("bg", _F), ("uss_tol", _F)] +
        [(n, _P) for n in ("xyzs", "dirs", "unit", "deltas", "ts", "enc", "sigmas", "rgbs", "ws", "d_sigmas", "d_rgbs",
                           "d_enc")] +
        [(n, _P) for n in ("vr_samples", "opacity", "depth", "rgb", "d_rgb", "d_depth", "d_opacity")] +
        [(n, _P) for n in ("flat_p", "flat_g", "flat_m", "flat_v")] + [("n_params", _I64)] +
        [("table_off", _I64), ("w_off", _I64 * 5)] +
        [("levels", HashLevels), ("hash_flags", _I32)] +
        [("loss_acc", _P), ("loss_out", _P)] +
        [("w_color", _F), ("w_uss", _F), ("w_tof", _F), ("w_rgbd", _F)] +
        [("scale_dev", _P), ("found_inf", _P), ("growth_tracker", _P)] +
        [("lr", _D), ("beta1", _D), ("beta2", _D), ("eps", _D), ("adam_step", _I32)] +
        [("ts_rows", _P), ("table_h", _P), ("step_dev", _P)])

    def set_ptrs(self, **tensors):
        for k, t in tensors.items():
            if t is None:
                setattr(self, k, None)
                continue
            if not (t.is_cuda and t.is_contiguous()):
                raise RuntimeError(f"vn_step_t.{k}: needs a contiguous CUDA tensor")
            setattr(self, k, t.data_ptr())


# spec characters: p device pointer (tensor / None), l int64, i int, f float, d double,
# h host pointer to a ctypes struct, s stream (filled in automatically)
_SPECS = {
    "vn_hash_levels_init": "ddilh",
    "vn_hash_encode_fwd_f32": "ppplhis",
    "vn_hash_encode_bwd_f32": "ppplhis",
    "vn_hash_encode_bwd_f32_levels": "ppplhiiis",
    "vn_hash_encode_fwd_f16": "ppplhis",
    "vn_hash_encode_bwd_f16": "ppplhis",
    "vn_f32_to_f16": "ppls",
    "vn_hash_indices": "plhpps",
    "vn_ray_aabb": "ppflps",
    "vn_march_train_count": "pppppliiffipppps",
    "vn_march_train_write": "ppppp" "liiff" "pl" "ppppp" "s",
    "vn_march_train_count_rows": "pppppliiffippppps",
    "vn_march_train_expand": "pppp" "liiff" "l" "ppppp" "s",
    "vn_march_train_expand_sh": "pppp" "liiff" "l" "ppppp" "pl" "s",
    "vn_march_test": "pppp" "lp" "iiffi" "ppppp" "s",
    "vn_march_test_compact": "plippppppppps",
    "vn_composite_train_fwd": "ppppp" "llf" "ppppp" "s",
    "vn_composite_train_bwd": "ppppp" "llf" "pppp" "pp" "s",
    "vn_composite_test": "pppppplfppps",
    "vn_sh_encode": "plps",
    "vn_morton3d": "plps",
    "vn_morton3d_invert": "plps",
    "vn_packbits": "plfps",
    "vn_occ_calc_pos_prob": "pppp" "liiif" "ffff" "ppppp" "s",
    "vn_occ_ray_prob": "pp" "lii" "fff" "pp" "s",
    "vn_occ_ray_prob_terms": "pp" "lii" "fff" "ppp" "s",
    "vn_occ_nerf_prob": "pldfppps",
    "vn_occ_bayes_update": "piplppp" "ps",
    "vn_occ_decay_pack": "pififps",
    "vn_occ_update": "pipp" "pppl" "pppl" "ii" "fffff" "df" "fif" "pp" "hi" "pp" "ff" "pl" "s",
    "vn_loss_fwd": "ppppppp" "lff" "pp" "s",
    "vn_composite_loss_fwd": "ppppp" "llf" "ppppp" "pppp" "ff" "pp" "s",
    "vn_composite_loss_bwd": "ppppp" "llf" "ppp" "pppp" "ff" "pp" "ffff" "p" "ppp" "s",
    "vn_loss_bwd": "ppppppp" "lff" "pp" "ffff" "p" "pppp" "s",
    "vn_grad_check": "plps",
    "vn_adam_step": "ppppl" "f" "dddd" "ipps",
    "vn_scaler_update": "pppffis",
    "vn_opt_state_init": "piddds",
    "vn_adam_step_dev": "ppppl" "dddd" "ppps",
    "vn_scaler_update_dev": "pppffi" "pddd" "s",
    "vn_umma_selftest": "iiippps",
    "vn_train_step_prepare": "hs",
    "vn_train_step_run": "hliis",
    "vn_train_step_expand": "hls",
    "vn_train_step_optim": "hs",
    "vn_p2p_allreduce": "ls",
    "vn_p2p_allreduce_small": "piis",
    "vn_p2p_reduce_adam": "lpp" "dddd" "i" "ppp" "s",
    "vn_p2p_step": "lpp" "dddd" "pppp" "s",
    "vn_batch_assemble": "ppl" "ppl" "pil" "pi" "pppp" "pp" "ppp" "pppp" "pp" "p" "s",
    "vn_pool_gather": "pplll" "pppppp" "pppppp" "p" "s",
    "vn_ngp_sample_occupied": "plfplpps",
    "vn_ngp_cell_positions": "ppliffps",
    "vn_ngp_grid_update": "ppplpplfps",
    "vn_ngp_threshold_pack": "plfppps",
    "vn_mlp_fwd": "pip" "ppppp" "li" "ppp" "s",
    "vn_mlp_bwd": "pip" "ppppp" "li" "pp" "p" "ppppp" "s",
    "vn_mlp_bwd_scatter": "pip" "ppppp" "l" "pp" "p" "hi" "p" "ppppp" "p" "s",
}

_CT = {"p": ctypes.c_void_p, "l": ctypes.c_int64, "i": ctypes.c_int, "f": ctypes.c_float,
       "d": ctypes.c_double, "h": ctypes.c_void_p, "s": ctypes.c_void_p}

_lib = None


def build(verbose=False):
    """compile csrc/*.cu for sm_100a into lib/libvirusnerf_sm100.so (nvcc cross-compiles
    without a GPU)"""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)], stdout=out)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"(or `make -C {CSRC}`).  virus-nerf_b200 has no CPU / PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.vn_last_error.restype = ctypes.c_char_p
        L.vn_abi_version.restype = ctypes.c_int
        L.vn_march_scan_tmp_ints.restype = ctypes.c_int64
        L.vn_ngp_select_tmp_ints.restype = ctypes.c_int64
        L.vn_ngp_threshold_tmp_bytes.restype = ctypes.c_int64
        L.vn_march_scan_tmp_ints.argtypes = [ctypes.c_int64]
        L.vn_launch_count.restype = ctypes.c_int64
        L.vn_occ_update_ws_floats.restype = ctypes.c_int64
        L.vn_occ_update_ws_floats.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        for name, spec in _SPECS.items():
            fn = getattr(L, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [_CT[c] for c in spec]
        _lib = L
    return _lib


def last_error():
    return lib().vn_last_error().decode("utf-8", "replace")


def _ptr(t, name, pos):
    if t is None:
        return None
    if isinstance(t, int):
        return ctypes.c_void_p(t)
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: argument {pos} must be a CUDA tensor or None, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: argument {pos} is on {t.device}; virus-nerf_b200 kernels need CUDA tensors "
                           "(there is no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: argument {pos} must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


KERNEL_NAMES = ["hash_encode_fwd", "hash_encode_bwd", "mlp_fwd", "mlp_bwd", "march_count", "march_write",
                "composite_fwd", "composite_bwd", "adam", "mlp_bwd_hash_scatter"]


def profile_start(kernels=None):
    """bracket the launches of the major kernels (all, or the named ones) with CUDA events on the
    launching stream (vn_profile_*)"""
    if kernels is None:
        lib().vn_profile_enable(1)
    else:
        mask = 0
        for k in kernels:
            mask |= 1 << KERNEL_NAMES.index(k)
        lib().vn_profile_enable_mask(ctypes.c_uint(mask))


def set_pdl(on):
    lib().vn_set_pdl(ctypes.c_int(1 if on else 0))


def profile_stop():
    """-> {kernel name: [(milliseconds, problem size), ...]}; synchronises"""
    L = lib()
    L.vn_profile_enable(0)
    out = {}
    kid, size, ms = ctypes.c_int(0), ctypes.c_int64(0), ctypes.c_float(0)
    for i in range(L.vn_profile_count()):
        if L.vn_profile_get(i, ctypes.byref(kid), ctypes.byref(size), ctypes.byref(ms)) != 0:
            raise RuntimeError(last_error())
        out.setdefault(KERNEL_NAMES[kid.value], []).append((ms.value, size.value))
    return out


def call(name, *args):
    """Invoke one C-ABI entry point on the current CUDA stream; raises RuntimeError with
    vn_last_error() on a non-zero return code."""
    L = lib()
    spec = _SPECS[name]
    conv = []
    it = iter(args)
    dev = None
    for pos, c in enumerate(spec):
        if c == "s":
            conv.append(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            continue
        try:
            a = next(it)
        except StopIteration:
            raise TypeError(f"{name}: expected {len(spec.replace('s', ''))} arguments, got {len(args)}")
        if c == "p":
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise RuntimeError(f"{name}: tensors on different devices ({dev} vs {a.device})")
            conv.append(_ptr(a, name, pos))
        elif c == "h":
            conv.append(ctypes.cast(ctypes.pointer(a), ctypes.c_void_p) if a is not None else None)
        elif c == "l":
            conv.append(ctypes.c_int64(int(a)))
        elif c == "i":
            conv.append(ctypes.c_int(int(a)))
        elif c == "f":
            conv.append(ctypes.c_float(float(a)))
        elif c == "d":
            conv.append(ctypes.c_double(float(a)))
    if len(list(it)) != 0:
        raise TypeError(f"{name}: too many arguments")
    fn = getattr(L, name)
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):
            conv = [ctypes.c_void_p(torch.cuda.current_stream().cuda_stream) if c == "s" else v
                    for c, v in zip(spec, conv)]
            rc = fn(*conv)
    else:
        rc = fn(*conv)
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {last_error()}")


def launch_count():
    """kernels launched by the library so far (for bench.py's gpu_launches)"""
    return int(lib().vn_launch_count())


def adam_config(lr, beta1, beta2, eps, step):
    """the six f32 constants of one Adam step (beta2, 1-beta1, 1-beta2, eps, step_size, bias_correction2_sqrt)
    as the kernels receive them (host only)"""
    out = (ctypes.c_float * 6)()
    L = lib()
    L.vn_adam_config.argtypes = [ctypes.c_double] * 4 + [ctypes.c_int, ctypes.c_void_p]
    rc = L.vn_adam_config(lr, beta1, beta2, eps, int(step), ctypes.cast(out, ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError(last_error())
    return list(out)


def ngp_select_tmp_ints(n_cells):
    return int(lib().vn_ngp_select_tmp_ints(ctypes.c_int64(int(n_cells))))


def ngp_threshold_tmp_bytes():
    return int(lib().vn_ngp_threshold_tmp_bytes())


def scan_tmp_ints(n):
    return int(lib().vn_march_scan_tmp_ints(ctypes.c_int64(int(n))))


def hash_levels(base_res, max_res, levels, max_params):
    """a1: host geometry (hash_encoder.py:183-208)"""
    lv = HashLevels()
    L = lib()
    rc = L.vn_hash_levels_init(ctypes.c_double(base_res), ctypes.c_double(max_res), ctypes.c_int(int(levels)),
                               ctypes.c_int64(int(max_params)), ctypes.cast(ctypes.pointer(lv), ctypes.c_void_p))
    if rc != 0:
        raise RuntimeError(f"vn_hash_levels_init failed (code {rc}): {last_error()}")
    return lv


def exported_symbols():
    """names declared in include/virusnerf.h (used by the CPU-side ABI test)"""
    return ["vn_last_error", "vn_abi_version", "vn_launch_count", "vn_ipc_get_handle", "vn_ipc_open", "vn_p2p_init", "vn_p2p_attach", "vn_p2p_shutdown", "vn_p2p_set_multicast", "vn_profile_enable", "vn_profile_enable_mask", "vn_set_pdl", "vn_profile_count", "vn_profile_get", "vn_device_info", "vn_march_scan_tmp_ints", "vn_adam_config",
            "vn_ngp_select_tmp_ints", "vn_ngp_threshold_tmp_bytes", "vn_occ_update_ws_floats"] + list(_SPECS)


def p2p_slice(n, rank, world):
    """[lo, hi) in floats of the slice of an n-float buffer (n % 4 == 0) that rank `rank` reduces and
    optimises in vn_p2p_reduce_adam / vn_p2p_allreduce: float4 chunks of ceil(n/4 / world), the last
    slices may be short or empty (csrc/p2p_allreduce.cu)"""
    n4 = n // 4
    chunk4 = (n4 + world - 1) // world
    lo = min(rank * chunk4, n4)
    return 4 * lo, 4 * min(lo + chunk4, n4)


def symmetric_empty(numel, device, group=None):
    """flat fp32 buffer in torch's symmetric memory (CUDA VMM allocation mapped into every rank, + an NVLink-multicast
    address when the fabric supports it).  Collective.  -> (tensor, peer pointers [world], multicast pointer or 0, handle)"""
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(int(numel), dtype=torch.float32, device=device)
    t.zero_()
    hdl = symm_mem.rendezvous(t, group if group is not None else dist.group.WORLD)
    mc = int(hdl.multicast_ptr) if hdl.multicast_ptr else 0
    return t, [int(x) for x in hdl.buffer_ptrs], mc, hdl


def p2p_setup_symmetric(grad_ptrs, param_ptrs, mc_grad, mc_params, flags, err, mbox, rank, world, group=None):
    """the peer-memory exchange over SYMMETRIC-MEMORY buffers: gradient / parameter peer pointers (and their multicast
    addresses) come from symmetric_empty(); the small flag / mailbox arrays are exchanged through CUDA IPC as in p2p_setup"""
    import torch.distributed as dist
    L = lib()

    def share(t):
        handle = ctypes.create_string_buffer(64)
        off = ctypes.c_int64(0)
        if L.vn_ipc_get_handle(ctypes.c_void_p(t.data_ptr()), handle, ctypes.byref(off)) != 0:
            raise RuntimeError(f"vn_ipc_get_handle failed: {last_error()}")
        return handle.raw, int(off.value)

    def open_peer(handle, off):
        base = _ipc_opened.get(handle)
        if base is None:
            out = ctypes.c_void_p()
            if L.vn_ipc_open(ctypes.c_char_p(handle), ctypes.c_int64(0), ctypes.byref(out)) != 0:
                raise RuntimeError(f"vn_ipc_open failed: {last_error()}")
            base = _ipc_opened[handle] = out.value
        return base + off

    mine = (share(flags), share(mbox))
    everyone = [None] * world
    dist.all_gather_object(everyone, mine, group=group)
    arrs = [(ctypes.c_void_p * world)() for _ in range(4)]
    for p in range(world):
        arrs[0][p] = grad_ptrs[p]
        arrs[1][p] = flags.data_ptr() if p == rank else open_peer(*everyone[p][0])
        arrs[2][p] = param_ptrs[p]
        arrs[3][p] = mbox.data_ptr() if p == rank else open_peer(*everyone[p][1])
    if L.vn_p2p_init(ctypes.c_int(rank), ctypes.c_int(world), arrs[0], arrs[1], ctypes.c_void_p(err.data_ptr())) != 0:
        raise RuntimeError(f"vn_p2p_init failed: {last_error()}")
    if L.vn_p2p_attach(arrs[2], arrs[3]) != 0:
        raise RuntimeError(f"vn_p2p_attach failed: {last_error()}")
    if mc_grad and mc_params:
        if L.vn_p2p_set_multicast(ctypes.c_void_p(mc_grad), ctypes.c_void_p(mc_params)) != 0:
            raise RuntimeError(f"vn_p2p_set_multicast failed: {last_error()}")
    dist.barrier(group=group)
    return arrs


def p2p_shutdown():
    """release the process-wide peer-memory exchange context (the next p2p_setup may then take it)"""
    lib().vn_p2p_shutdown()


_ipc_opened = {}   # 64-byte IPC handle -> mapped base pointer (a handle may be opened once per process)


def p2p_setup(grad, flags, err, rank, world, group=None, params=None, mbox=None):
    """Exchange CUDA-IPC handles of `grad` (flat fp32 gradient buffer) and `flags` (int32 [world],
    zeros) -- and, for the fused sharded optimiser, of `params` (flat fp32 parameter buffer) and
    `mbox` (fp32 zeros [2, world, 8]) -- across the ranks of a torch.distributed group and
    initialise the peer-memory exchange (vn_p2p_init / vn_p2p_attach).  Returns the lists of
    opened peer pointers (kept alive by the caller)."""
    import torch.distributed as dist
    L = lib()

    def share(t):
        handle = ctypes.create_string_buffer(64)
        off = ctypes.c_int64(0)
        rc = L.vn_ipc_get_handle(ctypes.c_void_p(t.data_ptr()), handle, ctypes.byref(off))
        if rc != 0:
            raise RuntimeError(f"vn_ipc_get_handle failed (code {rc}): {last_error()}")
        return handle.raw, int(off.value)

    def open_peer(handle, off):
        base = _ipc_opened.get(handle)
        if base is None:     # tensors that share one cudaMalloc block of the caching allocator share the handle
            out = ctypes.c_void_p()
            rc = L.vn_ipc_open(ctypes.c_char_p(handle), ctypes.c_int64(0), ctypes.byref(out))
            if rc != 0:
                raise RuntimeError(f"vn_ipc_open failed (code {rc}): {last_error()}")
            base = _ipc_opened[handle] = out.value
        return base + off

    tensors = [grad, flags] + ([params, mbox] if mbox is not None else [])
    if mbox is not None and params is None:
        raise ValueError("p2p_setup: mbox needs params")
    mine = tuple(share(t) for t in tensors)
    everyone = [None] * world
    dist.all_gather_object(everyone, mine, group=group)
    arrs = [(ctypes.c_void_p * world)() for _ in tensors]
    for p in range(world):
        for k, t in enumerate(tensors):
            arrs[k][p] = t.data_ptr() if p == rank else open_peer(*everyone[p][k])
    rc = L.vn_p2p_init(ctypes.c_int(rank), ctypes.c_int(world), arrs[0], arrs[1], ctypes.c_void_p(err.data_ptr()))
    if rc != 0:
        raise RuntimeError(f"vn_p2p_init failed (code {rc}): {last_error()}")
    if mbox is not None:
        rc = L.vn_p2p_attach(arrs[2], arrs[3])
        if rc != 0:
            raise RuntimeError(f"vn_p2p_attach failed (code {rc}): {last_error()}")
    dist.barrier(group=group)
    return arrs
