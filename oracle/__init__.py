"""ctypes/numpy front end of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs -- never by the product package.  See the header of
oracle.cpp for the parity status and the reference file:line each function follows.
"""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_DIR, "liboracle.so")

MAX_SAMPLES = 1024


def build(force=False):
    src = os.path.join(_DIR, "oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.vo_hash_levels.restype = ctypes.c_int64
        _lib.vo_march_train_write.restype = ctypes.c_int64
        _lib.vo_adam_step.restype = ctypes.c_int
        _lib.vo_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


f32, i32, i64, u8, u32 = np.float32, np.int32, np.int64, np.uint8, np.uint32
cf = ctypes.c_float
cd = ctypes.c_double
ci = ctypes.c_int
cl = ctypes.c_int64


def num_threads():
    return int(lib().vo_num_threads())


class HashLevels:
    """a1: per-level geometry (hash_encoder.py:183-208, 73-80; utils.py:19-42)."""

    def __init__(self, base_res=16.0, max_res=1024.0, levels=16, max_params=2 ** 19):
        self.levels = levels
        self.offsets = np.zeros(levels, i32)
        self.sizes = np.zeros(levels, i32)
        self.scales = np.zeros(levels, f32)
        self.res = np.zeros(levels, u32)
        self.res_host = np.zeros(levels, np.float64)
        bf = ctypes.c_int32(0)
        lb = ctypes.c_double(0)
        self.total = int(lib().vo_hash_levels(
            ctypes.c_double(base_res), ctypes.c_double(max_res), ci(levels), cl(int(max_params)),
            _p(self.offsets), _p(self.sizes), _p(self.scales), _p(self.res), _p(self.res_host),
            ctypes.byref(bf), ctypes.byref(lb)))
        self.begin_fast_hash_level = int(bf.value)
        self.log_b = float(lb.value)

    def _args(self):
        return (ci(self.levels), _p(self.offsets), _p(self.sizes), _p(self.scales), _p(self.res),
                ci(self.begin_fast_hash_level))


def hash_indices(xyz, lv, with_weights=False):
    xyz = _c(xyz, f32)
    S = xyz.shape[0]
    idx = np.zeros((S, lv.levels, 8), i32)
    w = np.zeros((S, lv.levels, 8), f32) if with_weights else None
    lib().vo_hash_indices(_p(xyz), cl(S), ci(lv.levels), _p(lv.sizes), _p(lv.scales), _p(lv.res),
                          ci(lv.begin_fast_hash_level), _p(idx), _p(w))
    return (idx, w) if with_weights else idx


def hash_fwd_f32(xyz, table, lv, threads=1):
    xyz, table = _c(xyz, f32), _c(table, f32)
    S = xyz.shape[0]
    out = np.zeros((S, 2 * lv.levels), f32)
    lib().vo_hash_fwd_f32(_p(xyz), _p(table), _p(out), cl(S), *lv._args(), ci(threads))
    return out


def hash_bwd_f32(xyz, dout, lv, threads=1):
    xyz, dout = _c(xyz, f32), _c(dout, f32)
    S = xyz.shape[0]
    grad = np.zeros(2 * lv.total, f32)
    lib().vo_hash_bwd_f32(_p(xyz), _p(dout), _p(grad), cl(S), *lv._args(), ci(threads))
    return grad


def hash_fwd_f16(xyz, table_h, lv, threads=1):
    xyz = _c(xyz, f32)
    table_h = _c(table_h, np.float16)
    S = xyz.shape[0]
    out = np.zeros((S, 2 * lv.levels), np.float16)
    lib().vo_hash_fwd_f16(_p(xyz), _p(table_h), _p(out), cl(S), *lv._args(), ci(threads))
    return out


def hash_bwd_f16(xyz, dout_h, lv, threads=1):
    xyz = _c(xyz, f32)
    dout_h = _c(dout_h, np.float16)
    S = xyz.shape[0]
    grad = np.zeros((lv.total, 2), f32)
    lib().vo_hash_bwd_f16(_p(xyz), _p(dout_h), _p(grad), cl(S), *lv._args(), ci(threads))
    return grad


def ray_aabb(rays_o, rays_d, scale):
    rays_o, rays_d = _c(rays_o, f32), _c(rays_d, f32)
    N = rays_o.shape[0]
    hits = np.zeros((N, 2), f32)
    lib().vo_ray_aabb(_p(rays_o), _p(rays_d), cf(scale), cl(N), _p(hits))
    return hits


def morton3d(coords):
    coords = _c(coords, i32)
    out = np.zeros(coords.shape[0], i32)
    lib().vo_morton3d(_p(coords), cl(coords.shape[0]), _p(out))
    return out


def morton3d_invert(indices):
    indices = _c(indices, i32)
    out = np.zeros((indices.shape[0], 3), i32)
    lib().vo_morton3d_invert(_p(indices), cl(indices.shape[0]), _p(out))
    return out


def packbits(grid, thr):
    grid = _c(grid, f32).reshape(-1)
    out = np.zeros(grid.shape[0] // 8, u8)
    lib().vo_packbits(_p(grid), cl(out.shape[0]), cf(thr), _p(out))
    return out


def march_train(rays_o, rays_d, hits_t, bitfield, noise, cascades, scale, esf, grid_size,
                max_samples=MAX_SAMPLES, threads=1):
    """a6: returns (rays_a [N,3], xyzs, dirs, deltas, ts, total) in canonical ray order."""
    rays_o, rays_d, hits_t = _c(rays_o, f32), _c(rays_d, f32), _c(hits_t, f32)
    bitfield, noise = _c(bitfield, u8), _c(noise, f32)
    N = rays_o.shape[0]
    counts = np.zeros(N, i32)
    lib().vo_march_train_count(_p(rays_o), _p(rays_d), _p(hits_t), _p(bitfield), _p(noise), cl(N),
                               ci(cascades), ci(grid_size), cf(scale), cf(esf), cf(max_samples),
                               _p(counts), ci(threads))
    total = int(counts.sum())
    rays_a = np.zeros((N, 3), i32)
    xyzs = np.zeros((total, 3), f32)
    dirs = np.zeros((total, 3), f32)
    deltas = np.zeros(total, f32)
    ts = np.zeros(total, f32)
    t2 = lib().vo_march_train_write(_p(rays_o), _p(rays_d), _p(hits_t), _p(bitfield), _p(noise), cl(N),
                                    ci(cascades), ci(grid_size), cf(scale), cf(esf), _p(counts),
                                    _p(rays_a), _p(xyzs), _p(dirs), _p(deltas), _p(ts), ci(threads))
    assert t2 == total
    return rays_a, xyzs, dirs, deltas, ts, total


def march_test(rays_o, rays_d, hits_t, alive, bitfield, cascades, scale, esf, grid_size, max_samples):
    """a7: kernel + wrapper (ray_march.py:271-335).  hits_t is mutated in place (must be a
    contiguous float32 array).  Returns (packed_info [A,2] i64, ray_indices, deltas, ts)."""
    rays_o, rays_d = _c(rays_o, f32), _c(rays_d, f32)
    assert hits_t.dtype == f32 and hits_t.flags["C_CONTIGUOUS"]
    alive = _c(alive, i64)
    bitfield = _c(bitfield, u8)
    A = alive.shape[0]
    ray_indices = np.zeros(A * max_samples, i64)
    valid = np.zeros(A * max_samples, u8)
    deltas = np.zeros(A * max_samples, f32)
    ts = np.zeros(A * max_samples, f32)
    counter = np.zeros(A, i32)
    lib().vo_march_test(_p(rays_o), _p(rays_d), _p(hits_t), _p(alive), cl(A), _p(bitfield), ci(cascades),
                        ci(grid_size), cf(scale), cf(esf), ci(max_samples), _p(ray_indices), _p(valid),
                        _p(deltas), _p(ts), _p(counter))
    valid = valid.astype(bool)
    cumsum = np.cumsum(counter.astype(i64))
    packed = np.stack([cumsum - counter, counter.astype(i64)], axis=-1)
    return packed, ray_indices[valid], deltas[valid], ts[valid]


def composite_train_fwd(sigmas, rgbs, deltas, ts, rays_a, T_thr, threads=1):
    sigmas, rgbs, deltas, ts = _c(sigmas, f32), _c(rgbs, f32), _c(deltas, f32), _c(ts, f32)
    rays_a = _c(rays_a, i32)
    N = rays_a.shape[0]
    total = np.zeros(N, i32)
    opacity = np.zeros(N, f32)
    depth = np.zeros(N, f32)
    rgb = np.zeros((N, 3), f32)
    ws = np.zeros(sigmas.shape[0], f32)
    lib().vo_composite_train_fwd(_p(sigmas), _p(rgbs), _p(deltas), _p(ts), _p(rays_a), cl(N), cf(T_thr),
                                 _p(total), _p(opacity), _p(depth), _p(rgb), _p(ws), ci(threads))
    return total, opacity, depth, rgb, ws


def composite_train_bwd(sigmas, rgbs, deltas, ts, rays_a, T_thr, dL_dopacity, dL_ddepth, dL_drgb,
                        dL_dws=None, threads=1):
    sigmas, rgbs, deltas, ts = _c(sigmas, f32), _c(rgbs, f32), _c(deltas, f32), _c(ts, f32)
    rays_a = _c(rays_a, i32)
    dO, dD, dC = _c(dL_dopacity, f32), _c(dL_ddepth, f32), _c(dL_drgb, f32)
    dW = None if dL_dws is None else _c(dL_dws, f32)
    N = rays_a.shape[0]
    dsig = np.zeros_like(sigmas)
    drgb = np.zeros_like(rgbs)
    lib().vo_composite_train_bwd(_p(sigmas), _p(rgbs), _p(deltas), _p(ts), _p(rays_a), cl(N), cf(T_thr),
                                 _p(dO), _p(dD), _p(dC), _p(dW), _p(dsig), _p(drgb), ci(threads))
    return dsig, drgb


def composite_test(sigmas, rgbs, deltas, ts, pack_info, alive, T_thr, opacity, depth, rgb):
    """a10: in place on alive / opacity / depth / rgb (contiguous arrays of the right dtype)."""
    sigmas, rgbs, deltas, ts = _c(sigmas, f32), _c(rgbs, f32), _c(deltas, f32), _c(ts, f32)
    pack_info = _c(pack_info, i64)
    for a, dt in ((alive, i64), (opacity, f32), (depth, f32), (rgb, f32)):
        assert a.dtype == dt and a.flags["C_CONTIGUOUS"]
    lib().vo_composite_test(_p(sigmas), _p(rgbs), _p(deltas), _p(ts), _p(pack_info), _p(alive),
                            cl(alive.shape[0]), cf(T_thr), _p(opacity), _p(depth), _p(rgb))


def sh_encode(dirs):
    dirs = _c(dirs, f32)
    out = np.zeros((dirs.shape[0], 16), f32)
    lib().vo_sh_encode(_p(dirs), cl(dirs.shape[0]), _p(out))
    return out


def mlp_fwd(enc, dirs, W1, W2, W3, W4, W5, threads=1, return_h=False):
    enc, dirs = _c(enc, f32), _c(dirs, f32)
    Ws = [_c(w, f32) for w in (W1, W2, W3, W4, W5)]
    S = enc.shape[0]
    sig = np.zeros(S, f32)
    rgb = np.zeros((S, 3), f32)
    h = np.zeros((S, 16), f32) if return_h else None
    lib().vo_mlp_fwd(_p(enc), _p(dirs), cl(S), *[_p(w) for w in Ws], _p(sig), _p(rgb), _p(h), ci(threads))
    return (sig, rgb, h) if return_h else (sig, rgb)


def dist_to_cube_border(rays_o, rays_d, cube_min, cube_max):
    rays_o, rays_d = _c(rays_o, f32), _c(rays_d, f32)
    out = np.zeros(rays_o.shape[0], f32)
    lib().vo_dist_to_cube_border(_p(rays_o), _p(rays_d), cl(rays_o.shape[0]), cf(cube_min), cf(cube_max),
                                 _p(out))
    return out


def occ_calc_pos(rays_o, rays_d, noise, M, grid_size, scale, noise_every_m):
    rays_o, rays_d = _c(rays_o, f32), _c(rays_d, f32)
    noise = None if noise is None else _c(noise, f32)
    N = rays_o.shape[0]
    dists = np.zeros((N, M), f32)
    pos = np.zeros((N * M, 3), f32)
    idxs = np.zeros((N * M, 3), i32)
    lib().vo_occ_calc_pos(_p(rays_o), _p(rays_d), _p(noise), cl(N), ci(M), ci(grid_size), cf(scale),
                          cf(noise_every_m), _p(dists), _p(pos), _p(idxs))
    return dists, pos, idxs


def occ_ray_prob(meas, dists, p_false, std_every_m, I=32, prob_min=0.03):
    meas, dists = _c(meas, f32), _c(dists, f32)
    N, M = dists.shape
    po = np.zeros((N, M), f32)
    pe = np.zeros((N, M), f32)
    lib().vo_occ_ray_prob(_p(meas), _p(dists), cl(N), ci(M), ci(I), cf(p_false), cf(std_every_m),
                          cf(prob_min), _p(po), _p(pe))
    return po, pe


def occ_nerf_prob(density, thr_max, slope):
    density = _c(density, f32)
    po = np.zeros_like(density)
    pe = np.zeros_like(density)
    lib().vo_occ_nerf_prob(_p(density), cl(density.shape[0]), ctypes.c_double(thr_max), cf(slope), _p(po), _p(pe))
    return po, pe


def occ_update_grid(grid, cell_idxs, probs_occ, probs_emp):
    """in place on grid [G,G,G] float32 contiguous"""
    assert grid.dtype == f32 and grid.flags["C_CONTIGUOUS"]
    cell_idxs = _c(cell_idxs, i32)
    po, pe = _c(probs_occ, f32).reshape(-1), _c(probs_emp, f32).reshape(-1)
    lib().vo_occ_update_grid(_p(grid), ci(grid.shape[0]), _p(cell_idxs), cl(cell_idxs.shape[0]), _p(po), _p(pe))


def occ_decay_pack(grid, decay, apply_decay, thr):
    """in place decay on grid; returns the Morton bitfield"""
    assert grid.dtype == f32 and grid.flags["C_CONTIGUOUS"]
    G = grid.shape[0]
    bf = np.zeros(G ** 3 // 8, u8)
    lib().vo_occ_decay_pack(_p(grid), ci(G), cf(decay), ci(1 if apply_decay else 0), cf(thr), _p(bf))
    return bf


def adam_step(p, g, m, v, inv_scale, lr, beta1, beta2, eps, step):
    for a in (p, m, v):
        assert a.dtype == f32 and a.flags["C_CONTIGUOUS"]
    g = _c(g, f32)
    return int(lib().vo_adam_step(_p(p), _p(g), _p(m), _p(v), cl(p.size), cf(inv_scale), cd(lr),
                                  cd(beta1), cd(beta2), cd(eps), ci(step)))
