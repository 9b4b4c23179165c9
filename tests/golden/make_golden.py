"""Generates tests/golden/golden_v1.npz by running THE REFERENCE'S OWN CODE from
/root/reference (read-only, only available in the build container):

  * every Taichi @ti.kernel on the hot path is executed from its original source under the
    pure-Python f32 shim in tests/golden/ti_shim (Taichi is not installable here);
  * the pure-torch parts (modules/occupancy_grid.py, helpers/geometric_fcts.py,
    training/loss.py) are imported and run as they are (third-party imports that are absent
    from the image -- kornia, alive_progress, ... -- are stubbed; none of them is on the path).

Run:  python tests/golden/make_golden.py        (needs /root/reference; ~1-2 minutes)
The fixture is committed; tests read only the .npz.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types
from types import SimpleNamespace
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

_STUBS = ("kornia", "alive_progress", "torchmetrics", "matplotlib", "imageio", "pypcd4", "robotathome", "nvidia_smi",
          "cv2", "mpl_toolkits", "skimage", "open3d", "rosbag", "rospy", "tf", "sensor_msgs", "cv_bridge", "tqdm",
          "scipy", "pandas", "PIL", "yaml", "seaborn")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _STUBS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__ = []
        m.__spec__ = spec
        m.__name__ = spec.name
        return m

    def exec_module(self, module):
        pass


def setup_reference_imports():
    sys.path.insert(0, os.path.join(HERE, "ti_shim"))
    sys.path.insert(0, REF)
    sys.meta_path.insert(0, _StubFinder())
    os.chdir(REF)   # the reference does sys.path.insert(0, os.getcwd())


def main():
    setup_reference_imports()
    sys.path.insert(0, ROOT)
    from virus_nerf_b200 import synthetic    # input generator only (host numpy)
    import taichi as ti
    assert "ti_shim" in ti.__file__
    from modules import utils as rutils
    from modules import hash_encoder as rhash
    from modules import hash_encoder_half as rhalf
    from modules import intersection as rinter
    from modules import ray_march as rmarch
    from modules import volume_train as rvol
    from modules import volume_render_test as rvt
    from modules import spherical_harmonics as rsh
    from modules.occupancy_grid import OccupancyGrid
    from helpers.geometric_fcts import distToCubeBorder
    from training.loss import Loss

    G = {}
    rng = np.random.default_rng(2024)
    t = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(dt)

    # ---- a1/a2: hash encoder (fp32) through the reference HashEncoder module ------------
    for tag, log2_T, max_res in (("T19", 19, 1024.0), ("T14", 14, 512.0)):
        enc = rhash.HashEncoder(max_params=2 ** log2_T, levels=16, base_res=16.0, max_res=max_res, feature_per_level=2)
        G[f"hash_{tag}_offsets"] = enc.offsets.numpy().copy()
        G[f"hash_{tag}_sizes"] = enc.hash_map_sizes.numpy().copy()
        G[f"hash_{tag}_begin_fast"] = np.int64(enc.begin_fast_hash_level)
        G[f"hash_{tag}_total"] = np.int64(enc.total_param_size)
        G[f"hash_{tag}_log_b"] = np.float64(enc.log_b)
        xyz = rng.random((40, 3)).astype(np.float32)
        xyz = np.concatenate([xyz, np.array([[0, 0, 0], [0.5, 0.5, 0.5], [1 - 2 ** -24] * 3, [0.999, 0.001, 0.5]], np.float32)])
        if log2_T == 14:
            table = rng.random(enc.total_param_size, dtype=np.float32)
            G[f"hash_{tag}_table"] = table
        else:   # 45 MB table: regenerate from a seed instead of storing it
            table = np.random.default_rng(7).random(enc.total_param_size, dtype=np.float32)
            G[f"hash_{tag}_table_seed"] = np.int64(7)
        out = torch.empty(xyz.shape[0], 32)
        enc._hash_encoder_kernel(t(xyz), t(table), out, enc.hash_map_sizes, enc.offsets, xyz.shape[0])
        G[f"hash_{tag}_xyz"] = xyz
        G[f"hash_{tag}_out"] = out.numpy().copy()

    # ---- a4: half encoder fwd + explicit backward kernel --------------------------------
    enc = rhalf.HashEncoder(max_params=2 ** 14, levels=16, base_res=16.0, max_res=512.0, feature_per_level=2)
    n_e = enc.hash_table.shape[0]
    table = (rng.random((n_e, 2), dtype=np.float32) * 2 - 1)
    xyz = rng.random((24, 3)).astype(np.float32)
    out = torch.empty(24, 16, 2, dtype=torch.float16)
    enc._hash_encoder_kernel(t(xyz), t(table).to(torch.float16), out, enc.hash_map_sizes, enc.offsets, 24)
    dout = rng.normal(size=(24, 16, 2)).astype(np.float16)
    dout[::5] = 0
    hg = torch.zeros(n_e, 2)
    enc._hash_encoder_backward_kernel(t(xyz), enc.hash_map_sizes, enc.offsets, t(dout, torch.float16), hg, 24)
    G.update(half_table=table, half_xyz=xyz, half_out=out.numpy().copy(), half_dout=dout, half_grad=hg.numpy().copy())

    # ---- rays: scene rays + edge cases ---------------------------------------------------
    sc = synthetic.RoomScene()
    ds = synthetic.SyntheticDataset(sc, pool_size=1 << 12, n_images=8, device="cpu")
    b = ds(20, {"pixs": {"valid_uss": 0.4, "valid_tof": 0.4}})
    so, sd = synthetic.scan_rays(6)
    ex_o = np.array([[0, 0, 0], [2, 2, 2], [2, 0, 0], [0.2, 0.1, 0.0], [-0.5, 0.0, 0.0]], np.float32)
    ex_d = np.array([[0, 0, -1], [1, 0, 0], [-1, 0, 0], [0.6, 0.8, 0.0], [1, 0, 0]], np.float32)
    ro = np.concatenate([b["rays_o"].numpy(), so, ex_o]).astype(np.float32)
    rd = np.concatenate([b["rays_d"].numpy(), sd, ex_d]).astype(np.float32)
    n = ro.shape[0]
    G.update(rays_o=ro, rays_d=rd)

    # ---- a5 ----------------------------------------------------------------------------
    for scale in (0.5, 2.0):
        G[f"aabb_{scale}"] = rinter.ray_aabb_intersection(t(ro), t(rd), scale).numpy().copy()

    # ---- a6: raymarching_train (wrapper draws torch.rand noise: fix it) ------------------
    bf_carved = synthetic.morton_pack(sc.occupancy_bitfield(128))
    bf_rand = np.random.default_rng(3).integers(0, 256, 128 ** 3 // 8).astype(np.uint8)
    G["bf_rand_seed"] = np.int64(3)
    noise = rng.random(n).astype(np.float32)
    G["march_noise"] = noise
    cases = {"carved": (bf_carved, 0.5, 0.0, 1), "rand": (bf_rand, 0.5, 0.0, 1),
             "casc": (np.concatenate([bf_carved, bf_rand, bf_rand[::-1].copy()]), 2.0, 1 / 256, 3)}
    for name, (bf, scale, esf, casc) in cases.items():
        hits = rinter.ray_aabb_intersection(t(ro), t(rd), scale)
        with mock.patch.object(torch, "rand_like", lambda x: t(noise)):
            rays_a, xyzs, dirs, deltas, ts, total = rmarch.raymarching_train(t(ro), t(rd), hits, t(bf, torch.uint8), casc,
                                                                             scale, esf, 128, 1024)
        total = int(total)
        G[f"march_{name}_rays_a"] = rays_a.numpy().copy()
        G[f"march_{name}_xyzs"] = xyzs[:total].numpy().copy()
        G[f"march_{name}_dirs"] = dirs[:total].numpy().copy()
        G[f"march_{name}_deltas"] = deltas[:total].numpy().copy()
        G[f"march_{name}_ts"] = ts[:total].numpy().copy()
        print("march", name, "total", total)

    # ---- a7: raymarching_test, two successive rounds --------------------------------------
    hits = rinter.ray_aabb_intersection(t(ro), t(rd), 0.5).contiguous()
    alive = torch.arange(0, n, 2, dtype=torch.long)
    G["mtest_alive"] = alive.numpy().copy()
    for rnd, ns in enumerate((1, 4)):
        pk, ri, de, ts_ = rmarch.raymarching_test(t(ro), t(rd), hits, alive, t(bf_carved, torch.uint8), 1, 0.5, 0.0, 128, ns)
        G[f"mtest{rnd}_pack"] = pk.numpy().copy(); G[f"mtest{rnd}_ri"] = ri.numpy().copy()
        G[f"mtest{rnd}_deltas"] = de.numpy().copy(); G[f"mtest{rnd}_ts"] = ts_.numpy().copy()
        G[f"mtest{rnd}_hits"] = hits.numpy().copy()

    # ---- a8: volume_rendering_kernel (forward; autodiff is not emulated) -------------------
    ra = torch.from_numpy(G["march_carved_rays_a"])
    S = G["march_carved_ts"].shape[0]
    sig = (rng.random(S).astype(np.float32) ** 2) * 8000
    rgbs = rng.random((S, 3)).astype(np.float32)
    T_rec = torch.zeros(S); tot = torch.zeros(n, dtype=torch.int32); op = torch.zeros(n); dp = torch.zeros(n)
    rgb = torch.zeros(n, 3); ws = torch.zeros(S)
    rvol.volume_rendering_kernel(t(sig), t(rgbs), t(G["march_carved_deltas"]), t(G["march_carved_ts"]), ra, 1e-4, T_rec, tot,
                                 op, dp, rgb, ws)
    G.update(comp_sigmas=sig, comp_rgbs=rgbs, comp_total=tot.numpy().copy(), comp_opacity=op.numpy().copy(),
             comp_depth=dp.numpy().copy(), comp_rgb=rgb.numpy().copy(), comp_ws=ws.numpy().copy())

    # ---- a10: composite_test --------------------------------------------------------------
    pack = torch.stack([ra[:, 1].long(), torch.clamp(ra[:, 2].long(), max=6)], -1)
    alive2 = torch.arange(n, dtype=torch.long)
    op2 = t(rng.random(n).astype(np.float32) * 0.5); dp2 = t(rng.random(n).astype(np.float32)); rgb2 = t(rng.random((n, 3)).astype(np.float32))
    G.update(ct_pack=pack.numpy().copy(), ct_op_in=op2.numpy().copy(), ct_dp_in=dp2.numpy().copy(), ct_rgb_in=rgb2.numpy().copy())
    rvt.composite_test(t(sig), t(rgbs), t(G["march_carved_deltas"]), t(G["march_carved_ts"]), pack, alive2, 1e-2, op2, dp2, rgb2)
    G.update(ct_alive=alive2.numpy().copy(), ct_op=op2.numpy().copy(), ct_dp=dp2.numpy().copy(), ct_rgb=rgb2.numpy().copy())

    # ---- a11 ------------------------------------------------------------------------------
    d = rng.random((32, 3)).astype(np.float32)
    emb = torch.empty(32, 16)
    rsh.dir_encoder(t(d), emb, 32)
    G.update(sh_in=d, sh_out=emb.numpy().copy())

    # ---- a15: morton / packbits -------------------------------------------------------------
    coords = rng.integers(0, 128, (64, 3)).astype(np.int32)
    m = rutils.morton3D(t(coords, torch.int32))
    G.update(morton_coords=coords, morton_idx=m.numpy().copy(), morton_inv=rutils.morton3D_invert(m).numpy().copy())
    grid = rng.random(512).astype(np.float32); grid[:8] = 0.5
    bits = torch.zeros(64, dtype=torch.uint8)
    rutils.packbits(t(grid), 0.5, bits)
    G.update(pack_grid=grid, pack_bits=bits.numpy().copy())

    # ---- a14: the reference OccupancyGrid (pure torch) at G = 32 ----------------------------
    args = synthetic.make_args(device="cpu")
    args.training.debug_mode = False
    torch.manual_seed(11)
    og = OccupancyGrid(args, 32, scene=None, dataset=None, fct_density=None)
    r = torch.arange(32, dtype=torch.int32)
    og.grid_coords = torch.stack(torch.meshgrid(r, r, r, indexing="ij"), -1).reshape(-1, 3)
    G["occ_grid0"] = og.occ_3d_grid.numpy().copy()
    G["occ_decay"] = np.float64(og.grid_decay)
    oro, ord_ = t(ro[:20]), t(rd[:20] * np.float32(1.3))
    dists, pos, idx = og._calcPos(oro, ord_, add_noise=False)
    G.update(occ_rays_o=oro.numpy().copy(), occ_rays_d=ord_.numpy().copy(), occ_dists=dists.numpy().copy(),
             occ_pos=pos.numpy().copy(), occ_idx=idx.numpy().copy())
    nz = rng.random((20, 32, 3)).astype(np.float32)
    with mock.patch.object(torch, "rand", lambda size, device=None, dtype=None: t(nz)):
        _, pos_n, idx_n = og._calcPos(oro, ord_, add_noise=True)
    G.update(occ_noise=nz, occ_pos_noise=pos_n.numpy().copy(), occ_idx_noise=idx_n.numpy().copy())
    meas = (rng.random(20) * 0.7 + 0.05).astype(np.float32)
    po, pe = og._rayProb(t(meas), dists)
    G.update(occ_meas=meas, occ_po=po.numpy().copy(), occ_pe=pe.numpy().copy())
    og._updateGrid(idx, po.reshape(-1), pe.reshape(-1))
    G["occ_grid1"] = og.occ_3d_grid.numpy().copy()
    rho = (rng.random(640) ** 3 * 40 + 1e-3).astype(np.float32)
    og.fct_density = lambda x: t(rho)
    po_n, pe_n = og._nerfProb(pos_n)
    G.update(occ_rho=rho, occ_po_nerf=po_n.numpy().copy(), occ_pe_nerf=pe_n.numpy().copy())
    og._updateGrid(idx_n, po_n, pe_n)
    og.occ_3d_grid *= og.grid_decay
    G["occ_grid2"] = og.occ_3d_grid.numpy().copy()
    og.updateBitfield(grid=og.occ_3d_grid, threshold=og.threshold, convert_cart2morton=True)
    G["occ_bitfield"] = og.getBitfield().numpy().copy()
    # debug round trip of trainer_plot.py:73-86
    cart = og.morton2cartesian(og.bitfield2morton(og.getBitfield()))
    assert torch.equal(cart, og.thresholdGrid(og.occ_3d_grid, og.threshold))
    # helpers/geometric_fcts.py known answers (test_scripts/helpers/test_geometric_fcts.py:8-15)
    kat = distToCubeBorder(torch.zeros(2, 3), torch.tensor([[0, 0, 2.0], [0, 1.5, -1.0]]), cube_min=-0.5, cube_max=0.5)
    G["dist_kat"] = kat.numpy().copy()

    # ---- training/loss.py ---------------------------------------------------------------------
    args.device = torch.device("cpu")
    scene = SimpleNamespace(w2c=lambda pos, only_scale=True, copy=True: pos)
    loss = Loss(args, scene=scene, sensors_dict={})
    res = {"rgb": t(rng.random((n, 3)).astype(np.float32)), "depth": t(rng.random(n).astype(np.float32))}
    uss = rng.random(n).astype(np.float32); uss[::3] = np.nan
    tof = rng.random(n).astype(np.float32); tof[1::2] = np.nan
    data = {"rgb": t(rng.random((n, 3)).astype(np.float32)), "depth": {"USS": t(uss), "ToF": t(tof)}}
    total, ld = loss(res, data, return_loss_dict=True)
    G.update(loss_res_rgb=res["rgb"].numpy().copy(), loss_res_depth=res["depth"].numpy().copy(), loss_rgb=data["rgb"].numpy().copy(),
             loss_uss=uss, loss_tof=tof, loss_total=np.float32(total), loss_color=np.float32(ld["color"]),
             loss_tof_w=np.float32(ld["ToF"]), loss_uss_w=np.float32(ld["USS"]))

    out = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB,", len(G), "arrays")


if __name__ == "__main__":
    main()
