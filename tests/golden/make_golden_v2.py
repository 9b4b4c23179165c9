"""Generates tests/golden/golden_v2.npz -- goldens for the "next" rows (SURVEY section 8(f) rows 2
and 3) by running THE REFERENCE'S OWN CODE from /root/reference (build container only):

  * datasets/dataset_base.py: DatasetBase._calcRayPoses + the gathers of __call__ (pure torch);
  * training/sampler.py: Sampler (pure torch / CPU generator, seeded);
  * modules/ngp_grid.py: NGPGrid.update (warm-up and sampled), with its Taichi helpers
    (morton3D, morton3D_invert, packbits) executed under tests/golden/ti_shim.  The random
    draws the reference makes inside (torch.randint, torch.rand_like) are recorded so that the
    restatement can be fed the same numbers.

Run:  python tests/golden/make_golden_v2.py     (needs /root/reference; ~1 minute)
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_golden import setup_reference_imports  # noqa: E402


def main():
    setup_reference_imports()
    import taichi as ti
    assert "ti_shim" in ti.__file__
    from datasets.dataset_base import DatasetBase as RefDataset
    from training.sampler import Sampler as RefSampler
    from modules.ngp_grid import NGPGrid as RefNGP
    import helpers.data_fcts as data_fcts

    G = {}
    rng = np.random.default_rng(77)
    logger = SimpleNamespace(error=print, warning=print, info=print)

    # ---------------- f2: batch assembly --------------------------------------------------
    W, H, n_img = 20, 12, 10
    HW = W * H
    cams = ["CAM1", "CAM3"]
    args = SimpleNamespace(device=torch.device("cpu"), seed=21, dataset=SimpleNamespace(name="ETHZ"), logger=logger,
                           training=SimpleNamespace(debug_mode=False, real_time_simulation=False))
    ids = {c: data_fcts.sensorName2ID(sensor_name=c, dataset="ETHZ") for c in cams}
    G["f2_cam_ids"] = np.array([ids[c] for c in cams], np.int64)
    sensor_ids = torch.tensor([ids[cams[k % 2]] for k in range(n_img)])
    # random rigid poses
    q = rng.normal(size=(n_img, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    a, b, c, d = q.T
    R = np.stack([np.stack([a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)], -1),
                  np.stack([2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)], -1),
                  np.stack([2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d], -1)], 1)
    poses = np.concatenate([R, rng.uniform(-0.4, 0.4, (n_img, 3, 1))], 2).astype(np.float32)
    directions = {}
    for k, cam in enumerate(cams):
        u, v = np.meshgrid(np.arange(W), np.arange(H), indexing="xy")
        dirs = np.stack([(u + 0.5 - W / 2) / (10.0 + k), (v + 0.5 - H / 2) / (11.0 + k), np.ones_like(u, float)], -1)
        dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
        directions[cam] = dirs.reshape(-1, 3).astype(np.float32)
    rgbs = rng.random((n_img, HW, 3)).astype(np.float32)
    depths = {"USS": rng.random((n_img, HW)).astype(np.float32), "ToF": rng.random((n_img, HW)).astype(np.float32)}
    depths["USS"][rng.random((n_img, HW)) < 0.5] = np.nan
    depths["ToF"][rng.random((n_img, HW)) < 0.8] = np.nan
    times = np.sort(rng.random(n_img)).astype(np.float32)

    ds = object.__new__(RefDataset)
    ds.args, ds.split = args, "train"
    ds.rgbs, ds.poses = torch.from_numpy(rgbs), torch.from_numpy(poses)
    ds.directions_dict = {k: torch.from_numpy(v) for k, v in directions.items()}
    ds.sensor_ids, ds.times = sensor_ids, torch.from_numpy(times)
    ds.depths_dict = {k: torch.from_numpy(v) for k, v in depths.items()}
    B = 257
    img_idxs = torch.from_numpy(rng.integers(0, n_img, B).astype(np.int32))
    pix_idxs = torch.from_numpy(rng.integers(0, HW, B).astype(np.int32))
    out = ds(img_idxs=img_idxs.long(), pix_idxs=pix_idxs.long())
    G.update({"f2_poses": poses, "f2_dirs": np.stack([directions[c] for c in cams]), "f2_rgbs": rgbs,
              "f2_uss": depths["USS"], "f2_tof": depths["ToF"], "f2_times": times, "f2_sensor_ids": sensor_ids.numpy(),
              "f2_img_idxs": img_idxs.numpy(), "f2_pix_idxs": pix_idxs.numpy(),
              "f2_rays_o": out["rays_o"].detach().numpy(), "f2_rays_d": out["rays_d"].detach().numpy(),
              "f2_rgb": out["rgb"].detach().numpy(), "f2_out_uss": out["depth"]["USS"].detach().numpy(),
              "f2_out_tof": out["depth"]["ToF"].detach().numpy(), "f2_out_ids": out["sensor_ids"].numpy(),
              "f2_out_time": out["time"].numpy()})

    # ---------------- f2: sampler -----------------------------------------------------------
    mask_uss = torch.from_numpy(rng.random(HW) < 0.3)
    mask_tof = torch.from_numpy(rng.random(HW) < 0.1)
    sensors = {"USS": SimpleNamespace(mask=mask_uss), "ToF": SimpleNamespace(mask=mask_tof)}
    G["f2_mask_uss"], G["f2_mask_tof"] = mask_uss.numpy(), mask_tof.numpy()
    sm = RefSampler(args=args, dataset_len=n_img, img_wh=(W, H), sensors_dict=sensors, times=torch.from_numpy(times))
    strategies = [{"imgs": "all", "pixs": {"valid_uss": 0.4, "valid_tof": 0.4}}, {"imgs": "same", "pixs": "random"},
                  {"imgs": "all", "pixs": "valid_tof"}, {"imgs": "all", "pixs": "entire_img"}]
    torch.manual_seed(1234)
    for k, st in enumerate(strategies):
        ii, pp = sm(batch_size=64, sampling_strategy=st, elapse_time=0.0)
        G[f"f2_sampler{k}_img"], G[f"f2_sampler{k}_pix"] = ii.numpy(), pp.numpy()

    # ---------------- f3: NGPGrid.update ------------------------------------------------------
    Gs = 16
    G3 = Gs ** 3
    nargs = SimpleNamespace(device=torch.device("cpu"), model=SimpleNamespace(scale=0.5), logger=logger)

    def density(x):          # analytic stand-in for NGP.density: positive, smooth, some large values
        return torch.exp(3.0 * torch.sin(7.0 * x[:, 0]) * torch.cos(5.0 * x[:, 1]) + 2.0 * x[:, 2])

    grid = object.__new__(RefNGP)
    grid.args, grid.grid_size, grid.scale, grid.fct_density = nargs, Gs, 0.5, density
    grid.cascades, grid.morton_structure, grid.threshold = 1, True, 0.5
    r = torch.arange(Gs, dtype=torch.int32)
    x, y, z = torch.meshgrid(r, r, r, indexing="ij")
    grid.grid_coords = torch.stack([x, y, z], -1).reshape(-1, 3)
    grid.bitfield = torch.zeros(G3 // 8, dtype=torch.uint8)
    occ0 = torch.from_numpy(rng.random(G3).astype(np.float32)) * 2.0
    occ0[torch.from_numpy(rng.random(G3) < 0.1)] = -1.0                 # invisible cells keep their mark
    occ0[torch.from_numpy(rng.random(G3) < 0.2)] = 0.0
    grid.occ_morton_grid = occ0.clone().reshape(1, G3)
    G["f3_occ0"] = occ0.numpy()

    rec = {}
    orig_randint, orig_rand_like = torch.randint, torch.rand_like

    def randint(*a, **k):
        o = orig_randint(*a, **k); rec.setdefault("randint", []).append(o.clone()); return o

    def rand_like(*a, **k):
        o = orig_rand_like(*a, **k); rec.setdefault("rand_like", []).append(o.clone()); return o

    torch.randint, torch.rand_like = randint, rand_like
    try:
        torch.manual_seed(99)
        thr = 0.01 * 1024 / 3 ** 0.5                                    # trainer.py:110
        G["f3_density_threshold"] = np.float64(thr)
        for tag, warm in (("warm", True), ("samp", False), ("samp2", False)):
            rec.clear()
            before = grid.occ_morton_grid.clone()
            grid.update(density_threshold=thr, warmup=warm, decay=0.95, erode=False)
            G[f"f3_{tag}_before"] = before.numpy().reshape(-1)
            G[f"f3_{tag}_after"] = grid.occ_morton_grid.numpy().reshape(-1).copy()
            G[f"f3_{tag}_threshold"] = np.float64(grid.threshold)
            G[f"f3_{tag}_bitfield"] = grid.bitfield.numpy().copy()
            G[f"f3_{tag}_noise"] = rec["rand_like"][0].numpy()
            if not warm:
                G[f"f3_{tag}_coords1"] = rec["randint"][0].numpy()
                G[f"f3_{tag}_rand_idx"] = rec["randint"][1].numpy()
    finally:
        torch.randint, torch.rand_like = orig_randint, orig_rand_like
    # the cells the reference visited and the positions / densities it evaluated, re-derived with its own helpers
    from modules.utils import morton3D, morton3D_invert
    G["f3_all_indices"] = morton3D(grid.grid_coords).long().numpy()
    G["f3_all_coords"] = grid.grid_coords.numpy()

    # ---------------- f4: evaluation inputs (helpers/geometric_fcts.py) ----------------------
    from helpers.geometric_fcts import createScanRays, createScanPos
    origins = rng.uniform(-0.3, 0.3, (3, 3)).astype(np.float32)
    so, sd = createScanRays(rays_o=torch.from_numpy(origins), angle_res=48)
    G["f4_origins"], G["f4_scan_o"], G["f4_scan_d"] = origins, so.numpy(), sd.numpy()
    G["f4_scan_pos"] = createScanPos(res_map=9, height_c=-0.05, num_avg_heights=3, tolerance_c=0.02, cube_min=-0.5,
                                     cube_max=0.5, device="cpu").numpy()

    out_path = os.path.join(HERE, "golden_v2.npz")
    np.savez_compressed(out_path, **G)
    print("wrote", out_path, len(G), "arrays")


if __name__ == "__main__":
    main()
