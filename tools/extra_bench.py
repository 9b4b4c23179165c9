"""Timings of the non-headline BASELINE.json configs (development / profiles/r1_extra.md):
config 3 (RH2-shaped, T=2^22, 2^18 rays/batch, fp32 and half encoder), config 5 (1920x1080
test-time render; occupancy-update sweep at 128^3 / 256^3)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from virus_nerf_b200 import _lib, synthetic  # noqa: E402
from virus_nerf_b200.engine import TrainEngine  # noqa: E402
from virus_nerf_b200.modules.occupancy_grid import OccupancyGrid  # noqa: E402
from virus_nerf_b200.modules.rendering import render  # noqa: E402

DEV = "cuda:0"
out = []


def rec(**kw):
    out.append(kw)
    print(json.dumps(kw), flush=True)


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


which = sys.argv[1:] or ["config3", "render", "occ", "next"]
scene = synthetic.RoomScene()
carved = torch.from_numpy(synthetic.morton_pack(scene.occupancy_bitfield(128))).to(DEV)

if "config3" in which:
    # RH2-shaped: RGBD + USS + ToF, T = 2^22 (table exceeds L2), 2^18 rays per batch
    N = 1 << 18
    args = synthetic.make_args(device=DEV, sensors=("RGBD", "USS", "ToF"), batch_size=N)
    ds = synthetic.SyntheticDataset(scene, kind="rh2", pool_size=1 << 19, device=DEV)
    for state in ("carved", "initial"):
        eng = TrainEngine(args, ds, DEV, log2_T=22)
        eng.grid_update_interval = 10 ** 9            # keep the occupancy state fixed
        eng._prep_step = eng.step_idx = 1
        if state == "carved":
            eng.model.occupancy_grid.bitfield = carved
        else:
            eng.model.occupancy_grid._decayAndPack(apply_decay=False)
        batches = [ds(N, {"pixs": "random"}) for _ in range(4)]
        it = [0]

        def one():
            b = batches[it[0] % 4]; it[0] += 1
            eng.step_fast(b, next_data=batches[it[0] % 4])
        ms = ev_time(one, reps=4, warm=2)
        rec(config="3: RH2-shaped, T=2^22 fp32 encoder, 2^18 rays", state=state, ms_per_step=round(ms, 3),
            rays_per_s=round(N / ms * 1e3), samples_per_step=eng.last_samples,
            mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 1))
        del eng
        torch.cuda.empty_cache()
    # half-precision encoder through the drop-in modules (autograd path)
    eng = TrainEngine(args, ds, DEV, log2_T=22, half_opt=True)
    eng.grid_update_interval = 10 ** 9; eng.step_idx = 1
    eng.model.occupancy_grid.bitfield = carved
    eng.model.pos_encoder._direct_grad = None
    b = ds(N, {"pixs": "random"})
    ms = ev_time(lambda: eng.step(b), reps=3, warm=1)
    rec(config="3: RH2-shaped, T=2^22 HALF encoder (modules + autograd), 2^18 rays", state="carved",
        ms_per_step=round(ms, 3), rays_per_s=round(N / ms * 1e3), samples_per_step=int(eng.last_samples))
    del eng
    torch.cuda.empty_cache()

if "render" in which:
    # config 5: 1920x1080 pinhole frame, fx = fy = 960, carved bitfield, reference chunking (8192) and one chunk
    args = synthetic.make_args(device=DEV)
    ds = synthetic.SyntheticDataset(scene, pool_size=1 << 12, n_images=8, device=DEV)
    eng = TrainEngine(args, ds, DEV)
    eng.model.occupancy_grid.bitfield = carved
    with torch.no_grad():
        eng.model.xyz_encoder.output_layer.weight[0].add_(0.25)      # opaque enough to terminate rays
    W, H = 1920, 1080
    u, v = torch.meshgrid(torch.arange(W, device=DEV, dtype=torch.float32), torch.arange(H, device=DEV, dtype=torch.float32), indexing="xy")
    d = torch.stack([(u + 0.5 - W / 2) / 960, torch.ones_like(u), -(v + 0.5 - H / 2) / 960], -1).reshape(-1, 3)
    d = d / d.norm(dim=1, keepdim=True)
    o = torch.tensor([0.0, -0.2, -0.05], device=DEV).expand_as(d).contiguous()
    for chunk, fused in ((8192, False), (8192, True), (1 << 18, True), (W * H, False), (W * H, True)):
        eng.model.fused_test_render = fused

        def frame():
            tot = 0
            for s in range(0, W * H, chunk):
                r = render(eng.model, o[s:s + chunk], d[s:s + chunk], test_time=True, exp_step_factor=0.0)
                tot += int(r["total_samples"])
            return tot
        t0 = time.perf_counter(); tot = frame(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        t0 = time.perf_counter(); tot = frame(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        rec(config="5: 1920x1080 test render", path="loop-free (march all + model + composite)" if fused else
            "reference round loop (raymarching_test + composite_test)", chunk=chunk,
            ms_per_frame=round(dt * 1e3, 1), rays_per_s=round(W * H / dt), samples=tot)

if "occ" in which:
    for G in (128, 256):
        for B in (1024, 8192, 65536):
            args = synthetic.make_args(device=DEV, occ_batch_size=B)
            ds = synthetic.SyntheticDataset(scene, pool_size=1 << 17, n_images=16, device=DEV)
            eng = TrainEngine(args, ds, DEV)
            og = OccupancyGrid(args, G, scene=None, dataset=ds.clone_with_seed(1), fct_density=eng.model.density)
            with torch.autocast(device_type="cuda", dtype=torch.float16):
                ms = ev_time(lambda: og.update(elapse_time=0.0), reps=5, warm=2)
            rec(config="5: occupancy-grid update sweep", grid=G, rays=B, ms_per_update=round(ms, 3),
                cells_per_s=round(B * 32 / ms * 1e3))
            del eng, og
            torch.cuda.empty_cache()

if "next" in which:
    # ---- batch assembly from ETHZ-shaped image storage (200 images, 640x480, 2 cameras) ----
    from types import SimpleNamespace
    from virus_nerf_b200.datasets.dataset_base import DatasetBase
    from virus_nerf_b200.training.sampler import Sampler
    n_img, W, H = 200, 640, 480
    HW = W * H
    g = torch.Generator(device=DEV).manual_seed(0)
    rgbs = torch.rand(n_img, HW, 3, device=DEV, generator=g)
    depths = {"USS": torch.rand(n_img, HW, device=DEV, generator=g), "ToF": torch.rand(n_img, HW, device=DEV, generator=g)}
    q = torch.randn(n_img, 3, 3, device=DEV, generator=g)
    R, _ = torch.linalg.qr(q)
    poses = torch.cat([R, torch.rand(n_img, 3, 1, device=DEV, generator=g) - 0.5], 2).contiguous()
    u, v = torch.meshgrid(torch.arange(W, device=DEV, dtype=torch.float32), torch.arange(H, device=DEV, dtype=torch.float32), indexing="xy")
    dirs = torch.stack([(u + 0.5 - W / 2) / 320, (v + 0.5 - H / 2) / 377, torch.ones_like(u)], -1).reshape(-1, 3)
    dirs = (dirs / dirs.norm(dim=1, keepdim=True)).contiguous()
    sensor_ids = (torch.arange(n_img, device=DEV) % 2) * 2 + 1            # CAM1 -> 1, CAM3 -> 3
    dargs = SimpleNamespace(device=torch.device(DEV), seed=21, logger=SimpleNamespace(error=print),
                            training=SimpleNamespace(debug_mode=False, real_time_simulation=False))
    masks = {"USS": SimpleNamespace(mask=torch.rand(HW, device=DEV, generator=g) < 0.2),
             "ToF": SimpleNamespace(mask=torch.rand(HW, device=DEV, generator=g) < 0.0002)}
    dsb = DatasetBase(dargs, rgbs=rgbs, poses=poses, directions_dict={1: dirs, 3: dirs.clone()}, sensor_ids=sensor_ids,
                      depths_dict=depths, times=torch.arange(n_img, device=DEV, dtype=torch.float32), img_wh=(W, H),
                      sampler=Sampler(dargs, n_img, (W, H), masks, None))
    strat = {"imgs": "all", "pixs": {"valid_uss": 0.4, "valid_tof": 0.4}}

    def torch_path(img_idxs, pix_idxs):
        """the reference's DatasetBase.__call__ / _calcRayPoses as torch ops (dataset_base.py:52-75, 194-243)"""
        img_idxs, pix_idxs = img_idxs.long(), pix_idxs.long()
        N = img_idxs.shape[0]
        ro = torch.full((N, 3), float("nan"), device=DEV); rd = torch.full((N, 3), float("nan"), device=DEV)
        for cam_id, directions in dsb.directions_dict.items():
            m = sensor_ids[img_idxs] == cam_id
            it, pt = img_idxs[m], pix_idxs[m]
            c2w = poses[it]
            d = (directions[pt].unsqueeze(1) @ c2w[..., :3].transpose(1, 2)).squeeze(1)
            ro[m] = c2w[..., 3].expand_as(d); rd[m] = d
        out = {"rays_o": ro.detach().clone(), "rays_d": rd.detach().clone(), "rgb": rgbs[img_idxs, pix_idxs, :3].detach().clone(),
               "depth": {k: v[img_idxs, pix_idxs].detach().clone() for k, v in depths.items()}}
        return out

    for B in (4096, 1 << 18):
        ii, pp = dsb.sampler(B, strat, 0.0)
        a, b = dsb(img_idxs=ii, pix_idxs=pp), torch_path(ii, pp)
        assert torch.allclose(a["rays_d"], b["rays_d"], rtol=1e-5, atol=1e-6) and torch.equal(a["rgb"], b["rgb"])
        ms_k = ev_time(lambda: dsb(img_idxs=ii, pix_idxs=pp), reps=20, warm=3)
        ms_t = ev_time(lambda: torch_path(ii, pp), reps=20, warm=3)
        ms_s = ev_time(lambda: dsb.sampler(B, strat, 0.0), reps=20, warm=3)
        t0 = time.perf_counter()
        for _ in range(20):
            dsb(batch_size=B, sampling_strategy=strat, elapse_time=0.0)
        torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 20 * 1e3
        rec(config="8(f) row 2: batch assembly, 200 x 640x480 images", rays=B, kernel_ms=round(ms_k, 4),
            torch_reference_path_ms=round(ms_t, 4), sampler_ms=round(ms_s, 4), sample_and_assemble_wall_ms=round(wall, 4),
            gbs=round(B * (8 + 12 + 48 + 12 + 8 + 36 + 8) / ms_k / 1e6, 1))
    del rgbs, depths, dsb
    torch.cuda.empty_cache()
    # ---- NGPGrid.update at 128^3 with the model's density (fused tcgen05 MLP) ----
    args = synthetic.make_args(device=DEV, grid_type="ngp")
    ds = synthetic.SyntheticDataset(scene, pool_size=1 << 12, n_images=8, device=DEV)
    eng = TrainEngine(args, ds, DEV)
    grid = eng.model.occupancy_grid
    thr = 0.01 * 1024 / 3 ** 0.5
    with torch.autocast(device_type="cuda", dtype=torch.float16):
        ms_w = ev_time(lambda: grid.update(thr, warmup=True), reps=5, warm=2)
        ms_s = ev_time(lambda: grid.update(thr, warmup=False), reps=5, warm=2)
    rec(config="8(f) row 3: NGPGrid.update, G=128", warmup_ms=round(ms_w, 3), warmup_cells_per_s=round(128 ** 3 / ms_w * 1e3),
        sampled_ms=round(ms_s, 3), sampled_cells_per_s=round(128 ** 3 / 2 / ms_s * 1e3), host_syncs_per_update=0)

with open(os.path.join(ROOT, "gpurun_out", "extra_bench.jsonl"), "a") as f:
    for l in out:
        f.write(json.dumps(l) + "\n")
