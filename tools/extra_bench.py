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


which = sys.argv[1:] or ["config3", "render", "occ"]
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

with open(os.path.join(ROOT, "gpurun_out", "extra_bench.jsonl"), "a") as f:
    for l in out:
        f.write(json.dumps(l) + "\n")
