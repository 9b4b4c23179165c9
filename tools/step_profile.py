"""Host-side attribution of one train step (development tool): per-section wall time with
device syncs, then a cProfile of un-synchronised steps."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from virus_nerf_b200 import _lib, synthetic  # noqa: E402
from virus_nerf_b200.engine import TrainEngine  # noqa: E402
from virus_nerf_b200.modules import rendering  # noqa: E402
from virus_nerf_b200.modules.intersection import ray_aabb_intersection  # noqa: E402

DEV = "cuda:0"
args = synthetic.make_args(device=DEV)
ds = synthetic.SyntheticDataset(pool_size=1 << 18, device=DEV)
eng = TrainEngine(args, ds, DEV)
sync = torch.cuda.synchronize


def tick(label, t0, acc):
    sync()
    t1 = time.perf_counter()
    acc[label] = acc.get(label, 0.0) + (t1 - t0) * 1e3
    return t1


for it in range(6):
    eng.step(ds(4096, args.training.sampling_strategy))
sync()
acc = {}
n_it = 8
for it in range(n_it):
    t = time.perf_counter()
    data = ds(4096, args.training.sampling_strategy); t = tick("dataset", t, acc)
    if eng.step_idx % 8 == 0:
        eng.occupancy_update(); t = tick("occ_update", t, acc)
    eng.flat_g.zero_(); t = tick("zero_grad", t, acc)
    with torch.autocast(device_type="cuda", dtype=torch.float16):
        hits = ray_aabb_intersection(data["rays_o"], data["rays_d"], 0.5); t = tick("aabb", t, acc)
        m = eng.model
        rays_a, xyzs, dirs, deltas, ts, total = rendering.raymarching_train(
            data["rays_o"], data["rays_d"], hits, m.occupancy_grid.getBitfield(), 1, 0.5, 0.0, 128, 1024)
        t = tick("march", t, acc)
        sig, rgbs = m(xyzs, dirs); t = tick("model_fwd", t, acc)
        vr, op, dp, rgb, ws = m.render_func(sig, rgbs, deltas, ts, rays_a, 1e-4); t = tick("composite_fwd", t, acc)
        res = {"rgb": rgb + (1 - op)[:, None], "depth": dp, "opacity": op}
        loss, terms = eng.loss_fn(res, data, 1); t = tick("loss", t, acc)
    (loss * eng.scale).sum().backward(); t = tick("backward", t, acc)
    eng.optimizer_step(); t = tick("optimizer", t, acc)
    eng.step_idx += 1
print("samples", int(total))
for k, v in acc.items():
    print(f"{k:14s} {v / n_it:8.3f} ms/step")
print("total", sum(acc.values()) / n_it)

pr = cProfile.Profile()
sync()
t0 = time.perf_counter()
pr.enable()
for it in range(8):
    eng.step(ds(4096, args.training.sampling_strategy))
sync()
pr.disable()
print("unsynced ms/step", (time.perf_counter() - t0) * 1e3 / 8)
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
