"""Diagnostic: fraction of samples whose output gradients are exactly zero (samples behind an
opaque surface: composite stops at T <= 1e-4) over the benchmark's first steps."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from virus_nerf_b200 import synthetic
from virus_nerf_b200.engine import TrainEngine
DEV = "cuda:0"
args = synthetic.make_args(device=DEV, batch_size=4096)
ds = synthetic.SyntheticDataset(synthetic.RoomScene(), pool_size=1 << 18, device=DEV)
eng = TrainEngine(args, ds, DEV)
for it in range(200):
    eng.step_fast(ds(4096, args.training.sampling_strategy))
    if it % 20 == 7 or it == 199:
        S = eng.last_samples
        ws = eng._ws
        dsig = ws.get("d_sig", S); drgb = ws.get("d_rgbs", S, 3)
        z = ((dsig == 0) & (drgb == 0).all(1))
        tiles = z[: S // 128 * 128].view(-1, 128).all(1).float().mean()
        print(f"step {it}: S={S} zero-gradient samples {float(z.float().mean()):.3f}, all-zero 128-tiles {float(tiles):.3f}", flush=True)
