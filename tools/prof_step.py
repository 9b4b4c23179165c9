"""Profiling target: a few train steps of the benchmark's configuration (4096 rays, T = 2^19, fast step) with the CUDA
profiler range around the last one (an occupancy-update step), so that `ncu --profile-from-start off` captures every kernel of a step -- march
count / expand, hash forward, fused MLP forward, composite forward / backward, loss, the fused MLP-backward + scatter
kernel (or, with --unfused, the MLP backward and the hash backward), Adam, scaler update -- plus, on the first captured
step, the occupancy update (vn_occ_update).  Usage: python tools/prof_step.py [--unfused] [--half]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from virus_nerf_b200 import synthetic  # noqa: E402
from virus_nerf_b200.engine import TrainEngine  # noqa: E402

DEV = "cuda:0"
unfused = "--unfused" in sys.argv
half = "--half" in sys.argv
args = synthetic.make_args(device=DEV, batch_size=4096)
ds = synthetic.SyntheticDataset(synthetic.RoomScene(), pool_size=1 << 18, device=DEV, seed=21)
eng = TrainEngine(args, ds, DEV, fused_scatter=False if unfused else "auto", half_opt=half, log2_T=22 if half else 19)
warm = 16                                  # the captured step is step 16: it carries an occupancy update
for it in range(warm):
    eng.step_fast(ds(4096, args.training.sampling_strategy))
torch.cuda.synchronize()
torch.cuda.profiler.start()
for it in range(1):
    eng.step_fast(ds(4096, args.training.sampling_strategy))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("samples", eng.last_samples, "fused_scatter", eng.fused_scatter, "half", half)
