"""Runs the hot kernels a few times at the benchmark's sample count (ray-coherent samples from
the synthetic scene, all-occupied grid) -- the target for `ncu --set full -k regex:...`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from virus_nerf_b200 import _lib, synthetic
from virus_nerf_b200.modules.intersection import ray_aabb_intersection
from virus_nerf_b200.modules.ray_march import raymarching_train
DEV = "cuda:0"
# default: the fast step's flags (planes, 2 levels per thread, 16-byte pair loads, 48-register bwd, zero skip)
# round 2: + 8192 = fp16 operand chunks out of the hash forward, bulk-copied by the fused MLP (enc_format 3)
flags = int(sys.argv[1]) if len(sys.argv) > 1 else (512 | 256 | 2048 | 1024 | 4096 | 8192)
fmt = 3 if flags & 8192 else (2 if flags & 512 else 0)
ds = synthetic.SyntheticDataset(pool_size=1 << 16, device=DEV)
b = ds(4096, {"pixs": {"valid_uss": 0.4, "valid_tof": 0.4}})
bf = torch.full((128 ** 3 // 8,), 255, dtype=torch.uint8, device=DEV)
hits = ray_aabb_intersection(b["rays_o"], b["rays_d"], 0.5)
rays_a, xyzs, dirs, deltas, ts, total = raymarching_train(b["rays_o"], b["rays_d"], hits, bf, 1, 0.5, 0.0, 128, 1024)
S = int(total)
x = (xyzs + 0.5).clamp(0, 1).contiguous()
lv = _lib.hash_levels(16, 1024, 16, 2 ** 19)
table = torch.rand(2 * lv.total_entries, device=DEV)
o = torch.empty(S, 32, device=DEV); dout = torch.randn(S, 32, device=DEV); grad = torch.zeros_like(table)
g = torch.Generator().manual_seed(0)
xav = lambda a, c: ((torch.rand(a, c, generator=g) * 2 - 1) * (6.0 / (a + c)) ** 0.5).to(DEV)
W = [xav(64, 32), xav(16, 64), xav(64, 32), xav(64, 64), xav(3, 64)]
sig = torch.empty(S, device=DEV); rgb = torch.empty(S, 3, device=DEV)
dsig = torch.randn(S, device=DEV); drgb = torch.randn(S, 3, device=DEV); denc = torch.empty(S, 32, device=DEV)
dW = [torch.zeros_like(w) for w in W]
for _ in range(3):
    _lib.call("vn_hash_encode_fwd_f32", x, table, o, S, lv, flags)
    _lib.call("vn_hash_encode_bwd_f32", x, dout, grad, S, lv, flags)
    _lib.call("vn_mlp_fwd", o, fmt, dirs, *W, S, 0, sig, rgb, None)
    _lib.call("vn_mlp_bwd", o, fmt, dirs, *W, S, 0, dsig, drgb, denc, *dW)
torch.cuda.synchronize()
print("S", S)
