"""Development tool: per-round cycle breakdown of one tile chain of the pipelined MLP backward (library built with
`make -C virus-nerf_b200/csrc EXTRA=-DVN_MLP_TIMING`): clock64 deltas of chain 0 / CTA 0, averaged per tile."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from virus_nerf_b200 import _lib
DEV = "cuda:0"
S = int(sys.argv[1]) if len(sys.argv) > 1 else 933574
g = torch.Generator().manual_seed(0)
xav = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) * (6.0 / (i + o)) ** 0.5).to(DEV)
W = [xav(64, 32), xav(16, 64), xav(64, 32), xav(64, 64), xav(3, 64)]
enc = torch.rand(S, 32, device=DEV); dirs = torch.randn(S, 3, device=DEV)
enc_c = enc.half().view(S, 4, 8).permute(1, 0, 2).contiguous()
FMT = int(os.environ.get("FMT", "5"))
if FMT == 5:
    enc_c = torch.cat([enc_c, torch.rand(2, S, 8, device=DEV).half()], 0).contiguous()
dsig = torch.randn(S, device=DEV); drgb = torch.randn(S, 3, device=DEV)
denc = torch.empty(S, 32, device=DEV); dW = [torch.zeros_like(w) for w in W]
L = _lib.lib()
out = (ctypes.c_ulonglong * 64)()
for _ in range(3):
    _lib.call("vn_mlp_bwd", enc_c, FMT, dirs, *W, S, 0, dsig, drgb, denc, *dW)
torch.cuda.synchronize()
L.vn_mlp_debug_timing(None, 1)
reps = 5
for _ in range(reps):
    _lib.call("vn_mlp_bwd", enc_c, FMT, dirs, *W, S, 0, dsig, drgb, denc, *dW)
torch.cuda.synchronize()
L.vn_mlp_debug_timing(out, 0)
n_tiles = (S + 127) // 128
tiles = (n_tiles + 3 * 148 - 1) // (3 * 148) * reps      # tiles of chain 0 / CTA 0
names = ["", "", "stage(+in wait)"] + sum([[f"wait r{r}", f"epi r{r}"] for r in range(1, 11)], [])
tot = 0
for i in range(2, 23):
    v = out[i] / tiles
    tot += v
    print(f"{names[i]:18s} {v:8.0f} cycles/tile")
print("total per tile-chain", round(tot))
for i, n in zip(range(32, 39), ["r6: ldtm+lds wait", "r6: mask+pack+tmem_st issue", "r6: wait::st", "r6: post_main", "r6: wait_wgrad", "r6: sts", "r6: post_wg"]):
    print(f"{n:30s} {out[i] / tiles:8.0f}")
