"""tcgen05 descriptor self-test against torch matmul (development tool)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from virus_nerf_b200 import _lib
dev = "cuda:0"
torch.manual_seed(0)
ok = True
for mode in (0, 1, 2):
    for N, K in ((64, 32), (64, 64), (16, 64), (32, 128), (64, 128), (16, 16)):
        if mode == 0:
            A = torch.randn(128, K, device=dev).half(); B = torch.randn(N, K, device=dev).half()
            ref = A.float() @ B.float().t()
        elif mode == 1:
            A = torch.randn(128, K, device=dev).half(); B = torch.randn(K, N, device=dev).half()
            ref = A.float() @ B.float()
        else:
            A = torch.randn(K, 128, device=dev).half(); B = torch.randn(K, N, device=dev).half()
            ref = A.float().t() @ B.float()
        D = torch.zeros(128, N, device=dev)
        _lib.call("vn_umma_selftest", mode, N, K, A, B, D)
        torch.cuda.synchronize()
        err = (D - ref).abs().max().item()
        good = err < 1e-2 * max(1.0, ref.abs().max().item())
        ok &= good
        print(f"mode {mode} N={N} K={K}: max err {err:.4g} ref max {ref.abs().max().item():.3g} {'OK' if good else 'FAIL'}")
print("ALL OK" if ok else "SOME FAILED")
