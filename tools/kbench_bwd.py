"""Micro-benchmark of the backward kernels of the fast step at the benchmark's sample count (ray-coherent samples of 4096
rays through the all-occupied grid): hash backward f32 (planes) / f16 (chunks) at T = 2^19 and 2^22, the fused MLP
backward, and the fused MLP-backward + scatter kernel.  CUDA events, L2 flushed between repetitions.  Development tool
(A/B of builds through VN_LIB_PATH); prints one JSON line per kernel."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from virus_nerf_b200 import _lib, synthetic  # noqa: E402
from virus_nerf_b200.modules.intersection import ray_aabb_intersection  # noqa: E402
from virus_nerf_b200.modules.ray_march import raymarching_train  # noqa: E402
from tools.kbench import timeit  # noqa: E402

DEV = "cuda:0"
FAST = _lib.VN_HASH_PLANAR | _lib.VN_HASH_LEVEL_GROUPS_2 | _lib.VN_HASH_TIGHT_REGS | _lib.VN_HASH_SKIP_ZERO_GRADS


def main():
    tag = os.environ.get("VN_LIB_PATH", "default").split("/")[-1]
    scene = synthetic.RoomScene()
    ds = synthetic.SyntheticDataset(scene, pool_size=1 << 18, device=DEV)
    b = ds(4096, {"pixs": {"valid_uss": 0.4, "valid_tof": 0.4}})
    bf = torch.full((128 ** 3 // 8,), 255, dtype=torch.uint8, device=DEV)
    hits = ray_aabb_intersection(b["rays_o"], b["rays_d"], 0.5)
    noise = torch.rand(4096, device=DEV)
    rays_a, xyzs, dirs, deltas, ts, total = raymarching_train(b["rays_o"], b["rays_d"], hits, bf, 1, 0.5, 0.0, 128, 1024, noise=noise)
    S = int(total)
    x = (xyzs + 0.5).clamp(0, 1).contiguous()
    g = torch.Generator().manual_seed(0)
    xav = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) * (6.0 / (i + o)) ** 0.5).to(DEV)
    W = [xav(64, 32), xav(16, 64), xav(64, 32), xav(64, 64), xav(3, 64)]
    dsig = torch.randn(S, device=DEV); drgb = torch.randn(S, 3, device=DEV)
    enc_c = torch.rand(4, S, 8, device=DEV).half().contiguous()
    dW = [torch.zeros_like(w) for w in W]

    def rec(name, ms, **kw):
        print(json.dumps({"lib": tag, "kernel": name, "ms": round(ms, 4), "S": S, **kw}), flush=True)

    for log2_T in (19, 22):
        lv = _lib.hash_levels(16, 1024, 16, 2 ** log2_T)
        table = torch.rand(2 * lv.total_entries, device=DEV)
        enc_o = torch.empty(S * 32, device=DEV)
        ffl = _lib.VN_HASH_PLANAR | _lib.VN_HASH_LEVEL_GROUPS_2 | _lib.VN_HASH_PAIR_LOADS | _lib.VN_HASH_F16_CHUNKS
        ms = timeit(lambda: _lib.call("vn_hash_encode_fwd_f32", x, table, enc_o, S, lv, ffl), reps=7)
        rec("hash_fwd_f32_chunks", ms, log2_T=log2_T, frac=round(S * 1164 / ms / 1e6 / 6454.9, 4))
        for extra, lpt_tag in ((_lib.VN_HASH_LEVEL_GROUPS_8, "lpt8"), (_lib.VN_HASH_LEVEL_GROUPS_16, "lpt16")):
            ms = timeit(lambda: _lib.call("vn_hash_encode_fwd_f32", x, table, enc_o, S, lv, (ffl & ~_lib.VN_HASH_LEVEL_GROUPS_2) | extra), reps=7)
            rec("hash_fwd_f32_chunks_" + lpt_tag, ms, log2_T=log2_T, frac=round(S * 1164 / ms / 1e6 / 6454.9, 4))
        del table
        grad = torch.zeros(2 * lv.total_entries, device=DEV)
        dout = torch.randn(8, S, 4, device=DEV)
        ms = timeit(lambda: _lib.call("vn_hash_encode_bwd_f32", x, dout, grad, S, lv, FAST), reps=7)
        rec("hash_bwd_f32_planes", ms, log2_T=log2_T, frac=round(S * 1164 / ms / 1e6 / 6454.9, 4))
        douth = torch.randn(4, S, 8, device=DEV).half().contiguous()
        ms = timeit(lambda: _lib.call("vn_hash_encode_bwd_f16", x, douth, grad, S, lv, FAST | _lib.VN_HASH_F16_CHUNKS), reps=7)
        rec("hash_bwd_f16_chunks", ms, log2_T=log2_T, frac=round(S * 1100 / ms / 1e6 / 6454.9, 4))
        for half in (0, 1):
            ms = timeit(lambda: _lib.call("vn_mlp_bwd_scatter", enc_c, 3, dirs, *W, S, dsig, drgb, x, lv, half, grad, *dW, None), reps=7)
            rec("mlp_bwd_hash_scatter" + ("_f16" if half else ""), ms, log2_T=log2_T, frac=round(S * 1148 / ms / 1e6 / 6454.9, 4))
        del grad
    denc = torch.empty(S * 32, device=DEV)
    ms = timeit(lambda: _lib.call("vn_mlp_bwd", enc_c, 3, dirs, *W, S, 0, dsig, drgb, denc, *dW), reps=7)
    rec("mlp_bwd_chunks", ms, tflops=round(S * 56448 / ms / 1e9, 1))


if __name__ == "__main__":
    main()
