"""Per-kernel micro-benchmark at the benchmark's sizes (CUDA events, L2 flushed between
repetitions).  Development tool: prints one line per kernel variant; not a bench value."""
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

DEV = "cuda:0"
flush_buf = None


def timeit(fn, reps=5, flush=True):
    global flush_buf
    if flush_buf is None:
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        if flush:
            flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    which = sys.argv[1:] or ["hash", "march", "composite", "adam"]
    scene = synthetic.RoomScene()
    ds = synthetic.SyntheticDataset(scene, pool_size=1 << 18, device=DEV)
    out = []

    def rec(name, ms, **kw):
        line = {"kernel": name, "ms": round(ms, 4), **kw}
        out.append(line)
        print(json.dumps(line), flush=True)

    for state in ("full", "carved"):
        bf = (torch.full((128 ** 3 // 8,), 255, dtype=torch.uint8, device=DEV) if state == "full"
              else torch.from_numpy(synthetic.morton_pack(scene.occupancy_bitfield(128))).to(DEV))
        for n_rays in (4096, 1 << 16):
            b = ds(n_rays, {"pixs": {"valid_uss": 0.4, "valid_tof": 0.4}})
            ro, rd = b["rays_o"], b["rays_d"]
            hits = ray_aabb_intersection(ro, rd, 0.5)
            noise = torch.rand(n_rays, device=DEV)
            rays_a, xyzs, dirs, deltas, ts, total = raymarching_train(ro, rd, hits, bf, 1, 0.5, 0.0, 128, 1024, noise=noise)
            S = int(total)
            if "march" in which:
                counter = torch.zeros(2, device=DEV, dtype=torch.int32)
                counts = torch.empty(n_rays, device=DEV, dtype=torch.int32)
                tmp = torch.empty(_lib.scan_tmp_ints(n_rays), device=DEV, dtype=torch.int32)
                ms = timeit(lambda: _lib.call("vn_march_train_count", ro, rd, hits, bf, noise, n_rays, 1, 128, 0.5, 0.0, 1024,
                                              counts, rays_a, counter, tmp))
                rec("march_count", ms, state=state, rays=n_rays, samples=S)
                ms = timeit(lambda: _lib.call("vn_march_train_write", ro, rd, hits, bf, noise, n_rays, 1, 128, 0.5, 0.0, rays_a,
                                              S, xyzs, dirs, deltas, ts, None))
                rec("march_write", ms, state=state, rays=n_rays, samples=S, gbs=round(S * 32 / ms / 1e6, 1))
            if "composite" in which:
                sig = torch.rand(S, device=DEV) * 20; rgbs = torch.rand(S, 3, device=DEV)
                tot = torch.empty(n_rays, dtype=torch.int32, device=DEV)
                op = torch.empty(n_rays, device=DEV); dp = torch.empty(n_rays, device=DEV)
                rgb = torch.empty(n_rays, 3, device=DEV); ws = torch.empty(S, device=DEV)
                ms = timeit(lambda: _lib.call("vn_composite_train_fwd", sig, rgbs, deltas, ts, rays_a, n_rays, S, 1e-4, tot, op,
                                              dp, rgb, ws))
                rec("composite_fwd", ms, state=state, rays=n_rays, samples=S, gbs=round(S * 28 / ms / 1e6, 1))
                dsig = torch.empty(S, device=DEV); drgb = torch.empty(S, 3, device=DEV)
                ms = timeit(lambda: _lib.call("vn_composite_train_bwd", sig, rgbs, deltas, ts, rays_a, n_rays, S, 1e-4, op, dp,
                                              rgb, None, dsig, drgb))
                rec("composite_bwd", ms, state=state, rays=n_rays, samples=S, gbs=round(S * 44 / ms / 1e6, 1))
            if "hash" in which and (n_rays == 4096 or state == "carved"):
                x = ((xyzs + 0.5)).clamp(0, 1).contiguous()
                for log2_T in (19, 22):
                    lv = _lib.hash_levels(16, 1024, 16, 2 ** log2_T)
                    table = torch.rand(2 * lv.total_entries, device=DEV)
                    o = torch.empty(S, 32, device=DEV)
                    dout = torch.randn(S, 32, device=DEV)
                    grad = torch.zeros(2 * lv.total_entries, device=DEV)
                    for flags in (0, 256):
                        ms = timeit(lambda: _lib.call("vn_hash_encode_fwd_f32", x, table, o, S, lv, flags))
                        rec("hash_fwd_f32", ms, state=state, S=S, log2_T=log2_T, flags=flags, gbs=round(S * 1164 / ms / 1e6, 1))
                    for flags in (0, 256, 32):
                        ms = timeit(lambda: _lib.call("vn_hash_encode_bwd_f32", x, dout, grad, S, lv, flags))
                        rec("hash_bwd_f32", ms, state=state, S=S, log2_T=log2_T, flags=flags, gbs=round(S * 1164 / ms / 1e6, 1))
                    # level-pair-plane layout ([8][S] float4) of the fast step
                    for flags in (512, 512 | 256, 512 | 2048, 512 | 256 | 2048):
                        ms = timeit(lambda: _lib.call("vn_hash_encode_fwd_f32", x, table, o, S, lv, flags))
                        rec("hash_fwd_f32", ms, state=state, S=S, log2_T=log2_T, flags=flags, gbs=round(S * 1164 / ms / 1e6, 1))
                    for flags in (512, 512 | 256, 512 | 1024, 512 | 256 | 1024):
                        ms = timeit(lambda: _lib.call("vn_hash_encode_bwd_f32", x, dout, grad, S, lv, flags))
                        rec("hash_bwd_f32", ms, state=state, S=S, log2_T=log2_T, flags=flags, gbs=round(S * 1164 / ms / 1e6, 1))
                    table_h = table.half().view(-1, 2)
                    oh = torch.empty(S, 16, 2, dtype=torch.float16, device=DEV)
                    ms = timeit(lambda: _lib.call("vn_hash_encode_fwd_f16", x, table_h, oh, S, lv, 0))
                    rec("hash_fwd_f16", ms, state=state, S=S, log2_T=log2_T, flags=0, gbs=round(S * 588 / ms / 1e6, 1))
                    douth = dout.half().view(S, 16, 2).contiguous()
                    ms = timeit(lambda: _lib.call("vn_hash_encode_bwd_f16", x, douth, grad, S, lv, 0))
                    rec("hash_bwd_f16", ms, state=state, S=S, log2_T=log2_T, flags=0, gbs=round(S * 1100 / ms / 1e6, 1))
                    del table, grad
    if "mlp" in which or "hash" in which:
        g = torch.Generator().manual_seed(0)
        xav = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) * (6.0 / (i + o)) ** 0.5).to(DEV)
        W = [xav(64, 32), xav(16, 64), xav(64, 32), xav(64, 64), xav(3, 64)]
        for S in (69710, 933574, 1 << 20):
            enc = torch.rand(S, 32, device=DEV); dirs = torch.randn(S, 3, device=DEV)
            sig = torch.empty(S, device=DEV); rgb = torch.empty(S, 3, device=DEV)
            ms = timeit(lambda: _lib.call("vn_mlp_fwd", enc, 0, dirs, *W, S, 0, sig, rgb, None))
            rec("mlp_fwd", ms, S=S, tflops=round(S * 18816 / ms / 1e9, 2), gbs=round(S * 156 / ms / 1e6, 1))
            dsig = torch.randn(S, device=DEV); drgb = torch.randn(S, 3, device=DEV)
            denc = torch.empty(S, 32, device=DEV); dW = [torch.zeros_like(w) for w in W]
            ms = timeit(lambda: _lib.call("vn_mlp_bwd", enc, 0, dirs, *W, S, 0, dsig, drgb, denc, *dW))
            rec("mlp_bwd", ms, S=S, tflops=round(S * 56448 / ms / 1e9, 2), gbs=round(S * 284 / ms / 1e6, 1))
            enc_c = enc.half().view(S, 4, 8).permute(1, 0, 2).contiguous()
            ms = timeit(lambda: _lib.call("vn_mlp_fwd", enc_c, 3, dirs, *W, S, 0, sig, rgb, None))
            rec("mlp_fwd_chunks", ms, S=S, tflops=round(S * 18816 / ms / 1e9, 2), gbs=round(S * 92 / ms / 1e6, 1))
            ms = timeit(lambda: _lib.call("vn_mlp_bwd", enc_c, 3, dirs, *W, S, 0, dsig, drgb, denc, *dW))
            rec("mlp_bwd_chunks", ms, S=S, tflops=round(S * 56448 / ms / 1e9, 2), gbs=round(S * 220 / ms / 1e6, 1))
            enc_c5 = torch.cat([enc_c, torch.rand(2, S, 8, device=DEV).half()], 0).contiguous()
            ms = timeit(lambda: _lib.call("vn_mlp_fwd", enc_c5, 5, None, *W, S, 0, sig, rgb, None))
            rec("mlp_fwd_chunks_sh", ms, S=S, tflops=round(S * 18816 / ms / 1e9, 2), gbs=round(S * 112 / ms / 1e6, 1))
            ms = timeit(lambda: _lib.call("vn_mlp_bwd", enc_c5, 5, None, *W, S, 0, dsig, drgb, denc, *dW))
            rec("mlp_bwd_chunks_sh", ms, S=S, tflops=round(S * 56448 / ms / 1e9, 2), gbs=round(S * 240 / ms / 1e6, 1))
            ms = timeit(lambda: _lib.call("vn_mlp_fwd", enc, 2, dirs, *W, S, 0, sig, rgb, None))
            rec("mlp_fwd_planar", ms, S=S, tflops=round(S * 18816 / ms / 1e9, 2), gbs=round(S * 156 / ms / 1e6, 1))
            ms = timeit(lambda: _lib.call("vn_mlp_bwd", enc, 2, dirs, *W, S, 0, dsig, drgb, denc, *dW))
            rec("mlp_bwd_planar", ms, S=S, tflops=round(S * 56448 / ms / 1e9, 2), gbs=round(S * 284 / ms / 1e6, 1))
    if "adam" in which:
        n = 11429472
        p, g, m, v = (torch.randn(n, device=DEV) for _ in range(4))
        v.abs_()
        found = torch.zeros(1, device=DEV); scale = torch.tensor([2.0 ** 19], device=DEV)
        ms = timeit(lambda: _lib.call("vn_adam_step", p, g, m, v, n, 1.0, 5e-3, 0.9, 0.999, 1e-15, 3, found, scale))
        rec("adam_step", ms, n=n, gbs=round(n * 28 / ms / 1e6, 1))
        ms = timeit(lambda: _lib.call("vn_grad_check", g, n, found))
        rec("grad_check", ms, n=n, gbs=round(n * 4 / ms / 1e6, 1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "kbench.jsonl"), "w") as f:
        for l in out:
            f.write(json.dumps(l) + "\n")


if __name__ == "__main__":
    main()
