"""world_size-2 gloo test (CPU) of the data-parallel host logic: with per-rank loss SUMS and
globally allreduced COUNTS, summing the per-rank gradients reproduces the single-process
gradient of the concatenated batch (masked means need global counts)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _batch(seed, n):
    g = torch.Generator().manual_seed(seed)
    uss = torch.rand(n, generator=g); uss[torch.rand(n, generator=g) < 0.4] = float("nan")
    tof = torch.rand(n, generator=g); tof[torch.rand(n, generator=g) < 0.7] = float("nan")
    return {"rgb": torch.rand(n, 3, generator=g), "depth": {"USS": uss, "ToF": tof},
            "x": torch.rand(n, 8, generator=g)}


def _model_out(w, data):
    h = data["x"] @ w
    return {"rgb": torch.sigmoid(h[:, :3]), "depth": h[:, 3].abs()}


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from virus_nerf_b200 import synthetic
    from virus_nerf_b200.engine import Loss
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    args = synthetic.make_args(device="cpu")
    torch.manual_seed(0)
    w = torch.randn(8, 4, requires_grad=True)
    data = _batch(100 + rank, 64 + 32 * rank)          # ragged shards
    loss, _ = Loss(args)(_model_out(w, data), data, world_size=world)
    loss.backward()
    g = w.grad.clone()
    dist.all_reduce(g)                                  # the engine's gradient exchange: SUM
    lt = loss.detach().clone()
    dist.all_reduce(lt)
    if rank == 0:
        q.put((g.numpy(), float(lt)))
    dist.destroy_process_group()


def test_dp_loss_and_gradient_match_single_process():
    from virus_nerf_b200 import synthetic
    from virus_nerf_b200.engine import Loss
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    g_dp, loss_dp = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    args = synthetic.make_args(device="cpu")
    torch.manual_seed(0)
    w = torch.randn(8, 4, requires_grad=True)
    a, b = _batch(100, 64), _batch(101, 96)
    data = {"rgb": torch.cat([a["rgb"], b["rgb"]]), "x": torch.cat([a["x"], b["x"]]),
            "depth": {k: torch.cat([a["depth"][k], b["depth"][k]]) for k in ("USS", "ToF")}}
    loss, _ = Loss(args)(_model_out(w, data), data, world_size=1)
    loss.backward()
    assert abs(loss_dp - float(loss)) <= 1e-5 * abs(float(loss))
    np.testing.assert_allclose(g_dp, w.grad.numpy(), rtol=1e-4, atol=1e-6)
