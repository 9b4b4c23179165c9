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


# ---- sharded optimiser protocol (vn_p2p_reduce_adam): slice partition + equivalence with the dense step ----------
def _sharded_worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import oracle
    from virus_nerf_b200 import _lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    p = rng.normal(size=n).astype(np.float32)                     # identical replicas
    m = np.zeros(n, np.float32); v = np.zeros(n, np.float32)      # only the own slice is ever touched
    lo, hi = _lib.p2p_slice(n, rank, world)
    for step in range(1, 4):
        g = torch.from_numpy(np.random.default_rng(100 * step + rank).normal(size=n).astype(np.float32))
        # reduce-scatter: rank r needs only slice r of the sum (gloo: allreduce, then keep the slice)
        dist.all_reduce(g)
        ps, ms, vs = p[lo:hi].copy(), m[lo:hi].copy(), v[lo:hi].copy()
        if hi > lo:
            oracle.adam_step(ps, g.numpy()[lo:hi].copy(), ms, vs, 1.0, 1e-2, 0.9, 0.999, 1e-15, step)
        m[lo:hi], v[lo:hi] = ms, vs
        # parameter push: every replica receives every owner's updated slice
        mine = torch.zeros(n); mine[lo:hi] = torch.from_numpy(ps)
        dist.all_reduce(mine)                                     # slices are disjoint: the sum is the concatenation
        p = mine.numpy().copy()
    q.put((rank, p, lo, hi))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [4096, 4 * 1001])                   # 1001 float4 chunks: ragged slices
def test_sharded_optimiser_protocol_equals_dense_adam(n, oracle_mod):
    from virus_nerf_b200 import _lib
    world = 2
    # the slices tile [0, n) exactly, for every world size the library supports
    for w in range(1, 9):
        edges = [_lib.p2p_slice(n, r, w) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:])) and all(lo % 4 == 0 and hi % 4 == 0 for lo, hi in edges)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + (n % 7)
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, n, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    outs = sorted([q.get(timeout=120) for _ in range(world)], key=lambda o: o[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    np.testing.assert_array_equal(outs[0][1], outs[1][1])                   # replicas bit-identical
    # dense reference: the same summed gradients through one Adam over the whole buffer
    rng = np.random.default_rng(5)
    p = rng.normal(size=n).astype(np.float32); m = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    for step in range(1, 4):
        g = sum(np.random.default_rng(100 * step + r).normal(size=n).astype(np.float32) for r in range(world))
        oracle_mod.adam_step(p, g.astype(np.float32), m, v, 1.0, 1e-2, 0.9, 0.999, 1e-15, step)
    np.testing.assert_array_equal(outs[0][1], p)


# ---- vn_p2p_step protocol: the inf flag is an INPUT (each rank's own check), combined with MAX in the start barrier -------
def _step_worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import oracle
    from virus_nerf_b200 import _lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(9)
    p = rng.normal(size=n).astype(np.float32)
    m = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    lo, hi = _lib.p2p_slice(n, rank, world)
    scale, applied, log = 2.0 ** 19, 0, []
    for step in range(1, 6):
        g = np.random.default_rng(100 * step + rank).normal(size=n).astype(np.float32) * np.float32(scale)
        if step == 2 and rank == world - 1:
            g[17] = np.inf                                          # ONE rank overflows
        if step == 4 and rank == 0:
            g[n - 3] = np.nan
        found = torch.tensor([0.0 if np.isfinite(g).all() else 1.0])   # the rank's OWN check (contributions are bounded)
        dist.all_reduce(found, op=dist.ReduceOp.MAX)                 # start barrier payload
        gt = torch.from_numpy(np.nan_to_num(g, nan=0.0, posinf=0.0, neginf=0.0))
        dist.all_reduce(gt)
        if float(found) != 0.0:
            scale *= 0.5                                             # GradScaler backoff, nothing applied, count unchanged
        else:
            applied += 1
            ps, ms, vs = p[lo:hi].copy(), m[lo:hi].copy(), v[lo:hi].copy()
            if hi > lo:
                oracle.adam_step(ps, gt.numpy()[lo:hi].copy(), ms, vs, 1.0 / scale, 1e-2, 0.9, 0.999, 1e-15, applied)
            m[lo:hi], v[lo:hi] = ms, vs
            mine = torch.zeros(n); mine[lo:hi] = torch.from_numpy(ps)
            dist.all_reduce(mine)
            p = mine.numpy().copy()
        log.append((float(found), scale, applied))
    q.put((rank, p, log))
    dist.destroy_process_group()


def test_one_kernel_step_protocol_skips_everywhere_and_keeps_the_step_count(oracle_mod):
    """what vn_p2p_step does, on CPU: every rank contributes its OWN inf flag, all ranks skip together (scale halves,
    Adam's step count -- the bias corrections -- does not advance), applied steps equal the dense optimiser"""
    n, world = 4 * 501, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_step_worker, args=(r, world, port, n, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    outs = sorted([q.get(timeout=120) for _ in range(world)], key=lambda o: o[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] == outs[1][2]
    assert [x[0] for x in outs[0][2]] == [0.0, 1.0, 0.0, 1.0, 0.0]
    assert [x[2] for x in outs[0][2]] == [1, 1, 2, 2, 3] and outs[0][2][-1][1] == 2.0 ** 17
    # dense reference: torch semantics of GradScaler.step + Adam over the whole buffer
    rng = np.random.default_rng(9)
    p = rng.normal(size=n).astype(np.float32); m = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    scale, applied = 2.0 ** 19, 0
    for step in range(1, 6):
        gs = [np.random.default_rng(100 * step + r).normal(size=n).astype(np.float32) * np.float32(scale) for r in range(world)]
        if step in (2, 4):
            scale *= 0.5
            continue
        applied += 1
        oracle_mod.adam_step(p, (gs[0] + gs[1]).astype(np.float32), m, v, 1.0 / scale, 1e-2, 0.9, 0.999, 1e-15, applied)
    np.testing.assert_array_equal(outs[0][1], p)


def test_strong_scaling_shards_tile_the_global_batch():
    """bench.py --scaling strong / SURVEY 8(e): rank r takes rays [r N / n, (r + 1) N / n) of the global batch"""
    for n_global in (4096, 262144, 1000, 7):
        for world in (1, 2, 3, 4, 8):
            edges = [((n_global * r) // world, (n_global * (r + 1)) // world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n_global
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
