"""GPU parity of the "next" rows through the C ABI: batch assembly
(vn_batch_assemble behind the DatasetBase mirror) and the NGPGrid kernels against the numpy
restatement in oracle/extras.py (itself pinned to the reference's code by golden_v2.npz).
Bars: everything bit-exact except the mean of the positive cells (rtol 1e-6, summation order)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
HERE = os.path.dirname(os.path.abspath(__file__))


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


@pytest.fixture(scope="module")
def g2():
    return np.load(os.path.join(HERE, "golden", "golden_v2.npz"))


@pytest.fixture(scope="module")
def vn():
    from virus_nerf_b200 import _lib
    _lib.lib()
    return _lib


def _density_cpu(x):
    x = x.detach().cpu()
    return torch.exp(3.0 * torch.sin(7.0 * x[:, 0]) * torch.cos(5.0 * x[:, 1]) + 2.0 * x[:, 2])


def _dataset(g2, extra_imgs=0):
    from virus_nerf_b200.datasets.dataset_base import DatasetBase
    cams = ["CAM1", "CAM3"]
    name2id = {c: int(i) for c, i in zip(cams, g2["f2_cam_ids"])}
    args = SimpleNamespace(device=torch.device(DEV))
    ds = DatasetBase(args, rgbs=T(g2["f2_rgbs"]), poses=T(g2["f2_poses"]),
                     directions_dict={c: T(g2["f2_dirs"][k]) for k, c in enumerate(cams)},
                     sensor_ids=T(g2["f2_sensor_ids"]), depths_dict={"USS": T(g2["f2_uss"]), "ToF": T(g2["f2_tof"])},
                     times=T(g2["f2_times"]), img_wh=(20, 12), sensor_name2id=lambda c: name2id[c])
    return ds


def test_batch_assemble_matches_oracle_and_reference(g2):
    from oracle import extras
    ds = _dataset(g2)
    rng = np.random.default_rng(5)
    for B, src in ((257, "golden"), (1, "rand"), (4096, "rand")):
        if src == "golden":
            ii, pp = g2["f2_img_idxs"], g2["f2_pix_idxs"]
        else:
            ii = rng.integers(0, 10, B).astype(np.int32); pp = rng.integers(0, 240, B).astype(np.int32)
        out = ds(img_idxs=T(ii), pix_idxs=T(pp))
        slot = np.full(10, -1, np.int64)
        for k, cid in enumerate(g2["f2_cam_ids"]):
            slot[g2["f2_sensor_ids"] == cid] = k
        ref = extras.batch_assemble(ii, pp, g2["f2_poses"], slot, g2["f2_dirs"], g2["f2_rgbs"],
                                    {"USS": g2["f2_uss"], "ToF": g2["f2_tof"]}, g2["f2_sensor_ids"], g2["f2_times"])
        for k in ("rays_o", "rays_d", "rgb", "time"):
            np.testing.assert_array_equal(out[k].cpu().numpy(), ref[k])
        np.testing.assert_array_equal(out["sensor_ids"].cpu().numpy(), ref["sensor_ids"])
        for s in ("USS", "ToF"):
            np.testing.assert_array_equal(out["depth"][s].cpu().numpy(), ref["depth"][s])
        if src == "golden":       # and directly against what the reference's DatasetBase returned
            np.testing.assert_allclose(out["rays_d"].cpu().numpy(), g2["f2_rays_d"], rtol=1e-6, atol=1e-7)
            np.testing.assert_array_equal(out["rays_o"].cpu().numpy(), g2["f2_rays_o"])
    assert not ds.indices_out_of_range()
    bad = ds(img_idxs=T(np.array([0, 99], np.int32)), pix_idxs=T(np.array([5, 5], np.int32)))
    assert torch.isnan(bad["rays_o"][1]).all() and not torch.isnan(bad["rays_o"][0]).any()
    assert ds.indices_out_of_range()
    empty = ds(img_idxs=T(np.zeros(0, np.int32)), pix_idxs=T(np.zeros(0, np.int32)))
    assert empty["rays_o"].shape == (0, 3)


def test_dataset_with_sampler_produces_engine_batches(g2):
    """Sampler (device RNG) + DatasetBase.__call__: the dict the train engine consumes"""
    from virus_nerf_b200.training.sampler import Sampler
    ds = _dataset(g2)
    args = SimpleNamespace(device=torch.device(DEV), seed=21, logger=SimpleNamespace(error=print),
                           training=SimpleNamespace(debug_mode=False, real_time_simulation=False))
    sensors = {"USS": SimpleNamespace(mask=T(g2["f2_mask_uss"])), "ToF": SimpleNamespace(mask=T(g2["f2_mask_tof"]))}
    ds.sampler = Sampler(args, dataset_len=10, img_wh=(20, 12), sensors_dict=sensors, times=ds.times)
    b = ds(batch_size=512, sampling_strategy={"imgs": "all", "pixs": {"valid_uss": 0.4, "valid_tof": 0.4}}, elapse_time=0.0)
    assert b["rays_o"].shape == (512, 3) and b["depth"]["USS"].shape == (512,)
    pix = b["pix_idxs"].cpu().numpy()
    assert g2["f2_mask_uss"][pix[:204]].all() and g2["f2_mask_tof"][pix[204:408]].all()      # int(0.4 * 512) = 204
    np.testing.assert_allclose(b["rays_d"].norm(dim=1).cpu().numpy(), 1.0, rtol=1e-5)


@pytest.mark.parametrize("G", [16, 64])
def test_ngp_kernels_match_oracle(vn, oracle_mod, G):
    from oracle import extras
    rng = np.random.default_rng(G)
    G3 = G ** 3
    occ = (rng.random(G3) * 2).astype(np.float32)
    occ[rng.random(G3) < 0.1] = -1.0
    occ[rng.random(G3) < 0.5] = 0.0
    thr = np.float32(0.7)
    # occupied-cell rank queries (incl. ranks beyond the count: reduced modulo it)
    M = 5000
    rand_idx = rng.integers(0, G3, M)
    tmp = torch.zeros(vn.ngp_select_tmp_ints(G3), dtype=torch.int32, device=DEV)
    out = torch.empty(M, dtype=torch.int64, device=DEV)
    vn.call("vn_ngp_sample_occupied", T(occ), G3, float(thr), T(rand_idx), M, tmp, out)
    np.testing.assert_array_equal(out.cpu().numpy(), extras.ngp_sample_occupied(occ, thr, rand_idx))
    vn.call("vn_ngp_sample_occupied", T(np.zeros(G3, np.float32)), G3, float(thr), T(rand_idx), M, tmp, out)
    assert (out == -1).all()
    # positions
    coords = rng.integers(0, G, (M, 3)).astype(np.int32)
    noise = rng.random((M, 3)).astype(np.float32)
    s = 0.5
    xyz = torch.empty(M, 3, device=DEV)
    vn.call("vn_ngp_cell_positions", T(coords), T(noise), M, G, float(np.float32(s - s / G)), float(np.float32(s / G)), xyz)
    np.testing.assert_array_equal(xyz.cpu().numpy(), extras.ngp_cell_positions(coords, noise, G, s))
    # scatter with duplicates and skipped (-1) entries + decayed maximum; tmp is re-zeroed, winner restored
    indices = rng.integers(0, G3, M).astype(np.int64)
    indices[::7] = indices[3]; indices[5::11] = -1
    sig = (rng.random(M) * 3).astype(np.float32)
    occ_d = T(occ); tmp_d = torch.zeros(G3, device=DEV); win = torch.full((G3,), -1, dtype=torch.int32, device=DEV)
    vn.call("vn_ngp_grid_update", occ_d, tmp_d, win, G3, T(indices), T(sig), M, 0.95, None)
    ref = extras.ngp_grid_update(occ, indices, sig, 0.95)
    np.testing.assert_array_equal(occ_d.cpu().numpy(), ref)
    assert float(tmp_d.abs().max()) == 0.0 and int((win != -1).sum()) == 0
    dc = (0.1 + 0.85 * rng.random(G3)).astype(np.float32)
    occ_d = T(occ)
    vn.call("vn_ngp_grid_update", occ_d, tmp_d, win, G3, T(indices), T(sig), M, 0.95, T(dc))
    np.testing.assert_array_equal(occ_d.cpu().numpy(), extras.ngp_grid_update(occ, indices, sig, 0.95, dc))
    # threshold + bitfield
    scratch = torch.zeros(vn.ngp_threshold_tmp_bytes() // 8, dtype=torch.float64, device=DEV)
    thr2 = torch.zeros(2, device=DEV); bf = torch.zeros(G3 // 8, dtype=torch.uint8, device=DEV)
    for dthr in (5.91, 0.3):
        vn.call("vn_ngp_threshold_pack", T(ref), G3, dthr, scratch, thr2, bf)
        mean, t, rbf = extras.ngp_threshold_pack(ref, dthr)
        np.testing.assert_allclose(thr2.cpu().numpy(), [mean, t], rtol=1e-6)
        if float(thr2[1]) == float(t):
            np.testing.assert_array_equal(bf.cpu().numpy(), rbf)
    vn.call("vn_ngp_threshold_pack", T(-np.ones(G3, np.float32)), G3, 0.3, scratch, thr2, bf)
    assert torch.isnan(thr2).all() and int(bf.sum()) == 0            # mean of nothing is nan; nothing is occupied


@pytest.mark.parametrize("tag", ["warm", "samp"])
def test_ngp_grid_module_reproduces_the_reference_update(g2, tag):
    """NGPGrid.update on the GPU, fed the random draws the reference made, ends in the reference's grid"""
    from virus_nerf_b200.modules.ngp_grid import NGPGrid
    args = SimpleNamespace(device=torch.device(DEV), model=SimpleNamespace(scale=0.5))
    grid = NGPGrid(args, 16, fct_density=lambda x: _density_cpu(x).to(DEV))
    grid.occ_morton_grid = T(g2[f"f3_{tag}_before"]).reshape(1, -1).clone()
    thr = float(g2["f3_density_threshold"])
    if tag == "warm":
        # the module enumerates the cells x-major, the reference's kornia meshgrid in another order: align the noise
        ref_index = {int(i): k for k, i in enumerate(g2["f3_all_indices"])}
        mine = grid.getAllCells()[0][0].cpu().numpy()
        noise = T(g2["f3_warm_noise"][[ref_index[int(i)] for i in mine]])
        grid.update(density_threshold=thr, warmup=True, noise=[noise])
    else:
        draws = [T(g2["f3_samp_coords1"]), T(g2["f3_samp_rand_idx"])]
        orig = torch.randint
        torch.randint = lambda *a, **k: draws.pop(0)
        try:
            grid.update(density_threshold=thr, warmup=False, noise=[T(g2["f3_samp_noise"])])
        finally:
            torch.randint = orig
    np.testing.assert_array_equal(grid.occ_morton_grid.cpu().numpy().reshape(-1), g2[f"f3_{tag}_after"])
    np.testing.assert_allclose(grid.threshold, float(g2[f"f3_{tag}_threshold"]), rtol=1e-6)
    bf, ref_bf = grid.getBitfield().cpu().numpy(), g2[f"f3_{tag}_bitfield"]
    if not np.array_equal(bf, ref_bf):
        diff = np.unpackbits(bf ^ ref_bf, bitorder="little").nonzero()[0]
        assert np.allclose(g2[f"f3_{tag}_after"][diff], grid.threshold, rtol=1e-6)


def test_engine_trains_with_the_ngp_grid():
    """grid_type 'ngp' (args/ethz_usstof_win.json): warm-up and sampled updates drive the fast step"""
    from virus_nerf_b200 import synthetic
    from virus_nerf_b200.engine import TrainEngine
    args = synthetic.make_args(device=DEV, batch_size=512, grid_type="ngp")
    args.ngp_grid.update_interval, args.ngp_grid.warmup_steps = 4, 8
    ds = synthetic.SyntheticDataset(pool_size=1 << 13, n_images=8, device=DEV)
    eng = TrainEngine(args, ds, DEV)
    losses = []
    for it in range(16):
        losses.append(float(eng.step_fast(ds(512, args.training.sampling_strategy))))
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    g = eng.model.occupancy_grid
    assert g.getBitfield().numel() == 128 ** 3 // 8 and 0 < int(g.getBitfield().count_nonzero())
    assert np.isfinite(g.threshold) and g.threshold <= float(np.float32(0.01 * 1024 / 3 ** 0.5))


@pytest.mark.parametrize("kind", ["ethz", "rh2"])
def test_pool_gather_equals_advanced_indexing(kind):
    """vn_pool_gather (one launch per batch segment) returns exactly the batches of the torch advanced-indexing path, for
    the training strategy (three segments) and the occupancy-update strategies (valid-sensor subsets)"""
    from virus_nerf_b200 import synthetic
    ds = synthetic.SyntheticDataset(kind=kind, pool_size=1 << 13, n_images=8, device=DEV, seed=3)
    strategies = [{"imgs": "all", "pixs": {"valid_uss": 0.4, "valid_tof": 0.4}}, {"imgs": "all", "pixs": "valid_uss"},
                  {"imgs": "all", "pixs": "valid_tof"}, {"imgs": "all", "pixs": "random"}, None]
    for strat in strategies:
        for B in (1, 257, 4096):
            ds.fast_gather = True; ds.gen.manual_seed(11)
            fast = ds(B, strat)
            ds.fast_gather = False; ds.gen.manual_seed(11)
            slow = ds(B, strat)
            for k in ("rays_o", "rays_d", "rgb"):
                assert fast[k].is_contiguous() and fast[k].shape == slow[k].shape and torch.equal(fast[k], slow[k]), (strat, B, k)
            assert set(fast["depth"]) == set(slow["depth"]) == set(ds.sensors)
            for k in ds.sensors:
                assert torch.equal(torch.nan_to_num(fast["depth"][k], nan=-1.0), torch.nan_to_num(slow["depth"][k], nan=-1.0))
            assert fast["depth_valid_by_construction"] == slow["depth_valid_by_construction"]
