"""CPU tests of the "next" rows: the numpy restatement in
oracle/extras.py against goldens produced by the reference's own code
(tests/golden/make_golden_v2.py -> golden_v2.npz), and the Sampler mirror against the reference's
index stream for the same torch seed."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def g2():
    return np.load(os.path.join(HERE, "golden", "golden_v2.npz"))


def _density(x):
    x = torch.from_numpy(np.ascontiguousarray(x))
    return torch.exp(3.0 * torch.sin(7.0 * x[:, 0]) * torch.cos(5.0 * x[:, 1]) + 2.0 * x[:, 2]).numpy()


def test_batch_assemble_oracle_matches_reference(g2):
    from oracle import extras
    cam_ids = g2["f2_cam_ids"]
    slot = np.full(len(g2["f2_sensor_ids"]), -1, np.int64)
    for k, cid in enumerate(cam_ids):
        slot[g2["f2_sensor_ids"] == cid] = k
    out = extras.batch_assemble(g2["f2_img_idxs"], g2["f2_pix_idxs"], g2["f2_poses"], slot, g2["f2_dirs"], g2["f2_rgbs"],
                                {"USS": g2["f2_uss"], "ToF": g2["f2_tof"]}, g2["f2_sensor_ids"], g2["f2_times"])
    np.testing.assert_array_equal(out["rays_o"], g2["f2_rays_o"])
    np.testing.assert_allclose(out["rays_d"], g2["f2_rays_d"], rtol=1e-6, atol=1e-7)     # torch bmm summation order
    np.testing.assert_array_equal(out["rgb"], g2["f2_rgb"])
    np.testing.assert_array_equal(out["depth"]["USS"], g2["f2_out_uss"])                 # NaN == NaN here
    np.testing.assert_array_equal(out["depth"]["ToF"], g2["f2_out_tof"])
    np.testing.assert_array_equal(out["sensor_ids"], g2["f2_out_ids"])
    np.testing.assert_array_equal(out["time"], g2["f2_out_time"])


def test_sampler_reproduces_the_reference_index_stream(g2):
    from virus_nerf_b200.training.sampler import Sampler
    args = SimpleNamespace(device=torch.device("cpu"), seed=21, logger=SimpleNamespace(error=print),
                           training=SimpleNamespace(debug_mode=False, real_time_simulation=False))
    sensors = {"USS": SimpleNamespace(mask=torch.from_numpy(g2["f2_mask_uss"])),
               "ToF": SimpleNamespace(mask=torch.from_numpy(g2["f2_mask_tof"]))}
    sm = Sampler(args=args, dataset_len=10, img_wh=(20, 12), sensors_dict=sensors, times=torch.from_numpy(g2["f2_times"]))
    strategies = [{"imgs": "all", "pixs": {"valid_uss": 0.4, "valid_tof": 0.4}}, {"imgs": "same", "pixs": "random"},
                  {"imgs": "all", "pixs": "valid_tof"}, {"imgs": "all", "pixs": "entire_img"}]
    torch.manual_seed(1234)
    for k, st in enumerate(strategies):
        ii, pp = sm(batch_size=64, sampling_strategy=st, elapse_time=0.0)
        np.testing.assert_array_equal(ii.numpy(), g2[f"f2_sampler{k}_img"])
        np.testing.assert_array_equal(pp.numpy(), g2[f"f2_sampler{k}_pix"])
        assert ii.dtype == torch.int32


@pytest.mark.parametrize("tag", ["warm", "samp", "samp2"])
def test_ngp_grid_oracle_matches_reference(g2, oracle_mod, tag):
    from oracle import extras
    Gs, s, thr = 16, 0.5, float(g2["f3_density_threshold"])
    before = g2[f"f3_{tag}_before"]
    if tag == "warm":
        indices, coords = g2["f3_all_indices"], g2["f3_all_coords"]
    else:
        coords1 = g2[f"f3_{tag}_coords1"]
        indices1 = oracle_mod.morton3d(coords1).astype(np.int64)
        indices2 = extras.ngp_sample_occupied(before, thr, g2[f"f3_{tag}_rand_idx"])
        assert (indices2 >= 0).all()
        coords2 = oracle_mod.morton3d_invert(indices2.astype(np.int32))
        indices, coords = np.concatenate([indices1, indices2]), np.concatenate([coords1, coords2])
    xyz = extras.ngp_cell_positions(coords, g2[f"f3_{tag}_noise"], Gs, s)
    after = extras.ngp_grid_update(before, indices, _density(xyz), 0.95)
    np.testing.assert_array_equal(after, g2[f"f3_{tag}_after"])
    mean, t, bf = extras.ngp_threshold_pack(after, thr)
    np.testing.assert_allclose(float(t), float(g2[f"f3_{tag}_threshold"]), rtol=1e-6)
    ref_bf = g2[f"f3_{tag}_bitfield"]
    if not np.array_equal(bf, ref_bf):       # a cell within rounding of the mean may flip: allow only those
        diff = np.unpackbits(bf ^ ref_bf, bitorder="little").nonzero()[0]
        assert np.allclose(after[diff], float(t), rtol=1e-6)


def test_evaluation_inputs_match_reference(g2):
    """createScanRays / createScanPos (helpers/geometric_fcts.py:77-150) as the reference computes them"""
    from virus_nerf_b200.training.evaluation import createScanRays, createScanPos
    so, sd = createScanRays(torch.from_numpy(g2["f4_origins"]), angle_res=48)
    np.testing.assert_array_equal(so.numpy(), g2["f4_scan_o"])
    np.testing.assert_array_equal(sd.numpy(), g2["f4_scan_d"])
    assert (sd[:, 2] == 0).all()
    pos = createScanPos(res_map=9, height_c=-0.05, num_avg_heights=3, tolerance_c=0.02, cube_min=-0.5, cube_max=0.5,
                        device="cpu")
    np.testing.assert_array_equal(pos.numpy(), g2["f4_scan_pos"])
