"""CPU restatement (numpy, strict f32 with one rounding per operation) of the "next" rows of
SURVEY section 8(f): batch assembly (datasets/dataset_base.py:23-76,194-243 + datasets/ray_utils.py:
51-80) and the Instant-NGP baseline grid (modules/ngp_grid.py:37-152).

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and nothing in the product package); pinned against
the reference's own code through tests/golden/golden_v2.npz (tests/golden/make_golden_v2.py).
"""
import numpy as np

f32 = np.float32


def batch_assemble(img_idxs, pix_idxs, poses, cam_slot, directions, rgbs, depth_maps, sensor_ids, times):
    """dataset_base.py:52-75: rays_o = c2w[:, 3]; rays_d = directions[pix] @ c2w[:, :3]^T (ray_utils.py:67-74,
    products and sums rounded separately, k = 0, 1, 2); rgb / depth / id / time gathers"""
    i = np.asarray(img_idxs, np.int64); p = np.asarray(pix_idxs, np.int64)
    poses = np.asarray(poses, f32)
    cam = np.asarray(cam_slot, np.int64)[i]
    d = np.asarray(directions, f32)[cam, p]                       # [B,3]
    R = poses[i][:, :, :3]                                        # [B,3,3]
    acc = d[:, None, 0] * R[:, :, 0]
    acc = (acc + d[:, None, 1] * R[:, :, 1]).astype(f32)
    acc = (acc + d[:, None, 2] * R[:, :, 2]).astype(f32)
    out = {"rays_o": poses[i][:, :, 3].copy(), "rays_d": acc, "rgb": np.asarray(rgbs, f32)[i, p, :3].copy(),
           "depth": {k: np.asarray(v, f32)[i, p].copy() for k, v in depth_maps.items()},
           "sensor_ids": np.asarray(sensor_ids)[i].copy(), "time": np.asarray(times, f32)[i].copy()}
    return out


def ngp_sample_occupied(occ, threshold, rand_idx):
    """ngp_grid.py:53-60: nonzero(occ > thr)[rand_idx % count] (ascending = Morton order); -1 if none"""
    nz = np.nonzero(np.asarray(occ, f32) > f32(threshold))[0]
    if nz.size == 0:
        return np.full(len(rand_idx), -1, np.int64)
    return nz[np.asarray(rand_idx, np.int64) % nz.size].astype(np.int64)


def ngp_cell_positions(coords, noise, grid_size, s):
    """ngp_grid.py:130-135"""
    hgs = s / grid_size
    span, half = f32(s - hgs), f32(hgs)
    v = (np.asarray(coords).astype(f32) / f32(grid_size - 1)).astype(f32)
    v = ((v * f32(2)).astype(f32) - f32(1)).astype(f32)
    v = (v * span).astype(f32)
    jit = ((np.asarray(noise, f32) * f32(2)).astype(f32) - f32(1)).astype(f32)
    jit = (jit * half).astype(f32)
    return (v + jit).astype(f32)


def ngp_grid_update(occ, indices, sigmas, decay, decay_cells=None):
    """ngp_grid.py:121,136,148-152: tmp = zeros; tmp[indices] = sigmas (last occurrence wins, CPU
    index_put_ order; negative indices are skipped); occ = where(occ < 0, occ, maximum(occ * decay, tmp))"""
    occ = np.asarray(occ, f32).copy()
    tmp = np.zeros_like(occ)
    idx = np.asarray(indices, np.int64)
    ok = idx >= 0
    tmp[idx[ok]] = np.asarray(sigmas, f32)[ok]          # numpy assigns in order: the last duplicate wins
    dec = f32(decay) if decay_cells is None else np.asarray(decay_cells, f32)
    d = (occ * dec).astype(f32)
    new = np.maximum(d, tmp)                            # propagates NaN like torch.maximum
    return np.where(occ < 0, occ, new).astype(f32)


def ngp_threshold_pack(occ_all, density_threshold):
    """ngp_grid.py:155-163: mean of the positive cells (float64 accumulation), python min(mean, thr), packbits"""
    occ_all = np.asarray(occ_all, f32).reshape(-1)
    pos = occ_all[occ_all > 0]
    mean = f32(pos.astype(np.float64).sum() / pos.size) if pos.size else f32(np.nan)
    thr = f32(density_threshold) if f32(density_threshold) < mean else mean
    bits = (occ_all > thr).astype(np.uint8).reshape(-1, 8)
    return mean, thr, np.packbits(bits, axis=1, bitorder="little").reshape(-1)
