"""Drop-in for the reference's modules/ray_march.py (raymarching_train / raymarching_test)."""
import torch

from .. import _lib
from .utils import torch_type


def raymarching_train(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, grid_size,
                      max_samples, noise=None):
    """ray_march.py:127-195.  Returns (rays_a, xyzs, dirs, deltas, ts, total_samples).

    Differences from the reference, both deliberate: rays_a rows are in ray order
    with start indices = exclusive scan of the per-ray counts (the reference's atomic counter
    gives a run-dependent permutation of the same segments), and the output buffers are sized
    by the counted total instead of N*max_samples rows.  `noise` may be passed in to reproduce
    a run; by default it is torch.rand_like as in the reference (:139)."""
    N = rays_o.shape[0]
    dev = rays_o.device
    if noise is None:
        noise = torch.rand_like(rays_o[:, 0])
    rays_o = rays_o.contiguous(); rays_d = rays_d.contiguous(); hits_t = hits_t.contiguous()
    noise = noise.contiguous()
    counter = torch.zeros(2, device=dev, dtype=torch.int32)
    rays_a = torch.empty(N, 3, device=dev, dtype=torch.int32)
    counts = torch.empty(N, device=dev, dtype=torch.int32)
    scan_tmp = torch.empty(_lib.scan_tmp_ints(N), device=dev, dtype=torch.int32)
    _lib.call("vn_march_train_count", rays_o, rays_d, hits_t, density_bitfield, noise, N, int(cascades),
              int(grid_size), float(scale), float(exp_step_factor), int(max_samples), counts, rays_a, counter,
              scan_tmp)
    total_samples = counter[0]
    total = int(total_samples.item())   # the reference syncs here too (slicing with a 0-dim tensor, :188-193)
    xyzs = torch.empty(total, 3, device=dev, dtype=torch_type)
    dirs = torch.empty(total, 3, device=dev, dtype=torch_type)
    deltas = torch.empty(total, device=dev, dtype=torch_type)
    ts = torch.empty(total, device=dev, dtype=torch_type)
    _lib.call("vn_march_train_write", rays_o, rays_d, hits_t, density_bitfield, noise, N, int(cascades),
              int(grid_size), float(scale), float(exp_step_factor), rays_a, total, xyzs, dirs, deltas, ts, None)
    return rays_a, xyzs, dirs, deltas, ts, total_samples


def raymarching_test(rays_o, rays_d, hits_t, alive_indices, density_bitfield, cascades, scale, exp_step_factor,
                     grid_size, max_samples):
    """ray_march.py:271-335.  Returns (packed_info, ray_indices, deltas, ts); mutates hits_t[:,0]
    (hits_t must be contiguous for the in-place update to be visible, as in the reference)."""
    A = alive_indices.size(0)
    dev = rays_o.device
    n_slots = A * int(max_samples)
    ray_indices = torch.empty(n_slots, device=dev, dtype=torch.long)
    valid_mask = torch.zeros(n_slots, device=dev, dtype=torch.uint8)
    deltas = torch.empty(n_slots, device=dev, dtype=rays_o.dtype)
    ts = torch.empty(n_slots, device=dev, dtype=rays_o.dtype)
    samples_counter = torch.empty(A, device=dev, dtype=torch.int32)
    assert hits_t.is_contiguous()
    _lib.call("vn_march_test", rays_o.contiguous(), rays_d.contiguous(), hits_t, alive_indices.contiguous(), A,
              density_bitfield, int(cascades), int(grid_size), float(scale), float(exp_step_factor),
              int(max_samples), ray_indices, valid_mask, deltas, ts, samples_counter)
    # compaction on device (the reference uses cumsum + three boolean-mask gathers, :328-335)
    packed_info = torch.empty(A, 2, device=dev, dtype=torch.long)
    ri_out = torch.empty(n_slots, device=dev, dtype=torch.long)
    de_out = torch.empty(n_slots, device=dev, dtype=rays_o.dtype)
    ts_out = torch.empty(n_slots, device=dev, dtype=rays_o.dtype)
    total = torch.zeros(1, device=dev, dtype=torch.long)
    scan_tmp = torch.empty(_lib.scan_tmp_ints(A), device=dev, dtype=torch.int32)
    _lib.call("vn_march_test_compact", samples_counter, A, int(max_samples), ray_indices, deltas, ts, packed_info,
              ri_out, de_out, ts_out, total, scan_tmp)
    n = int(total.item())
    return packed_info, ri_out[:n], de_out[:n], ts_out[:n]
