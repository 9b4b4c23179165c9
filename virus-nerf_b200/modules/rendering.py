"""Drop-in for the reference's modules/rendering.py (render driver, train and test paths)."""
import torch

from .intersection import ray_aabb_intersection
from .ray_march import raymarching_test, raymarching_train
from .volume_render_test import composite_test

MAX_SAMPLES = 1024
NEAR_DISTANCE = 0.01


def render(model, rays_o, rays_d, test_time=False, exp_step_factor=0, T_threshold=1e-4, max_samples=MAX_SAMPLES):
    """rendering.py:12-57"""
    hits_t = ray_aabb_intersection(rays_o.contiguous(), rays_d.contiguous(), model.scale)
    if test_time:
        if exp_step_factor == 0 and max_samples >= MAX_SAMPLES and getattr(model, "fused_test_render", True):
            return __render_rays_test_fused(model, rays_o, rays_d, hits_t, T_threshold=T_threshold)
        return __render_rays_test(model, rays_o, rays_d, hits_t, exp_step_factor=exp_step_factor,
                                  T_threshold=T_threshold, max_samples=max_samples)
    return __render_rays_train(model, rays_o, rays_d, hits_t, exp_step_factor=exp_step_factor,
                               T_threshold=T_threshold)


@torch.no_grad()
def __render_rays_test(model, rays_o, rays_d, hits_t, exp_step_factor=0, T_threshold=1e-4, max_samples=MAX_SAMPLES):
    """rendering.py:61-158: iterative march / model / composite with alive-ray compaction"""
    results = {}
    N_rays = len(rays_o)
    device = rays_o.device
    opacity = torch.zeros(N_rays, device=device)
    depth = torch.zeros(N_rays, device=device)
    rgb = torch.zeros(N_rays, 3, device=device)

    samples = total_samples = 0
    alive_indices = torch.arange(N_rays, device=device)
    min_samples = 1 if exp_step_factor == 0 else 4                       # :94

    while samples < max_samples:
        N_alive = len(alive_indices)
        if N_alive == 0:
            break
        N_samples = max(min(N_rays // N_alive, 64), min_samples)         # :102
        samples += N_samples

        pack_info, ray_indices, deltas, ts = raymarching_test(
            rays_o, rays_d, hits_t, alive_indices, model.occupancy_grid.getBitfield(), model.cascades,
            model.scale, exp_step_factor, model.grid_size, N_samples)
        if ray_indices.shape[0] == 0:
            break
        ray_o_local = rays_o[ray_indices, :3]
        ray_d_local = rays_d[ray_indices, :3]
        xyzs = ray_o_local + ts[:, None] * ray_d_local                   # :126
        dirs = ray_d_local

        sigmas, rgbs = model(xyzs, dirs)

        composite_test(sigmas, rgbs, deltas, ts, pack_info, alive_indices, T_threshold, opacity, depth, rgb)
        alive_indices = alive_indices[alive_indices >= 0]                # :144
        total_samples += pack_info[:, 1].sum()

    results['opacity'] = opacity
    results['depth'] = depth
    results['rgb'] = rgb
    results['total_samples'] = total_samples

    if exp_step_factor == 0:
        rgb_bg = torch.ones(3, device=device)
    else:
        rgb_bg = torch.zeros(3, device=device)
    results['rgb'] += rgb_bg * (1 - opacity)[:, None]                    # :156
    return results


@torch.no_grad()
def __render_rays_test_fused(model, rays_o, rays_d, hits_t, T_threshold=1e-4):
    """Test-time render without the round loop, for exp_step_factor == 0 (constant step).

    The reference's loop (rendering.py:96-145) marches every alive ray a few samples per round,
    evaluates the model and composites, up to ~1024 rounds of ~15 tiny launches per 8192-ray
    chunk.  With a constant step dt = sqrt(3)/1024 a ray inside the unit cube has at most 1024
    lattice points, so the per-ray budget (the sum of the round sizes, >= 1024) can never
    truncate it: the rounds concatenate to ONE continuous march (same lattice, no jitter), and a
    ray contributes exactly the samples whose incoming transmittance exceeds T_threshold -- which
    is what the training compositor computes.  So: march all (a6 kernels, noise = 0) -> model ->
    composite (a8 kernel): 6 launches per call.  total_samples reports the samples composited."""
    N = rays_o.shape[0]
    noise = torch.zeros(N, device=rays_o.device, dtype=torch.float32)
    rays_a, xyzs, dirs, deltas, ts, _ = raymarching_train(
        rays_o, rays_d, hits_t, model.occupancy_grid.getBitfield(), model.cascades, model.scale, 0.0,
        model.grid_size, MAX_SAMPLES, noise=noise)
    sigmas, rgbs = model(xyzs, dirs)
    vr, opacity, depth, rgb, _ws = model.render_func(sigmas, rgbs, deltas, ts, rays_a, T_threshold)
    rgb = rgb + (1 - opacity)[:, None]                                   # white background (:152-156)
    return {'opacity': opacity, 'depth': depth, 'rgb': rgb, 'total_samples': vr}


def __render_rays_train(model, rays_o, rays_d, hits_t, exp_step_factor=0, T_threshold=1e-4):
    """rendering.py:161-228: march -> model -> composite"""
    results = {}
    (rays_a, xyzs, dirs, results['deltas'], results['ts'], results['rm_samples']) = raymarching_train(
        rays_o, rays_d, hits_t, model.occupancy_grid.getBitfield(), model.cascades, model.scale,
        exp_step_factor, model.grid_size, MAX_SAMPLES)

    sigmas, rgbs = model(xyzs, dirs)

    (results['vr_samples'], results['opacity'], results['depth'], results['rgb'], results['ws']) = \
        model.render_func(sigmas, rgbs, results['deltas'], results['ts'], rays_a, T_threshold)
    results['rays_a'] = rays_a

    if exp_step_factor == 0:
        rgb_bg = torch.ones(3, device=rays_o.device)
    else:
        rgb_bg = torch.zeros(3, device=rays_o.device)
    results['rgb'] = results['rgb'] + rgb_bg * (1 - results['opacity'])[:, None]   # :225-226
    return results
