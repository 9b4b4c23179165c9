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
