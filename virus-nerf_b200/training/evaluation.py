"""Evaluation consumers of the test-time path, mirroring the host
drivers of the reference that feed the hot path with its real evaluation inputs:

  * createScanRays / createScanPos            helpers/geometric_fcts.py:77-150
  * batchify_render / batchify_density        training/trainer_base.py:183-254 (_batchifyRender/_batchifyDensity)
  * interfere_density_map                     training/trainer_base.py:92-140 (interfereDensityMap)
  * evaluation_depth_nerf                     training/trainer.py:574-629 (_getEvaluationDataNeRF): 2D lidar-like
                                              scans (d_z = 0) rendered with raymarching_test + composite_test
  * save_checkpoint / load_checkpoint         training/trainer_base.py:142-181: torch state dict with the
                                              reference's keys (.pth files interchange with the reference)

Host logic only (numpy / torch glue, identical call structure); every density / render call runs on
the sm_100a kernels through modules/.  Scene conversion (c2w / w2c) is the caller's: these functions
work in cube coordinates.
"""
import os

import numpy as np
import torch

from ..modules.rendering import render


def createScanRays(rays_o, angle_res: int, angle_min_max: tuple = (-np.pi, np.pi)):
    """helpers/geometric_fcts.py:77-111: M scan directions with d_z = 0 per origin -> (N*M, 3), (N*M, 3)"""
    is_tensor = isinstance(rays_o, torch.Tensor)
    if is_tensor:
        device = rays_o.device
        rays_o = rays_o.detach().clone().cpu().numpy()
    rays_d = np.zeros((angle_res, 3))
    angles = np.linspace(angle_min_max[0], angle_min_max[1], angle_res, endpoint=False)
    rays_d[:, 0] = np.cos(angles)
    rays_d[:, 1] = np.sin(angles)
    rays_d = np.tile(rays_d, (rays_o.shape[0], 1))
    rays_o = np.repeat(rays_o, angle_res, axis=0)
    if is_tensor:
        rays_o = torch.tensor(rays_o, dtype=torch.float32, device=device)
        rays_d = torch.tensor(rays_d, dtype=torch.float32, device=device)
    return rays_o, rays_d


def createScanPos(res_map: int, height_c: float, num_avg_heights: int, tolerance_c: float, cube_min: float,
                  cube_max: float, device):
    """helpers/geometric_fcts.py:113-150: (L*L*A, 3) slice positions at A heights around height_c"""
    pos = torch.linspace(cube_min, cube_max, res_map).to(device)
    m1, m2 = torch.meshgrid(pos, pos, indexing="ij")
    pos = torch.stack((m1.reshape(-1), m2.reshape(-1)), dim=1)
    pos_avg = torch.zeros(res_map * res_map, num_avg_heights, 3).to(device)
    for i, h in enumerate(np.linspace(height_c - tolerance_c, height_c + tolerance_c, num_avg_heights)):
        pos_avg[:, i, :2] = pos
        pos_avg[:, i, 2] = h
    return pos_avg.reshape(-1, 3)


def batchify_render(model, rays_o, rays_d, test_time: bool, batch_size: int, exp_step_factor: float = 0.0):
    """trainer_base.py:183-221"""
    N = rays_o.shape[0]
    ctx = torch.no_grad() if test_time else torch.enable_grad()
    with ctx:
        for s in range(0, N, batch_size):
            e = min(s + batch_size, N)
            yield render(model, rays_o=rays_o[s:e], rays_d=rays_d[s:e], test_time=test_time,
                         exp_step_factor=exp_step_factor)


def batchify_density(model, pos, test_time: bool, batch_size: int):
    """trainer_base.py:223-254"""
    N = pos.shape[0]
    ctx = torch.no_grad() if test_time else torch.enable_grad()
    with ctx:
        for s in range(0, N, batch_size):
            yield model.density(pos[s:min(s + batch_size, N)].contiguous())


def interfere_density_map(model, res_map: int, height_c: float, num_avg_heights: int, tolerance_c: float,
                          threshold: float, cube_min: float, cube_max: float, batch_size: int = 8192):
    """trainer_base.py:92-140 in cube coordinates -> (density_map (L, L), thresholded map (L, L))"""
    dev = next(model.parameters()).device
    pos_avg = createScanPos(res_map, height_c, num_avg_heights, tolerance_c, cube_min, cube_max, dev)
    parts = [d.to(torch.float32) for d in batchify_density(model, pos_avg, test_time=True, batch_size=batch_size)]
    density_map = torch.cat(parts, dim=0).detach().cpu().numpy().reshape(-1, num_avg_heights)
    density_map = np.nanmax(density_map, axis=1).reshape(res_map, res_map)
    density_map_thr = np.zeros_like(density_map)
    density_map_thr[density_map >= threshold] = 1.0
    return density_map, density_map_thr


def evaluation_depth_nerf(model, scan_origins, angle_res: int, batch_size: int = 8192, exp_step_factor: float = 0.0):
    """trainer.py:574-629: depth of `angle_res` horizontal scan rays per origin (cube coordinates)
    -> rays_o (N*M, 3), rays_d (N*M, 3), depths (N*M,) as numpy"""
    rays_o, rays_d = createScanRays(scan_origins, angle_res)
    depths = [r['depth'] for r in batchify_render(model, rays_o.contiguous(), rays_d.contiguous(), test_time=True,
                                                  batch_size=batch_size, exp_step_factor=exp_step_factor)]
    depths = torch.cat(depths, dim=0) if depths else torch.empty(0, device=rays_o.device)
    return rays_o.cpu().numpy(), rays_d.cpu().numpy(), depths.detach().cpu().numpy()


def save_checkpoint(model, save_dir: str, name: str = "model.pth"):
    """trainer_base.py:155-181: plain state dict, the reference's keys"""
    os.makedirs(save_dir, exist_ok=True)
    path = os.path.join(save_dir, name)
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, path)
    return path


def load_checkpoint(model, ckpt_path: str):
    """trainer_base.py:142-153"""
    state_dict = torch.load(ckpt_path, map_location=torch.device('cpu'))
    model.load_state_dict(state_dict)
    return model
