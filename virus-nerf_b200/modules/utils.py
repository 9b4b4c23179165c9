"""Drop-in for the reference's modules/utils.py (constants, level-geometry helpers, Morton
encode/decode, packbits, deployment export) -- Taichi kernels replaced by C-ABI calls."""
import os

import numpy as np
import torch

from .. import _lib

torch_type = torch.float32

# modules/utils.py:12-16
MAX_SAMPLES = 1024
NEAR_DISTANCE = 0.01
SQRT3 = 1.7320508075688772
SQRT3_MAX_SAMPLES = SQRT3 / 1024
SQRT3_2 = 1.7320508075688772 * 2


def res_in_level_np(level_i, base_res, log_per_level_scale):
    """modules/utils.py:19-29"""
    result = np.ceil(float(base_res) * np.exp(float(level_i) * log_per_level_scale) - 1.0)
    return float(result + 1)


def scale_in_level_np(base_res, max_res, levels):
    """modules/utils.py:31-40"""
    return np.log(float(max_res) / float(base_res)) / float(levels - 1)


def align_to(x, y):
    """modules/utils.py:42"""
    return int((x + y - 1) / y) * y


def morton3D_invert(indices):
    """modules/utils.py:130-137: Morton index -> (x, y, z) int32 coords"""
    indices = indices.contiguous().to(torch.int32)
    coords = torch.zeros(indices.size(0), 3, device=indices.device, dtype=torch.int32)
    _lib.call("vn_morton3d_invert", indices, indices.size(0), coords)
    return coords


def morton3D(coords1):
    """modules/utils.py:147-154: (x, y, z) -> Morton index int32"""
    coords1 = coords1.contiguous().to(torch.int32)
    indices = torch.zeros(coords1.size(0), device=coords1.device, dtype=torch.int32)
    _lib.call("vn_morton3d", coords1, coords1.size(0), indices)
    return indices


def packbits(density_grid, density_threshold, density_bitfield):
    """modules/utils.py:157-169: bit i of byte n = density_grid[8n+i] > threshold (in place)"""
    assert density_grid.dtype == torch.float32 and density_bitfield.dtype == torch.uint8
    _lib.call("vn_packbits", density_grid.contiguous(), density_bitfield.numel(), float(density_threshold),
              density_bitfield)


def depth2img(depth):
    """modules/utils.py:223-228"""
    import cv2
    depth = (depth - depth.min()) / (depth.max() - depth.min())
    return cv2.applyColorMap((depth * 255).astype(np.uint8), cv2.COLORMAP_TURBO)


def save_deployment_model(model, dataset, save_dir):
    """modules/utils.py:230-253: same dict keys / layouts as the reference's deployment.npy"""
    padding = torch.zeros(13, 16)
    rgb_out = model.rgb_net.output_layer.weight.detach().cpu()
    rgb_out = torch.cat([rgb_out, padding], dim=0)
    new_dict = {
        'poses': dataset.poses.cpu().numpy(),
        'model.density_bitfield': model.occupancy_grid.getBitfield().cpu().numpy(),
        'model.hash_encoder.params': model.pos_encoder.hash_table.detach().cpu().numpy(),
        'model.per_level_scale': model.pos_encoder.log_b,
        'model.xyz_encoder.params': torch.cat(
            [model.xyz_encoder.hidden_layers[0].weight.detach().cpu().reshape(-1),
             model.xyz_encoder.output_layer.weight.detach().cpu().reshape(-1)]).numpy(),
        'model.rgb_net.params': torch.cat(
            [model.rgb_net.hidden_layers[0].weight.detach().cpu().reshape(-1),
             rgb_out.reshape(-1)]).numpy(),
    }
    np.save(os.path.join(f'{save_dir}', 'deployment.npy'), new_dict)
