"""Drop-in for the reference's modules/intersection.py."""
import torch

from .. import _lib
from .utils import NEAR_DISTANCE  # noqa: F401  (re-exported like the reference)


def ray_aabb_intersection(rays_o, rays_d, scale):
    """intersection.py:40-55 -> hits_t [N,2] (t1 clamped to NEAR_DISTANCE; (-1,-1) on a miss)"""
    hits_t = torch.empty(rays_o.size(0), 2, device=rays_o.device, dtype=rays_o.dtype)
    _lib.call("vn_ray_aabb", rays_o.contiguous(), rays_d.contiguous(), float(scale), rays_o.size(0), hits_t)
    return hits_t
