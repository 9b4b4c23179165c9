"""Drop-in for the reference's modules/volume_train.py (VolumeRenderer)."""
import torch

from .. import _lib
from .utils import torch_type


class _VolumeRender(torch.autograd.Function):
    """volume_train.py:58-175"""

    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, ts, rays_a, T_threshold):
        n_rays = rays_a.shape[0]
        S = sigmas.shape[0]
        dev = rays_a.device
        total_samples = torch.empty(n_rays, dtype=torch.int32, device=dev)
        opacity = torch.empty(n_rays, dtype=torch_type, device=dev)
        depth = torch.empty(n_rays, dtype=torch_type, device=dev)
        rgb = torch.empty(n_rays, 3, dtype=torch_type, device=dev)
        ws = torch.empty(S, dtype=torch_type, device=dev)
        _lib.call("vn_composite_train_fwd", sigmas, rgbs, deltas, ts, rays_a, n_rays, S, float(T_threshold),
                  total_samples, opacity, depth, rgb, ws)
        ctx.save_for_backward(sigmas, rgbs, deltas, ts, rays_a)
        ctx.T_threshold = float(T_threshold)
        vr = total_samples.sum()
        ctx.mark_non_differentiable(vr)
        return vr, opacity, depth, rgb, ws

    @staticmethod
    def backward(ctx, dL_dtotal_samples, dL_dopacity, dL_ddepth, dL_drgb, dL_dws):
        sigmas, rgbs, deltas, ts, rays_a = ctx.saved_tensors
        n_rays, S = rays_a.shape[0], sigmas.shape[0]
        dsig = torch.empty_like(sigmas)
        drgb = torch.empty_like(rgbs)

        def _c(g, like_shape):
            if g is None:
                return torch.zeros(like_shape, dtype=torch_type, device=sigmas.device)
            return g.contiguous().float()

        _lib.call("vn_composite_train_bwd", sigmas, rgbs, deltas, ts, rays_a, n_rays, S, ctx.T_threshold,
                  _c(dL_dopacity, (n_rays,)), _c(dL_ddepth, (n_rays,)), _c(dL_drgb, (n_rays, 3)),
                  None if dL_dws is None else dL_dws.contiguous().float(), dsig, drgb)
        return dsig, drgb, None, None, None, None


class VolumeRenderer(torch.nn.Module):

    def __init__(self):
        super().__init__()

    def forward(self, sigmas, rgbs, deltas, ts, rays_a, T_threshold):
        """volume_train.py:179-195 -> (vr_samples, opacity, depth, rgb, ws)"""
        return _VolumeRender.apply(sigmas.contiguous().float(), rgbs.contiguous().float(), deltas.contiguous(),
                                   ts.contiguous(), rays_a.contiguous(), T_threshold)
