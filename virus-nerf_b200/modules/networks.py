"""Drop-in for the reference's modules/networks.py: NGP model (hash encoder -> 32-64-16 MLP
-> TruncExp; SH(16)+16 -> 64-64-3 sigmoid MLP, all bias-free), same sub-module names and
state-dict keys (pos_encoder.hash_table, xyz_encoder.hidden_layers.0.weight,
xyz_encoder.output_layer.weight, rgb_net.hidden_layers.{0,1}.weight, rgb_net.output_layer.weight,
buffers center / xyz_min / xyz_max / half_size)."""
from typing import Callable, Optional

import numpy as np
import torch
from torch import nn

from .. import _lib
from .spherical_harmonics import DirEncoder
from .volume_train import VolumeRenderer


class _FusedMLP(torch.autograd.Function):
    """density net + SH + colour net as one tcgen05 kernel (vn_mlp_fwd / vn_mlp_bwd):
    networks.py:142-145 and :160-162.  fp16 operands / fp32 accumulation, i.e. what the
    reference computes under torch.autocast(float16) on CUDA (trainer.py:104)."""

    @staticmethod
    def forward(ctx, enc, dirs, W1, W2, W3, W4, W5):
        S = enc.shape[0]
        sig = torch.empty(S, dtype=torch.float32, device=enc.device)
        rgb = torch.empty(S, 3, dtype=torch.float32, device=enc.device)
        half = 1 if enc.dtype == torch.float16 else 0
        Ws = [w.detach().float().contiguous() for w in (W1, W2, W3, W4, W5)]
        _lib.call("vn_mlp_fwd", enc, half, dirs, *Ws, S, 0, sig, rgb, None)
        ctx.save_for_backward(enc, dirs, *Ws)
        ctx.half = half
        return sig, rgb

    @staticmethod
    def backward(ctx, dsig, drgb):
        enc, dirs, *Ws = ctx.saved_tensors
        S = enc.shape[0]
        denc = torch.empty(S, 32, dtype=torch.float32, device=enc.device)
        dWs = [torch.zeros_like(w) for w in Ws]
        dsig = torch.zeros(S, device=enc.device) if dsig is None else dsig.contiguous().float()
        drgb = torch.zeros(S, 3, device=enc.device) if drgb is None else drgb.contiguous().float()
        _lib.call("vn_mlp_bwd", enc, ctx.half, dirs, *Ws, S, 0, dsig, drgb, denc, *dWs)
        return (denc.to(enc.dtype), None, *dWs)


class _FusedDensity(torch.autograd.Function):
    """density net only (NGP.density): vn_mlp_fwd / vn_mlp_bwd with density_only = 1"""

    @staticmethod
    def forward(ctx, enc, W1, W2, want_h):
        S = enc.shape[0]
        sig = torch.empty(S, dtype=torch.float32, device=enc.device)
        h = torch.empty(S, 16, dtype=torch.float32, device=enc.device) if want_h else None
        half = 1 if enc.dtype == torch.float16 else 0
        W1c, W2c = W1.detach().float().contiguous(), W2.detach().float().contiguous()
        _lib.call("vn_mlp_fwd", enc, half, None, W1c, W2c, None, None, None, S, 1, sig, None, h)
        ctx.save_for_backward(enc, W1c, W2c)
        ctx.half = half
        if want_h:
            ctx.mark_non_differentiable(h)
            return sig, h
        return sig

    @staticmethod
    def backward(ctx, dsig, *_):
        enc, W1, W2 = ctx.saved_tensors
        S = enc.shape[0]
        denc = torch.empty(S, 32, dtype=torch.float32, device=enc.device)
        dW1, dW2 = torch.zeros_like(W1), torch.zeros_like(W2)
        _lib.call("vn_mlp_bwd", enc, ctx.half, None, W1, W2, None, None, None, S, 1, dsig.contiguous().float(), None,
                  denc, dW1, dW2, None, None, None)
        return denc.to(enc.dtype), dW1, dW2, None


class TruncExp(torch.autograd.Function):
    """networks.py:17-29 (forward exp in fp32; backward dL * exp(clamp(x, -15, 15)))"""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dL_dout):
        x = ctx.saved_tensors[0]
        return dL_dout * torch.exp(x.clamp(-15, 15))


class MLP(nn.Module):
    """networks.py:195-282 (bias-free ReLU MLP; skip connections as in the reference)"""

    def __init__(self, input_dim: int, output_dim: int = None, net_depth: int = 8, net_width: int = 256,
                 skip_layer: int = 4, hidden_init: Callable = nn.init.xavier_uniform_,
                 hidden_activation: Callable = nn.ReLU(), output_enabled: bool = True,
                 output_init: Optional[Callable] = nn.init.xavier_uniform_,
                 output_activation: Optional[Callable] = nn.Identity(), bias_enabled: bool = True,
                 bias_init: Callable = nn.init.zeros_):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.net_depth = net_depth
        self.net_width = net_width
        self.skip_layer = skip_layer
        self.hidden_init = hidden_init
        self.hidden_activation = hidden_activation
        self.output_enabled = output_enabled
        self.output_init = output_init
        self.output_activation = output_activation
        self.bias_enabled = bias_enabled
        self.bias_init = bias_init

        self.hidden_layers = nn.ModuleList()
        in_features = self.input_dim
        for i in range(self.net_depth):
            self.hidden_layers.append(nn.Linear(in_features, self.net_width, bias=bias_enabled))
            if (self.skip_layer is not None) and (i % self.skip_layer == 0) and (i > 0):
                in_features = self.net_width + self.input_dim
            else:
                in_features = self.net_width
        if self.output_enabled:
            self.output_layer = nn.Linear(in_features, self.output_dim, bias=bias_enabled)
        else:
            self.output_dim = in_features
        self.initialize()

    def initialize(self):
        for m in list(self.hidden_layers) + ([self.output_layer] if self.output_enabled else []):
            init = self.hidden_init if m is not getattr(self, "output_layer", None) else self.output_init
            if init is not None:
                init(m.weight)
            if self.bias_enabled and self.bias_init is not None:
                self.bias_init(m.bias)

    def forward(self, x):
        inputs = x
        for i in range(self.net_depth):
            x = self.hidden_layers[i](x)
            x = self.hidden_activation(x)
            if (self.skip_layer is not None) and (i % self.skip_layer == 0) and (i > 0):
                x = torch.cat([x, inputs], dim=-1)
        if self.output_enabled:
            x = self.output_layer(x)
            x = self.output_activation(x)
        return x


class NGP(nn.Module):

    def __init__(self, scale: float = 0.5, pos_encoder_type: str = 'hash', levels: int = 16,
                 feature_per_level: int = 2, log2_T: int = 19, base_res: int = 16, max_res: int = 2048,
                 half_opt: bool = False, xyz_net_width: int = 64, xyz_net_depth: int = 1,
                 xyz_net_out_dim: int = 16, rgb_net_depth: int = 2, rgb_net_width: int = 64,
                 scene=None, dataset=None, args=None):
        super().__init__()

        # scene bounding box, networks.py:57-62
        self.scale = scale
        self.register_buffer('center', torch.zeros(1, 3))
        self.register_buffer('xyz_min', -torch.ones(1, 3) * scale)
        self.register_buffer('xyz_max', torch.ones(1, 3) * scale)
        self.register_buffer('half_size', (self.xyz_max - self.xyz_min) / 2)

        self.cascades = max(1 + int(np.ceil(np.log2(2 * scale))), 1)   # :65
        self.grid_size = 128                                            # :66
        self.half_opt = half_opt
        # fused tcgen05 MLP (fp16 operands, as the reference under autocast on CUDA).  Set to False
        # to run the layers as fp32 torch.nn.Linear (the reference's CUDA-less behaviour; used by
        # the fp32 parity tests).
        self.fused_mlp = (xyz_net_width, xyz_net_depth, xyz_net_out_dim, rgb_net_depth, rgb_net_width) == (64, 1, 16, 2, 64)

        if pos_encoder_type == 'hash':
            if half_opt:
                from .hash_encoder_half import HashEncoder
            else:
                from .hash_encoder import HashEncoder
            self.pos_encoder = HashEncoder(max_params=2 ** log2_T, base_res=base_res, max_res=max_res,
                                           levels=levels, feature_per_level=feature_per_level)
        else:
            raise NotImplementedError(f"pos_encoder_type {pos_encoder_type} not implemented "
                                      "(triplane is outside the hot path, SURVEY section 2.1 row 14)")

        self.xyz_encoder = MLP(input_dim=self.pos_encoder.out_dim, output_dim=xyz_net_out_dim,
                               net_depth=xyz_net_depth, net_width=xyz_net_width, bias_enabled=False)
        self.dir_encoder = DirEncoder()
        rgb_input_dim = self.dir_encoder.out_dim + self.xyz_encoder.output_dim
        self.rgb_net = MLP(input_dim=rgb_input_dim, output_dim=3, net_depth=rgb_net_depth, net_width=rgb_net_width,
                           bias_enabled=False, output_activation=nn.Sigmoid())
        self.render_func = VolumeRenderer()

        self.args = args
        if args is not None and self.args.model.grid_type == 'occ':
            from .occupancy_grid import OccupancyGrid
            self.occupancy_grid = OccupancyGrid(args=args, grid_size=self.grid_size, scene=scene, dataset=dataset,
                                                fct_density=self.density)
        elif args is not None and self.args.model.grid_type == 'ngp':                  # networks.py:117-122
            from .ngp_grid import NGPGrid
            self.occupancy_grid = NGPGrid(args=args, grid_size=self.grid_size, fct_density=self.density)
        elif args is not None:
            raise NotImplementedError(f"grid_type {self.args.model.grid_type} not implemented")

    def density(self, x, return_feat=False):
        """networks.py:134-148"""
        x = (x - self.xyz_min) / (self.xyz_max - self.xyz_min)
        embedding = self.pos_encoder(x)
        if self._use_fused(embedding):
            return _FusedDensity.apply(embedding.contiguous(), self.xyz_encoder.hidden_layers[0].weight,
                                       self.xyz_encoder.output_layer.weight, return_feat)
        h = self.xyz_encoder(embedding)
        sigmas = TruncExp.apply(h[:, 0])
        if return_feat:
            return sigmas, h
        return sigmas

    def _use_fused(self, t):
        return self.fused_mlp and t.is_cuda and self.pos_encoder.out_dim == 32

    def forward(self, x, d):
        """networks.py:150-164"""
        if self._use_fused(x):
            xn = (x - self.xyz_min) / (self.xyz_max - self.xyz_min)
            embedding = self.pos_encoder(xn)
            return _FusedMLP.apply(embedding.contiguous(), d.contiguous().float(),
                                   self.xyz_encoder.hidden_layers[0].weight, self.xyz_encoder.output_layer.weight,
                                   self.rgb_net.hidden_layers[0].weight, self.rgb_net.hidden_layers[1].weight,
                                   self.rgb_net.output_layer.weight)
        sigmas, h = self.density(x, return_feat=True)
        d = d / torch.norm(d, dim=1, keepdim=True)
        d = self.dir_encoder((d + 1) / 2)
        rgbs = self.rgb_net(torch.cat([d.to(h.dtype), h], 1))
        return sigmas, rgbs

    @torch.no_grad()
    def updateNeRFGrid(self, density_threshold, warmup=False, decay=0.95, erode=False):
        """networks.py:166-178"""
        self.occupancy_grid.update(density_threshold=density_threshold, warmup=warmup, decay=decay, erode=erode)

    def updateOccGrid(self, density_threshold: float, elapse_time: float):
        """networks.py:180-191"""
        self.occupancy_grid.update(elapse_time=elapse_time)
