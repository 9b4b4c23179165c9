"""Drop-in for the reference's modules/hash_encoder_half.py: fp16 table copy per call,
fp16 output [N, levels*2], f32 gradient buffer ``hash_grad`` with the zero-skip rule."""
import torch

from .. import _lib
from .utils import align_to, res_in_level_np, scale_in_level_np  # noqa: F401  (module-level names of the reference)

torch_type = torch.float16


class _HashEncodeHalf(torch.autograd.Function):
    """hash_encoder_half.py:321-359"""

    @staticmethod
    def forward(ctx, input_pos, params_f32, enc):
        S = input_pos.shape[0]
        # hash_table.to(torch.float16) re-materialised every call (:367)
        table_h = torch.empty(params_f32.shape, dtype=torch.float16, device=params_f32.device)
        _lib.call("vn_f32_to_f16", params_f32, table_h, params_f32.numel())
        out = torch.empty(S, enc.hash_level, enc.feature_per_level, dtype=torch_type, device=input_pos.device)
        _lib.call("vn_hash_encode_fwd_f16", input_pos, table_h, out, S, enc._levels, enc.kernel_flags)
        ctx.save_for_backward(input_pos)
        ctx.enc = enc
        return out

    @staticmethod
    def backward(ctx, doutput):
        (input_pos,) = ctx.saved_tensors
        enc = ctx.enc
        hash_grad = enc.hash_grad.zero_()               # :350-352
        _lib.call("vn_hash_encode_bwd_f16", input_pos, doutput.contiguous().to(torch.float16), hash_grad,
                  input_pos.shape[0], enc._levels, enc.kernel_flags)
        return None, hash_grad, None


class HashEncoder(torch.nn.Module):

    def __init__(self, max_params: float = 2 ** 19, levels: int = 16.0, base_res: float = 16.0,
                 max_res: float = 2048.0, feature_per_level: int = 2):
        super().__init__()
        if feature_per_level != 2:
            raise NotImplementedError("virus-nerf_b200 HashEncoder(half): feature_per_level must be 2")
        levels = int(levels)
        self.log_b = scale_in_level_np(base_res=base_res, max_res=max_res, levels=levels)
        self.base_res = base_res
        self.hash_level = levels
        self.max_params = max_params
        self.feature_per_level = feature_per_level
        self.out_dim = feature_per_level * levels
        self.kernel_flags = 0

        self._levels = _lib.hash_levels(base_res, max_res, levels, int(max_params))
        self.register_buffer('offsets', torch.tensor(list(self._levels.offsets)[:levels], dtype=torch.int32),
                             persistent=False)
        self.register_buffer('hash_map_sizes', torch.tensor(list(self._levels.sizes)[:levels], dtype=torch.int32),
                             persistent=False)
        self.begin_fast_hash_level = int(self._levels.begin_fast_hash_level)
        offset = int(self._levels.total_entries)
        self.total_param_size = offset * feature_per_level

        print(f'Hash Encoder: base_res={base_res} max_res={max_res} hash_level={levels} '
              f'feat_per_level={feature_per_level} per_level_scale={self.log_b} total_hash_size={offset} ')

        self.hash_table = torch.nn.Parameter(torch.zeros(offset, feature_per_level, dtype=torch.float32),
                                             requires_grad=True)
        torch.nn.init.uniform_(self.hash_table, -1e-4, 1e-4)          # :299
        self.register_buffer('hash_grad', torch.zeros_like(self.hash_table, dtype=torch.float32))   # :300-306

    def forward(self, positions):
        return _HashEncodeHalf.apply(positions.contiguous(), self.hash_table.contiguous(), self).view(-1, self.out_dim)
