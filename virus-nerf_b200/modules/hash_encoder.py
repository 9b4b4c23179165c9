"""Drop-in for the reference's modules/hash_encoder.py: fp32 multi-resolution hash grid.
Same constructor, attributes (hash_table, offsets, hash_map_sizes, log_b, out_dim,
begin_fast_hash_level, total_param_size) and state-dict keys; forward / backward are the
sm_100a kernels vn_hash_encode_fwd_f32 / vn_hash_encode_bwd_f32."""
import torch

from .. import _lib
from .utils import scale_in_level_np, torch_type


class _HashEncode(torch.autograd.Function):
    """hash_encoder.py:237-277 (_module_function)"""

    @staticmethod
    def forward(ctx, input_pos, params, enc):
        S = input_pos.shape[0]
        out = torch.empty(S, enc.out_dim, dtype=torch_type, device=input_pos.device)
        _lib.call("vn_hash_encode_fwd_f32", input_pos, params, out, S, enc._levels, enc.kernel_flags)
        ctx.save_for_backward(input_pos)
        ctx.enc = enc
        ctx.n_params = params.numel()
        return out

    @staticmethod
    def backward(ctx, doutput):
        (input_pos,) = ctx.saved_tensors
        enc = ctx.enc
        direct = getattr(enc, "_direct_grad", None)
        if direct is not None:
            # engine mode: accumulate straight into the (already zeroed) flat gradient buffer
            _lib.call("vn_hash_encode_bwd_f32", input_pos, doutput.contiguous().float(), direct,
                      input_pos.shape[0], enc._levels, enc.kernel_flags)
            return None, None, None
        grad = torch.zeros(ctx.n_params, dtype=torch_type, device=input_pos.device)
        _lib.call("vn_hash_encode_bwd_f32", input_pos, doutput.contiguous().float(), grad, input_pos.shape[0],
                  enc._levels, enc.kernel_flags)
        # the reference returns (None, params.grad): no gradient w.r.t. positions (:277)
        return None, grad, None


class HashEncoder(torch.nn.Module):

    def __init__(self, max_params: float = 2 ** 19, levels: int = 16.0, base_res: float = 16.0,
                 max_res: float = 2048.0, feature_per_level: int = 2):
        super().__init__()
        if feature_per_level != 2:
            raise NotImplementedError("virus-nerf_b200 HashEncoder: feature_per_level must be 2 "
                                      "(every shipped reference config uses 2)")
        levels = int(levels)
        self.log_b = scale_in_level_np(base_res=base_res, max_res=max_res, levels=levels)  # :160-164
        self.base_res = base_res
        self.hash_level = levels
        self.max_params = max_params
        self.feature_per_level = feature_per_level
        self.out_dim = feature_per_level * levels
        self.kernel_flags = 0

        # a1: level geometry from the C ABI (hash_encoder.py:183-208)
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

        self.hash_table = torch.nn.Parameter(torch.zeros(self.total_param_size, dtype=torch.float32),
                                             requires_grad=True)
        torch.nn.init.uniform_(self.hash_table)   # :227, U(0,1)

    def forward(self, positions):
        return _HashEncode.apply(positions.contiguous(), self.hash_table.contiguous(), self)
