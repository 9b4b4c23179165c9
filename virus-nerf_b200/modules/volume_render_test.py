"""Drop-in for the reference's modules/volume_render_test.py (composite_test, in place)."""
from .. import _lib


def composite_test(sigmas, rgbs, deltas, ts, pack_info, alive_indices, T_threshold, opacity, depth, rgb):
    """volume_render_test.py:4-54: accumulates into opacity/depth/rgb and sets finished rays'
    alive_indices entries to -1."""
    _lib.call("vn_composite_test", sigmas.contiguous().float(), rgbs.contiguous().float(), deltas.contiguous(),
              ts.contiguous(), pack_info.contiguous(), alive_indices, alive_indices.size(0), float(T_threshold),
              opacity, depth, rgb)
