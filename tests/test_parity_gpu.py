"""GPU parity tests: every C-ABI kernel of libvirusnerf_sm100.so against the CPU oracle on the
same seeded inputs.  Bars (BASELINE.json north_star): hash / Morton indices, bitfields,
per-ray sample counts and (t, dt) sequences bit-exact; encoder and composite forward rtol
1e-5; gradients rtol 1e-4 (+ atol scaled to the largest reference entry: float atomics
reorder); half encoder fwd rtol 2e-3 / atol 1e-3 vs the fp16 oracle, grads rtol 1e-2."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def N(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def vn():
    from virus_nerf_b200 import _lib
    _lib.lib()
    return _lib


@pytest.fixture(scope="module")
def scene_rays():
    from virus_nerf_b200 import synthetic
    sc = synthetic.RoomScene()
    ds = synthetic.SyntheticDataset(sc, pool_size=1 << 14, n_images=16, device="cpu")
    b = ds(3000, {"pixs": {"valid_uss": 0.4, "valid_tof": 0.4}})
    ro, rd = b["rays_o"].numpy().copy(), b["rays_d"].numpy().copy()
    # edge cases: axis aligned, zero components (scan rays), rays from outside, missing rays
    so, sd = synthetic.scan_rays(64)
    extra_o = np.array([[0, 0, 0], [0, 0, 0], [2, 2, 2], [2, 0, 0], [0.2, 0.1, 0.0], [-0.5, 0.0, 0.0]], np.float32)
    extra_d = np.array([[1, 0, 0], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0.6, 0.8, 0.0], [1, 0, 0]], np.float32)
    ro = np.concatenate([ro, so, extra_o]); rd = np.concatenate([rd, sd, extra_d])
    bitfields = {
        "carved": synthetic.morton_pack(sc.occupancy_bitfield(128)),
        "full": np.full(128 ** 3 // 8, 255, np.uint8),
        "empty": np.zeros(128 ** 3 // 8, np.uint8),
        "random": np.random.default_rng(3).integers(0, 256, 128 ** 3 // 8).astype(np.uint8),
    }
    return ro, rd, bitfields


def ray_coherent_points(n_rays=256, per_ray=64, seed=0):
    """consecutive samples along rays in [0,1]^3 (what the encoder sees in training)"""
    rng = np.random.default_rng(seed)
    o = rng.random((n_rays, 1, 3)).astype(np.float32) * 0.6 + 0.2
    d = rng.normal(size=(n_rays, 1, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    t = (np.arange(per_ray, dtype=np.float32) * np.float32(np.sqrt(3) / 1024))[None, :, None]
    return np.clip(o + d * t, 0.0, 1.0).reshape(-1, 3).astype(np.float32)


# ------------------------------------------------------------------------------------ a1/a2/a3
@pytest.mark.parametrize("log2_T,max_res", [(19, 1024), (22, 1024), (19, 2048)])
def test_hash_levels_and_indices_bit_exact(vn, oracle_mod, log2_T, max_res):
    lv_o = oracle_mod.HashLevels(16, max_res, 16, 2 ** log2_T)
    lv = vn.hash_levels(16, max_res, 16, 2 ** log2_T)
    assert lv.total_entries == lv_o.total and lv.begin_fast_hash_level == lv_o.begin_fast_hash_level
    np.testing.assert_array_equal(np.array(lv.offsets[:16]), lv_o.offsets)
    np.testing.assert_array_equal(np.array(lv.sizes[:16]), lv_o.sizes)
    np.testing.assert_array_equal(np.array(lv.scales[:16], np.float32), lv_o.scales)
    np.testing.assert_array_equal(np.array(lv.res[:16], np.uint32), lv_o.res)
    rng = np.random.default_rng(1)
    xyz = rng.random((4096, 3)).astype(np.float32)
    edge = np.array([[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1 - 2 ** -24] * 3, [1, 0, 0.5], [0.25, 0.75, 1.0]], np.float32)
    xyz = np.concatenate([xyz, edge, ray_coherent_points(8, 32)])
    S = xyz.shape[0]
    idx = torch.zeros(S, 16, 8, dtype=torch.int32, device=DEV)
    w = torch.zeros(S, 16, 8, dtype=torch.float32, device=DEV)
    vn.call("vn_hash_indices", T(xyz), S, lv, idx, w)
    idx_o, w_o = oracle_mod.hash_indices(xyz, lv_o, with_weights=True)
    np.testing.assert_array_equal(N(idx), idx_o)
    np.testing.assert_array_equal(N(w), w_o)


@pytest.mark.parametrize("flags", [0, 16, 32, 64, 128, 256])
@pytest.mark.parametrize("log2_T", [19, 22])
def test_hash_fwd_f32(vn, oracle_mod, flags, log2_T):
    lv_o = oracle_mod.HashLevels(16, 1024, 16, 2 ** log2_T)
    lv = vn.hash_levels(16, 1024, 16, 2 ** log2_T)
    rng = np.random.default_rng(2)
    table = rng.random(2 * lv_o.total, dtype=np.float32)
    xyz = np.concatenate([rng.random((5000, 3)).astype(np.float32), ray_coherent_points(64, 48),
                          np.array([[0, 0, 0], [1, 1, 1]], np.float32)])
    S = xyz.shape[0]
    out = torch.empty(S, 32, device=DEV)
    vn.call("vn_hash_encode_fwd_f32", T(xyz), T(table), out, S, lv, flags)
    ref = oracle_mod.hash_fwd_f32(xyz, table, lv_o)
    np.testing.assert_allclose(N(out), ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("flags", [0, 1, 16, 32, 33, 64, 128, 129, 256])
def test_hash_bwd_f32(vn, oracle_mod, flags):
    lv_o = oracle_mod.HashLevels(16, 1024, 16, 2 ** 19)
    lv = vn.hash_levels(16, 1024, 16, 2 ** 19)
    rng = np.random.default_rng(4)
    xyz = np.concatenate([rng.random((3001, 3)).astype(np.float32), ray_coherent_points(128, 64)])
    S = xyz.shape[0]
    dout = rng.normal(size=(S, 32)).astype(np.float32)
    grad = torch.zeros(2 * lv_o.total, device=DEV)
    vn.call("vn_hash_encode_bwd_f32", T(xyz), T(dout), grad, S, lv, flags)
    ref = oracle_mod.hash_bwd_f32(xyz, dout, lv_o)
    np.testing.assert_allclose(N(grad), ref, rtol=1e-4, atol=1e-6 * np.abs(ref).max())


def _to_planes(rows):
    """[S, 32] rows -> [8][S] float4 level-pair planes (VN_HASH_PLANAR)"""
    S = rows.shape[0]
    return np.ascontiguousarray(rows.reshape(S, 8, 4).transpose(1, 0, 2))


@pytest.mark.parametrize("flags", [0, 32, 64, 128, 256])
@pytest.mark.parametrize("S_extra", [0, 1, 31])
def test_hash_planar_layout(vn, oracle_mod, flags, S_extra):
    """the level-pair-plane layout of the fast step holds exactly the values of the row layout
    (forward: bit-identical to the row kernel; backward: same gradient as the oracle), and the
    48-register backward variant agrees with it"""
    lv_o = oracle_mod.HashLevels(16, 1024, 16, 2 ** 19)
    lv = vn.hash_levels(16, 1024, 16, 2 ** 19)
    rng = np.random.default_rng(11)
    table = rng.random(2 * lv_o.total, dtype=np.float32)
    xyz = np.concatenate([rng.random((1000 + S_extra, 3)).astype(np.float32), ray_coherent_points(96, 64)])
    S = xyz.shape[0]
    rows = torch.empty(S, 32, device=DEV)
    planes = torch.empty(8, S, 4, device=DEV)
    vn.call("vn_hash_encode_fwd_f32", T(xyz), T(table), rows, S, lv, flags)
    vn.call("vn_hash_encode_fwd_f32", T(xyz), T(table), planes, S, lv, flags | vn.VN_HASH_PLANAR)
    np.testing.assert_array_equal(N(planes), _to_planes(N(rows)))
    if flags in (0, 32, 256):
        vn.call("vn_hash_encode_fwd_f32", T(xyz), T(table), planes, S, lv, flags | vn.VN_HASH_PLANAR | vn.VN_HASH_PAIR_LOADS)
        np.testing.assert_array_equal(N(planes), _to_planes(N(rows)))
    np.testing.assert_allclose(N(rows), oracle_mod.hash_fwd_f32(xyz, table, lv_o), rtol=1e-5, atol=1e-6)
    dout = rng.normal(size=(S, 32)).astype(np.float32)
    dout[rng.random(S) < 0.4] = 0.0                      # samples behind an opaque surface: exactly zero gradient
    dout[::5, 4:8] = 0.0
    ref = oracle_mod.hash_bwd_f32(xyz, dout, lv_o)
    for extra in (0, vn.VN_HASH_TIGHT_REGS, vn.VN_HASH_TIGHT_REGS | vn.VN_HASH_SKIP_ZERO_GRADS, vn.VN_HASH_SKIP_ZERO_GRADS):
        grad = torch.zeros(2 * lv_o.total, device=DEV)
        vn.call("vn_hash_encode_bwd_f32", T(xyz), T(_to_planes(dout)), grad, S, lv, flags | vn.VN_HASH_PLANAR | extra)
        np.testing.assert_allclose(N(grad), ref, rtol=1e-4, atol=1e-6 * np.abs(ref).max())


@pytest.mark.parametrize("flags", [0, 64, 128, 2048])
@pytest.mark.parametrize("S_extra", [0, 1, 31])
def test_hash_f16_chunk_layout(vn, oracle_mod, flags, S_extra):
    """VN_HASH_F16_CHUNKS: plane c of the output holds levels 4c..4c+3 of every point as fp16 -- exactly the
    row kernel's fp32 values rounded to nearest fp16 (where autocast rounds the Linear input)"""
    lv = vn.hash_levels(16, 1024, 16, 2 ** 19)
    rng = np.random.default_rng(12)
    table = rng.random(2 * lv.total_entries, dtype=np.float32) * 2 - 1
    xyz = np.concatenate([rng.random((1000 + S_extra, 3)).astype(np.float32), ray_coherent_points(96, 64)])
    S = xyz.shape[0]
    rows = torch.empty(S, 32, device=DEV)
    chunks = torch.empty(4, S, 8, device=DEV, dtype=torch.float16)
    vn.call("vn_hash_encode_fwd_f32", T(xyz), T(table), rows, S, lv, 0)
    vn.call("vn_hash_encode_fwd_f32", T(xyz), T(table), chunks, S, lv, flags | vn.VN_HASH_F16_CHUNKS)
    np.testing.assert_array_equal(N(chunks), N(rows.half().view(S, 4, 8).permute(1, 0, 2)))
    # half-precision table (hash_encoder_half.py): same layout, the half kernel's own arithmetic
    table_h = T(table).half().view(-1, 2)
    rows_h = torch.empty(S, 16, 2, dtype=torch.float16, device=DEV)
    vn.call("vn_hash_encode_fwd_f16", T(xyz), table_h, rows_h, S, lv, 0)
    vn.call("vn_hash_encode_fwd_f16", T(xyz), table_h, chunks, S, lv, (flags & ~2048) | vn.VN_HASH_F16_CHUNKS)
    np.testing.assert_array_equal(N(chunks), N(rows_h.view(S, 4, 8).permute(1, 0, 2)))


def test_hash_planar_rejects_odd_groups(vn):
    lv = vn.hash_levels(16, 1024, 16, 2 ** 19)
    x = torch.rand(64, 3, device=DEV); t = torch.rand(2 * lv.total_entries, device=DEV); o = torch.empty(8, 64, 4, device=DEV)
    with pytest.raises(RuntimeError, match="planar"):
        vn.call("vn_hash_encode_fwd_f32", x, t, o, 64, lv, vn.VN_HASH_PLANAR | vn.VN_HASH_LEVEL_GROUPS_1)
    with pytest.raises(RuntimeError, match="planar"):
        vn.call("vn_hash_encode_bwd_f32", x, o, t, 64, lv, vn.VN_HASH_PLANAR | vn.VN_HASH_NO_WARP_AGG)


def test_hash_bwd_level_ranges(vn, oracle_mod):
    """the per-level-group launches used for the overlapped DP allreduce add up to the full backward"""
    lv_o = oracle_mod.HashLevels(16, 1024, 16, 2 ** 19)
    lv = vn.hash_levels(16, 1024, 16, 2 ** 19)
    xyz = ray_coherent_points(64, 64)
    S = xyz.shape[0]
    dout = np.random.default_rng(6).normal(size=(S, 32)).astype(np.float32)
    grad = torch.zeros(2 * lv_o.total, device=DEV)
    for lb, le in ((12, 16), (8, 12), (4, 8), (0, 4)):
        vn.call("vn_hash_encode_bwd_f32_levels", T(xyz), T(dout), grad, S, lv, 0, lb, le)
        part = N(grad)
        lo, hi = 2 * lv_o.offsets[lb], 2 * (lv_o.offsets[le] if le < 16 else lv_o.total)
        assert np.abs(part[lo:hi]).max() > 0
    ref = oracle_mod.hash_bwd_f32(xyz, dout, lv_o)
    np.testing.assert_allclose(N(grad), ref, rtol=1e-4, atol=1e-6 * np.abs(ref).max())
    with pytest.raises(RuntimeError, match="level range"):
        vn.call("vn_hash_encode_bwd_f32_levels", T(xyz), T(dout), grad, S, lv, 0, 8, 4)


def test_hash_module_autograd(vn, oracle_mod):
    from virus_nerf_b200.modules.hash_encoder import HashEncoder
    enc = HashEncoder(max_params=2 ** 19, levels=16, base_res=16, max_res=1024).to(DEV)
    assert enc.out_dim == 32 and enc.total_param_size == 11420064 and enc.begin_fast_hash_level == 6
    xyz = torch.rand(1000, 3, device=DEV)
    out = enc(xyz)
    g = torch.randn_like(out)
    out.backward(g)
    lv_o = oracle_mod.HashLevels(16, 1024, 16, 2 ** 19)
    np.testing.assert_allclose(N(out), oracle_mod.hash_fwd_f32(N(xyz), N(enc.hash_table), lv_o), rtol=1e-5, atol=1e-6)
    ref = oracle_mod.hash_bwd_f32(N(xyz), N(g), lv_o)
    np.testing.assert_allclose(N(enc.hash_table.grad), ref, rtol=1e-4, atol=1e-6 * np.abs(ref).max())


def test_hash_empty_and_ragged(vn):
    lv = vn.hash_levels(16, 1024, 16, 2 ** 19)
    table = torch.rand(2 * lv.total_entries, device=DEV)
    out = torch.empty(0, 32, device=DEV)
    vn.call("vn_hash_encode_fwd_f32", torch.empty(0, 3, device=DEV), table, out, 0, lv, 0)
    for S in (1, 31, 33, 257):
        xyz = torch.rand(S, 3, device=DEV)
        o1 = torch.empty(S, 32, device=DEV); o2 = torch.empty(S, 32, device=DEV)
        vn.call("vn_hash_encode_fwd_f32", xyz, table, o1, S, lv, 0)
        vn.call("vn_hash_encode_fwd_f32", xyz, table, o2, S, lv, 16)
        assert torch.equal(o1, o2)


# ------------------------------------------------------------------------------------ a4
def test_hash_half(vn, oracle_mod):
    lv_o = oracle_mod.HashLevels(16, 1024, 16, 2 ** 19)
    lv = vn.hash_levels(16, 1024, 16, 2 ** 19)
    rng = np.random.default_rng(5)
    table = (rng.random((lv_o.total, 2), dtype=np.float32) * 2 - 1)
    xyz = np.concatenate([rng.random((4000, 3)).astype(np.float32), ray_coherent_points(32, 64)])
    S = xyz.shape[0]
    table_h = torch.empty(lv_o.total, 2, dtype=torch.float16, device=DEV)
    vn.call("vn_f32_to_f16", T(table), table_h, table.size)
    np.testing.assert_array_equal(N(table_h), table.astype(np.float16))
    out = torch.empty(S, 16, 2, dtype=torch.float16, device=DEV)
    vn.call("vn_hash_encode_fwd_f16", T(xyz), table_h, out, S, lv, 0)
    ref = oracle_mod.hash_fwd_f16(xyz, table.astype(np.float16), lv_o)
    np.testing.assert_allclose(N(out).reshape(S, 32).astype(np.float32), ref.astype(np.float32), rtol=2e-3, atol=1e-3)
    # and against the fp32 oracle on the fp16-rounded table
    ref32 = oracle_mod.hash_fwd_f32(xyz, table.astype(np.float16).astype(np.float32).reshape(-1), lv_o)
    np.testing.assert_allclose(N(out).reshape(S, 32).astype(np.float32), ref32, rtol=2e-3, atol=2e-3)
    dout = rng.normal(size=(S, 16, 2)).astype(np.float16)
    dout[::7] = 0      # zero-skip rule
    grad = torch.zeros(lv_o.total, 2, device=DEV)
    vn.call("vn_hash_encode_bwd_f16", T(xyz), T(dout), grad, S, lv, 0)
    gref = oracle_mod.hash_bwd_f16(xyz, dout, lv_o)
    np.testing.assert_allclose(N(grad), gref, rtol=1e-2, atol=1e-5 * np.abs(gref).max())


def test_hash_half_module(vn, oracle_mod):
    from virus_nerf_b200.modules.hash_encoder_half import HashEncoder
    enc = HashEncoder(max_params=2 ** 19, levels=16, base_res=16, max_res=1024).to(DEV)
    assert enc.hash_table.shape == (5710032, 2) and enc.hash_grad.shape == (5710032, 2)
    lv_o = oracle_mod.HashLevels(16, 1024, 16, 2 ** 19)
    with torch.no_grad():                                  # the reference's U(-1e-4, 1e-4) init carries no signal
        torch.manual_seed(2)
        enc.hash_table.copy_(torch.rand_like(enc.hash_table) * 2 - 1)
    xyz = torch.cat([torch.rand(500, 3, device=DEV), T(ray_coherent_points(8, 64))])
    S = xyz.shape[0]
    out = enc(xyz)
    assert out.dtype == torch.float16 and out.shape == (S, 32)
    # forward of the module == oracle's half kernel on the fp16 copy of the table (hash_encoder_half.py:367)
    table_h = N(enc.hash_table).astype(np.float16)
    ref = oracle_mod.hash_fwd_f16(N(xyz), table_h, lv_o)
    np.testing.assert_allclose(N(out).astype(np.float32), ref.astype(np.float32).reshape(S, 32), rtol=2e-3, atol=1e-3)
    # backward of the module (autograd -> hash_grad, zero-skip rule) == oracle.hash_bwd_f16 of the same fp16 gradient
    g = torch.randn(S, 32, device=DEV).half()
    g[::7] = 0
    out.backward(g)
    gref = oracle_mod.hash_bwd_f16(N(xyz), N(g).reshape(S, 16, 2), lv_o)
    assert enc.hash_table.grad is not None and enc.hash_table.grad.dtype == torch.float32
    np.testing.assert_allclose(N(enc.hash_table.grad), gref.reshape(-1, 2), rtol=1e-2, atol=1e-5 * np.abs(gref).max())
    assert torch.equal(enc.hash_table.grad, enc.hash_grad)          # the reference hands out its hash_grad buffer (:352-358)


# ------------------------------------------------------------------------------------ a5
def test_ray_aabb_bit_exact(vn, oracle_mod, scene_rays):
    ro, rd, _ = scene_rays
    for scale in (0.5, 2.0):
        hits = torch.empty(ro.shape[0], 2, device=DEV)
        vn.call("vn_ray_aabb", T(ro), T(rd), scale, ro.shape[0], hits)
        np.testing.assert_array_equal(N(hits), oracle_mod.ray_aabb(ro, rd, scale))


# ------------------------------------------------------------------------------------ a6
@pytest.mark.parametrize("bf_name", ["carved", "full", "empty", "random"])
@pytest.mark.parametrize("scale,esf,cascades", [(0.5, 0.0, 1), (2.0, 1 / 256, 3)])
def test_march_train_bit_exact(vn, oracle_mod, scene_rays, bf_name, scale, esf, cascades):
    from virus_nerf_b200.modules.ray_march import raymarching_train
    ro, rd, bfs = scene_rays
    bf = bfs[bf_name]
    if cascades > 1:
        bf = np.concatenate([bf] + [np.random.default_rng(c).integers(0, 256, bf.shape[0]).astype(np.uint8)
                                    for c in range(1, cascades)])
    n = ro.shape[0]
    hits = oracle_mod.ray_aabb(ro, rd, scale)
    noise = np.random.default_rng(7).random(n).astype(np.float32)
    rays_a, xyzs, dirs, deltas, ts, total = raymarching_train(T(ro), T(rd), T(hits), T(bf), cascades, scale, esf, 128,
                                                              1024, noise=T(noise))
    o_ra, o_xyz, o_dirs, o_de, o_ts, o_total = oracle_mod.march_train(ro, rd, hits, bf, noise, cascades, scale, esf, 128)
    assert int(total) == o_total
    np.testing.assert_array_equal(N(rays_a), o_ra)
    np.testing.assert_array_equal(N(ts), o_ts)
    np.testing.assert_array_equal(N(deltas), o_de)
    np.testing.assert_array_equal(N(xyzs), o_xyz)
    np.testing.assert_array_equal(N(dirs), o_dirs)


@pytest.mark.parametrize("bf_name", ["carved", "full", "random"])
@pytest.mark.parametrize("scale,esf,cascades", [(0.5, 0.0, 1), (2.0, 1 / 256, 3)])
@pytest.mark.parametrize("tile", [1, 6])          # 6 x 3070 rays >= 16384: the thread-per-ray marcher
def test_march_single_pass_equals_two_pass(vn, scene_rays, bf_name, scale, esf, cascades, tile):
    """vn_march_train_count_rows + vn_march_train_expand (the fast step's single-pass march) produce
    bit-identical rays_a / xyzs / dirs / deltas / ts / unit positions to count + write (which the tests
    above pin to the oracle)"""
    ro, rd, bfs = scene_rays
    bf = bfs[bf_name]
    if cascades > 1:
        bf = np.concatenate([bf] + [np.random.default_rng(c).integers(0, 256, bf.shape[0]).astype(np.uint8)
                                    for c in range(1, cascades)])
    ro, rd = np.tile(ro, (tile, 1)), np.tile(rd, (tile, 1))
    n = ro.shape[0]
    o, d, b = T(ro), T(rd), T(bf)
    noise = T(np.random.default_rng(7).random(n).astype(np.float32))
    hits = torch.empty(n, 2, device=DEV)
    vn.call("vn_ray_aabb", o, d, scale, n, hits)
    res = []
    for single in (False, True):
        counts = torch.empty(n, dtype=torch.int32, device=DEV)
        rays_a = torch.empty(n, 3, dtype=torch.int32, device=DEV)
        counter = torch.zeros(2, dtype=torch.int32, device=DEV)
        tmp = torch.empty(vn.scan_tmp_ints(n), dtype=torch.int32, device=DEV)
        rows = torch.full((n, 1024), float("nan"), device=DEV)
        if single:
            vn.call("vn_march_train_count_rows", o, d, hits, b, noise, n, cascades, 128, scale, esf, 1024, counts, rays_a,
                    counter, tmp, rows)
        else:
            vn.call("vn_march_train_count", o, d, hits, b, noise, n, cascades, 128, scale, esf, 1024, counts, rays_a,
                    counter, tmp)
        S = int(counter[0])
        outs = [torch.full((S, 3), float("nan"), device=DEV), torch.full((S, 3), float("nan"), device=DEV),
                torch.full((S,), float("nan"), device=DEV), torch.full((S,), float("nan"), device=DEV),
                torch.full((S, 3), float("nan"), device=DEV)]
        if single:
            vn.call("vn_march_train_expand", o, d, rays_a, rows, n, 1024, 128, scale, esf, S, *outs)
        else:
            vn.call("vn_march_train_write", o, d, hits, b, noise, n, cascades, 128, scale, esf, rays_a, S, *outs)
        res.append((S, rays_a, outs))
    assert res[0][0] == res[1][0] > 0
    assert torch.equal(res[0][1], res[1][1])
    for a, c in zip(res[0][2], res[1][2]):
        assert not torch.isnan(c).any()
        assert torch.equal(a, c)


def test_march_train_max_samples_and_empty(vn, oracle_mod, scene_rays):
    from virus_nerf_b200.modules.ray_march import raymarching_train
    ro, rd, bfs = scene_rays
    hits = oracle_mod.ray_aabb(ro, rd, 0.5)
    noise = np.zeros(ro.shape[0], np.float32)
    rays_a, *_rest, total = raymarching_train(T(ro), T(rd), T(hits), T(bfs["full"]), 1, 0.5, 0.0, 128, 17, noise=T(noise))
    o = oracle_mod.march_train(ro, rd, hits, bfs["full"], noise, 1, 0.5, 0.0, 128, max_samples=17)
    np.testing.assert_array_equal(N(rays_a), o[0])
    assert N(rays_a)[:, 2].max() == 17
    e = torch.empty(0, 3, device=DEV)
    ra, x, d, de, ts, tot = raymarching_train(e, e, torch.empty(0, 2, device=DEV), T(bfs["full"]), 1, 0.5, 0.0, 128, 1024)
    assert ra.shape == (0, 3) and x.shape == (0, 3) and int(tot) == 0


# ------------------------------------------------------------------------------------ a7
@pytest.mark.parametrize("max_samples", [1, 4, 64])
def test_march_test_bit_exact(vn, oracle_mod, scene_rays, max_samples):
    from virus_nerf_b200.modules.ray_march import raymarching_test
    ro, rd, bfs = scene_rays
    bf = bfs["carved"]
    hits = oracle_mod.ray_aabb(ro, rd, 0.5)
    alive = np.random.default_rng(8).permutation(ro.shape[0])[: ro.shape[0] // 2].astype(np.int64)
    alive.sort()
    hits_g = T(hits.copy())
    hits_o = hits.copy()
    for _round in range(3):      # successive rounds continue from the mutated hits_t
        pk, ri, de, ts = raymarching_test(T(ro), T(rd), hits_g, T(alive), T(bf), 1, 0.5, 0.0, 128, max_samples)
        o_pk, o_ri, o_de, o_ts = oracle_mod.march_test(ro, rd, hits_o, alive, bf, 1, 0.5, 0.0, 128, max_samples)
        np.testing.assert_array_equal(N(pk), o_pk)
        np.testing.assert_array_equal(N(ri), o_ri)
        np.testing.assert_array_equal(N(de), o_de)
        np.testing.assert_array_equal(N(ts), o_ts)
        np.testing.assert_array_equal(N(hits_g), hits_o)


# ------------------------------------------------------------------------------------ a8/a9/a10
def _composite_inputs(oracle_mod, scene_rays, seed=9, sigma_scale=30.0):
    ro, rd, bfs = scene_rays
    hits = oracle_mod.ray_aabb(ro, rd, 0.5)
    noise = np.random.default_rng(seed).random(ro.shape[0]).astype(np.float32)
    rays_a, xyzs, dirs, deltas, ts, total = oracle_mod.march_train(ro, rd, hits, bfs["carved"], noise, 1, 0.5, 0.0, 128)
    rng = np.random.default_rng(seed)
    sigmas = (rng.random(total).astype(np.float32) ** 4) * sigma_scale * 20
    rgbs = rng.random((total, 3)).astype(np.float32)
    return rays_a, sigmas, rgbs, deltas, ts


@pytest.mark.parametrize("sigma_scale", [1.0, 30.0, 3000.0])
def test_composite_train_fwd_bwd(vn, oracle_mod, scene_rays, sigma_scale):
    from virus_nerf_b200.modules.volume_train import VolumeRenderer
    rays_a, sigmas, rgbs, deltas, ts = _composite_inputs(oracle_mod, scene_rays, sigma_scale=sigma_scale)
    n = rays_a.shape[0]
    sg = T(sigmas).requires_grad_(True)
    cg = T(rgbs).requires_grad_(True)
    vr, op, dp, rgb, ws = VolumeRenderer()(sg, cg, T(deltas), T(ts), T(rays_a), 1e-4)
    o_total, o_op, o_dp, o_rgb, o_ws = oracle_mod.composite_train_fwd(sigmas, rgbs, deltas, ts, rays_a, 1e-4)
    assert int(vr) == int(o_total.sum())
    np.testing.assert_allclose(N(op), o_op, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(dp), o_dp, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(rgb), o_rgb, rtol=1e-5, atol=1e-7)
    # 1 - exp(-x) is quantised to ulp(1) = 6e-8 near x = 0, so one ulp of expf shows up as an
    # absolute 6e-8 in w = a * T (T <= 1); the per-ray sums above are unaffected at rtol 1e-5
    np.testing.assert_allclose(N(ws), o_ws, rtol=1e-5, atol=1.5e-7)
    rng = np.random.default_rng(10)
    dO, dD, dC = (rng.normal(size=n).astype(np.float32), rng.normal(size=n).astype(np.float32),
                  rng.normal(size=(n, 3)).astype(np.float32))
    dW = rng.normal(size=sigmas.shape[0]).astype(np.float32) * 0.1
    (op * T(dO)).sum().add((dp * T(dD)).sum()).add((rgb * T(dC)).sum()).add((ws * T(dW)).sum()).backward()
    r_ds, r_dc = oracle_mod.composite_train_bwd(sigmas, rgbs, deltas, ts, rays_a, 1e-4, dO, dD, dC, dW)
    np.testing.assert_allclose(N(sg.grad), r_ds, rtol=1e-4, atol=1e-6 * np.abs(r_ds).max())
    np.testing.assert_allclose(N(cg.grad), r_dc, rtol=1e-4, atol=1e-6 * np.abs(r_dc).max())


@pytest.mark.parametrize("sigma_scale", [30.0, 3000.0])
def test_composite_with_fused_loss_equals_separate_kernels(vn, oracle_mod, scene_rays, sigma_scale):
    """vn_composite_loss_fwd / _bwd (what the native step runner enqueues) == vn_composite_train_fwd + vn_loss_fwd and
    vn_loss_bwd + vn_composite_train_bwd: per-ray outputs, valid counts and sample gradients bit-identical, loss sums equal
    up to the order of the additions"""
    rays_a, sigmas, rgbs, deltas, ts = _composite_inputs(oracle_mod, scene_rays, sigma_scale=sigma_scale)
    n, S = rays_a.shape[0], sigmas.shape[0]
    rng = np.random.default_rng(3)
    gt = T(rng.random((n, 3)).astype(np.float32))
    uss = rng.random(n).astype(np.float32) * 2; uss[::3] = np.nan
    tof = rng.random(n).astype(np.float32); tof[1::2] = np.nan
    uss, tof = T(uss), T(tof)
    sg, cg, de, tt, ra = T(sigmas), T(rgbs), T(deltas), T(ts), T(rays_a)
    scale = torch.tensor([2.0 ** 12], device=DEV)
    w = (1.0, 0.7, 0.3, 0.0)
    res = []
    for fused in (False, True):
        vr = torch.zeros(n, dtype=torch.int32, device=DEV)
        op = torch.zeros(n, device=DEV); dp = torch.zeros(n, device=DEV); rgb = torch.zeros(n, 3, device=DEV)
        ws = torch.zeros(S, device=DEV)
        acc = torch.zeros(8, device=DEV); loss = torch.zeros(1, device=DEV)
        ds = torch.zeros(S, device=DEV); dc = torch.zeros(S, 3, device=DEV)
        if fused:
            vn.call("vn_composite_loss_fwd", sg, cg, de, tt, ra, n, S, 1e-4, vr, op, dp, rgb, ws, gt, uss, tof, None, 1.0, 0.03,
                    acc[:4], acc[4:])
            vn.call("vn_composite_loss_bwd", sg, cg, de, tt, ra, n, S, 1e-4, rgb, op, dp, gt, uss, tof, None, 1.0, 0.03,
                    acc[:4], acc[4:], *w, scale, ds, dc, loss)
        else:
            vn.call("vn_composite_train_fwd", sg, cg, de, tt, ra, n, S, 1e-4, vr, op, dp, rgb, ws)
            vn.call("vn_loss_fwd", rgb, op, dp, gt, uss, tof, None, n, 1.0, 0.03, acc[:4], acc[4:])
            d_rgb = torch.zeros(n, 3, device=DEV); d_dp = torch.zeros(n, device=DEV); d_op = torch.zeros(n, device=DEV)
            vn.call("vn_loss_bwd", rgb, op, dp, gt, uss, tof, None, n, 1.0, 0.03, acc[:4], acc[4:], *w, scale, d_rgb, d_dp, d_op, loss)
            vn.call("vn_composite_train_bwd", sg, cg, de, tt, ra, n, S, 1e-4, d_op, d_dp, d_rgb, None, ds, dc)
        res.append((vr, op, dp, rgb, ws, acc.clone(), loss.clone(), ds, dc))
    a, b = res
    for k in range(5):
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(a[5][4:], b[5][4:]) and float(a[5][5]) > 0 and float(a[5][6]) > 0        # counts are exact
    np.testing.assert_allclose(N(a[5][:4]), N(b[5][:4]), rtol=1e-5)
    np.testing.assert_allclose(N(a[6]), N(b[6]), rtol=1e-5)
    np.testing.assert_allclose(N(a[7]), N(b[7]), rtol=1e-6, atol=1e-6 * float(a[7].abs().max()))
    np.testing.assert_allclose(N(a[8]), N(b[8]), rtol=1e-6, atol=1e-6 * float(a[8].abs().max()))
    assert float(a[7].abs().max()) > 0


def test_composite_closed_form(vn):
    """constant-sigma slab: opacity = 1 - exp(-sigma L); zero-sample rays give zeros"""
    from virus_nerf_b200.modules.volume_train import VolumeRenderer
    ns, sigma, delta = 200, 7.0, 0.004
    rays_a = torch.tensor([[0, 0, ns], [1, ns, 0], [2, ns, 3]], dtype=torch.int32, device=DEV)
    S = ns + 3
    sig = torch.full((S,), sigma, device=DEV); de = torch.full((S,), delta, device=DEV)
    ts = torch.arange(S, device=DEV, dtype=torch.float32) * delta
    rgbs = torch.ones(S, 3, device=DEV)
    vr, op, dp, rgb, ws = VolumeRenderer()(sig, rgbs, de, ts, rays_a, 1e-4)
    assert abs(float(op[0]) - (1 - np.exp(-sigma * delta * ns))) < 1e-5
    assert float(op[1]) == 0.0 and float(dp[1]) == 0.0 and int(vr) == ns + 3
    np.testing.assert_allclose(N(rgb[0]), N(op[0]).repeat(3), rtol=1e-6)


def test_composite_test_kernel(vn, oracle_mod, scene_rays):
    from virus_nerf_b200.modules.volume_render_test import composite_test
    rays_a, sigmas, rgbs, deltas, ts = _composite_inputs(oracle_mod, scene_rays, sigma_scale=300.0)
    n = rays_a.shape[0]
    alive = np.arange(n, dtype=np.int64)
    pack = np.stack([rays_a[:, 1], np.minimum(rays_a[:, 2], 8)], -1).astype(np.int64)
    rng = np.random.default_rng(11)
    op = (rng.random(n) * 0.5).astype(np.float32); dp = rng.random(n).astype(np.float32)
    rgb = rng.random((n, 3)).astype(np.float32)
    g_alive, g_op, g_dp, g_rgb = T(alive.copy()), T(op.copy()), T(dp.copy()), T(rgb.copy())
    composite_test(T(sigmas), T(rgbs), T(deltas), T(ts), T(pack), g_alive, 1e-2, g_op, g_dp, g_rgb)
    oracle_mod.composite_test(sigmas, rgbs, deltas, ts, pack, alive, 1e-2, op, dp, rgb)
    np.testing.assert_array_equal(N(g_alive), alive)
    assert (alive == -1).any() and (alive >= 0).any()
    np.testing.assert_allclose(N(g_op), op, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(g_dp), dp, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(g_rgb), rgb, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------ a11/a15
def test_sh_and_morton_and_packbits(vn, oracle_mod):
    from virus_nerf_b200.modules.spherical_harmonics import DirEncoder
    from virus_nerf_b200.modules import utils as U
    d = np.random.default_rng(12).random((1000, 3)).astype(np.float32)
    np.testing.assert_allclose(N(DirEncoder()(T(d))), oracle_mod.sh_encode(d), rtol=1e-6, atol=1e-7)
    r = np.arange(128, dtype=np.int32)
    coords = np.stack(np.meshgrid(r, r, r, indexing="ij"), -1).reshape(-1, 3)
    m = U.morton3D(T(coords))
    np.testing.assert_array_equal(N(m), oracle_mod.morton3d(coords))
    assert np.unique(N(m)).size == 128 ** 3                        # bijection
    np.testing.assert_array_equal(N(U.morton3D_invert(m)), coords)
    np.testing.assert_array_equal(N(U.morton3D_invert(m)), oracle_mod.morton3d_invert(N(m)))
    g = np.random.default_rng(13).random(128 ** 3).astype(np.float32)
    g[:16] = 0.5                                                   # strict > threshold
    bf = torch.zeros(128 ** 3 // 8, dtype=torch.uint8, device=DEV)
    U.packbits(T(g), 0.5, bf)
    np.testing.assert_array_equal(N(bf), oracle_mod.packbits(g, 0.5))
    assert N(bf)[0] == 0 and N(bf)[1] == 0


# ------------------------------------------------------------------------------------ a14
def _occ_grid(G=128, scale=0.5):
    from virus_nerf_b200 import synthetic
    from virus_nerf_b200.modules.occupancy_grid import OccupancyGrid
    args = synthetic.make_args(device=DEV, scale=scale)
    torch.manual_seed(0)
    return OccupancyGrid(args, G, scene=None, dataset=None, fct_density=None), args


@pytest.mark.parametrize("G", [32, 128])
def test_occ_calc_pos_and_ray_prob(vn, oracle_mod, scene_rays, G):
    ro, rd, _ = scene_rays
    ro, rd = ro[:777], rd[:777] * np.float32(1.7)      # un-normalised directions are normalised inside
    og, args = _occ_grid(G)
    noise = np.random.default_rng(14).random((777, 32, 3)).astype(np.float32)
    for nz in (None, noise):
        dists, pos, idx = og._calcPos(T(ro), T(rd), add_noise=nz is not None, noise=None if nz is None else T(nz))
        o_d, o_p, o_i = oracle_mod.occ_calc_pos(ro, rd, nz, 32, G, 0.5, og.nerf_pos_noise_every_m)
        np.testing.assert_array_equal(N(dists), o_d)
        np.testing.assert_array_equal(N(pos), o_p)
        np.testing.assert_array_equal(N(idx), o_i)
    meas = (np.random.default_rng(15).random(777) * 0.8 + 0.05).astype(np.float32)
    po, pe = og._rayProb(T(meas), dists)
    o_po, o_pe = oracle_mod.occ_ray_prob(meas, o_d, og.false_detection_prob_every_m, og.std_every_m)
    np.testing.assert_allclose(N(po), o_po, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(pe), o_pe, rtol=1e-5, atol=1e-7)
    # return_probs=True (occupancy_grid.py:387-388): the four factors, against the reference's own torch expressions
    # (:361-381) written with the mirror's _sensorEmptyPDF / _sensorOccupiedPDF
    po2, pe2, eq_emp, eq_occ, nl_emp, nl_occ = og._rayProb(T(meas), dists, return_probs=True)
    assert torch.equal(po2, po) and torch.equal(pe2, pe)
    m = T(meas)
    r_eq_emp = og._sensorEmptyPDF(shape=dists.shape)
    r_eq_occ = r_eq_emp + og._sensorOccupiedPDF(meas=m[:, None], dists=dists)
    r_nl_emp = (1 - r_eq_emp * dists).clamp(min=og.prob_min)
    y = torch.linspace(0, 1, og.I, device=DEV, dtype=torch.float32)[None, :] * m[:, None]
    integral = og._sensorOccupiedPDF(meas=y[:, None, :], dists=dists[:, :, None]).sum(dim=2) * (m / og.I)[:, None]
    r_nl_occ = (r_nl_emp - integral).clamp(min=og.prob_min)
    for got, ref in ((eq_emp, r_eq_emp), (eq_occ, r_eq_occ), (nl_emp, r_nl_emp), (nl_occ, r_nl_occ)):
        np.testing.assert_allclose(N(got), N(ref), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(N(eq_emp * nl_emp), N(pe), rtol=1e-6)
    np.testing.assert_allclose(N(eq_occ * nl_occ), N(po), rtol=1e-6)


def test_occ_bayes_update_and_pack_bit_exact(vn, oracle_mod, scene_rays):
    ro, rd, _ = scene_rays
    og, args = _occ_grid(128)
    grid0 = N(og.occ_3d_grid).copy()
    _, _, idx = og._calcPos(T(ro), T(rd), add_noise=False)
    n = idx.shape[0]
    rng = np.random.default_rng(16)
    po = (rng.random(n) * 0.9 + 0.05).astype(np.float32); pe = (rng.random(n) * 0.9 + 0.05).astype(np.float32)
    og._updateGrid(idx, T(po), T(pe))
    ref = grid0.copy()
    oracle_mod.occ_update_grid(ref, N(idx), po, pe)
    np.testing.assert_array_equal(N(og.occ_3d_grid), ref)            # duplicates: last index wins
    assert (N(og._winner) == -1).all()                               # scratch restored
    assert np.unique(N(idx), axis=0).shape[0] < n                    # the case did contain duplicates
    # fixed point: po == pe leaves p unchanged (up to rounding of p*po/(p*po+(1-p)*po))
    og.update_step = 0
    og._decayAndPack(apply_decay=True)
    bf_ref = oracle_mod.occ_decay_pack(ref, og.grid_decay, True, 0.5)
    np.testing.assert_array_equal(N(og.occ_3d_grid), ref)
    np.testing.assert_array_equal(N(og.getBitfield()), bf_ref)
    # debug round trip of the reference (trainer_plot.py:73-86)
    cart = og.morton2cartesian(og.bitfield2morton(og.getBitfield()))
    assert torch.equal(cart, og.getBinaryCartesianGrid(0.5))


def test_occ_nerf_prob(vn, oracle_mod):
    og, args = _occ_grid(32)
    rho = (np.random.default_rng(17).random(32 * 700) ** 3 * 40 + 1e-3).astype(np.float32)
    og.fct_density = lambda x: T(rho)
    for thr_max in (5.91, 1e9):
        args.occ_grid.nerf_threshold_max = thr_max
        po, pe = og._nerfProb(torch.zeros(rho.shape[0], 3, device=DEV))
        o_po, o_pe = oracle_mod.occ_nerf_prob(rho, thr_max, args.occ_grid.nerf_threshold_slope)
        np.testing.assert_allclose(N(po), o_po, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(N(pe), o_pe, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------ f1
def test_adam_step(vn, oracle_mod):
    n = 100003
    rng = np.random.default_rng(18)
    p = rng.normal(size=n).astype(np.float32); g = (rng.normal(size=n) * 2 ** 19).astype(np.float32)
    m = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    gp, gm, gv = T(p.copy()), T(m.copy()), T(v.copy())
    found = torch.zeros(1, device=DEV); scale = torch.tensor([2.0 ** 19], device=DEV)
    for step in (1, 2, 3):
        vn.call("vn_grad_check", T(g), n, found)
        vn.call("vn_adam_step", gp, T(g), gm, gv, n, 1.0, 5e-3, 0.9, 0.999, 1e-15, step, found, scale)
        assert oracle_mod.adam_step(p, g, m, v, 2.0 ** -19, 5e-3, 0.9, 0.999, 1e-15, step) == 0
    assert float(found) == 0.0
    np.testing.assert_allclose(N(gp), p, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(N(gm), m, rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(N(gv), v, rtol=1e-5, atol=1e-12)
    g[77] = np.inf                                                    # GradScaler: skip + back off
    before = N(gp).copy()
    vn.call("vn_grad_check", T(g), n, found)
    assert float(found) == 1.0
    vn.call("vn_adam_step", gp, T(g), gm, gv, n, 1.0, 5e-3, 0.9, 0.999, 1e-15, 4, found, scale)
    np.testing.assert_array_equal(N(gp), before)
    tracker = torch.zeros(1, dtype=torch.int32, device=DEV)
    vn.call("vn_scaler_update", scale, tracker, found, 2.0, 0.5, 2000)
    assert float(scale) == 2.0 ** 18 and float(found) == 0.0


# ------------------------------------------------------------------------------------ errors
def test_error_conventions(vn):
    lv = vn.hash_levels(16, 1024, 16, 2 ** 19)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        vn.call("vn_sh_encode", torch.zeros(4, 3), 4, torch.zeros(4, 16))
    with pytest.raises(RuntimeError, match="contiguous"):
        vn.call("vn_sh_encode", torch.zeros(3, 4, device=DEV).t(), 4, torch.zeros(4, 16, device=DEV))
    with pytest.raises(RuntimeError, match="null pointer"):
        vn.call("vn_hash_encode_fwd_f32", torch.zeros(4, 3, device=DEV), None, torch.zeros(4, 32, device=DEV), 4, lv, 0)
    with pytest.raises(RuntimeError, match="power of two"):
        vn.call("vn_occ_decay_pack", torch.zeros(100 ** 3, device=DEV), 100, 1.0, 0, 0.5,
                torch.zeros(100 ** 3 // 8, dtype=torch.uint8, device=DEV))
    assert vn.launch_count() > 0


# ------------------------------------------------------------------------------------ full-size properties
def test_full_size_properties(vn):
    """BASELINE sizes (T=2^22 table, 2^20 points): size-independent properties instead of the
    oracle -- linearity of the encoder in the table, adjointness <dout, enc(table)> ==
    <scatter(dout), table>, and agreement of the three level-grouping variants."""
    lv = vn.hash_levels(16, 1024, 16, 2 ** 22)
    S = 1 << 20
    g = torch.Generator(device=DEV).manual_seed(0)
    xyz = torch.rand(S, 3, device=DEV, generator=g)
    t1 = torch.rand(2 * lv.total_entries, device=DEV, generator=g)
    t2 = torch.rand(2 * lv.total_entries, device=DEV, generator=g)
    o1 = torch.empty(S, 32, device=DEV); o2 = torch.empty(S, 32, device=DEV); o12 = torch.empty(S, 32, device=DEV)
    vn.call("vn_hash_encode_fwd_f32", xyz, t1, o1, S, lv, 0)
    vn.call("vn_hash_encode_fwd_f32", xyz, t2, o2, S, lv, 16)
    vn.call("vn_hash_encode_fwd_f32", xyz, t1 + t2, o12, S, lv, 32)
    torch.testing.assert_close(o12, o1 + o2, rtol=1e-5, atol=1e-5)
    dout = torch.randn(S, 32, device=DEV, generator=g)
    grad = torch.zeros_like(t1)
    vn.call("vn_hash_encode_bwd_f32", xyz, dout, grad, S, lv, 0)
    lhs = (dout.double() * o1.double()).sum()
    rhs = (grad.double() * t1.double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-5 * abs(float(lhs)) + 1e-2


def test_march_properties_full_batch(vn, scene_rays):
    """2^18 rays: counts sum to the total, per-ray ts strictly increasing, every emitted sample
    lies in an occupied cell"""
    from virus_nerf_b200 import synthetic
    from virus_nerf_b200.modules.intersection import ray_aabb_intersection
    from virus_nerf_b200.modules.ray_march import raymarching_train
    ds = synthetic.SyntheticDataset(pool_size=1 << 18, n_images=16, device=DEV)
    ro, rd = ds.pool["rays_o"], ds.pool["rays_d"]
    bf = T(scene_rays[2]["carved"])
    hits = ray_aabb_intersection(ro, rd, 0.5)
    rays_a, xyzs, dirs, deltas, ts, total = raymarching_train(ro, rd, hits, bf, 1, 0.5, 0.0, 128, 1024)
    ra = rays_a.long()
    assert int(ra[:, 2].sum()) == int(total) == xyzs.shape[0]
    assert torch.equal(ra[:, 1], torch.cumsum(ra[:, 2], 0) - ra[:, 2])
    seg = torch.repeat_interleave(torch.arange(ra.shape[0], device=DEV), ra[:, 2])
    same = seg[1:] == seg[:-1]
    assert bool((ts[1:][same] > ts[:-1][same]).all())
    cell = torch.clamp((0.5 * (xyzs / 0.5 + 1) * 128), 0, 127).to(torch.int32)
    from virus_nerf_b200.modules.utils import morton3D
    m = morton3D(cell).long()
    occ = (bf[m // 8].long() >> (m % 8)) & 1
    assert bool(occ.all())
