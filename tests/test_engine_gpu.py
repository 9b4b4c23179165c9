"""GPU end-to-end parity: the render drivers, NGP module and the train-step engine against the
oracle pipeline (oracle/pipeline.py) with the same weights, rays, bitfield and jitter noise.
fp32 everywhere (autocast off) so that the MLP (cuBLAS fp32 stand-in / fused kernel in fp32
mode) is comparable at rtol 1e-4."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def N(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def setup():
    import oracle
    from oracle import pipeline
    from virus_nerf_b200 import synthetic
    from virus_nerf_b200.engine import TrainEngine
    oracle.build()
    args = synthetic.make_args(device=DEV, batch_size=256)
    ds = synthetic.SyntheticDataset(pool_size=1 << 13, n_images=8, device=DEV)
    eng = TrainEngine(args, ds, DEV, autocast=False)
    bf = synthetic.morton_pack(ds.scene.occupancy_bitfield(128))
    eng.model.occupancy_grid.bitfield = torch.from_numpy(bf).to(DEV)
    om = pipeline.OracleNGP(threads=4)
    om.load_from(eng.model)
    return eng, om, ds, bf, args, pipeline


def _data_cpu(data):
    return {"rays_o": data["rays_o"].cpu(), "rays_d": data["rays_d"].cpu(), "rgb": data["rgb"].cpu(),
            "depth": {k: v.cpu() for k, v in data["depth"].items()}}


def test_state_dict_keys(setup):
    eng = setup[0]
    keys = set(eng.model.state_dict().keys())
    assert {"pos_encoder.hash_table", "xyz_encoder.hidden_layers.0.weight", "xyz_encoder.output_layer.weight",
            "rgb_net.hidden_layers.0.weight", "rgb_net.hidden_layers.1.weight", "rgb_net.output_layer.weight",
            "center", "xyz_min", "xyz_max", "half_size"} <= keys
    assert eng.model.state_dict()["pos_encoder.hash_table"].shape == (11420064,)
    assert "pos_encoder.offsets" not in keys and "pos_encoder.hash_map_sizes" not in keys   # persistent=False


def test_ngp_forward_matches_oracle(setup):
    eng, om, ds, bf, args, pipeline = setup
    x = (torch.rand(4000, 3, device=DEV) - 0.5) * 0.98
    d = torch.randn(4000, 3, device=DEV)
    with torch.no_grad():
        sig, rgb = eng.model(x, d)
        o_sig, o_rgb = om.forward(x.cpu(), d.cpu())
    np.testing.assert_allclose(N(sig), N(o_sig), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(N(rgb), N(o_rgb), rtol=1e-4, atol=1e-6)


def test_render_train_and_grads_match_oracle(setup, monkeypatch):
    eng, om, ds, bf, args, pipeline = setup
    from virus_nerf_b200.modules import ray_march, rendering
    data = ds(256, args.training.sampling_strategy)
    noise = torch.rand(256, device=DEV)
    orig = ray_march.raymarching_train
    monkeypatch.setattr(rendering, "raymarching_train", lambda *a, **k: orig(*a, noise=noise, **k))
    eng.flat_g.zero_()
    loss, terms, res = eng.forward_loss(data)
    loss.backward()
    o_res = pipeline.render_train(om, data["rays_o"].cpu(), data["rays_d"].cpu(), bf, N(noise))
    o_loss = pipeline.loss_fn(o_res, _data_cpu(data))
    for p in om.parameters():
        p.grad = None
    o_loss.backward()
    assert int(res["rm_samples"]) == o_res["rm_samples"]
    np.testing.assert_array_equal(N(res["ts"]), o_res["ts"])
    for k in ("opacity", "depth", "rgb"):
        np.testing.assert_allclose(N(res[k]), N(o_res[k]), rtol=1e-4, atol=1e-6)
    assert abs(float(loss) - float(o_loss)) <= 1e-4 * abs(float(o_loss))
    g_hash = N(eng.model.pos_encoder.hash_table.grad)
    r_hash = N(om.hash_table.grad)
    np.testing.assert_allclose(g_hash, r_hash, rtol=1e-3, atol=1e-5 * np.abs(r_hash).max())
    names = ["xyz_encoder.hidden_layers.0.weight", "xyz_encoder.output_layer.weight", "rgb_net.hidden_layers.0.weight",
             "rgb_net.hidden_layers.1.weight", "rgb_net.output_layer.weight"]
    params = dict(eng.model.named_parameters())
    for n, w in zip(names, om.W):
        np.testing.assert_allclose(N(params[n].grad), N(w.grad), rtol=1e-3, atol=1e-5 * float(w.grad.abs().max()))


def test_render_test_matches_oracle(setup):
    eng, om, ds, bf, args, pipeline = setup
    from virus_nerf_b200.modules.rendering import render
    from virus_nerf_b200 import synthetic
    data = ds(300, {"pixs": "random"})
    so, sd = synthetic.scan_rays(64, height=-0.05)
    ro = torch.cat([data["rays_o"], torch.from_numpy(so).to(DEV)])
    rd = torch.cat([data["rays_d"], torch.from_numpy(sd).to(DEV)])
    # make the field opaque enough for early termination to matter
    with torch.no_grad():
        eng.model.xyz_encoder.output_layer.weight[0].add_(0.25)
    om.load_from(eng.model)
    eng.model.fused_test_render = False                     # the reference's round loop (a7 + a10 kernels)
    res = render(eng.model, ro, rd, test_time=True, exp_step_factor=0.0)
    o = pipeline.render_test(om, ro.cpu(), rd.cpu(), bf)
    assert int(res["total_samples"]) == o["total_samples"]
    for k in ("opacity", "depth", "rgb"):
        np.testing.assert_allclose(N(res[k]), o[k], rtol=1e-4, atol=1e-5)
    eng.model.fused_test_render = True                      # loop-free path: same image
    fast = render(eng.model, ro, rd, test_time=True, exp_step_factor=0.0)
    for k in ("opacity", "depth", "rgb"):
        np.testing.assert_allclose(N(fast[k]), o[k], rtol=1e-4, atol=1e-5)
    assert 0 < int(fast["total_samples"]) <= o["total_samples"]
    with torch.no_grad():
        eng.model.xyz_encoder.output_layer.weight[0].sub_(0.25)
    om.load_from(eng.model)


def test_evaluation_consumers_match_oracle(setup, tmp_path):
    """lidar-like scan rendering (d_z = 0 rays, trainer.py:574-629), the density-map slice
    (trainer_base.py:92-140) and the .pth checkpoint round trip, against the oracle pipeline"""
    eng, om, ds, bf, args, pipeline = setup
    from virus_nerf_b200.training import evaluation as ev
    from virus_nerf_b200.modules.networks import NGP
    with torch.no_grad():
        eng.model.xyz_encoder.output_layer.weight[0].add_(0.25)
    om.load_from(eng.model)
    origins = torch.tensor([[0.05, -0.1, -0.05], [-0.2, 0.15, -0.05]], device=DEV)
    ro, rd, depth = ev.evaluation_depth_nerf(eng.model, origins, angle_res=96, batch_size=64)
    assert ro.shape == (192, 3) and (rd[:, 2] == 0).all()
    o = pipeline.render_test(om, torch.from_numpy(ro), torch.from_numpy(rd), bf)
    np.testing.assert_allclose(depth, o["depth"], rtol=1e-4, atol=1e-5)
    dm, dm_thr = ev.interfere_density_map(eng.model, res_map=24, height_c=-0.05, num_avg_heights=3, tolerance_c=0.02,
                                          threshold=1.0, cube_min=-0.5, cube_max=0.5, batch_size=500)
    pos = ev.createScanPos(24, -0.05, 3, 0.02, -0.5, 0.5, "cpu")
    ref = om.density(pos).detach().numpy().reshape(-1, 3).max(1).reshape(24, 24)
    np.testing.assert_allclose(dm, ref, rtol=1e-4, atol=1e-6)
    assert set(np.unique(dm_thr)) <= {0.0, 1.0} and ((dm >= 1.0) == (dm_thr == 1.0)).all()
    # checkpoint: the reference's state-dict keys, readable by a freshly built model
    path = ev.save_checkpoint(eng.model, str(tmp_path))
    sd = torch.load(path, map_location="cpu")
    assert "pos_encoder.hash_table" in sd and "rgb_net.output_layer.weight" in sd
    fresh = NGP(scale=args.model.scale, pos_encoder_type='hash', levels=16, max_res=1024, log2_T=19, args=args,
                dataset=ds).to(DEV)
    ev.load_checkpoint(fresh, path)
    fresh.occupancy_grid.bitfield = eng.model.occupancy_grid.bitfield
    fresh.fused_mlp = eng.model.fused_mlp
    _, _, depth2 = ev.evaluation_depth_nerf(fresh, origins, angle_res=96, batch_size=64)
    np.testing.assert_array_equal(depth2, depth)
    with torch.no_grad():
        eng.model.xyz_encoder.output_layer.weight[0].sub_(0.25)
    om.load_from(eng.model)


def test_train_steps_track_oracle(setup, monkeypatch):
    """three optimiser steps (no grid update in between): losses and parameters track the CPU
    trainer (fp32, Adam eps 1e-15; the GPU side goes through GradScaler 2^19 + fused Adam)"""
    eng, om, ds, bf, args, pipeline = setup
    from virus_nerf_b200.modules import ray_march, rendering
    om.load_from(eng.model)
    tr = pipeline.OracleTrainer(om, lr=args.training.lr)
    eng.step_idx = 1   # skip the occupancy update (tested separately)
    orig = ray_march.raymarching_train
    for it in range(3):
        data = ds(256, args.training.sampling_strategy)
        noise = torch.rand(256, device=DEV)
        monkeypatch.setattr(rendering, "raymarching_train", lambda *a, **k: orig(*a, noise=noise, **k))
        loss = float(eng.step(data))
        o_loss, _ = tr.step(_data_cpu(data), bf, N(noise))
        assert abs(loss - o_loss) <= 2e-3 * abs(o_loss), (it, loss, o_loss)
    w_g = N(eng.model.rgb_net.output_layer.weight)
    np.testing.assert_allclose(w_g, N(om.W[4]), rtol=0, atol=2e-3)


@pytest.mark.parametrize("fused", [True, False])
def test_fast_step_fp16_tracks_oracle_directly(fused):
    """the path that is BENCHED -- step_fast with the fp16 tcgen05 MLP, operand-chunk layout, single-pass march, PDL and
    (fused=True) the fused MLP-backward + hash-scatter kernel with the in-kernel inf check -- against the fp32 CPU
    trainer of oracle/pipeline.py on the same rays, bitfield and jitter.  Stated tolerance of the fp16 MLP
    (tests/test_mlp_gpu.py): 2e-2 relative on the loss, 5e-2 of the largest entry on gradients."""
    import oracle
    from oracle import pipeline
    from virus_nerf_b200 import synthetic
    from virus_nerf_b200.engine import TrainEngine
    oracle.build()
    args = synthetic.make_args(device=DEV, batch_size=256)
    ds = synthetic.SyntheticDataset(pool_size=1 << 13, n_images=8, device=DEV)
    eng = TrainEngine(args, ds, DEV, autocast=True, fused_scatter=fused)
    assert eng.enc_chunks and eng.fused_scatter == fused and eng.model.fused_mlp
    bf = synthetic.morton_pack(ds.scene.occupancy_bitfield(128))
    eng.model.occupancy_grid.bitfield = torch.from_numpy(bf).to(DEV)
    eng.step_idx = eng._prep_step = 1                      # no occupancy update (tested separately)
    om = pipeline.OracleNGP(threads=4)
    om.load_from(eng.model)
    tr = pipeline.OracleTrainer(om, lr=args.training.lr)
    for it in range(3):
        data = ds(256, args.training.sampling_strategy)
        noise = torch.rand(256, device=DEV)
        p_before = eng.flat_p.clone()
        loss = float(eng.step_fast(data, noise=noise))
        o_loss, o_res = tr.step(_data_cpu(data), bf, N(noise))
        assert int(eng.last_samples) == int(o_res["rm_samples"]) > 0          # the march is bit-exact
        assert abs(loss - o_loss) <= 2e-2 * abs(o_loss), (it, loss, o_loss)
        # gradients of this step: GPU flat gradient (loss-scaled) vs the oracle's autograd gradients
        scale = float(eng.scale)
        g_tab = N(eng.model.pos_encoder.hash_table.grad).reshape(-1) / scale
        o_tab = N(om.hash_table.grad).reshape(-1)
        assert np.abs(g_tab - o_tab).max() <= 5e-2 * np.abs(o_tab).max(), (it, np.abs(g_tab - o_tab).max(), np.abs(o_tab).max())
        g_w = N(eng._mlp_g[4]) / scale
        assert np.abs(g_w - N(om.W[4].grad)).max() <= 5e-2 * np.abs(N(om.W[4].grad)).max()
        assert float(eng.found_inf) == 0.0 and not torch.equal(p_before, eng.flat_p)
    assert eng.applied_steps() == 3
    w_g = N(eng.model.rgb_net.output_layer.weight)
    np.testing.assert_allclose(w_g, N(om.W[4]), rtol=0, atol=5e-3)


@pytest.mark.parametrize("fused", [True, False])
def test_fast_step_skips_on_overflow(fused):
    """GradScaler semantics through the native step runner: a step whose gradients are not finite leaves parameters and
    Adam state untouched, halves the scale and does not advance Adam's step count -- with the separate vn_grad_check
    pass (fused=False) and with the check evaluated inside the fused backward kernel (fused=True)"""
    from virus_nerf_b200 import synthetic
    from virus_nerf_b200.engine import TrainEngine
    args = synthetic.make_args(device=DEV, batch_size=256)
    ds = synthetic.SyntheticDataset(pool_size=1 << 13, n_images=8, device=DEV)
    eng = TrainEngine(args, ds, DEV, fused_scatter=fused)
    bf = synthetic.morton_pack(ds.scene.occupancy_bitfield(128))
    eng.model.occupancy_grid.bitfield = torch.from_numpy(bf).to(DEV)
    eng.step_idx = eng._prep_step = 1
    good = ds(256, args.training.sampling_strategy)
    eng.step_fast(good)
    assert eng.applied_steps() == 1 and float(eng.scale) == 2.0 ** 19
    p, m, v = eng.flat_p.clone(), eng.flat_m.clone(), eng.flat_v.clone()
    bad = {k: (val.clone() if torch.is_tensor(val) else {kk: vv.clone() for kk, vv in val.items()}) for k, val in good.items()
           if k in ("rays_o", "rays_d", "rgb", "depth")}
    bad["rgb"][3, 1] = float("inf")                       # -> infinite colour loss gradient
    eng.step_fast(bad)
    assert torch.equal(p, eng.flat_p) and torch.equal(m, eng.flat_m) and torch.equal(v, eng.flat_v)
    assert eng.applied_steps() == 1 and float(eng.scale) == 2.0 ** 18 and float(eng.found_inf) == 0.0
    eng.step_fast(good)
    assert eng.applied_steps() == 2 and not torch.equal(p, eng.flat_p) and torch.isfinite(eng.flat_p).all()


def test_occupancy_update_runs_and_is_deterministic(setup):
    eng, om, ds, bf, args, pipeline = setup
    og = eng.model.occupancy_grid
    torch.manual_seed(5); og.dataset.gen.manual_seed(5)
    g0 = og.occ_3d_grid.clone(); step0 = og.update_step
    og.update(elapse_time=0.0)
    g1, b1 = og.occ_3d_grid.clone(), og.getBitfield().clone()
    og.occ_3d_grid.copy_(g0); og.update_step = step0
    torch.manual_seed(5); og.dataset.gen.manual_seed(5)
    og.update(elapse_time=0.0)
    assert torch.equal(og.occ_3d_grid, g1) and torch.equal(og.getBitfield(), b1)   # replicas stay bit-identical
    assert og.getBitfield().shape == (128 ** 3 // 8,) and og.getBitfield().dtype == torch.uint8
    assert torch.isfinite(og.occ_3d_grid).all()
    assert not torch.equal(g0, g1)


def test_engine_converts_world_lengths_with_the_scene():
    """loss.py:28-30 and occupancy_grid.py:50-62: the 3 cm USS tolerance and the sensor-model parameters are world
    lengths; with a scene they reach the kernels in cube units (scene.w2c(only_scale=True))"""
    from types import SimpleNamespace
    from virus_nerf_b200 import synthetic
    from virus_nerf_b200.engine import TrainEngine
    k = 0.25
    scene = SimpleNamespace(w2c=lambda pos, only_scale=False, copy=True: pos * k)
    args = synthetic.make_args(device=DEV, batch_size=256)
    ds = synthetic.SyntheticDataset(pool_size=1 << 13, n_images=8, device=DEV)
    eng = TrainEngine(args, ds, DEV, scene=scene)
    og = eng.model.occupancy_grid
    assert abs(eng.loss_fn.uss_depth_tol - 0.03 * k) < 1e-12
    assert abs(eng._structs[0].uss_tol - 0.03 * k) < 1e-9
    assert abs(og.std_every_m - args.occ_grid.std_every_m * k) < 1e-12
    assert abs(og.false_detection_prob_every_m - args.occ_grid.false_detection_prob_every_m / k) < 1e-12
    plain = TrainEngine(args, ds, DEV)
    assert plain.loss_fn.uss_depth_tol == 0.03
    loss = float(eng.step_fast(ds(256, args.training.sampling_strategy)))
    assert np.isfinite(loss)


@pytest.mark.parametrize("half_opt", [False, True])
def test_occupancy_update_single_call_is_bit_identical(half_opt):
    """vn_occ_update (the whole OccupancyGrid.update after sampling as one C call, caller-owned workspace) produces the
    same grid and bitfield, bit for bit, as the reference-shaped sequence _rayUpdate / _nerfUpdate / decay + repack"""
    from virus_nerf_b200 import synthetic
    from virus_nerf_b200.engine import TrainEngine
    args = synthetic.make_args(device=DEV, batch_size=256)
    ds = synthetic.SyntheticDataset(pool_size=1 << 13, n_images=8, device=DEV)
    eng = TrainEngine(args, ds, DEV, half_opt=half_opt)
    og = eng.model.occupancy_grid
    if half_opt:
        with torch.no_grad():
            torch.manual_seed(4)
            eng.model.pos_encoder.hash_table.copy_(torch.rand_like(eng.model.pos_encoder.hash_table) * 0.2 - 0.1)
    assert og._native_model(torch.device(DEV)) is eng.model
    results = []
    g0, step0 = og.occ_3d_grid.clone(), og.update_step
    for native in (True, False, True):
        og.native_update = native
        og.occ_3d_grid.copy_(g0); og.update_step = step0
        torch.manual_seed(5); og.dataset.gen.manual_seed(5)
        with torch.autocast(device_type="cuda", dtype=torch.float16):
            for _ in range(3):                                      # three updates: decay steps and re-marked cells
                og.update(elapse_time=0.0)
        results.append((og.occ_3d_grid.clone(), og.getBitfield().clone()))
    assert not torch.equal(results[0][0], g0)
    for k in (1, 2):
        assert torch.equal(results[0][0], results[k][0]) and torch.equal(results[0][1], results[k][1])
    assert (og._winner == -1).all()


@pytest.mark.parametrize("enc_layout,single_pass,fused", [("chunks", True, True), ("chunks", True, False),
                                                          ("planar", True, False), ("rows", False, False)])
def test_fast_step_matches_autograd_step(monkeypatch, enc_layout, single_pass, fused):
    """the hand-chained C-ABI step (engine.step_fast) and the autograd step through the drop-in
    modules produce the same loss, gradients and parameter update (same rays, same jitter)"""
    from virus_nerf_b200 import synthetic
    from virus_nerf_b200.engine import TrainEngine
    from virus_nerf_b200.modules import ray_march, rendering
    args = synthetic.make_args(device=DEV, batch_size=512)
    ds = synthetic.SyntheticDataset(pool_size=1 << 13, n_images=8, device=DEV)
    e1 = TrainEngine(args, ds, DEV)
    e2 = TrainEngine(args, ds, DEV, enc_layout=enc_layout, single_pass_march=single_pass, fused_scatter=fused)
    assert e2.fused_scatter == fused
    assert torch.equal(e1.flat_p, e2.flat_p)
    e1.step_idx = e2.step_idx = 1                       # no occupancy update (it draws random numbers)
    e1._prep_step = e2._prep_step = 1
    bf = torch.from_numpy(synthetic.morton_pack(ds.scene.occupancy_bitfield(128))).to(DEV)
    e1.model.occupancy_grid.bitfield = bf; e2.model.occupancy_grid.bitfield = bf
    orig = ray_march.raymarching_train
    for it in range(2):
        data = ds(512, args.training.sampling_strategy)
        noise = torch.rand(512, device=DEV)
        monkeypatch.setattr(rendering, "raymarching_train", lambda *a, **k: orig(*a, noise=noise, **k))
        l1 = float(e1.step(data))
        l2 = float(e2.step_fast(data, noise=noise))
        assert e2._ticket is None
        assert abs(l1 - l2) <= 1e-5 * abs(l1), (l1, l2)
        g1, g2 = e1.flat_g, e2.flat_g
        assert float((g1 - g2).norm() / g1.norm()) < 1e-4
        assert float((e1.flat_p - e2.flat_p).abs().max()) < 1e-4
    assert int(e1.last_samples) == int(e2.last_samples) > 0


@pytest.mark.parametrize("fused", [True, False])
def test_fast_step_half_encoder_matches_autograd_step(monkeypatch, fused):
    """BASELINE config 3's half-precision encoder (hash_encoder_half.py) inside the native step runner: fp16 table
    copy per step, half forward kernel, fp16 encoding gradients into the half backward kernel (zero-skip, fp32
    accumulation into the flat gradient) == the same step through the drop-in half module and torch autograd"""
    from virus_nerf_b200 import synthetic
    from virus_nerf_b200.engine import TrainEngine
    from virus_nerf_b200.modules import ray_march, rendering
    args = synthetic.make_args(device=DEV, batch_size=512)
    ds = synthetic.SyntheticDataset(pool_size=1 << 13, n_images=8, device=DEV)
    e1 = TrainEngine(args, ds, DEV, half_opt=True)
    e2 = TrainEngine(args, ds, DEV, half_opt=True, fused_scatter=fused)
    assert type(e2.model.pos_encoder).__module__.endswith("hash_encoder_half") and e2._table_h is not None
    # the reference initialises the half table with U(-1e-4, 1e-4): give the encoding some signal
    with torch.no_grad():
        torch.manual_seed(4)
        t = (torch.rand_like(e1.model.pos_encoder.hash_table) * 2 - 1) * 0.5
        e1.model.pos_encoder.hash_table.copy_(t); e2.model.pos_encoder.hash_table.copy_(t)
    assert torch.equal(e1.flat_p, e2.flat_p)
    e1.step_idx = e2.step_idx = 1
    e1._prep_step = e2._prep_step = 1
    bf = torch.from_numpy(synthetic.morton_pack(ds.scene.occupancy_bitfield(128))).to(DEV)
    e1.model.occupancy_grid.bitfield = bf; e2.model.occupancy_grid.bitfield = bf
    orig = ray_march.raymarching_train
    for it in range(2):
        data = ds(512, args.training.sampling_strategy)
        noise = torch.rand(512, device=DEV)
        monkeypatch.setattr(rendering, "raymarching_train", lambda *a, **k: orig(*a, noise=noise, **k))
        l1 = float(e1.step(data))
        l2 = float(e2.step_fast(data, noise=noise))
        assert abs(l1 - l2) <= 1e-5 * abs(l1), (l1, l2)
        g1, g2 = e1.flat_g, e2.flat_g
        assert float(g1.norm()) > 0
        assert float((g1 - g2).norm() / g1.norm()) < 1e-4
        assert float((e1.flat_p - e2.flat_p).abs().max()) < 1e-4
    assert int(e1.last_samples) == int(e2.last_samples) > 0


def test_half_engine_rejects_other_layouts():
    from virus_nerf_b200 import synthetic
    from virus_nerf_b200.engine import TrainEngine
    args = synthetic.make_args(device=DEV, batch_size=64)
    ds = synthetic.SyntheticDataset(pool_size=1 << 10, n_images=4, device=DEV)
    with pytest.raises(ValueError, match="half_opt"):
        TrainEngine(args, ds, DEV, half_opt=True, enc_layout="rows")


def test_skipped_step_does_not_advance_adam(monkeypatch):
    """torch semantics of GradScaler.step + Adam: a step with a non-finite gradient is skipped, the scale backs off,
    and Adam's step count (bias corrections) does NOT advance -- checked against torch.optim.Adam + torch GradScaler
    arithmetic on the same gradients (the host-side counter advanced on skipped steps)"""
    from virus_nerf_b200 import _lib
    torch.manual_seed(0)
    n = 4096
    p = torch.randn(n, device=DEV); g = torch.zeros(n, device=DEV)
    m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
    scale = torch.tensor([2.0 ** 19], device=DEV); tracker = torch.zeros(1, dtype=torch.int32, device=DEV)
    found = torch.zeros(1, device=DEV); state = torch.zeros(4, device=DEV)
    lr, b1, b2, eps = 5e-3, 0.9, 0.999, 1e-15
    _lib.call("vn_opt_state_init", state, 0, lr, b1, b2)
    ref_p = p.clone().cpu().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=lr, betas=(b1, b2), eps=eps)
    ref_scale = 2.0 ** 19
    grads = [torch.randn(n), torch.randn(n), torch.randn(n), torch.randn(n)]
    grads[0][17] = float("inf"); grads[2][5] = float("nan")          # steps 1 and 3 overflow
    applied = 0
    for k, gk in enumerate(grads):
        g.copy_((gk * ref_scale).to(DEV))
        _lib.call("vn_grad_check", g, n, found)
        _lib.call("vn_adam_step_dev", p, g, m, v, n, lr, b1, b2, eps, state, found, scale)
        _lib.call("vn_scaler_update_dev", scale, tracker, found, 2.0, 0.5, 2000, state, lr, b1, b2)
        if torch.isfinite(gk).all():
            ref_p.grad = (gk * ref_scale) * (1.0 / ref_scale)
            opt.step(); applied += 1
        else:
            ref_scale *= 0.5
        assert float(scale) == ref_scale
        assert int(state[2:3].view(torch.int32)) == applied
    assert applied == 2 and opt.state[ref_p]["step"] == 2
    np.testing.assert_allclose(p.cpu().numpy(), ref_p.detach().numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(m.cpu().numpy(), opt.state[ref_p]["exp_avg"].numpy(), rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(v.cpu().numpy(), opt.state[ref_p]["exp_avg_sq"].numpy(), rtol=2e-6, atol=1e-12)


def test_pipelined_steps_equal_unpipelined():
    """step_fast(data, next_data=...) (front half of the next step enqueued early) is the same
    computation as calling step_fast(data) step by step"""
    from virus_nerf_b200 import synthetic
    from virus_nerf_b200.engine import TrainEngine
    args = synthetic.make_args(device=DEV, batch_size=512)
    ds = synthetic.SyntheticDataset(pool_size=1 << 13, n_images=8, device=DEV)
    batches = [ds(512, args.training.sampling_strategy) for _ in range(10)]
    outs = []
    for pipelined in (False, True):
        torch.manual_seed(3)
        eng = TrainEngine(args, ds, DEV)
        losses = []
        for i, b in enumerate(batches):
            nxt = batches[i + 1] if (pipelined and i + 1 < len(batches)) else None
            losses.append(float(eng.step_fast(b, next_data=nxt)))
        outs.append((losses, eng.flat_p.clone(), eng.model.occupancy_grid.getBitfield().clone(), eng.last_samples))
    (l0, p0, b0, s0), (l1, p1, b1, s1) = outs
    assert s0 == s1 and torch.equal(b0, b1)
    np.testing.assert_allclose(l0, l1, rtol=1e-4)
    # Adam normalises the gradient, so atomic-order noise in a near-zero gradient entry can move that
    # entry by +-lr per step: compare in the mean, not entry-wise
    d = (p0 - p1).abs()
    assert float(d.mean()) < 1e-5 and float((d > 1e-3).float().mean()) < 1e-3
