"""CPU restatement of the reference's train / render drivers on top of oracle.cpp
(TEST INFRASTRUCTURE + the reported CPU baseline; never imported by the product package).

This is what the reference does on a CUDA-less host (training/trainer_base.py:37 selects
ti.cpu, args/args.py:69 torch.device("cpu"), autocast / GradScaler disable themselves): the
Taichi kernels are the C++ restatements of oracle.cpp (OpenMP when threads > 1), and the
MLPs, loss and Adam are PyTorch-CPU fp32 exactly as in modules/networks.py:195-282,
training/loss.py and training/trainer.py:53-57,138-141.
"""
import numpy as np
import torch

import oracle

MAX_SAMPLES = 1024


class _HashFn(torch.autograd.Function):
    """modules/hash_encoder.py:237-277"""

    @staticmethod
    def forward(ctx, x, table, lv, threads):
        ctx.save_for_backward(x)
        ctx.lv, ctx.threads = lv, threads
        return torch.from_numpy(oracle.hash_fwd_f32(x.numpy(), table.detach().numpy(), lv, threads))

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        g = oracle.hash_bwd_f32(x.numpy(), dout.contiguous().numpy(), ctx.lv, ctx.threads)
        return None, torch.from_numpy(g), None, None


class _CompositeFn(torch.autograd.Function):
    """modules/volume_train.py:58-175"""

    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, ts, rays_a, T_thr, threads):
        total, op, dp, rgb, ws = oracle.composite_train_fwd(sigmas.detach().numpy(), rgbs.detach().numpy(),
                                                            deltas.numpy(), ts.numpy(), rays_a.numpy(), T_thr, threads)
        ctx.save_for_backward(sigmas, rgbs, deltas, ts, rays_a)
        ctx.T_thr, ctx.threads = T_thr, threads
        return torch.from_numpy(op), torch.from_numpy(dp), torch.from_numpy(rgb), torch.from_numpy(ws), int(total.sum())

    @staticmethod
    def backward(ctx, dO, dD, dC, dW, _):
        sigmas, rgbs, deltas, ts, rays_a = ctx.saved_tensors
        n = rays_a.shape[0]
        z = lambda g, shape: np.zeros(shape, np.float32) if g is None else g.contiguous().numpy()
        ds, dc = oracle.composite_train_bwd(sigmas.detach().numpy(), rgbs.detach().numpy(), deltas.numpy(), ts.numpy(),
                                            rays_a.numpy(), ctx.T_thr, z(dO, n), z(dD, n), z(dC, (n, 3)),
                                            None if dW is None else dW.contiguous().numpy(), ctx.threads)
        return torch.from_numpy(ds), torch.from_numpy(dc), None, None, None, None, None


class OracleNGP:
    """NGP (modules/networks.py:32-164) with weights shared from / comparable to the CUDA model"""

    def __init__(self, scale=0.5, log2_T=19, max_res=1024, levels=16, threads=1, seed=21):
        self.scale, self.threads = scale, threads
        self.lv = oracle.HashLevels(16, max_res, levels, 2 ** log2_T)
        g = torch.Generator().manual_seed(seed)
        self.hash_table = torch.rand(2 * self.lv.total, generator=g).requires_grad_(True)   # hash_encoder.py:227
        xav = lambda o, i: ((torch.rand(o, i, generator=g) * 2 - 1) * np.sqrt(6.0 / (i + o))).requires_grad_(True)
        self.W = [xav(64, 32), xav(16, 64), xav(64, 32), xav(64, 64), xav(3, 64)]
        self.grid_size, self.cascades = 128, max(1 + int(np.ceil(np.log2(2 * scale))), 1)

    def load_from(self, model):
        """copy parameters of a virus_nerf_b200 NGP module"""
        sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
        self.hash_table = sd["pos_encoder.hash_table"].reshape(-1).clone().requires_grad_(True)
        names = ["xyz_encoder.hidden_layers.0.weight", "xyz_encoder.output_layer.weight",
                 "rgb_net.hidden_layers.0.weight", "rgb_net.hidden_layers.1.weight", "rgb_net.output_layer.weight"]
        self.W = [sd[n].clone().requires_grad_(True) for n in names]

    def parameters(self):
        return [self.hash_table] + self.W

    def density(self, x, return_feat=False):
        x = (x + self.scale) / (2 * self.scale)                                     # networks.py:142
        enc = _HashFn.apply(x.contiguous(), self.hash_table, self.lv, self.threads)
        h = torch.relu(enc @ self.W[0].t()) @ self.W[1].t()
        h0 = h[:, 0]
        sig = _TruncExp.apply(h0)
        return (sig, h) if return_feat else sig

    def forward(self, x, d):
        sig, h = self.density(x, True)
        d = d / torch.norm(d, dim=1, keepdim=True)
        sh = torch.from_numpy(oracle.sh_encode(((d + 1) / 2).contiguous().numpy()))  # networks.py:160-161
        z = torch.cat([sh, h], 1)
        rgb = torch.sigmoid(torch.relu(torch.relu(z @ self.W[2].t()) @ self.W[3].t()) @ self.W[4].t())
        return sig, rgb


class _TruncExp(torch.autograd.Function):
    """networks.py:17-29"""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        return g * torch.exp(ctx.saved_tensors[0].clamp(-15, 15))


def render_train(model, rays_o, rays_d, bitfield, noise, esf=0.0, T_thr=1e-4):
    """modules/rendering.py:12-57,161-228 (train path)"""
    ro, rd = rays_o.numpy(), rays_d.numpy()
    hits = oracle.ray_aabb(ro, rd, model.scale)
    rays_a, xyzs, dirs, deltas, ts, total = oracle.march_train(ro, rd, hits, bitfield, noise, model.cascades,
                                                               model.scale, esf, model.grid_size, MAX_SAMPLES,
                                                               model.threads)
    sig, rgbs = model.forward(torch.from_numpy(xyzs), torch.from_numpy(dirs))
    op, dp, rgb, ws, vr = _CompositeFn.apply(sig, rgbs, torch.from_numpy(deltas), torch.from_numpy(ts),
                                             torch.from_numpy(rays_a), T_thr, model.threads)
    bg = 1.0 if esf == 0 else 0.0
    rgb = rgb + bg * (1 - op)[:, None]
    return {"opacity": op, "depth": dp, "rgb": rgb, "ws": ws, "rm_samples": total, "vr_samples": vr,
            "rays_a": rays_a, "ts": ts, "deltas": deltas}


def render_test(model, rays_o, rays_d, bitfield, esf=0.0, T_thr=1e-4, max_samples=MAX_SAMPLES):
    """modules/rendering.py:61-158 (test path)"""
    ro, rd = np.ascontiguousarray(rays_o.numpy()), np.ascontiguousarray(rays_d.numpy())
    n = ro.shape[0]
    hits = oracle.ray_aabb(ro, rd, model.scale)
    opacity = np.zeros(n, np.float32); depth = np.zeros(n, np.float32); rgb = np.zeros((n, 3), np.float32)
    alive = np.arange(n, dtype=np.int64)
    samples = total = 0
    min_samples = 1 if esf == 0 else 4
    with torch.no_grad():
        while samples < max_samples:
            if alive.shape[0] == 0:
                break
            ns = max(min(n // alive.shape[0], 64), min_samples)
            samples += ns
            pk, ri, de, ts = oracle.march_test(ro, rd, hits, alive, bitfield, model.cascades, model.scale, esf,
                                               model.grid_size, ns)
            if ri.shape[0] == 0:
                break
            xyzs = ro[ri] + ts[:, None] * rd[ri]
            sig, rgbs = model.forward(torch.from_numpy(xyzs), torch.from_numpy(rd[ri]))
            oracle.composite_test(sig.numpy(), rgbs.numpy(), de, ts, pk, alive, T_thr, opacity, depth, rgb)
            alive = np.ascontiguousarray(alive[alive >= 0])
            total += int(pk[:, 1].sum())
    bg = 1.0 if esf == 0 else 0.0
    rgb = rgb + bg * (1 - opacity)[:, None]
    return {"opacity": opacity, "depth": depth, "rgb": rgb, "total_samples": total}


def loss_fn(results, data, sensors=("USS", "ToF"), w=None, uss_tol=0.03):
    """training/loss.py:34-198"""
    w = w or {"color": 1.0, "ToF": 50.0, "USS": 50.0, "RGBD": 100.0}
    loss = w["color"] * torch.nn.functional.mse_loss(results["rgb"], data["rgb"])
    depth = results["depth"]
    for s in sensors:
        meas = data["depth"][s]
        valid = ~torch.isnan(meas)
        if s == "USS":
            valid = valid & (depth < meas - uss_tol)
        if valid.any():
            loss = loss + w[s] * torch.mean((depth[valid] - meas[valid]) ** 2)
    return loss


class OracleTrainer:
    """training/trainer.py:87-165 inner loop on the CPU (Adam eps=1e-15, no GradScaler on CPU)"""

    def __init__(self, model, lr=5e-3):
        self.model = model
        self.opt = torch.optim.Adam(model.parameters(), lr, eps=1e-15)

    def step(self, data, bitfield, noise, sensors=("USS", "ToF")):
        res = render_train(self.model, data["rays_o"], data["rays_d"], bitfield, noise)
        loss = loss_fn(res, data, sensors)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.detach()), res


class OracleOccupancyGrid:
    """modules/occupancy_grid.py:12-105 on the CPU (torch.rand streams replaced by a numpy
    Generator; same update order: depth update, NeRF update, decay, bitfield)"""

    def __init__(self, model, dataset, args, seed=21):
        self.model, self.dataset, self.args = model, dataset, args
        self.G, self.M, self.I = 128, 32, 32
        self.rng = np.random.default_rng(seed)
        self.grid = (0.5 + 0.01 * self.rng.random(self.G ** 3, dtype=np.float32)).reshape(self.G, self.G, self.G)
        og = args.occ_grid
        decay = (0.5 / 0.51) ** (1 / (og.decay_warmup_steps / og.update_interval))
        self.grid_decay = ((decay * 1000) // 1) / 1000
        self.update_step = 0
        self.bitfield = oracle.occ_decay_pack(self.grid, 1.0, False, 0.5)

    def _batch(self, B, pixs, sensor):
        data = self.dataset(B, {"imgs": "all", "pixs": pixs})
        meas = data["depth"][sensor]
        ok = ~torch.isnan(meas)
        return data["rays_o"][ok].numpy(), data["rays_d"][ok].numpy(), meas[ok].numpy()

    def update(self):
        og = self.args.occ_grid
        B_ray = int(og.batch_size * og.batch_ratio_ray_update)
        ro, rd, meas = self._batch(B_ray, "valid_tof", "ToF")
        if ro.shape[0]:
            dists, _, idx = oracle.occ_calc_pos(ro, rd, None, self.M, self.G, self.model.scale, og.nerf_pos_noise_every_m)
            po, pe = oracle.occ_ray_prob(meas, dists, og.false_detection_prob_every_m, og.std_every_m, self.I)
            oracle.occ_update_grid(self.grid, idx, po, pe)
        ro, rd, meas = self._batch(og.batch_size - B_ray, "valid_uss", "USS")
        if ro.shape[0]:
            noise = self.rng.random((ro.shape[0], self.M, 3), dtype=np.float32)
            _, pos, idx = oracle.occ_calc_pos(ro, rd, noise, self.M, self.G, self.model.scale, og.nerf_pos_noise_every_m)
            with torch.no_grad():
                rho = self.model.density(torch.from_numpy(pos)).numpy()
            po, pe = oracle.occ_nerf_prob(rho, og.nerf_threshold_max, og.nerf_threshold_slope)
            oracle.occ_update_grid(self.grid, idx, po, pe)
        self.update_step += 1
        self.bitfield = oracle.occ_decay_pack(self.grid, self.grid_decay, self.update_step <= og.decay_warmup_steps, 0.5)
