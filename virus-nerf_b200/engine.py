"""Train-step driver for the hot path: the inner loop of the reference's Trainer.train()
(training/trainer.py:87-165) with the Loss of training/loss.py, re-hosted on the sm_100a
kernels and made data-parallel (one process per GPU, ray batches sharded, hash-table and MLP
gradients sum-allreduced with NCCL).

What is deliberately different from the reference's loop (DESIGN.md "engine"):
  * all parameters live in ONE flat fp32 buffer (hash table | MLP weights) with one flat
    gradient buffer, so the optimiser is a single fused unscale + inf-check + Adam pass
    (vn_grad_check / vn_adam_step) and the DP exchange is a single allreduce;
  * the hash-encoder backward scatters straight into that gradient buffer;
  * no per-step .item() logging syncs (the reference does five, loss.py:76,96,124,169,197);
  * masked depth-loss means are normalised by GLOBAL valid counts under DP so that N ranks
    reproduce the single-process gradient of the concatenated batch.
"""
import math

import torch
import torch.distributed as dist

from . import _lib
from .modules.networks import NGP
from .modules.rendering import render


class Loss:
    """training/loss.py:34-198 without the logging syncs.  Returns the total loss; every masked
    mean takes its denominator from `counts` (local, or allreduced under DP)."""

    def __init__(self, args, uss_depth_tol=0.03):
        self.args = args
        self.uss_depth_tol = uss_depth_tol   # loss.py:29 (already in cube units here)

    def terms(self, results, data):
        """per-rank SUMS and COUNTS of every loss term (so they can be reduced across ranks)"""
        t = self.args.training
        out = {}
        diff = results['rgb'] - data['rgb']
        out['color'] = ((diff * diff).sum(), torch.tensor(float(diff.numel()), device=diff.device), t.color_loss_w)
        depth = results['depth']
        for sensor in t.sensors:
            meas = data['depth'][sensor]
            valid = ~torch.isnan(meas)
            if sensor == 'USS':                                         # loss.py:171-198
                valid = valid & (depth < meas - self.uss_depth_tol)
                w = t.uss_loss_w
            elif sensor == 'ToF':                                       # loss.py:127-145
                w = t.tof_loss_w
            else:                                                       # RGBD, loss.py:101-125
                w = t.rgbd_loss_w
            d = torch.where(valid, depth - torch.nan_to_num(meas), torch.zeros_like(depth))
            out[sensor] = ((d * d).sum(), valid.sum().to(torch.float32), w)
        return out

    def __call__(self, results, data, world_size=1):
        terms = self.terms(results, data)
        sums = torch.stack([v[0] for v in terms.values()])
        cnts = torch.stack([v[1] for v in terms.values()])
        ws = torch.tensor([v[2] for v in terms.values()], device=sums.device, dtype=torch.float32)
        if world_size > 1:
            cnts = cnts.clone()
            dist.all_reduce(cnts)                                       # global normalisers
        # mean = sum / count; empty masks contribute 0 (loss.py:140-141, 186-190)
        per_term = torch.where(cnts > 0, sums / cnts.clamp(min=1.0), torch.zeros_like(sums))
        return (per_term * ws).sum(), per_term.detach()


class TrainEngine:
    def __init__(self, args, dataset, device, world_size=1, rank=0, log2_T=19, max_res=1024, half_opt=False,
                 autocast=True, seed=21, grad_scale=2.0 ** 19):
        self.args = args
        self.device = torch.device(device)
        self.world_size, self.rank = world_size, rank
        self.dataset = dataset
        self.autocast = autocast
        torch.manual_seed(seed)          # identical replicas on every rank
        self.model = NGP(scale=args.model.scale, pos_encoder_type='hash', levels=args.model.hash_levels,
                         max_res=max_res, log2_T=log2_T, half_opt=half_opt, args=args, dataset=dataset)
        self.model.to(self.device)
        self.model.fused_mlp = self.model.fused_mlp and autocast   # fp32 mode (tests): torch.nn.Linear fp32
        self.loss_fn = Loss(args)
        self.step_idx = 0
        self.grid_update_interval = args.occ_grid.update_interval
        self.lr, self.betas, self.eps = args.training.lr, (0.9, 0.999), 1e-15   # trainer.py:53-57

        # ---- flat parameter / gradient / Adam state ------------------------------------
        params = [p for p in self.model.parameters() if p.requires_grad]
        sizes = [p.numel() for p in params]
        pad = lambda n: (n + 3) // 4 * 4                                 # keep every slice 16-byte aligned
        total = sum(pad(n) for n in sizes)
        self.flat_p = torch.zeros(total, device=self.device)
        self.flat_g = torch.zeros(total, device=self.device)
        self.flat_m = torch.zeros(total, device=self.device)
        self.flat_v = torch.zeros(total, device=self.device)
        off = 0
        self.slices = []
        for p, n in zip(params, sizes):
            self.flat_p[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + n].view_as(p)
            p.grad = self.flat_g[off:off + n].view_as(p)
            self.slices.append((off, n))
            off += pad(n)
        self.n_params = total
        enc = self.model.pos_encoder
        enc._direct_grad = enc.hash_table.grad                           # scatter straight into flat_g
        # GradScaler(2**19) state on device (trainer.py:49-50)
        self.scale = torch.tensor([grad_scale], device=self.device)
        self.growth_tracker = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.found_inf = torch.zeros(1, device=self.device)
        self.adam_step = 0
        self.last_samples = 0

    # ------------------------------------------------------------------------------------
    def occupancy_update(self, elapse_time=0.0):
        with torch.autocast(device_type='cuda', dtype=torch.float16, enabled=self.autocast):
            self.model.updateOccGrid(density_threshold=0.5, elapse_time=elapse_time)

    def forward_loss(self, data):
        with torch.autocast(device_type='cuda', dtype=torch.float16, enabled=self.autocast):
            results = render(self.model, data['rays_o'], data['rays_d'], exp_step_factor=self.args.exp_step_factor)
            loss, terms = self.loss_fn(results, data, self.world_size)
        return loss, terms, results

    def step(self, data, elapse_time=0.0):
        """one full train step: (grid update) -> render -> loss -> backward -> allreduce -> Adam"""
        if self.step_idx % self.grid_update_interval == 0:              # trainer.py:106-117
            self.occupancy_update(elapse_time)
        self.flat_g.zero_()                                             # optimizer.zero_grad()
        loss, terms, results = self.forward_loss(data)
        (loss * self.scale).sum().backward()                            # grad_scaler.scale(loss).backward()
        if self.world_size > 1:
            dist.all_reduce(self.flat_g)                                # sum: terms are globally normalised
        self.optimizer_step()
        self.step_idx += 1
        self.last_samples = results['rm_samples']
        return loss.detach()

    def optimizer_step(self):
        """grad_scaler.step(optimizer); grad_scaler.update() (trainer.py:140-141), fused"""
        self.adam_step += 1
        _lib.call("vn_grad_check", self.flat_g, self.n_params, self.found_inf)
        _lib.call("vn_adam_step", self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.n_params,
                  1.0, self.lr, self.betas[0], self.betas[1], self.eps, self.adam_step, self.found_inf, self.scale)
        _lib.call("vn_scaler_update", self.scale, self.growth_tracker, self.found_inf, 2.0, 0.5, 2000)
