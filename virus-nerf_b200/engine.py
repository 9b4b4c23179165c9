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
import os

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


class _Workspace:
    """persistent per-step sample buffers (grown geometrically, never shrunk): the fast step
    never touches the caching allocator in steady state"""

    def __init__(self, device):
        self.device = device
        self.bufs = {}
        self.retired = []     # outgrown buffers stay alive for a while: kernels of another stream may still use them

    def get(self, name, rows, cols=None, dtype=torch.float32):
        key = (name, cols, dtype)
        t = self.bufs.get(key)
        if t is None or t.shape[0] < rows:
            cap = max(int(rows * 1.25) + 1024, 4096)
            shape = (cap,) if cols is None else (cap, cols)
            if t is not None:
                self.retired.append(t)
                del self.retired[:-16]
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self.bufs[key] = t
        return t[:rows]


class TrainEngine:
    def __init__(self, args, dataset, device, world_size=1, rank=0, log2_T=19, max_res=1024, half_opt=False,
                 autocast=True, seed=21, grad_scale=2.0 ** 19, comm="auto", enc_layout="chunks",
                 single_pass_march=True, fused_scatter="auto", early_expand=True, scene=None):
        self.args = args
        self.device = torch.device(device)
        self.world_size, self.rank = world_size, rank
        self.dataset = dataset
        self.autocast = autocast
        # layout of the encoding / its gradient INSIDE the fast step (never visible to the drop-in modules):
        # "planar" = [8][S] float4 level-pair planes (coalesced on both sides), "rows" = the reference's [S,32]
        assert enc_layout in ("chunks", "planar", "rows")
        self.enc_planar = enc_layout in ("planar", "chunks")
        self.enc_chunks = enc_layout == "chunks" and autocast
        # MLP backward and hash backward as ONE kernel (vn_mlp_bwd_scatter): d(enc) goes from tensor memory straight into
        # the table gradient; only with the operand-chunk layout
        # (measured, profiles/r2_kbench.md: faster than the two kernels while the table is L2 resident -- 0.539 vs 0.565 ms
        # at 1.3 M samples, T = 2^19 -- and slower when it is not -- T = 2^22: 0.68 vs 0.63 ms: "auto" decides on the size)
        if fused_scatter == "auto":
            fused_scatter = (2 ** log2_T) * 2 * 4 * 16 <= (96 << 20)
        self.fused_scatter = bool(fused_scatter) and self.enc_chunks
        # the sample expansion of step k+1 (no dependence on the parameters) runs on the side stream under step k's
        # backward / optimiser / gradient exchange; the per-sample arrays it writes are double-buffered
        self.early_expand = bool(early_expand)
        # single-pass march of the fast step: the count pass stores t of every sample in a [N, 1024] scratch and
        # the write pass only expands it (bit-identical to re-marching); capped at 64 Ki rays (256 MB scratch x 2)
        self.single_pass_march = single_pass_march
        torch.manual_seed(seed)          # identical replicas on every rank
        self.model = NGP(scale=args.model.scale, pos_encoder_type='hash', levels=args.model.hash_levels,
                         max_res=max_res, log2_T=log2_T, half_opt=half_opt, args=args, dataset=dataset, scene=scene)
        self.model.to(self.device)
        # the occupancy update is REPLICATED: every rank runs the same update from the same
        # ray pool with an identically seeded sampler, so the grids stay bit-identical without
        # any broadcast; the training sampler may differ per rank.
        if hasattr(dataset, "clone_with_seed"):
            self.model.occupancy_grid.dataset = dataset.clone_with_seed(seed)
        self.model.fused_mlp = self.model.fused_mlp and autocast   # fp32 mode (tests): torch.nn.Linear fp32
        # loss.py:28-30: the 3 cm USS tolerance is a WORLD length; with a scene it is converted to cube units like the
        # occupancy grid's sensor-model parameters (occupancy_grid.py:50-62); without one the numbers are taken as cube units
        uss_tol = 0.03 if scene is None else float(scene.w2c(pos=0.03, only_scale=True, copy=True))
        self.loss_fn = Loss(args, uss_depth_tol=uss_tol)
        self.step_idx = 0
        self.grid_type = getattr(args.model, "grid_type", "occ")
        # trainer_base.py:85-88
        self.grid_update_interval = (args.ngp_grid if self.grid_type == "ngp" else args.occ_grid).update_interval
        self._grid_updates = 0
        self.lr, self.betas, self.eps = args.training.lr, (0.9, 0.999), 1e-15   # trainer.py:53-57
        self.half_opt = bool(half_opt)
        if self.half_opt and not (enc_layout == "chunks" and autocast):
            raise ValueError("TrainEngine: half_opt needs enc_layout='chunks' and autocast (the fp16 operand-chunk path)")

        # ---- flat parameter / gradient / Adam state ------------------------------------
        params = [p for p in self.model.parameters() if p.requires_grad]
        sizes = [p.numel() for p in params]
        pad = lambda n: (n + 3) // 4 * 4                                 # keep every slice 16-byte aligned
        total = sum(pad(n) for n in sizes)
        # data parallel with the fused exchange: parameters and gradients live in torch's SYMMETRIC memory when it is
        # available (peer pointers without CUDA IPC and, on an NVSwitch fabric, NVLink-multicast addresses: vn_p2p_step
        # then reduces and broadcasts through the switch).  VN_P2P_NVLS=0 keeps plain allocations + CUDA IPC.
        self._symm = None
        # (measured: at 2 ranks the multicast path is slower than peer loads / stores -- one peer, no traffic to save:
        # 0.150 vs 0.102 ms -- so "auto" takes it from 4 ranks on; VN_P2P_NVLS=1 / 0 forces it on / off)
        nvls_env = os.environ.get("VN_P2P_NVLS", "auto")
        want_symm = nvls_env == "1" or (nvls_env not in ("0", "1") and world_size >= 4)
        if world_size > 1 and comm in ("auto", "p2p_fused") and want_symm:
            try:
                fp, p_ptrs, mc_p, hp = _lib.symmetric_empty(total, self.device)
                fg, g_ptrs, mc_g, hg = _lib.symmetric_empty(total, self.device)
                self._symm = {"p_ptrs": p_ptrs, "g_ptrs": g_ptrs, "mc_p": mc_p, "mc_g": mc_g, "handles": (hp, hg)}
                self.flat_p, self.flat_g = fp, fg
            except Exception as e:           # no symmetric memory on this build / box: CUDA IPC below
                self._symm = None
                self._symm_warning = repr(e)
        if self._symm is None:
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
        self.adam_step = 0        # host-side count of optimiser CALLS (diagnostic only: the bias corrections follow opt_state)
        # Adam's state['step'] on the device: advanced only by steps that are applied (GradScaler skips on overflow)
        self.opt_state = torch.zeros(4, device=self.device)
        if self.device.type == "cuda":
            _lib.call("vn_opt_state_init", self.opt_state, 0, self.lr, self.betas[0], self.betas[1])
        # fp16 copy of the table, re-materialised every forward (hash_encoder_half.py:367), for the fast step
        self._table_h = (torch.empty(self.model.pos_encoder.hash_table.numel(), dtype=torch.float16, device=self.device)
                         if self.half_opt else None)
        self.last_samples = 0
        # fast-path state
        self._ws = _Workspace(self.device)
        self._counter = torch.zeros(2, dtype=torch.int32, device=self.device)
        cuda = self.device.type == "cuda"
        self._counters = [torch.zeros(2, dtype=torch.int32, device=self.device) for _ in range(2)]
        self._counter_hosts = [torch.zeros(2, dtype=torch.int32).pin_memory() for _ in range(2)] if cuda else None
        self._prep_stream = torch.cuda.Stream(device=self.device, priority=-1) if cuda else None
        self._ticket = None
        self._prep_step = 0
        self._loss_acc = torch.zeros(8, device=self.device)
        self._loss_out = torch.zeros(1, device=self.device)
        names = ["xyz_encoder.hidden_layers.0.weight", "xyz_encoder.output_layer.weight", "rgb_net.hidden_layers.0.weight",
                 "rgb_net.hidden_layers.1.weight", "rgb_net.output_layer.weight"]
        pd = dict(self.model.named_parameters())
        self._mlp_w = [pd[n].data for n in names]           # views into flat_p
        self._mlp_g = [pd[n].grad for n in names]           # views into flat_g
        self._comm_stream = torch.cuda.Stream(device=self.device) if world_size > 1 else None
        # gradient exchange: NCCL allreduce, or the library's own peer-memory allreduce (csrc/p2p_allreduce.cu;
        # standalone on 8 B200: 0.193 ms vs 0.209 ms NCCL for the 45.7 MB buffer; slower than NCCL at 2-4 ranks)
        #   "p2p_fused": sharded optimiser fused with the exchange (csrc/p2p_allreduce.cu, vn_p2p_reduce_adam):
        #   rank r reduces slice r of the gradient over peer memory, runs Adam on it with rank-local m / v and
        #   pushes the updated parameters into every replica; the loss normalisers go through the same mailboxes
        self._p2p = None
        self.comm = comm if world_size > 1 else "none"
        if world_size > 1 and comm in ("auto", "p2p", "p2p_fused"):
            try:
                self._p2p_flags = torch.zeros(world_size, dtype=torch.int32, device=self.device)
                self._p2p_err = torch.zeros(1, dtype=torch.int32, device=self.device)
                self._p2p_mbox = torch.zeros(2, world_size, 8, device=self.device)
                if self._symm is not None:
                    y = self._symm
                    self._p2p = _lib.p2p_setup_symmetric(y["g_ptrs"], y["p_ptrs"], y["mc_g"], y["mc_p"], self._p2p_flags,
                                                         self._p2p_err, self._p2p_mbox, rank, world_size)
                else:
                    self._p2p = _lib.p2p_setup(self.flat_g, self._p2p_flags, self._p2p_err, rank, world_size,
                                               params=self.flat_p, mbox=self._p2p_mbox)
                ok = torch.ones(1, device=self.device)
            except RuntimeError as e:                  # e.g. no peer access between the GPUs of this box
                if comm != "auto":
                    raise
                self._p2p, ok = None, torch.zeros(1, device=self.device)
                self._p2p_warning = str(e)
            if comm == "auto":                         # measured (profiles/r1_bench.md): fused beats NCCL + dense Adam
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # every rank must take the same path
                self.comm = "p2p_fused" if float(ok) == 1.0 else "nccl"
                if self.comm == "nccl":
                    self._p2p = None
        self._p2p_fused = self.comm == "p2p_fused"
        self.nvls = bool(self._p2p_fused and self._symm is not None and self._symm["mc_g"] and self._symm["mc_p"])
        self._err_host = torch.zeros(1, dtype=torch.int32).pin_memory() if (self._p2p is not None) else None
        self._hash_slice_idx = [i for i, p in enumerate(params) if p is enc.hash_table][0]
        self._structs = [self._new_step_struct(), self._new_step_struct()] if self.device.type == "cuda" else None

    # ------------------------------------------------------------------------------------
    def occupancy_update(self, elapse_time=0.0):
        """trainer.py:106-119"""
        step = self._grid_updates * self.grid_update_interval          # the train step this update belongs to
        self._grid_updates += 1
        with torch.autocast(device_type='cuda', dtype=torch.float16, enabled=self.autocast):
            if self.grid_type == "ngp":
                self.model.updateNeRFGrid(density_threshold=0.01 * 1024 / 3 ** 0.5,
                                          warmup=step < self.args.ngp_grid.warmup_steps)
            else:
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
        self._prep_step = self.step_idx + 1
        self.flat_g.zero_()                                             # optimizer.zero_grad()
        loss, terms, results = self.forward_loss(data)
        (loss * self.scale).sum().backward()                            # grad_scaler.scale(loss).backward()
        if self.world_size > 1:
            dist.all_reduce(self.flat_g)                                # sum: terms are globally normalised
        self.optimizer_step()
        self.step_idx += 1
        self.last_samples = results['rm_samples']
        return loss.detach()

    # ------------------------------------------------------------------------------------
    def _new_step_struct(self):
        """vn_step_t with everything that does not change from step to step"""
        m, a = self.model, self.args
        enc = m.pos_encoder
        st = _lib.Step()
        st.cascades, st.grid_size, st.max_samples = m.cascades, m.grid_size, 1024
        st.scale, st.exp_step_factor, st.T_threshold = float(m.scale), float(a.exp_step_factor), 1e-4
        st.bg = 1.0 if a.exp_step_factor == 0 else 0.0                   # rendering.py:219-224
        st.uss_tol = float(self.loss_fn.uss_depth_tol)
        st.set_ptrs(flat_p=self.flat_p, flat_g=self.flat_g, flat_m=self.flat_m, flat_v=self.flat_v,
                    loss_acc=self._loss_acc, loss_out=self._loss_out, scale_dev=self.scale, found_inf=self.found_inf,
                    growth_tracker=self.growth_tracker)
        st.n_params = self.n_params
        esz = self.flat_p.element_size()
        st.table_off = (enc.hash_table.data_ptr() - self.flat_p.data_ptr()) // esz
        for i, w in enumerate(self._mlp_w):
            st.w_off[i] = (w.data_ptr() - self.flat_p.data_ptr()) // esz
        st.levels = enc._levels
        st.hash_flags = enc.kernel_flags
        if self.enc_planar:
            # measured on B200 (profiles/r1_kbench.md): planes + 2 levels per thread + 16-byte pair loads (fwd)
            # + the 48-register backward: fwd 0.240 -> 0.197 ms, bwd 0.417 -> 0.378 ms at 1.3 M samples
            # + no scatter of exactly-zero gradients (41 % of the samples after 200 steps: tools/zero_frac.py)
            st.hash_flags |= (_lib.VN_HASH_PLANAR | _lib.VN_HASH_LEVEL_GROUPS_2 | _lib.VN_HASH_PAIR_LOADS |
                              _lib.VN_HASH_TIGHT_REGS | _lib.VN_HASH_SKIP_ZERO_GRADS)
            # round 2: the forward encoding leaves the hash kernel as fp16 tensor-core operand chunks (rounded
            # where autocast rounds the Linear input) and the MLP kernels pull tiles in with bulk async copies
            if self.enc_chunks:
                st.hash_flags |= _lib.VN_HASH_F16_CHUNKS
            if self.fused_scatter:
                st.hash_flags |= _lib.VN_HASH_FUSED_SCATTER
        t = a.training
        st.w_color, st.w_uss, st.w_tof, st.w_rgbd = t.color_loss_w, t.uss_loss_w, t.tof_loss_w, t.rgbd_loss_w
        st.lr, st.beta1, st.beta2, st.eps = self.lr, self.betas[0], self.betas[1], self.eps
        st.set_ptrs(step_dev=self.opt_state, table_h=self._table_h)
        return st

    def prepare(self, data, elapse_time=0.0, noise=None, ready=None):
        """front half of a step, independent of the gradients in flight: (occupancy update if
        due) -> ray/AABB -> march count + scan (vn_train_step_prepare) -> async read-back of the
        sample total into pinned memory.  step_fast() enqueues it for the NEXT batch while the
        current step's backward / allreduce is still running, so the one host sync of a step
        never stalls."""
        m = self.model
        update = self._prep_step % self.grid_update_interval == 0
        self._prep_step += 1
        par = self._prep_step & 1                       # per-ray buffers are double-buffered
        ws = self._ws
        main = torch.cuda.current_stream()
        # the front half only needs the bitfield: run it on a high-priority side stream so that it
        # overlaps the previous step's kernels and S is on the host before that step has finished.
        # When the occupancy update is due it depends on the weights -> stay on the main stream.
        side = self._prep_stream if (self._prep_stream is not None and not update) else None
        rays_o, rays_d = data['rays_o'].contiguous(), data['rays_d'].contiguous()
        N = rays_o.shape[0]
        t = self.args.training
        depth = data['depth']
        st = self._structs[par]
        st.N = N
        # every tensor whose pointer goes into the struct stays alive in the ticket (a .contiguous() copy made on the side
        # stream would otherwise be freed before the main-stream kernels read it) and must be float32
        gt_rgb = data['rgb'].contiguous()
        dep = {k: (depth[k].contiguous() if (k in t.sensors and depth.get(k) is not None) else None) for k in ('USS', 'ToF', 'RGBD')}
        for name, ten in [('rays_o', rays_o), ('rays_d', rays_d), ('rgb', gt_rgb)] + [(k, v) for k, v in dep.items() if v is not None]:
            if ten.dtype != torch.float32:
                raise TypeError(f"TrainEngine.prepare: {name} must be float32, got {ten.dtype}")
        keep = [rays_o, rays_d]

        def body():
            nz = noise
            if update:
                self.occupancy_update(elapse_time)
            if nz is None:
                nz = ws.get(f"noise{par}", N)
                nz.uniform_()                                               # ray_march.py:139
            nz = nz.contiguous()
            keep.extend([nz, gt_rgb] + [d for d in dep.values() if d is not None])
            st.set_ptrs(rays_o=rays_o, rays_d=rays_d, noise=nz, gt_rgb=gt_rgb,
                        uss=dep['USS'], tof=dep['ToF'], rgbd=dep['RGBD'],
                        hits_t=ws.get(f"hits{par}", N, 2), counts=ws.get(f"counts{par}", N, None, torch.int32),
                        rays_a=ws.get(f"rays_a{par}", N, 3, torch.int32),
                        scan_tmp=ws.get(f"scan_tmp{par}", _lib.scan_tmp_ints(N), None, torch.int32),
                        counter=self._counters[par], bitfield=m.occupancy_grid.getBitfield(),
                        ts_rows=ws.get(f"ts_rows{par}", N * 1024) if (self.single_pass_march and N <= 65536) else None,
                        vr_samples=ws.get("vr", N, None, torch.int32), opacity=ws.get("op", N), depth=ws.get("dp", N),
                        rgb=ws.get("rgb", N, 3), d_rgb=ws.get("d_rgb", N, 3), d_depth=ws.get("d_dp", N),
                        d_opacity=ws.get("d_op", N))
            _lib.call("vn_train_step_prepare", st)
            self._counter_hosts[par].copy_(self._counters[par], non_blocking=True)
            if self._err_host is not None:       # time-out word of the peer-memory exchange rides on the same host sync
                self._err_host.copy_(self._p2p_err, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            return ev, nz

        if side is None:
            ev, nz = body()
        else:
            if ready is None:                                               # the batch itself (H2D copy / gather)
                ready = torch.cuda.Event(); ready.record(main)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                ev, nz = body()
        return {"data": data, "struct": st, "event": ev, "par": par, "side": side is not None, "keep": keep}

    def step_fast(self, data, elapse_time=0.0, noise=None, next_data=None, next_noise=None):
        """The same train step as step(), enqueued by the native step runner (csrc/step.cu): no
        autograd graph, no torch glue, persistent workspace, three host calls per step:
        [prepare: aabb, march count + scan] -> host reads S -> [run: march write (also emits
        unit-cube positions), hash fwd, fused MLP fwd, composite fwd, loss fwd, (allreduce
        counts), loss bwd (gradient seeds), composite bwd, fused MLP bwd (dW straight into the
        flat gradient), hash bwd (ditto)] -> (allreduce flat gradient || prepare(next_data)) ->
        [optim: grad check, Adam, scaler update].  Pass next_data to software-pipeline the front
        half of the next step."""
        ready_next = None
        if next_data is not None:
            # next_data was produced (gather / H2D copy) on this stream before this call: everything the
            # next front half may depend on is older than this point, so it can overlap this step's kernels
            ready_next = torch.cuda.Event(); ready_next.record()
        tk = self._ticket if (self._ticket is not None and self._ticket["data"] is data) else \
            self.prepare(data, elapse_time, noise)
        self._ticket = None
        st, ws = tk["struct"], self._ws
        tk["event"].synchronize()                                         # the one host sync of the step
        if tk["side"]:
            torch.cuda.current_stream().wait_event(tk["event"])           # main stream: front half done
        S = int(self._counter_hosts[tk["par"]][0])
        if self._err_host is not None and int(self._err_host[0]) != 0:
            raise RuntimeError("TrainEngine: a cross-rank wait of the peer-memory exchange timed out (VN_P2P_TIMEOUT_MS): "
                               "the replicas are no longer synchronised -- restart from the last checkpoint")
        self.last_samples = S
        par = tk["par"]
        enc_cols = 24 if self.enc_chunks else 32          # operand chunks: [4 hash + 2 SH planes][S] x 16 B
        st.set_ptrs(xyzs=ws.get(f"xyzs{par}", S, 3), dirs=ws.get(f"dirs{par}", S, 3), unit=ws.get(f"unit{par}", S, 3),
                    deltas=ws.get(f"deltas{par}", S), ts=ws.get(f"ts{par}", S), enc=ws.get(f"enc{par}", S, enc_cols),
                    sigmas=ws.get("sig", S), rgbs=ws.get("rgbs", S, 3), ws=ws.get("ws", S), d_sigmas=ws.get("d_sig", S),
                    d_rgbs=ws.get("d_rgbs", S, 3), d_enc=None if self.fused_scatter else ws.get("d_enc", S, 32))
        skip_expand = 0
        if self.early_expand and tk["side"]:
            # front half on the side stream => no occupancy update in between: the expansion can follow it there, under
            # the previous step's tail (its buffers have the other parity); the main stream picks it up before the encoder
            with torch.cuda.stream(self._prep_stream):
                _lib.call("vn_train_step_expand", st, S)
                expanded = torch.cuda.Event(); expanded.record()
            torch.cuda.current_stream().wait_event(expanded)
            skip_expand = _lib.VN_STEP_SKIP_EXPAND
        self.adam_step += 1
        st.adam_step = self.adam_step
        self.step_idx += 1
        update_due = self._prep_step % self.grid_update_interval == 0     # next front half needs the new weights
        if self.world_size == 1:
            _lib.call("vn_train_step_run", st, S, 0 | skip_expand, 1)
            if next_data is not None:
                self._ticket = self.prepare(next_data, elapse_time, noise=next_noise, ready=None if update_due else ready_next)
            return self._loss_out[0]
        # ---- data parallel ----------------------------------------------------------------
        _lib.call("vn_train_step_run", st, S, 1 | skip_expand, 0)
        if self._p2p_fused:                                               # global normalisers
            _lib.call("vn_p2p_allreduce_small", self._loss_acc[4:], 4, 0)
        else:
            dist.all_reduce(self._loss_acc[4:])
        _lib.call("vn_train_step_run", st, S, 2, 0)
        if self._p2p_fused:
            # reduce-scatter + inf check + sharded Adam + parameter push + scaler update, on the main stream: the next
            # step's main-stream work needs its result anyway, only the front half of the next step (side stream) overlaps
            # ONE kernel (csrc/p2p_allreduce.cu, vn_p2p_step); its inf flag is this rank's own check: evaluated inside the
            # fused backward kernel, else by a pass over the local gradient
            if not self.fused_scatter:
                _lib.call("vn_grad_check", self.flat_g, self.n_params, self.found_inf)
            _lib.call("vn_p2p_step", self.n_params, self.flat_m, self.flat_v, self.lr, self.betas[0], self.betas[1],
                      self.eps, self.opt_state, self.found_inf, self.scale, self.growth_tracker)
            if next_data is not None and not update_due:
                self._ticket = self.prepare(next_data, elapse_time, noise=next_noise, ready=ready_next)
        else:
            ev = torch.cuda.Event(); ev.record()
            self._comm_stream.wait_event(ev)
            with torch.cuda.stream(self._comm_stream):
                if self._p2p is not None:                                 # own two-shot allreduce over NVLink peer memory
                    _lib.call("vn_p2p_allreduce", self.n_params)
                    work = None
                    done = torch.cuda.Event(); done.record()
                else:
                    work = dist.all_reduce(self.flat_g, async_op=True)    # NCCL; overlaps prepare(next) below
            if next_data is not None and not update_due:
                self._ticket = self.prepare(next_data, elapse_time, noise=next_noise, ready=ready_next)
            if work is not None:
                work.wait()
            else:
                torch.cuda.current_stream().wait_event(done)
            _lib.call("vn_train_step_optim", st)
        if next_data is not None and update_due:
            self._ticket = self.prepare(next_data, elapse_time, noise=next_noise)
        return self._loss_out[0]

    @torch.no_grad()
    def render_frame(self, rays_o, rays_d, chunk=1 << 20):
        """Test-time render of a full frame, rays sharded over the ranks in contiguous bands with
        no collective: this rank renders rays [rank*n/world, (rank+1)*n/world)
        and returns (lo, hi, results) for its band."""
        n = rays_o.shape[0]
        lo = (n * self.rank) // self.world_size
        hi = (n * (self.rank + 1)) // self.world_size
        parts = []
        for s in range(lo, hi, chunk):
            e_ = min(s + chunk, hi)
            parts.append(render(self.model, rays_o[s:e_], rays_d[s:e_], test_time=True,
                                exp_step_factor=self.args.exp_step_factor))
        out = {k: torch.cat([p[k] for p in parts]) for k in ("opacity", "depth", "rgb")} if parts else {}
        return lo, hi, out

    def replica_checksum(self):
        """(bitfield, parameter) checksums used to verify that DP replicas are bit-identical"""
        bf = self.model.occupancy_grid.getBitfield().to(torch.int64)
        w = torch.arange(1, bf.numel() + 1, device=bf.device, dtype=torch.int64)
        return torch.stack([(bf * w).sum(), self.flat_p.view(torch.int32).to(torch.int64).sum()])

    def optimizer_step(self):
        """grad_scaler.step(optimizer); grad_scaler.update() (trainer.py:140-141), fused"""
        self.adam_step += 1
        _lib.call("vn_grad_check", self.flat_g, self.n_params, self.found_inf)
        _lib.call("vn_adam_step_dev", self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.n_params,
                  self.lr, self.betas[0], self.betas[1], self.eps, self.opt_state, self.found_inf, self.scale)
        _lib.call("vn_scaler_update_dev", self.scale, self.growth_tracker, self.found_inf, 2.0, 0.5, 2000,
                  self.opt_state, self.lr, self.betas[0], self.betas[1])

    def close(self):
        """release the process-wide peer-memory exchange (one engine may own it at a time)"""
        if getattr(self, "_p2p", None) is not None:
            torch.cuda.synchronize(self.device)
            _lib.p2p_shutdown()
            self._p2p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def applied_steps(self):
        """number of optimiser steps that were applied (torch Adam's state['step']); synchronises"""
        return int(self.opt_state[2:3].view(torch.int32).item())
