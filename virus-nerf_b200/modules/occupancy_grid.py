"""Drop-in for the reference's modules/occupancy_grid.py: the VIRUS-NeRF occupancy grid
(Bayesian per-cell update from depth sensors and from NeRF density).  The ~60 small torch
launches per update of the reference are 5 C-ABI kernels here; there is no host sync (the
reference's torch.mean(...).item(), occupancy_grid.py:402, is computed on device)."""
import numpy as np
import torch

from .. import _lib
from .grid import Grid


class OccupancyGrid(Grid):
    def __init__(self, args, grid_size: int, scene=None, dataset=None, fct_density: callable = None) -> None:
        self.args = args
        self.grid_size = grid_size
        self.dataset = dataset
        self.fct_density = fct_density
        self.cascades = max(1 + int(np.ceil(np.log2(2 * self.args.model.scale))), 1)   # :26

        super().__init__(args=args, grid_size=grid_size, cascades=self.cascades, morton_structure=False)

        self.update_step = 0

        # initialize occupancy grid, :36-42
        self.threshold = 0.5
        occ_init_max = 0.51
        grid = torch.rand(size=(self.grid_size ** 3,), device=self.args.device, dtype=torch.float32)
        grid = self.threshold + (occ_init_max - self.threshold) * grid
        self.occ_3d_grid = grid.reshape(self.grid_size, self.grid_size, self.grid_size).contiguous()

        # fixed parameters, :45-47
        self.I = 32
        self.M = 32
        self.prob_min = 0.03

        # variable parameters, :50-62
        decay_num_steps = self.args.occ_grid.decay_warmup_steps / self.args.occ_grid.update_interval
        grid_decay = (self.threshold / occ_init_max) ** (1 / decay_num_steps)
        self.grid_decay = ((grid_decay * 1000) // 1) / 1000
        self.cell_size = 2 * self.args.model.scale / grid_size

        if scene is not None:
            self.false_detection_prob_every_m = self.args.occ_grid.false_detection_prob_every_m / scene.w2c(pos=1, only_scale=True, copy=False)
            self.std_every_m = scene.w2c(pos=self.args.occ_grid.std_every_m, only_scale=True, copy=False)
            self.nerf_pos_noise_every_m = scene.w2c(pos=self.args.occ_grid.nerf_pos_noise_every_m, only_scale=True, copy=False)
        else:
            self.false_detection_prob_every_m = self.args.occ_grid.false_detection_prob_every_m
            self.std_every_m = self.args.occ_grid.std_every_m
            self.nerf_pos_noise_every_m = self.args.occ_grid.nerf_pos_noise_every_m

        # kernel scratch (caller-owned, see include/virusnerf.h a14)
        self._winner = None
        self._scratch = None
        self._ws = None              # workspace of vn_occ_update
        self._table_h = None
        self.native_update = True    # False: always the step-by-step path (one C-ABI call per reference method)

    def _buffers(self, device):
        if self._winner is None or self._winner.device != device:
            self._winner = torch.full((self.grid_size ** 3,), -1, dtype=torch.int32, device=device)
            self._scratch = torch.zeros(1024, dtype=torch.float32, device=device)
        return self._winner, self._scratch

    def _native_model(self, device):
        """the NGP whose density this grid queries, when NGP.density runs as hash forward + fused density MLP (then the
        whole update can be enqueued by ONE C call, vn_occ_update); None -> the step-by-step path below"""
        if not self.native_update or device.type != "cuda":
            return None
        m = getattr(self.fct_density, "__self__", None)
        if m is None or not hasattr(m, "pos_encoder") or not getattr(m, "fused_mlp", False):
            return None
        if m.pos_encoder.out_dim != 32 or not m.pos_encoder.hash_table.is_cuda:
            return None
        return m

    @torch.no_grad()
    def _updateNative(self, m, ray_update: dict, nerf_update: dict, apply_decay: bool):
        """:65-105 from the sampled batches on: _rayUpdate, _nerfUpdate, decay and repack through vn_occ_update"""
        grid = self.occ_3d_grid
        dev = grid.device
        f = lambda t: t.contiguous().float()
        n_ray, n_nerf = int(ray_update["batch_size"]), int(nerf_update["batch_size"])
        n_ray = ray_update["rays_o"].shape[0] if n_ray > 0 else 0
        n_nerf = nerf_update["rays_o"].shape[0] if n_nerf > 0 else 0
        need = _lib.lib().vn_occ_update_ws_floats(n_ray, n_nerf, self.M)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(int(need * 1.25) + 1024, dtype=torch.float32, device=dev)
        winner, _ = self._buffers(dev)
        bf = self.bitfield
        if bf.numel() != self.grid_size ** 3 // 8 or bf.device != dev:
            bf = torch.zeros(self.grid_size ** 3 // 8, dtype=torch.uint8, device=dev)   # grid.py:205
        noise = torch.rand(size=(n_nerf, self.M, 3), device=dev, dtype=torch.float32) if n_nerf > 0 else None   # :326
        enc = m.pos_encoder
        table_h = None
        if enc.hash_table.dim() == 2:          # hash_encoder_half: fp16 copy of the table per forward (:367)
            if self._table_h is None or self._table_h.numel() != enc.hash_table.numel():
                self._table_h = torch.empty(enc.hash_table.numel(), dtype=torch.float16, device=dev)
            table_h = self._table_h
        og = self.args.occ_grid
        _lib.call("vn_occ_update", grid, self.grid_size, bf, winner,
                  f(ray_update["rays_o"]) if n_ray else None, f(ray_update["rays_d"]) if n_ray else None,
                  f(ray_update["depth_meas"]) if n_ray else None, n_ray,
                  f(nerf_update["rays_o"]) if n_nerf else None, f(nerf_update["rays_d"]) if n_nerf else None, noise, n_nerf,
                  self.M, self.I, float(self.args.model.scale), float(self.nerf_pos_noise_every_m),
                  float(self.false_detection_prob_every_m), float(self.std_every_m), float(self.prob_min),
                  float(og.nerf_threshold_max), float(og.nerf_threshold_slope), float(self.grid_decay),
                  1 if apply_decay else 0, float(self.threshold), enc.hash_table.detach(), table_h, enc._levels,
                  enc.kernel_flags, m.xyz_encoder.hidden_layers[0].weight.detach(), m.xyz_encoder.output_layer.weight.detach(),
                  -float(m.scale), float(m.scale), self._ws, self._ws.numel())
        self.bitfield = bf

    @torch.no_grad()
    def update(self, elapse_time: float):
        """:65-105"""
        ray_update, nerf_update = self._sample(elapse_time=elapse_time)

        m = self._native_model(self.occ_3d_grid.device)
        if m is not None:
            self.update_step += 1
            self._updateNative(m, ray_update, nerf_update,
                               apply_decay=self.update_step <= self.args.occ_grid.decay_warmup_steps)
            return

        if ray_update["batch_size"] > 0:
            self._rayUpdate(rays_o=ray_update["rays_o"], rays_d=ray_update["rays_d"], meas=ray_update["depth_meas"])
        if nerf_update["batch_size"] > 0:
            self._nerfUpdate(rays_o=nerf_update["rays_o"], rays_d=nerf_update["rays_d"], meas=nerf_update["depth_meas"])

        # warmup decay (:95-98) fused with the bitfield repack (:101-105)
        self.update_step += 1
        apply_decay = self.update_step <= self.args.occ_grid.decay_warmup_steps
        self._decayAndPack(apply_decay=apply_decay)

    @torch.no_grad()
    def _decayAndPack(self, apply_decay: bool):
        grid = self.occ_3d_grid
        assert grid.is_contiguous()
        bf = self.bitfield
        if bf.numel() != self.grid_size ** 3 // 8 or bf.device != grid.device:
            bf = torch.zeros(self.grid_size ** 3 // 8, dtype=torch.uint8, device=grid.device)   # grid.py:205
        _lib.call("vn_occ_decay_pack", grid, self.grid_size, float(self.grid_decay), 1 if apply_decay else 0,
                  float(self.threshold), bf)
        self.bitfield = bf

    @torch.no_grad()
    def _sample(self, elapse_time: float):
        """:108-180"""
        B = self.args.occ_grid.batch_size
        B_ray = int(B * self.args.occ_grid.batch_ratio_ray_update)
        B_nerf = B - B_ray
        sensors = self.args.training.sensors
        if "RGBD" in sensors:
            plan = (("random", "RGBD"), ("random", "RGBD"))
        elif ("ToF" in sensors) and ("USS" in sensors):
            plan = (("valid_tof", "ToF"), ("valid_uss", "USS"))
        elif ("ToF" in sensors) and not ("USS" in sensors):
            plan = (("valid_tof", "ToF"), ("valid_tof", "ToF"))
        elif not ("ToF" in sensors) and ("USS" in sensors):
            plan = (("valid_uss", "USS"), ("valid_uss", "USS"))
        else:
            self.args.logger.error("occupancy grid sampling strategy does not exist")
            raise ValueError("occupancy grid sampling strategy does not exist")
        ray_update = self._sampleBatch(B=B_ray, pixel_strategy=plan[0][0], sensor=plan[0][1], elapse_time=elapse_time)
        nerf_update = self._sampleBatch(B=B_nerf, pixel_strategy=plan[1][0], sensor=plan[1][1], elapse_time=elapse_time)
        return ray_update, nerf_update

    @torch.no_grad()
    def _sampleBatch(self, B: int, pixel_strategy: str, sensor: str, elapse_time: float):
        """:183-222"""
        data = self.dataset(batch_size=B, sampling_strategy={"imgs": "all", "pixs": pixel_strategy},
                            elapse_time=elapse_time)
        rays_o = data['rays_o']
        rays_d = data['rays_d']
        depth_meas = data['depth'][sensor]
        if sensor in data.get('depth_valid_by_construction', ()):
            # the sampler drew only pixels where this sensor has a measurement: the NaN filter of
            # the reference (:216-222) is the identity and its boolean-mask gather (a host sync) is skipped
            return {"batch_size": B, "rays_o": rays_o, "rays_d": rays_d, "depth_meas": depth_meas}
        valid_depth = ~torch.isnan(depth_meas)
        return {"batch_size": B, "rays_o": rays_o[valid_depth], "rays_d": rays_d[valid_depth],
                "depth_meas": depth_meas[valid_depth]}

    @torch.no_grad()
    def _rayUpdate(self, rays_o: torch.Tensor, rays_d: torch.Tensor, meas: torch.Tensor):
        """:225-258: _calcPos + _rayProb in one kernel, then the Bayes update"""
        N = rays_o.shape[0]
        if N == 0:
            return
        dev = rays_o.device
        cell_idxs = torch.empty(N * self.M, 3, dtype=torch.int32, device=dev)
        probs_occ = torch.empty(N, self.M, dtype=torch.float32, device=dev)
        probs_emp = torch.empty(N, self.M, dtype=torch.float32, device=dev)
        _lib.call("vn_occ_calc_pos_prob", rays_o.contiguous().float(), rays_d.contiguous().float(), None,
                  meas.contiguous().float(), N, self.M, self.I, self.grid_size, float(self.args.model.scale),
                  float(self.nerf_pos_noise_every_m), float(self.false_detection_prob_every_m),
                  float(self.std_every_m), float(self.prob_min), None, None, cell_idxs, probs_occ, probs_emp)
        self._updateGrid(cell_idxs=cell_idxs, probs_occ=probs_occ.reshape(-1), probs_emp=probs_emp.reshape(-1))

    @torch.no_grad()
    def _nerfUpdate(self, rays_o: torch.Tensor, rays_d: torch.Tensor, meas: torch.Tensor):
        """:261-290"""
        if rays_o.shape[0] == 0:
            return
        _, cell_pos, cell_idxs = self._calcPos(rays_o=rays_o, rays_d=rays_d, add_noise=True)
        probs_occ, probs_emp = self._nerfProb(cell_pos=cell_pos)
        self._updateGrid(cell_idxs=cell_idxs, probs_occ=probs_occ, probs_emp=probs_emp)

    @torch.no_grad()
    def _calcPos(self, rays_o: torch.Tensor, rays_d: torch.Tensor, add_noise: bool, noise: torch.Tensor = None):
        """:293-335 -> (cell_dists (N,M), cell_pos (N*M,3), cell_idxs (N*M,3)).  `noise` (N,M,3)
        uniform [0,1) may be supplied to reproduce a run (default torch.rand as in :326)."""
        N = rays_o.shape[0]
        dev = rays_o.device
        if add_noise and noise is None:
            noise = torch.rand(size=(N, self.M, 3), device=dev, dtype=torch.float32)
        cell_dists = torch.empty(N, self.M, dtype=torch.float32, device=dev)
        cell_pos = torch.empty(N * self.M, 3, dtype=torch.float32, device=dev)
        cell_idxs = torch.empty(N * self.M, 3, dtype=torch.int32, device=dev)
        _lib.call("vn_occ_calc_pos_prob", rays_o.contiguous().float(), rays_d.contiguous().float(),
                  noise.contiguous() if add_noise else None, None, N, self.M, self.I, self.grid_size,
                  float(self.args.model.scale), float(self.nerf_pos_noise_every_m),
                  float(self.false_detection_prob_every_m), float(self.std_every_m), float(self.prob_min),
                  cell_dists, cell_pos, cell_idxs, None, None)
        return cell_dists, cell_pos, cell_idxs

    @torch.no_grad()
    def _rayProb(self, meas: torch.Tensor, dists: torch.Tensor, return_probs: bool = False):
        """:338-389 on given distances (N,M): P[meas@dist | occ], P[meas@dist | emp]"""
        N, M = dists.shape
        dev = dists.device
        probs_occ = torch.empty(N, M, dtype=torch.float32, device=dev)
        probs_emp = torch.empty(N, M, dtype=torch.float32, device=dev)
        if return_probs:       # :387-388: + P[meas=dist|emp], P[meas=dist|occ], P[meas not< dist|emp], P[meas not< dist|occ]
            terms = torch.empty(4, N, M, dtype=torch.float32, device=dev)
            _lib.call("vn_occ_ray_prob_terms", meas.contiguous().float(), dists.contiguous().float(), N, M, self.I,
                      float(self.false_detection_prob_every_m), float(self.std_every_m), float(self.prob_min),
                      probs_occ, probs_emp, terms)
            return probs_occ, probs_emp, terms[0], terms[1], terms[2], terms[3]
        _lib.call("vn_occ_ray_prob", meas.contiguous().float(), dists.contiguous().float(), N, M, self.I,
                  float(self.false_detection_prob_every_m), float(self.std_every_m), float(self.prob_min),
                  probs_occ, probs_emp)
        return probs_occ, probs_emp

    @torch.no_grad()
    def _nerfProb(self, cell_pos: torch.Tensor):
        """:392-408"""
        cell_density = self.fct_density(x=cell_pos).contiguous().float()
        n = cell_density.shape[0]
        _, scratch = self._buffers(cell_density.device)
        probs_occ = torch.empty(n, dtype=torch.float32, device=cell_density.device)
        probs_emp = torch.empty(n, dtype=torch.float32, device=cell_density.device)
        _lib.call("vn_occ_nerf_prob", cell_density, n, float(self.args.occ_grid.nerf_threshold_max),
                  float(self.args.occ_grid.nerf_threshold_slope), scratch, probs_occ, probs_emp)
        return probs_occ, probs_emp

    @torch.no_grad()
    def _updateGrid(self, cell_idxs: torch.Tensor, probs_occ: torch.Tensor, probs_emp: torch.Tensor):
        """:411-430.  Duplicate cells: the entry with the largest flat index wins (the CPU
        index_put_ order); the reference's CUDA scatter is unordered."""
        n = cell_idxs.shape[0]
        if n == 0:
            return
        winner, _ = self._buffers(cell_idxs.device)
        tmp = torch.empty(n, dtype=torch.float32, device=cell_idxs.device)
        _lib.call("vn_occ_bayes_update", self.occ_3d_grid, self.grid_size, cell_idxs.contiguous(), n,
                  probs_occ.contiguous(), probs_emp.contiguous(), winner, tmp)

    @torch.no_grad()
    def _sensorEmptyPDF(self, shape: tuple):
        """:433-446: P[meas=dist | cell=emp] (constant false-detection density; analysis helper -- the update path has it
        inside vn_occ_calc_pos_prob)"""
        return self.false_detection_prob_every_m * torch.ones(shape, device=self.args.device, dtype=torch.float32)

    @torch.no_grad()
    def _sensorOccupiedPDF(self, meas: torch.Tensor, dists: torch.Tensor):
        """:449-465: P[meas=dist | cell=occ], any broadcastable shapes (analysis helper, see _sensorEmptyPDF)"""
        stds = self.std_every_m * dists + 0.00001
        return torch.exp(-0.5 * (meas - dists) ** 2 / stds ** 2)

    @torch.no_grad()
    def _c2idx(self, pos: torch.Tensor):
        """:468-480"""
        map_idxs = (self.grid_size - 1) * (pos + self.args.model.scale) / (2 * self.args.model.scale)
        return torch.clamp(map_idxs.round().to(dtype=torch.int32), 0, self.grid_size - 1)

    @torch.no_grad()
    def _idx2c(self, idx: torch.Tensor):
        """:483-496"""
        pos = (2 * self.args.model.scale) * (idx + 0.5) / self.grid_size - self.args.model.scale
        return torch.clamp(pos, -self.args.model.scale, self.args.model.scale)

