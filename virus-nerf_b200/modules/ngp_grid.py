"""Drop-in for the reference's modules/ngp_grid.py: NGPGrid, the Instant-NGP baseline density
grid of the "occ vs ngp" ablation (args/ethz_usstof_win.json: grid_type "ngp"; trainer.py:108-112).

The update is six kernels without a host synchronisation (the reference does torch.nonzero,
len(...) and .mean().item() on the host side every update): occupied-cell sampling by rank
query, cell -> jittered world position, density query (the fused tcgen05 MLP through
`fct_density`), scatter + decayed maximum, deterministic mean / threshold, packbits.  Random
numbers come from torch on the host side of the boundary, in the reference's call order.
"""
import numpy as np
import torch

from .. import _lib
from .grid import Grid
from .utils import NEAR_DISTANCE, morton3D, morton3D_invert


class NGPGrid(Grid):
    def __init__(self, args, grid_size: int, fct_density: callable):
        self.args = args
        self.grid_size = grid_size
        self.scale = args.model.scale
        self.fct_density = fct_density
        self.cascades = max(1 + int(np.ceil(np.log2(2 * self.scale))), 1)     # ngp_grid.py:27
        super().__init__(args=args, grid_size=grid_size, cascades=self.cascades, morton_structure=True)
        self.threshold = 0.5
        self._bufs = None

    def _buffers(self):
        dev = self.occ_morton_grid.device
        if self._bufs is None or self._bufs["winner"].device != dev:
            G3 = self.grid_size ** 3
            self._bufs = {
                "winner": torch.full((G3,), -1, dtype=torch.int32, device=dev),
                "tmp": torch.zeros(G3, dtype=torch.float32, device=dev),
                "select": torch.zeros(_lib.ngp_select_tmp_ints(G3), dtype=torch.int32, device=dev),
                "scratch": torch.zeros(_lib.ngp_threshold_tmp_bytes() // 8, dtype=torch.float64, device=dev),
                "thr": torch.zeros(2, dtype=torch.float32, device=dev),
            }
        return self._bufs

    @torch.no_grad()
    def sample_uniform_and_occupied_cells(self, M, density_threshold):
        """ngp_grid.py:38-66.  Cells that the reference would not append (no occupied cell in the
        cascade) come back with index -1 and are ignored by update()."""
        dev = self.occ_morton_grid.device
        b = self._buffers()
        G3 = self.grid_size ** 3
        cells = []
        for c in range(self.cascades):
            coords1 = torch.randint(self.grid_size, (M, 3), dtype=torch.int32, device=dev)     # :50-52
            indices1 = morton3D(coords1).long()
            rand_idx = torch.randint(G3, (M,), device=dev)       # :58 draws below len(indices2); the kernel reduces modulo it
            indices2 = torch.empty(M, dtype=torch.int64, device=dev)
            _lib.call("vn_ngp_sample_occupied", self.occ_morton_grid[c], G3, float(density_threshold), rand_idx, M,
                      b["select"], indices2)
            coords2 = morton3D_invert(indices2.clamp(min=0).int())                            # :61
            cells += [(torch.cat([indices1, indices2]), torch.cat([coords1, coords2]))]
        return cells

    @torch.no_grad()
    def mark_invisible_cells(self, K, poses, img_wh, chunk=32 ** 3):
        """ngp_grid.py:68-112.  Never called by the reference's trainer (one-off initialisation in
        upstream ngp_pl); kept as the same torch expression sequence."""
        N_cams = poses.shape[0]
        self.count_grid = torch.zeros_like(self.occ_morton_grid)
        w2c_R = poses[:, :3, :3].transpose(1, 2)
        w2c_T = -w2c_R @ poses[:, :3, 3:]
        cells = self.getAllCells()
        for c in range(self.cascades):
            indices, coords = cells[c]
            for i in range(0, len(indices), chunk):
                xyzs = coords[i:i + chunk] / (self.grid_size - 1) * 2 - 1
                s = min(2 ** (c - 1), self.scale)
                half_grid_size = s / self.grid_size
                xyzs_w = (xyzs * (s - half_grid_size)).T
                xyzs_c = w2c_R @ xyzs_w + w2c_T
                uvd = K @ xyzs_c
                uv = uvd[:, :2] / uvd[:, 2:]
                in_image = (uvd[:, 2] >= 0) & (uv[:, 0] >= 0) & (uv[:, 0] < img_wh[0]) & (uv[:, 1] >= 0) & (uv[:, 1] < img_wh[1])
                covered_by_cam = (uvd[:, 2] >= NEAR_DISTANCE) & in_image
                self.count_grid[c, indices[i:i + chunk]] = count = covered_by_cam.sum(0) / N_cams
                too_near_to_any_cam = ((uvd[:, 2] < NEAR_DISTANCE) & in_image).any(0)
                valid_mask = (count > 0) & (~too_near_to_any_cam)
                self.occ_morton_grid[c, indices[i:i + chunk]] = torch.where(valid_mask, 0., -1.)

    @torch.no_grad()
    def update(self, density_threshold, warmup=False, decay=0.95, erode=False, noise=None):
        """ngp_grid.py:114-163.  `noise` (list of [M,3] uniform tensors per cascade) replaces
        torch.rand_like for reproducible tests."""
        b = self._buffers()
        G3 = self.grid_size ** 3
        if warmup:                                                                     # :122-126
            cells = self.getAllCells()
        else:
            cells = self.sample_uniform_and_occupied_cells(G3 // 4, density_threshold)
        decay_cells = None
        if erode:                                                                      # :146-147
            decay_cells = torch.clamp(decay ** (1 / self.count_grid), 0.1, 0.95).contiguous()
        self.occ_morton_grid = self.occ_morton_grid.contiguous()
        for c in range(self.cascades):
            indices, coords = cells[c]
            M = indices.shape[0]
            s = min(2 ** (c - 1), self.scale)                                          # :130
            half_grid_size = s / self.grid_size
            u = noise[c] if noise is not None else torch.rand(M, 3, device=coords.device)   # :135 rand_like
            xyzs_w = torch.empty(M, 3, device=coords.device)
            _lib.call("vn_ngp_cell_positions", coords.contiguous(), u.contiguous(), M, self.grid_size,
                      float(np.float32(s - half_grid_size)), float(np.float32(half_grid_size)), xyzs_w)
            sigmas = self.fct_density(xyzs_w).to(torch.float32).contiguous()            # :136
            _lib.call("vn_ngp_grid_update", self.occ_morton_grid[c], b["tmp"], b["winner"], G3, indices.contiguous(),
                      sigmas, M, float(decay), None if decay_cells is None else decay_cells[c])
        # :154-163: threshold from the mean of the positive cells (stays on the device), bitfield
        bf = torch.empty(self.cascades * G3 // 8, dtype=torch.uint8, device=self.occ_morton_grid.device)
        _lib.call("vn_ngp_threshold_pack", self.occ_morton_grid, self.cascades * G3, float(density_threshold), b["scratch"],
                  b["thr"], bf)
        self._threshold_dev = b["thr"]
        self.bitfield = bf

    @property
    def threshold(self):
        """min(mean density of the positive cells, density_threshold) of the last update (:155-156);
        reading it synchronises -- the update itself does not"""
        t = getattr(self, "_threshold_dev", None)
        return self._threshold_host if t is None else float(t[1])

    @threshold.setter
    def threshold(self, v):
        self._threshold_host = v
        self._threshold_dev = None
