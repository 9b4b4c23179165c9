"""Drop-in for the reference's modules/grid.py (Grid base class: Morton bitfield handling)."""
from abc import abstractmethod

import numpy as np
import torch

from .. import _lib
from .utils import morton3D, packbits


class Grid():
    def __init__(self, args, grid_size: int, cascades: int, morton_structure: bool) -> None:
        self.args = args
        self.grid_size = grid_size
        self.cascades = cascades
        self.morton_structure = morton_structure
        self._grid_coords = None   # built on first use (the reference uses kornia.create_meshgrid3d, grid.py:29-35)

        # grid.py:37: one cascade worth of bits per cascade
        self.bitfield = torch.zeros(self.cascades * self.grid_size ** 3 // 8, dtype=torch.uint8,
                                    device=self.args.device)
        if morton_structure:
            self.occ_morton_grid = torch.zeros(self.cascades, self.grid_size ** 3, device=self.args.device)
        else:
            self.occ_3d_grid = torch.zeros(self.grid_size, self.grid_size, self.grid_size)

    @property
    def grid_coords(self):
        """all G^3 integer coordinates, (N,3) int32.  Any full enumeration gives the same
        cartesian<->Morton permutation (grid.py:165-170, 187-189)."""
        if self._grid_coords is None:
            r = torch.arange(self.grid_size, dtype=torch.int32, device=self.args.device)
            x, y, z = torch.meshgrid(r, r, r, indexing="ij")
            self._grid_coords = torch.stack([x, y, z], dim=-1).reshape(-1, 3).contiguous()
        return self._grid_coords

    @abstractmethod
    def update(self):
        pass

    @torch.no_grad()
    def getBitfield(self, clone: bool = False):
        """grid.py:49-62"""
        if clone:
            return self.bitfield.clone().detach()
        return self.bitfield

    @torch.no_grad()
    def getOccupancyCartesianGrid(self, clone: bool = False):
        """grid.py:65-85"""
        if self.morton_structure:
            grid = self.morton2cartesian(grid_morton=self.occ_morton_grid[0])
        else:
            grid = self.occ_3d_grid
        if clone:
            return grid.clone().detach()
        return grid

    @torch.no_grad()
    def getBinaryCartesianGrid(self, threshold: float):
        """grid.py:88-109"""
        if self.morton_structure:
            grid = self.morton2cartesian(grid_morton=self.occ_morton_grid[0])
        else:
            grid = self.occ_3d_grid
        return self.thresholdGrid(grid=grid, threshold=threshold)

    @torch.no_grad()
    def getAllCells(self):
        """grid.py:112-125"""
        indices = morton3D(self.grid_coords).long()
        cells = [(indices, self.grid_coords)] * self.cascades
        return cells

    @torch.no_grad()
    def updateBitfield(self, grid: torch.tensor, threshold: float, convert_cart2morton: bool):
        """grid.py:128-151.  The cartesian path is one fused kernel (Morton permutation +
        packbits, vn_occ_decay_pack with decay disabled) instead of morton3D over all G^3
        coordinates + a permuting scatter + packbits."""
        if convert_cart2morton:
            bf = torch.zeros(self.grid_size ** 3 // 8, dtype=torch.uint8, device=grid.device)
            _lib.call("vn_occ_decay_pack", grid.contiguous(), self.grid_size, 1.0, 0, float(threshold), bf)
            self.bitfield = bf
        else:
            self.bitfield = self.morton2bitfield(occ_morton=grid, threshold=threshold)

    @torch.no_grad()
    def cartesian2morton(self, grid_3d: torch.tensor):
        """grid.py:154-170"""
        indices, coords = self.getAllCells()[0]
        grid_morton = torch.zeros(self.grid_size ** 3, dtype=grid_3d.dtype, device=self.args.device)
        grid_morton[indices] = grid_3d[coords[:, 0].long(), coords[:, 1].long(), coords[:, 2].long()]
        return grid_morton

    @torch.no_grad()
    def morton2cartesian(self, grid_morton: torch.tensor):
        """grid.py:173-189"""
        indices, coords = self.getAllCells()[0]
        grid_3d = torch.zeros(self.grid_size, self.grid_size, self.grid_size, dtype=grid_morton.dtype,
                              device=self.args.device)
        grid_3d[coords[:, 0].long(), coords[:, 1].long(), coords[:, 2].long()] = grid_morton[indices]
        return grid_3d

    @torch.no_grad()
    def morton2bitfield(self, occ_morton: torch.tensor, threshold: float) -> torch.tensor:
        """grid.py:192-211"""
        bin_bitfield = torch.zeros(self.grid_size ** 3 // 8, dtype=torch.uint8, device=self.args.device)
        packbits(density_grid=occ_morton.reshape(-1).contiguous(), density_threshold=threshold,
                 density_bitfield=bin_bitfield)
        return bin_bitfield

    @torch.no_grad()
    def bitfield2morton(self, bin_bitfield: torch.tensor):
        """grid.py:214-233 (debug inverse of packbits)"""
        bin_bitfield = bin_bitfield.clone().detach()
        mask = torch.tensor([[1, 2, 4, 8, 16, 32, 64, 128]], dtype=torch.uint8, device=bin_bitfield.device)
        bin_morton = (bin_bitfield.reshape(-1, 1) & mask).to(dtype=torch.bool)
        return bin_morton.reshape(-1)

    @torch.no_grad()
    def thresholdGrid(self, grid: torch.tensor, threshold: float):
        """grid.py:236-252: strict > threshold"""
        return grid > threshold

    @torch.no_grad()
    def c2oCoordinates(self, pos_c):
        """grid.py:255-269"""
        height_o = self.grid_size * (pos_c + self.args.model.scale) / (2 * self.args.model.scale)
        if torch.is_tensor(height_o):
            return torch.round(height_o).to(dtype=torch.int32)
        return np.round(height_o).astype(np.int32)
