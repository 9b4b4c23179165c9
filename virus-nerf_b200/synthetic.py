"""Synthetic ETHZ- / Robot@Home2-shaped scenes.

There are no datasets in the build image, so the benchmark and the tests feed the hot path
with rays of the same shape and dict layout the reference's ``DatasetBase.__call__`` produces
(datasets/dataset_base.py:23-76): unit ray directions from pinhole cameras on a planar
trajectory inside the cube, RGB targets, and USS / ToF depth targets with NaN where the
sensor has no measurement (USS: per-image minimum depth inside an elliptical field of view,
datasets/sensor_uss.py:219-239; ToF: an 8x8 pixel lattice, datasets/sensor_tof.py:80-109).
Depth is analytic: a room shell (walls at +-0.9*scale) plus a few boxes.

Everything here is host-side numpy; `SyntheticDataset` keeps a pre-generated pool of rays
either in HBM (`device="cuda"`) or in pinned host memory (`pinned=True`, used for the
end-to-end measurement where the per-step host->device copy is inside the timed region).
"""
import math
from types import SimpleNamespace

import numpy as np
import torch


def make_args(device="cuda", scale=0.5, sensors=("USS", "ToF"), grid_type="occ", batch_size=4096,
              occ_batch_size=1024, update_interval=8, decay_warmup_steps=80, lr=5e-3):
    """the slice of the reference's Args (args/args.py, args/h_params.py) the hot path reads;
    values = args/ethz_usstof_not_optimized_gpu.json"""
    return SimpleNamespace(
        device=torch.device(device),
        seed=21,
        exp_step_factor=(1 / 256 if scale > 0.5 else 0.0),          # args/args.py:84
        model=SimpleNamespace(scale=scale, grid_type=grid_type, encoder_type="hash", hash_levels=16,
                              hash_max_res=1024),
        training=SimpleNamespace(sensors=list(sensors), batch_size=batch_size, lr=lr, debug_mode=False,
                                 color_loss_w=1.0, tof_loss_w=50.0, uss_loss_w=50.0, rgbd_loss_w=100.0,
                                 sampling_strategy={"imgs": "all", "pixs": {"valid_uss": 0.4, "valid_tof": 0.4}}),
        ngp_grid=SimpleNamespace(update_interval=16, warmup_steps=256),     # args/ethz_usstof_win.json:61-65
        occ_grid=SimpleNamespace(batch_size=occ_batch_size, update_interval=update_interval,
                                 decay_warmup_steps=decay_warmup_steps, batch_ratio_ray_update=0.5,
                                 false_detection_prob_every_m=0.3, std_every_m=0.2, nerf_pos_noise_every_m=0.2,
                                 nerf_threshold_max=5.91, nerf_threshold_slope=0.01),
        logger=SimpleNamespace(error=print, warning=print, info=print),
    )


class RoomScene:
    """analytic room shell + boxes in cube coordinates [-scale, scale]^3"""

    def __init__(self, scale=0.5, seed=21, n_boxes=5):
        self.scale = scale
        rng = np.random.default_rng(seed)
        w = 0.9 * scale
        self.room_min = np.array([-w, -w, -0.6 * scale], np.float64)
        self.room_max = np.array([w, w, 0.6 * scale], np.float64)
        boxes = []
        for _ in range(n_boxes):
            c = rng.uniform(-0.6 * scale, 0.6 * scale, 3)
            c[2] = self.room_min[2] + rng.uniform(0.05, 0.3) * scale
            h = rng.uniform(0.06, 0.16, 3) * scale
            boxes.append((c - h, c + h))
        self.boxes = boxes

    def depth(self, o, d):
        """distance along unit rays (N,3) to the first surface"""
        o = np.asarray(o, np.float64); d = np.asarray(d, np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d
            # room: exit distance from inside
            t0 = (self.room_min - o) * inv
            t1 = (self.room_max - o) * inv
            t_exit = np.nanmin(np.maximum(t0, t1), axis=1)
            best = t_exit
            for lo, hi in self.boxes:
                a = (lo - o) * inv; b = (hi - o) * inv
                tn = np.nanmax(np.minimum(a, b), axis=1)
                tf = np.nanmin(np.maximum(a, b), axis=1)
                hit = (tn <= tf) & (tf > 0) & (tn > 0)
                best = np.where(hit & (tn < best), tn, best)
        return best

    def occupancy_bitfield(self, grid_size=128, thickness_cells=1.5):
        """'carved' state: cells within `thickness` of a surface are occupied (Morton order)"""
        G = grid_size
        s = self.scale
        c = (np.arange(G) + 0.5) / G * 2 * s - s
        X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
        P = np.stack([X, Y, Z], -1)
        th = thickness_cells * 2 * s / G

        def box_sdf(p, lo, hi):
            ctr = (lo + hi) / 2; half = (hi - lo) / 2
            q = np.abs(p - ctr) - half
            return np.linalg.norm(np.maximum(q, 0), axis=-1) + np.minimum(q.max(-1), 0)

        occ = np.abs(box_sdf(P, self.room_min, self.room_max)) < th
        for lo, hi in self.boxes:
            occ |= np.abs(box_sdf(P, lo, hi)) < th
        return occ  # cartesian bool [G,G,G]


def morton_pack(occ_cart):
    """cartesian bool [G,G,G] -> Morton-order uint8 bitfield (host restatement used only to
    build test / benchmark inputs; grid.py:165-170 + utils.py:157-169)"""
    G = occ_cart.shape[0]
    r = np.arange(G, dtype=np.uint32)

    def expand(v):
        v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
        v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
        v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
        v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
        return v
    with np.errstate(over="ignore"):
        e = expand(r)
    idx = e[:, None, None] | (e[None, :, None] << np.uint32(1)) | (e[None, None, :] << np.uint32(2))
    flat = np.zeros(G ** 3, np.uint8)
    flat[idx.reshape(-1)] = occ_cart.reshape(-1).astype(np.uint8)
    return np.packbits(flat.reshape(-1, 8), axis=1, bitorder="little").reshape(-1)


class SyntheticDataset:
    """ray pool with the reference's batch dict layout.  kind: 'ethz' (640x480, 2 cameras,
    USS + ToF) or 'rh2' (240x320, 4 cameras, RGBD + USS + ToF)."""

    def __init__(self, scene=None, kind="ethz", n_images=200, pool_size=1 << 18, device="cuda", pinned=False,
                 seed=21):
        self.scene = scene if scene is not None else RoomScene()
        self.kind = kind
        rng = np.random.default_rng(seed)
        s = self.scene.scale
        if kind == "ethz":
            W, H, n_cam, aov = 640, 480, 2, (90.0, 65.0)
        else:
            W, H, n_cam, aov = 240, 320, 4, (58.0, 73.0)
        self.img_wh = (W, H)
        fx = 0.5 * W / math.tan(math.radians(aov[0]) / 2)
        fy = 0.5 * H / math.tan(math.radians(aov[1]) / 2)
        # planar trajectory at fixed height, looking horizontally
        n_pose = n_images // n_cam
        ang = np.linspace(0, 2 * np.pi, n_pose, endpoint=False)
        centers = np.stack([0.35 * s * np.cos(ang), 0.35 * s * np.sin(ang), np.full(n_pose, -0.1 * s)], -1)
        poses = []
        for k in range(n_pose):
            for c in range(n_cam):
                yaw = ang[k] + np.pi / 2 + c * (2 * np.pi / n_cam)
                fwd = np.array([np.cos(yaw), np.sin(yaw), 0.0])
                up = np.array([0.0, 0.0, 1.0])
                right = np.cross(fwd, up)
                R = np.stack([right, -up, fwd], -1)          # camera x right, y down, z forward
                poses.append((R, centers[k]))
        self.n_images = len(poses)

        def rays_for(img, u, v):
            dirs = np.stack([(u + 0.5 - W / 2) / fx, (v + 0.5 - H / 2) / fy, np.ones_like(u, dtype=np.float64)], -1)
            dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
            Rm = np.stack([poses[i][0] for i in img])
            o = np.stack([poses[i][1] for i in img])
            d = np.einsum("nij,nj->ni", Rm, dirs)
            return o, d

        # USS: per-image minimum depth inside the ellipse (coarse lattice is enough for a target)
        ru, rv = (W * 55 / aov[0]) / 2, (H * 35 / aov[1]) / 2
        gu, gv = np.meshgrid(np.arange(0, W, 8), np.arange(0, H, 8), indexing="xy")
        ell = (((gu - W / 2) / ru) ** 2 + ((gv - H / 2) / rv) ** 2) <= 1.0
        eu, ev = gu[ell].astype(np.float64), gv[ell].astype(np.float64)
        uss_min = np.zeros(self.n_images)
        for i in range(self.n_images):
            o, d = rays_for(np.full(eu.shape, i), eu, ev)
            uss_min[i] = self.scene.depth(o, d).min()
        # ToF: 8x8 lattice spanning a centred window
        tu = np.round(np.linspace(W / 2 - W / 4, W / 2 + W / 4, 8)).astype(int)
        tv = np.round(np.linspace(H / 2 - H * 0.35, H / 2 + H * 0.35, 8)).astype(int)

        # pool: 40 % USS-valid pixels, 40 % ToF-valid pixels, 20 % random pixels
        n_uss = int(0.4 * pool_size); n_tof = int(0.4 * pool_size); n_rnd = pool_size - n_uss - n_tof
        img = rng.integers(0, self.n_images, pool_size)
        u = np.empty(pool_size); v = np.empty(pool_size)
        # uss pixels: rejection-free sampling inside the ellipse
        r = np.sqrt(rng.random(n_uss)); th = rng.random(n_uss) * 2 * np.pi
        u[:n_uss] = np.floor(W / 2 + r * ru * np.cos(th) * 0.999)
        v[:n_uss] = np.floor(H / 2 + r * rv * np.sin(th) * 0.999)
        u[n_uss:n_uss + n_tof] = tu[rng.integers(0, 8, n_tof)]
        v[n_uss:n_uss + n_tof] = tv[rng.integers(0, 8, n_tof)]
        u[n_uss + n_tof:] = rng.integers(0, W, n_rnd)
        v[n_uss + n_tof:] = rng.integers(0, H, n_rnd)
        o, d = rays_for(img, u, v)
        depth = self.scene.depth(o, d)
        in_ell = (((u - W / 2) / ru) ** 2 + ((v - H / 2) / rv) ** 2) <= 1.0
        on_tof = np.isin(u, tu) & np.isin(v, tv)
        d_uss = np.where(in_ell, uss_min[img], np.nan)
        d_tof = np.where(on_tof, depth, np.nan)
        hit = o + d * depth[:, None]
        rgb = 0.5 + 0.5 * np.sin(hit * (12.0 / s) + np.array([0.0, 2.0, 4.0]))

        self.pool_size = pool_size
        self.idx_uss = torch.from_numpy(np.nonzero(in_ell)[0])
        self.idx_tof = torch.from_numpy(np.nonzero(on_tof)[0])
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        pool = {"rays_o": f(o), "rays_d": f(d), "rgb": f(rgb), "USS": f(d_uss), "ToF": f(d_tof), "RGBD": f(depth)}
        self.device = torch.device(device)
        self.pinned = pinned
        if pinned:
            self.pool = {k: t.pin_memory() for k, t in pool.items()}
        else:
            self.pool = {k: t.to(self.device) for k, t in pool.items()}
            self.idx_uss = self.idx_uss.to(self.device)
            self.idx_tof = self.idx_tof.to(self.device)
        self.gen = torch.Generator(device="cpu" if pinned else self.device)
        self.gen.manual_seed(seed)
        self.sensors = ("RGBD", "USS", "ToF") if kind == "rh2" else ("USS", "ToF")
        self.fast_gather = True      # device pool: vn_pool_gather instead of torch advanced indexing (same batches)

    def __len__(self):
        return self.n_images

    def clone_with_seed(self, seed):
        """same ray pool, independent sampler state (shallow copy)"""
        import copy
        c = copy.copy(self)
        c.gen = torch.Generator(device=self.gen.device)
        c.gen.manual_seed(seed)
        return c

    def sample_indices(self, batch_size, sampling_strategy):
        pixs = sampling_strategy.get("pixs", "random") if sampling_strategy else "random"
        dev = self.gen.device
        ri = lambda hi, n: torch.randint(0, hi, (n,), generator=self.gen, device=dev)
        if isinstance(pixs, dict):
            n_u = int(batch_size * pixs.get("valid_uss", 0.0))
            n_t = int(batch_size * pixs.get("valid_tof", 0.0))
            n_r = batch_size - n_u - n_t
            return torch.cat([self.idx_uss[ri(len(self.idx_uss), n_u)], self.idx_tof[ri(len(self.idx_tof), n_t)],
                              ri(self.pool_size, n_r)])
        if pixs == "valid_uss":
            return self.idx_uss[ri(len(self.idx_uss), batch_size)]
        if pixs == "valid_tof":
            return self.idx_tof[ri(len(self.idx_tof), batch_size)]
        return ri(self.pool_size, batch_size)

    def gather(self, idx, non_blocking=True):
        """batch dict on self.device (host pool: gather on the host, then ONE pinned H2D copy per key)"""
        out = {}
        for k, t in self.pool.items():
            b = t[idx]
            if self.pinned:
                b = b.pin_memory().to(self.device, non_blocking=non_blocking)
            out[k] = b
        return {"rays_o": out["rays_o"], "rays_d": out["rays_d"], "rgb": out["rgb"],
                "depth": {s: out[s] for s in self.sensors}}

    def _fast_call(self, batch_size, sampling_strategy):
        """device-resident pool: the same random draws as sample_indices(), but ONE vn_pool_gather launch per segment
        writes all outputs (instead of index + six advanced-indexing gathers); bit-identical batches"""
        from . import _lib
        pixs = sampling_strategy.get("pixs", "random") if sampling_strategy else "random"
        dev = self.device
        if isinstance(pixs, dict):
            n_u = int(batch_size * pixs.get("valid_uss", 0.0))
            n_t = int(batch_size * pixs.get("valid_tof", 0.0))
            segs = [(self.idx_uss, n_u), (self.idx_tof, n_t), (None, batch_size - n_u - n_t)]
        else:
            segs = [({"valid_uss": self.idx_uss, "valid_tof": self.idx_tof}.get(pixs), batch_size)]
        B = batch_size
        ns = len(self.sensors)
        out = torch.empty((9 + ns) * B, dtype=torch.float32, device=dev)
        ro, rd, rgb = out[0:3 * B].view(B, 3), out[3 * B:6 * B].view(B, 3), out[6 * B:9 * B].view(B, 3)
        dep = [out[(9 + k) * B:(10 + k) * B] for k in range(ns)]
        src = [self.pool[s_] for s_ in self.sensors] + [None] * (3 - ns)
        off = 0
        for sel, n in segs:
            if n == 0:
                continue
            hi = len(sel) if sel is not None else self.pool_size
            draw = torch.randint(0, hi, (n,), generator=self.gen, device=self.gen.device)
            _lib.call("vn_pool_gather", draw, sel, hi if sel is not None else 0, self.pool_size, n, self.pool["rays_o"],
                      self.pool["rays_d"], self.pool["rgb"], src[0], src[1], src[2], ro[off:off + n], rd[off:off + n],
                      rgb[off:off + n], dep[0][off:off + n] if ns > 0 else None, dep[1][off:off + n] if ns > 1 else None,
                      dep[2][off:off + n] if ns > 2 else None, None)
            off += n
        return {"rays_o": ro, "rays_d": rd, "rgb": rgb, "depth": {s_: dep[k] for k, s_ in enumerate(self.sensors)}}

    def __call__(self, batch_size, sampling_strategy=None, elapse_time=None):
        """datasets/dataset_base.py:23-76 shape: dict(rays_o, rays_d, rgb, depth{sensor: (N,)})"""
        if self.fast_gather and not self.pinned and self.device.type == "cuda":
            batch = self._fast_call(batch_size, sampling_strategy)
        else:
            batch = self.gather(self.sample_indices(batch_size, sampling_strategy))
        pixs = sampling_strategy.get("pixs", "random") if sampling_strategy else "random"
        # every ray drawn from the valid-<sensor> subset carries a measurement of that sensor
        batch["depth_valid_by_construction"] = {"valid_uss": ("USS",), "valid_tof": ("ToF",)}.get(pixs, ()) \
            if isinstance(pixs, str) else ()
        return batch


def scan_rays(n=512, scale=0.5, height=0.0, origin=(0.0, 0.0)):
    """evaluation scan rays with d_z = 0 (helpers/geometric_fcts.py:77-111 shape)"""
    ang = np.linspace(-np.pi, np.pi, n, endpoint=False)
    d = np.stack([np.cos(ang), np.sin(ang), np.zeros(n)], -1).astype(np.float32)
    o = np.tile(np.array([origin[0], origin[1], height], np.float32), (n, 1))
    return o, d
