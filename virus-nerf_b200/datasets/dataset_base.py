"""Drop-in for the reference's datasets/dataset_base.py (DatasetBase.__call__: one training
batch from image-shaped storage).  File loading / scene conversion (dataset_ethz.py,
dataset_rh.py) stays the reference's: this class is fed the tensors those loaders produce and
keeps their attribute names (rgbs, poses, directions_dict, sensor_ids, depths_dict, times,
sampler).  The batch itself is ONE kernel (vn_batch_assemble) instead of the per-camera
boolean-mask loop of _calcRayPoses (:194-243) plus two advanced-indexing gathers per tensor.
"""
import torch

from .. import _lib

_MAX_DEPTH_SENSORS = 4


class DatasetBase(torch.utils.data.Dataset):
    def __init__(self, args, split='train', rgbs=None, poses=None, directions_dict=None, sensor_ids=None,
                 depths_dict=None, times=None, img_wh=None, sampler=None, sensor_name2id=None):
        """rgbs [N, H*W, >=3] f32; poses [N, 3, 4] f32; directions_dict {cam_id: [H*W, 3]}; sensor_ids [N] int;
        depths_dict {sensor: [N, H*W]} f32 with NaN = no measurement; times [N]; sensor_name2id: cam_id -> id
        (helpers/data_fcts.py:sensorName2ID of the reference)."""
        self.args = args
        self.split = split
        self.rgbs, self.poses, self.directions_dict = rgbs, poses, directions_dict
        self.sensor_ids, self.depths_dict, self.times = sensor_ids, depths_dict or {}, times
        self.img_wh = img_wh
        self.sampler = sampler
        self.sensor_name2id = sensor_name2id or (lambda cam_id: int(cam_id))
        self._packed = None
        if len(self.depths_dict) > _MAX_DEPTH_SENSORS:
            raise ValueError(f"at most {_MAX_DEPTH_SENSORS} depth sensors")

    def __len__(self):
        return len(self.poses)

    def to(self, device):
        """dataset_base.py:78-98"""
        self.rgbs = self.rgbs.to(device)
        self.poses = self.poses.to(device)
        self.times = self.times.to(device)
        self.sensor_ids = self.sensor_ids.to(device)
        for key in self.depths_dict.keys():
            self.depths_dict[key] = self.depths_dict[key].to(device)
        for cam_id, directions in self.directions_dict.items():
            self.directions_dict[cam_id] = directions.to(device)
        self._packed = None
        return self

    def _pack(self):
        """kernel-side views of the storage: stacked camera directions and the image -> camera row table
        (what the `sensor_ids[img_idxs] == id` masks of _calcRayPoses select, :218-221)"""
        if self._packed is None:
            dev = self.poses.device
            cams = list(self.directions_dict.keys())
            directions = torch.stack([self.directions_dict[c].to(torch.float32) for c in cams]).contiguous()
            slot = torch.full((len(self.poses),), -1, dtype=torch.int32, device=dev)
            for k, cam in enumerate(cams):
                slot[self.sensor_ids == self.sensor_name2id(cam)] = k
            self._packed = {
                "directions": directions, "slot": slot, "poses": self.poses.to(torch.float32).contiguous(),
                "rgbs": self.rgbs.to(torch.float32).contiguous(), "ids": self.sensor_ids.to(torch.int32).contiguous(),
                "times": self.times.to(torch.float32).contiguous(),
                "depths": {k: v.to(torch.float32).contiguous() for k, v in self.depths_dict.items()},
                "err": torch.zeros(1, dtype=torch.int32, device=dev),
            }
        return self._packed

    def __call__(self, batch_size: int = None, sampling_strategy: dict = None, elapse_time: float = None,
                 img_idxs: torch.Tensor = None, pix_idxs: torch.Tensor = None):
        """dataset_base.py:23-76: dict(img_idxs, pix_idxs, sensor_ids, time, rays_o, rays_d, rgb, depth{sensor})"""
        if img_idxs is None or pix_idxs is None:
            img_idxs, pix_idxs = self.sampler(batch_size=batch_size, sampling_strategy=sampling_strategy,
                                              elapse_time=elapse_time)
        pk = self._pack()
        dev = pk["poses"].device
        B = img_idxs.shape[0]
        ii = img_idxs.to(device=dev, dtype=torch.int32).contiguous()
        pi = pix_idxs.to(device=dev, dtype=torch.int32).contiguous()
        rays_o = torch.empty(B, 3, device=dev); rays_d = torch.empty(B, 3, device=dev); rgb = torch.empty(B, 3, device=dev)
        ids = torch.empty(B, dtype=torch.int32, device=dev); time = torch.empty(B, device=dev)
        names = list(pk["depths"].keys())
        maps = [pk["depths"][k] for k in names] + [None] * (_MAX_DEPTH_SENSORS - len(names))
        outs = [torch.empty(B, device=dev) for _ in names] + [None] * (_MAX_DEPTH_SENSORS - len(names))
        HW = pk["directions"].shape[1]
        _lib.call("vn_batch_assemble", ii, pi, B, pk["poses"], pk["slot"], pk["poses"].shape[0], pk["directions"],
                  pk["directions"].shape[0], HW, pk["rgbs"], pk["rgbs"].shape[2], *maps, pk["ids"], pk["times"],
                  rays_o, rays_d, rgb, *outs, ids, time, pk["err"])
        return {'img_idxs': img_idxs, 'pix_idxs': pix_idxs, 'sensor_ids': ids.to(self.sensor_ids.dtype), 'time': time,
                'rays_o': rays_o, 'rays_d': rays_d, 'rgb': rgb, 'depth': {k: o for k, o in zip(names, outs)}}

    def _calcRayPoses(self, directions_dict=None, poses=None, sensor_ids=None, img_idxs=None, pix_idxs=None):
        """dataset_base.py:194-243 (the arguments other than the indices are the dataset's own tensors)"""
        out = self(img_idxs=img_idxs, pix_idxs=pix_idxs)
        return out['rays_o'], out['rays_d']

    def getMeanHeight(self):
        """dataset_base.py:100-108"""
        return torch.mean(self.poses[:, 2, 3]).item()

    def indices_out_of_range(self):
        """True when any batch so far addressed an image / pixel / camera outside the storage (the reference
        leaves NaN rays and logs in debug mode, :233-236); reading it synchronises"""
        return self._packed is not None and int(self._packed["err"]) != 0
