"""Drop-in for the reference's training/sampler.py (image / pixel index sampling).

This is the RNG side of the boundary: the same torch.randint calls in the same order, with the
same shapes and dtypes, so a given torch seed reproduces the reference's index stream on the same
device type.  What changes: the valid-pixel index lists of the sensor masks (`torch.where(mask)`,
sampler.py:259, a host-synchronising nonzero per call) are built once, and the real-time filter
(:79-87) is a prefix count because the time stamps are sorted.
"""
import copy
import sys

import numpy as np
import torch


class Sampler():
    def __init__(self, args, dataset_len: int, img_wh: tuple, sensors_dict: dict = None, times: torch.Tensor = None) -> None:
        self.args = args
        self.dataset_len = dataset_len
        self.img_wh = img_wh
        self.sensors_dict = sensors_dict
        self.times = times
        self.rgn = np.random.default_rng(seed=self.args.seed)       # sampler.py:32
        self._mask_idxs = {}

    def __call__(self, batch_size: int, sampling_strategy: dict, elapse_time: float):
        """sampler.py:34-69 -> (img_idxs, pix_idxs), int32 (batch_size,)"""
        img_idxs = self._imgIdxs(batch_size=batch_size, img_strategy=sampling_strategy["imgs"], elapse_time=elapse_time)
        pix_idxs = self._pixIdxs(pix_strategy=sampling_strategy["pixs"], img_idxs=img_idxs)
        return img_idxs, pix_idxs

    def getValidImgIdxs(self, elapse_time: float):
        """sampler.py:71-94"""
        valid_img_idxs = torch.arange(self.dataset_len, device=self.args.device, dtype=torch.int32)
        if getattr(self.args.training, "real_time_simulation", False):
            mask = (self.times <= elapse_time)
            valid_img_idxs = valid_img_idxs[mask]
        if valid_img_idxs.shape[0] == 0:
            self.args.logger.error("no valid images found")
            sys.exit()
        return valid_img_idxs

    def _imgIdxs(self, batch_size: int, img_strategy: str, elapse_time: float):
        """sampler.py:96-125"""
        valid_img_idxs = self.getValidImgIdxs(elapse_time=elapse_time)
        if img_strategy == "all":
            idxs = torch.randint(0, valid_img_idxs.shape[0], size=(batch_size,), device=self.args.device, dtype=torch.int32)
            return valid_img_idxs[idxs.long()]
        if img_strategy == "same":
            idx = torch.randint(0, valid_img_idxs.shape[0], size=(1,), device=self.args.device, dtype=torch.int32)
            img_idx = valid_img_idxs[idx.long()]
            return img_idx * torch.ones(batch_size, device=self.args.device, dtype=torch.int32)
        self.args.logger.error(f"image sampling strategy must be either 'all' or 'same' but is {img_strategy}")

    def _pixIdxs(self, pix_strategy, img_idxs: torch.Tensor = None):
        """sampler.py:127-205"""
        pix_strategy = copy.deepcopy(pix_strategy)
        if pix_strategy == "entire_img":
            return self._pixStrategyEntireImg()
        if isinstance(pix_strategy, str):
            pix_strategy = {pix_strategy: 1.0}
        if getattr(self.args.training, "debug_mode", False):
            share_sum = sum(pix_strategy.values())
            if share_sum < 0.0 or share_sum > 1.0:
                self.args.logger.error(f"ray sampling strategy shares must be between 0 and 1 but sum is {share_sum}")
                return None
        B_sum = 0
        for strategy, share in pix_strategy.items():                    # :158-162
            B = int(share * img_idxs.shape[0])
            pix_strategy[strategy] = B
            B_sum += B
        B_rest = int(img_idxs.shape[0] - B_sum)                         # :165-167
        if B_rest > 0:
            pix_strategy["random"] = B_rest
        B_sum = 0
        pix_idxs = -1 * torch.ones(img_idxs.shape[0], device=self.args.device, dtype=torch.int32)
        for strategy, B in pix_strategy.items():
            if strategy == "random":
                tmp = self._pixStrategyRandom(B=B)
            elif strategy == "closest":
                tmp = self._pixStrategyClosest(img_idxs=img_idxs[B_sum:B_sum + B])
            elif strategy == "valid_uss":
                tmp = self._pixStrategyValidDepth(B=B, sensor_type="USS")
            elif strategy == "valid_tof":
                tmp = self._pixStrategyValidDepth(B=B, sensor_type="ToF")
            else:
                self.args.logger.error(f"ray sampling strategy = {strategy} not implemented")
                continue
            pix_idxs[B_sum:B_sum + B] = tmp
            B_sum += B
        return pix_idxs

    def _pixStrategyRandom(self, B: int):
        """sampler.py:207-218"""
        WH = self.img_wh[0] * self.img_wh[1]
        return torch.randint(0, WH, size=(B,), device=self.args.device, dtype=torch.int32)

    def _pixStrategyEntireImg(self):
        """sampler.py:220-228"""
        WH = self.img_wh[0] * self.img_wh[1]
        return torch.arange(0, WH, device=self.args.device, dtype=torch.int32)

    def _pixStrategyClosest(self, img_idxs: torch.tensor):
        """sampler.py:230-244"""
        pix_idxs, _, _ = self.sensors_dict["USS"].getStatsForBatch(batch_img_idxs=img_idxs)
        return pix_idxs

    def _pixStrategyValidDepth(self, B: int, sensor_type: str):
        """sampler.py:246-262; torch.where(mask) is cached per sensor (the masks are static)"""
        mask_idxs = self._mask_idxs.get(sensor_type)
        if mask_idxs is None:
            mask = self.sensors_dict[sensor_type].mask
            mask_idxs = self._mask_idxs[sensor_type] = torch.where(mask)[0].to(self.args.device)
        rand_ints = torch.randint(0, mask_idxs.shape[0], (B,), device=self.args.device, dtype=torch.int32)
        return mask_idxs[rand_ints.long()]
