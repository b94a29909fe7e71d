"""Device-resident pixel sampling for one view (datasets/scene_dataset.py:68-117 `__getitem__`, `collate_fn`,
`change_sampling_idx`).  The reference builds the full [H*W, 2] uv lattice on the host for every item, indexes it
with a host `randperm` and copies the slices; here the image, mask and lattice of every view live on the device
and a batch is three gathers.  Image decoding / camera loading (rend_util.load_rgb, load_K_Rt_from_P) stay with the
caller: this class takes ready tensors.
"""
from typing import Dict, Optional, Tuple

import numpy as np
import torch


def uv_lattice(img_res) -> torch.Tensor:
    """[H*W, 2] pixel coordinates (x, y) in the reference's order (scene_dataset.py:72-74)."""
    uv = np.mgrid[0:img_res[0], 0:img_res[1]].astype(np.int32)
    uv = torch.from_numpy(np.flip(uv, axis=0).copy()).float()
    return uv.reshape(2, -1).transpose(1, 0).contiguous()


class DevicePixelSampler:
    def __init__(self, rgb_images: torch.Tensor, object_masks: torch.Tensor, intrinsics: torch.Tensor,
                 poses: torch.Tensor, img_res, device="cuda"):
        """rgb_images [V, H*W, 3] in [-1, 1], object_masks [V, H*W] bool, intrinsics / poses [V, 4, 4]."""
        self.img_res = tuple(img_res)
        self.total_pixels = self.img_res[0] * self.img_res[1]
        if rgb_images.shape[1] != self.total_pixels or object_masks.shape[1] != self.total_pixels:
            raise ValueError("images must be flattened to H*W = %d pixels" % self.total_pixels)
        self.rgb = rgb_images.to(device)
        self.masks = object_masks.to(device)
        self.intrinsics = intrinsics.to(device)
        self.poses = poses.to(device)
        self.uv = uv_lattice(self.img_res).to(device)
        self.n_images = self.rgb.shape[0]
        self.sampling_idx: Optional[torch.Tensor] = None

    def __len__(self):
        return self.n_images

    def change_sampling_idx(self, sampling_size: int, generator: Optional[torch.Generator] = None):
        """-1: all pixels; else a fresh random subset, drawn with the HOST generator like the reference (:113-117)."""
        if sampling_size == -1:
            self.sampling_idx = None
        else:
            self.sampling_idx = torch.randperm(self.total_pixels, generator=generator)[:sampling_size].to(self.uv.device)

    def batch(self, idx) -> Tuple[torch.Tensor, Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
        """(indices, model_input, ground_truth) for the views `idx`, shaped like the collated reference batch."""
        idx = torch.as_tensor(idx, dtype=torch.long, device=self.uv.device).reshape(-1)
        sel = self.sampling_idx
        uv = self.uv if sel is None else self.uv[sel]
        rgb = self.rgb[idx] if sel is None else self.rgb[idx][:, sel]
        mask = self.masks[idx] if sel is None else self.masks[idx][:, sel]
        sample = {"object_mask": mask, "uv": uv.unsqueeze(0).expand(idx.numel(), -1, -1).contiguous(),
                  "intrinsics": self.intrinsics[idx], "pose": self.poses[idx]}
        return idx, sample, {"rgb": rgb}
