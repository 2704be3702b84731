"""Input/output specification objects a MargiPose model carries (`model.data_specs`).

Mirrors /root/reference/src/margipose/data_specs.py:6-64 as callers of the model use it: resolution, ImageNet
mean / stddev, skeleton, and the image conversion of the inference front-end (bin/infer_single.py:62-66 calls
`model.data_specs.input_specs.convert(image)` and `.unconvert(tensor)`).  `convert` is the host-side (CPU) form of the
input step; `convert_uint8` hands the raw pixels to the CUDA path instead, where /255 and the mean / stddev
normalisation run inside the stem conv's gather kernel (`model(uint8_batch)`, `InferStep(uint8=True)`).
"""
from collections.abc import Sequence

import torch


def normalize_pixels(tensor, mean, std):
    """(C, H, W) float tensor, in place: channel c becomes (x - mean[c]) / std[c] (data_specs.py:6-13)."""
    if mean is not None:
        for plane, m in zip(tensor, mean):
            plane.sub_(m)
    if std is not None:
        for plane, s in zip(tensor, std):
            plane.div_(s)
    return tensor


def denormalize_pixels(tensor, mean, std):
    """Inverse of normalize_pixels, in place (data_specs.py:16-23)."""
    if std is not None:
        for plane, s in zip(tensor, std):
            plane.mul_(s)
    if mean is not None:
        for plane, m in zip(tensor, mean):
            plane.add_(m)
    return tensor


def _rgb_bytes(img):
    """PIL image -> uint8 (H, W, 3) tensor of its RGB pixels."""
    import numpy as np
    if img.mode != 'RGB':
        img = img.convert('RGB')
    return torch.from_numpy(np.array(img, dtype=np.uint8, copy=True))


class ImageSpecs:
    IMAGENET_MEAN = [0.485, 0.456, 0.406]
    IMAGENET_STDDEV = [0.229, 0.224, 0.225]

    def __init__(self, resolution, mean=None, stddev=None):
        if isinstance(resolution, Sequence):
            self.height, self.width = resolution
        else:
            self.height = self.width = resolution
        self.mean = mean
        self.stddev = stddev

    def convert(self, img):
        """PIL image -> normalised float32 (3, H, W) tensor (data_specs.py:38-39: to_tensor, then mean / stddev)."""
        chw = _rgb_bytes(img).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
        return normalize_pixels(chw, self.mean, self.stddev)

    def convert_uint8(self, img):
        """PIL image -> raw uint8 (H, W, 3) pixels for the fused input step of the CUDA path."""
        return _rgb_bytes(img)

    def unconvert(self, tensor):
        """Normalised (3, H, W) tensor -> PIL RGB image (data_specs.py:41-42)."""
        import PIL.Image
        pixels = denormalize_pixels(tensor.detach().cpu().clone(), self.mean, self.stddev)
        hwc = pixels.mul(255).byte().permute(1, 2, 0).contiguous().numpy()
        return PIL.Image.fromarray(hwc, 'RGB')


class JointsSpecs:
    def __init__(self, skeleton_desc, n_dims=3):
        self.skeleton_desc = skeleton_desc
        self.n_dims = n_dims


class DataSpecs:
    def __init__(self, input_specs, output_specs):
        self._input_specs = input_specs
        self._output_specs = output_specs

    @property
    def input_specs(self):
        return self._input_specs

    @property
    def output_specs(self):
        return self._output_specs
