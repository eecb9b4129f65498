"""Host-side image processors with the ``.preprocess(image, ..., return_tensors="pt")`` surface the callers use
(``mm_utils.prepare_inputs``): the arithmetic of Qwen2VLImageProcessor (smart-resize to multiples of 28, bicubic, /255,
CLIP mean/std, 2x2-merge patch order; called at mm_utils.py:615) and of the reference's CLIPImageProcessor in the
aux tower (davit/image_processing_clip.py:222-367, config davit/configs.py:139-152; ``dynamic`` = no resize), written
against PIL + numpy so they do not depend on the installed transformers version."""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch
from PIL import Image

CLIP_MEAN = np.array([0.48145466, 0.4578275, 0.40821073], dtype=np.float32)
CLIP_STD = np.array([0.26862954, 0.26130258, 0.27577711], dtype=np.float32)
IMNET_MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
IMNET_STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def smart_resize(height: int, width: int, factor: int = 28, min_pixels: int = 56 * 56, max_pixels: int = 2048 * 2048):
    """Qwen2-VL ``smart_resize``: both sides multiples of ``factor``, pixel count within [min, max], aspect kept."""
    if max(height, width) / min(height, width) > 200:
        raise ValueError("absolute aspect ratio must be smaller than 200")
    h_bar = max(factor, round(height / factor) * factor)
    w_bar = max(factor, round(width / factor) * factor)
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, math.floor(height / beta / factor) * factor)
        w_bar = max(factor, math.floor(width / beta / factor) * factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = math.ceil(height * beta / factor) * factor
        w_bar = math.ceil(width * beta / factor) * factor
    return h_bar, w_bar


def _device_preprocessor(device, *args):
    """``device`` given -> the CUDA pre-processing of the engine (fo1_preprocess_*); None -> the PIL / numpy host path."""
    if device is None:
        return None
    from importlib import import_module
    import fo1_b200  # noqa: F401
    return import_module("vlm-fo1_b200.preprocess").DevicePreprocessor(device, *args)


class PrimaryImageProcessor:
    """-> {'pixel_values': fp32 [gh*gw, 3*2*14*14], 'image_grid_thw': int64 [1, 3]}"""

    def __init__(self, patch_size: int = 14, merge_size: int = 2, temporal_patch_size: int = 2, min_pixels: int = 56 * 56,
                 max_pixels: int = 2048 * 2048, device=None):
        self.patch_size, self.merge_size, self.temporal_patch_size = patch_size, merge_size, temporal_patch_size
        self.min_pixels, self.max_pixels = min_pixels, max_pixels
        self._dev = _device_preprocessor(device, patch_size, merge_size, temporal_patch_size, min_pixels, max_pixels)

    def preprocess(self, images, videos=None, return_tensors="pt", **kwargs) -> Dict[str, torch.Tensor]:
        img = images.convert("RGB")
        if self._dev is not None:     # uint8 over PCIe, resize / normalise / patchify on the GPU (bit-identical, tests/test_gpu_preprocess.py)
            px, (gh, gw) = self._dev.primary(torch.from_numpy(np.asarray(img)).to(self._dev.device))
            return {"pixel_values": px, "image_grid_thw": torch.tensor([[1, gh, gw]], dtype=torch.int64)}
        p, m, t = self.patch_size, self.merge_size, self.temporal_patch_size
        h, w = smart_resize(img.height, img.width, p * m, self.min_pixels, self.max_pixels)
        if (h, w) != (img.height, img.width):
            img = img.resize((w, h), Image.Resampling.BICUBIC)
        x = (np.asarray(img, dtype=np.float32) / 255.0 - CLIP_MEAN) / CLIP_STD
        x = x.transpose(2, 0, 1)
        gh, gw = h // p, w // p
        x = np.broadcast_to(x[None], (t,) + x.shape).reshape(1, t, 3, gh // m, m, p, gw // m, m, p)
        x = np.ascontiguousarray(x.transpose(0, 3, 6, 4, 7, 2, 1, 5, 8).reshape(gh * gw, 3 * t * p * p))
        return {"pixel_values": torch.from_numpy(x), "image_grid_thw": torch.tensor([[1, gh, gw]], dtype=torch.int64)}

    __call__ = preprocess


class AuxImageProcessor:
    """-> {'pixel_values': fp32 [1, 3, H, W]}; ``dynamic`` keeps the image size, otherwise a bicubic squash to size x size."""

    def __init__(self, image_size: int = 768, aspect_ratio: str = "squash", device=None):
        self.image_size, self.aspect_ratio = image_size, aspect_ratio
        self.do_resize = aspect_ratio != "dynamic"
        self._dev = _device_preprocessor(device)

    def preprocess(self, images, return_tensors="pt", **kwargs) -> Dict[str, torch.Tensor]:
        img = images.convert("RGB")
        if self._dev is not None:
            x = self._dev.aux(torch.from_numpy(np.asarray(img)).to(self._dev.device), self.image_size if self.do_resize else 0)
            return {"pixel_values": x[None]}
        if self.do_resize and img.size != (self.image_size, self.image_size):
            img = img.resize((self.image_size, self.image_size), Image.Resampling.BICUBIC)
        x = (np.asarray(img, dtype=np.float32) * np.float32(1 / 255) - IMNET_MEAN) / IMNET_STD
        return {"pixel_values": torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 1)[None]))}

    __call__ = preprocess
