"""Device-side image pre-processing over the C ABI (fo1_resize_bicubic_u8 / fo1_preprocess_primary_u8 / fo1_preprocess_aux_u8):
uint8 RGB images stay uint8 across PCIe (12x fewer bytes than the fp32 tensors the host processors emit) and are resized,
normalised and patchified on the GPU, bit-identically to ``Qwen2VLImageProcessor`` (mm_utils.py:615) and the reference's
``CLIPImageProcessor`` (davit/image_processing_clip.py:222-367).

The only host arithmetic is PIL's coefficient table (``ImagingResample``: bicubic a = -0.5, support scaled by the shrink
factor, normalised in double precision, rounded to 22 fractional bits) -- a few KB per (input size, output size), cached."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Tuple

import numpy as np
import torch

from ._lib import check, lib

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
IMNET_MEAN = (0.485, 0.456, 0.406)
IMNET_STD = (0.229, 0.224, 0.225)
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float, a: float = -0.5) -> float:
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coefficients(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """PIL ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the bicubic filter -> (bounds int32 [out, 2], coef int32 [out, ksize], ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    coef = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            coef[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, coef, ksize


def resize_bicubic_u8_model(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """numpy statement of the two integer passes the kernels run (CPU test model: must equal PIL byte for byte)."""
    x = img.astype(np.int64)
    H, W = img.shape[:2]
    half = 1 << (PRECISION_BITS - 1)
    if out_w != W:
        b, kk, _ = resample_coefficients(W, out_w)
        o = np.empty((H, out_w, 3), np.int64)
        for xx in range(out_w):
            lo, n = b[xx]
            o[:, xx] = (x[:, lo:lo + n] * kk[xx, :n][None, :, None]).sum(1) + half
        x = np.clip(o >> PRECISION_BITS, 0, 255)
    if out_h != H:
        b, kk, _ = resample_coefficients(H, out_h)
        o = np.empty((out_h, x.shape[1], 3), np.int64)
        for yy in range(out_h):
            lo, n = b[yy]
            o[yy] = (x[lo:lo + n] * kk[yy, :n][:, None, None]).sum(0) + half
        x = np.clip(o >> PRECISION_BITS, 0, 255)
    return x.astype(np.uint8)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class DevicePreprocessor:
    def __init__(self, device, patch: int = 14, merge: int = 2, temporal: int = 2, min_pixels: int = 56 * 56, max_pixels: int = 2048 * 2048):
        self.device = torch.device(device)
        self.patch, self.merge, self.temporal = patch, merge, temporal
        self.min_pixels, self.max_pixels = min_pixels, max_pixels
        self._tables: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor, int]] = {}
        L = lib()
        for fn in ("fo1_resize_bicubic_u8", "fo1_preprocess_primary_u8", "fo1_preprocess_aux_u8"):
            getattr(L, fn).restype = C.c_int

    def _table(self, n_in: int, n_out: int):
        key = (n_in, n_out)
        if key not in self._tables:
            b, k, ks = resample_coefficients(n_in, n_out)
            self._tables[key] = (torch.from_numpy(b).to(self.device), torch.from_numpy(k).to(self.device), ks)
        return self._tables[key]

    def resize(self, img: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
        """uint8 [H, W, 3] on the device -> uint8 [out_h, out_w, 3] (PIL bicubic)."""
        assert img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 3 and img.is_cuda and img.is_contiguous()
        H, W = img.shape[:2]
        if (H, W) == (out_h, out_w):
            return img
        out = torch.empty((out_h, out_w, 3), dtype=torch.uint8, device=img.device)
        hb = hk = vb = vk = None
        hks = vks = 0
        if out_w != W:
            hb, hk, hks = self._table(W, out_w)
        if out_h != H:
            vb, vk, vks = self._table(H, out_h)
        tmp = torch.empty((H, out_w, 3), dtype=torch.uint8, device=img.device) if (hb is not None and vb is not None) else None
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        check(lib().fo1_resize_bicubic_u8(p(img), H, W, out_h, out_w, p(hb), p(hk), hks, p(vb), p(vk), vks, p(tmp), p(out), _stream()),
              "fo1_resize_bicubic_u8")
        return out

    def primary(self, img: torch.Tensor):
        """uint8 [H, W, 3] -> (pixel_values fp32 [gh*gw, 3*temporal*patch^2], (gh, gw)) like Qwen2VLImageProcessor."""
        from vlm_fo1.processors import smart_resize
        H, W = img.shape[:2]
        h, w = smart_resize(H, W, self.patch * self.merge, self.min_pixels, self.max_pixels)
        img = self.resize(img.contiguous(), h, w)
        gh, gw = h // self.patch, w // self.patch
        out = torch.empty((gh * gw, 3 * self.temporal * self.patch * self.patch), dtype=torch.float32, device=img.device)
        mean = (C.c_float * 3)(*CLIP_MEAN); std = (C.c_float * 3)(*CLIP_STD)
        check(lib().fo1_preprocess_primary_u8(C.c_void_p(img.data_ptr()), h, w, self.patch, self.merge, self.temporal, mean, std,
                                              C.c_void_p(out.data_ptr()), _stream()), "fo1_preprocess_primary_u8")
        return out, (gh, gw)

    def aux(self, img: torch.Tensor, size: int = 0) -> torch.Tensor:
        """uint8 [H, W, 3] -> fp32 [3, H', W']: ``size`` > 0 squashes to size x size first (aspect ratio 'squash'); 0 = 'dynamic'."""
        if size > 0:
            img = self.resize(img.contiguous(), size, size)
        H, W = img.shape[:2]
        out = torch.empty((3, H, W), dtype=torch.float32, device=img.device)
        mean = (C.c_float * 3)(*IMNET_MEAN); std = (C.c_float * 3)(*IMNET_STD)
        check(lib().fo1_preprocess_aux_u8(C.c_void_p(img.contiguous().data_ptr()), H, W, mean, std, C.c_void_p(out.data_ptr()), _stream()),
              "fo1_preprocess_aux_u8")
        return out
