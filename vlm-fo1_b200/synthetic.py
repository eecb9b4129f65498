"""Synthetic workloads of BASELINE.json: seeded random images run through numpy
restatements of the two host-side preprocessors' arithmetic (normalise + patchify; sizes are already multiples
of 28 so no resize happens), seeded boxes and placeholder prompts.  No dataset, no tokenizer, no network."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from .pipeline import SampleInputs

CLIP_MEAN = np.array([0.48145466, 0.4578275, 0.40821073], dtype=np.float32)      # Qwen2VLImageProcessor defaults
CLIP_STD = np.array([0.26862954, 0.26130258, 0.27577711], dtype=np.float32)
IMNET_MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)                    # davit/configs.py:139-152
IMNET_STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def smart_size(S: int, factor: int = 28) -> int:
    return max(factor, int(round(S / factor)) * factor)


def patchify_primary(img_u8: np.ndarray, patch: int = 14, merge: int = 2, temporal: int = 2):
    """uint8 [H, W, 3] (H, W multiples of 28) -> (pixel_values fp32 [gh*gw, 3*temporal*patch*patch], (gh, gw)) in the
    processor's 2x2-merge order (Qwen2VLImageProcessor._preprocess)."""
    x = (img_u8.astype(np.float32) / 255.0 - CLIP_MEAN) / CLIP_STD
    x = x.transpose(2, 0, 1)                                                         # [3, H, W]
    H, W = x.shape[1:]
    gh, gw = H // patch, W // patch
    x = np.broadcast_to(x[None], (temporal,) + x.shape)                              # the frame repeated over the temporal patch
    x = x.reshape(1, temporal, 3, gh // merge, merge, patch, gw // merge, merge, patch)
    x = x.transpose(0, 3, 6, 4, 7, 2, 1, 5, 8)
    return np.ascontiguousarray(x.reshape(gh * gw, 3 * temporal * patch * patch)), (gh, gw)


def normalise_aux(img_u8: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(((img_u8.astype(np.float32) / 255.0 - IMNET_MEAN) / IMNET_STD).transpose(2, 0, 1))


def synthetic_boxes(i: int, S: int, n: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(2000 + i)
    w = torch.rand(n, generator=g) * (S / 2 - 32) + 32
    h = torch.rand(n, generator=g) * (S / 2 - 32) + 32
    x1 = torch.rand(n, generator=g) * (S - w)
    y1 = torch.rand(n, generator=g) * (S - h)
    return torch.stack([x1, y1, x1 + w, y1 + h], dim=1).to(torch.float32)


def synthetic_prompt(i: int, n_boxes: int) -> List[int]:
    g = torch.Generator().manual_seed(3000 + i)
    t = lambda k: torch.randint(0, 151643, (k,), generator=g).tolist()
    ids = [151644] + t(3) + [151645] + t(1) + [151644] + t(2) + [151652, -200, 151653] + t(1)
    for _ in range(n_boxes):
        ids += t(1) + [-300]
    ids += t(1) + t(25) + [151645] + t(1) + [151644] + t(2)
    return ids


def synthetic_sample(i: int, S: int, n_boxes: int) -> SampleInputs:
    rng = np.random.default_rng(1000 + i)
    Sp = smart_size(S)
    img = rng.integers(0, 256, size=(S, S, 3), dtype=np.uint8)
    # the primary processor resizes to a multiple of 28 (bicubic); for synthetic noise a fresh draw at that size is equivalent
    img_p = img if Sp == S else rng.integers(0, 256, size=(Sp, Sp, 3), dtype=np.uint8)
    px, grid = patchify_primary(img_p)
    return SampleInputs(input_ids=synthetic_prompt(i, n_boxes), pixel_values=torch.from_numpy(px), grid_hw=grid,
                        image_aux=torch.from_numpy(normalise_aux(img)), boxes=synthetic_boxes(i, S, n_boxes))


def synthetic_batch(start: int, count: int, S: int, n_boxes: int) -> List[SampleInputs]:
    return [synthetic_sample(start + k, S, n_boxes) for k in range(count)]


def synthetic_sample_u8(i: int, S: int, n_boxes: int) -> SampleInputs:
    """The same seeded image / boxes / prompt as synthetic_sample, handed over as the raw uint8 image: the pipeline's device-side
    pre-processing then does what the two host processors do (incl. the bicubic smart-resize when S is not a multiple of 28)."""
    rng = np.random.default_rng(1000 + i)
    img = rng.integers(0, 256, size=(S, S, 3), dtype=np.uint8)
    return SampleInputs(input_ids=synthetic_prompt(i, n_boxes), pixel_values=None, grid_hw=None, image_aux=None,
                        boxes=synthetic_boxes(i, S, n_boxes), image_u8=torch.from_numpy(img))
