"""Host-side pre-processing mirror (vlm_fo1/processors.py) vs fixtures produced by the reference's own CLIPImageProcessor
(aux tower, both resize modes) and by transformers' Qwen2VLImageProcessor (primary tower) -- oracle/gen_golden.py
stage ``processors``.  CPU only."""
import os

import numpy as np
import pytest
from PIL import Image

from vlm_fo1.processors import AuxImageProcessor, PrimaryImageProcessor, smart_resize


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "processors.npz"))


def _images(gold):
    rng = np.random.default_rng(11)
    return [Image.fromarray(rng.integers(0, 256, tuple(hw) + (3,), dtype=np.uint8)) for hw in gold["sizes"]]


@pytest.mark.parametrize("mode", ["dynamic", "squash"])
def test_aux_processor_matches_reference_clip_processor(gold, mode):
    """davit/image_processing_clip.py:222-367 with davit/configs.py:139-152: same PIL bicubic squash, /255, ImageNet mean/std;
    only the fp32 operation order differs (<= 1e-6)."""
    for i, img in enumerate(_images(gold)):
        got = AuxImageProcessor(768, mode).preprocess(img, return_tensors="pt")["pixel_values"][0].numpy()
        assert list(got.shape) == gold[f"aux_{mode}{i}_shape"].tolist()
        assert np.abs(got[:, ::7, ::5] - gold[f"aux_{mode}{i}_sample"]).max() <= 2e-6
        assert abs(got.astype(np.float64).sum() - float(gold[f"aux_{mode}{i}_sum"])) <= 1e-6 * got.size


def test_primary_processor_grid_patch_order_and_values(gold):
    """Qwen2VLImageProcessor (qwen2_5_vl_encoder.py:210-225): grid and patch order always exact; values exact (fp32 rounding)
    when smart_resize keeps the size, within 2 uint8 steps of the installed transformers' resize otherwise."""
    lsb = 1.0 / 255.0 / 0.26130258           # one uint8 step after CLIP normalisation (smallest std)
    for i, img in enumerate(_images(gold)):
        out = PrimaryImageProcessor().preprocess(img, videos=None, return_tensors="pt")
        px = out["pixel_values"].numpy()
        assert out["image_grid_thw"].numpy().tolist() == gold[f"prim{i}_grid"].tolist()
        assert list(px.shape) == gold[f"prim{i}_shape"].tolist()
        h, w = smart_resize(img.height, img.width)
        d = np.abs(px[::3, ::11] - gold[f"prim{i}_sample"])
        if (h, w) == (img.height, img.width):
            assert d.max() <= 1e-6
        else:
            assert d.max() <= 2.0 * lsb + 1e-6 and d.mean() <= 0.02 * lsb
        assert abs(px.astype(np.float64).sum() - float(gold[f"prim{i}_sum"])) <= 2e-4 * px.size


def test_smart_resize_contract():
    for h, w in [(399, 500), (20, 20), (3000, 2500), (37, 911), (4096, 4096), (56, 56)]:
        hb, wb = smart_resize(h, w)
        assert hb % 28 == 0 and wb % 28 == 0 and 56 * 56 <= hb * wb <= 2048 * 2048
    with pytest.raises(ValueError):
        smart_resize(10, 4000)
