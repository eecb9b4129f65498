"""CPU: the integer restatement of PIL's bicubic resize that the device kernels implement (vlm-fo1_b200/preprocess.py::
resample_coefficients + the two integer passes) equals PIL byte for byte, at the BASELINE sizes' resize cases."""
import numpy as np
import pytest
from PIL import Image


@pytest.mark.parametrize("H,W,oh,ow", [(399, 500, 392, 504), (640, 640, 644, 644), (37, 911, 28, 924), (300, 210, 768, 768),
                                       (100, 100, 28, 28), (57, 64, 56, 56)])
def test_integer_resize_equals_pil(H, W, oh, ow):
    from importlib import import_module
    import fo1_b200  # noqa: F401
    PP = import_module("vlm-fo1_b200.preprocess")
    img = np.random.default_rng(H * 1000 + W).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.Resampling.BICUBIC))
    got = PP.resize_bicubic_u8_model(img, oh, ow)
    assert np.array_equal(ref, got)
