"""-m gpu: device-side pre-processing (fo1_resize_bicubic_u8 / fo1_preprocess_primary_u8 / fo1_preprocess_aux_u8) vs the host
processors of the boundary mirror (themselves pinned to the reference's processors by tests/test_boundary_processors.py):
BIT-EXACT -- the resize is integer arithmetic, the normalisation follows the processors' float32 operation order."""
import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(399, 500), (640, 640), (896, 896), (37, 911), (1344, 1344), (56, 56)])
def test_device_preprocessing_is_bit_exact(hw):
    from importlib import import_module
    import fo1_b200  # noqa: F401
    PP = import_module("vlm-fo1_b200.preprocess")
    from vlm_fo1.processors import AuxImageProcessor, PrimaryImageProcessor
    img = np.random.default_rng(hw[0] + hw[1]).integers(0, 256, tuple(hw) + (3,), dtype=np.uint8)
    pil = Image.fromarray(img)
    dp = PP.DevicePreprocessor("cuda")
    d = torch.from_numpy(img).cuda()
    px, grid = dp.primary(d)
    ref = PrimaryImageProcessor().preprocess(pil, return_tensors="pt")
    assert list(grid) == ref["image_grid_thw"][0, 1:].tolist()
    assert torch.equal(px.cpu(), ref["pixel_values"])
    for size, mode in ((0, "dynamic"), (768, "squash")):
        a = dp.aux(d, size)
        r = AuxImageProcessor(768, mode).preprocess(pil, return_tensors="pt")["pixel_values"][0]
        assert torch.equal(a.cpu(), r), mode
