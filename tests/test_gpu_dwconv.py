"""-m gpu: the DaViT depth-wise 3x3 position-encoding conv (csrc/conv.cu, fo1_dwconv3x3_residual) against a plain torch fp32
restatement of PreNorm(None, DepthWiseConv2d) (modeling_davit.py:29-48, 72-99): y = x + bf16(conv(x) + bias), both kernel variants
(the strip kernel for C % 256 == 0 and the general one) at ragged sizes, and the two variants bit-identical to each other."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, w9, bias):
    B, H, W, C = x.shape
    w = w9.float().t().reshape(C, 1, 3, 3)
    conv = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w, bias.float(), padding=1, groups=C)
    return (x.float() + conv.permute(0, 2, 3, 1).bfloat16().float()).bfloat16()


@pytest.mark.parametrize("shape", [(2, 48, 48, 256), (1, 37, 29, 512), (3, 24, 24, 1024), (2, 7, 9, 2048), (2, 13, 11, 64), (1, 1, 1, 256), (1, 50, 3, 256)])
def test_dwconv_matches_torch(shape, monkeypatch):
    from importlib import import_module
    import fo1_b200  # noqa: F401
    ops = import_module("vlm-fo1_b200.ops")
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(shape, device="cuda", generator=g).bfloat16()
    w9 = (torch.randn(9, shape[3], device="cuda", generator=g) * 0.3).bfloat16()
    bias = (torch.randn(shape[3], device="cuda", generator=g) * 0.1).bfloat16()
    y = ops.dwconv3x3_residual(x, w9, bias)
    ref = _ref(x, w9, bias)
    # fp32 accumulation of 9 products in a different order than cuDNN: at most one bf16 ulp of the conv term, then the same add
    err = (y.float() - ref.float()).abs()
    tol = 2.0 ** -7 * (ref.float().abs() + (ref.float() - x.float()).abs()) + 1e-6
    assert bool((err <= tol).all()), float((err - tol).max())
    assert float((y != ref).float().mean()) < 0.02
    monkeypatch.setenv("FO1_DWCONV_SIMPLE", "1")
    y2 = ops.dwconv3x3_residual(x, w9, bias)
    assert torch.equal(y, y2)
