"""Pin the HFRE oracle against outputs of the reference's own HFREModule (tests/golden/hfre_*.npz)
and against its literal numpy restatement of roi_align / bilinear up-sampling."""
import os

import numpy as np
import pytest
import torch

from oracle import hfre as O


def _load(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, f"hfre_{tag}.npz"))
    aux = [torch.from_numpy(z[f"aux{i}"]) for i in range(4)]
    taps = [torch.from_numpy(z[f"tap{i}"]) for i in range(4)]
    pyr = [torch.from_numpy(z[f"pyr{i}"]) for i in range(4)]
    return z, aux, taps, pyr


@pytest.mark.parametrize("tag", ["small", "rect"])
def test_oracle_matches_reference_module(golden_dir, tag):
    z, aux, taps, pyr = _load(golden_dir, tag)
    boxes = torch.from_numpy(z["boxes"]); vt_boxes = torch.from_numpy(z["vt_boxes"])
    gh, gw = (int(v) for v in z["grid_hw"])
    out_a = O.hfre_forward(aux, boxes, taps, vt_boxes, vt_mode="concat", vt_grid_hw=(gh, gw))
    out_b = O.hfre_forward(aux, boxes, pyr, vt_boxes, vt_mode="fpn", vt_grid_hw=(gh, gw))
    # same third-party kernels, same op order: expect (near) bit equality
    np.testing.assert_allclose(out_a.numpy(), z["out_concat"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(out_b.numpy(), z["out_fpn"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("tag", ["small", "rect"])
def test_literal_roi_align_and_upsample(golden_dir, tag):
    z, aux, taps, pyr = _load(golden_dir, tag)
    boxes = z["boxes"]
    H, W = aux[0].shape[1:]
    up = [aux[0].numpy()] + [O.upsample_bilinear_literal(a.numpy(), H, W) for a in aux[1:]]
    ref_up = torch.nn.functional.interpolate(aux[2].unsqueeze(0), size=(H, W), mode="bilinear", align_corners=False)[0]
    np.testing.assert_allclose(up[2], ref_up.numpy(), rtol=1e-6, atol=1e-6)
    lit = O.roi_align_avg_literal(np.concatenate(up, 0), boxes, 0.25)
    ref = O.hfre_aux_branch(aux, torch.from_numpy(boxes))
    np.testing.assert_allclose(lit, ref.numpy(), rtol=2e-5, atol=2e-6)
    # edge rows of the fixture: degenerate / border boxes must agree too (skip rule y<-1 or y>H)
    lit_vt = O.roi_align_avg_literal(np.concatenate([t.numpy() for t in taps], 0), z["vt_boxes"], 1.0 / 14.0)
    ref_vt = O.hfre_vt_branch(taps, torch.from_numpy(z["vt_boxes"]), "concat")
    np.testing.assert_allclose(lit_vt, ref_vt.numpy(), rtol=2e-5, atol=2e-6)


def test_sine_embedding_layout():
    b = torch.tensor([[10.0, 20.0, 50.0, 80.0]])
    e = O.sine_box_embedding(b, 100.0, 200.0, 32)[0]
    q = 8
    cy = (20 / 200 + 80 / 200) / 2
    assert abs(e[0].item() - np.sin(2 * np.pi * cy)) < 1e-6       # first block is cy, sin first
    assert abs(e[1].item() - np.cos(2 * np.pi * cy)) < 1e-6
    cx = (10 / 100 + 50 / 100) / 2
    assert abs(e[q].item() - np.sin(2 * np.pi * cx)) < 1e-6
