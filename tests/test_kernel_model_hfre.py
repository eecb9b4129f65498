"""The separable-window algorithm the CUDA kernels implement == the oracle's
interpolate -> concat -> roi_align -> mean (CPU-only check of the math)."""
import os

import numpy as np
import pytest
import torch

from oracle import hfre as O
from tests import kernel_models as KM


@pytest.mark.parametrize("tag", ["small", "rect"])
def test_separable_weights_equal_oracle(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, f"hfre_{tag}.npz"))
    aux = [z[f"aux{i}"] for i in range(4)]
    boxes = z["boxes"]
    H, W = aux[0].shape[1:]
    got = np.concatenate([KM.hfre_level(a.transpose(1, 2, 0), boxes, 0.25, 7, (H, W)) for a in aux], axis=1)
    ref = O.hfre_aux_branch([torch.from_numpy(a) for a in aux], torch.from_numpy(boxes)).numpy()
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 2e-5 * scale
    pyr = [z[f"pyr{i}"] for i in range(4)]
    vtb = z["vt_boxes"]
    got = np.concatenate([KM.hfre_level(p.transpose(1, 2, 0), vtb, 1.0 / s, 7, p.shape[1:]) for p, s in zip(pyr, O.FPN_STRIDES)], axis=1)
    ref = O.hfre_vt_branch([torch.from_numpy(p) for p in pyr], torch.from_numpy(vtb), "fpn").numpy()
    assert np.abs(got - ref).max() <= 2e-5 * np.abs(ref).max()
