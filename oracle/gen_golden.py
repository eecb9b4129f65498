"""Generate ``tests/golden/*.npz`` by running the REFERENCE's own modules -- TEST INFRASTRUCTURE ONLY.

Run in the build container (needs ``/root/reference``):  ``python oracle/gen_golden.py [stage ...]``
Stages: hfre fpn davit vit proj llm splice (default: all that are implemented).
The fixtures are small, committed, and replayed by ``tests/`` on machines without the reference.
Inputs are stored next to the outputs so the tests do not depend on RNG reproducibility.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
# make sure ``import vlm_fo1`` can only resolve to the reference, never to the repo's boundary mirror
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != REPO]
sys.path.insert(0, HERE)

import numpy as np
import torch

import ref_shim

GOLD = os.path.join(REPO, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def save(name: str, **arrays) -> None:
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


# ------------------------------------------------------------------------------------------ HFRE
def hfre_boxes(S: int, n_rand: int, gen: torch.Generator) -> torch.Tensor:
    """random interior boxes + the border / degenerate cases SURVEY.md appendix A lists."""
    w = torch.rand(n_rand, generator=gen) * (S / 2 - 4) + 4
    h = torch.rand(n_rand, generator=gen) * (S / 2 - 4) + 4
    x1 = torch.rand(n_rand, generator=gen) * (S - w)
    y1 = torch.rand(n_rand, generator=gen) * (S - h)
    rnd = torch.stack([x1, y1, x1 + w, y1 + h], dim=1)
    S_ = float(S)
    edge = torch.tensor([
        [0.0, 0.0, S_, S_],                     # whole image
        [0.0, 0.0, 1.0, 1.0],                   # < one cell, top-left corner
        [S_ - 0.1, S_ - 0.1, S_, S_],           # degenerate, bottom-right corner (max(size,1) pushes samples out)
        [2.0, S_ - 2.5, 9.0, S_ - 0.1],         # thin, on the bottom border
        [S_ - 0.8, 10.0, S_, 30.0],             # thin, on the right border
        [10.0, 10.0, 10.0, 10.0],               # zero-area
        [5.3, 7.7, 5.9, S_ - 3.0],              # very thin vertical
        [0.0, S_ / 2, S_, S_ / 2 + 0.5],        # very thin horizontal, full width
    ])
    return torch.cat([rnd, edge], dim=0).to(torch.float32)


def gen_hfre() -> None:
    hf, _ = ref_shim.load_hfre_only()
    gen = torch.Generator().manual_seed(1234)
    for tag, S, vt_S, chans, vt_c in (("small", 96, 98, (8, 16, 32, 64), 24), ("rect", 128, 126, (16, 16, 32, 32), 16)):
        # aux pyramid at strides 4/8/16/32 of an S x (S or 3S/4) image
        Sh, Sw = (S, S) if tag == "small" else (S, (S * 3) // 4)
        aux = [bf16_round(torch.randn(1, c, Sh // (4 << i), Sw // (4 << i), generator=gen)) for i, c in enumerate(chans)]
        boxes = hfre_boxes(min(Sh, Sw), 12, gen)
        gh, gw = vt_S // 14, (vt_S // 14 if tag == "small" else (vt_S // 14) - 2)
        sx, sy = (gw * 14) / Sw, (gh * 14) / Sh
        vt_boxes = boxes * torch.tensor([sx, sy, sx, sy])
        # variant A: four tap maps, concatenated by the module
        taps = [bf16_round(torch.randn(1, vt_c, gh, gw, generator=gen)) for _ in range(4)]
        D_a = sum(chans) + 4 * vt_c
        mod_a = hf.HFREModule(roi_output_size=7, region_feature_dim=D_a, apply_position_embedding=True,
                              pos_embedding_strategy="bbox_based", use_vision_tower_region_feature=True,
                              region_feature_combination="concat", vision_tower_region_feature_dim=4 * vt_c,
                              vision_tower_spatial_scale=1 / 14, use_simpleFPN_for_vt=False,
                              aux_vision_tower_spatial_scale=0.25)
        out_a = mod_a(aux_multi_level_features=[a.clone() for a in aux], aux_boxes=[boxes.clone()],
                      vt_multi_level_features=[t.clone() for t in taps], vt_boxes=[vt_boxes.clone()])
        # variant B: the module's SimpleFP is replaced by a fixed pyramid (the FPN itself is pinned
        # by the separate ``fpn`` fixture); everything after it is the reference's own code
        pyr = [bf16_round(torch.randn(1, vt_c, int(gh * f), int(gw * f), generator=gen)) for f in (4, 2, 1, 0.5)]
        D_b = sum(chans) + 4 * vt_c
        mod_b = hf.HFREModule(roi_output_size=7, region_feature_dim=D_b, apply_position_embedding=True,
                              pos_embedding_strategy="bbox_based", use_vision_tower_region_feature=True,
                              region_feature_combination="concat", vision_tower_region_feature_dim=4 * vt_c,
                              vision_tower_spatial_scale=1 / 14, use_simpleFPN_for_vt=False,
                              aux_vision_tower_spatial_scale=0.25)
        mod_b.use_simpleFPN_for_vt = True
        mod_b.simple_fpn = lambda x, _p=pyr: [p.clone() for p in _p]
        last_tap = taps[-1]
        out_b = mod_b(aux_multi_level_features=[a.clone() for a in aux], aux_boxes=[boxes.clone()],
                      vt_multi_level_features=last_tap.clone(), vt_boxes=[vt_boxes.clone()])
        # (aux-only, use_vision_tower_region_feature=False, raises UnboundLocalError in the reference
        #  at hybrid_finegrained_region_encoder.py:456 -- dead configuration, not pinned)
        save(f"hfre_{tag}", boxes=boxes, vt_boxes=vt_boxes, grid_hw=np.array([gh, gw]),
             **{f"aux{i}": a[0] for i, a in enumerate(aux)}, **{f"tap{i}": t[0] for i, t in enumerate(taps)},
             **{f"pyr{i}": p[0] for i, p in enumerate(pyr)},
             out_concat=out_a[0], out_fpn=out_b[0])


STAGES = {"hfre": gen_hfre}

if __name__ == "__main__":
    torch.set_grad_enabled(False)
    wanted = sys.argv[1:] or list(STAGES)
    for s in wanted:
        STAGES[s]()
