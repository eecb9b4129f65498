"""-m gpu: the batched device-side integer bookkeeping (csrc/prompt.cu) equals the per-sample host functions bit for bit:
fo1_splice_plan_batch vs fo1_splice_plan (itself pinned to the reference's get_rope_index by tests/test_oracle_llm.py), and
fo1_parse_predictions vs the text regexes of vlm_fo1.mm_utils.extract_predictions_to_indexes."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _eng():
    from importlib import import_module
    import fo1_b200  # noqa: F401
    return import_module("vlm-fo1_b200.engine"), import_module("vlm-fo1_b200.synthetic")


def test_splice_plan_batch_matches_host_bit_exact():
    E, SY = _eng()
    rng = random.Random(3)
    t = lambda k: [rng.randrange(0, 151640) for _ in range(k)]
    prompts, grids, nreg = [], [], []
    # the BASELINE prompts (C3 / C4 / C5 geometry) ...
    for i, (S, n) in enumerate([(896, 64), (644, 100), (1344, 300), (448, 0)]):
        prompts.append(SY.synthetic_prompt(i, n)); grids.append([(S // 14, S // 14)]); nreg.append(n)
    # ... two images in one prompt, text between and after, regions; no image at all; image at the very start; a vision_start
    # that is NOT followed by an image (counts as text, modeling_qwen2_5_vl.py:1628-1630)
    prompts.append(t(5) + [151652, -200, 151653] + t(3) + [151652, -200, 151653] + [x for _ in range(7) for x in (t(1)[0], -300)] + t(4))
    grids.append([(8, 12), (6, 4)]); nreg.append(7)
    prompts.append(t(9) + [-300, -300] + t(2)); grids.append([]); nreg.append(2)
    prompts.append([151652, -200, 151653] + t(6)); grids.append([(4, 4)]); nreg.append(0)
    prompts.append(t(2) + [151652] + t(3) + [151652, -200] + t(5)); grids.append([(10, 6)]); nreg.append(0)
    out = E.splice_plan_batch(prompts, grids, nreg, "cuda")
    torch.cuda.synchronize()
    assert out["status"].cpu().tolist() == [0] * len(prompts)
    off = np.cumsum([0] + out["lens"])
    img_row = reg_row = 0
    for b, (p, g, n) in enumerate(zip(prompts, grids, nreg)):
        ref = E.splice_plan(p, g, n)
        sl = slice(int(off[b]), int(off[b + 1]))
        assert out["lens"][b] == len(ref["kind"])
        assert np.array_equal(out["new_ids"][sl].cpu().numpy().astype(np.int64), ref["new_ids"])
        assert np.array_equal(out["kind"][sl].cpu().numpy(), ref["kind"])
        idx = ref["index"].copy()
        idx[ref["kind"] == 1] += img_row; idx[ref["kind"] == 2] += reg_row          # the batch packs feature rows back to back
        assert np.array_equal(out["index"][sl].cpu().numpy(), idx)
        assert np.array_equal(out["position_ids"][:, sl].cpu().numpy(), ref["position_ids"])
        assert int(out["rope_delta"][b]) == ref["rope_delta"]
        img_row += sum(gh * gw // 4 for gh, gw in g); reg_row += n


def test_splice_plan_batch_reports_bad_samples_individually():
    E, SY = _eng()
    good = SY.synthetic_prompt(0, 3)
    bad = [5, -200, 6, -200, 7]                       # two image placeholders, one image
    out = E.splice_plan_batch([good, bad, good], [[(8, 8)], [(4, 4)], [(8, 8)]], [3, 0, 3], "cuda")
    st = out["status"].cpu().tolist()
    assert st[0] == 0 and st[2] == 0 and st[1] != 0


def test_parse_predictions_matches_text_regex():
    E, SY = _eng()
    from vlm_fo1.mm_utils import extract_predictions_to_indexes
    GS, GE, OS, OE, NL = 1000, 1001, 1002, 1003, 10
    REG = [2000 + i for i in range(120)]
    words = {i: f"w{i} " for i in range(20, 60)}
    text = {GS: "<ground>", GE: "</ground>", OS: "<objects>", OE: "</objects>", NL: "\n", **{r: f"<region{i}>" for i, r in enumerate(REG)}, **words}
    rng = random.Random(11)
    w = lambda k: [rng.randrange(20, 60) for _ in range(k)]
    seqs = [
        w(3) + [GS] + w(2) + [GE, OS, REG[3], REG[17], OE] + w(1) + [GS] + w(1) + [GE, OS, REG[0], OE],
        [GS] + w(1) + [GE, OS, OE] + [GS, 25, GE, OS, REG[5], REG[5], REG[99], OE],                       # empty list; duplicate region
        [GS] + w(1) + [GE] + w(1) + [OS, REG[1], OE] + [GS, 30, GE, OS, REG[2], OE],                       # </ground> not followed by <objects>
        [GS, 31, NL, GE, OS, REG[4], OE, GS, 32, GE, OS, REG[6], NL, OE, GS, 33, GE, OS, REG[7], OE],      # newline inside a group kills that match
        [GS, 34, GS, 35, GE, OS, REG[8], w(1)[0], REG[9], OE] + w(2),                                      # nested <ground> stays in the label
        w(5),
        [GS, 40, GE, OS, REG[10]],                                                                         # unterminated
        [GS, 41, GE, OS, REG[11], OE, GS, 41, GE, OS, REG[12], OE],                                        # same label twice: union
    ]
    T = max(len(s) for s in seqs)
    tok = torch.full((len(seqs), T), 0, dtype=torch.int32)
    for b, s in enumerate(seqs):
        tok[b, :len(s)] = torch.tensor(s, dtype=torch.int32)
    lens = torch.tensor([len(s) for s in seqs], dtype=torch.int32)
    recs = E.parse_predictions(tok.cuda(), lens.cuda(), (GS, GE), (OS, OE), REG, newline_token_ids=[NL], vocab=4096)
    for b, s in enumerate(seqs):
        ref = extract_predictions_to_indexes("".join(text[t] for t in s))
        got = {}
        for ls, le, n in recs[b]:
            label = "".join(text[t] for t in s[ls:le]).strip()
            got.setdefault(label, set())
            if n >= 0:
                got[label].add(n)
        assert got == ref, (b, got, ref)
