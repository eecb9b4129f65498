"""Pin oracle/llm.py against the reference's own decoder-layer modules and get_rope_index (tests/golden/llm_small.npz),
and the engine's host-side splice / M-RoPE bookkeeping (C ABI, CPU) against the oracle -- bit-exact integer work."""
import ctypes as C

import numpy as np
import torch

from oracle import llm as OL
from tests.golden_io import load, nerr


def _cases(z):
    for ci in range(int(z["n_rope_cases"])):
        yield ci, z[f"rope_ids_{ci}"].tolist(), [tuple(int(v) for v in g) for g in z[f"rope_grids_{ci}"]], z[f"rope_pos_{ci}"], int(z[f"rope_delta_{ci}"])


def test_rope_index_bit_exact_vs_reference():
    _, z = load("llm_small")
    for ci, ids, grids, pos, delta in _cases(z):
        p, d = OL.rope_index(ids, grids)
        assert np.array_equal(p.numpy(), pos), ci
        assert d == delta, ci


def test_decoder_oracle_matches_reference_layers():
    sd, z = load("llm_small")
    cfg = z["cfg"]
    ids = z["rope_ids_0"].tolist()
    pos = torch.from_numpy(z["rope_pos_0"]); delta = int(z["rope_delta_0"])
    toks, step_logits, prompt_logits = OL.generate(sd, cfg, torch.from_numpy(z["llm_embeds"]), pos, delta, 3, stop_ids=[],
                                                   forced=z["llm_tokens"].tolist())
    assert nerr(prompt_logits, torch.from_numpy(z["llm_prompt_logits"])) < 2e-5
    assert nerr(step_logits, torch.from_numpy(z["llm_step_logits"])) < 2e-5
    free, _, _ = OL.generate(sd, cfg, torch.from_numpy(z["llm_embeds"]), pos, delta, 3, stop_ids=[])
    assert free == z["llm_tokens"].tolist()


class SpliceCfg(C.Structure):
    _fields_ = [("image_token_id", C.c_int32), ("video_token_id", C.c_int32), ("vision_start_token_id", C.c_int32),
                ("merge", C.c_int32), ("image_placeholder", C.c_int32), ("region_placeholder", C.c_int32)]


def _plan(L, ids, grids, n_regions, cap):
    cfg = SpliceCfg(151655, 151656, 151652, 2, -200, -300)
    a_ids = (C.c_int64 * len(ids))(*ids)
    a_g = (C.c_int32 * max(2 * len(grids), 1))(*[x for g in grids for x in g])
    new_ids = (C.c_int64 * cap)(); kind = (C.c_int32 * cap)(); idx = (C.c_int32 * cap)(); pos = (C.c_int32 * (3 * cap))()
    delta = C.c_int32(); n = C.c_int32()
    L.fo1_splice_plan.restype = C.c_int
    rc = L.fo1_splice_plan(a_ids, len(ids), a_g, len(grids), n_regions, C.byref(cfg), new_ids, kind, idx, pos, C.byref(delta), C.byref(n), cap)
    return rc, n.value, delta.value, np.frombuffer(new_ids, dtype=np.int64), np.frombuffer(kind, dtype=np.int32), np.frombuffer(idx, dtype=np.int32), np.frombuffer(pos, dtype=np.int32).reshape(3, cap)


def test_splice_plan_cabi_bit_exact():
    import fo1_b200
    L = fo1_b200.lib()
    _, z = load("llm_small")
    for ci, ids, grids, pos, delta in _cases(z):
        # rebuild the placeholder form of the prompt: the run of image tokens collapses to one -200
        raw, i = [], 0
        while i < len(ids):
            if ids[i] == 151655:
                raw.append(-200)
                while i < len(ids) and ids[i] == 151655:
                    i += 1
            else:
                raw.append(ids[i]); i += 1
        n_reg = sum(1 for t in raw if t == -300)
        o_ids, o_kind, o_idx = OL.splice_plan(raw, grids, n_reg)
        assert o_ids == ids
        rc, n, d, new_ids, kind, idx, p = _plan(L, raw, grids, n_reg, len(ids) + 7)
        assert rc == 0, L.fo1_last_error()
        assert n == len(ids) and d == delta
        assert np.array_equal(new_ids[:n], np.array(ids)), ci
        assert np.array_equal(kind[:n], np.array(o_kind)) and np.array_equal(idx[:n], np.array(o_idx)), ci
        assert np.array_equal(p[:, :n], pos), ci
    # capacity too small -> FO1_ERR_WORKSPACE with the needed length reported; too few region features -> invalid arg
    rc, n, *_ = _plan(L, [1, -200, 2], [(4, 4)], 0, 3)
    assert rc == -4 and n == 6
    rc, *_ = _plan(L, [1, -300, 2], [], 0, 8)
    assert rc == -1
