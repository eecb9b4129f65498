"""CPU: the boundary mirror ``vlm_fo1.mm_utils`` (prompt assembly, placeholder tokenisation, box clamp/rescale/cap,
stop ids, output parsing) is bit-exact against the REFERENCE's own mm_utils.prepare_inputs / extract_predictions_*
(tests/golden/mm_utils.npz, produced by oracle/gen_golden.py with the same fabricated tokenizer and processors)."""
import json
import os
import tempfile
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch
from PIL import Image


@pytest.fixture(scope="module")
def env():
    from importlib import import_module
    import fo1_b200  # noqa: F401
    FB = import_module("vlm-fo1_b200.fabricate")
    d = os.path.join(tempfile.mkdtemp(), "VLM-FO1_Qwen2.5-VL-3B-v01")
    os.makedirs(d)
    FB.write_tokenizer(d)
    from transformers import AutoTokenizer
    tok = AutoTokenizer.from_pretrained(d, use_fast=False)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "mm_utils.npz"))
    Image.fromarray(z["img_a"]).save(os.path.join(d, "a.png"))
    rng = np.random.default_rng(7)
    rng.integers(0, 256, (399, 500, 3), dtype=np.uint8)                 # same draw order as the generator
    Image.fromarray(rng.integers(0, 256, (2300, 1200, 3), dtype=np.uint8)).save(os.path.join(d, "b.png"))
    return d, tok, z


def test_prepare_inputs_bit_exact(env):
    d, tok, z = env
    import vlm_fo1.mm_utils as MU
    from vlm_fo1.processors import AuxImageProcessor, PrimaryImageProcessor
    assert "reference" not in MU.__file__
    procs = (PrimaryImageProcessor(), AuxImageProcessor(768, "dynamic"))
    model = NS(config=NS(mm_use_region_index_token=True))
    for case in json.loads(bytes(z["cases_json"]).decode()):
        tag = case["tag"]
        msgs = []
        if case["system"]:
            msgs.append({"role": "system", "content": case["system"]})
        m = {"role": "user", "content": [{"type": "image_url", "image_url": {"url": os.path.join(d, case["img"])}},
                                         {"type": "text", "text": case["text"]}]}
        if case["boxes"] is not None:
            m["bbox_list"] = [list(b) for b in case["boxes"]]
        msgs.append(m)
        kw = MU.prepare_inputs(d, model, procs, tok, msgs, device="cpu", max_tokens=64)
        assert np.array_equal(kw["inputs"].numpy(), z[f"{tag}_inputs"]), tag                     # token ids incl. -200 / -300
        assert np.array_equal(kw["image_grid_thws"][0].numpy(), z[f"{tag}_grid"]), tag
        assert list(kw["images_aux"][0].shape) == z[f"{tag}_aux_shape"].tolist(), tag
        assert np.array_equal(kw["bbox_list"][0].numpy(), z[f"{tag}_boxes"]), tag                 # clamp + rescale + cap 100, fp32 bit-exact
        assert float(kw["images"][0].double().sum()) == float(z[f"{tag}_px_sum"]), tag
        assert float(kw["images_aux"][0].double().sum()) == float(z[f"{tag}_aux_sum"]), tag
        stop = [int(k.item()) for k in kw["stopping_criteria"][0].keyword_ids if k.numel() == 1]
        assert stop == z[f"{tag}_stop"].tolist() == [151645]
        assert kw["max_new_tokens"] == 64 and kw["do_sample"] is False and kw["use_cache"] is True
        assert set(kw) == {"inputs", "images", "images_aux", "image_grid_thws", "bbox_list", "do_sample", "temperature", "max_new_tokens",
                           "streamer", "top_p", "use_cache", "stopping_criteria", "pad_token_id"}


def test_prediction_parsing_matches_reference(env):
    _, _, z = env
    import vlm_fo1.mm_utils as MU
    ref = json.loads(bytes(z["parse_json"]).decode())
    boxes7 = [[161.0, 11.0, 292.0, 127.0], [268.0, 61.0, 428.0, 226.0], [12.0, 100.0, 140.0, 227.0], [205.0, 188.0, 332.0, 320.0],
              [326.0, 202.0, 478.0, 357.0], [136.0, 106.0, 269.0, 233.0], [25.0, 206.0, 200.0, 383.0]]
    idx = MU.extract_predictions_to_indexes(ref["pred"])
    assert {k: sorted(v) for k, v in idx.items()} == ref["indexes"]
    bxs = MU.extract_predictions_to_bboxes(ref["pred"], boxes7)
    assert {k: sorted(v) for k, v in bxs.items()} == ref["bboxes"]


def test_templates_and_upn_stub_importable():
    from vlm_fo1.task_templates import OD_template
    from vlm_fo1 import constants as K
    assert OD_template.format("orange").startswith("Please detect orange")
    assert (K.IMAGE_TOKEN_INDEX, K.DEFAULT_REGION_INDEX) == (-200, -300)
    from detect_tools.upn import UPNWrapper                      # inference.py:3 imports it without using it
    with pytest.raises(NotImplementedError):
        UPNWrapper("x")
