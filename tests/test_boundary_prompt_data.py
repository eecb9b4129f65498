"""The pure-data mirrors (special tokens, placeholder ids, task prompts) carry exactly the reference's names and values
(tests/golden/prompt_data.json is dumped from the reference's own modules by oracle/gen_golden.py stage ``prompt_data``)."""
import json
import os


def _public(mod, skip=()):
    return {k: v for k, v in vars(mod).items() if not k.startswith("_") and isinstance(v, (str, int)) and k not in skip}


def test_constants_and_templates_equal_reference(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "prompt_data.json")))
    import vlm_fo1.constants as C
    import vlm_fo1.task_templates as T
    assert _public(C) == gold["constants"]
    imported = set(gold["constants"])                      # names task_templates imports from constants
    assert _public(T, skip=imported) == gold["task_templates"]
    assert T.OD_template.format("cats") == "Please detect cats in this image. Answer the question with object indexes."
