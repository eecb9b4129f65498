"""Placeholder ids and special-token strings of the FO1 prompt format.  Pure data: every caller and the tokenisation
helpers rely on these names and values (reference: vlm_fo1/constants.py:1-29); they are generated here from two small
tables and pinned against the reference's module by tests/test_boundary_prompt_data.py."""

_INTS = {
    "IGNORE_INDEX": -100,                     # label value the loss skips
    "IMAGE_TOKEN_INDEX": -200,                # placeholder id of an <image> slot in input_ids
    "DEFAULT_REGION_INDEX": -300,             # placeholder id of a <regionfeat> slot
    "QWEN2_5_VL_IMAGE_TOKEN_INDEX": 151655,   # id of <|image_pad|> in the Qwen2.5 vocabulary
}
_SINGLE_TAGS = {                              # NAME -> text between the angle brackets
    "DEFAULT_IMAGE_TOKEN": "image",
    "DEFAULT_IMAGE_PATCH_TOKEN": "im_patch",
    "DEFAULT_IM_START_TOKEN": "im_start",
    "DEFAULT_IM_END_TOKEN": "im_end",
    "QWEN2_5_VL_IMAGE_TOKEN": "|image_pad|",
    "DEFAULT_REGION_TOKEN": "region<i>",
    "DEFAULT_REGION_FEATURE_TOKEN": "regionfeat",
}
_PAIRED_TAGS = {"GROUNDING": "ground", "GROUNDING_OBJECTS": "objects", "THINK": "think"}   # -> DEFAULT_<KEY>_START / _END

LOGDIR = "."
globals().update(_INTS)
globals().update({name: f"<{tag}>" for name, tag in _SINGLE_TAGS.items()})
for _key, _tag in _PAIRED_TAGS.items():
    globals()[f"DEFAULT_{_key}_START"] = f"<{_tag}>"
    globals()[f"DEFAULT_{_key}_END"] = f"</{_tag}>"
del _key, _tag

__all__ = ["LOGDIR", *_INTS, *_SINGLE_TAGS, *(f"DEFAULT_{k}_{e}" for k in _PAIRED_TAGS for e in ("START", "END"))]
