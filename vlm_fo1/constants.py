"""Placeholder ids and special-token strings of the FO1 prompt format (reference: vlm_fo1/constants.py:1-29).
Pure data every caller and the tokenisation helpers rely on; pinned against the reference's module by
tests/test_boundary_prompt_data.py."""

LOGDIR = "."
IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
DEFAULT_REGION_INDEX = -300
QWEN2_5_VL_IMAGE_TOKEN_INDEX = 151655
DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_IMAGE_PATCH_TOKEN = "<im_patch>"
DEFAULT_IM_START_TOKEN = "<im_start>"
DEFAULT_IM_END_TOKEN = "<im_end>"
QWEN2_5_VL_IMAGE_TOKEN = "<|image_pad|>"
DEFAULT_REGION_TOKEN = "<region<i>>"
DEFAULT_REGION_FEATURE_TOKEN = "<regionfeat>"
DEFAULT_GROUNDING_START = "<ground>"
DEFAULT_GROUNDING_END = "</ground>"
DEFAULT_GROUNDING_OBJECTS_START = "<objects>"
DEFAULT_GROUNDING_OBJECTS_END = "</objects>"
DEFAULT_THINK_START = "<think>"
DEFAULT_THINK_END = "</think>"
