"""The nine task prompts -- the model's trained text interface, so the VALUES are fixed (reference:
vlm_fo1/task_templates.py:1-17, including its spelling ``Viusal_Region_Reasoning_template`` and the doubled word in that
prompt).  They are assembled from their shared phrases and pinned against the reference's module by
tests/test_boundary_prompt_data.py."""
from .constants import DEFAULT_THINK_END, DEFAULT_THINK_START

_IN_IMAGE = "in this image"
_WITH_INDEXES = "with object indexes"
_ANSWER_INDEXED = f"Answer the question {_WITH_INDEXES}."

_detect = f"Please detect {{}} {_IN_IMAGE}. {_ANSWER_INDEXED}"
OD_template = _detect
REC_template = _detect
OD_Counting_template = (f"How many {{}} are there {_IN_IMAGE}? Count each instance of the target object. "
                        f"Locate them {_WITH_INDEXES} and then answer the question with the number of objects.")
Region_OCR_template = "Please provide the ocr results of {} in the image."
Brief_Region_Caption_template, Detailed_Region_Caption_template = (f"Provide a {kind} description for {{}}." for kind in ("brief", "detailed"))
Grounding_template = "Briefly describe this image and detect all mentioned objects. Answer with grounded object indexes."
Visual_Prompt_OD_template = (f"Using the provided object {{}} as a reference, identify all other objects of the same category "
                             f"{_IN_IMAGE}. Respond {_WITH_INDEXES}.")
_RP, _ANS = "reasoning process", "answer"
_tags = f"{DEFAULT_THINK_START} {DEFAULT_THINK_END} and <{_ANS}> </{_ANS}>"
_example = f"{DEFAULT_THINK_START} {_RP} here {DEFAULT_THINK_END}<{_ANS}> {_ANS} here </{_ANS}>"
Viusal_Region_Reasoning_template = " ".join([
    f"First thinks about the {_RP} in the mind and then provides the user with the {_ANS}.",
    f"The {_RP} and {_ANS} are enclosed within {_tags} tags, respectively, i.e., {_example}.",
    f"Please give a detailed {_RP} process and provide image regions that can help you {_ANS} the question better.",
    "{}"])
