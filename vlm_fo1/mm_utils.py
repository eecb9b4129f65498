"""Host-side prompt / image / box preparation and output parsing -- the reference's ``vlm_fo1.mm_utils`` surface
(file:line cited per function), re-implemented; all of it runs on the host and is integer / string work that must be
(and is tested to be) bit-exact with the reference."""
from __future__ import annotations

import base64
import io
import random
import re
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from PIL import Image, ImageDraw

from vlm_fo1 import constants as K
from vlm_fo1.constants import (DEFAULT_IM_END_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_REGION_FEATURE_TOKEN, DEFAULT_REGION_INDEX,
                               DEFAULT_REGION_TOKEN, IMAGE_TOKEN_INDEX)

_GROUND_RE = re.compile(r"<ground>(.*?)<\/ground><objects>(.*?)<\/objects>")
_REGION_RE = re.compile(r"<region(\d+)>")


def _as_tensor(ids: List[int], return_tensors: Optional[str]):
    if return_tensors is None:
        return ids
    if return_tensors != "pt":
        raise ValueError(f"Unsupported tensor type: {return_tensors}")
    return torch.tensor(ids, dtype=torch.long)


def tokenizer_image_token(prompt, tokenizer, image_token_index=IMAGE_TOKEN_INDEX, return_tensors=None):
    """mm_utils.py:28-81 -- text pieces tokenised separately, ``image_token_index`` between them (``-200`` for the
    indexed ``<image_N>`` form); a leading BOS is kept once."""
    if "<image_0>" in prompt:
        pieces = re.split(r"<image_[0-9]+>", prompt)
        n_tags = len(re.findall(r"<image_(\d+)>", prompt))
        ids: List[int] = []
        for i, piece in enumerate(pieces):
            ids += tokenizer(piece).input_ids
            if i < n_tags:
                ids.append(-200)
        return _as_tensor(ids, return_tensors)
    chunks = [tokenizer(piece).input_ids for piece in prompt.split("<image>")]
    ids = []
    skip = 0
    if chunks and chunks[0] and chunks[0][0] == tokenizer.bos_token_id:
        skip = 1
        ids.append(chunks[0][0])
    for i, chunk in enumerate(chunks):
        ids += chunk[skip:]
        if i + 1 < len(chunks):
            ids += ([image_token_index] * (skip + 1))[skip:]
    return _as_tensor(ids, return_tensors)


def tokenizer_image_region_token(prompt, tokenizer, image_token_index=IMAGE_TOKEN_INDEX, region_token_index=DEFAULT_REGION_INDEX,
                                 return_tensors=None):
    """mm_utils.py:83-135 -- split on ``<image>`` then ``<regionfeat>``; one ``region_token_index`` per region slot,
    one ``image_token_index`` between image chunks."""
    groups = [[tokenizer(sub).input_ids for sub in chunk.split("<regionfeat>")] for chunk in prompt.split("<image>")]
    ids: List[int] = []
    skip = 0
    if groups and groups[0] and groups[0][0] and groups[0][0][0] == tokenizer.bos_token_id:
        skip = 1
        ids.append(groups[0][0][0])
    for gi, group in enumerate(groups):
        if group:
            ids += group[0][skip:]
        for sub in group[1:]:
            ids.append(region_token_index)
            ids += sub
        if gi + 1 < len(groups):
            ids.append(image_token_index)
    return _as_tensor(ids, return_tensors)


try:  # the callers pass this object through ``stopping_criteria=[...]``; keep HF's base class when available
    from transformers import StoppingCriteria as _StopBase
except Exception:  # pragma: no cover
    _StopBase = object


class KeywordsStoppingCriteria(_StopBase):
    """mm_utils.py:137-181.  The engine does not call this per token (that would be a host sync per step); it reads
    ``keyword_ids`` once and checks single-token keywords on the device (``fo1_llm_generate`` stop ids).  The callable
    form is kept for callers that use it directly."""

    def __init__(self, keywords, tokenizer, input_ids):
        self.keywords = keywords
        self.tokenizer = tokenizer
        self.keyword_ids = []
        self.max_keyword_len = 0
        for kw in keywords:
            ids = tokenizer(kw).input_ids
            if len(ids) > 1 and ids[0] == tokenizer.bos_token_id:
                ids = ids[1:]
            self.max_keyword_len = max(self.max_keyword_len, len(ids))
            self.keyword_ids.append(torch.tensor(ids))
        self.start_len = input_ids.shape[1]

    def call_for_batch(self, output_ids: torch.LongTensor, scores, **kwargs) -> bool:
        self.keyword_ids = [k.to(output_ids.device) for k in self.keyword_ids]
        for k in self.keyword_ids:
            if torch.equal(output_ids[0, -k.shape[0]:], k):
                return True
        tail = min(output_ids.shape[1] - self.start_len, self.max_keyword_len)
        text = self.tokenizer.batch_decode(output_ids[:, -tail:], skip_special_tokens=True)[0]
        return any(kw in text for kw in self.keywords)

    def __call__(self, output_ids: torch.LongTensor, scores, **kwargs) -> bool:
        return all(self.call_for_batch(output_ids[i].unsqueeze(0), scores) for i in range(output_ids.shape[0]))


def load_image(image_file):
    """mm_utils.py:183-213 -- path / URL / data-URI -> RGB PIL image of at least 28x28.  (The reference crashes on a PIL
    input because it calls ``.startswith`` on it; here a PIL image is accepted.)"""
    if isinstance(image_file, Image.Image):
        image = image_file
    elif image_file.startswith("http"):
        import requests
        image = Image.open(io.BytesIO(requests.get(image_file).content))
    elif image_file.startswith("data:image/"):
        image = Image.open(io.BytesIO(base64.b64decode(image_file.replace("data:image/jpeg;base64,", ""))))
    else:
        image = Image.open(image_file).convert("RGB")
    if image.width < 28 or image.height < 28:
        image = image.resize((max(28, image.width), max(28, image.height)))
    return image


def image_to_base64(img_pil):
    """mm_utils.py:215-228"""
    with io.BytesIO() as buf:
        img_pil.save(buf, format="JPEG")
        return base64.b64encode(buf.getvalue()).decode("utf-8")


def draw_bboxes_and_save(image: Image.Image, fo1_bboxes: dict = {}, detection_bboxes: List[Tuple[int, int, int, int]] = [],
                         output_path: str = "output.jpg", color: str = "red", total_color: str = "green", width: int = 2) -> None:
    """mm_utils.py:230-279 -- detection boxes in ``total_color``, labelled FO1 boxes in ``color``; saves to ``output_path``."""
    draw = ImageDraw.Draw(image)
    for box in detection_bboxes:
        if len(box) != 4:
            print(f"warning: skipping malformed box {box}")
            continue
        draw.rectangle([(box[0], box[1]), (box[2], box[3])], outline=total_color, width=width)
    for label, boxes in fo1_bboxes.items():
        for box in boxes:
            if len(box) != 4:
                print(f"warning: skipping malformed box {box}")
                continue
            draw.rectangle([(box[0], box[1]), (box[2], box[3])], outline=color, width=width)
            draw.text((box[0], box[1]), label, fill=color)
    try:
        image.save(output_path)
        print(f"image saved to: {output_path}")
    except IOError as exc:
        print(f"error: could not save image to {output_path}: {exc}")


def adjust_bbox(bbox_list, original_h, original_w, resize_h, resize_w):
    """mm_utils.py:281-312 -- clamp to the original image, then rescale to the processed size (same float op order:
    ``v * resize / original``)."""
    out = []
    for x1, y1, x2, y2 in bbox_list:
        x1 = max(0, min(original_w, x1)); y1 = max(0, min(original_h, y1))
        x2 = max(0, min(original_w, x2)); y2 = max(0, min(original_h, y2))
        out.append([x1 * resize_w / original_w, y1 * resize_h / original_h, x2 * resize_w / original_w, y2 * resize_h / original_h])
    return out


def extract_predictions_to_indexes(prediction: str) -> Dict[str, set]:
    """mm_utils.py:346-369 -- ``<ground>label</ground><objects><regionN>...</objects>`` -> {label: {N, ...}}."""
    found: Dict[str, set] = {}
    for label, body in _GROUND_RE.findall(prediction):
        idx = {int(n) for n in _REGION_RE.findall(body)}
        label = label.strip()
        found[label] = found[label] | idx if label in found else idx
    return found


def extract_predictions_to_bboxes(prediction: str, bbox_list):
    """mm_utils.py:314-344"""
    return {label: [bbox_list[i] for i in idx] for label, idx in extract_predictions_to_indexes(prediction).items()}


def resize_shortest_edge_images_and_bboxes(image_list: List[Image.Image], bbox_lists: List, candidate_sizes: List[int] = [],
                                           max_size: int = 2048):
    """mm_utils.py:371-462 -- optional random short-edge target, long edge capped at ``max_size``, min 28 px, BICUBIC;
    boxes scaled by the realised width/height ratios.  A single [N,4] list comes back as a single list."""
    single = len(torch.tensor(bbox_lists).shape) == 2 and torch.tensor(bbox_lists).shape[1] == 4
    if single:
        bbox_lists = [bbox_lists]
    if not image_list or not bbox_lists:
        raise ValueError("Input lists cannot be empty.")
    if len(image_list) != len(bbox_lists):
        raise ValueError("The lengths of the image list and the bounding box list must be the same.")
    target = random.choice(candidate_sizes) if len(candidate_sizes) > 0 else None
    out_imgs, out_boxes = [], []
    for img, boxes in zip(image_list, bbox_lists):
        ow, oh = img.size
        scale = target / min(ow, oh) if target else 1.0
        nh, nw = int(oh * scale), int(ow * scale)
        if max(nh, nw) > max_size:
            scale = max_size / max(nh, nw)
            nh, nw = int(nh * scale), int(nw * scale)
        nw, nh = max(28, nw), max(28, nh)
        out_imgs.append(img if (nw == ow and nh == oh) else img.resize((nw, nh), Image.Resampling.BICUBIC))
        sx, sy = nw / ow, nh / oh
        out_boxes.append([[x1 * sx, y1 * sy, x2 * sx, y2 * sy] for x1, y1, x2, y2 in boxes])
    return (out_imgs, out_boxes[0]) if single else (out_imgs, out_boxes)


def make_message_context(tokenizer, message, chat_format="chatml"):
    """mm_utils.py:464-528 -- one chat message -> (prompt text, token ids with -200/-300 placeholders, image urls, boxes)."""
    image_urls = []
    if chat_format != "chatml":
        return None
    im_start, im_end = "<|im_start|>", "<|im_end|>"
    role, content = message["role"], message["content"]
    bbox_list = message.get("bbox_list", None)
    nl = tokenizer.encode("\n")

    def plain(text):
        ids = tokenizer.encode(role, allowed_special=set()) + nl + tokenizer.encode(text, allowed_special=set())
        return f"{im_start}{role}\n{text}{im_end}\n", [151644] + ids + [151645]

    inp, tokens = None, None
    if role == "system" or (role == "user" and isinstance(content, str)):
        inp, tokens = plain(content)
    if role == "user" and isinstance(content, list):
        inp = f"{im_start}{role}\n"
        for part in content:
            if part["type"] == "text":
                inp += f"{part['text']}"
            if part["type"] == "image_url":
                inp += DEFAULT_IM_START_TOKEN + "<image>" + DEFAULT_IM_END_TOKEN + "\n"
                if bbox_list and len(bbox_list) > 0:
                    inp += "".join(DEFAULT_REGION_TOKEN.replace("<i>", str(i)) + DEFAULT_REGION_FEATURE_TOKEN for i in range(len(bbox_list)))
                    inp += "\n"
                image_urls.append(part["image_url"]["url"])
        inp += f"{im_end}\n"
        if bbox_list and len(bbox_list) > 0:
            tokens = tokenizer_image_region_token(inp, tokenizer)
        else:
            tokens = tokenizer_image_token(inp, tokenizer, image_token_index=IMAGE_TOKEN_INDEX)
    return inp, tokens, image_urls, bbox_list


def prepare_inputs(model_name, model, image_processors, tokenizer, messages, device="cuda", max_tokens=512, top_p=1.0, temperature=0.0,
                   do_sample=False):
    """mm_utils.py:530-655 -- messages -> kwargs for ``model.generate``: ``inputs`` (int64 [1, P] with -200 / -300
    placeholders), ``images`` / ``image_grid_thws`` (primary processor), ``images_aux`` (aux processor), ``bbox_list``
    (boxes clamped to the image, capped at 100, rescaled to the aux tensor), stopping criteria, streamer, sampling flags."""
    global DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN
    if "qwen2.5-vl" in model_name.lower() or "qwen2_5_vl" in model_name.lower():
        DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN = "<|vision_start|>", "<|vision_end|>"
    primary_proc, aux_proc = image_processors

    prompt, input_tokens, image_urls, bbox_list = "", [], [], None
    for message in messages:
        inp, ctx, image_urls, bbox_list = make_message_context(tokenizer, message)
        prompt += inp
        input_tokens.extend(ctx)
    if "system" not in prompt:
        sys_text = "system\nYou are a helpful assistant."
        prompt = "<|im_start|>" + sys_text + "<|im_end|>" + "\n" + prompt
        input_tokens = [151644] + tokenizer(sys_text).input_ids + [151645] + tokenizer("\n").input_ids + input_tokens
    if not prompt.endswith("<|im_start|>assistant"):
        prompt += "<|im_start|>" + "assistant" + "\n"
        input_tokens.extend([151644] + tokenizer("assistant\n").input_ids)

    aux_tensors = None
    if image_urls:
        images = [load_image(u) for u in image_urls]
        images, bbox_list = resize_shortest_edge_images_and_bboxes(images, bbox_list, max_size=2048)
        if getattr(model.config, "mm_use_region_index_token", False):
            sizes = [im.size for im in images]
            aux_tensors = [aux_proc.preprocess(im, return_tensors="pt")["pixel_values"][0].to(device) for im in images.copy()]
            if bbox_list and len(bbox_list) > 0:
                bbox_list = bbox_list[:100]                                       # the reference's silent cap (:600)
                rh, rw = aux_tensors[0].shape[-2:]
                ow, oh = sizes[0]
                bbox_list = [torch.tensor(adjust_bbox(bbox_list, oh, ow, rh, rw))]
            else:
                bbox_list = None
    primary, grids = [], []
    for im in images:
        data = primary_proc.preprocess(im, videos=None, return_tensors="pt")
        primary.append(data["pixel_values"].to(device))
        grids.append(data["image_grid_thw"])

    if "qwen" in model_name.lower():
        input_ids = torch.tensor([input_tokens]).to(device)
        keywords = ["<|im_end|>"]
    stopping = KeywordsStoppingCriteria(keywords, tokenizer, input_ids)
    try:
        from transformers import TextStreamer
        streamer = TextStreamer(tokenizer, skip_prompt=True, skip_special_tokens=True)
    except Exception:  # pragma: no cover
        streamer = None
    print("question:================\n", prompt, "\n=================")
    return dict(inputs=input_ids, images=primary, images_aux=aux_tensors, image_grid_thws=grids, bbox_list=bbox_list,
                do_sample=(temperature != 0.0), temperature=temperature, max_new_tokens=max_tokens, streamer=streamer, top_p=top_p,
                use_cache=True, stopping_criteria=[stopping], pad_token_id=tokenizer.pad_token_id)
