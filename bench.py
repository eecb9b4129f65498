#!/usr/bin/env python
"""bench.py -- contract benchmark of the FO1 hot path on N B200s (one rank per GPU).

Workload (BASELINE.json configs[2], the configuration the headline metric is quoted on): per GPU a batch of 32
synthetic 896x896 images, 64 boxes each, random-init 3B FO1 weights: dual vision towers (ViT + DaViT) -> SimpleFPN ->
HFRE region tokens -> projector -> splice -> Qwen2.5-3B prefill -> 64-token greedy decode; weak scaling (per-GPU batch
fixed), one NCCL all-gather of the decoded ids per step.  A "step" is one pass of that pipeline over one batch.

  value : images/sec, inputs already resident in HBM, CUDA-event timed, max over ranks.
  e2e   : the same through the public pipeline API from pinned HOST buffers (H2D of pixel rows / aux images /
          boxes and D2H of the token ids inside the timed region).
  roofline     : the dominant kernel (tcgen05 GEMM) -- algorithmic FLOPs / CUDA-event launch time (one extra
                 instrumented step after the timed region) vs the measured cuBLAS bf16 sustained peak.
  roofline_hfre: HFRE gather -- SURVEY 8d algorithmic bytes / launch time vs the measured HBM copy bandwidth.
  cpu_baseline : the CPU oracle port of the same path (oracle/pipeline.py) timed on the host cores on ONE image.
``--impl reference`` times only that CPU arm (the reference's PyTorch path restated; the original modules cannot
travel to the GPU box)."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "images/sec (prefill+64-tok decode) 3B FO1"
UNIT = "images/s"


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.proc = None
        self.lines = []
        self.index = index

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


CPU_SAMPLE_NOTE = ("1 image at full resolution through oracle/pipeline.py (fp32, all host threads) on a REDUCED-DEPTH copy of the model with "
                   "the real layer widths: ViT 1 windowed + 1 full-attention block, DaViT 1 block per stage, LLM 2 layers, {tok} decode steps; "
                   "per-block / per-layer wall times are measured and extrapolated to the real depths (ViT 28+4 blocks, DaViT 1/1/9/1, LLM 36 "
                   "layers, {T}-token decode); patch-embed, merger, conv-embeds, FPN, HFRE, projector and lm_head are measured at full size")


def reduced_cfg(E, cfg):
    """Same widths, minimal depth: what the CPU arm actually executes."""
    r = E.EngineConfig()
    r.vit = dict(cfg.vit, depth=2, fullatt_block_indexes=[1])
    r.davit = dict(cfg.davit, depths=[1, 1, 1, 1])
    r.llm = dict(cfg.llm, num_hidden_layers=2)
    r.fpn_out, r.region_dim, r.proj_aux_layers = cfg.fpn_out, cfg.region_dim, cfg.proj_aux_layers
    return r


def cpu_arm(args, CK, E, cfg, steps):
    """Time the CPU oracle port.  Bounded sample (CPU_SAMPLE_NOTE): returns (images/s, stage seconds, per-step seconds)."""
    from importlib import import_module
    from oracle import pipeline as OP
    SY = import_module("vlm-fo1_b200.synthetic")
    # thread count: all host cores unless a quick matmul calibration shows fewer threads are faster (oversubscribed or
    # shared hosts make 128-thread GEMVs pathologically slow); the count used is reported as `cores`
    best_t, best_s = os.cpu_count() or 1, None
    a = torch.randn(1195, 2048); w = torch.randn(11008, 2048); v = torch.randn(1, 2048)
    for nt in sorted({os.cpu_count() or 1, 64, 32, 16}, reverse=True):
        if nt > (os.cpu_count() or 1):
            continue
        torch.set_num_threads(nt)
        t0 = time.perf_counter()
        for _ in range(2):
            (a @ w.t()); [(v @ w.t()) for _ in range(8)]
        dt = time.perf_counter() - t0
        if best_s is None or dt < best_s:
            best_t, best_s = nt, dt
    torch.set_num_threads(best_t)
    rc = reduced_cfg(E, cfg)
    sds = CK.random_state_dicts(rc, "cpu", 0)
    sds = {k: {n: t.float() for n, t in v.items()} for k, v in sds.items()}
    s = SY.synthetic_batch(0, 1, args.size, args.boxes)[0]
    times, detail = [], None
    for _ in range(steps):
        with torch.no_grad():
            out = OP.run_sample(sds, rc.vit, rc.davit, rc.llm, input_ids=s.input_ids, pixel_values=s.pixel_values, grid_hw=s.grid_hw,
                                image_aux=s.image_aux, boxes=s.boxes, region_dim=cfg.region_dim, max_new_tokens=args.cpu_tokens)
        t = out["timings"]
        vd, dd = t["vit_detail"], t["davit_detail"]
        t_win = statistics.mean(x for f, x in vd["blocks"] if not f)
        t_full = statistics.mean(x for f, x in vd["blocks"] if f)
        n_full = len(cfg.vit["fullatt_block_indexes"])
        vit = vd["embed_s"] + vd["merger_s"] + (cfg.vit["depth"] - n_full) * t_win + n_full * t_full
        davit = sum(dd["embed_s"]) + sum(cfg.davit["depths"][i] * statistics.mean(dd["block_s"][i]) for i in range(4))
        L = cfg.llm["num_hidden_layers"]
        prefill = L * statistics.mean(t["llm_prefill_layer_s"]) + t["llm_head_s"]
        per_tok = L * statistics.mean(t["llm_decode_layer_s"]) + t["llm_head_s"]
        total = vit + davit + t["fpn_s"] + t["hfre_s"] + t["proj_s"] + prefill + (args.tokens - 1) * per_tok
        times.append(total)
        detail = {"vit_s": vit, "davit_s": davit, "fpn_s": t["fpn_s"], "hfre_s": t["hfre_s"], "proj_s": t["proj_s"], "llm_prefill_s": prefill,
                  "llm_decode_s_per_token": per_tok, "vit_block_windowed_s": t_win, "vit_block_full_s": t_full}
    return 1.0 / statistics.mean(times), detail, times


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fo1", choices=["fo1", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--size", type=int, default=896)
    ap.add_argument("--boxes", type=int, default=64)
    ap.add_argument("--tokens", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-tokens", type=int, default=3, help="decode steps the CPU arm actually runs (rest extrapolated)")
    ap.add_argument("--profile-run", action="store_true", help="warm-up exactly as given, one timed pass, nothing else (for ncu)")
    ap.add_argument("--small", action="store_true", help="tiny architecture (plumbing check only; NOT a valid bench number)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from importlib import import_module
    import fo1_b200  # noqa: F401
    E = import_module("vlm-fo1_b200.engine"); CK = import_module("vlm-fo1_b200.checkpoint")
    P = import_module("vlm-fo1_b200.pipeline"); SY = import_module("vlm-fo1_b200.synthetic"); HF = import_module("vlm-fo1_b200.hfre")
    cfg = E.EngineConfig()
    if args.small:
        cfg.vit = dict(cfg.vit, depth=4, fullatt_block_indexes=[1, 3])
        cfg.davit = dict(cfg.davit, depths=[1, 1, 1, 1])
        cfg.llm = dict(cfg.llm, num_hidden_layers=2)
    config = {"workload": f"full prefill + {args.tokens}-token decode: batch {args.batch}/GPU, {args.size}x{args.size}, {args.boxes} boxes/img, "
                          f"random-init 3B FO1 (variant B: SimpleFPN, D=5888, mlp2x_gelu projector)",
              "global_batch": args.batch * world, "per_gpu_batch": args.batch, "image_size": args.size, "boxes_per_image": args.boxes,
              "decode_tokens": args.tokens, "parallelism": f"dp{world}", "l2_policy": "working set (8.3 GB weights + GBs of activations per step) >> 126 MB L2; no flush needed",
              "small_arch": bool(args.small)}

    # ------------------------------------------------------------------ reference (CPU) arm
    if args.impl == "reference":
        if rank != 0:
            return
        args.steps = max(1, min(args.steps, 3)); args.warmup = 0   # bounded: each step is tens of seconds of CPU work
        ips, detail, times = cpu_arm(args, CK, E, cfg, args.steps)
        cores = torch.get_num_threads()
        line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": 0,
                "ms_per_step": 1000.0 / ips, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": CPU_SAMPLE_NOTE.format(tok=args.cpu_tokens, T=args.tokens),
                                 "stage_seconds": detail},
                "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=out, flush=True)
        return

    # ------------------------------------------------------------------ fo1 arm
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    sds = CK.random_state_dicts(cfg, dev, 0)
    eng = CK.load_engine(cfg, sds, dev)
    keep_cpu = (rank == 0 and world == 1 and not args.no_cpu_baseline)
    del sds
    torch.cuda.empty_cache()
    pipe = P.Fo1Pipeline(eng)
    B, T = args.batch, args.tokens
    host = SY.synthetic_batch(rank * B, B, args.size, args.boxes)
    for s in host:
        s.pixel_values = s.pixel_values.pin_memory(); s.image_aux = s.image_aux.pin_memory(); s.boxes = s.boxes.pin_memory()
    resident = [P.SampleInputs(s.input_ids, s.pixel_values.to(dev), s.grid_hw, s.image_aux.to(dev), s.boxes.to(dev)) for s in host]
    gathered = torch.empty((world * B, T), dtype=torch.int32, device=dev) if world > 1 else None
    host_tokens = torch.empty((B, T), dtype=torch.int32).pin_memory()

    def step(samples):
        out = pipe.generate(samples, T, stop_ids=[], early_exit_interval=0)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out["tokens"])   # the path's only collective: decoded ids over NVLink
        return out

    def step_e2e():
        moved = [P.SampleInputs(s.input_ids, s.pixel_values.to(dev, non_blocking=True), s.grid_hw, s.image_aux.to(dev, non_blocking=True),
                                s.boxes.to(dev, non_blocking=True)) for s in host]
        out = step(moved)
        host_tokens.copy_(out["tokens"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out

    def timed(fn, steps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    L = fo1_b200.lib()
    if args.profile_run:
        for _ in range(args.warmup):
            step(resident)
        ms = timed(lambda: step(resident), args.steps)
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms / args.steps, "launches": int(L.fo1_launch_count())}), file=out, flush=True)
        return
    for _ in range(max(args.warmup, 3)):
        step(resident)
    torch.cuda.synchronize()
    L.fo1_launch_count_reset()
    with ClockSampler(local) as cs:
        ms = timed(lambda: step(resident), args.steps)
    launches = int(L.fo1_launch_count())
    clocks = cs.summary()
    value = world * B * args.steps / (ms / 1000.0)
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e = world * B * args.steps / (ms_e2e / 1000.0)
    h2d = sum(s.pixel_values.numel() * 4 + s.image_aux.numel() * 4 + s.boxes.numel() * 4 for s in host)
    d2h = B * T * 4

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches}

    if rank == 0:
        # ---- one extra instrumented step: per-kernel CUDA-event times for the roofline objects ----
        peaks = load_peaks()
        # (rank-0-only passes must not contain the collective: the other ranks are already at the final barrier)
        local_step = lambda: pipe.generate(resident, T, stop_ids=[], early_exit_interval=0)
        pipe.profile_stages = True
        local_step()
        line["stage_ms"] = pipe.stage_ms()          # un-instrumented kernels, CUDA events between the stages
        pipe.profile_stages = False
        L.fo1_profile_enable(1)
        local_step()
        buf = (__import__("ctypes").c_char * (1 << 20))()
        L.fo1_profile_collect(buf, 1 << 20)
        L.fo1_profile_enable(0)
        raw = json.loads(buf.value.decode())
        # fold the per-shape GEMM records into two totals, keep the 12 heaviest shapes for the report
        prof, shapes = {}, []
        for k, v in raw.items():
            base = k.split(":")[0]
            if base != k:
                shapes.append({"shape": k, **v, "tflops": (v["flops"] / v["ms"] / 1e9) if v["ms"] > 0 else 0.0,
                               "GBs": (v["bytes"] / v["ms"] / 1e6) if v["ms"] > 0 else 0.0})
            a = prof.setdefault(base, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0, "max_ms": 0.0})
            a["launches"] += v["launches"]; a["ms"] += v["ms"]; a["flops"] += v["flops"]; a["bytes"] += v["bytes"]
            a["max_ms"] = max(a["max_ms"], v["max_ms"])
        line["kernel_profile"] = prof
        line["gemm_shapes_top"] = sorted(shapes, key=lambda r: -r["ms"])[:14]
        g = prof.get("gemm")
        if g and g["ms"] > 0:
            ach = g["flops"] / g["ms"] / 1e9
            line["roofline"] = {"kernel": "gemm_bf16_tcgen05_kernel (M > 128 launches of one step)", "bound": "tensor", "achieved": ach,
                                "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": ach / peaks["tflops"], "traffic": None,
                                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
                                "launches": g["launches"], "ms_per_launch": g["ms"] / g["launches"], "share_of_step": g["ms"] / (ms / args.steps)}
        h = prof.get("hfre_sweep_mma") or prof.get("hfre_sweep") or prof.get("hfre_gather")
        if h and h["ms"] > 0:
            tot = 0
            for s in host:
                H0 = args.size // 4
                gh, gw = s.grid_hw
                shapes = [(args.size // (4 << i), args.size // (4 << i), c) for i, c in enumerate(cfg.davit["dim_embed"])] + \
                         [(int(gh * f), int(gw * f), cfg.fpn_out) for f in (4, 2, 1, 0.5)]
                sc = gh * 14 / args.size
                bl = [s.boxes.numpy()] * 4 + [s.boxes.numpy() * sc] * 4
                scales = [0.25] * 4 + [1 / x for x in HF.FPN_STRIDES]
                ups = [H0 // sh[0] for sh in shapes[:4]] + [1] * 4
                tot += HF.algorithmic_bytes(shapes, bl, scales, ups, s.boxes.shape[0], cfg.region_dim)["unique_bytes"]
            ach = tot / h["ms"] / 1e6
            line["roofline_hfre"] = {"kernel": "hfre_sweep_mma_kernel" if "hfre_sweep_mma" in prof else ("hfre_sweep_kernel" if "hfre_sweep" in prof else "hfre_gather_kernel"), "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                     "frac": ach / peaks["hbm_gbs"], "traffic": None, "algorithmic_bytes": tot, "ms": h["ms"],
                                     "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['source']})"}
        # ---- CPU baseline (rank 0, N = 1 only) ----
        if keep_cpu:
            try:
                ips, detail, times = cpu_arm(args, CK, E, cfg, 1)
                line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                        "sample": CPU_SAMPLE_NOTE.format(tok=args.cpu_tokens, T=args.tokens), "stage_seconds": detail}
            except Exception as exc:  # the baseline is a report, never a reason to lose the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc!r}"}
        print(json.dumps(line), file=out, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _claim_stdout():
    """stdout carries exactly ONE JSON line: libraries that print from C (NCCL's version banner) go to stderr instead.
    Returns a text handle on the original stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(os.dup(2), "w", buffering=1)
    return os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    main()
