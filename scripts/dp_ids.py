"""Decode a fixed list of synthetic samples data-parallel over WORLD_SIZE ranks (one GPU each) and have rank 0 write the gathered
ids: the same list must give the same ids for every rank count (tests/test_gpu_dp_ids.py).
Reduced depth, real widths (the point is the plumbing + batch-composition independence of the kernels, not the model size)."""
import argparse
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--samples", type=int, default=6)
    ap.add_argument("--tokens", type=int, default=12)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    from importlib import import_module
    import fo1_b200  # noqa: F401
    E = import_module("vlm-fo1_b200.engine"); CK = import_module("vlm-fo1_b200.checkpoint"); DP = import_module("vlm-fo1_b200.dp")
    P = import_module("vlm-fo1_b200.pipeline"); SY = import_module("vlm-fo1_b200.synthetic")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = E.EngineConfig()
    cfg.vit = dict(cfg.vit, depth=4, fullatt_block_indexes=[1, 3])
    cfg.davit = dict(cfg.davit, depths=[1, 1, 1, 1])
    cfg.llm = dict(cfg.llm, num_hidden_layers=4)
    eng = CK.load_engine(cfg, CK.random_state_dicts(cfg, dev, 0), dev)
    pipe = P.Fo1Pipeline(eng)
    sizes = [448, 644, 448, 896, 644, 448, 448, 644]          # ragged grids incl. the 46x46 one (ragged windows, unaligned tiles)
    lo, hi = DP.shard_range(args.samples, rank, world)
    per = max(DP.shard_range(args.samples, r, world)[1] - DP.shard_range(args.samples, r, world)[0] for r in range(world))
    host = [SY.synthetic_sample(i, sizes[i % len(sizes)], 5 + 3 * i) for i in range(lo, hi)]
    res = pipe.generate([P.SampleInputs(s.input_ids, s.pixel_values.to(dev), s.grid_hw, s.image_aux.to(dev), s.boxes.to(dev)) for s in host],
                        args.tokens, stop_ids=[], early_exit_interval=0)
    toks, lens = DP.gather_ids(res["tokens"], res["lens"], world, per)
    torch.cuda.synchronize()
    if rank == 0:
        rows = []
        for r in range(world):                      # drop the padding rows of ragged shards, keep global sample order
            a, b = DP.shard_range(args.samples, r, world)
            rows.append(toks[r * per: r * per + (b - a)].cpu().numpy())
        np.save(args.out, np.concatenate(rows, 0))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
