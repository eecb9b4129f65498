"""CPU, world_size 2, gloo: the data-parallel host logic (sample sharding + the single all-gather of decoded ids)
gives the same ids in the same order as one process."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_decode(sample_index: int, T: int):
    g = torch.Generator().manual_seed(sample_index)
    n = int(torch.randint(1, T + 1, (1,), generator=g))
    row = torch.full((T,), 7, dtype=torch.int32)
    row[:n] = torch.randint(0, 1000, (n,), generator=g, dtype=torch.int32)
    return row, n


def _worker(rank, world, port, n_total, T, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from importlib import import_module
    import fo1_b200  # noqa: F401
    DP = import_module("vlm-fo1_b200.dp")
    a, b = DP.shard_range(n_total, rank, world)
    rows = [_fake_decode(i, T) for i in range(a, b)]
    toks = torch.stack([r for r, _ in rows]) if rows else torch.zeros((0, T), dtype=torch.int32)
    lens = torch.tensor([n for _, n in rows], dtype=torch.int32)
    per = max(DP.shard_range(n_total, r, world)[1] - DP.shard_range(n_total, r, world)[0] for r in range(world))
    at, al = DP.gather_ids(toks, lens, world, per)
    if rank == 0:
        q.put((at.tolist(), al.tolist(), per))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process():
    n_total, T, world = 7, 12, 2          # odd count: ragged shards (4 + 3)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    at, al, per = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from importlib import import_module
    import fo1_b200  # noqa: F401
    DP = import_module("vlm-fo1_b200.dp")
    assert [DP.shard_range(n_total, r, world) for r in range(world)] == [(0, 4), (4, 7)]
    got = []
    for r in range(world):
        a, b = DP.shard_range(n_total, r, world)
        for k in range(b - a):
            got.append((at[r * per + k], al[r * per + k]))
        for k in range(b - a, per):
            assert al[r * per + k] == 0 and all(v == -1 for v in at[r * per + k])
    ref = [_fake_decode(i, T) for i in range(n_total)]
    assert [(r.tolist(), n) for r, n in ref] == got
