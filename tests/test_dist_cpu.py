"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch of clouds, run the (oracle) per-cloud path on
their shard with no data-path collective, and the gathered result equals the single-process result; the
max-over-ranks timing reduction returns the slowest rank."""

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from si_mamba_b200.shard import gather_predictions, max_over_ranks, shard_range


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 32, 33):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle import tokenizer
    xyz = tokenizer.synthetic_clouds(6, 128, 3, "ball")        # every rank builds the same global batch
    lo, hi = shard_range(xyz.shape[0], rank, world)
    local = tokenizer.fps(xyz[lo:hi], 8)                       # per-cloud work on the shard: no collective
    full = gather_predictions(local)
    slowest = max_over_ranks(10.0 + rank)
    if rank == 0:
        q.put((full, slowest))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, slowest = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import tokenizer
    ref = tokenizer.fps(tokenizer.synthetic_clouds(6, 128, 3, "ball"), 8)
    assert torch.equal(full, ref)
    assert slowest == 11.0
