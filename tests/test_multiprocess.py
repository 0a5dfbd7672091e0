"""N > 1 host path on CPU: two processes (gloo), event-range shards, gather of the CSR clouds on rank 0."""

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from attpc_engine_b200.detector.engine import SimBatch
from attpc_engine_b200.detector.sharding import concat_batches, gather_to_rank0, shard_range

N_EVENTS = 11


def _shard_batch(start, stop):
    """Deterministic stand-in for a simulated shard: content depends only on the GLOBAL event number."""
    counts = [(ev * 7) % 5 for ev in range(start, stop)]
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    cloud = np.concatenate([np.full((c, 3), float(ev)) for ev, c in zip(range(start, stop), counts)] + [np.zeros((0, 3))])
    labels = np.concatenate([np.full(c, ev, dtype=np.int64) for ev, c in zip(range(start, stop), counts)] + [np.zeros(0, np.int64)])
    return SimBatch(start, offsets, cloud, labels, stats={"n_points": int(offsets[-1])})


def _worker(rank, world, port, out):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    start, stop = shard_range(N_EVENTS, rank, world)
    merged = gather_to_rank0(_shard_batch(start, stop), dist)
    dist.barrier()
    if rank == 0:
        out.put((merged.first_event, merged.offsets, merged.cloud, merged.labels, merged.stats))
    else:
        assert merged is None
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (np.random.default_rng().integers(0, 2000))
    procs = [ctx.Process(target=_worker, args=(r, 2, int(port), out)) for r in range(2)]
    for p in procs:
        p.start()
    first, offsets, cloud, labels, stats = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    whole = concat_batches([_shard_batch(0, N_EVENTS)])
    assert first == 0
    assert np.array_equal(offsets, whole.offsets)
    assert np.array_equal(cloud, whole.cloud) and np.array_equal(labels, whole.labels)
    assert stats["n_points"] == whole.stats["n_points"]
