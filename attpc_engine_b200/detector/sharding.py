"""Event-range sharding across the GPUs of one box and host-side gather of the point clouds.

Events are independent (`detector/simulator.py:93-95` creates a fresh dict per event) and every random draw
is addressed by the global event number, so shard ``g`` of ``G`` simply simulates events
``[start_g, stop_g)`` on its own GPU; no collective touches the data path.  The only exchange is the
gather of the finished CSR clouds to the writer, done here on the host.
"""

from __future__ import annotations

import os
import subprocess

import numpy as np

from .engine import SimBatch


def parse_cpu_list(text: str) -> set[int]:
    """``"0-3,8,10-11"`` (the format of ``/sys/.../local_cpulist``) -> ``{0, 1, 2, 3, 8, 10, 11}``."""
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device: int) -> set[int] | None:
    """Pin the calling process to the CPUs next to GPU ``device`` (Linux; best effort, returns the CPU set or None).

    A call that brings its rows to the host is PCIe-bound, and with one process per GPU the pinned result buffers
    should live on the NUMA node the GPU's PCIe link ends in: otherwise half the ranks of a two-socket box push their
    rows across the socket interconnect.  Call it before the first simulation of the process (the library allocates
    its pinned buffers in the calling thread, so first touch puts them on that node).
    """
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(int(device))],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()  # fmt: skip
        if not bus:
            return None
        if len(bus.split(":")[0]) == 8:  # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        cpus = parse_cpu_list(open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read())
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except (OSError, ValueError, subprocess.SubprocessError):
        return None


def shard_range(n_events: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced event range of shard ``rank`` (the first ``n % world`` shards get one more)."""
    if not 0 <= rank < world:
        raise ValueError("rank outside 0..world-1")
    base, extra = divmod(int(n_events), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def concat_batches(batches: list[SimBatch]) -> SimBatch:
    """Join consecutive shards (ascending ``first_event``, no gaps) into one CSR batch."""
    batches = sorted(batches, key=lambda b: b.first_event)
    expect = batches[0].first_event
    for b in batches:
        if b.first_event != expect:
            raise ValueError(f"shards are not contiguous: expected first_event {expect}, got {b.first_event}")
        expect += len(b)

    def join(offsets_list, rows_list, labels_list, width):
        offs, base = [np.zeros(1, dtype=np.int64)], 0
        for o in offsets_list:
            offs.append(np.asarray(o[1:], dtype=np.int64) + base)
            base += int(o[-1])
        rows = np.concatenate([np.asarray(r).reshape(-1, width) for r in rows_list])
        return np.concatenate(offs), rows, np.concatenate([np.asarray(l, dtype=np.int64) for l in labels_list])

    offsets, cloud, labels = join([b.offsets for b in batches], [b.cloud for b in batches], [b.labels for b in batches], 3)
    out = SimBatch(batches[0].first_event, offsets, cloud, labels)
    if all(b.rows is not None for b in batches):
        out.row_offsets, out.rows, out.row_labels = join(
            [b.row_offsets for b in batches], [b.rows for b in batches], [b.row_labels for b in batches], 8
        )
    keys = set().union(*(b.stats.keys() for b in batches))
    out.stats = {k: sum(b.stats.get(k, 0) for b in batches) for k in keys if not k.startswith("ms_")}
    return out


def gather_to_rank0(batch: SimBatch, dist) -> SimBatch | None:
    """Gather every rank's shard on rank 0 through ``torch.distributed`` (host objects; any backend)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    payload = dict(first_event=batch.first_event, offsets=batch.offsets, cloud=batch._cloud, labels=batch._labels,
                   row_offsets=batch.row_offsets, rows=batch.rows, row_labels=batch.row_labels, stats=batch.stats,
                   columns=batch.columns)  # fmt: skip
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if rank != 0:
        return None
    return concat_batches([SimBatch(**p) for p in gathered])
