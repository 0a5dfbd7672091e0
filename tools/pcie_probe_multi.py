#!/usr/bin/env python
"""Concurrent device-to-host bandwidth of a multi-GPU box: is the end-to-end limit the host, as DESIGN.md section 5 says?

    python tools/pcie_probe_multi.py [--gpus 8] [--mib 1024] [--reps 5]

Three measurements, each the aggregate over all GPUs of pinned-memory D2H copies running at the same time:

* ``processes``   one process per GPU (what torchrun / bench.py do), every process with its own pinned buffer;
* ``threads``     one process, one thread per GPU, separate pinned buffers;
* ``one by one``  the same copies, one GPU after the other (the per-link number, no contention).

Prints one JSON object.  With 8 GPUs behind one host memory system the first two saturate well below
8 x (one by one): that ceiling, not the GPUs, bounds `e2e` at 4 and 8 GPUs.
"""
import argparse
import json
import multiprocessing as mp
import threading
import time


def _copy_loop(dev, nbytes, reps, barrier, out, key):
    import torch

    torch.cuda.set_device(dev)
    d = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{dev}")
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.copy_(d)
    torch.cuda.synchronize(dev)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(dev)
    out[key] = (t0, time.perf_counter())


def _proc(dev, nbytes, reps, barrier, q):
    out = {}
    _copy_loop(dev, nbytes, reps, barrier, out, dev)
    q.put((dev,) + out[dev])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch

    n = args.gpus or torch.cuda.device_count()
    nbytes = args.mib << 20
    total = n * nbytes * args.reps
    res = {"gpus": n, "bytes_per_copy": nbytes, "copies_per_gpu": args.reps}

    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(n), ctx.Queue()
    procs = [ctx.Process(target=_proc, args=(d, nbytes, args.reps, barrier, q)) for d in range(n)]
    for p in procs:
        p.start()
    spans = [q.get() for _ in procs]
    for p in procs:
        p.join()
    res["processes_GBps"] = round(total / (max(s[2] for s in spans) - min(s[1] for s in spans)) / 1e9, 2)
    res["processes_per_gpu_GBps"] = [round(nbytes * args.reps / (s[2] - s[1]) / 1e9, 2) for s in sorted(spans)]

    out, tb = {}, threading.Barrier(n)
    threads = [threading.Thread(target=_copy_loop, args=(d, nbytes, args.reps, tb, out, d)) for d in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    res["threads_GBps"] = round(total / (max(v[1] for v in out.values()) - min(v[0] for v in out.values())) / 1e9, 2)

    single = []
    for d in range(n):
        o, b1 = {}, threading.Barrier(1)
        _copy_loop(d, nbytes, args.reps, b1, o, d)
        single.append(round(nbytes * args.reps / (o[d][1] - o[d][0]) / 1e9, 2))
    res["one_by_one_GBps"] = single
    res["sum_of_one_by_one_GBps"] = round(sum(single), 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
