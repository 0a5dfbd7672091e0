"""Event drivers of the detector simulation (reference: `detector/simulator.py`).

``simulate`` and ``run_simulation`` keep the reference's signatures; ``simulate_batch`` is the
array-level entry the CUDA path is built around.  All three run on the GPU through
``engine.Engine``; there is no CPU implementation here.

Randomness: the reference threads one unseeded PCG64 generator through all events
(`simulator.py:169`).  Here every draw is a Philox4x32-10 value addressed by
``(seed; event number, nucleus index, grid step)`` or ``(seed; event number, pad/time key)``, so a
given ``seed`` reproduces a run bit-for-bit however it is batched or sharded over GPUs.
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np
from numpy.random import Generator, default_rng

import attpc_engine_b200 as _pkg
from .engine import SimBatch, engine_for
from .parameters import Config
from .writer import SimulationWriter


@dataclass
class SimEvent:
    """One simulated event: the ``(cloud, labels)`` pair `simulate` returns, plus its number."""

    event_number: int
    cloud: np.ndarray  # [N, 3] pad id, time bucket, electrons
    labels: np.ndarray  # [N] index of the nucleus that produced the point

    def __iter__(self):
        return iter((self.cloud, self.labels))


def default_indices(n_nuclei: int) -> list[int]:
    """All final products: ejectile, every decay's first product, and the last nucleus (`simulator.py:152-158`)."""
    picks = list(range(2, n_nuclei, 2))
    picks.append(n_nuclei - 1)
    return picks


def _nuclei_for(proton_numbers, mass_numbers, indices, nmap) -> list:
    out = []
    for idx in indices:
        if proton_numbers[idx] == 0:  # neutrons leave no track (`simulator.py:97`)
            continue
        out.append(nmap.get_data(int(proton_numbers[idx]), int(mass_numbers[idx])))
    return out


def simulate_batch(
    momenta: np.ndarray,
    vertices: np.ndarray,
    proton_numbers: np.ndarray,
    mass_numbers: np.ndarray,
    config: Config,
    seed: int,
    indices: list[int],
    first_event: int = 0,
    device: int = 0,
    spyral_rows: bool = False,
    copy: bool = True,
    nuclear_data=None,
    rows_only: bool = False,
    columns: bool = False,
    row_columns: bool = False,
    engine_instance: int = 0,
    packed: bool = True,
    **tuning,
) -> SimBatch:
    """Detector simulation of ``B`` kinematics events in one call.

    ``momenta [B, K, 4]`` (px, py, pz, E in MeV), ``vertices [B, 3]`` (m).  Event ``e`` of the batch
    is global event ``first_event + e`` for the random streams.  Rows of each event come out in
    ascending (time bucket, pad) order (the reference's order is dict insertion order).
    """
    nmap = nuclear_data if nuclear_data is not None else _pkg.nuclear_map  # looked up per call: replaceable
    momenta = np.asarray(momenta, dtype=np.float64)
    if momenta.ndim != 3:
        raise ValueError("momenta must have shape [n_events, n_nuclei, 4]")
    charged = _nuclei_for(proton_numbers, mass_numbers, indices, nmap)
    if not charged or momenta.shape[0] == 0:
        n = momenta.shape[0]
        empty = SimBatch(first_event, np.zeros(n + 1, np.int64), np.zeros((0, 3)), np.zeros(0, np.int64))
        if spyral_rows:
            empty.row_offsets, empty.rows, empty.row_labels = empty.offsets.copy(), np.zeros((0, 8)), empty.labels
        return empty
    engine = engine_for(config, charged, device=device, instance=engine_instance, **tuning)
    return engine.simulate_batch(
        momenta, vertices, proton_numbers, mass_numbers, indices, seed=seed, first_event=first_event,
        spyral_rows=spyral_rows, copy=copy, rows_only=rows_only, columns=columns, row_columns=row_columns, packed=packed,
    )  # fmt: skip


def simulate(
    momenta: np.ndarray,
    vertex: np.ndarray,
    proton_numbers: np.ndarray,
    mass_numbers: np.ndarray,
    config: Config,
    rng: Generator,
    indices: list[int],
) -> tuple[np.ndarray, np.ndarray]:
    """Apply the detector simulation to one kinematics event (`simulator.py:52-115`).

    Same arguments and return value as the reference: ``(cloud [N, 3], labels [N])`` with rows
    ``[pad id, time bucket, electrons]``.  ``rng`` only supplies the 63-bit seed of the event's
    counter-based streams (one ``rng.integers`` draw per call).
    """
    seed = int(rng.integers(0, 2**63 - 1))
    batch = simulate_batch(
        np.asarray(momenta, dtype=np.float64)[None], np.asarray(vertex, dtype=np.float64)[None],
        proton_numbers, mass_numbers, config, seed, indices,
    )  # fmt: skip
    return batch.event(0)


class _ArrayKinematics:
    """Kinematics events held as arrays: ``data [n, K, 4]``, ``vertices [n, 3]`` (+ Z, A)."""

    def __init__(self, data, vertices, proton_numbers, mass_numbers):
        self.data, self.vertices = data, vertices
        self.proton_numbers, self.mass_numbers = np.asarray(proton_numbers), np.asarray(mass_numbers)
        self.n_events = len(data)

    def read(self, start: int, stop: int):
        return self.data[start:stop], self.vertices[start:stop]


class _Hdf5Kinematics:
    """The reference's kinematics file (`kinematics/pipeline.py:449-493`), read in event ranges."""

    def __init__(self, path: Path):
        try:
            import h5py
        except ImportError as exc:  # pragma: no cover - depends on the environment
            raise ImportError(
                "reading HDF5 kinematics needs h5py; use a .npz kinematics file "
                "(attpc_engine_b200.kinematics.save_kinematics_npz) where h5py is unavailable"
            ) from exc
        self.file = h5py.File(path, "r")
        self.group = self.file["data"]
        self.proton_numbers = np.asarray(self.group.attrs["proton_numbers"])
        self.mass_numbers = np.asarray(self.group.attrs["mass_numbers"])
        self.n_events = int(self.group.attrs["n_events"])
        self.chunk_size = int(self.group.attrs["chunk_size"])

    def read(self, start: int, stop: int):
        k = len(self.proton_numbers)
        data = np.empty((stop - start, k, 4))
        vertices = np.empty((stop - start, 3))
        for i, ev in enumerate(range(start, stop)):
            dset = self.group[f"chunk_{ev // self.chunk_size}"][f"event_{ev}"]
            data[i] = dset[:]
            vertices[i] = (dset.attrs["vertex_x"], dset.attrs["vertex_y"], dset.attrs["vertex_z"])
        return data, vertices


def _open_kinematics(input_path):
    if isinstance(input_path, _ArrayKinematics):
        return input_path
    path = Path(input_path)
    if path.suffix == ".npz":
        with np.load(path) as f:
            return _ArrayKinematics(f["data"], f["vertices"], f["proton_numbers"], f["mass_numbers"])
    return _Hdf5Kinematics(path)


class _OrderedFanIn:
    """Batches simulated by one worker thread per GPU, handed to the writer in ascending batch order.

    Batch ``k`` goes to device ``k mod G``.  A worker starts its next batch only after the writer has taken the
    previous one, so a batch may hold views of its engine's pinned buffers until then (no copy), and at most one
    finished batch per GPU waits in host memory.  An exception in a worker is re-raised in the caller.
    """

    def __init__(self, n_batches: int, devices: list[int], work):
        import threading

        self.n_batches, self.devices, self.work = n_batches, list(devices), work
        self.cond = threading.Condition()
        self.ready: dict[int, object] = {}
        self.taken = -1  # highest batch index the writer is done with
        self.error: BaseException | None = None
        self.threads = [threading.Thread(target=self._run, args=(g,), daemon=True) for g in range(len(self.devices))]
        for t in self.threads:
            t.start()

    def _run(self, g: int) -> None:
        try:
            for k in range(g, self.n_batches, len(self.devices)):
                with self.cond:  # the writer must be done with this worker's previous batch (it may be a view)
                    self.cond.wait_for(lambda: self.taken >= k - len(self.devices) or self.error is not None)
                    if self.error is not None:
                        return
                batch = self.work(k, g)
                with self.cond:
                    self.ready[k] = batch
                    self.cond.notify_all()
        except BaseException as exc:  # noqa: BLE001 - handed to the caller
            with self.cond:
                self.error = exc
                self.cond.notify_all()

    def __iter__(self):
        for k in range(self.n_batches):
            with self.cond:
                self.cond.wait_for(lambda: k in self.ready or self.error is not None)
                if self.error is not None:
                    raise self.error
                batch = self.ready.pop(k)
            yield k, batch
            with self.cond:
                self.taken = k
                self.cond.notify_all()

    def close(self) -> None:
        with self.cond:
            if self.error is None:
                self.error = GeneratorExit()
            self.cond.notify_all()
        for t in self.threads:
            t.join(timeout=60)


def _workers(devices, engines_per_device: int) -> list[int]:
    """Device of every worker: the GPUs in turn, `engines_per_device` times (consecutive batches go to different GPUs)."""
    if engines_per_device < 1:
        raise ValueError("engines_per_device must be at least 1")
    return [int(d) for _ in range(int(engines_per_device)) for d in devices]


def simulate_stream(
    batches,
    proton_numbers: np.ndarray,
    mass_numbers: np.ndarray,
    config: Config,
    seed: int,
    indices: list[int],
    devices=(0,),
    engines_per_device: int = 2,
    **batch_options,
):
    """Pipelined `simulate_batch` over a sequence of batches: yields ``(k, SimBatch)`` in ascending ``k``.

    ``batches[k]`` is ``(momenta [B, K, 4], vertices [B, 3], first_event)`` or a callable returning that (called in
    the worker, e.g. to read a chunk of a file; one at a time unless it has a true ``thread_safe`` attribute).
    Batch ``k`` goes to worker ``k mod W``, ``W = len(devices) * engines_per_device`` workers, each a host thread
    with its own engine (its own streams and pinned buffers).  Two
    engines on one GPU overlap the device-to-host copy of one batch with the kernels of the next: the end-to-end
    rate of a single GPU is otherwise bounded by ``compute + copy`` of one batch at a time.  A yielded batch is valid
    until the consumer asks for the next one (its worker starts on batch ``k + W`` only then), so ``copy=False``
    views are safe; ``batch_options`` are those of `simulate_batch`.  Every random draw is addressed by the global
    event number: the output does not depend on ``devices`` or ``engines_per_device``.
    """
    import threading

    batches = list(batches) if not hasattr(batches, "__getitem__") else batches
    workers = _workers(devices, engines_per_device)
    read_lock = threading.Lock()  # (h5py handles are not thread-safe)

    def work(k: int, g: int) -> SimBatch:
        item = batches[k]
        if callable(item):
            if getattr(item, "thread_safe", False):
                item = item()
            else:
                with read_lock:
                    item = item()
        momenta, vertices, first_event = item
        return simulate_batch(momenta, vertices, proton_numbers, mass_numbers, config, seed, indices,
                              first_event=first_event, device=workers[g], engine_instance=g, **batch_options)  # fmt: skip

    if len(workers) == 1:
        for k in range(len(batches)):
            yield k, work(k, 0)
        return
    fan_in = _OrderedFanIn(len(batches), workers, work)
    try:
        yield from fan_in
    finally:
        fan_in.close()


def run_simulation(
    config: Config,
    input_path: Path,
    writer: SimulationWriter,
    indices: list[int] | None = None,
    seed: int | None = None,
    batch_size: int = 16384,
    device: int = 0,
    verbose: bool = True,
    devices: list[int] | None = None,
    engines_per_device: int = 2,
) -> None:
    """Run the detector simulation over a kinematics file (`simulator.py:118-210`).

    ``input_path``: the reference's HDF5 kinematics file (needs h5py) or an ``.npz`` with ``data``,
    ``vertices``, ``proton_numbers``, ``mass_numbers``.  ``writer.write(cloud, labels, config,
    event_number)`` is called once per non-empty event in ascending event order and
    ``writer.close()`` at the end, exactly like the reference; a writer that also defines
    ``write_batch(batch, config)`` receives whole :class:`SimBatch` objects instead (with the Spyral
    rows already computed on the GPU when it sets ``wants_spyral_rows = True``).  The arrays of that
    batch are the writer's to keep.  A writer that is done with them when ``write_batch`` returns can set
    ``accepts_views = True``: it is then handed views of the engine's pinned host buffers, valid only until
    the next batch is simulated (no copy; the built-in writers do this).
    ``seed=None`` draws one from the OS, like the reference's unseeded generator.

    ``devices=[0, 1, ...]``: the event ranges (batches of ``batch_size`` events) are dealt round-robin to one worker
    thread and one engine per listed GPU; the writer still sees every event in ascending order, and -- every random
    draw being addressed by the global event number -- the output does not depend on the number of GPUs.
    ``engines_per_device`` (default 2): engines, each with its own worker thread, per GPU, so that the copy of one batch
    to the host and the writer overlap the kernels of the next batch (`simulate_stream`).
    """
    kin = _open_kinematics(input_path)
    devices = [int(device)] if not devices else [int(d) for d in devices]
    if verbose:
        print("------- AT-TPC Simulation Engine (B200) -------")
        print(f"Applying detector effects to kinematics from file: {input_path}")
        print(f"Found {kin.n_events} kinematics events.")
        print(f"Output will be written to {writer.get_directory_name()}.")
        if len(devices) * engines_per_device > 1:
            print(f"Event ranges of {batch_size} events are dealt to GPUs {devices}, {engines_per_device} engine(s) each.")
    nuclei_to_sim = list(indices) if indices is not None else default_indices(len(kin.proton_numbers))
    if seed is None:
        seed = int(default_rng().integers(0, 2**63 - 1))
    batched = hasattr(writer, "write_batch")
    want_rows = bool(getattr(writer, "wants_spyral_rows", False)) and batched
    views_ok = not batched or bool(getattr(writer, "accepts_views", False))  # per-event arrays are built fresh anyway
    starts = list(range(0, kin.n_events, batch_size))

    def reader(k: int):  # batch k, read in the worker that simulates it
        def read():
            start, stop = starts[k], min(starts[k] + batch_size, kin.n_events)
            momenta, vertices = kin.read(start, stop)
            return momenta, vertices, start

        return read

    stream = simulate_stream(
        [reader(k) for k in range(len(starts))], kin.proton_numbers, kin.mass_numbers, config, seed, nuclei_to_sim,
        devices=devices, engines_per_device=engines_per_device, spyral_rows=want_rows, copy=not views_ok,
        rows_only=want_rows and bool(getattr(writer, "rows_only", False)),
        row_columns=want_rows,  # 13 instead of 72 B/row over PCIe; `SimBatch.event_rows` rebuilds the float64 rows
        # per-event writers get their arrays built event by event anyway; batch writers say if they want columns
        columns=not batched or bool(getattr(writer, "wants_columns", False)),
    )  # fmt: skip

    def consume(k: int, batch: SimBatch) -> None:
        if batched:
            writer.write_batch(batch, config)
            return
        for e in range(len(batch)):
            cloud, labels = batch.event(e)
            if len(cloud) == 0:  # `simulator.py:204`
                continue
            writer.write(np.array(cloud), np.array(labels), config, starts[k] + e)

    try:
        for k, batch in stream:
            consume(k, batch)
    finally:
        stream.close()
    writer.close()
    if verbose:
        print("Done.")
        print("----------------------------------------")
