"""Writers for simulated point clouds (reference: `detector/writer.py`).

* ``SimulationWriter`` -- the plug-in protocol `run_simulation` drives (`writer.py:12-58`).
* ``convert_to_spyral`` -- same signature as the reference's njit function (`writer.py:61-112`),
  computed on the GPU.
* ``SpyralWriter`` -- same constructor, file layout, rollover and attributes as the reference
  (`writer.py:115-281`); the per-point response / threshold / z-sort runs on the GPU.  It also
  implements the optional batch hook ``write_batch`` so `run_simulation` can hand it whole batches
  whose Spyral rows were already produced by the finalize kernels.  Needs h5py at construction.
* ``ArrayWriter`` -- h5py-free bulk writer: one ``.npz`` per file with CSR rows, for hosts without
  HDF5 and for the 10M-event configuration where one dataset per event is the bottleneck.
"""

from __future__ import annotations

from pathlib import Path
from typing import Protocol

import numpy as np

from .engine import Engine, SimBatch
from .parameters import Config
from .response import get_response


class SimulationWriter(Protocol):
    """What `run_simulation` needs from a writer (`writer.py:12-58`)."""

    def write(self, data: np.ndarray, labels: np.ndarray, config: Config, event_number: int) -> None:
        """Store one event: ``data [N, 3]`` = pad id, time bucket, electrons; ``labels [N]``."""
        ...

    def get_directory_name(self) -> Path:
        """Directory the output goes to."""
        ...

    def close(self) -> None:
        """Flush and close."""
        ...


_conversion_engines: dict = {}


def _conversion_engine(window_edge, mm_edge, length, response, pad_centers, pad_sizes, threshold, device=0) -> Engine:
    response = np.ascontiguousarray(response, dtype=np.float64)
    from .engine import _digest

    key = (
        int(window_edge), int(mm_edge), float(length), float(threshold), int(device),
        _digest(response, pad_centers, pad_sizes),  # content, not id(): ids are reused once an array is freed
    )  # fmt: skip
    eng = _conversion_engines.get(key)
    if eng is None:
        eng = Engine.for_conversion(window_edge, mm_edge, length, response, pad_centers, pad_sizes, threshold, device)
        if len(_conversion_engines) > 8:
            _conversion_engines.clear()
        _conversion_engines[key] = eng
    return eng


def convert_to_spyral(
    points: np.ndarray,
    window_edge: int,
    mm_edge: int,
    length: float,
    response: np.ndarray,
    pad_centers: np.ndarray,
    pad_sizes: np.ndarray,
) -> np.ndarray:
    """``[N, 3]`` cloud -> ``[N, 8]`` Spyral rows, same order, no threshold (`writer.py:61-112`).

    Columns: x, y (pad centre, mm), z (mm), amplitude, integral, pad id, time bucket, pad size.
    """
    points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    eng = _conversion_engine(window_edge, mm_edge, length, response, pad_centers, pad_sizes, -np.inf)
    offsets = np.array([0, len(points)], dtype=np.int64)
    out = eng.convert_to_spyral(offsets, points, np.zeros(len(points), np.int64), keep_all=True)
    return out.rows


class SpyralWriter:
    """Point-cloud writer for the Spyral analysis (`writer.py:115-281`).

    Output: ``run_{n:04d}.h5`` files with group ``cloud`` holding ``cloud_{event}`` (float64
    ``[n, 8]``, attrs ``orig_run``, ``orig_event``, ``ic_amplitude`` = ``ic_multiplicity`` =
    ``ic_integral`` = ``ic_centroid`` = -1.0) and ``labels_{event}`` (int64 ``[n]``); group attrs
    ``min_event`` / ``max_event``; a new file every ``max_events_per_file`` events.
    """

    wants_spyral_rows = True
    rows_only = True  # `write_batch` never looks at the raw cloud: it can stay on the GPU
    accepts_views = True  # `write_batch` stores every array before it returns (h5py copies on create_dataset)

    def __init__(
        self,
        directory_path: Path,
        config: Config,
        max_events_per_file: int = 5_000,
        first_run_number: int = 0,
    ):
        try:
            import h5py
        except ImportError as exc:
            raise ImportError("SpyralWriter writes HDF5 and needs h5py; use ArrayWriter where it is missing") from exc
        self._h5 = h5py
        self.directory_path: Path = directory_path
        self.response: np.ndarray = get_response(config).copy()
        self.max_events_per_file: int = max_events_per_file
        self.run_number = first_run_number
        self.starting_event = 0
        self.last_event = 0
        self.events_written = 0
        self._open_file()

    def _open_file(self) -> None:
        path: Path = self.directory_path / f"run_{self.run_number:04d}.h5"
        self.file = self._h5.File(path, "w")
        self.cloud_group = self.file.create_group("cloud")

    def create_next_file(self) -> None:
        """Move to the next run number and open its file (`writer.py:180-192`)."""
        self.run_number += 1
        self._open_file()

    def _store(self, rows: np.ndarray, labels: np.ndarray, event_number: int) -> None:
        if self.events_written == self.max_events_per_file:  # `writer.py:214-218`
            self.close()
            self.create_next_file()
            self.starting_event = event_number
            self.events_written = 0
        dset = self.cloud_group.create_dataset(f"cloud_{event_number}", data=rows)
        dset.attrs["orig_run"] = self.run_number
        dset.attrs["orig_event"] = event_number
        dset.attrs["ic_amplitude"] = -1.0
        dset.attrs["ic_multiplicity"] = -1.0
        dset.attrs["ic_integral"] = -1.0
        dset.attrs["ic_centroid"] = -1.0
        self.cloud_group.create_dataset(f"labels_{event_number}", data=labels)
        self.last_event = event_number
        self.events_written += 1

    def write(self, data: np.ndarray, labels: np.ndarray, config: Config, event_number: int) -> None:
        """Rows, ADC threshold and z-sort on the GPU, then the two datasets (`writer.py:194-255`)."""
        if config.pad_centers is None:
            raise ValueError("Pad centers are not assigned at write!")
        eng = _conversion_engine(
            config.elec_params.windows_edge, config.elec_params.micromegas_edge, config.det_params.length,
            self.response, config.pad_centers, config.pad_sizes, config.elec_params.adc_threshold,
        )  # fmt: skip
        data = np.ascontiguousarray(data, dtype=np.float64).reshape(-1, 3)
        out = eng.convert_to_spyral(np.array([0, len(data)], dtype=np.int64), data, labels)
        self._store(out.rows, out.row_labels, event_number)

    def write_batch(self, batch: SimBatch, config: Config) -> None:
        """Batch hook of `run_simulation`: rows were already produced on the GPU."""
        if batch.row_offsets is None:
            raise ValueError("SpyralWriter.write_batch needs a batch simulated with spyral_rows=True")
        for e in range(len(batch)):
            if batch.offsets[e + 1] == batch.offsets[e]:  # empty clouds are skipped (`simulator.py:204`)
                continue
            rows, labels = batch.event_rows(e)
            self._store(rows, labels, batch.first_event + e)

    def set_number_of_events(self) -> None:
        """First / last event number of the current file (`writer.py:257-263`)."""
        self.cloud_group.attrs["min_event"] = self.starting_event
        self.cloud_group.attrs["max_event"] = self.last_event

    def get_directory_name(self) -> Path:
        return self.directory_path

    def close(self) -> None:
        self.set_number_of_events()
        self.file.close()


class ArrayWriter:
    """Bulk CSR writer without HDF5: ``run_{n:04d}.npz`` with ``event_numbers``, ``offsets``, ``rows``, ``labels``.

    ``rows`` are the Spyral 8-column rows (or the raw 3-column cloud with ``spyral=False``).  With
    ``directory_path=None`` nothing is written and the arrays are kept in ``self.files`` (tests).
    """

    def __init__(self, directory_path: Path | None, config: Config, max_events_per_file: int = 100_000,
                 first_run_number: int = 0, spyral: bool = True):  # fmt: skip
        self.directory_path = directory_path
        self.max_events_per_file = max_events_per_file
        self.run_number = first_run_number
        self.wants_spyral_rows = spyral
        self.rows_only = spyral  # `write_batch` then needs the offsets and the rows, not the raw cloud
        self.accepts_views = True  # `write_batch` copies what it keeps
        self.response = get_response(config).copy()
        self.files: list[dict] = []
        self._reset()

    def _reset(self) -> None:
        self._events: list[int] = []
        self._rows: list[np.ndarray] = []
        self._labels: list[np.ndarray] = []

    def _append(self, rows: np.ndarray, labels: np.ndarray, event_number: int) -> None:
        if len(self._events) == self.max_events_per_file:
            self._flush()
            self.run_number += 1
        self._events.append(int(event_number))
        self._rows.append(rows)
        self._labels.append(labels)

    def write(self, data: np.ndarray, labels: np.ndarray, config: Config, event_number: int) -> None:
        if self.wants_spyral_rows:
            eng = _conversion_engine(
                config.elec_params.windows_edge, config.elec_params.micromegas_edge, config.det_params.length,
                self.response, config.pad_centers, config.pad_sizes, config.elec_params.adc_threshold,
            )  # fmt: skip
            data = np.ascontiguousarray(data, dtype=np.float64).reshape(-1, 3)
            out = eng.convert_to_spyral(np.array([0, len(data)], dtype=np.int64), data, labels)
            self._append(out.rows, out.row_labels, event_number)
        else:
            self._append(np.array(data), np.array(labels), event_number)

    def write_batch(self, batch: SimBatch, config: Config) -> None:
        for e in range(len(batch)):
            if batch.offsets[e + 1] == batch.offsets[e]:
                continue
            rows, labels = batch.event_rows(e) if self.wants_spyral_rows else batch.event(e)
            self._append(np.array(rows), np.array(labels), batch.first_event + e)

    def _flush(self) -> None:
        if not self._events:
            return
        width = 8 if self.wants_spyral_rows else 3
        counts = np.array([len(r) for r in self._rows], dtype=np.int64)
        offsets = np.zeros(len(counts) + 1, dtype=np.int64)
        np.cumsum(counts, out=offsets[1:])
        payload = dict(
            event_numbers=np.array(self._events, dtype=np.int64),
            offsets=offsets,
            rows=np.concatenate(self._rows) if len(self._rows) else np.zeros((0, width)),
            labels=np.concatenate(self._labels) if len(self._labels) else np.zeros(0, np.int64),
            run_number=np.int64(self.run_number),
        )
        if self.directory_path is None:
            self.files.append(payload)
        else:
            Path(self.directory_path).mkdir(parents=True, exist_ok=True)
            np.savez(Path(self.directory_path) / f"run_{self.run_number:04d}.npz", **payload)
        self._reset()

    def get_directory_name(self) -> Path:
        return Path(self.directory_path) if self.directory_path is not None else Path(".")

    def close(self) -> None:
        self._flush()


class ParquetCloudWriter:
    """Bulk columnar writer of the raw point clouds: ``run_{n:04d}.parquet`` via pyarrow.

    One table row per cloud point, in event order and, inside an event, in the engine's canonical (time bucket, pad)
    order: ``event`` int64, ``pad`` int16, ``tb`` float64 (time bucket + wiggle), ``electrons`` int64, ``label`` int8.
    It takes whole batches (``write_batch``) as typed columns, so no Python loop over events or points stands between
    the GPU and the file; the per-event ``write`` of the `SimulationWriter` protocol works too.  Files roll over at
    ``max_events_per_file`` events like the reference's writer (`writer.py:214-218`).  `read_parquet_clouds` turns a
    file back into CSR arrays.
    """

    wants_columns = True  # `run_simulation`: bring the rows to the host as typed columns (11 B/row over PCIe)
    accepts_views = True  # the table is written before `write_batch` returns

    def __init__(self, directory_path: Path, config: Config | None = None, max_events_per_file: int = 1_000_000,
                 first_run_number: int = 0, compression: str = "zstd"):  # fmt: skip
        import pyarrow  # noqa: F401  (fail at construction, not at the first write)

        self.directory_path = Path(directory_path)
        self.max_events_per_file = int(max_events_per_file)
        self.run_number = int(first_run_number)
        self.compression = compression
        self._writer = None
        self._events_in_file = 0
        self.directory_path.mkdir(parents=True, exist_ok=True)

    def _table(self, event, pad, tb, electrons, label):
        import pyarrow as pa

        return pa.table({
            "event": pa.array(event, type=pa.int64()), "pad": pa.array(pad, type=pa.int16()),
            "tb": pa.array(tb, type=pa.float64()), "electrons": pa.array(electrons, type=pa.int64()),
            "label": pa.array(label, type=pa.int8()),
        })  # fmt: skip

    def _emit(self, table, n_events: int) -> None:
        import pyarrow.parquet as pq

        if self._writer is not None and self._events_in_file + n_events > self.max_events_per_file:
            self._close_file()
            self.run_number += 1
        if self._writer is None:
            path = self.directory_path / f"run_{self.run_number:04d}.parquet"
            self._writer = pq.ParquetWriter(path, table.schema, compression=self.compression)
            self._events_in_file = 0
        self._writer.write_table(table)
        self._events_in_file += n_events

    def write(self, data: np.ndarray, labels: np.ndarray, config: Config, event_number: int) -> None:
        data = np.asarray(data, dtype=np.float64).reshape(-1, 3)
        n = len(data)
        self._emit(self._table(np.full(n, int(event_number), np.int64), data[:, 0].astype(np.int16), data[:, 1],
                               data[:, 2].astype(np.int64), np.asarray(labels).astype(np.int8)), 1)  # fmt: skip

    def write_batch(self, batch: SimBatch, config: Config | None = None) -> None:
        counts = np.diff(batch.offsets)
        event = np.repeat(batch.first_event + np.arange(len(batch), dtype=np.int64), counts)
        c = batch.columns
        if c is not None:
            tb = c["tb_q16"].astype(np.float64)
            tb *= 1.0 / 65536.0
            electrons = c["electrons"] if "electrons" in c else batch._electrons().astype(np.int64)
            table = self._table(event, c["pad"], tb, electrons, c["label8"])
        else:
            cloud = batch.cloud
            table = self._table(event, cloud[:, 0].astype(np.int16), cloud[:, 1], cloud[:, 2].astype(np.int64),
                                batch.labels.astype(np.int8))  # fmt: skip
        self._emit(table, len(batch))

    def _close_file(self) -> None:
        if self._writer is not None:
            self._writer.close()
            self._writer = None

    def get_directory_name(self) -> Path:
        return self.directory_path

    def close(self) -> None:
        self._close_file()


def read_parquet_clouds(path: Path) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """``(event_numbers [E], offsets [E + 1], cloud [N, 3] float64, labels [N] int64)`` of a `ParquetCloudWriter` file
    (events without points do not appear, like in the reference's files)."""
    import pyarrow.parquet as pq

    t = pq.read_table(path)
    event = t["event"].to_numpy()
    cloud = np.empty((len(event), 3), dtype=np.float64)
    cloud[:, 0] = t["pad"].to_numpy()
    cloud[:, 1] = t["tb"].to_numpy()
    cloud[:, 2] = t["electrons"].to_numpy()
    labels = t["label"].to_numpy().astype(np.int64)
    starts = np.flatnonzero(np.r_[True, event[1:] != event[:-1]]) if len(event) else np.zeros(0, np.int64)
    offsets = np.r_[starts, len(event)].astype(np.int64)
    return event[starts].astype(np.int64), offsets, cloud, labels
