"""CPU tests of the host side: C-ABI surface, device-constant baking, mirrors of the reference's helpers."""

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from attpc_engine_b200 import _lib, nuclear_map
from attpc_engine_b200.detector import pairing
from attpc_engine_b200.detector.beam_pads import BEAM_PADS
from attpc_engine_b200.detector.engine import SimBatch, build_pad_lut, default_freeze_ke
from attpc_engine_b200.detector.response import get_response
from attpc_engine_b200.detector.sharding import concat_batches, shard_range
from attpc_engine_b200.detector.simulator import default_indices
from attpc_engine_b200.target import AnalyticGasTarget, TableGasTarget, interpolation_error
from tests.common import gas, make_config

ROOT = Path(__file__).resolve().parent.parent


def test_library_loads_and_exports_every_declared_symbol():
    """No compute call: the box running this has no GPU.  Every prototype of include/attpc_b200.h must resolve."""
    header = (ROOT / "include" / "attpc_b200.h").read_text()
    declared = set(re.findall(r"\b(attpc_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.attpc_abi_version() == _lib.ABI_VERSION
    assert re.search(r"#define ATTPC_ABI_VERSION (\d+)", header).group(1) == str(_lib.ABI_VERSION)


def test_struct_layouts_match_header():
    """ctypes mirrors have the field order of the C structs (sizes are checked against the C compiler's)."""
    header = (ROOT / "include" / "attpc_b200.h").read_text()
    for cls, cname in ((_lib.AttpcConfig, "AttpcConfig"), (_lib.AttpcSpecies, "AttpcSpecies"),
                       (_lib.AttpcReplay, "AttpcReplay"), (_lib.AttpcResult, "AttpcResult")):  # fmt: skip
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
        assert names == [f[0] for f in cls._fields_], cname


def test_struct_sizes_match_c_compiler(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include "attpc_b200.h"\nint main(void){printf("%zu %zu %zu %zu\\n", sizeof(AttpcConfig),'
        " sizeof(AttpcSpecies), sizeof(AttpcReplay), sizeof(AttpcResult)); return 0;}\n"
    )
    import subprocess

    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), "-o", str(exe), str(src)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(c) for c in (_lib.AttpcConfig, _lib.AttpcSpecies, _lib.AttpcReplay, _lib.AttpcResult)]


def test_no_gpu_means_loud_failure():
    """The product path has no CPU fallback: without a device, engine creation raises."""
    lib = _lib.load()
    if lib.attpc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from attpc_engine_b200.detector.engine import Engine

    with pytest.raises((RuntimeError, ValueError)):
        Engine(make_config(), [nuclear_map.get_data(1, 2)])


def test_pad_lut_reproduces_reference_lookup(golden_misc):
    """The 1 mm LUT (host-built, uploaded to the GPU) answers exactly like position_to_index + grid + beam veto."""
    cfg = make_config()
    lut, origin = build_pad_lut(cfg.pad_grid, cfg.pad_grid_edges)
    assert lut.shape == (559, 559) and origin == -280
    assert not np.isin(lut, BEAM_PADS).any()
    xy = golden_misc["pad_lookup/xy"]
    f = np.floor(xy * 1000.0)
    inside = np.all((f >= cfg.pad_grid_edges[0]) & (f < cfg.pad_grid_edges[1]), axis=1)
    idx = (f - origin).astype(np.int64)
    got = np.full(len(xy), -1, dtype=np.int16)
    got[inside] = lut[idx[inside, 0], idx[inside, 1]]
    assert np.array_equal(got, golden_misc["pad_lookup/pad"])


def test_response_and_pairing_mirrors(golden_misc):
    assert np.array_equal(get_response(make_config()), golden_misc["response/default"])
    assert pairing.pair(56, 937) == 937**2 + 56 and pairing.pair(937, 56) == 937**2 + 937 + 56
    assert pairing.unpair(937**2 + 56) == (56, 937) and pairing.unpair(937**2 + 937 + 56) == (937, 56)
    tb, pad, key = golden_misc["pairing/tb"], golden_misc["pairing/pad"], golden_misc["pairing/key"]
    assert np.array_equal(pairing.pair(tb, pad), key)
    utb, upad = pairing.unpair(key)
    assert np.array_equal(utb, tb) and np.array_equal(upad, pad)
    assert pairing.pair(-1, 3) == -1


def test_config_surface():
    cfg = make_config()
    assert cfg.pad_grid.shape == (5600, 5600) and cfg.pad_grid.dtype == np.int16
    assert list(cfg.pad_grid_edges) == [-280.0, 279.0, 0.1]
    assert cfg.pad_centers.shape == (10240, 2) and cfg.pad_sizes.shape == (10240,)
    assert cfg.drift_velocity == 1.0 / 550.0
    assert set(np.unique(cfg.pad_sizes)) == {0.5, 1.0}
    assert default_indices(4) == [2, 3] and default_indices(6) == [2, 4, 5] and default_indices(8) == [2, 4, 6, 7]


def test_dedx_table_scalar_and_vector_agree():
    target = gas("D2_600")
    deuteron = nuclear_map.get_data(1, 2)
    table = target.table_for(deuteron)
    ke = np.concatenate([np.logspace(-9, 3, 4001), table.nodes()[::37], [0.0, 1e-30, 5e3]])
    vec = table.evaluate(ke)
    assert np.array_equal(vec, np.array([table(float(k)) for k in ke]))
    assert np.array_equal(table.evaluate(table.nodes()), table.values)
    # the table reproduces its analytic source to the documented interpolation budget
    assert interpolation_error(AnalyticGasTarget([(1, 2, 2)], 600.0), deuteron, table) < 5e-5
    assert abs(target.density - 1.3128e-4) < 1e-7
    # range / energy-loss helpers are consistent
    loss = target.get_energy_loss(deuteron, 5.0, np.array([0.0, 0.1, 1e3]))
    assert loss[0] == 0.0 and 0.0 < loss[1] < 5.0 and loss[2] == 5.0


def test_freeze_budget_formula():
    b = default_freeze_ke(0.2, 34.0)
    n = b / 34.0e-6
    assert n + np.sqrt(0.2 * n) * 8.66 < 1.0 < (n * 1.01) + np.sqrt(0.2 * n * 1.01) * 8.66 + 0.01


def test_shard_ranges_partition_the_events():
    for n, world in ((10, 1), (10, 3), (7, 8), (1_000_000, 8), (0, 4)):
        ranges = [shard_range(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1


def _fake_batch(first, counts, seed):
    rng = np.random.default_rng(seed)
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    n = int(offsets[-1])
    return SimBatch(first, offsets, rng.random((n, 3)), rng.integers(0, 4, n), stats={"n_points": n, "ms_total": 1.0})


def test_concat_batches_rebases_offsets():
    a, b, c = _fake_batch(0, [3, 0, 2], 1), _fake_batch(3, [0, 4], 2), _fake_batch(5, [1], 3)
    whole = concat_batches([c, a, b])
    assert whole.first_event == 0 and len(whole) == 6
    assert list(whole.offsets) == [0, 3, 3, 5, 5, 9, 10]
    assert np.array_equal(whole.event(4)[0], b.event(1)[0]) and np.array_equal(whole.event(5)[1], c.event(0)[1])
    assert whole.stats["n_points"] == 10 and "ms_total" not in whole.stats
    with pytest.raises(ValueError):
        concat_batches([a, c])


def test_cpu_list_parser_and_numa_binding_is_best_effort():
    """`bind_to_gpu_numa_node` reads sysfs CPU lists; without a GPU (here) it must return None and change nothing."""
    import os

    from attpc_engine_b200.detector.sharding import bind_to_gpu_numa_node, parse_cpu_list

    assert parse_cpu_list("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert parse_cpu_list("") == set()
    before = os.sched_getaffinity(0)
    got = bind_to_gpu_numa_node(0)
    assert got is None or got <= before
    if got is None:
        assert os.sched_getaffinity(0) == before
    else:
        os.sched_setaffinity(0, before)


def test_parquet_cloud_writer_round_trip(tmp_path):
    """Bulk columnar output: typed columns -> parquet -> the same CSR clouds, file roll-over included."""
    pytest.importorskip("pyarrow")
    from attpc_engine_b200.detector import ParquetCloudWriter, read_parquet_clouds
    from attpc_engine_b200.detector.engine import SimBatch

    rng = np.random.default_rng(1)

    def batch(first, counts):
        off = np.r_[0, np.cumsum(counts)].astype(np.int64)
        n = int(off[-1])
        cols = dict(pad=rng.integers(0, 10240, n).astype(np.int16),
                    tb_q16=rng.integers(10 << 16, 500 << 16, n).astype(np.uint32),
                    label8=rng.integers(2, 4, n).astype(np.int8),
                    electrons_u32=rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32),
                    big_rows=np.array([1], np.int64), big_electrons=np.array([2**33 + 5], np.int64))  # fmt: skip
        return SimBatch(first, off, columns=cols)

    a, b = batch(100, [3, 0, 5, 2]), batch(104, [4, 4])
    w = ParquetCloudWriter(tmp_path, max_events_per_file=4)
    w.write_batch(a)
    w.write_batch(b)  # does not fit the first file any more
    w.write(b.event(1)[0], b.event(1)[1], None, 200)  # the per-event protocol works too
    w.close()
    ev, off, cloud, labels = read_parquet_clouds(tmp_path / "run_0000.parquet")
    assert list(ev) == [100, 102, 103] and list(off) == [0, 3, 8, 10]  # the empty event leaves no trace
    assert np.array_equal(cloud, a.cloud) and np.array_equal(labels, a.labels)
    assert cloud[1, 2] == 2**33 + 5
    ev, off, cloud, labels = read_parquet_clouds(tmp_path / "run_0001.parquet")
    assert list(ev) == [104, 105, 200] and list(off) == [0, 4, 8, 12]
    assert np.array_equal(cloud[:8], b.cloud) and np.array_equal(cloud[8:], b.event(1)[0])


def test_dedx_table_file_round_trip(tmp_path):
    """tabulate a gas target once (pycatima where it is installed), validate, serialise, reload."""
    from attpc_engine_b200.target import load_table_target, tabulate_gas_target

    d, c16 = nuclear_map.get_data(1, 2), nuclear_map.get_data(6, 16)
    source = AnalyticGasTarget([(1, 2, 2)], 600.0)
    made = tabulate_gas_target(source, [d, c16], tmp_path / "d2_600.npz", max_error=1e-4)
    assert set(made.errors) == {(1, 2), (6, 16)} and max(made.errors.values()) < 1e-4
    with pytest.raises(ValueError):
        tabulate_gas_target(source, [d], max_error=1e-9)
    back = load_table_target(tmp_path / "d2_600.npz")
    assert back.density == made.density and set(back.tables) == {(1, 2), (6, 16)}
    for key in back.tables:
        assert np.array_equal(back.tables[key].values, made.tables[key].values)
    assert back.get_dedx(d, 3.7) == made.get_dedx(d, 3.7)
    with pytest.raises(KeyError):
        back.get_dedx(nuclear_map.get_data(2, 4), 1.0)  # not in the file, and no source to tabulate from


def test_unknown_nuclei_raise_instead_of_getting_estimated_masses():
    """a liquid-drop mass is off by MeV; the reference's AME table raises for unknown nuclei."""
    import warnings

    from attpc_engine_b200.nuclear import NuclearDataMap

    nm = NuclearDataMap()
    assert nm.get_data(6, 16).A == 16 and nm.has_tabulated_mass(6, 16)
    with pytest.raises(KeyError, match="No tabulated mass"):
        nm.get_data(20, 48)
    nm.add_mass(20, 48, 47.95252276)
    assert abs(nm.get_data(20, 48).mass - (47.95252276 * 931.49410242 - 20 * 0.51099895)) < 1e-9
    loose = NuclearDataMap(allow_liquid_drop=True)
    with warnings.catch_warnings(record=True) as seen:
        warnings.simplefilter("always")
        est = loose.get_data(20, 48)
    assert seen and "ESTIMATE" in str(seen[0].message)
    assert abs(est.mass - nm.get_data(20, 48).mass) < 20.0  # an estimate, MeV-level error


def test_global_nuclear_map_is_looked_up_at_call_time(monkeypatch):
    """Replacing ``attpc_engine_b200.nuclear_map`` reaches the kinematics and detector front ends."""
    import attpc_engine_b200
    from attpc_engine_b200.kinematics import Reaction
    from attpc_engine_b200.nuclear import NuclearDataMap

    class Spy(NuclearDataMap):
        asked = []

        def get_data(self, z, a):
            Spy.asked.append((int(z), int(a)))
            return super().get_data(z, a)

    spy = Spy()
    monkeypatch.setattr(attpc_engine_b200, "nuclear_map", spy)
    Reaction(target=spy.get_data(1, 2), projectile=spy.get_data(6, 16), ejectile=spy.get_data(1, 2))
    assert (6, 16) in Spy.asked[3:]  # the residual was resolved through the replaced global


def test_config_fingerprint_sees_every_baked_value():
    """`engine_for` must not reuse an engine whose baked constants went stale."""
    from attpc_engine_b200 import nuclear_map
    from attpc_engine_b200.detector.engine import config_fingerprint
    from tests.common import make_config

    cfg = make_config()
    nuclei = [nuclear_map.get_data(1, 2), nuclear_map.get_data(6, 16)]
    base = config_fingerprint(cfg, nuclei)
    assert config_fingerprint(cfg, nuclei) == base
    for obj, field, value in (
        (cfg.det_params, "bfield", 2.5), (cfg.det_params, "efield", 50000.0), (cfg.det_params, "diffusion", 0.3),
        (cfg.det_params, "fano_factor", 0.25), (cfg.det_params, "w_value", 30.0), (cfg.det_params, "mpgd_gain", 1000),
        (cfg.elec_params, "adc_threshold", 10), (cfg.elec_params, "shaping_time", 500), (cfg.elec_params, "amp_gain", 500),
    ):  # fmt: skip
        old = getattr(obj, field)
        setattr(obj, field, value)
        assert config_fingerprint(cfg, nuclei) != base, field
        setattr(obj, field, old)
    assert config_fingerprint(cfg, nuclei) == base
    cfg.pad_sizes = cfg.pad_sizes * 2.0  # a replaced array (what load_pad_grid does) is seen
    assert config_fingerprint(cfg, nuclei) != base
    other_gas = make_config("He4_600")
    assert config_fingerprint(other_gas, nuclei) != base


def test_ordered_fan_in_of_the_multi_gpu_driver():
    """`run_simulation(devices=[...])`: batches come back in order whatever the workers' speed; a worker starts its
    next batch only after the writer took its previous one (the batch may be a view); errors reach the caller."""
    import threading
    import time

    from attpc_engine_b200.detector.simulator import _OrderedFanIn

    log, lock = [], threading.Lock()

    def work(k, g):
        time.sleep(0.002 * ((7 * k) % 5))  # uneven speeds
        with lock:
            log.append(("start", k, g))
        return k * 10

    fan = _OrderedFanIn(23, [0, 1, 2], work)
    seen = []
    for k, batch in fan:
        with lock:
            log.append(("take", k))
        seen.append((k, batch))
    fan.close()
    assert seen == [(k, 10 * k) for k in range(23)]
    for k in range(3, 23):  # worker g = k % 3 started batch k only after batch k - 3 had been taken
        assert log.index(("take", k - 3)) < log.index(("start", k, k % 3))
    assert {g for _, k, g in (e for e in log if e[0] == "start") if True} == {0, 1, 2}

    def failing(k, g):
        if k == 5:
            raise ValueError("boom")
        return k

    fan = _OrderedFanIn(9, [0, 1], failing)
    with pytest.raises(ValueError, match="boom"):
        for _ in fan:
            pass
    fan.close()


def test_hdf5_kinematics_reader_on_a_reference_written_file(monkeypatch):
    """`_Hdf5Kinematics` (`detector/simulator.py:146-196` of the reference) on a file the REFERENCE's own
    `run_kinematics_pipeline` wrote (tests/golden/make_kinematics_file.py; h5py replaced by the in-memory stand-in,
    23 events in four chunk groups): attributes, chunk arithmetic, datasets and vertex attributes."""
    import sys
    import types

    from attpc_engine_b200.detector.simulator import _Hdf5Kinematics, default_indices
    from tests.common import GOLDEN

    sys.path.insert(0, str(GOLDEN))
    import ref_shim

    with np.load(GOLDEN / "kinematics_file.npz") as f:
        flat = {k: f[k] for k in f.files}
    root = ref_shim.MemGroup()

    def node(path):  # "/data/chunk_0" -> the MemGroup, created on the way
        cur = root
        for part in [p for p in path.split("/") if p]:
            if part not in cur:
                cur.create_group(part)
            cur = cur[part]
        return cur

    for key in sorted(flat):
        kind, path, *rest = key.split("|")
        if kind == "group":
            node(path)
        elif kind == "data":
            parent, name = path.rsplit("/", 1)
            node(parent).create_dataset(name, data=flat[key])
    for key in sorted(flat):
        kind, path, *rest = key.split("|")
        if kind == "attr":
            parent, name = path.rsplit("/", 1) if "/" in path.strip("/") else ("", path.strip("/"))
            target = node(parent)[name] if name in node(parent) else node(path)
            value = flat[key]
            target.attrs[rest[0]] = value.item() if value.ndim == 0 else value
    fake = types.ModuleType("h5py")
    fake.File = lambda path, mode="r": root
    monkeypatch.setitem(sys.modules, "h5py", fake)
    kin = _Hdf5Kinematics("whatever.h5")
    assert kin.n_events == 23 and kin.chunk_size == 7
    assert list(kin.proton_numbers) == [1, 6, 1, 6, 6, 0] and list(kin.mass_numbers) == [2, 14, 1, 15, 14, 1]
    assert default_indices(len(kin.proton_numbers)) == [2, 4, 5]
    data, vertices = kin.read(0, 23)
    assert np.array_equal(data, flat["expected|momenta"]) and np.array_equal(vertices, flat["expected|vertices"])
    data, vertices = kin.read(5, 16)  # a range that crosses two chunk boundaries
    assert np.array_equal(data, flat["expected|momenta"][5:16]) and np.array_equal(vertices, flat["expected|vertices"][5:16])


def test_simulate_stream_deals_batches_to_workers_and_keeps_order(monkeypatch):
    """`simulate_stream`: batch k -> worker k mod W (GPUs in turn, `engines_per_device` engines each), results in
    ascending order, callables are read inside the worker; one worker means no threads at all."""
    import threading

    from attpc_engine_b200.detector import simulator

    calls = []

    def fake(momenta, vertices, zs, as_, config, seed, indices, first_event=0, device=0, engine_instance=0, **kw):
        calls.append((first_event, device, engine_instance, threading.current_thread() is threading.main_thread(), kw))
        return ("batch", first_event)

    monkeypatch.setattr(simulator, "simulate_batch", fake)
    assert simulator._workers([0, 1], 2) == [0, 1, 0, 1]
    with pytest.raises(ValueError):
        simulator._workers([0], 0)
    batches = [(lambda k=k: (np.zeros((4, 2, 4)), np.zeros((4, 3)), 100 * k)) if k % 2 else
               (np.zeros((4, 2, 4)), np.zeros((4, 3)), 100 * k) for k in range(11)]
    out = list(simulator.simulate_stream(batches, [1, 6], [2, 16], None, 7, [0, 1], devices=[3, 5], engines_per_device=2,
                                         columns=True))
    assert out == [(k, ("batch", 100 * k)) for k in range(11)]
    by_event = {c[0]: c for c in calls}
    for k in range(11):
        first, device, instance, on_main, kw = by_event[100 * k]
        assert (device, instance) == ([3, 5, 3, 5][k % 4], k % 4) and not on_main and kw == {"columns": True}
    calls.clear()
    out = list(simulator.simulate_stream(batches[:3], [1, 6], [2, 16], None, 7, [0, 1], devices=[2], engines_per_device=1))
    assert [o[0] for o in out] == [0, 1, 2] and all(c[3] and c[1:3] == (2, 0) for c in calls)


def test_packed_columns_decode():
    """`SimBatch` with the packed wire form (pad id + rank in 16 bits, wiggle, rows per time bucket and event): the
    typed columns, the float64 cloud and single events decoded from it."""
    from attpc_engine_b200.detector.engine import SimBatch

    rng = np.random.default_rng(3)
    n_ev = 5
    counts = np.zeros((n_ev, 512), dtype=np.uint16)
    for e in range(n_ev):
        tbs = rng.choice(512, size=rng.integers(0, 40), replace=False)
        counts[e, tbs] = rng.integers(1, 9, size=len(tbs))
    counts[3] = 0  # an empty event
    per_event = counts.sum(axis=1).astype(np.int64)
    offsets = np.concatenate([[0], np.cumsum(per_event)])
    n = int(offsets[-1])
    pad = rng.integers(0, 10240, size=n).astype(np.uint16)
    rank = rng.integers(0, 3, size=n).astype(np.uint16)
    wiggle = rng.integers(0, 65536, size=n).astype(np.uint16)
    e32 = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    big_rows = np.sort(rng.choice(n, size=7, replace=False)).astype(np.int64)
    big_e = (rng.integers(1, 2**10, size=7).astype(np.int64) << 32) | e32[big_rows].astype(np.int64)
    packed = dict(pad_rank=pad | (rank << 14), wiggle=wiggle, tb_counts=counts, rank_shift=14,
                  labels_of_rank=np.array([2, 3, 5, 0], dtype=np.int8), electrons_u32=e32, big_rows=big_rows,
                  big_electrons=big_e)  # fmt: skip
    tb = np.concatenate([np.repeat(np.arange(512), counts[e]) for e in range(n_ev)])
    electrons = e32.astype(np.float64)
    electrons[big_rows] = big_e
    batch = SimBatch(100, offsets, packed=packed)
    for e in range(n_ev):  # single events first: nothing else is decoded
        a, b = offsets[e], offsets[e + 1]
        cloud, labels = batch.event(e)
        assert batch._columns is None
        assert np.array_equal(cloud[:, 0], pad[a:b]) and np.array_equal(cloud[:, 1], tb[a:b] + wiggle[a:b] / 65536.0)
        assert np.array_equal(cloud[:, 2], electrons[a:b])
        assert np.array_equal(labels, np.array([2, 3, 5])[rank[a:b]]) and labels.dtype == np.int64
    from attpc_engine_b200.detector import _decode

    if _decode.HAVE_NUMBA:  # the one-pass decoder and the numpy route give the same arrays
        fast = SimBatch(100, offsets, packed=packed)
        assert fast._decode_packed() and fast._columns is None
        saved, _decode.HAVE_NUMBA = _decode.HAVE_NUMBA, False
        try:
            slow = SimBatch(100, offsets, packed=packed)
            assert not slow._decode_packed()
            assert np.array_equal(fast.cloud, slow.cloud) and np.array_equal(fast.labels, slow.labels)
            assert fast.cloud.dtype == slow.cloud.dtype == np.float64 and fast.labels.dtype == slow.labels.dtype == np.int64
        finally:
            _decode.HAVE_NUMBA = saved
    cols = batch.columns
    assert cols["pad"].dtype == np.int16 and cols["tb_q16"].dtype == np.uint32 and cols["label8"].dtype == np.int8
    assert np.array_equal(cols["pad"], pad) and np.array_equal(cols["tb_q16"], (tb.astype(np.uint32) << 16) | wiggle)
    assert np.array_equal(batch.cloud[:, 2], electrons) and np.array_equal(batch.labels, np.array([2, 3, 5])[rank])
