"""GPU parity, part (a): identical trajectory points + replayed random numbers -> bit-exact pad ids,
time buckets and pairing keys, charges within 1e-6 relative (they come out equal).

Every case compares the CUDA path, called through the C ABI, with fixtures recorded from the
UNMODIFIED reference (tests/golden/make_golden.py)."""

import numpy as np
import pytest

from attpc_engine_b200.detector.pairing import unpair
from tests.common import (
    WORKLOAD_NAMES, case_config, case_names, case_tracks, cloud_keys, load_workload, nuclei_of, reference_cloud_from_dict,
    sort_cloud, workload_config, workload_event_dict, workload_tracks,
)  # fmt: skip

pytestmark = pytest.mark.gpu

CHARGE_RTOL = 1e-6  # north_star tolerance for charges; pad / tb / key must be exact


def _replay(ev, name, tuning=None, **kw):
    from attpc_engine_b200.detector.engine import engine_for

    cfg = case_config(name)
    tracks = case_tracks(ev, name)
    if not tracks:
        pytest.skip("no charged track")
    eng = engine_for(cfg, nuclei_of(tracks), **(tuning or {}))
    batch, electrons = eng.simulate_replay(
        [t["rows"] for t in tracks], [t["normals"] for t in tracks], [0] * len(tracks),
        [t["rank"] for t in tracks], [t["idx"] for t in tracks], [t["za"] for t in tracks], 1,
        uniforms=[(ev[f"{name}/keys"], ev[f"{name}/uniforms"])], **kw,
    )  # fmt: skip
    return cfg, tracks, batch, electrons


@pytest.mark.parametrize("name", case_names())
def test_electrons_bit_exact(golden_events, name):
    """`generate_electrons` with replayed normals (`solver.py:308-347`)."""
    _, tracks, _, electrons = _replay(golden_events, name, keep_all_tb=True)
    for t, got in zip(tracks, electrons):
        assert np.array_equal(got, t["electrons"])


# default work split, and a stress split: each event over many deposit CTAs, each CTA's table appended in many segments
SPLITS = [None, dict(unit_points=40, table_spill_keys=120)]
SPLIT_IDS = ["default", "stress-split"]


@pytest.mark.parametrize("tuning", SPLITS, ids=SPLIT_IDS)
@pytest.mark.parametrize("name", case_names())
def test_dict_keys_charges_labels(golden_events, name, tuning):
    """The (pad, tb) -> (charge, label) map after all tracks (`transporter.py:252-317`)."""
    ev = golden_events
    _, _, batch, _ = _replay(ev, name, tuning=tuning, keep_all_tb=True)
    cloud, labels = batch.event(0)
    tb, pad = unpair(ev[f"{name}/keys"])
    order = np.lexsort((pad, tb))  # canonical order of the CUDA path: ascending (time bucket, pad)
    want_keys = ev[f"{name}/keys"][order]
    assert np.array_equal(cloud_keys(cloud), want_keys)
    want_charge = ev[f"{name}/charges"][order].astype(np.float64)
    assert np.allclose(cloud[:, 2], want_charge, rtol=CHARGE_RTOL, atol=0.0)
    assert np.array_equal(cloud[:, 2], want_charge), "charges are integers and come out identical"
    assert np.array_equal(labels, ev[f"{name}/key_labels"][order])
    assert np.array_equal(cloud[:, 1], np.floor(cloud[:, 1]) + ev[f"{name}/uniforms"][order])


@pytest.mark.parametrize("tuning", SPLITS, ids=SPLIT_IDS)
@pytest.mark.parametrize("name", case_names())
def test_simulate_cloud(golden_events, name, tuning):
    """Final `simulate` output (`simulator.py:104-115`), compared in canonical key order."""
    ev = golden_events
    _, _, batch, _ = _replay(ev, name, tuning=tuning)
    cloud, labels = batch.event(0)
    want_cloud, want_labels = sort_cloud(ev[f"{name}/cloud"], ev[f"{name}/labels"])
    assert cloud.shape == want_cloud.shape
    assert np.array_equal(cloud[:, 0], want_cloud[:, 0])  # pad ids
    assert np.array_equal(cloud[:, 1], want_cloud[:, 1])  # time buckets incl. replayed wiggle
    assert np.array_equal(cloud_keys(cloud), cloud_keys(want_cloud))
    assert np.allclose(cloud[:, 2], want_cloud[:, 2], rtol=CHARGE_RTOL, atol=0.0)
    assert np.array_equal(labels, want_labels)


@pytest.mark.parametrize("name", case_names())
def test_spyral_rows(golden_events, name):
    """`SpyralWriter.write` rows: response, ADC threshold, z-sort (`writer.py:61-112, 232-238`)."""
    ev = golden_events
    _, _, batch, _ = _replay(ev, name, spyral_rows=True)
    rows, labels = batch.event_rows(0)
    want, want_labels = ev[f"{name}/spyral_rows"], ev[f"{name}/spyral_labels"]
    assert rows.shape == want.shape
    if len(want) == 0:
        return
    assert np.all(np.diff(rows[:, 2]) >= 0)
    for col in (0, 1, 2, 3, 5, 6, 7):
        assert np.array_equal(rows[:, col], want[:, col]), f"column {col}"
    # the integral is a 512-term float sum in the reference; the device uses prefix sums
    assert np.allclose(rows[:, 4], want[:, 4], rtol=1e-12, atol=0.0)
    assert np.array_equal(labels, want_labels)


def test_pad_lookup(golden_misc):
    """`position_to_index` + grid read + beam veto for 68k positions incl. every edge case."""
    from attpc_engine_b200 import nuclear_map
    from attpc_engine_b200.detector.engine import engine_for
    from tests.common import make_config

    eng = engine_for(make_config(), [nuclear_map.get_data(1, 2)])
    got = eng.lookup_pads(golden_misc["pad_lookup/xy"])
    assert np.array_equal(got, golden_misc["pad_lookup/pad"].astype(np.int32))


def test_three_events_in_one_replay_call(golden_events):
    """Several events per call: per-event tables, labels and replayed uniforms stay separate."""
    from attpc_engine_b200.detector.engine import engine_for

    ev = golden_events
    names = ["dd_exit", "dd_stop", "dd_back"]  # same Config
    cfg = case_config(names[0])
    rows, normals, events, ranks, labels, zas, uniforms = [], [], [], [], [], [], []
    for e, name in enumerate(names):
        for t in case_tracks(ev, name):
            rows.append(t["rows"])
            normals.append(t["normals"])
            events.append(e)
            ranks.append(t["rank"])
            labels.append(t["idx"])
            zas.append(t["za"])
        uniforms.append((ev[f"{name}/keys"], ev[f"{name}/uniforms"]))
    eng = engine_for(cfg, nuclei_of([{"za": za} for za in zas]))
    batch, _ = eng.simulate_replay(rows, normals, events, ranks, labels, zas, len(names), uniforms=uniforms)
    for e, name in enumerate(names):
        cloud, lab = batch.event(e)
        want_cloud, want_labels = sort_cloud(ev[f"{name}/cloud"], ev[f"{name}/labels"])
        assert np.array_equal(cloud, want_cloud), name
        assert np.array_equal(lab, want_labels), name


# ------------------------------------------------------------- 32 reference events per bench workload, one call
def _workload_replay(name, tuning=None, **kw):
    from attpc_engine_b200.detector.engine import engine_for

    fx = load_workload(name)
    cfg = workload_config(name)
    tracks = workload_tracks(fx)
    n_events = len(fx["digests"])
    eng = engine_for(cfg, nuclei_of(tracks), **(tuning or {}))
    batch, electrons = eng.simulate_replay(
        [t["rows"] for t in tracks], [t["normals"] for t in tracks], [t["event"] for t in tracks],
        [t["rank"] for t in tracks], [t["idx"] for t in tracks], [t["za"] for t in tracks], n_events,
        uniforms=[workload_event_dict(fx, e)[::3] for e in range(n_events)], **kw,
    )  # fmt: skip
    return fx, cfg, tracks, batch, electrons


@pytest.mark.parametrize("tuning", SPLITS, ids=SPLIT_IDS)
@pytest.mark.parametrize("name", WORKLOAD_NAMES)
def test_workload_replay_dict(name, tuning):
    """Events drawn from bench.build_workload: electrons per row and the whole (pad, tb) -> (charge, label) map equal
    the reference's, for all 32 events of a workload in ONE replay call (`solver.py:308-347`, `transporter.py:252-317`)."""
    fx, _, tracks, batch, electrons = _workload_replay(name, tuning, keep_all_tb=True)
    for t, got in zip(tracks, electrons):
        assert np.array_equal(got, t["electrons"])
    for e in range(len(fx["digests"])):
        keys, charges, labels, uniforms = workload_event_dict(fx, e)
        tb, pad = unpair(keys)
        order = np.lexsort((pad, tb))
        cloud, lab = batch.event(e)
        assert np.array_equal(cloud_keys(cloud), keys[order]), (name, e)
        assert np.array_equal(cloud[:, 0], pad[order]) and np.array_equal(np.floor(cloud[:, 1]), tb[order])
        assert np.allclose(cloud[:, 2], charges[order].astype(np.float64), rtol=CHARGE_RTOL, atol=0.0)
        assert np.array_equal(cloud[:, 2], charges[order].astype(np.float64)), (name, e)
        assert np.array_equal(lab, labels[order]), (name, e)
        assert np.array_equal(cloud[:, 1], np.floor(cloud[:, 1]) + uniforms[order])


@pytest.mark.parametrize("name", WORKLOAD_NAMES)
def test_workload_replay_cloud_and_spyral(name):
    """Final `simulate` cloud (`simulator.py:104-115`) and `SpyralWriter.write` rows (`writer.py:61-112, 232-238`) of
    the same 32 events.  The expected arrays are rebuilt from the reference's recorded dict; the CPU test
    `test_workload_fixture_digests` pins that reconstruction to the digests of the reference's actual output."""
    from oracle import attpc_oracle as oracle

    fx, cfg, _, batch, _ = _workload_replay(name, spyral_rows=True)
    resp = oracle.get_response(cfg)
    for e, want in enumerate(fx["digests"]):
        want_cloud, want_labels = reference_cloud_from_dict(*workload_event_dict(fx, e))
        assert len(want_cloud) == want["n_cloud"]
        cloud, lab = batch.event(e)
        sorted_cloud, sorted_labels = sort_cloud(want_cloud, want_labels)
        assert np.array_equal(cloud, sorted_cloud) and np.array_equal(lab, sorted_labels), (name, e)
        rows, row_labels = batch.event_rows(e)
        assert len(rows) == want["n_spyral"], (name, e)
        if len(rows) == 0:
            continue
        want_rows, want_row_labels = oracle.spyral_event(want_cloud, want_labels, cfg, resp)
        assert np.all(np.diff(rows[:, 2]) >= 0)
        for col in (0, 1, 2, 3, 5, 6, 7):
            assert np.array_equal(rows[:, col], want_rows[:, col]), (name, e, col)
        assert np.allclose(rows[:, 4], want_rows[:, 4], rtol=1e-12, atol=0.0)
        assert np.array_equal(row_labels, want_row_labels)
