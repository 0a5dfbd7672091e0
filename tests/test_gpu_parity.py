"""GPU parity, part (a): identical trajectory points + replayed random numbers -> bit-exact pad ids,
time buckets and pairing keys, charges within 1e-6 relative (they come out equal).

Every case compares the CUDA path, called through the C ABI, with fixtures recorded from the
UNMODIFIED reference (tests/golden/make_golden.py)."""

import numpy as np
import pytest

from attpc_engine_b200.detector.pairing import unpair
from tests.common import case_config, case_names, case_tracks, cloud_keys, nuclei_of, sort_cloud

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


@pytest.mark.parametrize("tuning", [None, dict(unit_points=40, table_spill_keys=120)], ids=["default", "stress-split"])
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


# default work split, and a stress split: each event over many deposit CTAs, each CTA's table appended in many segments
SPLITS = [None, dict(unit_points=40, table_spill_keys=120)]


@pytest.mark.parametrize("tuning", SPLITS, ids=["default", "stress-split"])
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
