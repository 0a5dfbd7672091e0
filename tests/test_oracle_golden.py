"""Pin the CPU oracle (oracle/attpc_oracle.py) against fixtures recorded from the UNMODIFIED reference
(tests/golden/make_golden.py) and against the reference's own known-answer vectors."""

import numpy as np
import pytest

from attpc_engine_b200 import nuclear_map
from oracle import attpc_oracle as oracle
from tests.common import (
    WORKLOAD_NAMES, case_config, case_names, case_tracks, digest, load_workload, make_config, reference_cloud_from_dict,
    workload_config, workload_event_dict, workload_tracks,
)  # fmt: skip


def test_pairing_known_answers():
    """The reference's own KATs: tests/test_pairing.py:4-26."""
    assert oracle.szudzik_pair(56, 937) == 937**2 + 56
    assert oracle.szudzik_unpair(937**2 + 56) == (56, 937)
    assert oracle.szudzik_pair(937, 56) == 937**2 + 937 + 56
    assert oracle.szudzik_unpair(937**2 + 937 + 56) == (937, 56)


def test_pairing_golden(golden_misc):
    tb, pad, key = golden_misc["pairing/tb"], golden_misc["pairing/pad"], golden_misc["pairing/key"]
    for a, b, k, un in zip(tb, pad, key, golden_misc["pairing/unpaired"]):
        assert oracle.szudzik_pair(int(a), int(b)) == k
        assert oracle.szudzik_unpair(int(k)) == tuple(un)


def test_response_golden(golden_misc):
    assert np.array_equal(oracle.get_response(make_config()), golden_misc["response/default"])


@pytest.mark.parametrize("name", case_names())
def test_electrons_replay(golden_events, name):
    """`generate_electrons` (`solver.py:308-347`) with the recorded standard normals."""
    cfg = case_config(name)
    for t in case_tracks(golden_events, name):
        mass = nuclear_map.get_data(*t["za"]).mass
        got = oracle.fano_electrons(
            t["rows"], mass, cfg.det_params.w_value, cfg.det_params.fano_factor, normals=t["normals"]
        )
        assert np.array_equal(got, t["electrons"])


@pytest.mark.parametrize("name", case_names())
def test_simulate_replay(golden_events, name):
    """Whole `simulate` from recorded trajectories + random numbers: dict, cloud, labels identical."""
    ev = golden_events
    cfg = case_config(name)
    tracks = case_tracks(ev, name)
    rec = {}
    cloud, labels = oracle.simulate_event(
        ev[f"{name}/momenta"], ev[f"{name}/vertex"], ev[f"{name}/Z"], ev[f"{name}/A"], cfg, None,
        list(ev[f"{name}/indices"]), nuclear_map, record=rec, tracks=[t["rows"] for t in tracks],
        normals=[t["normals"] for t in tracks], uniforms=ev[f"{name}/uniforms"],
    )  # fmt: skip
    assert np.array_equal(rec["keys"], ev[f"{name}/keys"])  # insertion order included
    assert np.array_equal(rec["charges"], ev[f"{name}/charges"])
    assert np.array_equal(rec["key_labels"], ev[f"{name}/key_labels"])
    assert np.array_equal(cloud, ev[f"{name}/cloud"])
    assert np.array_equal(labels, ev[f"{name}/labels"])


@pytest.mark.parametrize("name", case_names())
def test_spyral_rows(golden_events, name):
    """Rows + threshold + z-sort of `SpyralWriter.write` (`writer.py:220-238`)."""
    ev = golden_events
    cfg = case_config(name)
    if len(ev[f"{name}/cloud"]) == 0:
        pytest.skip("empty cloud is never written")
    rows, labels = oracle.spyral_event(ev[f"{name}/cloud"], ev[f"{name}/labels"], cfg, oracle.get_response(cfg))
    assert np.array_equal(rows, ev[f"{name}/spyral_rows"])
    assert np.array_equal(labels, ev[f"{name}/spyral_labels"])


@pytest.mark.parametrize("name", ["dd_exit", "dp_decay", "alpha_breakup", "outside"])
def test_trajectory_default_radau(golden_events, name):
    """`generate_trajectory` (`solver.py:243-305`): same scipy Radau call -> same rows as the reference."""
    ev = golden_events
    cfg = case_config(name)
    for t in case_tracks(ev, name):
        nucleus = nuclear_map.get_data(*t["za"])
        track = oracle.integrate_track(ev[f"{name}/vertex"], ev[f"{name}/momenta"][t["idx"]], nucleus, cfg.det_params)
        n = len(t["rows"])
        assert len(track) >= n
        assert np.allclose(track[:n], t["rows"], rtol=1e-9, atol=1e-12)


def test_pad_lookup_matches_reference(golden_misc):
    """The oracle's grid index + veto against the reference chain for 68k positions."""
    cfg = make_config()
    xy = golden_misc["pad_lookup/xy"]
    got = np.empty(len(xy), dtype=np.int16)
    for i, (x, y) in enumerate(xy):
        ix, iy = oracle._grid_index(cfg.pad_grid_edges, x, y)
        pad = -1 if ix == -1 or iy == -1 else int(cfg.pad_grid[ix, iy])
        got[i] = -1 if (pad == -1 or pad in oracle.BEAM_PAD_IDS) else pad
    assert np.array_equal(got, golden_misc["pad_lookup/pad"])


@pytest.mark.parametrize("name", WORKLOAD_NAMES)
def test_workload_fixture_digests(name):
    """32 events per bench workload (tests/golden/make_workload_golden.py): the oracle, replaying the recorded
    trajectories / normals / uniforms, reproduces the reference's dict in insertion order, and the cloud and Spyral
    rows rebuilt from that dict hash to the digests of the reference's own `simulate` / `SpyralWriter.write` output."""
    fx = load_workload(name)
    cfg = workload_config(name)
    resp = oracle.get_response(cfg)
    tracks = workload_tracks(fx)
    indices = list(fx["indices"])
    for e, want in enumerate(fx["digests"]):
        mine = [t for t in tracks if t["event"] == e]
        keys, charges, labels, uniforms = workload_event_dict(fx, e)
        rec = {}
        cloud, lab = oracle.simulate_event(
            fx["momenta"][e], fx["vertices"][e], fx["Z"], fx["A"], cfg, None, indices, nuclear_map, record=rec,
            tracks=[t["rows"] for t in mine], normals=[t["normals"] for t in mine], uniforms=uniforms,
        )  # fmt: skip
        for t, got in zip(mine, rec["electrons"]):
            assert np.array_equal(got, t["electrons"])
        assert np.array_equal(rec["keys"], keys) and np.array_equal(rec["charges"], charges)
        assert np.array_equal(rec["key_labels"], labels)
        rebuilt, rebuilt_lab = reference_cloud_from_dict(keys, charges, labels, uniforms)
        assert np.array_equal(cloud, rebuilt) and np.array_equal(lab, rebuilt_lab)
        assert len(cloud) == want["n_cloud"] and digest(cloud, lab) == want["cloud"], (name, e)
        if len(cloud):
            rows, row_labels = oracle.spyral_event(cloud, lab, cfg, resp)
        else:
            rows, row_labels = np.zeros((0, 8)), np.zeros(0, np.int64)
        assert len(rows) == want["n_spyral"] and digest(rows, row_labels) == want["spyral"], (name, e)
