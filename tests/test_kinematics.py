"""Kinematics front end: the reference's own assertions (tests/test_kinematics.py) re-expressed with
pytest idioms, plus conservation checks of the batched implementation."""

import numpy as np
import pytest

from attpc_engine_b200 import nuclear_map
from attpc_engine_b200.kinematics import (
    Decay,
    ExcitationGaussian,
    ExcitationUniform,
    KinematicsPipeline,
    KinematicsTargetMaterial,
    PipelineError,
    PolarUniform,
    Reaction,
    run_kinematics_pipeline,
)
from attpc_engine_b200.kinematics.reaction import invariant_mass
from tests.common import gas


def nuc(z, a):
    return nuclear_map.get_data(z, a)


def b10_he3_chain():
    return [
        Reaction(target=nuc(5, 10), projectile=nuc(2, 3), ejectile=nuc(2, 4)),
        Decay(parent=nuc(5, 9), residual_1=nuc(2, 4)),
        Decay(parent=nuc(3, 5), residual_1=nuc(2, 4)),
    ]


def test_reaction_lise_value():
    """12C(d,p) at 16 MeV, 20 deg in the c.m.: ejectile KE = 18.391 MeV (reference tests/test_kinematics.py:13-36)."""
    rxn = Reaction(nuc(6, 12), nuc(1, 2), nuc(1, 1))
    result = rxn.calculate(16.0, np.deg2rad(20.0), 0.0, residual_excitation=0.0)
    assert np.round(result[2].E - result[2].M, decimals=3) == 18.391
    total = sum(v.as_array() for v in result[2:])
    assert np.allclose(total, result[0].as_array() + result[1].as_array(), atol=1e-9)


def test_pipeline_runs_and_reports_nuclei():
    pipeline = KinematicsPipeline(
        b10_he3_chain(),
        [ExcitationGaussian(16.8, 0.2), ExcitationGaussian(0.0, 1.25), ExcitationGaussian(0.0, 0.0)],
        [PolarUniform(0.0, np.pi)] * 3,
        24.0,
    )
    vertex, result = pipeline.run()
    assert np.all(pipeline.get_proton_numbers() == np.array([5, 2, 2, 5, 2, 3, 2, 1]))
    assert np.all(pipeline.get_mass_numbers() == np.array([10, 3, 4, 9, 4, 5, 4, 1]))
    assert len(result) == 8
    assert np.all(vertex == 0.0)


@pytest.mark.parametrize(
    "steps, n_ex, n_pol",
    [
        (lambda: b10_he3_chain()[:2], 1, 2),  # excitations shorter than steps
        (lambda: b10_he3_chain()[:2], 2, 1),  # polar distributions shorter than steps
        (lambda: [b10_he3_chain()[0], Decay(parent=nuc(4, 8), residual_1=nuc(2, 4))], 2, 2),  # broken chain
        (lambda: [b10_he3_chain()[1], b10_he3_chain()[0]], 2, 2),  # decay before reaction
        (lambda: [], 0, 0),
    ],
)
def test_pipeline_validation(steps, n_ex, n_pol):
    with pytest.raises(PipelineError):
        KinematicsPipeline(steps(), [ExcitationGaussian(16.8, 0.2)] * n_ex, [PolarUniform(0.0, np.pi)] * n_pol, 24.0)


def test_pipeline_sample_limit():
    """An excitation that is never energetically allowed stops at the sample limit."""
    pipeline = KinematicsPipeline(
        [b10_he3_chain()[0]], [ExcitationGaussian(16.8, 0.2)], [PolarUniform(0.0, np.pi)], 2.0, event_sample_limit=20
    )
    with pytest.raises(PipelineError):
        pipeline.run()


def test_batch_conserves_four_momentum():
    pipeline = KinematicsPipeline(
        b10_he3_chain(),
        [ExcitationGaussian(16.8, 0.2), ExcitationGaussian(0.0, 1.25), ExcitationGaussian(0.0, 0.0)],
        [PolarUniform(0.0, np.pi)] * 3,
        24.0,
    ).seed(5)
    vertices, p = pipeline.run_batch(2000)
    assert p.shape == (2000, 8, 4) and vertices.shape == (2000, 3)
    assert np.allclose(p[:, 0] + p[:, 1], p[:, 2] + p[:, 3], atol=1e-8)
    assert np.allclose(p[:, 3], p[:, 4] + p[:, 5], atol=1e-8)
    assert np.allclose(p[:, 5], p[:, 6] + p[:, 7], atol=1e-8)
    for col, (z, a) in zip((2, 4, 6, 7), ((2, 4), (2, 4), (2, 4), (1, 1))):
        assert np.allclose(invariant_mass(p[:, col]), nuc(z, a).mass, rtol=1e-9)


def test_target_material_sampling_and_npz_roundtrip(tmp_path):
    target = KinematicsTargetMaterial(material=gas("D2_600"), z_range=(0.0, 1.0), rho_sigma=0.007)
    pipeline = KinematicsPipeline(
        [Reaction(target=nuc(1, 2), projectile=nuc(6, 16), ejectile=nuc(1, 2))],
        [ExcitationUniform(0.0, 2.0)], [PolarUniform(0.0, np.pi)], 184.131, target_material=target,
    ).seed(3)  # fmt: skip
    vertices, p = pipeline.run_batch(500)
    assert np.all((vertices[:, 2] >= 0.0) & (vertices[:, 2] < 1.0))
    beam_ke = p[:, 1, 3] - nuc(6, 16).mass
    assert np.all(beam_ke < 184.131) and np.all(beam_ke > 100.0)
    order = np.argsort(vertices[:, 2])
    assert np.all(np.diff(beam_ke[order]) <= 1e-9)  # deeper vertex, more energy lost
    out = tmp_path / "kin.npz"
    run_kinematics_pipeline(pipeline.seed(3), 50, out, verbose=False)
    with np.load(out) as f:
        assert f["data"].shape == (50, 4, 4) and f["vertices"].shape == (50, 3)
        assert list(f["proton_numbers"]) == [1, 6, 1, 6] and list(f["mass_numbers"]) == [2, 16, 2, 16]
