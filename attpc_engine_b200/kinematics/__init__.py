"""Kinematics phase space (mirror of `attpc_engine.kinematics`, reference `kinematics/__init__.py:3-33`)."""

from .angle import PolarArbitrary, PolarDistribution, PolarUniform
from .excitation import ExcitationBreitWigner, ExcitationDistribution, ExcitationGaussian, ExcitationUniform
from .pipeline import (
    KinematicsPipeline,
    KinematicsTargetMaterial,
    PipelineError,
    run_kinematics_pipeline,
    save_kinematics_npz,
)
from .reaction import Decay, Reaction

__all__ = [
    "KinematicsPipeline",
    "run_kinematics_pipeline",
    "save_kinematics_npz",
    "KinematicsTargetMaterial",
    "PipelineError",
    "ExcitationDistribution",
    "ExcitationGaussian",
    "ExcitationUniform",
    "ExcitationBreitWigner",
    "PolarDistribution",
    "PolarArbitrary",
    "PolarUniform",
    "Reaction",
    "Decay",
]
