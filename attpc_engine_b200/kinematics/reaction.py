"""Relativistic two-body reaction / decay kinematics (reference: `kinematics/reaction.py`).

The reference builds `vector` 4-vector objects event by event; here the algebra is written out
on arrays (``*_batch`` methods, ``[n, 4]`` rows of px, py, pz, E in MeV) and the scalar methods
of the reference's API are thin wrappers that return :class:`FourVector` records.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

import attpc_engine_b200 as _pkg


@dataclass
class FourVector:
    """Minimal stand-in for ``vector.MomentumObject4D``: px, py, pz, E (+ invariant mass ``M``)."""

    px: float
    py: float
    pz: float
    E: float

    @property
    def M(self) -> float:
        return float(np.sqrt(max(self.E**2 - self.px**2 - self.py**2 - self.pz**2, 0.0)))

    def as_array(self) -> np.ndarray:
        return np.array([self.px, self.py, self.pz, self.E])

    @classmethod
    def from_array(cls, row) -> "FourVector":
        return cls(float(row[0]), float(row[1]), float(row[2]), float(row[3]))


def invariant_mass(p4: np.ndarray) -> np.ndarray:
    return np.sqrt(np.maximum(p4[..., 3] ** 2 - np.sum(p4[..., :3] ** 2, axis=-1), 0.0))


def boost_from_rest_frame(p4_cm: np.ndarray, frame: np.ndarray) -> np.ndarray:
    """Boost rows given in the rest frame of ``frame`` (4-momenta, rows) into the lab."""
    mass = invariant_mass(frame)
    beta = frame[..., :3] / frame[..., 3:4]
    gamma = frame[..., 3] / mass
    bp = np.sum(beta * p4_cm[..., :3], axis=-1)
    out = np.empty_like(p4_cm)
    coeff = gamma * (gamma / (gamma + 1.0) * bp + p4_cm[..., 3])
    out[..., :3] = p4_cm[..., :3] + beta * coeff[..., None]
    out[..., 3] = gamma * (p4_cm[..., 3] + bp)
    return out


def _two_body(parent: np.ndarray, m_light: float, m_heavy: np.ndarray, theta: np.ndarray, phi: np.ndarray):
    """Split ``parent`` into (light, heavy); the light one goes to (theta, phi) in the parent frame."""
    e_cm = invariant_mass(parent)
    e1 = (m_light**2 - m_heavy**2 + e_cm**2) / (2.0 * e_cm)
    p1 = np.sqrt(np.maximum(e1**2 - m_light**2, 0.0))
    cm = np.stack(
        [p1 * np.sin(theta) * np.cos(phi), p1 * np.sin(theta) * np.sin(phi), p1 * np.cos(theta), e1], axis=-1
    )
    light = boost_from_rest_frame(cm, parent)
    return light, parent - light


class Reaction:
    """target(projectile, ejectile)residual; the residual is deduced (`reaction.py:8-68`)."""

    def __init__(self, target, projectile, ejectile):
        self.projectile = projectile
        self.target = target
        self.ejectile = ejectile
        resid_z = projectile.Z + target.Z - ejectile.Z
        resid_a = projectile.A + target.A - ejectile.A
        if resid_z < 0:
            raise ValueError("Reaction calculated a residual Z (proton number) < 0, illegal reaction!")
        if resid_a < 0:
            raise ValueError("Reaction calculated a residual A (mass number) < 0, illegal reaction!")
        self.residual = _pkg.nuclear_map.get_data(resid_z, resid_a)
        self.reaction_symbol = f"{self.target}({self.projectile},{self.ejectile}){self.residual}"

    def __str__(self) -> str:
        return self.reaction_symbol

    def _entrance(self, projectile_energy):
        t = np.asarray(projectile_energy, dtype=np.float64)
        pz = np.sqrt(t * (t + 2.0 * self.projectile.mass))
        zero = np.zeros_like(pz)
        proj = np.stack([zero, zero, pz, t + self.projectile.mass], axis=-1)
        targ = np.stack([zero, zero, zero, zero + self.target.mass], axis=-1)
        return targ, proj

    def is_excitation_allowed(self, projectile_energy, residual_excitation):
        """Enough centre-of-mass energy for the exit channel? (`reaction.py:70-101`); works on arrays."""
        targ, proj = self._entrance(projectile_energy)
        e_cm = invariant_mass(targ + proj)
        ok = self.ejectile.mass + self.residual.mass + np.asarray(residual_excitation) < e_cm
        return bool(ok) if np.ndim(ok) == 0 else ok

    def threshold(self, residual_excitation):
        q = self.target.mass + self.projectile.mass - (self.ejectile.mass + self.residual.mass + residual_excitation)
        m_out = self.ejectile.mass + self.residual.mass
        return -q * m_out / (m_out - self.projectile.mass)

    def calculate_batch(self, projectile_energy, ejectile_polar, ejectile_azimuthal, residual_excitation) -> np.ndarray:
        """``[n, 4, 4]``: target, projectile, ejectile, residual 4-momenta in the lab."""
        targ, proj = self._entrance(projectile_energy)
        eject, resid = _two_body(
            targ + proj, self.ejectile.mass, self.residual.mass + np.asarray(residual_excitation, dtype=np.float64),
            np.asarray(ejectile_polar, dtype=np.float64), np.asarray(ejectile_azimuthal, dtype=np.float64),
        )  # fmt: skip
        return np.stack([targ, proj, eject, resid], axis=-2)

    def calculate(self, projectile_energy, ejectile_polar, ejectile_azimuthal, residual_excitation) -> list[FourVector]:
        """Scalar API of the reference (`reaction.py:103-178`), same ValueError below threshold."""
        if projectile_energy < self.threshold(residual_excitation):
            raise ValueError("Beam energy below kinematic threshold!")
        rows = self.calculate_batch(
            np.array([projectile_energy]), np.array([ejectile_polar]), np.array([ejectile_azimuthal]),
            np.array([residual_excitation]),
        )[0]  # fmt: skip
        return [FourVector.from_array(r) for r in rows]


class Decay:
    """parent -> residual_1 + residual_2; residual_2 is deduced (`reaction.py:181-228`)."""

    def __init__(self, parent, residual_1):
        self.parent = parent
        self.residual_1 = residual_1
        resid_2_z = parent.Z - residual_1.Z
        resid_2_a = parent.A - residual_1.A
        if resid_2_z < 0:
            raise ValueError("Decay calculated a residual2 Z (proton number) < 0, illegal decay!")
        if resid_2_a < 0:
            raise ValueError("Decay calculated a residual2 A (mass number) < 0, illegal decay!")
        self.residual_2 = _pkg.nuclear_map.get_data(resid_2_z, resid_2_a)
        self.decay_symbol = f"{self.parent}->{self.residual_1}+{self.residual_2}"

    def __str__(self) -> str:
        return self.decay_symbol

    @staticmethod
    def _rows(parent_vector) -> np.ndarray:
        if isinstance(parent_vector, FourVector):
            return parent_vector.as_array()
        return np.asarray(parent_vector, dtype=np.float64)

    def is_excitation_allowed(self, parent_vector, residual_2_excitation):
        """Parent invariant mass above the exit-channel masses? (`reaction.py:230-250`)."""
        q = invariant_mass(self._rows(parent_vector)) - (
            self.residual_1.mass + self.residual_2.mass + np.asarray(residual_2_excitation)
        )
        ok = q > 0.0
        return bool(ok) if np.ndim(ok) == 0 else ok

    def calculate_batch(self, parent_rows, residual_1_polar, residual_1_azimuthal, residual_2_excitation):
        """``(residual_1 [n, 4], residual_2 [n, 4])`` in the lab."""
        return _two_body(
            np.asarray(parent_rows, dtype=np.float64), self.residual_1.mass,
            self.residual_2.mass + np.asarray(residual_2_excitation, dtype=np.float64),
            np.asarray(residual_1_polar, dtype=np.float64), np.asarray(residual_1_azimuthal, dtype=np.float64),
        )  # fmt: skip

    def calculate(self, parent_vector, residual_1_polar, residual_1_azimuthal, residual_2_excitation):
        """Scalar API of the reference (`reaction.py:252-303`)."""
        rows = self._rows(parent_vector)
        q = invariant_mass(rows) - (self.residual_1.mass + self.residual_2.mass + residual_2_excitation)
        if q < 0.0:
            raise ValueError("Parent doesn't have enough energy to decay!")
        r1, r2 = self.calculate_batch(
            rows[None], np.array([residual_1_polar]), np.array([residual_1_azimuthal]),
            np.array([residual_2_excitation]),
        )  # fmt: skip
        return [FourVector.from_array(rows), FourVector.from_array(r1[0]), FourVector.from_array(r2[0])]
