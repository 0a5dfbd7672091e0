"""Excitation-energy distributions (reference: `kinematics/excitation.py`).

Same classes and ``sample(rng)`` contract; each also has ``sample_n(rng, n)`` for the batched
pipeline.  All energies in MeV.
"""

from __future__ import annotations

from typing import Protocol

import numpy as np
from numpy.random import Generator


class ExcitationDistribution(Protocol):
    def sample(self, rng: Generator) -> float: ...


class ExcitationGaussian:
    """Normal distribution given by centroid and FWHM (sigma = FWHM / 2.355, `excitation.py:65`)."""

    def __init__(self, centroid: float = 0.0, width: float = 0.0):
        self.centroid = centroid
        self.width = width
        self.sigma = self.width / 2.355

    def sample(self, rng: Generator) -> float:
        return rng.normal(self.centroid, self.sigma)

    def sample_n(self, rng: Generator, n: int) -> np.ndarray:
        return rng.normal(self.centroid, self.sigma, size=n)


class ExcitationUniform:
    """Uniform on [min_value, max_value) (`excitation.py:83-128`)."""

    def __init__(self, min_value: float = 0.0, max_value: float = 0.0):
        self.min_value = min_value
        self.max_value = max_value

    def sample(self, rng: Generator) -> float:
        return rng.uniform(self.min_value, self.max_value)

    def sample_n(self, rng: Generator, n: int) -> np.ndarray:
        return rng.uniform(self.min_value, self.max_value, size=n)


class ExcitationBreitWigner:
    """Relativistic Breit-Wigner in total energy, returned as excitation (`excitation.py:131-188`)."""

    def __init__(self, rest_mass: float, centroid: float, width: float):
        self.rest_mass = rest_mass
        self.centroid = centroid
        self.width = width

    def _draw(self, rng: Generator, size):
        from scipy.stats import rel_breitwigner

        rho = (self.rest_mass + self.centroid) / self.width
        return rel_breitwigner.rvs(rho, scale=self.width, size=size, random_state=rng) - self.rest_mass

    def sample(self, rng: Generator) -> float:
        return float(self._draw(rng, None))

    def sample_n(self, rng: Generator, n: int) -> np.ndarray:
        return np.asarray(self._draw(rng, n), dtype=np.float64)
