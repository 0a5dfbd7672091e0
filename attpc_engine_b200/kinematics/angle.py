"""Polar-angle distributions in the centre-of-mass frame (reference: `kinematics/angle.py`)."""

from __future__ import annotations

from typing import Protocol

import numpy as np
from numpy.random import Generator


class PolarDistribution(Protocol):
    def sample(self, rng: Generator) -> float: ...


class PolarUniform:
    """Isotropic between two polar angles: uniform in cos(theta) (`angle.py:35-80`)."""

    def __init__(self, angle_min: float, angle_max: float):
        self.cos_angle_min = np.cos(angle_max)  # cos flips the order
        self.cos_angle_max = np.cos(angle_min)

    def sample(self, rng: Generator) -> float:
        return np.arccos(rng.uniform(self.cos_angle_min, self.cos_angle_max))

    def sample_n(self, rng: Generator, n: int) -> np.ndarray:
        return np.arccos(rng.uniform(self.cos_angle_min, self.cos_angle_max, size=n))


class PolarArbitrary:
    """Tabulated distribution: pick a bin by probability, smear uniformly inside it (`angle.py:83-152`)."""

    def __init__(self, angles: np.ndarray, probabilities: np.ndarray, angle_bin_width: float):
        if np.sum(probabilities) > 1.0:
            raise ValueError(
                f"The sum of the probabilities passed to PolarArbitrary should be 1.0. Yours sum to {np.sum(probabilities)}"
            )
        self.angle_width = angle_bin_width
        self.probs = probabilities
        self.angles = angles

    def sample(self, rng: Generator) -> float:
        return rng.choice(self.angles, p=self.probs) + rng.uniform(0.0, 1.0) * self.angle_width

    def sample_n(self, rng: Generator, n: int) -> np.ndarray:
        return rng.choice(self.angles, p=self.probs, size=n) + rng.uniform(0.0, 1.0, size=n) * self.angle_width
