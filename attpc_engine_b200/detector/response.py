"""GET electronics response (reference: `detector/response.py:8-32`).

``get_response`` is evaluated once per ``Config`` on the host and uploaded to the GPU; the
per-point amplitude / integral of `response.py:35-57` (``apply_response``) is computed by the
CUDA finalize stage (`csrc/attpc_kernels.cuh: shaped`).
"""

import numpy as np

from .constants import E_CHARGE, NUM_TB
from .parameters import Config


def get_response(config: Config) -> np.ndarray:
    """Theoretical GET shaper response sampled at ``NUM_TB`` points (negative lobes clipped to 0).

    r(t) = 4095 e / (gain fC) * exp(-3 t/tau) (t/tau)^3 sin(t/tau), tau = shaping_time * clock * 1e-3,
    on ``linspace(0, NUM_TB, NUM_TB)`` exactly as the reference samples it.
    """
    scale = 4095 * E_CHARGE / config.elec_params.amp_gain / 1e-15
    t = np.linspace(0.0, NUM_TB, NUM_TB)
    x = t / (config.elec_params.shaping_time * config.elec_params.clock_freq * 0.001)
    shape = scale * np.exp(-3.0 * x) * (x**3) * np.sin(x)
    shape[shape < 0] = 0
    return shape
