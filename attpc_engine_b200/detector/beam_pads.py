"""Pads of the beam region, vetoed by the transport stage.

Reference: `detector/beam_pads.py:11-137` (a literal list of 122 ids scanned linearly per
pixel at `detector/transporter.py:165,237`).  Here the list is generated from its contiguous
runs and folded into the device lookup table once (`engine.build_pad_lut`), so the veto
costs nothing per pixel.
"""

import numpy as np

_RUNS = (
    (134, 164), (166, 166), (435, 457), (459, 459), (733, 733), (735, 735), (738, 738),
    (740, 741), (5254, 5284), (5286, 5286), (5555, 5577), (5579, 5579), (5853, 5853),
    (5855, 5855), (5858, 5858), (5860, 5861),
)  # fmt: skip

BEAM_PADS: list[int] = [p for lo, hi in _RUNS for p in range(lo, hi + 1)]
BEAM_PADS_ARRAY: np.ndarray = np.array(BEAM_PADS)
