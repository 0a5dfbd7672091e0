"""Host-side decode of the packed wire columns (`SimBatch.packed`) into the reference's arrays.

`simulate` of the reference returns ``cloud float64 [N, 3]`` (pad, time bucket, electrons) and ``labels int64 [N]``
(`simulator.py:104-115`).  The GPU ships 8 B/row + 1 KB/event instead of 32 B/row; this module rebuilds the arrays in one
pass over the rows, one host thread per group of events (numba, the reference's own JIT dependency).  Without numba the
numpy route of `SimBatch` does the same, an order of magnitude slower.  Format conversion only: nothing of the
simulation runs here.
"""

from __future__ import annotations

import numpy as np

try:  # numba is a dependency of the reference itself (`detector/solver.py`, `transporter.py`), not of this package
    import numba

    @numba.njit(parallel=True, cache=True)
    def _decode(pad_rank, wiggle, electrons_u32, tb_counts, offsets, labels_of_rank, shift, cloud, labels):  # pragma: no cover
        mask = (1 << shift) - 1
        n_tb = tb_counts.shape[1]
        for e in numba.prange(tb_counts.shape[0]):
            r = offsets[e]
            for tb in range(n_tb):
                c = tb_counts[e, tb]
                for _ in range(c):
                    pr = int(pad_rank[r])
                    cloud[r, 0] = float(pr & mask)
                    cloud[r, 1] = float(tb) + float(wiggle[r]) * (1.0 / 65536.0)  # exact: 9 + 16 bits
                    cloud[r, 2] = float(electrons_u32[r])
                    labels[r] = labels_of_rank[pr >> shift]
                    r += 1

    HAVE_NUMBA = True
except Exception:  # noqa: BLE001 - any import / compile problem: the numpy route stays
    HAVE_NUMBA = False


def decode_packed(packed: dict, offsets: np.ndarray):
    """``(cloud [N, 3] float64, labels [N] int64)`` from the packed columns, or None when numba is not usable or the
    electrons travel as int64 (heavy ions: the numpy route handles that rare case)."""
    if not HAVE_NUMBA or "electrons_u32" not in packed:
        return None
    n = len(packed["pad_rank"])
    cloud = np.empty((n, 3), dtype=np.float64)
    labels = np.empty(n, dtype=np.int64)
    if n:
        _decode(packed["pad_rank"], packed["wiggle"], packed["electrons_u32"], packed["tb_counts"],
                np.ascontiguousarray(offsets, dtype=np.int64), packed["labels_of_rank"].astype(np.int64),
                int(packed["rank_shift"]), cloud, labels)  # fmt: skip
        rows = packed["big_rows"]
        if len(rows):  # the few counts >= 2^32
            cloud[rows, 2] = packed["big_electrons"]
    return cloud, labels
