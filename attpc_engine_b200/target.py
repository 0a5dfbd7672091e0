"""Gas targets and stopping-power tables.

The reference evaluates ``target.get_dedx(ejectile, KE)`` inside every right-hand-side
call of the equation of motion (`detector/solver.py:64-66`), where ``target`` is a
``spyral_utils.nuclear.target.GasTarget`` backed by pycatima 1.96 (CATIMA, C++), neither
of which is vendored in the reference nor installed here.  This module provides

* :class:`DedxTable` -- the table format the CUDA integrator stages in shared memory: a
  *pseudo-logarithmic* grid whose nodes are the doubles ``2^e * (1 + m/M)``, so the cell
  index comes straight from the exponent/mantissa bits of KE (no ``log``) and the value is
  a linear interpolation inside the cell.  Both the device code and the CPU oracle evaluate
  exactly this piecewise-linear function, so trajectory parity never depends on CATIMA.
* :class:`TableGasTarget` -- duck-types the two members the hot path uses (``get_dedx``,
  ``density``) on top of tables sampled from *any* source target (real spyral_utils
  ``GasTarget`` included) at first use.
* :class:`AnalyticGasTarget` -- a self-contained Bethe + Lindhard stopping model, clearly
  **not CATIMA**, so tests and benchmarks run without spyral_utils.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .nuclear import AMU_2_MEV, NucleusData

GAS_CONSTANT = 62363.0  # cm^3 Torr / (K mol)
ROOM_TEMPERATURE = 293.15  # K


# --------------------------------------------------------------------------------------
# Table format
# --------------------------------------------------------------------------------------
@dataclass
class DedxTable:
    """dE/dx [MeV/(g/cm^2)] of one ion species on a pseudo-log KE grid.

    Node ``i`` sits at ``x_i = 2^(e_min + i // M) * (1 + (i % M) / M)`` MeV, ``M = 2**lm``.
    ``values`` has ``n_oct * M + 1`` entries.  Evaluation (same on host and device):

    * ``KE <  x_0``    -> ``values[0] * sqrt(KE / x_0)``   (velocity-proportional stopping)
    * ``KE >= x_last`` -> ``values[-1]``
    * else            -> ``values[i] + f * (values[i+1] - values[i])`` with ``i, f`` read off
      the binary representation of KE (``f`` is the cell fraction, exact in floating point).
    """

    z: int
    a: int
    mass: float  # nuclear mass MeV/c^2 the table was built for
    lm: int
    e_min: int
    n_oct: int
    values: np.ndarray = field(repr=False)

    @property
    def nodes_per_octave(self) -> int:
        return 1 << self.lm

    @property
    def ke_min(self) -> float:
        return math.ldexp(1.0, self.e_min)

    @property
    def ke_max(self) -> float:
        return math.ldexp(1.0, self.e_min + self.n_oct)

    def nodes(self) -> np.ndarray:
        return table_nodes(self.lm, self.e_min, self.n_oct)

    def __call__(self, ke: float) -> float:
        """Scalar evaluation (the oracle calls this once per RHS evaluation)."""
        if not ke >= self.ke_min:
            if ke <= 0.0 or ke != ke:
                return 0.0
            return float(self.values[0]) * math.sqrt(ke / self.ke_min)
        mant, exp = math.frexp(ke)  # ke = mant * 2^exp, mant in [0.5, 1)
        octave = exp - 1 - self.e_min
        if octave >= self.n_oct:
            return float(self.values[-1])
        pos = (mant * 2.0 - 1.0) * (1 << self.lm)
        m = int(pos)
        i = (octave << self.lm) + m
        lo = self.values[i]
        return float(lo + (pos - m) * (self.values[i + 1] - lo))

    def evaluate(self, ke: np.ndarray) -> np.ndarray:
        """Vectorised evaluation (identical arithmetic to ``__call__``)."""
        ke = np.asarray(ke, dtype=np.float64)
        out = np.empty_like(ke)
        low = ~(ke >= self.ke_min)
        with np.errstate(invalid="ignore", divide="ignore"):
            out[low] = np.where(
                ke[low] > 0.0, self.values[0] * np.sqrt(ke[low] / self.ke_min), 0.0
            )
        mant, exp = np.frexp(ke[~low])
        octave = exp - 1 - self.e_min
        hi = octave >= self.n_oct
        pos = (mant * 2.0 - 1.0) * (1 << self.lm)
        m = pos.astype(np.int64)
        i = np.where(hi, 0, (octave.astype(np.int64) << self.lm) + m)
        lo = self.values[i]
        val = lo + (pos - m) * (self.values[i + 1] - lo)
        out[~low] = np.where(hi, self.values[-1], val)
        return out


def table_nodes(lm: int, e_min: int, n_oct: int) -> np.ndarray:
    m = 1 << lm
    i = np.arange(n_oct * m + 1)
    return np.ldexp(1.0 + (i % m) / m, e_min + i // m)


DEFAULT_LM = 6  # 64 nodes per octave
DEFAULT_E_MIN = -24  # 2^-24 MeV ~ 0.06 eV
DEFAULT_N_OCT = 36  # up to 2^12 = 4096 MeV


def build_dedx_table(
    source,
    nucleus: NucleusData,
    lm: int = DEFAULT_LM,
    e_min: int = DEFAULT_E_MIN,
    n_oct: int = DEFAULT_N_OCT,
) -> DedxTable:
    """Sample ``source.get_dedx(nucleus, ke)`` on the pseudo-log grid."""
    xs = table_nodes(lm, e_min, n_oct)
    if hasattr(source, "get_dedx_array"):
        vals = np.asarray(source.get_dedx_array(nucleus, xs), dtype=np.float64)
    else:
        vals = np.array([source.get_dedx(nucleus, float(x)) for x in xs], dtype=np.float64)
    if not np.all(np.isfinite(vals)) or np.any(vals < 0.0):
        raise ValueError(f"get_dedx returned non-finite/negative values for {nucleus}")
    return DedxTable(
        z=int(nucleus.Z),
        a=int(nucleus.A),
        mass=float(nucleus.mass),
        lm=lm,
        e_min=e_min,
        n_oct=n_oct,
        values=vals,
    )


def interpolation_error(source, nucleus: NucleusData, table: DedxTable, per_cell: int = 4) -> float:
    """Max relative error of the table against its source at cell-interior probes."""
    xs = table.nodes()
    worst = 0.0
    for k in range(1, per_cell):
        probe = xs[:-1] + (xs[1:] - xs[:-1]) * (k / per_cell)
        if hasattr(source, "get_dedx_array"):
            truth = np.asarray(source.get_dedx_array(nucleus, probe))
        else:
            truth = np.array([source.get_dedx(nucleus, float(x)) for x in probe])
        approx = table.evaluate(probe)
        ok = truth > 0
        worst = max(worst, float(np.max(np.abs(approx[ok] - truth[ok]) / truth[ok])))
    return worst


# --------------------------------------------------------------------------------------
# Table-backed target (what Config hands to the CUDA engine and to the oracle)
# --------------------------------------------------------------------------------------
class TableGasTarget:
    """``get_dedx`` / ``density`` facade over per-species :class:`DedxTable` objects."""

    def __init__(
        self,
        source=None,
        density: float | None = None,
        tables: dict[tuple[int, int], DedxTable] | None = None,
        lm: int = DEFAULT_LM,
        e_min: int = DEFAULT_E_MIN,
        n_oct: int = DEFAULT_N_OCT,
    ) -> None:
        if source is None and density is None:
            raise ValueError("TableGasTarget needs a source target or an explicit density")
        self.source = source
        self.density = float(source.density if density is None else density)  # g/cm^3
        self.tables: dict[tuple[int, int], DedxTable] = dict(tables or {})
        self.lm, self.e_min, self.n_oct = lm, e_min, n_oct
        for name in ("compound", "pressure", "molar_mass", "ugly_string", "pretty_string"):
            if source is not None and hasattr(source, name):
                setattr(self, name, getattr(source, name))

    def table_for(self, nucleus: NucleusData) -> DedxTable:
        key = (int(nucleus.Z), int(nucleus.A))
        tab = self.tables.get(key)
        if tab is None:
            if self.source is None:
                raise KeyError(f"No dE/dx table for Z={key[0]} A={key[1]} and no source target")
            tab = build_dedx_table(self.source, nucleus, self.lm, self.e_min, self.n_oct)
            self.tables[key] = tab
        return tab

    def get_dedx(self, projectile_data: NucleusData, projectile_energy: float) -> float:
        return self.table_for(projectile_data)(projectile_energy)

    def get_dedx_array(self, projectile_data: NucleusData, energies: np.ndarray) -> np.ndarray:
        return self.table_for(projectile_data).evaluate(energies)

    # -- range / energy-loss helpers (used by the kinematics front end only) ------------
    def _range_table(self, nucleus: NucleusData) -> tuple[np.ndarray, np.ndarray]:
        key = ("range", int(nucleus.Z), int(nucleus.A))
        cached = getattr(self, "_range_cache", {}).get(key)
        if cached is not None:
            return cached
        tab = self.table_for(nucleus)
        xs = tab.nodes()
        inv = 1.0 / (tab.values * self.density * 100.0)  # m / MeV
        seg = 0.5 * (inv[1:] + inv[:-1]) * np.diff(xs)
        # below x_0 stopping ~ sqrt(E): integral_0^x0 dE/(s0 sqrt(E/x0)) = 2 x0 / s0
        r0 = 2.0 * xs[0] * inv[0]
        rng = np.concatenate(([r0], r0 + np.cumsum(seg)))
        if not hasattr(self, "_range_cache"):
            self._range_cache = {}
        self._range_cache[key] = (xs, rng)
        return xs, rng

    def get_range(self, projectile_data: NucleusData, projectile_energy: float) -> float:
        xs, rng = self._range_table(projectile_data)
        return float(np.interp(projectile_energy, xs, rng))

    def get_energy_loss(
        self, projectile_data: NucleusData, projectile_energy: float, distances: np.ndarray
    ) -> np.ndarray:
        """Energy lost [MeV] after travelling each of ``distances`` [m]."""
        xs, rng = self._range_table(projectile_data)
        r_start = np.interp(projectile_energy, xs, rng)
        remaining = np.clip(r_start - np.asarray(distances, dtype=np.float64), 0.0, None)
        e_end = np.interp(remaining, rng, xs, left=0.0)
        return projectile_energy - e_end


# --------------------------------------------------------------------------------------
# Analytic stand-in for CATIMA
# --------------------------------------------------------------------------------------
_MEAN_EXCITATION_EV = {1: 19.2, 2: 41.8, 6: 81.0, 7: 82.0, 8: 95.0, 9: 115.0, 10: 137.0, 18: 188.0}
_K_BETHE = 0.307075  # MeV cm^2 / mol
_ME_C2_EV = 510998.95
_ALPHA = 1.0 / 137.035999084
_AVOGADRO = 6.02214076e23
_LINDHARD_EV_CM2 = 1.9154e-14  # 8 pi e^2 a0 in eV cm^2


class AnalyticGasTarget:
    """Bethe (effective charge) + Lindhard-Scharff electronic stopping.  NOT CATIMA.

    Same constructor shape as ``spyral_utils.nuclear.target.GasTarget``:
    ``compound = [(Z, A, S), ...]``, ``pressure`` in Torr.  ``density`` follows the same
    ideal-gas expression spyral_utils uses (sum(A*S) * P / (R * 293.15 K)).
    """

    def __init__(self, compound: list[tuple[int, int, int]], pressure: float, nuclear_map=None):
        self.compound = [(int(z), int(a), int(s)) for z, a, s in compound]
        self.pressure = float(pressure)
        self.molar_mass = float(sum(a * s for _, a, s in self.compound))
        self.density = self.molar_mass * self.pressure / (GAS_CONSTANT * ROOM_TEMPERATURE)
        self.ugly_string = "".join(f"{a}Z{z}x{s}" for z, a, s in self.compound)
        self.pretty_string = self.ugly_string

    def get_dedx_array(self, projectile_data: NucleusData, energies: np.ndarray) -> np.ndarray:
        ke = np.asarray(energies, dtype=np.float64)
        z1 = float(projectile_data.Z)
        mass = float(projectile_data.mass)
        gamma = 1.0 + ke / mass
        beta2 = 1.0 - 1.0 / (gamma * gamma)
        beta2 = np.maximum(beta2, 1e-300)
        beta = np.sqrt(beta2)
        total = np.zeros_like(ke)
        for z2, a2, s in self.compound:
            weight = a2 * s / self.molar_mass  # mass fraction
            i_ev = _MEAN_EXCITATION_EV.get(z2, 16.0 * z2**0.9)
            zeff = z1 * (1.0 - np.exp(-125.0 * beta / z1 ** (2.0 / 3.0))) if z1 > 0 else 0.0
            arg = 2.0 * _ME_C2_EV * beta2 * gamma * gamma / i_ev
            high = _K_BETHE * z2 / a2 * zeff * zeff / beta2 * np.maximum(np.log1p(arg) - beta2, 1e-12)
            low_atom = (
                _LINDHARD_EV_CM2
                * z1 ** (7.0 / 6.0)
                * z2
                / (z1 ** (2.0 / 3.0) + z2 ** (2.0 / 3.0)) ** 0.75
                * (beta / _ALPHA)
            )  # eV cm^2 / atom
            low = low_atom * 1e-6 * _AVOGADRO / a2  # MeV cm^2 / g
            total += weight * (low * high) / (low + high)
        return total

    def get_dedx(self, projectile_data: NucleusData, projectile_energy: float) -> float:
        return float(self.get_dedx_array(projectile_data, np.array([projectile_energy]))[0])

    def as_table_target(self, **kw) -> TableGasTarget:
        return TableGasTarget(self, **kw)

    def get_range(self, projectile_data: NucleusData, projectile_energy: float) -> float:
        return self.as_table_target().get_range(projectile_data, projectile_energy)

    def get_energy_loss(self, projectile_data, projectile_energy, distances):
        if not hasattr(self, "_tt"):
            self._tt = self.as_table_target()
        return self._tt.get_energy_loss(projectile_data, projectile_energy, distances)


# Name kept so user scripts written against spyral_utils keep working when it is absent.
GasTarget = AnalyticGasTarget


# --------------------------------------------------------------------------------------
# Table files: tabulate once where pycatima is installed, run anywhere
# --------------------------------------------------------------------------------------
def save_table_target(path, target: TableGasTarget) -> None:
    """Write the tables of ``target`` (every species tabulated so far) and its density to one ``.npz`` file."""
    payload = {"density": np.float64(target.density), "grid": np.array([target.lm, target.e_min, target.n_oct])}
    for (z, a), tab in sorted(target.tables.items()):
        if (tab.lm, tab.e_min, tab.n_oct) != (target.lm, target.e_min, target.n_oct):
            raise ValueError("all tables of a file share one grid")
        payload[f"dedx_{z}_{a}"] = tab.values
        payload[f"mass_{z}_{a}"] = np.float64(tab.mass)
    np.savez(path, **payload)


def load_table_target(path) -> TableGasTarget:
    """A `TableGasTarget` without a source: exactly the species stored in the file (`save_table_target`)."""
    with np.load(path) as f:
        lm, e_min, n_oct = (int(v) for v in f["grid"])
        tables = {}
        for name in f.files:
            if name.startswith("dedx_"):
                z, a = (int(v) for v in name.split("_")[1:])
                tables[(z, a)] = DedxTable(z=z, a=a, mass=float(f[f"mass_{z}_{a}"]), lm=lm, e_min=e_min, n_oct=n_oct,
                                           values=np.array(f[name], dtype=np.float64))  # fmt: skip
        return TableGasTarget(density=float(f["density"]), tables=tables, lm=lm, e_min=e_min, n_oct=n_oct)


def tabulate_gas_target(gas_target, nuclei, path=None, max_error: float | None = None) -> TableGasTarget:
    """Tabulate ``gas_target.get_dedx`` (e.g. a real ``spyral_utils`` `GasTarget` backed by pycatima) for ``nuclei``,
    check the interpolation error against ``max_error`` (relative; None = report only, see ``.errors``) and
    optionally write the file.  The returned target is what `DetectorParams.gas_target` takes."""
    target = TableGasTarget(gas_target)
    target.errors = {}
    for nucleus in nuclei:
        table = target.table_for(nucleus)
        err = interpolation_error(gas_target, nucleus, table)
        target.errors[(int(nucleus.Z), int(nucleus.A))] = err
        if max_error is not None and err > max_error:
            raise ValueError(f"dE/dx table of Z={nucleus.Z} A={nucleus.A}: interpolation error {err:.2e} > {max_error:.2e}")
    if path is not None:
        save_table_target(path, target)
    return target


def ensure_table_target(gas_target) -> TableGasTarget:
    """Wrap any duck-typed gas target (``get_dedx``, ``density``) into a table target."""
    if isinstance(gas_target, TableGasTarget):
        return gas_target
    cached = getattr(gas_target, "_attpc_b200_table_target", None)
    if cached is None:
        cached = TableGasTarget(gas_target)
        try:
            gas_target._attpc_b200_table_target = cached
        except Exception:
            pass
    return cached


__all__ = [
    "AMU_2_MEV",
    "DedxTable",
    "TableGasTarget",
    "AnalyticGasTarget",
    "GasTarget",
    "build_dedx_table",
    "interpolation_error",
    "ensure_table_target",
    "save_table_target",
    "load_table_target",
    "tabulate_gas_target",
    "table_nodes",
]
