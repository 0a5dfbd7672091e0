"""Minimal nuclear-data provider (no spyral_utils dependency).

The reference obtains masses through ``spyral_utils.nuclear.NuclearDataMap``
(`src/attpc_engine/__init__.py:1-3`, call sites `detector/simulator.py:99`,
`kinematics/reaction.py:54,217`).  spyral_utils is not available in this image,
so this module ships a small atomic-mass table (AME values, in u) that covers the
reactions named in BASELINE.json.  A nucleus that is NOT in the table raises ``KeyError``,
like the reference's full AME table does for a nucleus that does not exist: a
Bethe-Weizsaecker estimate is off by MeV and would silently corrupt Q-values and
4-momenta.  ``add_mass`` extends the table; ``NuclearDataMap(allow_liquid_drop=True)``
opts into the estimate (with a warning per nucleus) for exploratory work.
Objects are duck-type compatible with what the hot path reads: ``.mass``
(nuclear mass, MeV/c^2), ``.Z``, ``.A``, ``.isotopic_symbol``.

When spyral_utils IS importable, ``attpc_engine_b200.nuclear_map`` is its full-AME
``NuclearDataMap`` (`attpc_engine_b200/__init__.py`).  Either way the global can be replaced
at run time (``attpc_engine_b200.nuclear_map = my_map``): every consumer looks it up through
the package attribute when it is called and only uses ``get_data(z, a)``.
"""

from __future__ import annotations

import warnings
from dataclasses import dataclass

AMU_2_MEV = 931.49410242  # MeV/c^2 per u
ELECTRON_MASS = 0.51099895000  # MeV/c^2

_SYMBOLS = (
    "n H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn "
    "Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce "
    "Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn "
    "Fr Ra Ac Th Pa U"
).split()

# (Z, A) -> atomic mass [u]
_ATOMIC_MASS_U: dict[tuple[int, int], float] = {
    (0, 1): 1.00866491595,
    (1, 1): 1.00782503190,
    (1, 2): 2.01410177784,
    (1, 3): 3.01604928132,
    (2, 3): 3.01602932197,
    (2, 4): 4.00260325413,
    (2, 5): 5.012057,
    (2, 6): 6.018885889,
    (2, 8): 8.03393439,
    (3, 5): 5.0125378,
    (3, 6): 6.0151228874,
    (3, 7): 7.0160034366,
    (3, 8): 8.02248625,
    (3, 9): 9.02679019,
    (4, 7): 7.016928717,
    (4, 8): 8.005305102,
    (4, 9): 9.012183065,
    (4, 10): 10.013534695,
    (4, 11): 11.02166108,
    (4, 12): 12.0269221,
    (5, 9): 9.01332965,
    (5, 10): 10.01293695,
    (5, 11): 11.00930536,
    (5, 12): 12.0143526,
    (5, 13): 13.0177800,
    (6, 10): 10.01685331,
    (6, 11): 11.0114336,
    (6, 12): 12.0,
    (6, 13): 13.00335483507,
    (6, 14): 14.0032419884,
    (6, 15): 15.0105993,
    (6, 16): 16.014701,
    (6, 17): 17.022577,
    (7, 13): 13.00573861,
    (7, 14): 14.00307400443,
    (7, 15): 15.00010889888,
    (7, 16): 16.0061019,
    (8, 15): 15.0030656,
    (8, 16): 15.99491461957,
    (8, 17): 16.99913175650,
    (8, 18): 17.99915961286,
    (18, 40): 39.9623831237,
    (50, 131): 130.9170450,
    (50, 132): 131.9178267,
    (50, 133): 132.9239134,
}


def _liquid_drop_atomic_mass_u(z: int, a: int) -> float:
    """Bethe-Weizsaecker estimate; only used for nuclei missing from the table."""
    n = a - z
    av, as_, ac, aa, ap = 15.75, 17.8, 0.711, 23.7, 11.18
    binding = av * a - as_ * a ** (2.0 / 3.0) - ac * z * (z - 1) / a ** (1.0 / 3.0)
    binding -= aa * (a - 2 * z) ** 2 / a
    if a % 2 == 0:
        binding += ap / a**0.5 if (z % 2 == 0) else -ap / a**0.5
    m_h = _ATOMIC_MASS_U[(1, 1)] * AMU_2_MEV
    m_n = _ATOMIC_MASS_U[(0, 1)] * AMU_2_MEV
    return (z * m_h + n * m_n - binding) / AMU_2_MEV


@dataclass
class NucleusData:
    """Same fields the reference reads from spyral_utils' NucleusData."""

    mass: float = 0.0  # nuclear mass, MeV/c^2
    atomic_mass: float = 0.0  # u
    element_symbol: str = ""
    isotopic_symbol: str = ""
    pretty_iso_symbol: str = ""
    Z: int = 0
    A: int = 0

    def __str__(self) -> str:
        return self.isotopic_symbol


class NuclearDataMap:
    """``get_data(z, a) -> NucleusData`` from the packaged table (``KeyError`` for anything else)."""

    def __init__(self, allow_liquid_drop: bool = False) -> None:
        self._cache: dict[tuple[int, int], NucleusData] = {}
        self._extra: dict[tuple[int, int], float] = {}
        self.allow_liquid_drop = bool(allow_liquid_drop)

    def add_mass(self, z: int, a: int, atomic_mass_u: float) -> None:
        """Register (or override) the atomic mass [u] of a nucleus, e.g. from AME2020."""
        key = (int(z), int(a))
        self._extra[key] = float(atomic_mass_u)
        self._cache.pop(key, None)

    def get_data(self, z: int, a: int) -> NucleusData:
        key = (int(z), int(a))
        hit = self._cache.get(key)
        if hit is not None:
            return hit
        zz, aa = key
        if aa <= 0 or zz < 0 or zz > aa:
            raise KeyError(f"Nucleus Z={zz}, A={aa} does not exist")
        atomic = self._extra.get(key, _ATOMIC_MASS_U.get(key))
        if atomic is None:
            if not self.allow_liquid_drop:
                raise KeyError(
                    f"No tabulated mass for Z={zz}, A={aa}: the packaged table only covers the benchmark reactions. "
                    "Use NuclearDataMap.add_mass(z, a, atomic_mass_u), install spyral_utils (full AME table), or "
                    "construct NuclearDataMap(allow_liquid_drop=True) to accept a Bethe-Weizsaecker estimate (MeV-level error)."
                )
            warnings.warn(f"mass of Z={zz}, A={aa} is a liquid-drop ESTIMATE (MeV-level error)", stacklevel=2)
            atomic = _liquid_drop_atomic_mass_u(zz, aa)
        sym = _SYMBOLS[zz] if zz < len(_SYMBOLS) else f"Z{zz}"
        data = NucleusData(
            mass=atomic * AMU_2_MEV - zz * ELECTRON_MASS,
            atomic_mass=atomic,
            element_symbol=sym,
            isotopic_symbol=f"{aa}{sym}",
            pretty_iso_symbol=f"<sup>{aa}</sup>{sym}",
            Z=zz,
            A=aa,
        )
        self._cache[key] = data
        return data

    def has_tabulated_mass(self, z: int, a: int) -> bool:
        return (int(z), int(a)) in _ATOMIC_MASS_U or (int(z), int(a)) in self._extra
