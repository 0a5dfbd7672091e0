"""attpc_engine_b200 -- B200-native detector-simulation hot path of the AT-TPC engine.

Mirrors the reference's top-level surface (`src/attpc_engine/__init__.py:1-3`): a global
``nuclear_map`` with ``get_data(z, a)``.  Unlike the reference, importing this package does
not require spyral_utils; when spyral_utils is importable its full-AME map is the global.
The global may be replaced at run time: consumers read ``attpc_engine_b200.nuclear_map`` when called.
"""

from .nuclear import NuclearDataMap, NucleusData

__version__ = "0.2.0"

try:  # the reference's own provider (`src/attpc_engine/__init__.py:1-3`)
    from spyral_utils.nuclear import NuclearDataMap as _SpyralNuclearDataMap

    nuclear_map = _SpyralNuclearDataMap()
except Exception:  # not installed (this image): the packaged table, KeyError outside it
    nuclear_map = NuclearDataMap()

__all__ = ["nuclear_map", "NuclearDataMap", "NucleusData", "__version__"]
