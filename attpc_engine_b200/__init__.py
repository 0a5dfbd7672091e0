"""attpc_engine_b200 -- B200-native detector-simulation hot path of the AT-TPC engine.

Mirrors the reference's top-level surface (`src/attpc_engine/__init__.py:1-3`): a global
``nuclear_map`` with ``get_data(z, a)``.  Unlike the reference, importing this package does
not require spyral_utils.
"""

from .nuclear import NuclearDataMap, NucleusData

__version__ = "0.1.0"

nuclear_map = NuclearDataMap()

__all__ = ["nuclear_map", "NuclearDataMap", "NucleusData", "__version__"]
