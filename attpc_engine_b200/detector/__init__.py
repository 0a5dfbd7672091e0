"""Detector effects (mirror of `attpc_engine.detector`, reference `detector/__init__.py:3-21`)."""

from .parameters import Config, DetectorParams, ElectronicsParams, PadParams
from .simulator import SimEvent, run_simulation, simulate, simulate_batch, simulate_stream
from .writer import (
    ArrayWriter,
    ParquetCloudWriter,
    SimulationWriter,
    SpyralWriter,
    convert_to_spyral,
    read_parquet_clouds,
)

__all__ = [
    "run_simulation",
    "simulate",
    "simulate_batch",
    "simulate_stream",
    "SimEvent",
    "DetectorParams",
    "ElectronicsParams",
    "PadParams",
    "Config",
    "SpyralWriter",
    "ArrayWriter",
    "ParquetCloudWriter",
    "read_parquet_clouds",
    "SimulationWriter",
    "convert_to_spyral",
]
