"""Import the UNMODIFIED reference (`/root/reference/src/attpc_engine`) in this container.

Only used by `tests/golden/make_golden.py` (and ad-hoc checks) on the build box -- never at
test/bench run time, because `/root/reference` does not exist on the GPU box.  The reference
needs spyral_utils (pycatima) and h5py, which are not installed; this registers minimal
stand-ins for those *third-party* modules so that the reference's own `solver.py`,
`transporter.py`, `pairing.py`, `simulator.simulate`, `response.py` and
`writer.convert_to_spyral` run as shipped.  Masses and dE/dx come from
`attpc_engine_b200.nuclear` / `attpc_engine_b200.target`, i.e. the same tables the CUDA path
and the oracle read.
"""

import sys
import types

REFERENCE_SRC = "/root/reference/src"


class MemDataset:
    """In-memory stand-in for h5py.Dataset (records what the reference's writer stores)."""

    def __init__(self, data):
        import numpy as np

        self.data = np.array(data)
        self.attrs = {}

    def __getitem__(self, item):
        return self.data[item]


class MemGroup(dict):
    def __init__(self):
        super().__init__()
        self.attrs = {}

    def create_group(self, name):
        self[name] = MemGroup()
        return self[name]

    def create_dataset(self, name, data=None):
        self[name] = MemDataset(data)
        return self[name]


class MemFile(MemGroup):
    opened = []  # every file the reference opened, in order

    def __init__(self, path, mode="r"):
        super().__init__()
        self.path, self.mode, self.closed = path, mode, False
        MemFile.opened.append(self)

    def close(self):
        self.closed = True


def install() -> types.ModuleType:
    if "attpc_engine" in sys.modules:
        return sys.modules["attpc_engine"]
    from attpc_engine_b200.nuclear import NuclearDataMap, NucleusData
    from attpc_engine_b200.target import AnalyticGasTarget, TableGasTarget

    class GasTarget(TableGasTarget):
        """spyral_utils-shaped constructor on top of the shared dE/dx tables."""

        def __init__(self, compound, pressure, nuclear_map=None):
            super().__init__(AnalyticGasTarget(compound, pressure))

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    su = mod("spyral_utils")
    su.__path__ = []
    nuc = mod("spyral_utils.nuclear", NucleusData=NucleusData, NuclearDataMap=NuclearDataMap)
    nuc.__path__ = []
    mod("spyral_utils.nuclear.target", GasTarget=GasTarget)
    mod("spyral_utils.nuclear.nuclear_map", NucleusData=NucleusData, NuclearDataMap=NuclearDataMap)
    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            mod("h5py", File=MemFile, Group=MemGroup, Dataset=MemDataset)
    if "vector" not in sys.modules:
        try:
            import vector  # noqa: F401
        except ImportError:
            pass
    sys.path.insert(0, REFERENCE_SRC)
    import attpc_engine.detector.simulator  # noqa: F401  (pulls solver, transporter, writer)

    return sys.modules["attpc_engine"]
