"""A kinematics file written by the REFERENCE's own `run_kinematics_pipeline`, as a fixture for the HDF5 reader.

    python tests/golden/make_kinematics_file.py      # writes tests/golden/kinematics_file.npz

h5py is not installed in this image, so the reference's unmodified function (`kinematics/pipeline.py:429-495`) writes
into the in-memory stand-in of tests/golden/ref_shim.py; the resulting tree (groups, datasets, attributes) is flattened
into an .npz.  `tests/test_host_logic.py::test_hdf5_kinematics_reader_on_a_reference_written_file` rebuilds the tree and
reads it back through `attpc_engine_b200.detector.simulator._Hdf5Kinematics`.  The sampling itself (the `vector`
package is not installed either) is replaced by a stub pipeline that returns prepared events: the layout of the file
is what is under test, not the physics.  CHUNK_SIZE is lowered so that 23 events span four chunk groups.
"""

import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))

import ref_shim  # noqa: E402

if "vector" not in sys.modules:
    try:
        import vector  # noqa: F401
    except ImportError:
        class _AnyAttr(types.ModuleType):  # imported by kinematics/reaction.py for annotations; never called here
            def __getattr__(self, name):
                return object

        sys.modules["vector"] = _AnyAttr("vector")
ref_shim.install()
import attpc_engine.kinematics.pipeline as ref_pipeline  # noqa: E402

N_EVENTS, K, CHUNK = 23, 6, 7


class StubPipeline:
    def __init__(self):
        rng = np.random.default_rng(77)
        self.vertices = rng.uniform(-0.01, 1.0, size=(N_EVENTS, 3))
        self.momenta = rng.normal(0.0, 500.0, size=(N_EVENTS, K, 4))
        self.at = 0

    def run(self):
        i = self.at
        self.at += 1
        return self.vertices[i], self.momenta[i]

    def get_proton_numbers(self):
        return np.array([1, 6, 1, 6, 6, 0], dtype=np.int64)

    def get_mass_numbers(self):
        return np.array([2, 14, 1, 15, 14, 1], dtype=np.int64)

    def __str__(self):
        return "stub"


def flatten(group, prefix, out):
    for k, v in group.attrs.items():
        out[f"attr|{prefix}|{k}"] = np.asarray(v)
    for name, child in group.items():
        path = f"{prefix}/{name}"
        if isinstance(child, ref_shim.MemDataset):
            out[f"data|{path}"] = child.data
            for k, v in child.attrs.items():
                out[f"attr|{path}|{k}"] = np.asarray(v)
        else:
            out[f"group|{path}"] = np.zeros(0)
            flatten(child, path, out)


def main():
    ref_pipeline.CHUNK_SIZE = CHUNK
    ref_pipeline.h5 = types.SimpleNamespace(File=ref_shim.MemFile)
    ref_shim.MemFile.opened.clear()
    stub = StubPipeline()
    ref_pipeline.run_kinematics_pipeline(stub, N_EVENTS, Path("/nonexistent/kinematics.h5"))
    tree = ref_shim.MemFile.opened[-1]
    out = {}
    flatten(tree, "", out)
    out["expected|momenta"] = stub.momenta
    out["expected|vertices"] = stub.vertices
    np.savez_compressed(HERE / "kinematics_file.npz", **out)
    print(sorted(k for k in out if not k.startswith("data|"))[:12], len(out))


if __name__ == "__main__":
    main()
