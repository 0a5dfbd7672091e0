"""Physical constants of the detector path (reference: `detector/constants.py:23-35`).

Values are the CODATA-2018 exact SI definitions the reference pulls from scipy.constants;
they are spelled out so that the device code, the oracle and the host agree to the bit.
"""

NUM_TB: int = 512  # GET time buckets per trace

E_CHARGE: float = 1.602176634e-19  # C (exact)
C: float = 299792458.0  # m/s (exact)
MEV_2_JOULE: float = E_CHARGE * 1.0e6  # J / MeV
MEV_2_KG: float = (E_CHARGE / (C * C)) * 1.0e6  # kg per MeV/c^2

# Integration grid of the reference (`detector/solver.py:14-16`)
KE_LIMIT: float = 1.0e-6  # MeV
TIME_STEP: float = 1.0e-10  # s
MAX_TIME_STEPS: int = 10001  # grid points 0 .. 1 us

# Hard-coded detector bounds of the reference's terminal events (`solver.py:160,200,240`)
Z_BOUND_HI: float = 1.0  # m
Z_BOUND_LO: float = 0.0  # m
RHO_BOUND: float = 0.292  # m

DIFFUSION_MESH_STEPS: int = 10  # `detector/transporter.py:8`
NUM_PADS: int = 10240
