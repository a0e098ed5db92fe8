"""B200-native implementation of the SEEKR2 OpenMM plugin's integrator step.

Only the per-timestep hot path of the reference (``MmvtLangevinIntegrator`` /
``ElberLangevinIntegrator``: Langevin update, milestone CV evaluation, crossing test, MMVT bounce,
crossing log) lives here, as hand-written sm_100a CUDA kernels behind the C ABI of
``include/seekr2_b200.h``.  Importing the package loads ``libseekr2_b200.so``; there is no CPU path.
"""
from ._lib import (  # noqa: F401
    Seekr2Error, LIB_PATH, SINGLE, MIXED, DOUBLE, VARIANT_CUDA, VARIANT_REFERENCE,
    KIND_LANGEVIN, KIND_MMVT, KIND_ELBER,
)
from .system import System, HarmonicBondForce, TetherForce, CustomCentroidBondForce  # noqa: F401
from .context import Context, State, BOLTZ  # noqa: F401
from .integrators import (  # noqa: F401
    MmvtLangevinIntegrator, ElberLangevinIntegrator, LangevinIntegrator,
    MmvtLangevinMiddleIntegrator, ElberLangevinMiddleIntegrator, LangevinMiddleIntegrator,
)

__all__ = [
    "Seekr2Error", "System", "HarmonicBondForce", "TetherForce", "CustomCentroidBondForce", "Context", "State",
    "MmvtLangevinIntegrator", "ElberLangevinIntegrator", "LangevinIntegrator", "BOLTZ",
    "MmvtLangevinMiddleIntegrator", "ElberLangevinMiddleIntegrator", "LangevinMiddleIntegrator",
]
