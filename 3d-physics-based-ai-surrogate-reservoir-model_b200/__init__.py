"""B200-native physics-loss path of the 3-D physics-based AI surrogate reservoir model.

Only the hot path is here: PVT splines, well sources, the finite-difference
mass-balance residual and its hand-written adjoint, as sm_100a CUDA kernels behind a C ABI
(``include/srm_physics.h``, ``libsrm_physics.so``), plus the host-side mirror of the reference's
``PhysicsLoss`` / ``PVTLayer`` / ``WellRatesPressure`` / ``HardLayer`` call contracts.

The directory name is not a Python identifier; import it with
``importlib.import_module("3d-physics-based-ai-surrogate-reservoir-model_b200")`` or through the
``srm_b200`` shim at the repository root.
"""
from . import _lib, config, pvt, synth  # noqa: F401
from .config import PhysicsSpec, spec_from_reference_configs  # noqa: F401
from .pvt import PVTLayer, build_polynomial_tables, build_spline_tables, load_default_pvt_table  # noqa: F401

__all__ = ["_lib", "config", "pvt", "synth", "PhysicsSpec", "spec_from_reference_configs", "PVTLayer",
           "build_spline_tables", "build_polynomial_tables", "load_default_pvt_table"]


def __getattr__(name):
    # engine / physics_loss import torch.cuda-facing code lazily
    import importlib
    if name in ("engine", "physics_loss", "wells", "dist", "hard_layer", "batching", "data"):
        return importlib.import_module(f"{__name__}.{name}")
    if name == "SrmPhysics":
        return importlib.import_module(f"{__name__}.engine").SrmPhysics
    if name in ("PhysicsLoss",):
        return getattr(importlib.import_module(f"{__name__}.physics_loss"), name)
    if name == "BatchGenerator":
        return importlib.import_module(f"{__name__}.batching").BatchGenerator
    if name in ("HardLayer", "CompleteTrainableModule"):
        return getattr(importlib.import_module(f"{__name__}.hard_layer"), name)
    if name in ("WellRatesPressure", "WellDataProcessor"):
        return getattr(importlib.import_module(f"{__name__}.wells"), name)
    raise AttributeError(name)
