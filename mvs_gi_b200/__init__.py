"""mvs_gi_b200 — B200-native (sm_100a) spherical sweep / cost-volume / soft-argmin path of castacks/mvs_gi.

Drop-in modules (same interfaces as the reference's dsta_mvs/model components):
    SphericalSweepStdMasked, SphericalSweep, DistanceRegressorWithFixedCandidates, patch_model
Host-side calibration for the fused in-kernel projection:
    calibration.Rig, EquirectCamera, DoubleSphereCamera, LinearSphereCamera
The CUDA kernels live in csrc/ behind the C ABI declared in include/mvsgi_b200.h.
"""
from . import abi, calibration  # noqa: F401  (torch-free)

__all__ = ["abi", "calibration", "SphericalSweepStdMasked", "SphericalSweep",
           "DistanceRegressorWithFixedCandidates", "patch_model"]


def __getattr__(name):  # lazy: importing torch can take a minute on a cold box
    if name in ("SphericalSweepStdMasked", "SphericalSweep", "DistanceRegressorWithFixedCandidates", "patch_model"):
        from . import modules

        return getattr(modules, name)
    if name in ("ops", "modules", "synthetic", "sharding"):
        import importlib

        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
