"""gasoline_b200 -- B200-native (sm_100a) tree-gravity force evaluation behind Gasoline's pkd call surface.

Only what the hot path needs lives here: csrc/ (CUDA kernels + the C ABI of include/gasoline_b200.h),
pkd.py (host-side mirror of the reference's pkd interface for this path, ctypes over the C ABI),
ics.py (synthetic Tipsy workloads), domain.py (multi-GPU domain split + tree exchange)."""
from .pkd import PKD, GravityParams, GasolineB200Error, load_library  # noqa: F401

__all__ = ["PKD", "GravityParams", "GasolineB200Error", "load_library"]
