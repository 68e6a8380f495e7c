"""flashvtg_b200 - the FlashVTG inference hot path as hand-written sm_100a CUDA kernels behind a
C-ABI (include/flashvtg_b200.h), with a host mirror of the reference's operator interface.

Importing the package never builds or loads native code; the first compute call loads
libflashvtg_b200.so and fails loudly if it is missing (no CPU / eager fallback)."""
from .config import PRESETS, ModelConfig, postprocessor_preset  # noqa: F401

__all__ = ["PRESETS", "ModelConfig", "postprocessor_preset", "FlashVTGB200", "build_model_b200"]


def __getattr__(name):
    if name in ("FlashVTGB200", "build_model_b200", "FvtgResult"):
        from . import model
        return getattr(model, name)
    raise AttributeError(name)
