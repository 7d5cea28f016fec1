"""esr_b200: Blackwell-native x4 efficient-SR inference (IMDN / RFDN / RLFN / BSRN forward) behind the
reference's `select_model` / `forward` API.  Compute lives in csrc/ (sm_100a CUDA, C ABI in
include/esr_b200.h); this package is the host-side mirror of the reference interface."""
from .engine import Engine, EsrError  # noqa: F401
from .demo_api import B200SRModel, build_model, forward, forward_uint8, select_model  # noqa: F401

__all__ = ["Engine", "EsrError", "B200SRModel", "build_model", "select_model", "forward", "forward_uint8"]
