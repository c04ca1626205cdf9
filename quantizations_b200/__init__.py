"""quantizations_b200 -- Blackwell-native (sm_100a) 4-bit weight-only Linear engine.

Drop-in for the hot path of kkbwilldo/quantizations: `core` mirrors the reference's core.py, `modules` its modules.py;
both call hand-written CUDA through the C ABI declared in include/quantizations_b200.h.  No CPU fallback.
"""
from . import _lib, core, modules  # noqa: F401
from .core import (  # noqa: F401
    Params4bit,
    QuantState,
    create_dynamic_map,
    dequantize_4bit,
    dequantize_blockwise,
    gemm_4bit,
    gemv_4bit_fused,
    gemv_4bit_chain,
    gemv_4bit_batch,
    decode_attention,
    gemv_4bit,
    get_4bit_type,
    quantize_4bit,
    quantize_blockwise,
)
from .modules import Linear4bit, Linear4bitGroup, matmul_4bit  # noqa: F401

__version__ = "0.1.0"
