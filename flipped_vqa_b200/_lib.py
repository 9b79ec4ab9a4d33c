"""ctypes binding of the C ABI declared in include/fvqa.h (+ the test hooks of include/fvqa_debug.h), built in-tree.

The 16-bit tensor-core operand format is a property of the LIBRARY (tcgen05.mma rejects mixed fp16 x bf16 operands):
`libfvqa.so` = fp16 (default: the reference's dtype, `llama_vqa.py:63`, and the format that meets the full-depth parity
bound), `libfvqa_bf16.so` = bf16, selected with the environment variable FVQA_DTYPE=bf16 before the first import.
`H16` is the matching torch dtype of every `fvqa_h16` tensor.

There is NO CPU fallback: if the library is missing, does not export a symbol the header declares,
or `fvqa_init()` fails (no sm_100 GPU), the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
DTYPE_NAME = os.environ.get("FVQA_DTYPE", "fp16").lower()
if DTYPE_NAME not in ("fp16", "bf16"):
    raise ValueError(f"FVQA_DTYPE must be fp16 or bf16, got {DTYPE_NAME!r}")
H16 = torch.float16 if DTYPE_NAME == "fp16" else torch.bfloat16
LIB_PATH = os.path.join(_HERE, "libfvqa.so" if DTYPE_NAME == "fp16" else "libfvqa_bf16.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "fvqa.h")
DEBUG_HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "fvqa_debug.h")

_p, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64

# name -> argtypes (all functions return int unless listed in _RESTYPES)
_SIGNATURES = {
    "fvqa_abi_version": [],
    "fvqa_operand_dtype": [],
    "fvqa_last_error": [],
    "fvqa_init": [],
    "fvqa_rmsnorm_fwd": [_p, _p, _p, _p, _i, _i, _f, _p],
    "fvqa_rmsnorm_bwd": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    "fvqa_rmsnorm_gather_fwd": [_p, _p, _p, _p, _p, _i, _i, _f, _p],
    "fvqa_rmsnorm_scatter_bwd": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    "fvqa_swiglu_fwd": [_p, _p, _i, _i, _p],
    "fvqa_swiglu_bwd": [_p, _p, _p, _i, _i, _p],
    "fvqa_gemm_nt": [_p, _i, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _p],
    "fvqa_gemm_nt_rope": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _p, _p, _i, _i, _i, _p],
    "fvqa_gemm_nt_rope_pos": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _p, _p, _i, _i, _p, _p],
    "fvqa_gemm_skinny_grouped": [_p, _i64, _i, _p, _i, _p, _i64, _i, _i, _i, _i, _i, _i, _p],
    "fvqa_gemm_swiglu_fwd": [_p, _i, _p, _i, _p, _i, _p, _i, _i, _i, _i, _p],
    "fvqa_gemm_swiglu_bwd": [_p, _i, _p, _i, _p, _i, _p, _i, _i, _i, _i, _p],
    "fvqa_gemm_nn": [_p, _i, _p, _i, _p, _i, _i, _i, _i, _p],
    "fvqa_gemm_swiglu_bwd_nn": [_p, _i, _p, _i, _p, _i, _p, _i, _i, _i, _i, _p],
    "fvqa_gemm_debug_force_bn": [_i],
    "fvqa_gemm_debug_skinny_nt": [_i],
    "fvqa_gemm_debug_quad": [_i],
    "fvqa_gemm_quad_clusters": [],
    "fvqa_gemm_debug_epilogue_warps": [_i],
    "fvqa_gemm_debug_l2_hints": [_i],
    "fvqa_gemm_debug_mixed_a": [_i],
    "fvqa_attn_debug_use_tc": [_i],
    "fvqa_debug_pdl": [_i],
    "fvqa_attn_uses_tc": [_i, _i, _i],
    "fvqa_attn_fwd": [_p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "fvqa_attn_bwd_ws_bytes": [_i, _i, _i, _i, _i],
    "fvqa_attn_bwd": [_p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "fvqa_visual_proj_fwd": [_p, _p, _p, _i, _i, _i, _p],
    "fvqa_visual_proj_bwd": [_p, _p, _p, _i, _i, _i, _p],
    "fvqa_linear_f32": [_p, _p, _p, _p, _p, _i, _i, _i, _p],
    "fvqa_cross_attn_fwd": [_p, _p, _p, _p, _i, _i, _i, _i, _p],
    "fvqa_build_h0_fwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "fvqa_build_h0_bwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "fvqa_video_grad_finish": [_p, _p, _p, _i, _i, _i, _p],
    "fvqa_video_grad": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "fvqa_ce_fwd": [_p, _i, _p, _p, _p, _i, _i, _p],
    "fvqa_ce_bwd": [_p, _i, _p, _p, _p, _f, _p, _i, _i, _i, _p],
    "fvqa_sum_scale": [_p, _i, _f, _p, _p],
    "fvqa_qav_loss_fwd": [_p, _p, _p, _p, _f, _p, _p, _i, _i, _i, _p],
    "fvqa_qav_loss_bwd": [_p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _i, _i, _i, _i, _p],
    "fvqa_scatter_rows": [_p, _p, _p, _i, _p],
    "fvqa_greedy_next": [_p, _i, _i, _p, _i, _p, _i, _p, _p, _i, _i, _p, _p, _i, _p],
    "fvqa_option_score": [_p, _p, _p, _i, _i, _i, _p],
    "fvqa_f32_to_h16": [_p, _p, _i64, _p],
    "fvqa_grad_scale_prepare": [_p, _f, _p, _p, _p],
    "fvqa_scale_f32": [_p, _p, _i64, _p],
    "fvqa_gather_rows": [_p, _p, _p, _i, _i, _p],
    "fvqa_scatter_row_vectors": [_p, _p, _p, _i, _i, _p],
    "fvqa_expand_rows": [_p, _p, _p, _i, _i, _p],
}
_RESTYPES = {"fvqa_last_error": C.c_char_p, "fvqa_attn_bwd_ws_bytes": _i64}

_lib: Optional[C.CDLL] = None
_initialised = False


class FvqaError(RuntimeError):
    pass


def header_symbols(debug: bool = True) -> list:
    """Every function the public header (and, with `debug`, the test-hook header) declares (symbol-export test)."""
    text = ""
    for path in (HEADER_PATH,) + ((DEBUG_HEADER_PATH,) if debug else ()):
        with open(path) as f:
            text += f.read()
    return sorted(set(re.findall(r"\b(fvqa_[a-z0-9_]+)\s*\(", text)))


def load(build_if_missing: bool = False) -> C.CDLL:
    """dlopen libfvqa.so and declare signatures. Does not touch the GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _build
            _build.build()
        else:
            raise FvqaError(f"{LIB_PATH} not found: run `python -m flipped_vqa_b200.build` "
                            "(there is no CPU/PyTorch fallback for the training step)")
    lib = C.CDLL(LIB_PATH)
    for name, args in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise FvqaError(f"libfvqa.so does not export {name}") from e
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.fvqa_abi_version() != 2:
        raise FvqaError(f"{LIB_PATH}: ABI version mismatch; rebuild")
    if lib.fvqa_operand_dtype() != (0 if DTYPE_NAME == "fp16" else 1):
        raise FvqaError(f"{LIB_PATH} was not built for {DTYPE_NAME} operands; rebuild")
    _lib = lib
    return lib


def lib() -> C.CDLL:
    """Library handle with the GPU side initialised (raises without an sm_100 device)."""
    global _initialised
    l = load()
    if not _initialised:
        if not torch.cuda.is_available():
            raise FvqaError("flipped_vqa_b200 needs a CUDA (sm_100a) device; no CPU path exists")
        torch.cuda.init()
        torch.zeros(1, device="cuda")          # make sure the primary context is current
        rc = l.fvqa_init()
        if rc != 0:
            raise FvqaError(f"fvqa_init failed: {l.fvqa_last_error().decode()}")
        _initialised = True
    return l


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise FvqaError(f"{what} failed ({rc}): {load().fvqa_last_error().decode()}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def stream() -> int:
    """cudaStream_t of torch's CURRENT stream on the current device. Called once per kernel launch (~640 times per 7B step):
    `torch.cuda.current_stream()` costs ~14 us per call in Python-side device-index bookkeeping (7 ms of host time per step,
    tools/host_profile.py); the raw accessor underneath it is ~0.3 us."""
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream
