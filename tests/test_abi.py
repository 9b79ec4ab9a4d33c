"""CPU: the C-ABI library builds, loads and exports every symbol include/fvqa.h declares.
No compute calls here (there is no GPU in this container)."""
import ctypes
import os

import pytest

from flipped_vqa_b200 import _lib, build


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def test_library_exports_every_header_symbol(lib_path):
    syms = _lib.header_symbols()
    assert len(syms) >= 25
    for variant in ("fp16", "bf16"):                 # both operand-format builds export the same ABI
        dll = ctypes.CDLL(build.lib_path(variant))
        missing = [s for s in syms if not hasattr(dll, s)]
        assert not missing, (variant, missing)
        assert dll.fvqa_operand_dtype() == {"fp16": 0, "bf16": 1}[variant]
    # and the ctypes binding declares a signature for each of them
    assert sorted(_lib._SIGNATURES) == syms
    # the product boundary (fvqa.h) holds no test / tuning hook: those live in fvqa_debug.h
    assert not [s for s in _lib.header_symbols(debug=False) if "debug" in s]


def test_abi_version_and_load(lib_path):
    l = _lib.load()
    assert l.fvqa_abi_version() == 2
    assert l.fvqa_operand_dtype() == (0 if _lib.DTYPE_NAME == "fp16" else 1)
    assert isinstance(l.fvqa_last_error(), bytes)
    assert l.fvqa_attn_bwd_ws_bytes(24, 128, 32, 128, 10) > 0


def test_sass_is_blackwell_native(lib_path):
    """tcgen05.mma / TMA / TMEM loads must be present in the sm_100a SASS (B200_PROFILING.md table)."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([exe, "-sass", lib_path], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.FvqaError):
        _lib.lib()
