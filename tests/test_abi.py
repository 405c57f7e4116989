"""The C-ABI boundary without a GPU: the library builds/loads, exports every symbol
include/infinisst_b200.h declares, the ctypes structs match the C layout, and the product
path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest
import torch

from infinisst_b200 import _lib, build, tiny_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "infinisst_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(isst_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes binding"
    assert sorted(_lib.SYMBOLS) == declared


def test_struct_layout_matches_header():
    # isst_config: 4 + 3*8 + 4 + 2 + 1 + 3*8 + 7 ints + float + 7 ints, all 4-byte fields
    n_fields = 1 + 3 * 8 + 4 + 2 + 1 + 3 * 8 + 7 + 1 + 7
    assert C.sizeof(_lib.IsstConfig) == 4 * n_fields
    # isst_gen_params: int, int, float, int, int[8], int, (pad), pointer, int, (pad)
    assert C.sizeof(_lib.IsstGenParams) == 4 * 12 + 4 + 4 + 8 + 8 or C.sizeof(_lib.IsstGenParams) == 72


def test_library_is_sm100a_tcgen05_tma():
    """SASS evidence that the shipped library carries the Blackwell-native GEMM (UTC*MMA, TMA)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from infinisst_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(tiny_config())
    lib = _lib.load()
    cfg = _lib.IsstConfig()
    h = C.c_void_p()
    assert lib.isst_create(C.byref(cfg), 0, C.byref(h)) != 0
    assert b"no CPU fallback" in lib.isst_last_error() or b"CUDA" in lib.isst_last_error()
