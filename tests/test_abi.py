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


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of every struct field as the C compiler lays out include/infinisst_b200.h, against the
    ctypes mirrors in _lib.py (field names must match too)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    structs = {"isst_config": _lib.IsstConfig, "isst_gen_params": _lib.IsstGenParams,
               "isst_beam_follow": _lib.IsstBeamFollow, "isst_beam_trace": _lib.IsstBeamTrace}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(src)], check=True)      # the header is plain C
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
    # every field the header declares is mirrored (count the declarators of isst_config)
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct isst_config \{(.*?)\} isst_config;", hdr, flags=re.S).group(1)
    names = re.findall(r"\b([a-z_0-9]+)(?:\[[A-Z_0-9]+\])?\s*[,;]", body)
    assert names == [f for f, _ in _lib.IsstConfig._fields_]


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
    # the CTA-pair GEMM: cta_group::2 MMAs, pair-wide TMA loads, commits multicast to both CTAs' barriers
    assert "UTCHMMA.2CTA" in sass and "UTMALDG.2D.2CTA" in sass and "UTCBAR.2CTA.MULTICAST" in sass


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
