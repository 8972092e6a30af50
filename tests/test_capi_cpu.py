"""No GPU needed: the C-ABI library builds, loads, and exports every symbol the header declares;
argument validation that happens before any launch."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "probpose_b200.h")).read()
    return sorted(set(re.findall(r"PP_API\s+[\w\s\*]+?\b(pp_\w+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = _header_symbols()
    for s in ("pp_decode", "pp_gemm", "pp_engine_create", "pp_engine_infer", "pp_engine_load", "pp_last_error"):
        assert s in syms
    assert len(syms) >= 17


def test_library_exports_every_declared_symbol():
    from probpose_code_b200 import _lib, build

    build.build()
    lib = _lib.lib()
    missing = [s for s in _header_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == _header_symbols()
    assert lib.pp_version().decode().endswith("sm_100a")


def test_ctypes_structs_match_header_sizes():
    """The ctypes mirrors must have the C layout (sizes from the C compiler)."""
    import subprocess
    import tempfile

    from probpose_code_b200 import _lib

    prog = r'''
#include <stdio.h>
#include "probpose_b200.h"
int main(void) { printf("%zu %zu %zu %zu\n", sizeof(pp_decode_cfg), sizeof(pp_gemm_args), sizeof(pp_engine_cfg), sizeof(pp_profile)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(_lib.DecodeCfg), C.sizeof(_lib.GemmArgs), C.sizeof(_lib.EngineCfg), C.sizeof(_lib.Profile)]


def test_validation_without_a_device():
    from probpose_code_b200 import _lib

    lib = _lib.lib()
    # NULL cfg -> PP_ERR_INVALID and a message, no CUDA call needed
    assert lib.pp_decode(None, None, None, None, None, None, 1, None, None, None) == -1
    assert b"cfg" in lib.pp_last_error()
    cfg = _lib.DecodeCfg(18, 64, 48, 0, 0.5, 1.0, 0.0)
    assert lib.pp_decode(C.byref(cfg), 1, None, None, None, None, 1, 1, None, None) == -1  # K > 17
    cfg = _lib.DecodeCfg(17, 32, 32, 0, 0.5, 1.0, 0.0)
    assert lib.pp_decode(C.byref(cfg), 1, None, None, None, None, 1, 1, None, None) == -3  # unsupported map size
    cfg = _lib.DecodeCfg(17, 64, 48, 0, 0.5, 1.0, 0.0)
    assert lib.pp_decode(C.byref(cfg), None, None, None, None, None, 0, None, None, None) == 0  # empty batch is a no-op
    ecfg = _lib.EngineCfg(0, 4, 256, 192, 16, 2, 100, 12, 12, 1536, 17, 256, 1e-6, 1e-5, 0.5, 1.0,
                          (C.c_float * 3)(1, 1, 1), (C.c_float * 3)(1, 1, 1))
    assert lib.pp_engine_workspace_bytes(C.byref(ecfg)) == 0  # embed_dim 100 not built
    assert b"embed_dim" in lib.pp_last_error()
    ecfg.embed_dim = 384
    assert lib.pp_engine_workspace_bytes(C.byref(ecfg)) > 100 << 20
    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    with pytest.raises(_lib.PPError):
        _lib.check(-2, "x")


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no file of the product package may reference it."""
    pkg = os.path.join(ROOT, "probpose_code_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
