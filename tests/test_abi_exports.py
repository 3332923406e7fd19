"""CPU: the C-ABI shared library builds (nvcc cross-compiles sm_100a without a GPU), loads, and exports every
symbol include/kvquant_b200.h declares; the Python surface has exactly the reference's 34 operator names."""
import ctypes
import os
import re
import pytest

from _util import ROOT

REFERENCE_OPS = """
vecquant4matmul_nuq_perchannel_transposed_mha_batched_fused_opt vecquant4matmul_nuq_perchannel_transposed_mha_batched_fused_opt2
vecquant4matmul_nuq_perchannel_transposed_rope_mha_batched_fused_opt vecquant4matmul_nuq_perchannel_transposed_rope_mha_batched_fused_opt2
vecquant4appendvecK vecquant4appendvecKsparse vecquant4appendvecKsparseParallel
vecquant4appendvecV vecquant4appendvecVsparse vecquant4appendvecVsparseParallel
vecquant3matmul_nuq_perchannel_transposed_mha_batched_fused_opt vecquant3matmul_nuq_perchannel_transposed_mha_batched_fused_opt2
vecquant3matmul_nuq_perchannel_transposed_rope_mha_batched_fused_opt vecquant3matmul_nuq_perchannel_transposed_rope_mha_batched_fused_opt2
vecquant3appendvecK vecquant3appendvecKsparse vecquant3appendvecKsparseParallel
vecquant3appendvecV vecquant3appendvecVsparse vecquant3appendvecVsparseParallel
vecquant2matmul_nuq_perchannel_transposed_mha_batched_fused_opt vecquant2matmul_nuq_perchannel_transposed_mha_batched_fused_opt2
vecquant2matmul_nuq_perchannel_transposed_rope_mha_batched_fused_opt vecquant2matmul_nuq_perchannel_transposed_rope_mha_batched_fused_opt2
vecquant2appendvecK vecquant2appendvecKsparse vecquant2appendvecKsparseParallel
vecquant2appendvecV vecquant2appendvecVsparse vecquant2appendvecVsparseParallel
vecquant4matmul_nuq_perchannel_transposed_rope_mha_batched_fused_opt2_orig vecquant4appendvecKsparseorig
vecquant4matmul_nuq_perchannel_transposed_mha_batched_fused_opt2_orig vecquant4appendvecVsparseorig
""".split()  # the m.def list of reference deployment/kvquant/quant_cuda.cpp:401-436


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "kvquant_b200.h")).read()
    return sorted(set(re.findall(r"KVQ_API\s+[\w\s\*]+?\b(kvq_\w+)\s*\(", hdr)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from kvquant_b200 import build, _lib
    path = build.build(verbose=False)
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert sorted(_lib.SIGNATURES) == declared          # the ctypes table covers the header one to one
    assert _lib.load().kvq_abi_version() == 1
    assert b"bits" in _lib.load().kvq_error_string(-1)


def test_python_surface_has_exactly_the_reference_ops():
    import quant_cuda  # top-level drop-in shim
    assert len(REFERENCE_OPS) == 34
    assert sorted(quant_cuda.OP_NAMES) == sorted(REFERENCE_OPS)
    for n in REFERENCE_OPS:
        assert callable(getattr(quant_cuda, n))


def test_argument_validation_is_loud_not_a_fallback():
    import pytest
    import torch
    import quant_cuda
    mat = torch.zeros((32, 16, 64), dtype=torch.int32)
    lut = torch.zeros((32, 128, 16))
    with pytest.raises(ValueError):  # CPU tensors are rejected, never computed on the host
        quant_cuda.vecquant4appendvecK(mat, lut, torch.zeros(4096), 0)
    with pytest.raises(ValueError):  # wrong packed height for the bit width
        quant_cuda.vecquant3appendvecK(mat, lut, torch.zeros(4096), 0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kvquant_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/_ref", ""), os.path.join(dirpath, f)


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/kvquant_b200.h is a C header (no C++ or torch types): it compiles as C99 and as C++11, and a C program
    that only includes it links against libkvquant_b200.so and runs without a GPU (the cgo / JNI / ctypes view)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi_c.c"
    src.write_text('#include "kvquant_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%d %s\\n", kvq_abi_version(), kvq_error_string(KVQ_E_SHAPE)); return 0; }\n')
    inc = os.path.join(root, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    subprocess.check_call(["g++", "-std=c++11", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)])
    lib = os.path.join(root, "kvquant_b200", "libkvquant_b200.so")
    if not os.path.exists(lib) or shutil.which("gcc") is None:
        pytest.skip("library not built")
    exe = tmp_path / "abi_c"
    subprocess.check_call(["gcc", "-std=c99", "-I", inc, str(src), lib, "-o", str(exe),
                           "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath,/usr/local/cuda/lib64"])
    out = subprocess.check_output([str(exe)], text=True).split(None, 1)
    assert out[0] == "1" and out[1].strip()
