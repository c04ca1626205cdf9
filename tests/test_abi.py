"""The C-ABI library loads and exports every symbol include/quantizations_b200.h declares; argument validation
returns the documented negative codes before anything is launched (so this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "quantizations_b200.h")


@pytest.fixture(scope="module")
def L():
    from quantizations_b200 import _lib, build

    build.build()
    return _lib.lib()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|int64_t|char\s*\*|void)\s*\*?\s*(\w+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_reference_entry_points():
    names = declared_functions()
    for ref in ("cgemm_4bit_inference_naive_fp32", "cquantize_blockwise_fp16_fp4", "cdequantize_blockwise_fp16_fp4",
                "cquantize_blockwise_fp32", "cdequantize_blockwise_fp32"):  # reference pythonInterface.cpp:154-161
        assert ref in names
    assert len(names) >= 15 and "q4_gemm_4bit" in names


def test_library_exports_every_declared_symbol(L):
    for name in declared_functions():
        assert hasattr(L, name), f"{name} is declared in include/quantizations_b200.h but not exported"


def test_python_binding_covers_the_header(L):
    from quantizations_b200 import _lib

    assert sorted(_lib.exported_symbols()) == declared_functions()
    assert L.q4_abi_version() == 1


def test_argument_errors_are_reported_not_launched(L):
    from quantizations_b200._lib import AbsmaxStats

    st = AbsmaxStats(1, None, None, None, None, 0)
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(4096)
    # bad blocksize (core.py:350,408,549,603)
    assert L.q4_quantize_blockwise_4bit(one, one, one, 100, 64, 1, 1, null) == -1
    assert L.q4_dequantize_blockwise_4bit(one, ctypes.byref(st), one, 48, 64, 1, 1, null) == -1
    assert L.q4_quantize_blockwise_8bit(one, one, one, one, 3, 64, null) == -1
    # bad dtype / quant type
    assert L.q4_quantize_blockwise_4bit(one, one, one, 64, 64, 1, 9, null) == -2
    assert L.q4_quantize_blockwise_4bit(one, one, one, 64, 64, 7, 1, null) == -3
    assert L.q4_dequantize_blockwise_4bit(one, ctypes.byref(st), one, 64, 64, 0, 1, null) == -3
    # shapes: negative n, odd K, n != 1 for the reference-named GEMV
    assert L.q4_quantize_blockwise_4bit(one, one, one, 64, -1, 1, 1, null) == -4
    assert L.q4_gemv_4bit(one, one, ctypes.byref(st), one, null, one, 8, 7, 64, 1, 0, null, 0, null) == -4
    assert L.cgemm_4bit_inference_naive_fp32(8, 2, 64, one, one, one, one, one, 8, 32, 8, 64) == -4
    assert L.q4_gemm_4bit(one, one, ctypes.byref(st), one, null, one, 4, 8, 100, 64, 1, null) == -4   # K % 64
    assert L.q4_gemm_4bit(one, one, ctypes.byref(st), one, null, one, 4, 8, 128, 128, 1, null) == -1  # blocksize != 64
    assert L.q4_gemm_4bit(one, one, ctypes.byref(st), one, null, one, 4, 8, 128, 64, 0, null) == -2   # fp32 not supported
    # NULL pointers
    assert L.q4_quantize_blockwise_4bit(null, one, one, 64, 64, 1, 1, null) == -5
    assert L.q4_gemv_4bit(one, one, None, one, null, one, 8, 64, 64, 1, 0, null, 0, null) == -5
    nested_missing = AbsmaxStats(None, 4096, None, None, None, 256)
    assert L.q4_gemv_4bit(one, one, ctypes.byref(nested_missing), one, null, one, 8, 64, 64, 1, 0, null, 0, null) == -5
    # alignment
    assert L.q4_quantize_blockwise_4bit(ctypes.c_void_p(4100), one, one, 64, 64, 1, 1, null) == -6
    # empty inputs are a no-op success (n == 0)
    assert L.q4_quantize_blockwise_4bit(null, null, null, 64, 0, 1, 1, null) == 0
    assert L.q4_dequantize_blockwise_4bit(null, ctypes.byref(st), null, 64, 0, 1, 1, null) == 0
    assert L.q4_gemv_4bit(one, one, ctypes.byref(st), one, null, one, 0, 64, 64, 1, 0, null, 0, null) == 0
    assert b"blocksize" in L.q4_error_string(-1) and L.q4_error_string(0) == b"success"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from quantizations_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.Q4Error, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "quantizations_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "q4_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_ctypes_structs_have_the_headers_layout(tmp_path):
    """The Python mirror of q4_gemv_fused_t / q4_absmax_t / q4_allreduce_t (quantizations_b200/_lib.py) must have exactly the C
    layout of include/quantizations_b200.h: compile a probe against the header and compare sizes and every field offset."""
    import ctypes
    import subprocess

    from quantizations_b200 import _lib

    structs = {"q4_gemv_fused_t": _lib.GemvFused, "q4_absmax_t": _lib.AbsmaxStats, "q4_allreduce_t": _lib.AllReduce}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "quantizations_b200.h")}"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "probe")
    subprocess.check_call(["gcc", "-std=c11", str(src), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out.strip().splitlines()}
    for cname, cls in structs.items():
        assert got[(cname, "size")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, f"{cname}.{fname}"
