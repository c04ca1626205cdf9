"""The five entry points the reference's FFI exports (pythonInterface.cpp:154-161), by their reference names and argument
order, EXECUTED on the GPU through ctypes on the inputs of the golden fixtures (tests/golden/*.npz = outputs of the reference's
own object code, see make_golden.py) -- the calls a maintainer of the reference would make after re-pointing `kbkim_lib` at
libquantizations_b200.so.  They launch on the legacy default stream like the reference (ops.cu:170).

Bars: quantize / dequantize bit-exact; the fp32 GEMV within 1e-5 of the reference kernel's output (fp32 accumulation in a
different order)."""
import numpy as np
import pytest
import torch

from conftest import iter_cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def L():
    from quantizations_b200 import _lib

    return _lib.lib()


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def same_bits(got, want):
    got, want = np.ascontiguousarray(got, dtype=np.float32), np.ascontiguousarray(want, dtype=np.float32)
    return ((got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))).all()


def test_cquantize_blockwise_fp16_fp4_on_the_golden_inputs(L, golden):
    g = golden("quantize_fp4")
    ran = 0
    for k, (dtype, blocksize, n, kind) in iter_cases(g):
        if dtype != "float16":  # the exported name is the fp16 instance (pythonInterface.cpp:82)
            continue
        blocksize, n = int(blocksize), int(n)
        A = dev(g[k + "_in"], torch.float16)
        absmax = torch.zeros(-(n // -blocksize), device=DEV, dtype=torch.float32)
        out = torch.zeros((n + 1) // 2, device=DEV, dtype=torch.uint8)
        with torch.cuda.stream(torch.cuda.default_stream()):
            rc = L.cquantize_blockwise_fp16_fp4(None, A.data_ptr(), absmax.data_ptr(), out.data_ptr(), blocksize, n)
            torch.cuda.synchronize()
        assert rc == 0
        assert np.array_equal(out.cpu().numpy(), g[k + "_packed"]), (k, kind)
        assert same_bits(absmax.cpu().numpy(), g[k + "_absmax"]), (k, kind)
        ran += 1
    assert ran >= 10


def test_cdequantize_blockwise_fp16_fp4_on_the_golden_inputs(L, golden):
    g = golden("dequantize_fp4")
    ran = 0
    for k, (dtype, blocksize, n) in iter_cases(g):
        if dtype != "float16":
            continue
        blocksize, n = int(blocksize), int(n)
        P, AM = dev(g[k + "_packed"]), dev(g[k + "_absmax"])
        out = torch.zeros(n, device=DEV, dtype=torch.float16)
        with torch.cuda.stream(torch.cuda.default_stream()):
            rc = L.cdequantize_blockwise_fp16_fp4(None, P.data_ptr(), AM.data_ptr(), out.data_ptr(), blocksize, n)
            torch.cuda.synchronize()
        assert rc == 0
        assert same_bits(out.float().cpu().numpy(), g[k + "_out"]), k
        ran += 1
    assert ran >= 5


def test_cquantize_and_cdequantize_blockwise_fp32_on_the_golden_inputs(L, golden):
    g = golden("blockwise_8bit")
    code = dev(g["code"])
    ran = 0
    for k, (name, blocksize) in iter_cases(g):
        blocksize = int(blocksize)
        a = g[k + "_in"]
        n = a.size
        A = dev(a)
        absmax = torch.zeros(-(n // -blocksize), device=DEV, dtype=torch.float32)
        out = torch.zeros(n, device=DEV, dtype=torch.uint8)
        deq = torch.zeros(n, device=DEV, dtype=torch.float32)
        with torch.cuda.stream(torch.cuda.default_stream()):
            rc = L.cquantize_blockwise_fp32(code.data_ptr(), A.data_ptr(), absmax.data_ptr(), out.data_ptr(), blocksize, n)
            rc2 = L.cdequantize_blockwise_fp32(code.data_ptr(), out.data_ptr(), absmax.data_ptr(), deq.data_ptr(), blocksize, n)
            torch.cuda.synchronize()
        assert rc == 0 and rc2 == 0
        assert np.array_equal(out.cpu().numpy(), g[k + "_q"]), (k, name)
        assert same_bits(absmax.cpu().numpy(), g[k + "_absmax"]), (k, name)
        assert same_bits(deq.cpu().numpy(), g[k + "_deq"]), (k, name)
        ran += 1
    assert ran >= 10


def test_cgemm_4bit_inference_naive_fp32_on_the_golden_inputs(L, golden):
    g = golden("gemv")
    ran = 0
    for k, (dtype, N, K, code_name) in iter_cases(g):
        if dtype != "float32":  # the exported name is the fp32 instance (pythonInterface.cpp:60-64)
            continue
        N, K = int(N), int(K)
        x, P, AM, C = dev(g[k + "_x"]), dev(g[k + "_packed"]), dev(g[k + "_absmax"]), dev(g[k + "_code"])
        out = torch.zeros(N, device=DEV, dtype=torch.float32)
        with torch.cuda.stream(torch.cuda.default_stream()):
            rc = L.cgemm_4bit_inference_naive_fp32(N, 1, K, x.data_ptr(), P.data_ptr(), AM.data_ptr(), C.data_ptr(), out.data_ptr(),
                                                   N, (K + 1) // 2, N, 64)
            torch.cuda.synchronize()
        assert rc == 0
        want = g[k + "_out"]
        err = np.abs(out.cpu().numpy() - want).max() / np.abs(want).max()
        assert err <= 1e-5, (k, code_name, err)
        ran += 1
    assert ran >= 10
