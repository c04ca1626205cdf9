"""Register / spill budgets of the hot kernels, read from the ptxas logs the in-tree build leaves next to the objects
(quantizations_b200/_build/*.o.log, `-Xptxas -v`).  The decode GEMV runs two 256-thread CTAs per SM, i.e. at most 128 registers
per thread and no spills; a change that pushes the PLAIN instantiation past its 125 shows up as a slower loop before it shows up
anywhere else (a run-time SwiGLU flag once cost it 3 registers and 5.6 % -- it is a template parameter since)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "quantizations_b200", "_build")


def _kernels(log):
    path = os.path.join(BUILD, log)
    if not os.path.exists(path):
        pytest.skip(f"{log} not found: run __graft_entry__.build() first")
    text = open(path).read()
    out = {}
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?(\d+) bytes spill stores, (\d+) bytes spill loads.*?Used (\d+) registers",
                         text, re.S):
        out[m.group(1)] = (int(m.group(4)), int(m.group(2)) + int(m.group(3)))
    return out


def test_decode_gemv_register_budget():
    ks = {k: v for k, v in _kernels("q4_gemv.o.log").items() if "gemv_mma_kernel" in k}
    assert ks, "no gemv_mma_kernel instantiation in the build log"
    for name, (regs, spill) in ks.items():
        chained = "Lb0ELb1ELb0EEEv" in name  # <.., TAIL = 0, CHAIN = 1, SWIGLU = 0>: the opt-in chained launch spills 12 bytes
        assert regs <= 128 and (spill == 0 or (chained and spill <= 32)), (name, regs, spill)   # two 256-thread CTAs per SM
    # <T, NESTED, COMPACT, TAIL, CHAIN, SWIGLU> = <bf16 / half, 1, 1, 0, 0, 0>: the instantiation every Llama shape runs
    plain = [v for k, v in ks.items() if "Lb1ELb1ELb0ELb0ELb0E" in k]
    assert len(plain) == 2 and all(regs <= 126 for regs, _ in plain), plain


def test_blockwise_kernels_register_budget():
    q = {k: v for k, v in _kernels("q4_quantize.o.log").items() if "quantize_4bit_lut_kernel" in k}
    assert q and all(regs <= 64 and spill == 0 for regs, spill in q.values()), q   # four 256-thread CTAs per SM
    d = {k: v for k, v in _kernels("q4_dequantize.o.log").items() if "dequantize_4bit_kernel" in k}
    assert d and all(regs <= 64 and spill == 0 for regs, spill in d.values()), d


def test_ring_tokens_and_gemm_register_budget():
    """The persistent ring GEMV is compiled for 640 threads per CTA (96 registers, raised to 112 for the consumers by setmaxnreg): its
    nested instantiations must not spill more than they do today -- every value that lives across the slot loop costs more than it saves.  The
    small-batch kernel runs 512 threads at 128 registers, the prefill GEMM 512 threads (three dequantise sets) below 128."""
    ring = {k: v for k, v in _kernels("q4_gemv_ring.o.log").items() if "gemv_ring_kernel" in k}
    assert ring and all(regs <= 96 for regs, _ in ring.values()), ring
    # <T, NESTED = 1, 16, 2> keeps one 4-byte value on the stack outside the slot loop (4 bytes stored, 8 loaded); anything more is a
    # value that lives across the loop
    assert all(spill <= 16 for k, (_, spill) in ring.items() if "Lb1ELi16ELi2E" in k), ring
    assert all(spill == 0 for k, (_, spill) in ring.items() if "Lb0ELi16ELi2E" in k), ring
    tok = {k: v for k, v in _kernels("q4_gemv_tokens.o.log").items() if "gemv_tokens_kernel" in k}
    assert len(tok) == 4 and all(regs <= 128 and spill == 0 for regs, spill in tok.values()), tok
    gemm = {k: v for k, v in _kernels("q4_gemm.o.log").items() if "gemm_dequant_tcgen05_kernel" in k}
    assert len(gemm) == 4 and all(regs <= 128 and spill == 0 for regs, spill in gemm.values()), gemm
