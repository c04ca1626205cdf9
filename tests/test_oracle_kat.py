"""Known-answer tests for the oracle, from the bit-level facts of the reference's source.
These hold whether or not golden fixtures exist.  CPU only."""
import hashlib
import struct

import numpy as np
import pytest


def f32_bits(x):
    return struct.unpack("<I", struct.pack("<f", float(np.float32(x))))[0]


def from_bits(u):
    return np.float32(struct.unpack("<f", struct.pack("<I", u))[0])


FP4_THRESHOLDS = {  # reference csrc/kernels.cu:141-159, as f32 bit patterns
    0.29166667: 0x3E955555, 0.583333: 0x3F155550, 0.8333333: 0x3F555555, 0.4166667: 0x3ED55556,
    0.0859375: 0x3DB00000, 0.20833333: 0x3E555555, 0.00260417: 0x3B2AAAB9,
}  # fmt: skip


def test_fp4_threshold_bit_patterns():
    for lit, pattern in FP4_THRESHOLDS.items():
        assert f32_bits(lit) == pattern
    assert f32_bits(0.583333) != f32_bits(7.0 / 12.0)  # the reference's literal is NOT 7/12


def test_fp4_quantize_codes_around_every_threshold(oracle):
    # (threshold, code for x <= t, code for x > t)  -- kernels.cu:141-162
    table = [(0.00260417, 0b0000, 0b0001), (0.0859375, 0b0001, 0b0110), (0.20833333, 0b0110, 0b0111),
             (0.29166667, 0b0111, 0b0100), (0.4166667, 0b0100, 0b0101), (0.583333, 0b0101, 0b0010),
             (0.8333333, 0b0010, 0b0011)]
    for t, below, above in table:
        t32 = np.float32(t)
        up = np.nextafter(t32, np.float32(2))
        assert oracle.quantize_fp4_scalar(t32) == below  # strict '>'
        assert oracle.quantize_fp4_scalar(up) == above
        assert oracle.quantize_fp4_scalar(-t32) == below + 8
        assert oracle.quantize_fp4_scalar(-up) == above + 8
    assert oracle.quantize_fp4_scalar(1.0) == 0b0011 and oracle.quantize_fp4_scalar(-1.0) == 0b1011
    assert oracle.quantize_fp4_scalar(0.0) == 0 and oracle.quantize_fp4_scalar(-0.0) == 0  # sign from x < 0
    assert oracle.quantize_fp4_scalar(float("nan")) == 0


def test_fp4_dequantize_constants(oracle):
    want = {0: 0x00000000, 1: 0x3BAAAAAB, 2: 0x3F2AAAAB, 3: 0x3F800000, 4: 0x3EAAAAAB, 5: 0x3F000000, 6: 0x3E2AAAAB,
            7: 0x3E800000}
    for nib, pattern in want.items():
        assert f32_bits(oracle.dequantize_fp4_scalar(nib, 1.0)) == pattern
        assert f32_bits(oracle.dequantize_fp4_scalar(nib + 8, 1.0)) == pattern | 0x80000000
    assert f32_bits(oracle.dequantize_fp4_scalar(8, 3.0)) == 0x80000000  # nibble 0b1000 -> -0.0
    # get_4bit_type's table equals the tree constants for 0..7; entry 8 is +0.0 there (core.py:218)
    code = oracle.fp4_table()
    for nib in range(16):
        if nib != 8:
            assert f32_bits(code[nib]) == f32_bits(oracle.dequantize_fp4_scalar(nib, 1.0))
    assert f32_bits(code[8]) == 0


def test_fp4_roundtrip_of_code_values(oracle):
    code = oracle.fp4_table()
    for nib in range(16):
        if nib in (0, 8):
            continue
        assert oracle.quantize_fp4_scalar(code[nib]) == nib


def test_nf4_table_and_midpoints(oracle):
    t = oracle.nf4_table()
    assert t[0] == -1 and t[7] == 0 and t[15] == 1 and np.all(np.diff(t) > 0)
    for nib in range(16):
        assert oracle.quantize_nf4_scalar(t[nib]) == nib
    mids = (t[1:].astype(np.float64) + t[:-1].astype(np.float64)) / 2
    for i, m in enumerate(mids):
        m32 = np.float32(m)  # the literal's own float32 rounding may sit one ulp either side: step two ulps
        up = np.nextafter(np.nextafter(m32, np.float32(2)), np.float32(2))
        down = np.nextafter(np.nextafter(m32, np.float32(-2)), np.float32(-2))
        assert oracle.quantize_nf4_scalar(up) == i + 1
        assert oracle.quantize_nf4_scalar(down) == i


def test_dynamic_map_pins(oracle):
    m = oracle.dynamic_map()
    assert m.shape == (256,) and m.dtype == np.float32
    assert np.all(np.diff(m) > 0)
    assert m[0] == np.float32(-0.99296874) and m[127] == 0 and m[128] == np.float32(5.5000004e-07) and m[255] == 1
    assert hashlib.sha256(m.astype("<f4").tobytes()).hexdigest() == oracle.DYNAMIC_MAP_SHA256


def test_8bit_quantize_is_not_plain_nearest_on_ties(oracle):
    """dQuantize<0> (kernels.cu:183-237) == nearest code except on exact midpoints; the oracle follows the bisection."""
    m = oracle.dynamic_map()
    for i, c in enumerate(m):
        assert oracle.quantize_8bit_scalar(m, c) == i
    mids = ((m[1:] + m[:-1]) * np.float32(0.5)).astype(np.float32)
    for i, mid in enumerate(mids):
        got = oracle.quantize_8bit_scalar(m, mid)
        assert got in (i, i + 1)
        assert oracle.quantize_8bit_scalar(m, np.nextafter(mid, np.float32(2))) in (i + 1, got)
    rng = np.random.default_rng(0)
    xs = rng.uniform(-1, 1, 20000).astype(np.float32)
    near = np.abs(xs[:, None].astype(np.float64) - m[None, :].astype(np.float64)).argmin(axis=1)
    got = np.array([oracle.quantize_8bit_scalar(m, x) for x in xs])
    assert (got == near).mean() > 0.999
    assert oracle.quantize_8bit_scalar(m, float("nan")) == 0
    assert oracle.quantize_8bit_scalar(m, 2.0) == 255 and oracle.quantize_8bit_scalar(m, -2.0) == 0


def test_packing_order_and_tails(oracle):
    # element 2i -> HIGH nibble, element 2i+1 -> LOW nibble (kernels.cu:467-468)
    a = np.zeros(64, dtype=np.float32)
    a[0], a[1] = 1.0, -1.0 / 3
    packed, absmax = oracle.quantize_blockwise_4bit(a, 64, "fp4")
    assert absmax[0] == 1 and packed[0] == (0b0011 << 4 | 0b1100)
    # odd n: (n+1)//2 bytes, missing slot quantised as 0 (kernels.cu:410,476)
    packed, absmax = oracle.quantize_blockwise_4bit(np.array([0.5, 0.25, -0.5], dtype=np.float32), 64, "fp4")
    assert packed.tolist() == [0b0011 << 4 | 0b0101, 0b1011 << 4 | 0] and absmax.tolist() == [0.5]
    # absmax == 0 -> inv = inf, 0*inf = NaN -> every compare false -> code 0
    packed, absmax = oracle.quantize_blockwise_4bit(np.zeros(64, dtype=np.float32), 64, "fp4")
    assert absmax[0] == 0 and not packed.any()
    # dequantising an odd n drops the last low nibble
    out = oracle.dequantize_blockwise_4bit(np.array([0x35, 0xB0], dtype=np.uint8), np.array([2.0], dtype=np.float32), 3, 64, "fp4",
                                           "float32")
    assert out.tolist() == [2.0, 1.0, -2.0]


def test_blocksize_1024_packing_quirk(oracle):
    """kernels.cu:450,465-470: for blocksize >= 1024 every odd byte is OR-ed with the byte before it."""
    rng = np.random.default_rng(3)
    a = rng.uniform(-1, 1, 2048).astype(np.float32)
    p512, _ = oracle.quantize_blockwise_4bit(a[:512], 512, "fp4")
    p1024, am = oracle.quantize_blockwise_4bit(a[:1024], 1024, "fp4")
    clean = []
    inv = np.float32(1.0) / am[0]
    for i in range(0, 1024, 2):
        clean.append(oracle.quantize_fp4_scalar(a[i] * inv) << 4 | oracle.quantize_fp4_scalar(a[i + 1] * inv))
    clean = np.array(clean, dtype=np.uint8)
    want = clean.copy()
    want[1::2] |= want[0::2]
    assert np.array_equal(p1024, want) and not np.array_equal(p1024, clean)
    assert p512.size == 256


@pytest.mark.parametrize("dtype,tol", [("float32", 2e-6), ("float16", 3e-3), ("bfloat16", 3e-2)])
def test_gemv_reference_order_close_to_truth(oracle, dtype, tol):
    rng = np.random.default_rng(5)
    N, K = 48, 1088
    st = oracle.quantize_4bit(oracle.round_to((rng.standard_normal((N, K)) * 0.02).astype(np.float32), "float16"), 64, "nf4")
    x = oracle.round_to(rng.standard_normal(K).astype(np.float32), dtype)
    am = oracle.state_absmax(st)
    y = oracle.gemv_4bit(x, st["packed"], am, st["code"], N, K, 64, dtype)
    truth = oracle.gemv_4bit_f64(x, st["packed"], am, st["code"], N, K, 64)
    assert np.abs(y - truth).max() <= tol * np.abs(truth).max()
    # and the fp64 truth equals a plain dequantize -> matmul
    w = oracle.dequantize_4bit(st, "float32")
    np.testing.assert_allclose(truth, w.astype(np.float64) @ x.astype(np.float64), rtol=1e-6, atol=1e-9)


def test_oracle_results_do_not_depend_on_thread_count(oracle, monkeypatch):
    rng = np.random.default_rng(9)
    a = (rng.standard_normal(64 * 4096) * 0.02).astype(np.float32)
    monkeypatch.setenv("Q4O_THREADS", "1")
    p1, m1 = oracle.quantize_blockwise_4bit(a, 64, "fp4")
    monkeypatch.setenv("Q4O_THREADS", "7")
    p7, m7 = oracle.quantize_blockwise_4bit(a, 64, "fp4")
    assert np.array_equal(p1, p7) and np.array_equal(m1, m7)
