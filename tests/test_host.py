"""Host-side logic of the reference-shaped API (quantizations_b200/core.py, modules.py) that needs no GPU."""
import hashlib

import numpy as np
import pytest
import torch

import quantizations_b200 as q


def test_dynamic_map_is_the_reference_table():
    m = q.create_dynamic_map()
    assert m.dtype == torch.float32 and m.shape == (256,)
    assert hashlib.sha256(m.numpy().astype("<f4").tobytes()).hexdigest() == \
        "e732639a65f497b4ad684bb166a4467708255edd5207757de8b8f0c7e1fda89c"


def test_get_4bit_type(oracle):
    fp4 = q.get_4bit_type("fp4", device="cpu")
    nf4 = q.get_4bit_type("nf4", device="cpu")
    assert np.array_equal(fp4.numpy().view(np.uint32), oracle.fp4_table().view(np.uint32))
    assert np.array_equal(nf4.numpy().view(np.uint32), oracle.nf4_table().view(np.uint32))
    with pytest.raises(NotImplementedError):
        q.get_4bit_type("int4", device="cpu")


def test_cpu_tensors_raise_like_the_reference():
    with pytest.raises(NotImplementedError):  # core.py:530-531
        q.quantize_4bit(torch.randn(64, 64, dtype=torch.float16))
    with pytest.raises(NotImplementedError):  # core.py:533-534 (checked after the device in the reference too)
        q.quantize_4bit(torch.randn(64, 64, dtype=torch.float16), quant_type="int4")
    with pytest.raises(ValueError):  # core.py:603-606
        q.dequantize_4bit(torch.zeros(32, 1, dtype=torch.uint8), q.QuantState(absmax=torch.zeros(1)), blocksize=100)
    with pytest.raises(ValueError):  # core.py:453-454
        q.gemv_4bit(torch.randn(1, 1, 64), torch.zeros(32, 1, dtype=torch.uint8), state=None)
    st = q.QuantState(absmax=torch.zeros(1), shape=torch.Size((1, 64)), code=torch.zeros(16), blocksize=64, quant_type="fp4")
    with pytest.raises(ValueError):  # core.py:457-460: A must be a single vector
        q.gemv_4bit(torch.randn(2, 64), torch.zeros(32, 1, dtype=torch.uint8), state=st)


def test_params4bit_constructor_contract():
    """HF/accelerate rebuild the parameter as Params4bit(value, requires_grad=False, **old.__dict__)."""
    w = torch.randn(8, 16)
    p = q.Params4bit(w, requires_grad=False, quant_type="nf4")
    assert type(p).__name__ == "Params4bit" and isinstance(p, torch.nn.Parameter)
    assert p.blocksize == 64 and p.quant_type == "nf4" and p.quant_storage == torch.uint8 and not p.bnb_quantized
    assert p.quant_state is None and p.module is None and p.compress_statistics is True
    p2 = q.Params4bit(p.data, requires_grad=False, **p.__dict__)
    assert p2.quant_type == "nf4" and torch.equal(p2.data, w)
    p3 = p.to("cpu")  # a CPU move does not quantise (core.py:176-190)
    assert not p3.bnb_quantized and torch.equal(p3.data, w)
    assert q.Params4bit().numel() == 0  # data=None -> empty (core.py:125-126)


def test_linear4bit_constructor_contract():
    lin = q.Linear4bit(32, 16, bias=False, compute_dtype=torch.bfloat16, compress_statistics=True, quant_type="nf4",
                       quant_storage=torch.uint8)
    assert isinstance(lin, torch.nn.Linear) and type(lin.weight).__name__ == "Params4bit"
    assert lin.weight.module is lin and lin.compute_dtype == torch.bfloat16 and lin.quant_state is None
    assert lin.in_features == 32 and lin.out_features == 16 and lin.bias is None and not lin.compute_type_is_set
    lin.set_compute_type(torch.zeros(1, dtype=torch.float32))
    assert lin.compute_dtype == torch.float32
    lin.set_compute_type(torch.zeros(1, dtype=torch.float16))  # fp16 input leaves it alone (modules.py:120-122)
    assert lin.compute_dtype == torch.float32
    with pytest.raises(RuntimeError):
        lin(torch.zeros(1, 1, 32))


def test_quant_state_serialisation_roundtrip():
    s2 = q.QuantState(absmax=torch.rand(3), code=q.create_dynamic_map(), blocksize=256, dtype=torch.float32)
    st = q.QuantState(absmax=torch.randint(0, 255, (600,), dtype=torch.uint8), shape=torch.Size((20, 1920)),
                      code=q.get_4bit_type("nf4", "cpu"), blocksize=64, quant_type="nf4", dtype=torch.bfloat16,
                      offset=torch.tensor(0.031), state2=s2)
    for packed in (False, True):
        d = st.as_dict(packed=packed)
        if packed:
            assert set(d) == {"absmax", "quant_map", "nested_absmax", "nested_quant_map", "quant_state.bitsandbytes__nf4"}
            d = {"weight." + k: v for k, v in d.items()}
        r = q.QuantState.from_dict(d, device="cpu")
        assert r.nested and r.quant_type == "nf4" and r.blocksize == 64 and r.dtype == torch.bfloat16
        assert tuple(r.shape) == (20, 1920) and r.state2.blocksize == 256 and r.state2.dtype == torch.float32
        assert torch.equal(r.absmax, st.absmax) and torch.equal(r.state2.absmax, s2.absmax)
        assert torch.equal(r.code, st.code) and torch.equal(r.state2.code, s2.code)
        assert abs(r.offset.item() - 0.031) < 1e-7
    with pytest.raises(ValueError):
        q.QuantState.from_dict({"foo": torch.zeros(1)}, device="cpu")


def test_native_stats_validation():
    st = q.QuantState(absmax=torch.zeros(4, dtype=torch.float16), blocksize=64)
    with pytest.raises(ValueError):
        st.native_stats()
    st = q.QuantState(absmax=torch.zeros(4, dtype=torch.float32), blocksize=64)
    s = st.native_stats()
    assert s.absmax == st.absmax.data_ptr() and not s.qabsmax and s.blocksize2 == 0
    assert st.native_stats() is s  # cached until the tensors move
    st.to("cpu")
    assert st._stats is None


def test_binned_4bit_encoder_equals_compare_tree_on_every_float(tmp_path):
    """csrc/q4_encode_lut.h (what the fast quantize kernel encodes with) against the reference's compare trees
    (kernels.cu:113-163; the 15 NF4 midpoints) on ALL 2^32 float bit patterns of its domain -- plain C on the host."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "encode_lut_check")
    subprocess.check_call(["gcc", "-O2", "-fopenmp", os.path.join(root, "tests", "encode_lut_check.c"), "-lm", "-o", exe])
    r = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


def _fake_quantised_linear(N, K, seed, nested=True):
    """A Linear4bit whose packed bytes / statistics are arbitrary CPU tensors of the right shapes (layout logic only)."""
    g = torch.Generator().manual_seed(seed)
    lin = q.Linear4bit(K, N, bias=False, compute_dtype=torch.bfloat16, quant_type="nf4", device="meta")
    nb = N * K // 64
    packed = torch.randint(0, 256, (N * K // 2, 1), dtype=torch.uint8, generator=g)
    code = q.get_4bit_type("nf4", device="cpu")
    if nested:
        s2 = q.QuantState(absmax=torch.rand(nb // 256, generator=g), code=q.create_dynamic_map(), blocksize=256, dtype=torch.float32)
        qs = q.QuantState(absmax=torch.randint(0, 256, (nb,), dtype=torch.uint8, generator=g), shape=torch.Size((N, K)), code=code,
                          blocksize=64, quant_type="nf4", dtype=torch.bfloat16, offset=torch.tensor(0.01 * seed), state2=s2)
    else:
        qs = q.QuantState(absmax=torch.rand(nb, generator=g), shape=torch.Size((N, K)), code=code, blocksize=64, quant_type="nf4",
                          dtype=torch.bfloat16)
    lin.weight = q.Params4bit(packed, requires_grad=False, quant_state=qs, quant_type="nf4", bnb_quantized=True, module=lin)
    return lin


@pytest.mark.parametrize("nested", [True, False])
@pytest.mark.parametrize("N,K", [(16, 4096), (8, 8192)])
def test_swiglu_group_interleaves_four_row_chunks(nested, N, K):
    """Linear4bitGroup(swiglu=True): rows 8t..8t+3 of the shared buffers are gate rows 4t..4t+3, rows 8t+4..8t+7 the up rows --
    packed bytes, block statistics and second-level statistics alike (what q4_gemv_fused_t's Q4_GEMV_SWIGLU flag expects)."""
    gate, up = _fake_quantised_linear(N, K, 1, nested), _fake_quantised_linear(N, K, 2, nested)
    grp = q.Linear4bitGroup([gate, up], swiglu=True)
    assert grp.swiglu and grp.out_features == 2 * N and list(grp._row_end) == [N, 2 * N]
    pk = grp.packed.view(2 * N, K // 2)
    am = grp.absmax.view(2 * N, K // 64)
    for r in range(2 * N):
        src, row = (gate, up)[(r // 4) % 2], (r // 8) * 4 + r % 4
        assert torch.equal(pk[r], src.weight.data.view(N, K // 2)[row])
        assert torch.equal(am[r], src.weight.quant_state.absmax.view(N, K // 64)[row])
    if nested:
        per = 4 * K // 64 // 256   # second-level entries per 4-row chunk
        a2 = grp.absmax2.view(2 * N // 4, per)
        for c in range(2 * N // 4):
            src = (gate, up)[c % 2]
            assert torch.equal(a2[c], src.weight.quant_state.state2.absmax.view(N // 4, per)[c // 2])
    # the members keep their own storage (prefill and the plain grouped launch do not see the interleaved copy)
    assert gate.weight.data.data_ptr() != grp.packed.data_ptr()
    with pytest.raises(ValueError):
        q.Linear4bitGroup([gate, up, gate], swiglu=True)
    with pytest.raises(ValueError):
        q.Linear4bitGroup([gate, _fake_quantised_linear(N, 2048, 3, nested)], swiglu=True)


def test_params4bit_survives_deepcopy_and_pickle():
    """copy.deepcopy / pickle of a (here: not yet quantised) Linear4bit keep the Params4bit class, its attributes and the
    module back-reference (the GPU counterpart with a quant_state is tests/test_gpu_parity.py::test_state_dict_round_trip...)."""
    import copy
    import io
    import pickle

    import torch

    import quantizations_b200 as q

    lin = q.Linear4bit(64, 32, bias=True, compute_dtype=torch.bfloat16, quant_type="nf4", compress_statistics=False)
    c = copy.deepcopy(lin)
    assert isinstance(c.weight, q.Params4bit) and c.weight is not lin.weight and c.weight.module is c
    assert (c.weight.quant_type, c.weight.blocksize, c.weight.compress_statistics, c.weight.bnb_quantized) == ("nf4", 64, False, False)
    buf = io.BytesIO()
    pickle.dump(lin.weight, buf)
    buf.seek(0)
    w = pickle.load(buf)
    assert isinstance(w, q.Params4bit) and w.quant_type == "nf4" and torch.equal(w.data, lin.weight.data)
