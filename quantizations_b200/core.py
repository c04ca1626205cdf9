"""Functional layer of the B200 4-bit Linear engine -- host-side mirror of the reference's core.py.

Same public names, signatures, return layouts and error behaviour as /root/reference/core.py (cited per function) so
that callers -- including HF transformers' bnb-linear replacement -- can switch by changing one import.  What differs:

* every native call goes through the C ABI of libquantizations_b200.so (quantizations_b200/_lib.py), on the current
  torch CUDA stream, with its return code checked (the reference launches on the legacy stream and checks nothing);
* `quant_type` accepts "nf4" as well as "fp4" and `compress_statistics=False` is honoured (the reference is
  FP4-only and always nests, core.py:27,533,563-565);
* `gemv_4bit` is ONE launch: the 8-bit absmax decode (+ offset) the reference runs as two extra launches on every
  call (core.py:467-468) is fused into the GEMV kernel; `dequantize_4bit` likewise fuses it (core.py:613-617);
* fp16 / bf16 / fp32 inputs are dispatched by dtype (the reference reinterprets every input as fp16 for quantize and
  as fp32 for GEMV regardless of the tensor, pythonInterface.cpp:60-64,82).

There is no CPU path: a non-CUDA tensor raises NotImplementedError exactly where the reference does (core.py:530).
"""
from __future__ import annotations

import contextlib
import os
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import AbsmaxStats, Q4Error, check

name2qmap = {}

dtype2bytes = {}
dtype2bytes[torch.uint8] = 1

_VALID_BLOCKSIZES = [4096, 2048, 1024, 512, 256, 128, 64]
_DTYPE_CODE = {torch.float32: _lib.Q4_F32, torch.float16: _lib.Q4_F16, torch.bfloat16: _lib.Q4_BF16}
_QUANT_CODE = {"fp4": _lib.Q4_FP4, "nf4": _lib.Q4_NF4}

# NF4 table: the reference's only NF4 artefact, csrc/kernels.cu:851
_NF4_VALUES = [
    -1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
    -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
    0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941,
    0.7229568362236023, 1.0,
]  # fmt: skip


def _stream(t: Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


_NO_GUARD = contextlib.nullcontext()


def _on_device(dev: torch.device):
    """Context for a launch whose pointers live on `dev`: the library sizes grids from, and sets kernel attributes on, the CURRENT
    device, so a launch for another GPU of the same process (HF device_map sharding, multi-GPU tests) must switch to it first.
    Free when `dev` is already current (the common case): no context manager is entered."""
    idx = dev.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(idx)


# Split-K workspace of the tcgen05 decode GEMV (include/quantizations_b200.h: q4_gemv_fused_t.workspace): zeroed once, one per
# (device, stream) because launches on different streams may run concurrently.
_gemv_workspaces = {}


# The tcgen05 kernel is opt-in (Q4_GEMV_TC=1): on B200 it currently measures slower than the mma.sync kernel.
_USE_TC = os.environ.get("Q4_GEMV_TC", "0") == "1"


def _ws_args(device):
    """(pointer, bytes) of the split-K workspace to pass with a decode GEMV: (None, 0) selects the mma.sync kernel."""
    if not _USE_TC:
        return None, 0
    return gemv_workspace(device).data_ptr(), _lib.Q4_GEMV_WORKSPACE_BYTES


def gemv_workspace(device) -> Tensor:
    key = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream)
    ws = _gemv_workspaces.get(key)
    if ws is None:
        ws = torch.zeros(_lib.Q4_GEMV_WORKSPACE_BYTES, dtype=torch.uint8, device=device)
        _gemv_workspaces[key] = ws
    return ws


# Decode-GEMV table images (include/quantizations_b200.h: q4_gemv_lut_build), one per (device, dtype, table contents): every
# Linear4bit of a model shares a single 64-KB image, so it stays L2-resident and each launch fetches it with one bulk copy.
_gemv_luts = {}


def gemv_lut(code: Tensor, code2: Optional[Tensor], dtype: torch.dtype) -> Tensor:
    """Table image for the given 4-bit code / 8-bit absmax code / activation dtype; built once per distinct contents.
    The first call for a given pair of tensors reads them back to the host (1 KB) to key the cache -- do it outside CUDA-graph
    capture (quantize_4bit does it at load time)."""
    key = (code.device, dtype, code.data_ptr(), 0 if code2 is None else code2.data_ptr(), code._version,
           0 if code2 is None else code2._version)
    hit = _gemv_luts.get(key)
    if hit is not None:
        return hit[0]
    content = (code.device, dtype, code.detach().float().cpu().numpy().tobytes(),
               b"" if code2 is None else code2.detach().float().cpu().numpy().tobytes())
    hit = _gemv_luts.get(content)
    if hit is None:
        lut = torch.empty(_lib.Q4_GEMV_LUT_BYTES, dtype=torch.uint8, device=code.device)
        c32 = code.detach().float().contiguous()
        c2 = None if code2 is None else code2.detach().float().contiguous()
        with torch.cuda.device(code.device):
            check(_lib.lib().q4_gemv_lut_build(c32.data_ptr(), None if c2 is None else c2.data_ptr(), _DTYPE_CODE[dtype],
                                               lut.data_ptr(), _stream(code)), "gemv_lut_build")
        hit = (lut,)
        _gemv_luts[content] = hit
    _gemv_luts[key] = (hit[0], code, code2)  # keeps the keyed tensors alive: their addresses cannot be recycled
    return hit[0]


class QuantState:
    """Container for the quantisation statistics of one tensor; layout of reference core.py:23-88.

    absmax  uint8 [nblocks] when nested (8-bit codes of absmax - offset), else float32 [nblocks]
    shape   torch.Size of the original [N, K] weight          code    float32 [16] 4-bit code table
    dtype   dtype of the original weight                      blocksize  64 by default
    offset  0-dim float32 tensor (mean of absmax)             state2  QuantState of the 8-bit level
                                                                       (absmax float32 [ceil(nblocks/256)],
                                                                        code float32 [256], blocksize 256)
    """

    valid_quant_types = ("fp4", "nf4")
    valid_qs_type_keys = [f"bitsandbytes__{x}" for x in valid_quant_types]
    valid_qs_keys = [
        "absmax",
        "quant_map",
        "nested_absmax",
        "nested_quant_map",
        "quant_state",
        "quant_type",
        "blocksize",
        "dtype",
        "shape",
        "nested_blocksize",
        "nested_dtype",
        "nested_offset",
    ]

    def __init__(self, absmax, shape=None, code=None, blocksize=None, quant_type=None, dtype=None, offset=None, state2=None):
        self.absmax = absmax
        self.shape = shape
        self.code = code
        self.dtype = dtype
        self.blocksize = blocksize
        self.quant_type = quant_type
        self.offset = offset
        self.state2 = state2
        self.nested = state2 is not None
        self._stats = None  # cached C struct of pointers (see native_stats)
        self._luts = {}     # decode-GEMV table image per activation dtype (see gemv_lut)

    def __getstate__(self):
        """copy / pickle: the tensors travel, the caches (C struct of raw pointers, table images) are rebuilt on first use"""
        state = self.__dict__.copy()
        state["_stats"] = None
        state["_luts"] = {}
        return state

    def to(self, device):
        """Move the statistics to `device` (reference core.py:78-88; also handles a non-nested state)."""
        self.absmax = self.absmax.to(device)
        if self.code is not None:
            self.code = self.code.to(device)
        if self.nested:
            self.offset = self.offset.to(device)
            self.state2.absmax = self.state2.absmax.to(device)
            self.state2.code = self.state2.code.to(device)
            self.state2._stats = None
        self._stats = None
        self._luts = {}

    def lut(self, dtype: torch.dtype) -> Tensor:
        """Decode-GEMV table image for activations of `dtype` (fp16 / bf16)."""
        t = self._luts.get(dtype)
        if t is None:
            t = gemv_lut(self.code, self.state2.code if self.nested else None, dtype)
            self._luts[dtype] = t
        return t

    def native_stats(self) -> AbsmaxStats:
        """q4_absmax_t view of this state (include/quantizations_b200.h); cached until the tensors move."""
        key = (self.absmax.data_ptr(), self.state2.absmax.data_ptr() if self.nested else 0)
        if self._stats is None or self._stats[0] != key:
            if self.nested:
                s2 = self.state2
                if self.absmax.dtype != torch.uint8 or s2.absmax.dtype != torch.float32:
                    raise ValueError("nested QuantState needs uint8 absmax and float32 state2.absmax")
                st = AbsmaxStats(None, self.absmax.data_ptr(), s2.code.data_ptr(), s2.absmax.data_ptr(),
                                 self.offset.data_ptr(), int(s2.blocksize))
            else:
                if self.absmax.dtype != torch.float32:
                    raise ValueError("QuantState.absmax must be float32 when not nested")
                st = AbsmaxStats(self.absmax.data_ptr(), None, None, None, None, 0)
            self._stats = (key, st)
        return self._stats[1]

    # ---- bnb-compatible serialisation (the on-disk format of this path; keys listed but unused in the
    # reference, core.py:29-42; HF's Bnb4bitDeserialize / save_pretrained need them) -------------------------

    def as_dict(self, packed: bool = False) -> dict:
        """Tensors + metadata under bitsandbytes' state-dict key names."""
        qs_dict = {
            "quant_type": self.quant_type,
            "absmax": self.absmax,
            "blocksize": self.blocksize,
            "quant_map": self.code,
            "dtype": str(self.dtype).strip("torch."),
            "shape": tuple(self.shape),
        }
        if self.nested:
            qs_dict.update(
                {
                    "nested_absmax": self.state2.absmax,
                    "nested_blocksize": self.state2.blocksize,
                    "nested_quant_map": self.state2.code.clone(),
                    "nested_dtype": str(self.state2.dtype).strip("torch."),
                    "nested_offset": self.offset.item(),
                }
            )
        if not packed:
            return qs_dict
        import json

        tensors = {k: v for k, v in qs_dict.items() if isinstance(v, torch.Tensor)}
        meta = {k: v for k, v in qs_dict.items() if not isinstance(v, torch.Tensor)}
        blob = torch.tensor(list(json.dumps(meta).encode("utf-8")), dtype=torch.uint8)
        tensors["quant_state." + "bitsandbytes__" + self.quant_type] = blob
        return tensors

    @classmethod
    def from_dict(cls, qs_dict: dict, device) -> "QuantState":
        """Inverse of as_dict(packed=True/False); accepts keys with or without a leading module prefix."""
        import json

        qs_key = [k for k, v in qs_dict.items() if "quant_state" in k and isinstance(v, torch.Tensor)]
        if not len(qs_key) and "quant_type" not in qs_dict:
            raise ValueError("Expected packed or unpacked quant_state items, found neither")
        elif len(qs_key) > 1 or (len(qs_key) == 1 and qs_key[0].split(".")[-1] not in cls.valid_qs_type_keys):
            raise ValueError(f"There should be exactly one `quant_state` item with ending from {cls.valid_qs_type_keys}.")
        qs_dict = dict(qs_dict)
        if len(qs_key) == 1:
            blob = qs_dict.pop(qs_key[0])
            qs_dict.update(json.loads(bytes(blob.cpu().tolist()).decode("utf-8")))
        qs_dict = {k.split(".")[-1]: v for k, v in qs_dict.items()}
        if not set(qs_dict.keys()).issubset(cls.valid_qs_keys):
            raise ValueError(f"unexpected quant_state keys: {set(qs_dict) - set(cls.valid_qs_keys)}")

        if "nested_absmax" in qs_dict:
            offset = torch.tensor(float(qs_dict["nested_offset"]), dtype=torch.float32, device=device)
            state2 = cls(
                absmax=qs_dict["nested_absmax"].to(device),
                blocksize=qs_dict["nested_blocksize"],
                code=qs_dict["nested_quant_map"].to(device),
                dtype=getattr(torch, qs_dict["nested_dtype"]),
            )
        else:
            offset, state2 = None, None
        return cls(
            quant_type=qs_dict["quant_type"],
            absmax=qs_dict["absmax"].to(device),
            blocksize=qs_dict["blocksize"],
            code=qs_dict["quant_map"].to(device),
            dtype=getattr(torch, qs_dict["dtype"]),
            shape=torch.Size(qs_dict["shape"]) if qs_dict["shape"] is not None else None,
            offset=offset,
            state2=state2,
        )


class Params4bit(torch.nn.Parameter):
    """4-bit quantised parameter; same constructor and `.to()` behaviour as reference core.py:91-190.

    HF / accelerate rebuild the parameter as `Params4bit(value, requires_grad=False, **old.__dict__).to(device)`, so
    `__new__` accepts exactly the attributes it sets (plus `compress_statistics`, which the reference's Linear4bit
    accepts and ignores, modules.py:80).
    """

    def __new__(
        cls,
        data: Optional[Tensor] = None,
        requires_grad=False,
        quant_state: Optional[QuantState] = None,
        blocksize: int = 64,
        quant_type: str = "fp4",
        quant_storage: torch.dtype = torch.uint8,
        module: Optional["Linear4bit"] = None,  # noqa: F821
        bnb_quantized: bool = False,
        compress_statistics: bool = True,
    ) -> "Params4bit":
        if data is None:
            data = torch.empty(0)
        self = Tensor._make_subclass(cls, data, requires_grad)
        self.blocksize = blocksize
        self.compress_statistics = compress_statistics
        self.quant_type = quant_type
        self.quant_state = quant_state
        self.quant_storage = quant_storage
        self.bnb_quantized = bnb_quantized
        self.data = data
        self.module = module
        return self

    @classmethod
    def from_prequantized(cls, data: Tensor, quantized_stats: dict, requires_grad: bool = False, device="cuda",
                          module=None, **kwargs) -> "Params4bit":
        """Rebuild a parameter from packed bytes + serialised statistics (bnb checkpoint format; absent in the
        reference, needed by HF's Bnb4bitDeserialize)."""
        self = Tensor._make_subclass(cls, data.to(device), requires_grad)
        self.quant_state = QuantState.from_dict(qs_dict=quantized_stats, device=device)
        self.blocksize = self.quant_state.blocksize
        self.compress_statistics = self.quant_state.nested
        self.quant_type = self.quant_state.quant_type
        self.quant_storage = data.dtype
        self.bnb_quantized = True
        self.module = module
        if module is not None:
            module.quant_state = self.quant_state
        return self

    # ---- copying / pickling keep the quantisation state (upstream bitsandbytes semantics; the reference has neither: a deepcopy or
    # pickle of its parameter silently drops quant_state and re-quantises nothing) ------------------------------------------------

    def __getstate__(self):
        state = self.__dict__.copy()
        state["data"] = self.data
        state["requires_grad"] = self.requires_grad
        state["module"] = None  # the owning module is re-attached by whoever rebuilds it
        return state

    def __setstate__(self, state):
        self.requires_grad = state["requires_grad"]
        self.blocksize = state["blocksize"]
        self.compress_statistics = state["compress_statistics"]
        self.quant_type = state["quant_type"]
        self.quant_state = state["quant_state"]
        self.data = state["data"]
        self.quant_storage = state["quant_storage"]
        self.bnb_quantized = state["bnb_quantized"]
        self.module = state.get("module")

    def __reduce_ex__(self, proto):
        return (_rebuild_params4bit, (self.__getstate__(),))

    def __deepcopy__(self, memo):
        import copy

        state = self.__getstate__()
        state["data"] = copy.deepcopy(state["data"], memo)
        state["quant_state"] = copy.deepcopy(state["quant_state"], memo)
        new = _rebuild_params4bit(state)
        new.module = memo.get(id(self.module)) if self.module is not None else None
        return new

    def __copy__(self):
        return _rebuild_params4bit(self.__getstate__())

    def _quantize(self, device):
        """reference core.py:139-161"""
        w = self.data.contiguous().cuda(device)
        w_4bit, quant_state = quantize_4bit(
            w,
            blocksize=self.blocksize,
            quant_type=self.quant_type,
            quant_storage=self.quant_storage,
            compress_statistics=self.compress_statistics,
        )
        self.data = w_4bit
        self.quant_state = quant_state
        if self.module is not None:
            self.module.quant_state = quant_state
        self.bnb_quantized = True
        return self

    def cuda(self, device=None, non_blocking: bool = False):
        return self.to(device="cuda" if device is None else device, non_blocking=non_blocking)

    def to(self, *args, **kwargs):
        """First move to a CUDA device quantises (reference core.py:164-190); later moves carry the state along."""
        device, dtype, non_blocking, convert_to_format = torch._C._nn._parse_to(*args, **kwargs)

        if device is not None and device.type == "cuda" and not self.bnb_quantized:
            return self._quantize(device)
        if self.quant_state is not None and device is not None:
            self.quant_state.to(device)
        new_param = Params4bit(
            super().to(device=device, dtype=dtype, non_blocking=non_blocking),
            requires_grad=self.requires_grad,
            quant_state=self.quant_state,
            blocksize=self.blocksize,
            quant_type=self.quant_type,
            quant_storage=self.quant_storage,
            module=self.module,
            bnb_quantized=self.bnb_quantized,
            compress_statistics=self.compress_statistics,
        )
        return new_param


def _rebuild_params4bit(state: dict) -> "Params4bit":
    p = Tensor._make_subclass(Params4bit, state["data"], state["requires_grad"])
    p.__setstate__(state)
    return p


def get_4bit_type(typename, device=None, blocksize=64):
    """16-entry code table, float32, normalised to max |v| = 1.  reference core.py:193-229 ("fp4"); "nf4" is the table
    at reference csrc/kernels.cu:851."""
    if device is None:
        device = "cuda"
    if typename == "fp4":
        # index = nibble: 0b000 0, 0b001 0.0625, 0b010 8, 0b011 12, 0b100 4, 0b101 6, 0b110 2, 0b111 3 (x1/12);
        # bit 3 = sign.  Entry 8 is +0.0 as in the reference (its list holds the integer literal -0).
        magnitudes = [0.0, 0.0625, 8.0, 12.0, 4.0, 6.0, 2.0, 3.0]
        data = magnitudes + [0.0] + [-m for m in magnitudes[1:]]
    elif typename == "nf4":
        data = list(_NF4_VALUES)
    else:
        raise NotImplementedError(f"Typename {typename} not supported")
    data = torch.tensor(data, device=device)
    data.div_(data.abs().max())
    assert data.numel() == 16
    return data


def get_ptr(A: Optional[Tensor]) -> int:
    """reference core.py:232-248"""
    return 0 if A is None else A.data.data_ptr()


def create_dynamic_map(signed=True, max_exponent_bits=7, total_bits=8):
    """Dynamic 8-bit quantisation map (256 sorted float32 codes).  reference core.py:251-314.

    Built with the same torch CPU ops (torch.linspace in float32, Python-float scaling, sort) so the table is
    bit-identical to the reference's: tests pin its sha256.
    """
    non_sign_bits = total_bits - 1  # the reference subtracts 1 whether or not `signed` (core.py:274)
    extra = 2 ** (non_sign_bits - max_exponent_bits) - 1
    signs = (1.0, -1.0) if signed else (1.0,)
    values = []
    i = 0
    for i in range(max_exponent_bits):
        n_edges = int(2 ** (i + non_sign_bits - max_exponent_bits + (0 if signed else 1)) + 1)
        edges = torch.linspace(0.1, 1, n_edges)
        centres = (edges[:-1] + edges[1:]) / 2.0
        for sgn in signs:
            values += ((sgn * 10 ** (-(max_exponent_bits - 1) + i)) * centres).tolist()
    if extra > 0:
        edges = torch.linspace(0.1, 1, extra + 1)
        centres = (edges[:-1] + edges[1:]) / 2.0
        for sgn in signs:
            values += ((sgn * 10 ** (-(max_exponent_bits - 1) + i)) * centres).tolist()
    values += [0, 1.0]
    assert len(values) == 2**total_bits
    values += [0] * (256 - len(values))
    values.sort()
    return Tensor(values)


def _dynamic_map(device) -> Tensor:
    key = ("dynamic", str(device))
    if key not in name2qmap:
        if "dynamic" not in name2qmap:
            name2qmap["dynamic"] = create_dynamic_map()
        name2qmap[key] = name2qmap["dynamic"].to(device)
    return name2qmap[key]


def quantize_blockwise(A: Tensor, blocksize=4096) -> Tuple[Tensor, QuantState]:
    """8-bit blockwise quantisation of a float32 tensor with the dynamic map.  reference core.py:317-366."""
    if A.device.type != "cuda":
        raise NotImplementedError(f"Device type not supported for blockwise quantization: {A.device.type}")
    if A.dtype != torch.float32:
        raise NotImplementedError(f"quantize_blockwise expects float32 input, got {A.dtype}")
    code = _dynamic_map(A.device)
    A = A.contiguous()
    n = A.numel()
    blocks = -(n // -blocksize)
    assert blocksize in _VALID_BLOCKSIZES
    absmax = torch.zeros((blocks,), device=A.device, dtype=torch.float32)
    out = torch.zeros_like(A, dtype=torch.uint8)
    with torch.cuda.device(A.device):
        check(_lib.lib().q4_quantize_blockwise_8bit(code.data_ptr(), A.data_ptr(), absmax.data_ptr(), out.data_ptr(),
                                                    blocksize, n, _stream(A)), "quantize_blockwise")
    return out, QuantState(absmax=absmax, code=code, blocksize=blocksize, dtype=A.dtype)


def dequantize_blockwise(
    A: Tensor,
    quant_state: Optional[QuantState] = None,
    absmax: Optional[Tensor] = None,
    code: Optional[Tensor] = None,
    out: Optional[Tensor] = None,
    blocksize: int = 4096,
    nested=False,
) -> Tensor:
    """8-bit blockwise dequantisation to float32: out[i] = code[A[i]] * absmax[i // blocksize].  reference core.py:369-423
    (which requires quant_state despite the assert; here absmax+code+blocksize are honoured when it is None)."""
    assert quant_state is not None or absmax is not None
    if quant_state is None:
        quant_state = QuantState(absmax=absmax, code=_dynamic_map(A.device) if code is None else code,
                                 blocksize=blocksize, dtype=torch.float32)
    if quant_state.blocksize not in _VALID_BLOCKSIZES:
        raise ValueError(
            f"The blockwise of {quant_state.blocksize} is not supported. Supported values: [2048, 4096, 1024, 512, 256, 128, 64]",
        )
    if out is None:
        out = torch.empty(A.shape, dtype=quant_state.dtype, device=A.device)
    if out.dtype != torch.float32:
        raise NotImplementedError(f"dequantize_blockwise writes float32, got out dtype {out.dtype}")
    qcode = quant_state.code.to(A.device)
    A = A.contiguous()
    with torch.cuda.device(A.device):
        check(_lib.lib().q4_dequantize_blockwise_8bit(qcode.data_ptr(), A.data_ptr(), quant_state.absmax.data_ptr(),
                                                      out.data_ptr(), quant_state.blocksize, A.numel(), _stream(A)),
              "dequantize_blockwise")
    return out


def gemv_4bit(
    A: Tensor,
    B: Tensor,
    out: Optional[Tensor] = None,
    transposed_A=False,
    transposed_B=False,
    state=None,
    bias: Optional[Tensor] = None,
    flags: int = _lib.Q4_GEMV_DEFAULT,
    prefetch: Optional[Tensor] = None,
):
    """out[..., n] = sum_k A[..., k] * dequant(B)[n, k]  for a single activation vector.  reference core.py:426-504.

    One kernel launch (fused double-quant decode, optional fused bias) instead of the reference's three.
    `prefetch` (optional) is a tensor the next call will stream -- usually the packed weight of the following layer; it is
    pulled into L2 while this call computes (a hint only, see include/quantizations_b200.h).
    """
    prefetch_k = 0
    if isinstance(prefetch, tuple):  # (packed weight of the next call, its in_features): the exact form of the hint
        prefetch, prefetch_k = prefetch
    if state is None:
        raise ValueError("state cannot None. gem_4bit( ) requires the state from quantize_4bit( )")
    if A.numel() != A.shape[-1]:
        raise ValueError(
            'Dimensions of A are invalid. Must be a vector with the leading dimensions of "1", e.g. [1, 1, 2048]',
        )
    if A.dtype not in _DTYPE_CODE:
        raise NotImplementedError(f"Matmul not implemented for data type {A.dtype}")
    if B.dtype not in [torch.uint8, torch.bfloat16, torch.float16, torch.float32]:
        raise NotImplementedError(f"Matmul not implemented for data type {B.dtype}")
    Bshape = state.shape
    bout, k = Bshape[0], Bshape[1]
    if A.shape[-1] != k:
        raise ValueError(f"A has {A.shape[-1]} features but the quantised weight expects {k}")
    if out is None:
        out = torch.empty(A.shape[:-1] + (bout,), dtype=A.dtype, device=A.device)
    if not A.is_contiguous():
        A = A.contiguous()
    lib = _lib.lib()
    if A.dtype != torch.float32 and state.blocksize == 64 and not (flags & _lib.Q4_GEMV_EXACT_F32):
        # 16-bit decode path: same kernel, but with the prebuilt table image (one bulk copy instead of a per-launch build)
        import ctypes

        f = _lib.GemvFused(
            A.data_ptr(), None, None, 0.0, B.data_ptr(), ctypes.pointer(state.native_stats()), None, None, 1, state.code.data_ptr(),
            None if bias is None else bias.data_ptr(), out.data_ptr(), bout, k, state.blocksize, _DTYPE_CODE[A.dtype], flags,
            None if prefetch is None else prefetch.data_ptr(), 0 if prefetch is None else prefetch.numel() * prefetch.element_size(),
            state.lut(A.dtype).data_ptr(), *_ws_args(A.device), None, int(prefetch_k),
        )
        with _on_device(A.device):
            rc = lib.q4_gemv_4bit_fused(ctypes.byref(f), torch.cuda.current_stream(A.device).cuda_stream)
        if rc != 0:
            check(rc, "gemv_4bit")
        return out
    with _on_device(A.device):
      code = lib.q4_gemv_4bit(
        A.data_ptr(), B.data_ptr(), state.native_stats(), state.code.data_ptr(),
        None if bias is None else bias.data_ptr(), out.data_ptr(), bout, k, state.blocksize,
        _DTYPE_CODE[A.dtype], flags,
        None if prefetch is None else prefetch.data_ptr(), 0 if prefetch is None else prefetch.numel() * prefetch.element_size(),
        torch.cuda.current_stream(A.device).cuda_stream,
      )
    if code != 0:
        check(code, "gemv_4bit")
    return out


def gemm_4bit(A: Tensor, B: Tensor, state: QuantState, bias: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    """out[..., n] = sum_k A[..., k] * dequant(B)[n, k] (+ bias) for any number of rows of A, with the dequantisation fused
    into a tcgen05 tensor-core GEMM (the dense weight is never materialised).  Replaces the prefill branch of reference
    modules.py:63-64.  Requires fp16/bf16 activations, blocksize 64 and K % 64 == 0 (`fused_gemm_supported`)."""
    if state is None:
        raise ValueError("state cannot be None")
    N, K = state.shape[0], state.shape[1]
    if A.shape[-1] != K:
        raise ValueError(f"A has {A.shape[-1]} features but the quantised weight expects {K}")
    if not fused_gemm_supported(A, state):
        raise NotImplementedError("fused 4-bit GEMM needs fp16/bf16 activations, blocksize 64 and K % 64 == 0")
    A2 = A.reshape(-1, K)
    if not A2.is_contiguous():
        A2 = A2.contiguous()
    M = A2.shape[0]
    if out is None:
        out = torch.empty(A.shape[:-1] + (N,), dtype=A.dtype, device=A.device)
    with torch.cuda.device(A.device):
        check(_lib.lib().q4_gemm_4bit(A2.data_ptr(), B.data_ptr(), state.native_stats(), state.code.data_ptr(),
                                      None if bias is None else bias.data_ptr(), out.data_ptr(), M, N, K, state.blocksize,
                                      _DTYPE_CODE[A.dtype], _stream(A)), "gemm_4bit")
    return out


def fused_gemm_supported(A: Tensor, state: QuantState) -> bool:
    return (A.dtype in (torch.float16, torch.bfloat16) and state.blocksize == 64 and len(state.shape) == 2
            and state.shape[1] % 64 == 0 and A.data_ptr() % 16 == 0)


def quantize_4bit(
    A: Tensor,
    blocksize=64,
    quant_type="fp4",
    quant_storage=torch.uint8,
    compress_statistics=True,
) -> Tuple[Tensor, QuantState]:
    """Blockwise 4-bit quantisation.  reference core.py:507-578.

    Returns (packed uint8 [(n+1)//2, 1], QuantState).  With compress_statistics (the reference's only mode) absmax is
    itself quantised: offset = absmax.mean(); 8-bit blockwise (blocksize 256) of absmax - offset.  mean() and the
    subtraction are the same torch ops the reference calls, so `offset` is bit-identical on the same device.
    """
    if A.device.type != "cuda":
        raise NotImplementedError(f"Device type not supported for FP4 quantization: {A.device.type}")
    if quant_type not in _QUANT_CODE:
        raise NotImplementedError(f"4-bit quantization data type {quant_type} is not implemented.")
    if A.dtype not in _DTYPE_CODE:
        raise NotImplementedError(f"4-bit quantization is not implemented for input dtype {A.dtype}")

    n = A.numel()
    input_shape = A.shape
    blocks = -(n // -blocksize)
    absmax = torch.zeros((blocks,), device=A.device, dtype=torch.float32)
    mod = dtype2bytes[quant_storage] * 2  # KeyError for anything but uint8, like the reference (core.py:545)
    out = torch.zeros(((n + 1) // mod, 1), dtype=quant_storage, device=A.device)
    assert blocksize in _VALID_BLOCKSIZES
    A = A.contiguous()
    with torch.cuda.device(A.device):
        check(_lib.lib().q4_quantize_blockwise_4bit(A.data_ptr(), absmax.data_ptr(), out.data_ptr(), blocksize, n,
                                                    _QUANT_CODE[quant_type], _DTYPE_CODE[A.dtype], _stream(A)),
              "quantize_4bit")
    code = get_4bit_type(quant_type, device=A.device)

    if compress_statistics:
        offset = absmax.mean()
        absmax -= offset
        qabsmax, state2 = quantize_blockwise(absmax, blocksize=256)
        del absmax
        state = QuantState(absmax=qabsmax, shape=input_shape, dtype=A.dtype, blocksize=blocksize, code=code,
                           quant_type=quant_type, offset=offset, state2=state2)
    else:
        state = QuantState(absmax=absmax, shape=input_shape, dtype=A.dtype, blocksize=blocksize, code=code,
                           quant_type=quant_type)
    if blocksize == 64 and A.dtype in (torch.float16, torch.bfloat16):
        state.lut(A.dtype)  # decode table for the likely compute dtype, keyed now so that decode can run under graph capture
    return out, state


def _dequantize_4bit_into(A: Tensor, quant_state: QuantState, out: Tensor) -> Tensor:
    """Dequantise packed `A` into preallocated contiguous `out` (any of fp16/bf16/fp32), one launch."""
    with torch.cuda.device(A.device):
        check(_lib.lib().q4_dequantize_blockwise_4bit(A.data_ptr(), quant_state.native_stats(), out.data_ptr(),
                                                      quant_state.blocksize, out.numel(),
                                                      _QUANT_CODE[quant_state.quant_type], _DTYPE_CODE[out.dtype],
                                                      _stream(A)), "dequantize_4bit")
    return out


def dequantize_4bit(
    A: Tensor,
    quant_state: Optional[QuantState] = None,
    blocksize: int = 64,
    quant_type="fp4",
) -> Tensor:
    """Blockwise 4-bit dequantisation.  reference core.py:581-634.

    Returns the TRANSPOSE (a [K, N] view) of the dequantised [N, K] weight in quant_state.dtype, exactly like the
    reference (`return out.t()`, core.py:634).  The quantisation type is taken from the state.
    """
    if blocksize not in _VALID_BLOCKSIZES:
        raise ValueError(
            f"The blockwise of {blocksize} is not supported. Supported values: [2048, 4096, 1024, 512, 256, 128, 64]",
        )
    if quant_state is None:
        raise ValueError("quant_state is required")
    qt = quant_state.quant_type if quant_state.quant_type is not None else quant_type
    if qt not in _QUANT_CODE:
        raise NotImplementedError(f"4-bit quantization data type {qt} is not implemented.")
    if quant_state.dtype not in _DTYPE_CODE:
        raise NotImplementedError(f"dequantize_4bit cannot produce dtype {quant_state.dtype}")
    out = torch.empty(quant_state.shape, dtype=quant_state.dtype, device=A.device)
    _dequantize_4bit_into(A, quant_state, out)
    return out.t()


def gemv_4bit_fused(
    A: Tensor,
    B: Tensor,
    state: Optional[QuantState] = None,
    *,
    group=None,
    gate: Optional[Tensor] = None,
    rms_weight: Optional[Tensor] = None,
    rms_eps: float = 1e-5,
    residual: Optional[Tensor] = None,
    out: Optional[Tensor] = None,
    flags: int = _lib.Q4_GEMV_PDL,
    prefetch: Optional[Tensor] = None,
    allreduce=None,
    _defer=None,
) -> Tensor:
    """Decode GEMV with a transformer block's elementwise glue fused in (include/quantizations_b200.h: q4_gemv_4bit_fused):

        x_eff = A                                   (default)
              = rms_norm(A) * rms_weight            (rms_weight given)
              = silu(gate) * A                      (gate given: A is the up-projection output)
        out   = x_eff @ dequant(B)^T  (+ residual)  (residual may be `out` itself: in-place residual stream)

    `group` (a Linear4bitGroup) runs the grouped launch over its members instead of a single (B, state).
    `allreduce` (a tp.FusedAllReduce) sums the output over the tensor-parallel ranks inside the kernel's epilogue, before the
    residual is added: the row-parallel layers' all-reduce without a collective call."""
    prefetch_k = 0
    if isinstance(prefetch, tuple):  # (packed weight of the next launch, its in_features): the exact form of the hint
        prefetch, prefetch_k = prefetch
    if A.numel() != A.shape[-1]:
        raise ValueError("gemv_4bit_fused needs a single activation vector")
    if A.dtype not in (torch.float16, torch.bfloat16):
        raise NotImplementedError("fused decode GEMV needs fp16/bf16 activations")
    import ctypes

    if group is not None:
        rows, K, stats, code, packed = group.out_features, group.in_features, group._stats, group.code, group.packed
        offsets, row_end, nmat, blocksize = group._offsets, group._row_end, len(group.splits), 64
        lut = group.lut(A.dtype)
    else:
        if state is None:
            raise ValueError("state cannot be None")
        rows, K, stats, code, packed = state.shape[0], state.shape[1], state.native_stats(), state.code, B
        offsets, row_end, nmat, blocksize = None, None, 1, state.blocksize
        lut = state.lut(A.dtype)
    if A.shape[-1] != K:
        raise ValueError(f"A has {A.shape[-1]} features but the quantised weight expects {K}")
    for t in (gate, rms_weight):
        if t is not None and (t.dtype != A.dtype or t.numel() != K):
            raise ValueError("gate / rms_weight must match the activation's dtype and length")
    out_rows = rows
    if flags & _lib.Q4_GEMV_SWIGLU:  # interleaved (gate, up) pair: the launch stores silu(gate) * up
        if group is None or not getattr(group, "swiglu", False) or residual is not None or allreduce is not None:
            raise ValueError("Q4_GEMV_SWIGLU needs a SwiGLU Linear4bitGroup and no residual / allreduce")
        out_rows = rows // 2
    if residual is not None and (residual.dtype != A.dtype or residual.numel() != rows):
        raise ValueError("residual must match the output's dtype and length")
    if out is None:
        out = torch.empty(A.shape[:-1] + (out_rows,), dtype=A.dtype, device=A.device)
    elif out.numel() != out_rows or out.dtype != A.dtype:
        raise ValueError("out must match the output's dtype and length")
    f = _lib.GemvFused(
        A.data_ptr(), None if gate is None else gate.data_ptr(), None if rms_weight is None else rms_weight.data_ptr(), float(rms_eps),
        packed.data_ptr(), ctypes.pointer(stats), offsets, row_end, nmat, code.data_ptr(),
        None if residual is None else residual.data_ptr(), out.data_ptr(), rows, K, blocksize, _DTYPE_CODE[A.dtype], flags,
        None if prefetch is None else prefetch.data_ptr(), 0 if prefetch is None else prefetch.numel() * prefetch.element_size(),
        lut.data_ptr(), *_ws_args(A.device), None if allreduce is None else ctypes.pointer(allreduce.struct), int(prefetch_k),
    )
    if _defer is not None:  # gemv_4bit_chain collects the stage instead of launching it
        _defer.append((f, (A, gate, rms_weight, residual, out, lut, stats)))
        return out
    with _on_device(A.device):
        rc = _lib.lib().q4_gemv_4bit_fused(ctypes.byref(f), torch.cuda.current_stream(A.device).cuda_stream)
    if rc:
        check(rc, "gemv_4bit_fused")
    return out


def gemv_4bit_batch(A: Tensor, B: Tensor, state: QuantState, bias: Optional[Tensor] = None, out: Optional[Tensor] = None,
                    flags: int = _lib.Q4_GEMV_DEFAULT) -> Tensor:
    """2..16 tokens in ONE pass over the packed weight (include/quantizations_b200.h: q4_gemv_4bit_batch): out[.., t, n] =
    sum_k A[.., t, k] * dequant(B)[n, k] (+ bias[n]).  Raises Q4Error when the shape is not supported by that kernel."""
    N, K = state.shape[0], state.shape[1]
    M = A.numel() // A.shape[-1]
    if A.shape[-1] != K or not 1 <= M <= 16 or A.dtype not in (torch.float16, torch.bfloat16):
        raise ValueError("gemv_4bit_batch needs 1..16 fp16/bf16 tokens of the weight's in_features")
    A2 = A.reshape(M, K)
    if not A2.is_contiguous():
        A2 = A2.contiguous()
    if out is None:
        out = torch.empty(A.shape[:-1] + (N,), dtype=A.dtype, device=A.device)
    ws = gemv_workspace(A.device)
    with _on_device(A.device):
      rc = _lib.lib().q4_gemv_4bit_batch(A2.data_ptr(), B.data_ptr(), state.native_stats(), state.code.data_ptr(),
                                       None if bias is None else bias.data_ptr(), out.data_ptr(), M, N, K, state.blocksize,
                                       _DTYPE_CODE[A.dtype], flags, state.lut(A.dtype).data_ptr(), ws.data_ptr(), ws.numel(),
                                       torch.cuda.current_stream(A.device).cuda_stream)
    if rc:
        check(rc, "gemv_4bit_batch")
    return out


_chain_barriers = {}


class gemv_4bit_chain:
    """Up to four DEPENDENT decode GEMVs in one persistent launch (include/quantizations_b200.h: q4_gemv_4bit_chain):

        with gemv_4bit_chain() as ch:
            h = ch.add(a, W_o, st_o, residual=h, out=h)                            # every add() takes gemv_4bit_fused's arguments
            gu = ch.add(h, None, group=gate_up, rms_weight=ln2)                    # and returns the (not yet written) output
            h = ch.add(gu[..., I:], W_d, st_d, gate=gu[..., :I], residual=h, out=h)
        # leaving the block launches the chain

    Results equal those of the same gemv_4bit_fused calls issued one after the other."""

    def __init__(self, flags: int = _lib.Q4_GEMV_PDL):
        self.flags, self.stages = flags, []

    def __enter__(self):
        return self

    def add(self, A, B, state=None, **kw) -> Tensor:
        kw.setdefault("flags", self.flags)
        return gemv_4bit_fused(A, B, state, _defer=self.stages, **kw)

    def __exit__(self, exc_type, exc, tb):
        if exc_type is None:
            self.launch()
        return False

    def launch(self):
        import ctypes

        if not self.stages:
            return
        dev = self.stages[0][1][0].device
        stream = torch.cuda.current_stream(dev)
        lib = _lib.lib()
        stages, self.stages = self.stages, []
        with _on_device(dev):
            i = 0
            while i < len(stages):
                # the persistent ring kernel takes up to 32 dependent stages; what it does not support (Q4_ERR_SHAPE / Q4_ERR_ALIGN,
                # nothing launched) goes to the older chained launch, four stages at a time
                if _USE_RING:
                    part = stages[i:i + _lib.Q4_GEMV_RING_MAX_STAGES]
                    ws = ring_workspace(dev)
                    arr = (_lib.GemvFused * len(part))(*[f for f, _ in part])
                    rc = lib.q4_gemv_4bit_ring(arr, len(part), ws.data_ptr(), ws.numel(), stream.cuda_stream)
                    if rc == 0:
                        i += len(part)
                        continue
                    if rc not in (_lib.Q4_ERR_SHAPE, _lib.Q4_ERR_ALIGN):
                        check(rc, "gemv_4bit_ring")
                if not _USE_OLD_CHAIN:  # measured: single launches under programmatic dependent launch beat the older chained kernel
                    f = stages[i][0]
                    rc = lib.q4_gemv_4bit_fused(ctypes.byref(f), stream.cuda_stream)
                    if rc:
                        check(rc, "gemv_4bit_fused")
                    i += 1
                    continue
                part = stages[i:i + 4]
                key = (dev.index, stream.cuda_stream)
                bar = _chain_barriers.get(key)
                if bar is None:
                    bar = _chain_barriers[key] = torch.zeros(64, dtype=torch.int32, device=dev)
                arr = (_lib.GemvFused * len(part))(*[f for f, _ in part])
                rc = lib.q4_gemv_4bit_chain(arr, len(part), bar.data_ptr(), stream.cuda_stream)
                if rc:
                    check(rc, "gemv_4bit_chain")
                i += len(part)


_USE_RING = os.environ.get("Q4_GEMV_RING", "1") != "0"
_USE_OLD_CHAIN = os.environ.get("Q4_GEMV_OLD_CHAIN", "0") == "1"  # fall back to q4_gemv_4bit_chain instead of single launches
_ring_workspaces = {}


def ring_workspace(device) -> Tensor:
    """Workspace of the persistent ring GEMV (include/quantizations_b200.h: q4_gemv_4bit_ring): zeroed ONCE, one per device -- it
    carries the exchange words and the per-CTA launch epochs, so it must neither be re-zeroed by a replayed CUDA graph nor be used
    by two streams at once.  Allocate it (one eager call) before capturing a graph that contains ring launches."""
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    ws = _ring_workspaces.get(idx)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            raise Q4Error("the ring GEMV workspace must exist before stream capture: run the step once eagerly first")
        ws = _ring_workspaces[idx] = torch.zeros(_lib.Q4_GEMV_RING_WS_BYTES, dtype=torch.uint8, device=torch.device("cuda", idx))
    return ws


def decode_attention(qkv: Tensor, cos: Tensor, sin: Tensor, k_cache: Tensor, v_cache: Tensor, pos: Tensor, nh: int, nkv: int,
                     out: Optional[Tensor] = None, flags: int = _lib.Q4_GEMV_PDL) -> Tensor:
    """Decode-step glue in one launch (include/quantizations_b200.h: q4_decode_attention): RoPE on the new token's q / k,
    KV-cache append at `pos` (a device scalar), attention of the one query over positions [0, pos].
    qkv [.., (nh + 2*nkv) * 128] as the grouped q/k/v GEMV leaves it; cos / sin [max_len, 64]; caches [nkv, max_len, 128]."""
    hd = k_cache.shape[-1]
    if qkv.dtype not in (torch.float16, torch.bfloat16) or qkv.numel() != (nh + 2 * nkv) * hd:
        raise ValueError("decode_attention needs one fp16/bf16 token of q|k|v")
    for t in (cos, sin, k_cache, v_cache):
        if t.dtype != qkv.dtype or not t.is_contiguous():
            raise ValueError("cos / sin / caches must be contiguous and of the activation dtype")
    if pos.dtype != torch.int64 or pos.numel() != 1:
        raise ValueError("pos must be a one-element int64 device tensor")
    if out is None:
        out = torch.empty(qkv.shape[:-1] + (nh * hd,), dtype=qkv.dtype, device=qkv.device)
    with _on_device(qkv.device):
      rc = _lib.lib().q4_decode_attention(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(),
                                        pos.data_ptr(), out.data_ptr(), nh, nkv, hd, k_cache.shape[-2], _DTYPE_CODE[qkv.dtype], flags,
                                        torch.cuda.current_stream(qkv.device).cuda_stream)
    if rc:
        check(rc, "decode_attention")
    return out


_ARGMAX_WS = {}


def argmax(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """Index of the largest element of a contiguous fp16/bf16/fp32 vector as a one-element int64 tensor (greedy sampling of the
    decode harness; include/quantizations_b200.h: q4_argmax).  One ~3 us launch, capturable in a CUDA graph."""
    if x.dtype not in _DTYPE_CODE or not x.is_contiguous():
        raise ValueError("argmax needs a contiguous fp16 / bf16 / fp32 tensor")
    key = (x.device.type, x.device.index)
    ws = _ARGMAX_WS.get(key)
    if ws is None:
        ws = _ARGMAX_WS[key] = torch.zeros(4096, dtype=torch.uint8, device=x.device)
    if out is None:
        out = torch.empty(1, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().q4_argmax(x.data_ptr(), x.numel(), _DTYPE_CODE[x.dtype], out.data_ptr(), ws.data_ptr(), _stream(x)), "argmax")
    return out
