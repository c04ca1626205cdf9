"""Tensor-parallel sharding of Linear4bit weights, one process per GPU, torch.distributed for the exchange.

Megatron-style split inside one NVSwitch domain:
  column-parallel (q/k/v/gate/up): rank r owns output rows [r*N/p, (r+1)*N/p) -- a contiguous byte range of the packed weight;
                                   no communication, the output stays sharded.
  row-parallel    (o/down):        rank r owns input columns [r*K/p, (r+1)*K/p); it quantises that [N, K/p] slice as its own
                                   tensor (own absmax / nested statistics), computes a partial [.., N] and the partials are
                                   summed with ONE all-reduce of a hidden-sized vector per layer.
Every rank derives its slice from the same full weight (shared seed or checkpoint), so no weight is ever sent.
"""
from __future__ import annotations

import torch

COLUMN_PARALLEL = ("q_proj", "k_proj", "v_proj", "gate_proj", "up_proj")
ROW_PARALLEL = ("o_proj", "down_proj")


def kind_of(name: str) -> str:
    leaf = name.split(".")[-1]
    if leaf in COLUMN_PARALLEL:
        return "col"
    if leaf in ROW_PARALLEL:
        return "row"
    raise ValueError(f"{name}: not a tensor-parallel projection")


def shard_shape(n: int, k: int, kind: str, world: int):
    if kind == "col":
        if n % world:
            raise ValueError(f"out_features {n} not divisible by tp={world}")
        return n // world, k
    if k % (world * 64):
        raise ValueError(f"in_features {k} must be a multiple of 64*tp={64 * world} (quantisation blocks must not straddle ranks)")
    return n, k // world


def shard_weight(W: torch.Tensor, kind: str, rank: int, world: int) -> torch.Tensor:
    """Slice of the full [N, K] weight owned by `rank` (contiguous copy, ready for quantize_4bit)."""
    n, k = shard_shape(W.shape[0], W.shape[1], kind, world)
    if kind == "col":
        return W[rank * n:(rank + 1) * n].contiguous()
    return W[:, rank * k:(rank + 1) * k].contiguous()


def shard_input(x: torch.Tensor, kind: str, rank: int, world: int) -> torch.Tensor:
    """Activation slice a row-parallel layer consumes (its producer, a column-parallel layer, already emits exactly this)."""
    if kind == "col":
        return x
    k = x.shape[-1] // world
    return x[..., rank * k:(rank + 1) * k]


def combine_output(y: torch.Tensor, kind: str, group=None) -> torch.Tensor:
    """Row-parallel partials are summed in place with one all-reduce; column-parallel outputs stay sharded."""
    if kind == "row":
        torch.distributed.all_reduce(y, group=group)
    return y


class FusedAllReduce:
    """Exchange area for the all-reduce fused into the row-parallel decode GEMV (include/quantizations_b200.h: q4_allreduce_t).

    One symmetric allocation per rank (torch.distributed._symmetric_memory: every rank can address every peer's copy over
    NVLink), zeroed once; `rows` is the output size of the row-parallel layers (the hidden size).  Pass the object as
    `allreduce=` to core.gemv_4bit_fused: the kernel then leaves the SUM over ranks (+ residual) in `out` -- no NCCL call, no
    extra launch.  NCCL (`combine_output`) stays the reference the tests compare against."""

    def __init__(self, rows: int, group=None, device=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib

        group = dist.group.WORLD if group is None else group
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.rows, self.world, self.rank = int(rows), dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("the fused all-reduce covers one NVSwitch domain: at most 8 ranks")
        nbytes = _lib.ar_bytes(self.rows)
        self.area = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.area.zero_()
        self.handle = symm_mem.rendezvous(self.area, group)
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every rank's area is zeroed before anyone's kernel can write into it
        self.struct = _lib.AllReduce(self.handle.buffer_ptrs_dev, self.world, self.rank, self.rows)
