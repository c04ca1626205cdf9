"""Multi-process host logic of the tensor-parallel path (quantizations_b200/tp.py) on CPU: world_size 2, gloo.

Each rank derives its shard of a full weight from the shared seed, quantises it with the ORACLE (there is no CPU product
path), runs its part of a gate/up -> down MLP and the partials are combined with one all-reduce -- the result must equal the
unsharded computation up to the quantisation error of the row-parallel split and exactly match a per-shard recomputation."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import q4_oracle as orc
        from quantizations_b200 import tp

        h, inter = 256, 512
        g = torch.Generator().manual_seed(0)
        Wg = torch.randn(inter, h, generator=g) * 0.05
        Wu = torch.randn(inter, h, generator=g) * 0.05
        Wd = torch.randn(h, inter, generator=g) * 0.05
        x = torch.randn(1, 1, h, generator=g)

        def qdq(W):  # quantise + dequantise a shard with the oracle (NF4, double-quant)
            st = orc.quantize_4bit(W.numpy(), 64, "nf4")
            return torch.from_numpy(orc.dequantize_4bit(st, "float32"))

        ws = {n: qdq(tp.shard_weight(W, tp.kind_of(n), rank, world)) for n, W in (("gate_proj", Wg), ("up_proj", Wu), ("down_proj", Wd))}
        assert ws["gate_proj"].shape == (inter // world, h) and ws["down_proj"].shape == (h, inter // world)
        act = torch.nn.functional.silu(x @ ws["gate_proj"].t()) * (x @ ws["up_proj"].t())      # column-parallel: stays sharded
        y = act @ ws["down_proj"].t()                                                          # row-parallel partial
        partial = y.clone()
        tp.combine_output(y, "row")
        # exact: the all-reduce is the sum of both ranks' partials
        gathered = [torch.zeros_like(partial) for _ in range(world)]
        dist.all_gather(gathered, partial)
        assert torch.allclose(y, sum(gathered), atol=1e-6)
        # close to the unsharded quantised MLP (row-parallel shards have their own absmax statistics)
        full = (torch.nn.functional.silu(x @ qdq(Wg).t()) * (x @ qdq(Wu).t())) @ qdq(Wd).t()
        err = (y - full).abs().max().item() / full.abs().max().item()
        assert err < 2e-2, err
        # the slices tile the full weight exactly
        cols = [torch.zeros(inter // world, h) for _ in range(world)]
        dist.all_gather(cols, tp.shard_weight(Wg, "col", rank, world))
        assert torch.equal(torch.cat(cols, 0), Wg)
        rows = [torch.zeros(h, inter // world) for _ in range(world)]
        dist.all_gather(rows, tp.shard_weight(Wd, "row", rank, world))
        assert torch.equal(torch.cat(rows, 1), Wd)
        assert tp.shard_input(act.new_zeros(1, 1, inter), "row", rank, world).shape[-1] == inter // world
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        ret[rank] = repr(e)
    finally:
        dist.destroy_process_group()


def test_tp_mlp_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_shapes_and_errors():
    from quantizations_b200 import tp

    assert tp.shard_shape(4096, 4096, "col", 8) == (512, 4096)
    assert tp.shard_shape(4096, 14336, "row", 2) == (4096, 7168)
    assert tp.kind_of("model.layers.3.self_attn.o_proj") == "row" and tp.kind_of("L0.up_proj") == "col"
    with pytest.raises(ValueError):
        tp.shard_shape(1000, 4096, "col", 3)
    assert tp.shard_shape(4096, 14336, "row", 8) == (4096, 1792)  # 28 whole quantisation blocks per rank
    with pytest.raises(ValueError):
        tp.shard_shape(4096, 4096 + 64, "row", 8)  # 65 blocks do not split over 8 ranks
