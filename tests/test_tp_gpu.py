"""Tensor-parallel decode on real GPUs (needs >= 2): the all-reduce fused into the row-parallel GEMV's epilogue
(q4_allreduce_t, quantizations_b200/tp.py: FusedAllReduce) against NCCL's all-reduce of the same partials."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import quantizations_b200 as q
        from quantizations_b200 import graphs, tp

        hidden, inter = 4096, 14336
        ar = tp.FusedAllReduce(hidden, device=dev)
        layers = []
        for name, (N, K) in (("o_proj", (hidden, hidden)), ("down_proj", (hidden, inter))):
            g = torch.Generator(device=dev).manual_seed(7 + len(layers))
            W = (torch.randn(N, K, device=dev, generator=g) * 0.02).to(torch.bfloat16)   # same full weight on every rank
            shard = tp.shard_weight(W, "row", rank, world)
            packed, st = q.quantize_4bit(shard, quant_type="nf4")
            layers.append((packed, st, K // world))
        res = torch.randn(1, 1, hidden, device=dev, dtype=torch.bfloat16, generator=torch.Generator(device=dev).manual_seed(99))
        worst = 0.0
        for rep in range(6):  # alternating layers: both halves of the double buffer, epochs beyond the first
            packed, st, k = layers[rep % 2]
            x = torch.randn(1, 1, k, device=dev, dtype=torch.bfloat16)
            part = q.gemv_4bit(x, packed, state=st).float()
            dist.all_reduce(part)
            want = (part + res.float())
            got = q.gemv_4bit_fused(x, packed, st, residual=res, allreduce=ar).float()
            worst = max(worst, ((got - want).abs().max() / want.abs().max()).item())
        assert worst <= 1e-2, worst
        # replayed CUDA graph: the device-side epochs keep the ranks in step
        packed, st, k = layers[0]
        x = torch.randn(1, 1, k, device=dev, dtype=torch.bfloat16)
        out = torch.empty(1, 1, hidden, device=dev, dtype=torch.bfloat16)
        g = graphs.capture(lambda: q.gemv_4bit_fused(x, packed, st, residual=res, out=out, allreduce=ar))
        part = q.gemv_4bit(x, packed, state=st).float()
        dist.all_reduce(part)
        for _ in range(5):
            out.zero_()
            g.replay()
            torch.cuda.synchronize()
            err = ((out.float() - (part + res.float())).abs().max() / part.abs().max()).item()
            assert err <= 1e-2, err
        # every rank holds the same bits (the sum is taken in rank order everywhere)
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        assert all(torch.equal(gathered[0], t) for t in gathered)
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback

        ret[rank] = traceback.format_exc()
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_fused_allreduce_matches_nccl():
    import torch.multiprocessing as mp

    world = 2
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, 29533, ret), nprocs=world, join=True)
        assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)
