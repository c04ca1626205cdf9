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
        # ---- the same exchange inside the persistent ring kernel: a tensor-parallel decoder-layer chain (row-parallel o_proj and
        #      down_proj shards with the all-reduce in their epilogues, column-parallel gate/up shard with SwiGLU in pair mode) in ONE
        #      launch against the same stages as single launches, repeated and replayed
        I2 = inter // world
        mk = lambda n, k, seed: q.Linear4bit(k, n, bias=False, compute_dtype=torch.bfloat16, quant_type="nf4", device="meta")
        def shard_lin(N, K, kind, seed):
            gg = torch.Generator(device=dev).manual_seed(seed)
            W = (torch.randn(N, K, device=dev, generator=gg) * K ** -0.5).to(torch.bfloat16)
            sh = tp.shard_weight(W, kind, rank, world)
            lin = mk(sh.shape[0], sh.shape[1], seed)
            lin.weight = q.Params4bit(sh, requires_grad=False, quant_type="nf4", module=lin).to(dev)
            return lin
        o_l, gate_l, up_l, down_l = shard_lin(hidden, hidden, "row", 31), shard_lin(inter, hidden, "col", 32), shard_lin(inter, hidden, "col", 33), shard_lin(hidden, inter, "row", 34)
        gu = q.Linear4bitGroup([gate_l, up_l])
        ln2 = (1 + 0.1 * torch.randn(hidden, device=dev, generator=torch.Generator(device=dev).manual_seed(5))).to(torch.bfloat16)
        a_in = torch.randn(1, 1, hidden // world, device=dev, dtype=torch.bfloat16, generator=torch.Generator(device=dev).manual_seed(50 + rank))
        h0 = torch.randn(1, 1, hidden, device=dev, dtype=torch.bfloat16, generator=torch.Generator(device=dev).manual_seed(51))

        h = h0.clone()
        q.gemv_4bit_fused(a_in, o_l.weight.data, o_l.weight.quant_state, residual=h, out=h, allreduce=ar)
        g_ref = q.gemv_4bit_fused(h, None, group=gu, rms_weight=ln2)
        q.gemv_4bit_fused(g_ref[..., I2:], down_l.weight.data, down_l.weight.quant_state, gate=g_ref[..., :I2], residual=h, out=h, allreduce=ar)
        h_ref = h.clone()

        h = h0.clone()
        g_u = torch.empty(1, 1, 2 * I2, device=dev, dtype=torch.bfloat16)

        def ring_chain():
            st = []
            q.gemv_4bit_fused(a_in, o_l.weight.data, o_l.weight.quant_state, residual=h, out=h, allreduce=ar, _defer=st)
            q.gemv_4bit_fused(h, None, group=gu, rms_weight=ln2, out=g_u, _defer=st)
            q.gemv_4bit_fused(g_u[..., I2:], down_l.weight.data, down_l.weight.quant_state, gate=g_u[..., :I2], residual=h, out=h, allreduce=ar, _defer=st)
            ws = q.core.ring_workspace(dev)
            arr = (q._lib.GemvFused * len(st))(*[f for f, _ in st])
            rc = q._lib.lib().q4_gemv_4bit_ring(arr, len(st), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
            assert rc == 0, rc
            return st

        for rep in range(3):
            h.copy_(h0)
            keep = ring_chain()
            torch.cuda.synchronize()
            err = ((h.float() - h_ref.float()).abs().max() / h_ref.float().abs().max()).item()
            assert err <= 1e-2, ("ring tp chain", rep, err)
            assert ((g_u.float() - g_ref.float()).abs().max() / g_ref.float().abs().max()).item() <= 1e-2
        gathered = [torch.empty_like(h) for _ in range(world)]
        dist.all_gather(gathered, h)
        assert all(torch.equal(gathered[0], t) for t in gathered), "ring: ranks hold different bits"
        h.copy_(h0)
        gr = graphs.capture(ring_chain)
        for _ in range(3):
            h.copy_(h0)
            gr.replay()
            torch.cuda.synchronize()
            err = ((h.float() - h_ref.float()).abs().max() / h_ref.float().abs().max()).item()
            assert err <= 1e-2, ("ring tp chain replay", err)
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
