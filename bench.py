#!/usr/bin/env python
"""bench.py -- the contract benchmark (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): the NF4 + double-quant Linear4bit stack of Llama-3-8B at batch 1 -- per layer
q_proj 4096x4096, k_proj/v_proj 1024x4096, o_proj 4096x4096, gate_proj/up_proj 14336x4096, down_proj 4096x14336,
x 32 layers = 224 GEMVs per step over 3.5 GB of DISTINCT packed weights (random-init, synthetic), far larger than the
126 MB L2, so every step streams every byte from HBM.  A step = one decode token's worth of Linear4bit forwards.

metric / unit   BASELINE.json's: 4-bit GEMV HBM GB/s = algorithmic bytes of the step / step time
value           device-timed (CUDA events around K graph replays), inputs resident in HBM, launches through the C ABI
e2e             the same step through the public Python API (Linear4bit.forward under quantizations_b200.graphs), with
                the activations copied from pinned host memory and the outputs copied back inside the timed region
roofline        the GEMV kernel: algorithmic bytes per launch / average launch duration vs MEASURED_PEAKS.json hbm_gbs
                (roofline.traffic: DRAM bytes per launch from the committed ncu launch list, profiles/gemv_traffic.json)
blockwise       the path's other two HBM-bound kernels, quantize and dequantize of a 14336x4096 weight (kernel-only, N = 1)
decode          whole-model batch-1 greedy decode tok/s in the quantizations_b200.llama harness (CUDA-graph step); with
                --gpus N the model is tensor-parallel over the ranks (all-reduces fused into the GEMV epilogues)
cpu_baseline    oracle/q4_torch_cpu.py (torch-CPU dequantize -> matmul port) on the box's host cores, bounded sample
--impl reference   the reference's own kernels (oracle/_ref, built from /root/reference for sm_100a) driven with the
                reference's gemv_4bit call sequence (core.py:467-499) on the same workload
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "4-bit GEMV HBM GB/s (% of ~8 TB/s); Llama-3-8B bs=1 decode tok/s"
LLAMA3_8B = dict(hidden=4096, inter=14336, kv=1024, layers=32)
LLAMA3_70B = dict(hidden=8192, inter=28672, kv=1024, layers=80)


def layer_shapes(cfg, tp: int = 1):
    """(name, N, K, parallel) of one decoder layer's Linear4bit modules, Megatron-style TP split."""
    h, i, kv = cfg["hidden"], cfg["inter"], cfg["kv"]
    return [
        ("q_proj", h // tp, h, "col"), ("k_proj", kv // tp, h, "col"), ("v_proj", kv // tp, h, "col"),
        ("o_proj", h, h // tp, "row"), ("gate_proj", i // tp, h, "col"), ("up_proj", i // tp, h, "col"),
        ("down_proj", h, i // tp, "row"),
    ]  # fmt: skip


def algo_bytes(N: int, K: int, xbytes: int = 2, nested: bool = True) -> int:
    """packed + absmax (+ nested tables) + code + x + y."""
    n = N * K
    b = n // 2 + 64 + xbytes * K + xbytes * N
    b += (n // 64 + 4 * -(-n // 16384) + 1024 + 4) if nested else 4 * (n // 64)
    return b


# ------------------------------------------------------------------------------------------------ clocks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workload


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def build_stack(cfg, tp, rank, device, dtype, quant_type, layers, q):
    """Random-init, quantised Linear4bit stack.  Every rank draws the full weight from a shared seed and quantises its own
    slice (row-parallel shards are quantised as their own [N, K/p] tensors)."""
    mods = []
    g = torch.Generator(device=device)
    for layer in range(layers):
        for j, (name, N, K, par) in enumerate(layer_shapes(cfg, tp)):
            g.manual_seed(1000 * layer + j)
            fullN, fullK = (N * tp, K) if par == "col" else (N, K * tp)
            # N(0, 1/in_features): every GEMV keeps the activation's scale, so the outputs of a step that really feeds each Linear's
            # output into the next stay finite over 128 dependent GEMVs
            if tp == 1:
                W = torch.randn(N, K, device=device, dtype=torch.float32, generator=g).mul_(fullK ** -0.5).to(dtype)
            else:
                Wf = torch.randn(fullN, fullK, device=device, dtype=torch.float32, generator=g).mul_(fullK ** -0.5).to(dtype)
                W = (Wf[rank * N:(rank + 1) * N] if par == "col" else Wf[:, rank * K:(rank + 1) * K]).contiguous()
                del Wf
            lin = q.Linear4bit(K, N, bias=False, compute_dtype=dtype, compress_statistics=True, quant_type=quant_type,
                               device="meta")
            lin.weight = q.Params4bit(W, requires_grad=False, quant_type=quant_type, module=lin).to(device)
            lin.parallel, lin.name_ = par, name
            mods.append(lin)
            del W
    return mods


def run_ours(args, rank, world, device):
    import quantizations_b200 as q
    from quantizations_b200 import _lib, graphs

    L = _lib.lib()
    dtype = torch.bfloat16
    cfg = LLAMA3_8B if args.model == "llama3-8b" else LLAMA3_70B
    layers = args.layers or cfg["layers"]
    tp = world
    mods = build_stack(cfg, tp, rank, device, dtype, "nf4", layers, q)
    # launch units: with grouping (default) q/k/v and gate/up of a layer are one grouped launch each (Linear4bitGroup)
    units = []
    if args.group:
        for l0 in range(0, len(mods), 7):
            qm, km, vm, om, gm, um, dm = mods[l0:l0 + 7]
            for u in (q.Linear4bitGroup([qm, km, vm]), om, q.Linear4bitGroup([gm, um]), dm):
                units.append(u)
        for u, name, par in zip(units, ["qkv_proj", "o_proj", "gate_up_proj", "down_proj"] * layers, ["col", "row", "col", "row"] * layers):
            u.name_, u.parallel = name, par
    else:
        units = list(mods)
    step_bytes_local = sum(algo_bytes(m.out_features, m.in_features) for m in mods)
    step_bytes = step_bytes_local * world  # every rank streams its own shard
    packed_bytes = sum(m.weight.numel() for m in mods)

    h, inter = cfg["hidden"], cfg["inter"]
    torch.manual_seed(1)
    x_in = {h: torch.randn(1, 1, h, device=device, dtype=dtype)}  # the token's hidden state entering layer 0; everything else is produced by the step
    outs = {}
    for m in units:
        outs.setdefault((m.name_, m.out_features), torch.empty(1, 1, m.out_features, device=device, dtype=dtype))
    comm = None
    fused_ar = None
    if world > 1:
        import torch.distributed as dist

        comm = dist
        if not args.nccl_allreduce:  # the product path: all-reduce inside the row-parallel GEMV's epilogue over NVLink peer memory
            from quantizations_b200 import tp as tpmod

            fused_ar = tpmod.FusedAllReduce(cfg["hidden"], device=device)

    # ---- tensor-parallel correctness, before anything is timed: every row-parallel unit once through the epilogue all-reduce
    #      (the product path) and once as a plain GEMV of the same shard + NCCL all-reduce of the partials
    tp_check = None
    if fused_ar is not None:
        worst = 0.0
        row_units = [m for m in units if m.parallel == "row"]
        picked = row_units[:4] + row_units[-2:]  # both layer kinds, both halves of the double buffer, epochs beyond the first
        for rep, m in enumerate(picked):
            xk = torch.randn(1, 1, m.in_features, device=device, dtype=dtype, generator=torch.Generator(device=device).manual_seed(77 + rep + 1000 * rank))
            st = m.weight.quant_state
            part = q.gemv_4bit(xk, m.weight.data, state=st).float()
            comm.all_reduce(part)
            got = q.gemv_4bit_fused(xk, m.weight.data, st, allreduce=fused_ar)
            err = ((got.float() - part).abs().max() / part.abs().max()).item()
            worst = max(worst, err)
            if not err <= 1e-2:
                raise RuntimeError(f"tp_check: fused all-reduce of {m.name_} differs from NCCL's by {err} (rank {rank})")
            gathered = [torch.empty_like(got) for _ in range(world)]
            comm.all_gather(gathered, got)
            if not all(torch.equal(gathered[0], t) for t in gathered):
                raise RuntimeError(f"tp_check: ranks hold different bits after the fused all-reduce of {m.name_}")
        tp_check = {"status": "ok", "units": len(picked), "max_rel_err_vs_nccl": worst, "bit_identical_across_ranks": True}

    # ---- what one all-reduce costs (SURVEY 8e "Reporting"): the same row-parallel GEMV of a shard, 32 launches per CUDA-graph replay, with
    #      the all-reduce fused into its epilogue and without; and NCCL's all-reduce of the same 8 KB vector, stream-ordered, for comparison
    ar_latency = None
    if fused_ar is not None and not getattr(args, "no_ar_latency", False):
        from quantizations_b200 import graphs as _graphs

        ar_latency = {}
        row_units = [m for m in units if m.parallel == "row"]
        for m in row_units[:2]:  # o_proj and down_proj shards of the first layer
            st = m.weight.quant_state
            xk = torch.randn(1, 1, m.in_features, device=device, dtype=dtype)
            yk = torch.empty(1, 1, m.out_features, device=device, dtype=dtype)

            def timed(kw):
                def body():
                    for _ in range(32):
                        q.gemv_4bit_fused(xk, m.weight.data, st, out=yk, **kw)
                g = _graphs.capture(body)
                comm.barrier()
                g.replay()
                torch.cuda.synchronize(device)
                comm.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    g.replay()
                e1.record()
                torch.cuda.synchronize(device)
                t = torch.tensor([e0.elapsed_time(e1) * 1e3 / (5 * 32)], device=device)
                comm.all_reduce(t, op=comm.ReduceOp.MAX)
                return t.item()

            with_ar, without = timed({"allreduce": fused_ar}), timed({})
            ar_latency[f"{m.name_} {m.out_features}x{m.in_features}"] = {"gemv_with_fused_allreduce_us": round(with_ar, 3), "gemv_alone_us": round(without, 3),
                                                                         "allreduce_us": round(with_ar - without, 3)}
        v = torch.randn(cfg["hidden"], device=device, dtype=dtype)
        for _ in range(5):
            comm.all_reduce(v)
        torch.cuda.synchronize(device)
        comm.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(64):
            comm.all_reduce(v)
        e1.record()
        torch.cuda.synchronize(device)
        t = torch.tensor([e0.elapsed_time(e1) * 1e3 / 64], device=device)
        comm.all_reduce(t, op=comm.ReduceOp.MAX)
        ar_latency["nccl_all_reduce_us"] = round(t.item(), 3)
        ar_latency["note"] = ("per-launch times, max over ranks, CUDA events around 5 replays of a 32-launch graph (L2-resident shard: latency, not bandwidth); "
                              f"NCCL: 64 stream-ordered all_reduce calls of {cfg['hidden']} bf16 values")

    pdl = _lib.Q4_GEMV_PDL if args.pdl else 0
    # ---- data flow of a step: a decode token's Linear4bit forwards are DEPENDENT -- every Linear consumes what the previous one
    #      produced (q/k/v <- the previous layer's down_proj, o_proj <- the first in_features values of q/k/v [attention is not part
    #      of this stack], gate/up <- o_proj, down_proj <- the first in_features values of gate/up), so a launch can never start its
    #      arithmetic before its predecessor's last row exists.  All launch paths below run this same flow.
    per_layer = len(units) // layers
    first_in = {"qkv_proj", "q_proj", "k_proj", "v_proj"}
    qkey = ("qkv_proj", (h + 2 * cfg["kv"]) // tp) if args.group else ("q_proj", h // tp)
    gkey = ("gate_up_proj", 2 * inter // tp) if args.group else ("gate_proj", inter // tp)

    def source_of(i, m, bufs, x0):
        """the tensor unit i reads: its first in_features values"""
        if m.name_ in first_in:
            t = x0 if i // per_layer == 0 else bufs[("down_proj", h)]
        elif m.name_ == "o_proj":
            t = bufs[qkey]
        elif m.name_ in ("gate_up_proj", "gate_proj", "up_proj"):
            t = bufs[("o_proj", h)]
        else:
            t = bufs[gkey]
        return t[..., :m.in_features]

    # A decoder layer's Linear4bit calls form a small DAG: q/k/v share their input and are independent of each other, so do
    # gate/up; o_proj and down_proj each depend on what precedes them.  The step is launched exactly like that: independent
    # GEMVs go to parallel streams (parallel branches of the CUDA graph) with Q4_GEMV_SHARE_SM so they are co-resident on
    # every SM; dependent ones follow in stream order with programmatic dependent launch.
    side = [torch.cuda.Stream(device=device) for _ in range(2)] if args.branches else []
    branch_of = {"q_proj": 0, "k_proj": 1, "v_proj": 2, "o_proj": 0, "gate_proj": 0, "up_proj": 1, "down_proj": 0,
                 "qkv_proj": 0, "gate_up_proj": 0}
    forks = {"q_proj", "gate_proj"}            # a parallel section starts here ...
    joins = {"o_proj", "down_proj"}            # ... and has ended before these

    def run_stack(launch):
        """walk the stack in model order, placing each Linear on its branch; `launch(i, m, flags)` issues the work"""
        main = torch.cuda.current_stream()
        for i, m in enumerate(units):
            b = branch_of[m.name_] if side else 0
            if side and m.name_ in forks:
                for st in side:
                    st.wait_stream(main)
            if side and m.name_ in joins:
                for st in side:
                    main.wait_stream(st)
            shared = side and m.name_ not in joins
            flags = pdl | (_lib.Q4_GEMV_SHARE_SM if shared else 0)
            if b == 0:
                launch(i, m, flags)
            else:
                with torch.cuda.stream(side[b - 1]):
                    launch(i, m, flags)
        for st in side:
            main.wait_stream(st)

    def packed_of(u):
        return u.packed if isinstance(u, q.Linear4bitGroup) else u.weight

    fused_args = {}
    ws_ptr, ws_bytes = q.core._ws_args(device)  # (None, 0) unless Q4_GEMV_TC=1: the mma.sync kernel is the default

    def launch_struct(i, m, flags):
        """argument struct of unit i (q4_gemv_fused_t, include/quantizations_b200.h) with the prebuilt table image; built once"""
        key = (i, flags)
        f = fused_args.get(key)
        if f is None:
            out = outs[(m.name_, m.out_features)]
            nxt = packed_of(units[(i + 1) % len(units)]) if args.prefetch else None  # the following launch's packed weight
            npt, nby = (None, 0) if nxt is None else (nxt.data_ptr(), nxt.numel())
            nk = units[(i + 1) % len(units)].in_features if args.prefetch else 0  # exact hint: the next launch's in_features
            if isinstance(m, q.Linear4bitGroup):
                f = _lib.GemvFused(source_of(i, m, outs, x_in[h]).data_ptr(), None, None, 0.0, m.packed.data_ptr(), ctypes.pointer(m._stats), m._offsets,
                                   m._row_end, len(m.splits), m.code.data_ptr(), None, out.data_ptr(), m.out_features, m.in_features, 64,
                                   _lib.Q4_BF16, flags, npt, nby, m.lut(dtype).data_ptr(), ws_ptr, ws_bytes, prefetch_K=nk)
            else:
                st = m.weight.quant_state
                ar = ctypes.pointer(fused_ar.struct) if (fused_ar is not None and m.parallel == "row") else None
                f = _lib.GemvFused(source_of(i, m, outs, x_in[h]).data_ptr(), None, None, 0.0, m.weight.data_ptr(), ctypes.pointer(st.native_stats()), None,
                                   None, 1, st.code.data_ptr(), None, out.data_ptr(), m.out_features, m.in_features, st.blocksize,
                                   _lib.Q4_BF16, flags, npt, nby, st.lut(dtype).data_ptr(), ws_ptr, ws_bytes, ar, nk)
            fused_args[key] = f
        return f

    def launch_cabi(i, m, flags):
        f = launch_struct(i, m, flags)
        rc = L.q4_gemv_4bit_fused(ctypes.byref(f), torch.cuda.current_stream().cuda_stream)
        if rc:
            _lib.check(rc, "q4_gemv_4bit_fused")
        if comm is not None and fused_ar is None and m.parallel == "row":
            comm.all_reduce(outs[(m.name_, m.out_features)])

    chain_arrays = []
    chain_bar = torch.zeros(64, dtype=torch.int32, device=device)

    def build_chains():
        """Chained launches (q4_gemv_4bit_chain): the decode data flow is attention(L) -> [o_L -> gate/up_L -> down_L -> qkv_{L+1}]
        -> attention(L+1), so the stack is one single launch (qkv_0) plus one persistent launch of up to four dependent GEMVs per
        layer.  Stage argument structs are the same ones the per-launch path uses."""
        flags = pdl
        for i, m in enumerate(units):
            launch_struct(i, m, flags)
        order = [[0]]
        per = 4
        nl = len(units) // per
        for l in range(nl):
            grp = [l * per + 1, l * per + 2, l * per + 3]
            if l + 1 < nl:
                grp.append((l + 1) * per)
            order.append(grp)
        for grp in order:
            arr = (_lib.GemvFused * len(grp))(*[fused_args[(i, flags)] for i in grp])
            chain_arrays.append((arr, len(grp)))

    def step_chained():
        stream = torch.cuda.current_stream().cuda_stream
        for arr, n in chain_arrays:
            rc = L.q4_gemv_4bit_chain(arr, n, chain_bar.data_ptr(), stream)
            if rc:
                _lib.check(rc, "q4_gemv_4bit_chain")

    use_chain = args.chain and args.group and not side and (comm is None or fused_ar is not None) and len(units) % 4 == 0
    # the persistent ring kernel (q4_gemv_4bit_ring): up to 32 dependent GEMVs (eight decoder layers) per launch
    # (tensor-parallel: measured slower than single launches at tp2 -- 1.03 vs 0.94 ms/step -- so there it is opt-in: --ring-tp)
    use_ring = args.ring and not use_chain and args.group and not side and (comm is None or (fused_ar is not None and args.ring_tp))
    ring_arrays = []
    ring_state = {"launches": 0, "fallbacks": 0}

    def build_rings():
        ws = q.core.ring_workspace(device)
        n_max = _lib.Q4_GEMV_RING_MAX_STAGES
        for c0 in range(0, len(units), n_max):
            idx = list(range(c0, min(c0 + n_max, len(units))))
            arr = (_lib.GemvFused * len(idx))(*[launch_struct(i, units[i], pdl) for i in idx])
            ring_arrays.append((arr, idx, ws))

    def step_ring():
        stream = torch.cuda.current_stream().cuda_stream
        for arr, idx, ws in ring_arrays:
            rc = L.q4_gemv_4bit_ring(arr, len(idx), ws.data_ptr(), ws.numel(), stream)
            if rc in (_lib.Q4_ERR_SHAPE, _lib.Q4_ERR_ALIGN):  # nothing was launched: the single-GEMV kernel takes these stages
                ring_state["fallbacks"] += 1
                for i in idx:
                    launch_cabi(i, units[i], pdl)
            elif rc:
                _lib.check(rc, "q4_gemv_4bit_ring")
            else:
                ring_state["launches"] += 1

    def step_cabi():
        """one decode token's worth of Linear4bit GEMVs straight through the C ABI"""
        if use_ring:
            if not ring_arrays:
                build_rings()
            step_ring()
        elif use_chain:
            if not chain_arrays:
                build_chains()
            step_chained()
        else:
            run_stack(launch_cabi)

    # ---- value: CUDA-graph replay of the step, device-timed
    n0 = _lib.launch_count()
    step_cabi()
    launches_per_step = _lib.launch_count() - n0
    torch.cuda.synchronize()
    ring_check = None
    if use_ring and ring_state["fallbacks"] == 0:
        # the ring launches against one single-GEMV launch per stage on the same data flow (all ranks, before anything is timed)
        got = {k: v.clone() for k, v in outs.items()}
        run_stack(launch_cabi)
        torch.cuda.synchronize()
        worst = 0.0
        for k, v in outs.items():
            ref = v.float()
            worst = max(worst, ((got[k].float() - ref).abs().max() / ref.abs().max().clamp_min(1e-20)).item())
        # the two paths round identically per tile, but a cut row group adds its pieces in a different association: over 128 DEPENDENT
        # stages single bf16 rounding flips propagate, so the bound here is loose -- the per-stage checks are tests/test_ring_gpu.py
        if not worst <= 6e-2:
            raise RuntimeError(f"ring_check: the ring kernel's step differs from the single-launch step by {worst} (rank {rank})")
        ring_check = {"status": "ok", "max_rel_err_vs_single_launches": worst, "stages": len(units)}
    graph = None
    if not args.no_graph:
        graph = graphs.capture(step_cabi)
    run = graph.replay if graph is not None else step_cabi

    def barrier():
        if comm is not None:
            comm.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        run()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(device.index) as clk:
        pad_until = time.time() + 0.35  # let the sampler see the load: repeat the timed block until 0.35 s have passed
        times = []
        while True:
            barrier()
            e0.record()
            for _ in range(args.steps):
                run()
            e1.record()
            barrier()
            times.append(e0.elapsed_time(e1))
            if time.time() > pad_until:
                break
    ms_total = times[0]  # the contract's number: the FIRST timed block of exactly K steps
    if comm is not None:
        t = torch.tensor([ms_total], device=device)
        comm.all_reduce(t, op=comm.ReduceOp.MAX)
        ms_total = t.item()
    ms_per_step = ms_total / args.steps
    value = step_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: public Python API (Linear4bit.forward captured by quantizations_b200.graphs), host buffers in and out
    xin_host = {K: v.cpu().pin_memory() for K, v in x_in.items()}
    x_static = {K: torch.empty_like(v) for K, v in x_in.items()}
    out_keys = list(outs)
    y_host = {k: torch.empty(outs[k].shape, dtype=dtype).pin_memory() for k in out_keys}
    y_static = {}

    def launch_api(i, m, flags):
        m.gemv_flags = flags
        m.prefetch_next = (packed_of(units[(i + 1) % len(units)]), units[(i + 1) % len(units)].in_features) if args.prefetch else None
        x = source_of(i, m, y_static, x_static[h])
        if fused_ar is not None and m.parallel == "row":
            y = q.gemv_4bit_fused(x, m.weight.data, m.weight.quant_state, flags=flags, allreduce=fused_ar)
        else:
            y = m.forward_fused(x) if isinstance(m, q.Linear4bitGroup) else m(x)
            if comm is not None and m.parallel == "row":
                comm.all_reduce(y)
        y_static[(m.name_, m.out_features)] = y

    y_ring = {k: torch.empty_like(v) for k, v in outs.items()} if use_ring else {}

    def step_api():
        if not use_ring:
            return run_stack(launch_api)
        # the public chain API: quantizations_b200.gemv_4bit_chain collects dependent Linear4bit forwards and issues them as
        # persistent ring launches (eight stages each)
        y_static.update(y_ring)
        with q.gemv_4bit_chain(flags=pdl) as ch:
            for i, m in enumerate(units):
                x = source_of(i, m, y_ring, x_static[h])
                out = y_ring[(m.name_, m.out_features)]
                if isinstance(m, q.Linear4bitGroup):
                    ch.add(x, None, group=m, out=out)
                else:
                    kw = {"allreduce": fused_ar} if (fused_ar is not None and m.parallel == "row") else {}
                    ch.add(x, m.weight.data, m.weight.quant_state, out=out, **kw)

    step_api()
    torch.cuda.synchronize()
    g2 = graphs.capture(step_api) if not args.no_graph else None

    def e2e_step():
        for K, v in xin_host.items():
            x_static[K].copy_(v, non_blocking=True)
        (g2.replay if g2 is not None else step_api)()
        for k in out_keys:
            y_host[k].copy_(y_static[k], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if comm is not None:
        t = torch.tensor([e2e_ms], device=device)
        comm.all_reduce(t, op=comm.ReduceOp.MAX)
        e2e_ms = t.item()
    h2d = sum(v.numel() * v.element_size() for v in xin_host.values())
    d2h = sum(v.numel() * v.element_size() for v in y_host.values())

    # sanity: the e2e outputs equal the C-ABI outputs (same kernels)
    for k in out_keys:
        if not torch.equal(y_host[k].to(device), outs[k]):
            raise RuntimeError(f"e2e output {k} differs from the C-ABI output")
        if not torch.isfinite(outs[k].float()).all():
            raise RuntimeError(f"output {k} of the chained step is not finite")

    if rank != 0:
        if not args.no_decode:  # the tensor-parallel decode leg runs on every rank
            del mods, graph, g2
            torch.cuda.empty_cache()
            decode_tok_s(args, device, "ours", rank, world)
        return None
    peak, peak_src = measured_peak()
    per_launch_us = ms_per_step * 1e3 / launches_per_step
    # DRAM traffic per launch from the committed ncu capture (profiles/gemv_traffic.json), averaged over a layer's launches;
    # only meaningful for the configuration it was captured on
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "gemv_traffic.json")))
        if world == 1 and args.group and args.model == "llama3-8b":
            traffic = int(sum(tj["per_launch_bytes"].values()) / len(tj["per_launch_bytes"]))
    except Exception:
        traffic = None
    res = {
        "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_string(args.model, layers, len(mods) if world == 1 else 7 * layers),
                   "parallelism": f"tp{world}" if world > 1 else "single GPU"},
        "launch": ((f"{launches_per_step} persistent ring launches/step ({len(units)} dependent GEMV stages, <= {_lib.Q4_GEMV_RING_MAX_STAGES} per launch: "
                    "TMA weight ring streaming across stage boundaries, inter-CTA split-K, tagged activation exchange instead of grid barriers)"
                    if use_ring and ring_state["fallbacks"] == 0 else f"{len(units)} launches/step")
                   + (" (q/k/v and gate/up grouped: they share their input)" if args.group else "")
                   + (", CUDA graph replay" if graph is not None else ", eager") + (" + programmatic dependent launch" if args.pdl else "")
                   + (", chained: 1 + layers persistent launches per step (o -> gate/up -> down -> next q/k/v per launch)" if use_chain else "")
                   + (", q/k/v and gate/up as parallel graph branches (co-resident CTAs)" if args.branches else "")
                   + (", next launch's first tiles prefetched into L2" if args.prefetch else "")
                   + (f", tensor-parallel tp{world} (" + ("NCCL all-reduce after" if fused_ar is None else "all-reduce fused into the epilogue of")
                      + " o_proj/down_proj)" if world > 1 else "")),
        "workload_detail": {
            "shapes": sorted({f"{m.out_features}x{m.in_features}" for m in mods}),
            "packed_weight_bytes_per_gpu": packed_bytes, "algorithmic_bytes_per_step": step_bytes,
            "l2_policy": "inputs larger than L2: every layer has its own weights (3.5 GB/step >> 126 MB L2), no flush needed",
        },
        "pct_of_8TBs": round(value / world / 8000 * 100, 2),
        "decode_linear_tok_s": round(1e3 / ms_per_step, 1),
        "e2e": {"value": round(step_bytes / (e2e_ms * 1e-3) / 1e9, 1), "unit": "GB/s", "ms_per_step": round(e2e_ms, 4),
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "quantizations_b200.Linear4bit / Linear4bitGroup forward x%d under quantizations_b200.graphs.capture, pinned-host x in / y out" % len(units)},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clk.summary(),
        "roofline": {"bound": "hbm", "kernel": ("q4::ring::gemv_ring_kernel<bf16, nested, 16, 2>" if use_ring and ring_state["fallbacks"] == 0
                                                else "q4::gemv_mma_kernel<bf16, nested>"), "achieved": round(value / world, 1), "peak": peak,
                     "unit": "GB/s", "frac": round(value / world / peak, 4), "peak_source": peak_src,
                     "avg_launch_us": round(per_launch_us, 3), "algorithmic_bytes_per_launch": step_bytes_local // launches_per_step,
                     "gemv_stages_per_step": len(units), "avg_stage_us": round(ms_per_step * 1e3 / len(units), 3),
                     "traffic": traffic, "timed_blocks_ms": [round(t, 3) for t in times[:5]]},
    }
    if tp_check is not None:
        res["tp_check"] = tp_check
    if ar_latency is not None:
        res["allreduce"] = ar_latency
    if ring_check is not None:
        res["ring_check"] = ring_check
    if world == 1 and not args.no_blockwise:
        res["blockwise"] = blockwise_rates(device, peak)
    if world == 1 and not args.no_sweep:
        res["gemv_sweep"] = gemv_sweep(device, peak, q)
    if world == 1 and not args.no_cpu:
        res["cpu_baseline"] = cpu_baseline(mods[:7], x_in, budget_s=args.cpu_seconds)
    if not args.no_decode:
        del mods, graph, g2
        torch.cuda.empty_cache()
        res["decode"] = decode_tok_s(args, device, "ours", rank, world)
    return res


def gemv_sweep(device, peak, q, pool_bytes=640 << 20, launches=96, replays=10):
    """BASELINE configs[1]: the four Llama-3-8B Linear shapes, NF4 + double quantisation, bf16, batch 1 -- per shape the time of one
    q4_gemv_4bit_fused launch (the kernel behind gemv_4bit / Linear4bit.forward), taken from a CUDA graph of `launches`
    stream-ordered launches with programmatic dependent launch, rotating over a pool of distinct weights larger than L2 so that
    every packed byte comes from HBM.  This is the per-shape figure the north star's 85 % target is defined on."""
    import ctypes

    from quantizations_b200 import _lib

    L = _lib.lib()
    dt = torch.bfloat16
    out = {}
    stream = torch.cuda.Stream(device=device)
    for N, K in [(4096, 4096), (1024, 4096), (14336, 4096), (4096, 14336)]:
        nw = max(3, min(48, pool_bytes // (N * K // 2)))
        lins = []
        for i in range(nw):
            lin = q.Linear4bit(K, N, bias=False, compute_dtype=dt, compress_statistics=True, quant_type="nf4", device="meta")
            W = torch.randn(N, K, device=device, dtype=torch.float32).mul_(K ** -0.5).to(dt)
            lin.weight = q.Params4bit(W, requires_grad=False, quant_type="nf4", module=lin).to(device)
            lins.append(lin)
            del W
        x = torch.randn(1, 1, K, device=device, dtype=dt)
        y = torch.empty(1, 1, N, device=device, dtype=dt)
        structs = []
        for lin in lins:
            st = lin.weight.quant_state
            structs.append(_lib.GemvFused(x.data_ptr(), None, None, 0.0, lin.weight.data_ptr(), ctypes.pointer(st.native_stats()), None, None, 1,
                                          st.code.data_ptr(), None, y.data_ptr(), N, K, st.blocksize, _lib.Q4_BF16, _lib.Q4_GEMV_PDL, None, 0,
                                          st.lut(dt).data_ptr(), None, 0))
        # small batch (speculative / multi-sequence decode): M tokens in ONE pass over the packed weight (q4_gemv_4bit_batch)
        xm = torch.randn(8, K, device=device, dtype=dt)
        ym = torch.empty(8, N, device=device, dtype=dt)

        def issue_single():
            for i in range(launches):
                rc = L.q4_gemv_4bit_fused(ctypes.byref(structs[i % nw]), stream.cuda_stream)
                if rc:
                    _lib.check(rc, "q4_gemv_4bit_fused")

        def issue_tokens(M):
            def issue():
                for i in range(launches):
                    st = lins[i % nw].weight.quant_state
                    rc = L.q4_gemv_4bit_batch(xm.data_ptr(), lins[i % nw].weight.data_ptr(), st.native_stats(), st.code.data_ptr(), None, ym.data_ptr(),
                                              M, N, K, st.blocksize, _lib.Q4_BF16, _lib.Q4_GEMV_PDL, st.lut(dt).data_ptr(), None, 0, stream.cuda_stream)
                    if rc:
                        _lib.check(rc, "q4_gemv_4bit_batch")
            return issue

        def time_graph(issue):
            issue()
            torch.cuda.synchronize(device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                issue()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(replays):
                g.replay()
            e1.record()
            torch.cuda.synchronize(device)
            del g
            return e0.elapsed_time(e1) * 1e3 / (replays * launches)

        with torch.cuda.stream(stream):
            us = time_graph(issue_single)
            us4, us8 = time_graph(issue_tokens(4)), time_graph(issue_tokens(8))
        b = algo_bytes(N, K)
        out[f"{N}x{K}"] = {"us": round(us, 3), "achieved": round(b / us / 1e3, 1), "unit": "GB/s", "frac": round(b / us / 1e3 / peak, 4),
                           "algorithmic_bytes": b, "pool_weights": nw,
                           "tokens4_us": round(us4, 3), "tokens8_us": round(us8, 3), "tokens4_vs_1": round(us4 / us, 3)}
        del lins, structs
        torch.cuda.empty_cache()
    return {"workload": "single Linear4bit NF4 + double-quant GEMV, bf16, bs=1; one q4_gemv_4bit_fused launch per Linear, "
                        f"{launches} stream-ordered launches per CUDA-graph replay (programmatic dependent launch), weights rotating over a pool > L2",
            "kernel": "q4::gemv_mma_kernel<bf16, nested>",
            "small_batch": "tokens4_us / tokens8_us: 4 / 8 tokens in ONE pass over the packed weight (q4_gemv_4bit_batch -> q4::tok::gemv_tokens_kernel), "
                           "measured the same way; the reference sends > 1 token to dequantise + dense GEMM (modules.py:56-64)",
            "shapes": out}


def blockwise_rates(device, peak, N=14336, K=4096, pool=6, iters=30):
    """The other two HBM-bound kernels of the path: blockwise quantize (bf16 -> NF4, fp32 absmax) and dequantize
    (NF4 + double-quant statistics -> bf16) of one 14336x4096 weight, kernel-only through the C ABI, rotating over `pool`
    distinct matrices (inputs + outputs >> L2), CUDA events on the launching stream.  Algorithmic bytes: quantize reads 2n and
    writes n/2 + n/16; dequantize reads n/2 + n/64 + nested statistics and writes 2n."""
    import quantizations_b200 as q
    from quantizations_b200 import _lib

    L = _lib.lib()
    n = N * K
    Ws = [(torch.randn(N, K, device=device) * 0.02).to(torch.bfloat16) for _ in range(pool)]
    packs, states = zip(*[q.quantize_4bit(W, quant_type="nf4") for W in Ws])
    absmaxs = [torch.empty(n // 64, device=device, dtype=torch.float32) for _ in range(pool)]
    outs = [torch.empty_like(p) for p in packs]
    deq = [torch.empty(N, K, device=device, dtype=torch.bfloat16) for _ in range(pool)]
    stats = [s.native_stats() for s in states]
    stream = torch.cuda.current_stream(device).cuda_stream

    def kq(i):
        j = i % pool
        _lib.check(L.q4_quantize_blockwise_4bit(Ws[j].data_ptr(), absmaxs[j].data_ptr(), outs[j].data_ptr(), 64, n, _lib.Q4_NF4,
                                                  _lib.Q4_BF16, stream), "quantize")

    def kd(i):
        j = i % pool
        _lib.check(L.q4_dequantize_blockwise_4bit(packs[j].data_ptr(), stats[j], deq[j].data_ptr(), 64, n, _lib.Q4_NF4, _lib.Q4_BF16,
                                                    stream), "dequantize")

    def timed(fn):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) * 1e3 / iters

    q_bytes = 2 * n + n // 2 + 4 * (n // 64)
    d_bytes = n // 2 + n // 64 + 4 * -(-(n // 64) // 256) + 1092 + 2 * n
    tq, td = timed(kq), timed(kd)
    assert torch.equal(outs[0], packs[0])
    return {"workload": f"{N}x{K} bf16 <-> NF4, blocksize 64, kernel-only, {pool} matrices in rotation (> L2)",
            "quantize": {"us": round(tq, 2), "achieved": round(q_bytes / tq / 1e3, 1), "unit": "GB/s", "frac": round(q_bytes / tq / 1e3 / peak, 3),
                         "algorithmic_bytes": q_bytes, "kernel": "q4::quantize_4bit_lut_kernel<bf16, 64, NF4>"},
            "dequantize": {"us": round(td, 2), "achieved": round(d_bytes / td / 1e3, 1), "unit": "GB/s", "frac": round(d_bytes / td / 1e3 / peak, 3),
                           "algorithmic_bytes": d_bytes, "kernel": "q4::dequantize_4bit_kernel<bf16, NF4, nested>"}}


def decode_tok_s(args, device, impl, rank=0, world=1):
    """Second half of BASELINE's metric: Llama-3 batch-1 greedy decode tok/s (random-init, 32-token prompt, 60 new tokens,
    best of 3) inside quantizations_b200.llama -- ours under a CUDA graph; the reference's kernels eagerly on the legacy
    default stream (they cannot be captured), same NF4 / bf16 configuration.  world > 1: the whole model tensor-parallel over the
    ranks (BASELINE config 5), row-parallel all-reduces inside the GEMV epilogues, time = max over ranks."""
    import dataclasses

    from quantizations_b200 import llama

    base = llama.LLAMA3_70B if getattr(args, "model", "llama3-8b") == "llama3-70b" else llama.LLAMA3_8B
    cfg = dataclasses.replace(base, layers=args.layers or base.layers)
    prompt = torch.arange(1, 33, device=device)
    if impl == "ours":
        model = llama.Llama(cfg, llama.linear4bit_factory(device, torch.bfloat16, "nf4", tp_rank=rank, tp=world), device, torch.bfloat16,
                            tp=world)
        if world > 1:
            from quantizations_b200 import tp as tpmod

            model.fused_ar = None if args.nccl_allreduce else tpmod.FusedAllReduce(cfg.hidden, device=device)
            if model.fused_ar is not None and getattr(args, "ring", True) and getattr(args, "ring_tp", False) and cfg.inter // world <= 16384:
                model.chain = True  # the all-reduces run in the epilogues of the ring kernel's row-parallel stages
        elif (getattr(args, "ring", True) and cfg.inter <= 16384) or getattr(args, "chain", False):
            model.chain = True  # o -> gate/up -> down -> next q/k/v as one persistent launch per layer (ring kernel, else the older chain)
        ctx, graph = torch.cuda.stream(torch.cuda.Stream(device=device)), True
    else:
        from oracle import ref_linear

        model = llama.Llama(cfg, ref_linear.ref_factory(device, torch.bfloat16, "nf4", torch.bfloat16), device, torch.bfloat16)
        ctx, graph = torch.cuda.stream(torch.cuda.default_stream(device)), False
    with ctx:
        model.generate(prompt, 8, use_graph=False)
        times = []
        for _ in range(3):
            dt = model.generate(prompt, 60, use_graph=graph)[1]
            if world > 1:
                t = torch.tensor([dt], device=device, dtype=torch.float64)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                dt = float(t.item())
            times.append(dt)
        best = min(times)
    return {"tok_s": round(60 / best, 1), "ms_per_token": round(best / 60 * 1e3, 3),
            "config": f"Llama-3 hidden {cfg.hidden} shapes, {cfg.layers} layers, random-init, NF4 + double-quant Linear4bit, bf16, bs=1, 32-token "
                      f"prompt, 60 new tokens, greedy, {'CUDA-graph decode step' if graph else 'eager (legacy default stream)'}"
                      + (f", tensor-parallel tp{world} ({'NCCL all-reduce' if model.fused_ar is None else 'all-reduce fused into the o_proj / down_proj epilogues'})" if world > 1 else "")}


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle/ port)


def cpu_baseline(mods, x_in, budget_s=12.0):
    """torch-CPU dequantize -> matmul of the same packed format on the host cores, on ONE decoder layer (7 matrices)."""
    from oracle import q4_torch_cpu as cpu

    ncores = os.cpu_count() or 1
    torch.set_num_threads(ncores)
    host = []
    for m in mods:
        st = m.weight.quant_state
        host.append(dict(packed=m.weight.data.cpu(), qabs=st.absmax.cpu(), code2=st.state2.code.cpu(), absmax2=st.state2.absmax.cpu(),
                         offset=st.offset.cpu(), code=st.code.cpu(), shape=tuple(st.shape),
                         x=torch.randn(1, 1, m.in_features, dtype=m.weight.quant_state.dtype)))  # timing does not depend on the values
    nbytes = sum(algo_bytes(h["shape"][0], h["shape"][1]) for h in host)

    def one_pass():
        for h in host:
            am = cpu.decode_absmax(h["qabs"], h["code2"], h["absmax2"], h["offset"])
            cpu.linear(h["x"], h["packed"], am, h["code"], h["shape"])

    one_pass()
    t0, n = time.perf_counter(), 0
    while True:
        one_pass()
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 20:
            break
    dt = (time.perf_counter() - t0) / n
    return {"value": round(nbytes / dt / 1e9, 3), "unit": "GB/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"one Llama-3-8B decoder layer (7 Linear4bit matrices, {nbytes} algorithmic bytes) x {n} passes, "
                      f"oracle/q4_torch_cpu.py: unpack -> code lookup -> x decoded absmax -> F.linear fp32",
            "ms_per_layer": round(dt * 1e3, 2)}


# ------------------------------------------------------------------------------------------------ reference arm


def workload_string(model: str, layers: int, ngemv: int) -> str:
    """config.workload: the SAME string for both arms (what is computed); how each arm launches it is in `launch`."""
    return (f"{model} Linear4bit stack, NF4 + double-quant, blocksize 64, bf16, bs=1 decode: {layers} layers x "
            f"(q,k,v,o,gate,up,down) = {ngemv} GEMVs/step")


def run_reference(args, rank, world, device):
    """The reference's own CUDA kernels (oracle/_ref/kbkim_lib.so + ref_shim.so, compiled from /root/reference for sm_100a)
    on the same workload, driven with the reference's gemv_4bit call sequence (core.py:467-499): dequantize_blockwise of the
    8-bit absmax, torch `+= offset`, then the GEMV kernel -- three launches per Linear.  bf16 uses the kernel instance the
    reference instantiates but does not export (ops.cu:177), NF4 its code-table argument.

    Nothing of quantizations_b200 is imported or loaded on this arm: the weights are quantised by the reference's own kernels
    (core.py:536-576 re-enacted: blockwise 4-bit quantise, absmax.mean(), 8-bit blockwise quantise of the shifted absmax with
    the dynamic map).  The reference has no NF4 quantiser (ops.cuh:6-10), so the nibbles come from its FP4 quantiser and are
    DECODED with the NF4 table (kernels.cu:851) -- the GEMV's time does not depend on the nibble values."""
    if rank != 0:
        return None
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    cfg = LLAMA3_8B
    layers = args.layers or cfg["layers"]
    ngemv = 7 * layers
    base = {"impl": "reference", "metric": METRIC, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic"}
    have_gpu = torch.cuda.is_available()
    if have_gpu and os.path.exists(os.path.join(ref_dir, "ref_shim.so")) and os.path.exists(os.path.join(ref_dir, "kbkim_lib.so")):
        import numpy as np

        from oracle import q4_oracle as orc  # tables only (dynamic map, NF4 values): numpy, no compute

        sys.path.insert(0, ref_dir)
        import kbkim_lib

        shim = ctypes.CDLL(os.path.join(ref_dir, "ref_shim.so"))
        vp, i32 = ctypes.c_void_p, ctypes.c_int
        shim.ref_gemv_bf16_async.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32]
        shim.ref_gemv_bf16_async.restype = None
        shim.ref_quant_fp4_bf16.argtypes = [vp, vp, vp, i32, i32]
        shim.ref_quant_fp4_bf16.restype = i32
        dtype = torch.bfloat16
        dyn = torch.from_numpy(np.ascontiguousarray(orc.dynamic_map())).to(device)
        code = torch.from_numpy(np.ascontiguousarray(orc.nf4_table())).to(device)
        mods = []
        g = torch.Generator(device=device)
        for layer in range(layers):
            for j, (name, N, K, _par) in enumerate(layer_shapes(cfg, 1)):
                g.manual_seed(1000 * layer + j)  # the same weights as the other arm's build_stack
                W = torch.randn(N, K, device=device, dtype=torch.float32, generator=g).mul_(0.02).to(dtype)
                n = N * K
                absmax = torch.zeros(n // 64, device=device, dtype=torch.float32)
                packed = torch.zeros(n // 2, device=device, dtype=torch.uint8)
                if shim.ref_quant_fp4_bf16(W.data_ptr(), absmax.data_ptr(), packed.data_ptr(), 64, n):   # core.py:552
                    raise RuntimeError("reference quantize kernel failed")
                offset = absmax.mean()                                                                   # core.py:563
                absmax -= offset                                                                         # core.py:564
                nb = absmax.numel()
                qabs = torch.zeros(nb, device=device, dtype=torch.uint8)
                absmax2 = torch.zeros(-(nb // -256), device=device, dtype=torch.float32)
                kbkim_lib.cquantize_blockwise_fp32(dyn.data_ptr(), absmax.data_ptr(), absmax2.data_ptr(), qabs.data_ptr(), 256, nb)  # :565
                torch.cuda.synchronize()
                mods.append(dict(name=name, N=N, K=K, packed=packed, qabs=qabs, absmax2=absmax2, offset=offset))
                del W, absmax
        step_bytes = sum(algo_bytes(m["N"], m["K"]) for m in mods)
        torch.manual_seed(1)
        x_in = {K: torch.randn(1, 1, K, device=device, dtype=dtype) for K in sorted({m["K"] for m in mods})}
        outs = {(m["name"], m["N"]): torch.empty(1, 1, m["N"], device=device, dtype=dtype) for m in mods}
        absmax_buf = {m["qabs"].numel(): torch.empty(m["qabs"].numel(), device=device) for m in mods}

        def step():
            for m in mods:
                nb = m["qabs"].numel()
                am = absmax_buf[nb]
                kbkim_lib.cdequantize_blockwise_fp32(dyn.data_ptr(), m["qabs"].data_ptr(), m["absmax2"].data_ptr(), am.data_ptr(), 256, nb)  # core.py:467
                am += m["offset"]                                                                 # core.py:468
                N, K = m["N"], m["K"]
                shim.ref_gemv_bf16_async(N, 1, K, x_in[K].data_ptr(), m["packed"].data_ptr(), am.data_ptr(), code.data_ptr(),
                                         outs[(m["name"], N)].data_ptr(), N, (K + 1) // 2, N, 64)   # core.py:486-499

        # the reference launches on the legacy default stream (ops.cu:170): run everything there
        with torch.cuda.stream(torch.cuda.default_stream(device)):
            for _ in range(max(3, args.warmup)):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with ClockSampler(device.index) as clk:
                e0.record()
                for _ in range(args.steps):
                    step()
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step()
            torch.cuda.synchronize()
            wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        v = step_bytes / (ms * 1e-3) / 1e9
        return {**base, "value": round(v, 1), "ms_per_step": round(ms, 4),
                "config": {"workload": workload_string("llama3-8b", layers, ngemv), "parallelism": "single GPU"},
                "launch": "reference gemv_4bit call sequence, 3 launches per Linear on the legacy default stream, eager; reference CUDA "
                          "kernels rebuilt for sm_100a (oracle/_ref), weights quantised by the reference's own kernels",
                "algorithmic_bytes_per_step": step_bytes,
                "cpu_baseline": {"value": round(v, 1), "unit": "GB/s", "cores": 0, "kind": "reference",
                                 "sample": "whole step; the reference's implementation of this path is CUDA (it has no CPU path), "
                                           "so it is timed on the same B200 rather than on host cores"},
                "e2e": {"value": round(step_bytes / (wall_ms * 1e-3) / 1e9, 1), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "clocks": clk.summary(), "gpu_launches": 3 * len(mods) * args.steps}
    # no GPU / no compiled reference: the oracle port on host cores
    from oracle import q4_oracle as orc  # noqa: F401
    import numpy as np

    rng = np.random.default_rng(0)
    N, K = 4096, 4096
    st = orc.quantize_4bit((rng.standard_normal((N, K)) * 0.02).astype(np.float32), 64, "nf4")
    x = rng.standard_normal(K).astype(np.float32)
    am = orc.state_absmax(st)
    t0, n = time.perf_counter(), 0
    while time.perf_counter() - t0 < 5 or n < 1:
        orc.gemv_4bit(x, st["packed"], am, st["code"], N, K, 64, "bfloat16")
        n += 1
    dt = (time.perf_counter() - t0) / n
    v = algo_bytes(N, K) / dt / 1e9
    return {**base, "value": round(v, 3), "ms_per_step": round(dt * 1e3, 3), "config": {"workload": "4096x4096 NF4 GEMV, oracle port on host cores"},
            "cpu_baseline": {"value": round(v, 3), "unit": "GB/s", "cores": orc.num_threads(), "kind": "port",
                             "sample": f"{n} x 4096x4096 GEMV (oracle/q4_oracle.c, reference summation order)"},
            "e2e": {"value": round(v, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="llama3-8b", choices=["llama3-8b", "llama3-70b"])
    ap.add_argument("--layers", type=int, default=0, help="decoder layers in the stack (0 = the model's own count)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-pdl", dest="pdl", action="store_false")
    ap.add_argument("--chain", action="store_true",
                    help="one persistent launch per layer (q4_gemv_4bit_chain: o -> gate/up -> down -> next q/k/v) instead of one "
                         "launch per (grouped) Linear; measured SLOWER on B200 (1.65 vs 1.47 ms/step: DESIGN.md 4.1c), hence opt-in")
    ap.add_argument("--no-ring", dest="ring", action="store_false",
                    help="single GPU: one launch per (grouped) Linear (q4_gemv_4bit_fused) instead of the persistent ring kernel "
                         "(q4_gemv_4bit_ring: up to 32 dependent GEMVs per launch, weights streamed through a TMA ring across stage boundaries)")
    ap.add_argument("--ring-tp", action="store_true", help="tensor-parallel runs: the ring kernel with the all-reduce in its row-parallel stages' epilogues")
    ap.add_argument("--nccl-allreduce", action="store_true",
                    help="tensor-parallel runs: NCCL all-reduce after the row-parallel GEMVs (the baseline) instead of the fused epilogue exchange")
    ap.add_argument("--no-prefetch", dest="prefetch", action="store_false",
                    help="do not pass the next launch's packed weight + in_features as the L2 prefetch hint (q4_gemv_fused_t.prefetch_K: "
                         "each CTA pulls the first tiles the next launch's CTAs will ask for into L2; measured 1.27 vs 1.30 ms/step)")
    ap.add_argument("--branches", action="store_true",
                    help="launch q/k/v and gate/up as parallel graph branches (measured slower than one stream + PDL: the graph's "
                         "cross-stream edges cost more than the co-residency gains on 1-5 us kernels)")
    ap.add_argument("--no-group", dest="group", action="store_false", help="one launch per Linear instead of grouped q/k/v and gate/up")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-blockwise", action="store_true", help="skip the quantize / dequantize kernel rates")
    ap.add_argument("--no-sweep", action="store_true", help="skip the per-shape GEMV sweep (BASELINE configs[1])")
    ap.add_argument("--no-ar-latency", action="store_true", help="skip the per-all-reduce latency measurement at --gpus > 1")
    ap.add_argument("--no-tp70b", action="store_true", help="skip the Llama-3-70B leg (BASELINE configs[4]: the 70B Linear stack and decode at this --gpus)")
    ap.add_argument("--no-decode", action="store_true", help="skip the end-to-end Llama-3-8B decode tok/s leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return 0
        device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
        if device.type == "cuda":
            torch.cuda.set_device(device)
        res = run_reference(args, 0, world, device)
        print(json.dumps(res), flush=True)
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: quantizations_b200 has no CPU path (use --impl reference for the CPU oracle timing)")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist

        import datetime

        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=180))  # a rank that went missing fails the run fast
    try:
        res = run_ours(args, rank, world, device)
        if args.model == "llama3-8b" and not args.no_tp70b and not args.layers:
            # BASELINE configs[4]: the Llama-3-70B stack and whole-model decode on the same N GPUs (tensor parallel for N > 1), as an
            # extra block of the same line; `value` stays the 8B stack so that N = 1 equals the plain single-GPU run
            import copy
            import gc

            gc.collect()
            torch.cuda.empty_cache()
            a70 = copy.copy(args)
            a70.model, a70.no_cpu, a70.no_blockwise, a70.no_sweep, a70.no_tp70b = "llama3-70b", True, True, True, True
            a70.no_ar_latency = True
            a70.steps = min(args.steps, 10)
            r70 = run_ours(a70, rank, world, device)
            if rank == 0:
                res["tp70b"] = {k: r70.get(k) for k in ("value", "unit", "ms_per_step", "n_gpus", "steps", "config", "tp_check", "decode")}
                res["tp70b"]["roofline_frac_per_gpu"] = r70["roofline"]["frac"]
                res["tp70b"]["e2e"] = r70["e2e"]["value"]
                res["tp70b"]["workload"] = r70["config"]["workload"]
        if rank == 0:
            print(json.dumps(res), flush=True)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
