// q4_gemv_ring.cuh -- the batch-1 decode GEMV as a persistent, warp-specialised streaming kernel over up to four DEPENDENT
// GEMVs ("stages": o_proj -> gate/up -> down_proj -> next layer's q/k/v), fp16 / bf16 activations, blocksize 64, whole tiles only
// (K % 256 == 0, rows % 32 == 0, grouped matrices end on 32-row boundaries; other shapes stay on q4_gemv_mma.cuh).
//
//   out[r] = sum_b absmax[r,b] * sum_{k in block b} x[k] * code[nib(r,k)]            (+ bias[r])
//
// Replaces reference csrc/kernels.cu:1061-1219 (kgemm_4bit_inference_naive), its launcher ops.cu:167-171 and the two launches
// core.py:467-468 issues before every call.  HBM-bound: every packed byte is read once.
//
// One CTA per SM, (NC consumer warps + a producer warp):
//
//   producer   a few lanes stream the CTA's share of the packed weight HBM -> shared memory with 2-D TMA (cp.async.bulk.tensor,
//              128-byte swizzle) into a ring of 8-KB slots guarded by full / empty mbarriers.  The weight does not depend on the
//              activation, so the producer never waits for anything but a free slot: it runs ahead ACROSS the stages of a launch
//              -- while the consumers wait for the previous stage's output, the ring is already full of this stage's first tiles
//              -- and it starts before griddepcontrol.wait under programmatic dependent launch.
//   consumers  take slots round-robin.  A slot is a 32-row x 512-k tile; a warp takes 8-row sub-tiles of it: a lane copies its 64
//              packed bytes out of the slot (4 conflict-free LDS.128, the swizzle spreads the 8 rows over the banks), decodes ONE
//              packed byte per shared-memory lookup (256-row table of lane-private {code[b>>4], code[b&15]} pairs, address spliced
//              by one PRMT) and feeds the pairs to mma.sync.m16n8k16 as A-fragments: the tensor pipe does the multiply-accumulate
//              and the k-reduction, its 8 B columns are used as 8 quantisation blocks so that D holds the 64 per-(row, block) sums of
//              the tile, each scaled by its decoded absmax (double-quant decode fused: code2[q] * absmax2 + offset, fp32 multiply
//              then add as kernels.cu:552 + core.py:468).
//
// Work split ("inter-CTA split-K").  A stage's slots form a flat list, k fastest; CTA b owns a contiguous range of it, so every SM
// streams the same number of bytes (+-1 slot) whatever the row count.  A row group cut by a range boundary is finished by the CTA
// that holds its first k tile: the other CTA meets that row group FIRST in its own range, publishes the 32 partial sums (value +
// epoch tag, one 8-byte store each) as soon as its warps are through with it, and the owner picks them up at the end of its range
// -- fixed order, no atomics on data, no extra grid-wide step.
//
// Stage boundaries ("one-hop exchange").  There is no grid barrier.  The epilogue of a stage whose output feeds the next stage stores
// every pair of outputs ALSO as one 8-byte {2 x T, epoch tag} word into an exchange buffer (workspace); the next stage's staging
// polls those words directly -- value and "it is there" arrive in the same load, so a boundary costs one store -> L2 -> load trip
// instead of (stores, release, atomic, acquire-spin, loads).  Everything a stage needs to know about its share of the work (range,
// first slot of every warp group, rows it owns) is computed once per launch into a shared-memory plan while the previous kernel is
// still running, so a warp enters a stage with a handful of instructions.
//
// SwiGLU.  A gate/up stage followed by down_proj runs in PAIR mode: a slot holds 16 gate rows and the 16 up rows of the same
// indices (four 16-row TMA boxes), so the CTA that owns a row group owns gate AND up of those rows, and its epilogue publishes
// silu(gate) * up -- the down_proj stage stages 14336 activations instead of 2 x 14336 and computes no exp.
#pragma once

#include <cuda.h>

#include <type_traits>

#include "q4_common.cuh"
#include "q4_gemv_mma.cuh"
#include "q4_tma.h"

namespace q4 {
namespace ring {

constexpr int kSub = 4;                     // 8-row sub-tiles per ring slot
constexpr int kSlotBytes = kSub * 2048;     // 32 rows x 256 packed bytes = two 128-byte-wide, 32-row TMA boxes (pair mode: four 16-row boxes)
constexpr int kSlots = 16;                  // ring depth (a power of two: position / phase of a sequence number are a mask and a shift)
#ifndef Q4_RING_PRODLANES
#define Q4_RING_PRODLANES 4
#endif
constexpr int kProdLanes = Q4_RING_PRODLANES;  // producer lanes issuing slots in lockstep
constexpr int kMaxStages = 32;              // stages per launch (32 x 384-byte descriptors: inside the 32-KB kernel-parameter limit of CUDA >= 12.1;
                                            // their plans fill what is left of shared memory beside the table, the ring and the activation)
constexpr int kXchBufs = 8;                 // exchange buffers: stage s publishes into buffer s % 8 (see the epilogue for why 8 is plenty)
constexpr int kConsumerBar = 1;             // named barrier of the consumer warps
constexpr int kTraceSlots = 16;             // developer trace: globaltimer marks per (stage, CTA)
// workspace layout (Q4_GEMV_RING_WS_BYTES, zeroed once by the caller, owned by one stream at a time)
constexpr int kWsEpochOff = 1024;           // u32 [kWsMaxCtas]: stages run so far, per CTA (the tags of both exchanges)
constexpr int kWsFixOff = 8192;             // {f32, u32} [kWsMaxCtas][32]: partial sums of the row group a CTA shares with its predecessor
constexpr int kWsMaxCtas = 256;
constexpr int kWsXchOff = kWsFixOff + kWsMaxCtas * kSub * 8 * 8;  // {2 x T, u32 tag} [kXchBufs][kXchMaxRows / 2]: stage outputs
constexpr int kXchMaxRows = 32768;
constexpr int kWsBytes = kWsXchOff + kXchBufs * kXchMaxRows * 4;

struct Stage {
    alignas(64) CUtensorMap map;  // packed weight as u8 [rows, K/2], box {128 bytes, 32 rows (pair mode: 16)}, SWIZZLE_128B
    const void* x;
    const void* x_gate;      // optional (plain staging only): effective activation = silu(x_gate[k]) * x[k]
    const void* rms_weight;  // optional: effective activation = x * rsqrt(mean(x^2) + eps) * rms_weight
    AbsmaxView s;
    const float* offsets[kMaxMats];  // nested: per-matrix offset scalars (device pointers)
    int row_end[kMaxMats];           // exclusive end row of each matrix (INT_MAX for unused slots)
    void* out;
    const void* bias;        // [rows] or nullptr (may alias out: residual stream updated in place)
    float rms_eps;
    int rows, K;
    int bias_stage;          // >= 0: bias is the output of that EARLIER stage of this launch -- read from its tagged copy (other CTAs wrote it)
    int KT;                  // ceil(K / 512): k tiles per row group
    int units;               // row groups x KT: length of the flat slot list
    int multi;               // grouped launch with per-matrix offsets
    int pair;                // PAIR mode: rows = 2 * half; a row group is gate rows [16 rg, +16) and up rows half + [16 rg, +16)
    int half;                // pair mode: rows of one member
    int x_tagged;            // the activation is the previous stage's tagged output (exchange buffer stage - 1), not x
    int publish;             // 0: out only; 1: out + tagged out; 2 (pair mode): out + tagged silu(gate) * up
    int skip_out;            // a later stage of this launch overwrites the same `out` (in-place residual stream): only the tagged copy is stored
    int d_rg, d_kt;          // consumer groups per launch configuration: NP / KT, NP % KT (plan_stage)
    int gran;                // slots per assignment quantum: 1 (a row group may be split between two CTAs) or KT (never split)
    int active, per, rem;    // CTA b < active owns quanta [b*per + min(b, rem), +per + (b < rem)); the others idle in this stage
    // fused one-shot all-reduce over tensor-parallel ranks (q4_allreduce_t), ar_world <= 1: off
    void* const* ar_peer_bases;
    int ar_world, ar_rank, ar_max_rows;
};

struct Args {
    Stage st[kMaxStages];
    int n;
    const void* lut;      // prebuilt 64-KB table image, or nullptr: built in the kernel from code / code2
    const float* code;
    unsigned* ws;         // workspace (see above); required when n > 1 or any stage has gran == 1
    int x_bytes;          // shared-memory bytes reserved for the activation vector: max over the stages of KT * 1024
    int part_bytes;       // ... for the per-slot partial sums: max over the stages of (slots per CTA) * 128
    unsigned long long* trace;  // developer: [stage][cta][kTraceSlots] globaltimer marks
};

// what a CTA needs to know about its share of a stage; computed once per launch (consumer warp `stage`), read by everybody
struct Plan {                 // 12 bytes: 32 of them have to fit beside the table, the ring and a 14336-element activation
    int S0;                   // the CTA's range of the flat slot list: [S0, S0 + nloc)
    unsigned short nloc;
    unsigned short rg_own0;   // row groups it owns (= holds the first k tile of): [rg_own0, rg_own0 + nrows_own / 32)
    unsigned short nrows_own;
    unsigned char head;       // its first `head` slots belong to a row group that began in the previous CTA
    unsigned char base;       // ring sequence number of its first slot (slots taken in earlier stages) mod 32: position and phase of a
                              // slot in the 16-deep ring, and the consumer group / producer lane it falls to, depend on nothing else
};
static_assert(sizeof(Plan) == 12, "plan size");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// blocks until the barrier's phase with the given parity has completed.  A pipeline bug must not hang the GPU: after ~2^25 failed
// probes (seconds; a probe itself suspends the thread for a while) the kernel traps.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    for (uint32_t spin = 0;; spin++) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if (spin > (1u << 25)) __trap();
    }
}
// the common case -- the phase has completed -- is one probe and a branch; the bounded spin lives out of line
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}


__device__ __forceinline__ int range_start(const Stage& a, int b) { return (b * a.per + (b < a.rem ? b : a.rem)) * a.gran; }

// the CTA's range of a stage's flat slot list
__device__ __forceinline__ void cta_range(const Stage& a, int b, int& u0, int& u1)
{
    if (b >= a.active) {
        u0 = u1 = 0;
        return;
    }
    const int q0 = b * a.per + (b < a.rem ? b : a.rem);
    const int q1 = q0 + a.per + (b < a.rem ? 1 : 0);
    u0 = q0 * a.gran;
    u1 = q1 * a.gran;
    if (u1 > a.units) u1 = a.units;  // gran == KT never overshoots; kept for safety
}

// first weight row of sub-tile `sub` of row group `rg`
__device__ __forceinline__ int sub_row(const Stage& a, int rg, int sub)
{
    return a.pair ? ((sub >> 1) * a.half + rg * 16 + (sub & 1) * 8) : (rg * (kSub * 8) + sub * 8);
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ld_relaxed_2xu64(const unsigned long long* p, unsigned long long& a, unsigned long long& b)
{
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// second half of the RMSNorm glue: `ss` is the thread's sum of squares over the chunks it staged itself (c = tid, tid + nthr, ...);
// x_eff = x * rsqrt(mean(x^2) + eps) * weight in fp32, rounded to T once
template <typename T>
__device__ __noinline__ void rms_rescale(const void* rms_weight, float rms_eps, int K, uint4* s_x, float* s_red, int nchunk, float ss,
                                         int tid, int nthr)
{
    const int lane = tid & 31, warp = tid >> 5;
    auto slot = [](int c) { return (c & ~7) | ((c ^ (c >> 3)) & 7); };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) s_red[warp] = ss;
    bar_sync(kConsumerBar, nthr);
    ss = 0.0f;
    for (int i = 0; i < (nthr >> 5); i++) ss += s_red[i];
    const float rs = rsqrtf(ss / (float)K + rms_eps);
#pragma unroll 1
    for (int c = tid; c < nchunk; c += nthr) {  // each thread rescales the chunks it wrote itself
        const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(rms_weight) + c);
        const uint4 v = s_x[slot(c)];
        const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
        uint32_t xw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q2 = 0; q2 < 4; q2++) {
            const float2 f = unpack2<T>(xw[q2]), gm = unpack2<T>(ww[q2]);
            xw[q2] = pack2<T>(f.x * rs * gm.x, f.y * rs * gm.y);
        }
        s_x[slot(c)] = make_uint4(xw[0], xw[1], xw[2], xw[3]);
    }
}

// Activation staging by the consumer threads, decode glue fused in (see q4_gemv_fused_t):
//   x_gate:      x_eff = silu(gate) * x, F.silu rounded to T, then the product rounded to T (as the separate torch kernels round)
//   rms_weight:  x_eff = x * rsqrt(mean(x^2) + eps) * weight in fp32, rounded to T once
// chunk c (8 activations) = (block c>>3, piece c&7) lands at piece (c&7) ^ (block&7): the eight lanes that later fetch eight
// different blocks piece by piece hit eight different bank groups
template <typename T>
__device__ __noinline__ void stage_x_glue(const void* x, const void* x_gate, const void* rms_weight, float rms_eps, int K, uint4* s_x,
                                          float* s_red, int nchunk, int npad, int tid, int nthr)
{
    const int lane = tid & 31, warp = tid >> 5;
    auto slot = [](int c) { return (c & ~7) | ((c ^ (c >> 3)) & 7); };
    float ss = 0.0f;
#pragma unroll 1
    for (int c = tid; c < npad; c += nthr) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (c < nchunk) {
            v = __ldcg(reinterpret_cast<const uint4*>(x) + c);
            uint32_t uw[4] = {v.x, v.y, v.z, v.w};
            if (x_gate) {
                const uint4 g4 = __ldcg(reinterpret_cast<const uint4*>(x_gate) + c);
                const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int q2 = 0; q2 < 4; q2++) {
                    const float2 gg = unpack2<T>(gw[q2]), u = unpack2<T>(uw[q2]);
                    // every CTA recomputes the whole vector: the fast exp / divide (~1e-6 relative, far below the rounding to T)
                    const float2 sg = unpack2<T>(pack2<T>(__fdividef(gg.x, 1.0f + __expf(-gg.x)), __fdividef(gg.y, 1.0f + __expf(-gg.y))));
                    uw[q2] = pack2<T>(sg.x * u.x, sg.y * u.y);
                }
                v = make_uint4(uw[0], uw[1], uw[2], uw[3]);
            }
#pragma unroll
            for (int q2 = 0; q2 < 4; q2++) {
                const float2 f = unpack2<T>(uw[q2]);
                ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
            }
        }
        s_x[slot(c)] = v;
    }
    if (rms_weight) rms_rescale<T>(rms_weight, rms_eps, K, s_x, s_red, nchunk, ss, tid, nthr);
}

// NC consumer warps; WPS of them share a ring slot (each takes kSub / WPS of its sub-tiles)
// With 16 consumer warps a 17th (producer) warp would cost every thread registers (5 warps on one scheduler: 96 per thread, and the
// consumer loop spills).  The block then carries a whole producer warpgroup (warps NC .. NC+3, one of them working) that hands
// its registers to the consumers with setmaxnreg: 24 for the producers, kConsumerRegs for the consumers.
constexpr int ring_threads(int nc) { return nc >= 16 ? (nc + 4) * 32 : (nc + 1) * 32; }
template <typename T, bool NESTED, int NC, int WPS>
__global__ void __launch_bounds__(ring_threads(NC), 1)
gemv_ring_kernel(const __grid_constant__ Args c)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    // Shared-memory plan: [table 64 KB][ring: kSlots x 8 KB][x][partials][barriers][scratch][plans].  The PRMT splice of the lookups
    // needs the table at (64-KB aligned window address) + (compile-time immediate): it is the first thing in dynamic shared memory,
    // which starts kDynBase into the CTA's window (probed on the host, trapped here if violated).  kDynBase + 64 KB is a multiple
    // of 1024, which the 128-byte TMA swizzle of the ring slots needs.
    constexpr int kImm = kDynBase;
    constexpr int kCons = NC * 32;
    constexpr int kWarpsPerSlot = WPS;
    constexpr int NP = NC / kWarpsPerSlot;  // slots in work at once: ring sequence number q goes to warp group q % NP
    constexpr int D = kSlots;
    static_assert(NC % kWarpsPerSlot == 0 && kSub % kWarpsPerSlot == 0, "warps per slot");
    static_assert(NP <= 8 && D % NP == 0 && D % kProdLanes == 0 && (D & (D - 1)) == 0, "ring depth / consumer groups");
    constexpr int kSubPerWarp = kSub / kWarpsPerSlot;
    const uint32_t smem_saddr = smem_u32(smem);
    if (smem_saddr != kDynBase) __trap();
    uint8_t* lut = smem;
    uint8_t* ringp = smem + kLutBytes;
    uint4* s_x = reinterpret_cast<uint4*>(ringp + (size_t)D * kSlotBytes);
    float* s_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_x) + c.x_bytes);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_part) + c.part_bytes);  // [0] table, [1 + i] full, [1 + D + i] empty
    float* s_red = reinterpret_cast<float*>(s_bar + 1 + 2 * D);                                          // 16 floats (one per consumer warp)
    uint32_t* s_misc = reinterpret_cast<uint32_t*>(s_red + 16);  // [0] head-piece counter, [1] epoch base, [2] all-reduce epoch
    Plan* s_plan = reinterpret_cast<Plan*>(s_misc + 8);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int bx = blockIdx.x, G = gridDim.x;
    const uint32_t bar0 = smem_u32(s_bar);
    const uint32_t ring_saddr = smem_saddr + kLutBytes;
    auto full = [&](int i) { return bar0 + 8u * (uint32_t)(1 + i); };
    auto empty = [&](int i) { return bar0 + 8u * (uint32_t)(1 + D + i); };
    auto mark = [&](int stage, int slot) {
        if (c.trace) c.trace[((size_t)stage * G + bx) * kTraceSlots + slot] = gtime();
    };

    pdl_launch_dependents();
    if (tid == NC * 32) {  // producer lane 0: barriers first
        mbar_init(bar0, 1);
        for (int i = 0; i < D; i++) {
            mbar_init(full(i), 1);
            mbar_init(empty(i), kWarpsPerSlot);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_misc[0] = 0;
        mark(0, 0);
    }
    for (int ps = warp; ps < c.n && warp < NC; ps += NC) {
        // ---- the plan of stage `ps` (the integer divisions of the CTA's ranges happen here, once per launch, off the critical path)
        const Stage& a = c.st[ps];
        int base = 0;
        for (int s = 0; s < ps; s++) {
            int u0, u1;
            cta_range(c.st[s], bx, u0, u1);
            base += u1 - u0;
        }
        int S0, S1;
        cta_range(a, bx, S0, S1);
        const int nloc = S1 - S0, KT = a.KT;
        if (lane == 0) {
            Plan& p = s_plan[ps];
            p.S0 = S0;
            p.nloc = (unsigned short)nloc;
            p.base = (unsigned char)(base & 31);
            p.head = (unsigned char)((nloc > 0 && (S0 % KT) != 0) ? (KT - S0 % KT < nloc ? KT - S0 % KT : nloc) : 0);
            const int rg_own0 = (S0 + KT - 1) / KT, rg_own1 = (S1 + KT - 1) / KT;
            p.rg_own0 = (unsigned short)rg_own0;
            p.nrows_own = (unsigned short)(nloc > 0 ? (rg_own1 - rg_own0) * (kSub * 8) : 0);
        }
    }
    __syncthreads();

    constexpr bool kRegSplit = NC >= 16;
    if (warp >= NC) {
        if (kRegSplit) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        // ============================================================ producer (warp NC; the rest of its warpgroup only donates registers)
        if (warp != NC) return;
        if (lane == 0 && c.lut) {  // table image: one expect, four bulk copies
            mbar_expect_tx(bar0, kLutBytes);
#pragma unroll
            for (int i = 0; i < 4; i++)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_saddr + i * (kLutBytes / 4)),
                             "l"(reinterpret_cast<const uint8_t*>(c.lut) + i * (kLutBytes / 4)), "r"(kLutBytes / 4), "r"(bar0)
                             : "memory");
        }
        if (lane < kProdLanes) {
            // kProdLanes lanes in lockstep, lane l taking slots l, l + kProdLanes, ... of the CTA's sequence: the latency of the
            // empty-barrier probe and the issue cost of the copies are paid once per kProdLanes slots.  D is a multiple of kProdLanes,
            // so a ring position is always filled by the same lane, fill after fill: its parity wait can never alias an older phase.
            for (int s = 0; s < c.n; s++) {
                const Stage& a = c.st[s];
                const int S0 = s_plan[s].S0, n = s_plan[s].nloc, base = s_plan[s].base;
                const int KT = a.KT, halfK = a.K >> 1;
                if (lane == 0) {
                    asm volatile("prefetch.tensormap [%0];" ::"l"(&a.map) : "memory");
                    mark(s, 6);
                }
                int i = lane - base % kProdLanes;
                if (i < 0) i += kProdLanes;
                int rg = (S0 + i) / KT, kt = (S0 + i) - rg * KT;
                const int d_rg = kProdLanes / KT, d_kt = kProdLanes - d_rg * KT;
                const bool pair = a.pair != 0;
                const int hrows = a.half;
                while (i < n) {
                    const int q = base + i;
                    const int pos = q & (D - 1);
                    const uint32_t ph = (uint32_t)(q / D) & 1u;
                    mbar_wait(empty(pos), ph ^ 1u);
                    const uint32_t dst = ring_saddr + (uint32_t)pos * kSlotBytes;
                    // K % 512 == 256 (tensor-parallel shards): the last k tile holds one 128-byte box only; the other half of the slot
                    // keeps stale bytes, which meet zero activations (the staging pads x with zeros up to whole tiles)
                    const bool two = kt * 256 + 128 < halfK;
                    mbar_expect_tx(full(pos), two ? kSlotBytes : kSlotBytes / 2);
                    if (!pair) {
                        tma_load_2d(dst, &a.map, kt * 256, rg * 32, full(pos));
                        if (two) tma_load_2d(dst + kSlotBytes / 2, &a.map, kt * 256 + 128, rg * 32, full(pos));
                    } else {  // 16 gate rows then the 16 up rows of the same indices, per 128-byte half
                        tma_load_2d(dst, &a.map, kt * 256, rg * 16, full(pos));
                        tma_load_2d(dst + kSlotBytes / 4, &a.map, kt * 256, hrows + rg * 16, full(pos));
                        if (two) {
                            tma_load_2d(dst + kSlotBytes / 2, &a.map, kt * 256 + 128, rg * 16, full(pos));
                            tma_load_2d(dst + 3 * kSlotBytes / 4, &a.map, kt * 256 + 128, hrows + rg * 16, full(pos));
                        }
                    }
                    i += kProdLanes;
                    rg += d_rg;
                    kt += d_kt;
                    if (kt >= KT) {
                        kt -= KT;
                        rg++;
                    }
                }
                if (lane == 0) mark(s, 7);
            }
        }
        return;
    }

    // ================================================================ consumers
    if (kRegSplit) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int g = lane >> 2, t4 = lane & 3;
    const int grp = warp / kWarpsPerSlot, wh = warp % kWarpsPerSlot;  // warp group (slot owner) and the warp's share of the slot's sub-tiles
    // byte offsets of the lane's 16-byte pieces inside a sub-tile: row g of the 8 x 128-byte block, chunks 2*t4, 2*t4+1 (block t4
    // of the half) as the 128-byte swizzle stores them (chunk ^ row)
    const uint32_t off_a0 = (uint32_t)(g * 128 + (((2 * t4) ^ g) << 4)), off_a1 = (uint32_t)(g * 128 + (((2 * t4 + 1) ^ g) << 4));
    const uint32_t lane_base = (uint32_t)(lane * 4);  // table window address is 0: the PRMT splice needs only the lane's word
    const bool xrole = t4 == (g & 3);  // this lane feeds column g of the B operand: x of the tile's block g
    const uint32_t x_saddr = smem_u32(s_x), xswz = (uint32_t)(g * 16);
    uint32_t epoch_base = 0;  // stages this CTA ran in earlier launches (workspace counter): read after griddepcontrol.wait
    uint2* const xch0 = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(c.ws) + kWsXchOff);

    // (Instantiating the stage body once per stage index -- every parameter a constant-bank operand at a fixed address -- was
    // measured: 1 % faster with 16 x 2 warps, 30 % slower with the other layouts; the code no longer fits the instruction cache.)
    for (int stage = 0; stage < c.n; stage++) {
        const Stage& a = c.st[stage];
        const Plan& pl = s_plan[stage];
        const int K = a.K, R = a.rows, KT = a.KT;
        const int bpr = K >> 6;
        const int S0 = pl.S0, nloc = pl.nloc, head = pl.head;
        const bool MULTI = a.multi != 0;
        float off[kMaxMats];
#pragma unroll
        for (int m = 0; m < kMaxMats; m++) off[m] = (NESTED && (MULTI || m == 0) && a.offsets[m]) ? __ldg(a.offsets[m]) : 0.0f;
        // ---- the warp's slots: ring sequence numbers q = base_seq + i with q % NP == grp, i.e. local index i = i0, i0 + NP, ...
        // D is a multiple of NP: a ring position always belongs to the same warp group, which meets every fill of it in order -- a
        // parity wait can never alias an older phase.  Within a stage the group's k tile is constant whenever NP % KT == 0.
        // Everything that moves from slot to slot is a running value advanced by per-stage constants (no multiplications, no
        // parameter reads in the loop): the flat index i, its row group / k tile, and the index `sb` of the lane's absmax pair.
        int ci = grp - (int)(pl.base % NP);  // ring sequence numbers q with q % NP == grp are this group's: local index of its first slot
        if (ci < 0) ci += NP;
        int crg = (S0 + ci) / KT, ckt = (S0 + ci) - crg * KT;
        int left = ci < nloc ? (nloc - ci + NP - 1) / NP : 0;  // slots this warp group still takes in this stage
        const int d_rg = a.d_rg, d_kt = a.d_kt;                // NP / KT, NP % KT (host)
        const int row_mul = a.pair ? kSub * 4 : kSub * 8;
        // first row of the warp's first sub-tile relative to rg * row_mul (its sub-tiles are consecutive 8-row groups)
        const int sub0 = wh * kSubPerWarp;
        const int row_off = (a.pair ? ((sub0 >> 1) * a.half + (sub0 & 1) * 8) : sub0 * 8) + g;
        int sb_sub[kSubPerWarp];  // absmax index of the warp's sub-tile q2 relative to its first one
#pragma unroll
        for (int q2 = 0; q2 < kSubPerWarp; q2++) {
            const int sub = sub0 + q2;
            sb_sub[q2] = ((a.pair ? ((sub >> 1) * a.half + (sub & 1) * 8) : sub * 8) + g - row_off) * bpr;
        }
        const int sb_step = d_rg * row_mul * bpr + d_kt * 8;        // ... between consecutive slots of the warp group
        const int sb_wrap = row_mul * bpr - KT * 8;                 // ... extra when the k tile wraps into the next row group
        int sb = (crg * row_mul + row_off) * bpr + ckt * 8 + 2 * t4;
        int pos;
        uint32_t ph;
        {
            const int seq = (int)pl.base + ci;
            pos = seq & (D - 1);
            ph = (uint32_t)(seq / D) & 1u;
        }
        const uint8_t* const qabs = a.s.qabsmax;
        const float* const am2 = NESTED ? a.s.absmax2 : a.s.absmax;
        const int shift2 = a.s.shift2;
        // grouped launch: the matrix (= offset) a row group belongs to; boundaries are whole row groups (host).  Pair mode: the
        // warp's sub-tiles are all gate rows or all up rows.
        const int mb0 = a.row_end[0] / (kSub * 8), mb1 = a.row_end[1] / (kSub * 8), mb2 = a.row_end[2] / (kSub * 8);

        // absmax of the lane's two blocks (2*t4, 2*t4+1 of row g of a sub-tile), fetched one slot ahead
        struct Stat { uint32_t q; float s0; };  // nested: two 8-bit codes + their second-level absmax; else the two fp32 absmax values
        const int sb_max = R * bpr - 2;  // a ragged last k tile asks for block pairs past its row's end: they meet zero activations, but
        auto load_stat = [&](int sbi) {  // the index must stay inside the array
            Stat r;
            sbi = sbi < sb_max ? sbi : sb_max;
            if (NESTED) {
                r.q = __ldg(reinterpret_cast<const unsigned short*>(qabs + sbi));
                r.s0 = __ldg(am2 + (sbi >> shift2));
            } else {
                const float2 f2 = __ldg(reinterpret_cast<const float2*>(am2 + sbi));
                r.q = __float_as_uint(f2.x);
                r.s0 = f2.y;
            }
            return r;
        };
        Stat st_cur[kSubPerWarp];
#pragma unroll
        for (int q2 = 0; q2 < kSubPerWarp; q2++) {
            st_cur[q2].q = 0;
            st_cur[q2].s0 = 0.0f;
            if (left > 0) st_cur[q2] = load_stat(sb + sb_sub[q2]);  // statistics do not depend on the previous stage: before the exchange
        }

        // ---- everything below may read the previous kernel's (stage 0) or the previous stage's output
        if (tid == 0) {
            mark(stage, 1);
            s_misc[0] = 0;  // head-piece counter of this stage (its last use was before the previous stage's closing barrier)
        }
        if (stage == 0) {
            pdl_wait();
            if (tid == 0) s_misc[1] = c.ws ? __ldcg(c.ws + kWsEpochOff / 4 + bx) : 0u;  // written by this CTA index of the previous launch
        }
        const int nchunk = K >> 3;  // 16-byte chunks of x
        const int npad = KT * 64;   // staged chunks (zero tail up to whole tiles)
        if (a.x_tagged) {
            // One-hop exchange: the previous stage's epilogue stored every pair of its outputs as one 64-bit word {2 x T, tag}; value
            // and arrival are one load.  All of a thread's loads are issued before the first is examined: one round trip when the data
            // is there.  A thread owns whole 16-byte chunks (8 activations), exactly as in the plain staging, so the sum of squares of
            // the RMSNorm is accumulated in the same order.
            const unsigned long long* src = reinterpret_cast<const unsigned long long*>(xch0) + (size_t)((stage - 1) & (kXchBufs - 1)) * (kXchMaxRows / 2);
            const uint32_t tag = epoch_base + (uint32_t)stage;  // the previous stage's epoch
            constexpr int kU = 4;
            float ss = 0.0f;
            for (int cb = tid; cb < npad; cb += kU * kCons) {
                unsigned long long w[kU][4];
#pragma unroll
                for (int j = 0; j < kU; j++) {
                    const int cc = cb + j * kCons;
#pragma unroll
                    for (int q2 = 0; q2 < 4; q2++) w[j][q2] = (unsigned long long)tag << 32;
                    if (cc < nchunk) {
                        ld_relaxed_2xu64(src + (size_t)cc * 4, w[j][0], w[j][1]);
                        ld_relaxed_2xu64(src + (size_t)cc * 4 + 2, w[j][2], w[j][3]);
                    }
                }
#pragma unroll
                for (int j = 0; j < kU; j++) {
                    const int cc = cb + j * kCons;
                    if (cc < npad) {
                        // a chunk is one 32-byte sector written by one store instruction of one CTA: its four words arrive together.
                        // Not there yet: back off before asking again -- 75 000 threads polling flat out keep the L2 slices busier
                        // than the stores they are waiting for.
#ifndef Q4_RING_BACKOFF0
#define Q4_RING_BACKOFF0 32
#define Q4_RING_BACKOFF_CAP 256
#endif
                        unsigned ns = Q4_RING_BACKOFF0;
                        for (long long spin = 0; (uint32_t)(w[j][0] >> 32) != tag || (uint32_t)(w[j][1] >> 32) != tag ||
                                                 (uint32_t)(w[j][2] >> 32) != tag || (uint32_t)(w[j][3] >> 32) != tag; spin++) {
                            __nanosleep(ns);
                            if (ns < Q4_RING_BACKOFF_CAP) ns *= 2;
                            ld_relaxed_2xu64(src + (size_t)cc * 4, w[j][0], w[j][1]);
                            ld_relaxed_2xu64(src + (size_t)cc * 4 + 2, w[j][2], w[j][3]);
                            if (spin > (1ll << 24)) __trap();  // a CTA never published: fail loudly instead of hanging the GPU
                        }
                        uint32_t v[4];
#pragma unroll
                        for (int q2 = 0; q2 < 4; q2++) {
                            v[q2] = (uint32_t)w[j][q2];
                            const float2 f = unpack2<T>(v[q2]);
                            ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
                        }
                        s_x[(cc & ~7) | ((cc ^ (cc >> 3)) & 7)] = make_uint4(v[0], v[1], v[2], v[3]);
                    }
                }
            }
            if (a.rms_weight) rms_rescale<T>(a.rms_weight, a.rms_eps, K, s_x, s_red, nchunk, ss, tid, kCons);
        } else if (a.x_gate || a.rms_weight) {
            stage_x_glue<T>(a.x, a.x_gate, a.rms_weight, a.rms_eps, K, s_x, s_red, nchunk, npad, tid, kCons);
        } else {
            constexpr int kU = NC >= 16 ? 4 : 8;  // loads in flight per thread: the whole vector in one round trip
            for (int cb = tid; cb < npad; cb += kU * kCons) {
                uint4 v[kU];
#pragma unroll
                for (int j = 0; j < kU; j++) {
                    const int cc = cb + j * kCons;
                    v[j] = make_uint4(0, 0, 0, 0);
                    if (cc < nchunk) v[j] = __ldcg(reinterpret_cast<const uint4*>(a.x) + cc);  // coherent: the previous kernel's output
                }
#pragma unroll
                for (int j = 0; j < kU; j++) {
                    const int cc = cb + j * kCons;
                    if (cc < npad) s_x[(cc & ~7) | ((cc ^ (cc >> 3)) & 7)] = v[j];
                }
            }
        }
        if (stage == 0 && !c.lut) {  // callers without a prebuilt image: build the table here
            const float* code2 = a.s.code2;
            for (int cc = tid; cc < kLutBytes / 16; cc += kCons) {
                const int seg = cc >> 3, b = seg >> 1;
                uint32_t word;
                if (seg & 1) word = NESTED ? __float_as_uint(__ldg(code2 + b)) : 0u;
                else word = pack2<T>(__ldg(c.code + (b >> 4)), __ldg(c.code + (b & 15)));
                *reinterpret_cast<uint4*>(lut + cc * 16) = make_uint4(word, word, word, word);
            }
        }
        if (tid == 0) mark(stage, 2);
        bar_sync(kConsumerBar, kCons);
        if (stage == 0 && c.lut) mbar_wait(bar0, 0);  // table landed?
        if (stage == 0) epoch_base = s_misc[1];
        const uint32_t epoch = epoch_base + (uint32_t)stage + 1u;  // tag of this stage's exchanges
        if (tid == 0) mark(stage, 3);

        // ---- main loop: the warp's sub-tiles as ONE continuous stream.  The table lookups run kAhead MMAs ahead of the tensor pipe
        // and keep running across sub-tile and slot boundaries: the next sub-tile's packed bytes replace the current one's in the
        // same registers as soon as those are dead (words 0-3 after MMA 7, words 4-7 after the lookups of MMA 15), so a warp never
        // drains its pipeline between sub-tiles and a slot is released as soon as its last byte has been copied out.
        uint32_t xr[32];
#pragma unroll
        for (int i = 0; i < 32; i++) xr[i] = 0;
        int kt_loaded = -1;
        constexpr int kAhead = 3;  // the reload points below (j == 8, j == 12) and the wrap of the fragment ring across sub-tiles are built on 3
        uint32_t wa[8], wb[8];        // packed bytes of (row g, block t4) and (row g, block 4 + t4) of the sub-tile in work
        uint32_t f[kAhead + 1][4];    // looked-up A fragments in flight
#pragma unroll
        for (int i = 0; i < 8; i++) wa[i] = wb[i] = 0;
        auto fetch = [&](uint32_t (&d)[4], int j) {
            const uint32_t va = wa[j >> 1], vb = wb[j >> 1];
#ifdef Q4_RING_EXPERIMENT_SKIPLOOKUP  // developer experiment (wrong results): every 4th MMA step decodes without shared-memory lookups
            if ((j & 3) == 3) {
                d[0] = va ^ 0x3c003c00u; d[1] = vb ^ 0x3c003c00u; d[2] = (va >> 1) ^ 0x3c003c00u; d[3] = (vb >> 1) ^ 0x3c003c00u;
                return;
            }
#endif
            if (j & 1) {
                d[0] = lut_lookup<2, kImm>(va, lane_base); d[1] = lut_lookup<2, kImm>(vb, lane_base);
                d[2] = lut_lookup<3, kImm>(va, lane_base); d[3] = lut_lookup<3, kImm>(vb, lane_base);
            } else {
                d[0] = lut_lookup<0, kImm>(va, lane_base); d[1] = lut_lookup<0, kImm>(vb, lane_base);
                d[2] = lut_lookup<1, kImm>(va, lane_base); d[3] = lut_lookup<1, kImm>(vb, lane_base);
            }
        };
        auto load_lo = [&](uint32_t sa) {  // words 0-3 of both pieces
            const uint4 a0 = lds128(sa + off_a0), b0 = lds128(sa + kSlotBytes / 2 + off_a0);
            wa[0] = a0.x; wa[1] = a0.y; wa[2] = a0.z; wa[3] = a0.w;
            wb[0] = b0.x; wb[1] = b0.y; wb[2] = b0.z; wb[3] = b0.w;
        };
        auto load_hi = [&](uint32_t sa) {  // words 4-7
            const uint4 a1 = lds128(sa + off_a1), b1 = lds128(sa + kSlotBytes / 2 + off_a1);
            wa[4] = a1.x; wa[5] = a1.y; wa[6] = a1.z; wa[7] = a1.w;
            wb[4] = b1.x; wb[5] = b1.y; wb[6] = b1.z; wb[7] = b1.w;
        };
        auto sub_addr = [&](int p, int q2) { return ring_saddr + (uint32_t)p * kSlotBytes + (uint32_t)((wh * kSubPerWarp + q2) * 1024); };

        if (left > 0) {  // prologue: the first sub-tile's bytes and its first lookups
            mbar_wait_fast(full(pos), ph);
            const uint32_t sa = sub_addr(pos, 0);
            load_lo(sa);
            load_hi(sa);
            if (kSubPerWarp == 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(empty(pos));
            }
#pragma unroll
            for (int j = 0; j < kAhead; j++) fetch(f[j], j);
        }

        while (left > 0) {
            // the warp group's next slot
            int nrg = crg + d_rg, nkt = ckt + d_kt, nsb = sb + sb_step;
            if (nkt >= KT) {
                nkt -= KT;
                nrg++;
                nsb += sb_wrap;
            }
            const int pos_n = (pos + NP) & (D - 1);
            const uint32_t ph_n = ph ^ (uint32_t)(pos + NP >= D);
            const bool has_next = left > 1;
            Stat st_nxt[kSubPerWarp];
#pragma unroll
            for (int q2 = 0; q2 < kSubPerWarp; q2++) {
                st_nxt[q2].q = 0;
                st_nxt[q2].s0 = 0.0f;
                if (has_next) st_nxt[q2] = load_stat(nsb + sb_sub[q2]);
            }
            float o_slot = off[0];
            if (MULTI && !a.pair) o_slot = crg < mb0 ? off[0] : (crg < mb1 ? off[1] : (crg < mb2 ? off[2] : off[3]));

            if (ckt != kt_loaded) {  // warp-uniform; the lookups in flight do not depend on it
                if (xrole) {
                    // chunk i of block kt*8+g sits at piece i ^ g: byte offset (kt*1024 + g*128 + i*16) ^ (g*16)
                    const uint32_t xb = x_saddr + (uint32_t)(ckt * 1024 + g * 128);
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint4 v = lds128((xb | (uint32_t)(i * 16)) ^ xswz);
                        xr[4 * i] = v.x; xr[4 * i + 1] = v.y; xr[4 * i + 2] = v.z; xr[4 * i + 3] = v.w;
                    }
                }
                kt_loaded = ckt;
            }

#pragma unroll
            for (int q2 = 0; q2 < kSubPerWarp; q2++) {
                const int sub = wh * kSubPerWarp + q2;
                const bool last_sub = q2 == kSubPerWarp - 1;
                // what follows this sub-tile in the warp's stream: the slot's next sub-tile, or the first one of the warp's next slot
                const uint32_t sa_next = last_sub ? sub_addr(pos_n, 0) : sub_addr(pos, q2 + 1);
                const bool follows = last_sub ? has_next : true;
                float ce[4] = {0.0f, 0.0f, 0.0f, 0.0f}, co[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#ifdef Q4_RING_EXPERIMENT_NOCOMPUTE  // developer experiment (wrong results): the ring's streaming rate without the decode
                {
                    uint32_t acc = 0;
#pragma unroll
                    for (int j = 0; j < 8; j++) acc ^= wa[j] ^ wb[j];
                    ce[0] = __uint_as_float(acc & 0x3fffffffu);
                    if (follows) {
                        if (last_sub) mbar_wait_fast(full(pos_n), ph_n);
                        load_lo(sa_next);
                        load_hi(sa_next);
                        const bool rel = last_sub ? kSubPerWarp == 1 : q2 + 1 == kSubPerWarp - 1;
                        if (rel) {
                            __syncwarp();
                            if (lane == 0) mbar_arrive(empty(last_sub ? pos_n : pos));
                        }
                    }
                }
#else
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    // MMA j covers bytes 2j, 2j+1 of both pieces = k 4j .. 4j+3 of the lane's blocks; lookups kAhead MMAs ahead, wrapping
                    // into the following sub-tile (whose words 0-3 are in place from j == 8 on)
                    fetch(f[(j + kAhead) % (kAhead + 1)], (j + kAhead) % 16);
                    uint32_t(&a4)[4] = f[j % (kAhead + 1)];
#ifdef Q4_RING_EXPERIMENT_NOMMA  // developer experiment (wrong results): the lookups without the tensor pipe
                    ce[j & 3] = __uint_as_float((__float_as_uint(ce[j & 3]) ^ a4[0] ^ a4[1] ^ a4[2] ^ a4[3] ^ xr[2 * j]) & 0x3fffffffu);
#else
                    if (j & 1) Hmma<T>::run(co, a4[0], a4[1], a4[2], a4[3], xr[2 * j], xr[2 * j + 1]);
                    else Hmma<T>::run(ce, a4[0], a4[1], a4[2], a4[3], xr[2 * j], xr[2 * j + 1]);
#endif
                    if (j == 8 && follows) {  // words 0-3 are dead (last used by the lookups of MMA 7, issued at j == 4)
                        if (last_sub) mbar_wait_fast(full(pos_n), ph_n);
                        load_lo(sa_next);
                    }
                    if (j == 12 && follows) {  // words 4-7 are dead (the lookups of MMA 15 were issued above)
                        load_hi(sa_next);
                        // the warp holds every byte it needs of a slot once the bytes of its last sub-tile there are in registers
                        const bool rel = last_sub ? kSubPerWarp == 1 : q2 + 1 == kSubPerWarp - 1;
                        if (rel) {
                            __syncwarp();
                            if (lane == 0) mbar_arrive(empty(last_sub ? pos_n : pos));
                        }
                    }
                }
#endif
                // the lane's two useful sums: blocks 2t, 2t+1 of the tile (columns 2t, 2t+1; first half -> rows 0-7, second -> 8-15)
                const float u0s = t4 < 2 ? ce[0] + co[0] : ce[2] + co[2];
                const float u1s = t4 < 2 ? ce[1] + co[1] : ce[3] + co[3];
                float am0, am1;
                if (NESTED) {
                    const float q0 = __uint_as_float(lut_lookup<0, kImm + 128>(st_cur[q2].q, lane_base));
                    const float q1 = __uint_as_float(lut_lookup<1, kImm + 128>(st_cur[q2].q, lane_base));
                    const float o = (MULTI && a.pair) ? (((sub0 + q2) >> 1) ? off[1] : off[0]) : o_slot;  // pair mode: gate / up sub-tile
                    am0 = __fadd_rn(__fmul_rn(q0, st_cur[q2].s0), o);  // reference: kernels.cu:552 then core.py:468
                    am1 = __fadd_rn(__fmul_rn(q1, st_cur[q2].s0), o);
                } else {
                    am0 = __uint_as_float(st_cur[q2].q);
                    am1 = st_cur[q2].s0;
                }
                float part = fmaf(u0s, am0, u1s * am1);
                part += __shfl_xor_sync(0xffffffffu, part, 1);
                part += __shfl_xor_sync(0xffffffffu, part, 2);
                if (t4 == 0) s_part[(ci * kSub + sub) * 8 + g] = part;
            }

            if (ci < head) {
                // this slot belongs to the row group shared with the previous CTA: the warp that completes the head piece publishes
                // its 32 partial sums (fixed order over the k tiles) for the owner
                __syncwarp();
                unsigned old = 0;
                if (lane == 0) {
                    __threadfence_block();
                    old = atomicAdd(&s_misc[0], 1u);
                }
                old = __shfl_sync(0xffffffffu, old, 0);
                if (old == (unsigned)(head * kWarpsPerSlot) - 1u) {
                    __threadfence_block();
                    float total = 0.0f;
                    for (int j = 0; j < head; j++) total += reinterpret_cast<volatile float*>(s_part)[j * (kSub * 8) + lane];
                    uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(c.ws) + kWsFixOff) + (size_t)bx * (kSub * 8) + lane;
                    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(total)), "r"(epoch) : "memory");
                }
            }

            ci += NP;
            crg = nrg;
            ckt = nkt;
            sb = nsb;
            left--;
#pragma unroll
            for (int q2 = 0; q2 < kSubPerWarp; q2++) st_cur[q2] = st_nxt[q2];
            pos = pos_n;
            ph = ph_n;
        }
        if (tid == 0) mark(stage, 4);
        // ---- what the epilogue needs from L2 is requested BEFORE the closing barrier (the wait for the CTA's slowest warp hides the
        //      round trip): the neighbour's piece of a cut row group, published long ago, and the residual / bias of the thread's row
        const int nrows_own = pl.nrows_own;
        constexpr int kRows = kSub * 8;  // rows per row group
        uint32_t pre_val = 0, pre_e = 0;
        unsigned long long pre_bias = 0;  // bias_stage >= 0: the tagged word; else the value's bits
        if (tid < nrows_own) {
            const int rg = pl.rg_own0 + tid / kRows, v = tid % kRows;
            if (rg * KT - S0 + KT > nloc) {
                const uint2* fix = reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(c.ws) + kWsFixOff) + v;
                asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(pre_val), "=r"(pre_e) : "l"(fix + (size_t)(bx + 1) * kRows) : "memory");
            }
            if (a.bias) {
                const int r = sub_row(a, rg, v >> 3) + (v & 7);
                if (a.bias_stage >= 0) pre_bias = ld_relaxed_u64(reinterpret_cast<const unsigned long long*>(xch0) + (size_t)(a.bias_stage & (kXchBufs - 1)) * (kXchMaxRows / 2) + (r >> 1));
                else if (r < R) pre_bias = __ldcg(reinterpret_cast<const unsigned short*>(a.bias) + r);
            }
        }
        bar_sync(kConsumerBar, kCons);
        if (tid == 0) {
            mark(stage, 8);
            if (c.trace) {
                unsigned smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                c.trace[((size_t)stage * G + bx) * kTraceSlots + 9] = smid;
            }
        }

        // ---- fixed-order sum over the k tiles of every row group this CTA owns (= holds the first k tile of), [the partner's
        //      partials of a split row group,] [all-reduce over tensor-parallel ranks,] bias / residual, store, publish
        uint32_t ar_epoch = 0;
        if (a.ar_world > 1) {
            // exchanges are counted per CTA in the rank's own exchange area, by EVERY CTA whether or not it owns rows in this stage: all
            // counters of all ranks stay equal, so a tag identifies one exchange whatever the row -> CTA mapping of the layer is
            if (tid == 0) {
                uint32_t* ep = reinterpret_cast<uint32_t*>(a.ar_peer_bases[a.ar_rank]) + bx;
                const uint32_t e = *ep + 1;
                *ep = e;
                s_misc[2] = e;
            }
            bar_sync(kConsumerBar, kCons);
            ar_epoch = s_misc[2];
        }
        if (nrows_own > 0) {
            const int rg_own0 = pl.rg_own0;
            auto row_of = [&](int i) { return sub_row(a, rg_own0 + i / kRows, (i % kRows) >> 3) + (i & 7); };
            auto row_total = [&](int i) {  // i = row index relative to the first owned row group
                const int rg = rg_own0 + i / kRows, v = i % kRows;
                const int j0 = rg * KT - S0;
                const int j1 = j0 + KT < nloc ? j0 + KT : nloc;
                const bool cut = j0 + KT > nloc;  // split row group: its remaining k tiles are the head pieces of the following CTAs
                const uint2* fix = reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(c.ws) + kWsFixOff) + v;
                uint32_t val = pre_val, e = pre_e;  // the thread's first row: requested before the closing barrier
                if (cut && i != tid)
                    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(val), "=r"(e) : "l"(fix + (size_t)(bx + 1) * kRows) : "memory");
                float total = 0.0f;
                for (int j = j0; j < j1; j += 4) {  // the partials of four k tiles are fetched together, added in k order
                    float p4[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) p4[u] = j + u < j1 ? s_part[(j + u) * kRows + v] : 0.0f;
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (j + u < j1) total += p4[u];
                }
                if (cut) {
                    const int gend = (rg + 1) * KT;
                    for (int b2 = bx + 1; b2 < a.active && range_start(a, b2) < gend; b2++) {  // pieces added in CTA order
                        const uint2* src = fix + (size_t)b2 * kRows;
                        for (long long spin = 0;; spin++) {
                            if (spin || b2 != bx + 1) asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(val), "=r"(e) : "l"(src) : "memory");
                            if (e == epoch) break;
                            if (spin > (1ll << 26)) __trap();
                        }
                        total += __uint_as_float(val);
                    }
                }
                return total;
            };
            // y = T(total) (+ bias), stored; published for the next stage when it consumes this output (whole warps: shuffles inside)
            auto emit = [&](int i, float total) {
                const int r = row_of(i);
                const bool valid = r < R;
                T y = Elem<T>::from_f32(total);
                if (a.bias && valid) {  // torch `out += bias`
                    float bv;
                    if (a.bias_stage >= 0) {
                        // written by other CTAs in an earlier stage of this launch: its tagged copy says when it is there
                        const unsigned long long* bsrc = reinterpret_cast<const unsigned long long*>(xch0) + (size_t)(a.bias_stage & (kXchBufs - 1)) * (kXchMaxRows / 2) + (r >> 1);
                        unsigned long long w = i == tid ? pre_bias : ld_relaxed_u64(bsrc);
                        const uint32_t btag = epoch_base + (uint32_t)a.bias_stage + 1u;
                        for (long long spin = 0; (uint32_t)(w >> 32) != btag; spin++) {
                            w = ld_relaxed_u64(bsrc);
                            if (spin > (1ll << 26)) __trap();
                        }
                        const float2 f = unpack2<T>((uint32_t)w);
                        bv = (r & 1) ? f.y : f.x;
                    } else {
                        const unsigned short bits = i == tid ? (unsigned short)pre_bias : __ldcg(reinterpret_cast<const unsigned short*>(a.bias) + r);
                        bv = Elem<T>::to_f32(*reinterpret_cast<const T*>(&bits));
                    }
                    y = Elem<T>::from_f32(Elem<T>::to_f32(y) + bv);
                }
                if (valid && !a.skip_out) reinterpret_cast<T*>(a.out)[r] = y;
                if (a.publish) {
                    // buffer stage % 8: a CTA that publishes stage s has read ALL of stage s - 1's output, so every CTA has published
                    // s - 1 and therefore read all of s - 2: nothing older than two stages is still being read (a residual comes from
                    // at most a few stages back, checked on the host), and a stale word carries another epoch tag anyway
                    unsigned long long* dst = reinterpret_cast<unsigned long long*>(xch0) + (size_t)(stage & (kXchBufs - 1)) * (kXchMaxRows / 2);
                    const float yf = Elem<T>::to_f32(y);
                    if (a.publish == 1) {
                        const float yo = __shfl_xor_sync(0xffffffffu, yf, 1);
                        if (!(lane & 1)) st_relaxed_u64(dst + (r >> 1), ((unsigned long long)epoch << 32) | pack2<T>(yf, yo));
                    } else {
                        // pair mode: lanes 0-15 hold gate rows, lanes 16-31 the up rows of the same indices.  F.silu rounded to T,
                        // then the product rounded to T -- the arithmetic of the staging glue (stage_x_glue), done once instead of per CTA
                        const float up = __shfl_xor_sync(0xffffffffu, yf, 16);
                        const float sg = Elem<T>::to_f32(Elem<T>::from_f32(__fdividef(yf, 1.0f + __expf(-yf))));
                        const float h = sg * up;
                        const float ho = __shfl_xor_sync(0xffffffffu, h, 1);
                        if (lane < 16 && !(lane & 1)) st_relaxed_u64(dst + (r >> 1), ((unsigned long long)epoch << 32) | pack2<T>(h, ho));
                    }
                }
            };
            if (a.ar_world > 1) {
                // One-shot all-reduce in the epilogue (include/quantizations_b200.h: q4_allreduce_t): every partial travels as ONE 8-byte
                // store {value, epoch} into every peer's exchange area; the owner of a row polls the W slots of that row until they
                // carry the current epoch.  The same CTA owns the same rows on every rank (identical launch), slots are double-buffered
                // by epoch parity (a fast peer may start its next exchange while this rank still reads the current one).
                const int W = a.ar_world, me = a.ar_rank;
                uint8_t* mine = reinterpret_cast<uint8_t*>(a.ar_peer_bases[me]);
                const size_t halfsel = (size_t)(ar_epoch & 1) * W * a.ar_max_rows;
                for (int i = tid; i < nrows_own; i += kCons) {
                    const int r = row_of(i);
                    if (r >= R) continue;
                    const float total = row_total(i);
                    for (int p = 0; p < W; p++) {
                        uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(a.ar_peer_bases[p]) + kArDataOffset) + halfsel +
                                     (size_t)me * a.ar_max_rows + r;
                        asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(total)), "r"(ar_epoch) : "memory");
                    }
                }
                const uint2* slots = reinterpret_cast<const uint2*>(mine + kArDataOffset) + halfsel;
                for (int i = tid; i < nrows_own; i += kCons) {
                    const int r = row_of(i);
                    float total = 0.0f;
                    if (r < R) {
                        for (int p = 0; p < W; p++) {  // rank order: every rank computes the same sum
                            const uint2* src = slots + (size_t)p * a.ar_max_rows + r;
                            uint32_t v, e;
                            for (long long spin = 0;; spin++) {
                                asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(e) : "l"(src) : "memory");
                                if (e == ar_epoch) break;
                                if (spin > (1ll << 26)) __trap();  // a peer never arrived: fail loudly instead of hanging the GPU
                            }
                            total += __uint_as_float(v);
                        }
                    }
                    emit(i, total);
                }
            } else {
                for (int i = tid; i < nrows_own; i += kCons) emit(i, row_total(i));
            }
        }
        if (tid == 0) mark(stage, 5);
    }  // stage loop

    // ---- leave the workspace ready for the next launch: this CTA's epoch advanced
    if (c.ws && tid == 0) c.ws[kWsEpochOff / 4 + bx] = epoch_base + (uint32_t)c.n;
}

// ---------------------------------------------------------------------------------------------- host-side planning

// tensor map of a packed weight for the ring's slots: u8 [rows, K/2], box = one half (128 bytes) of a k tile x one row group
// (pair mode: the gate or the up half of a row group, 16 rows)
inline bool make_weight_map(CUtensorMap* m, const void* B, int64_t rows, int64_t K, bool pair = false)
{
    return make_map_2d(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, B, (uint64_t)K / 2, (uint64_t)rows, (uint64_t)K / 2, 128, pair ? kSub * 4 : kSub * 8,
                       CU_TENSOR_MAP_SWIZZLE_128B);
}

// how a stage's flat slot list is dealt to G CTAs.  split_ok (a workspace is available): slot granularity, a row group may be cut
// between neighbouring CTAs (each CTA publishes at most one head piece); else whole row groups.
inline void plan_stage(Stage& a, int rows, int K, int G, bool split_ok, int np = 8)
{
    a.rows = rows;
    a.K = K;
    a.KT = (K + 511) / 512;
    const int RT = a.pair ? a.half / (kSub * 4) : (rows + kSub * 8 - 1) / (kSub * 8);  // row groups (one slot = a row group x one k tile)
    const long long units = (long long)RT * a.KT;
    a.units = (int)units;
    long long Q;
    // Splitting row groups between CTAs balances the bytes per SM, but the owner of a cut row group has to pick the other piece
    // up.  That is free when the other CTA finishes the piece rounds before the owner needs it (it meets it first), and a memory
    // round trip on the critical path when both CTAs are through after a single round of their consumer groups (np slots are in
    // work at once): 4096 x 4096 on 148 SMs is 6.9 slots per CTA split, 8 unsplit -- one round either way -- so it stays whole.
    if (split_ok) {
        const long long whole = ((RT + G - 1) / G) * (long long)a.KT;  // longest slot list, whole row groups
        const long long cut = (units + G - 1) / G;                     // ... cut at slot granularity
        if (whole <= np || whole - cut < 2) split_ok = false;
    }
    if (split_ok) {
        a.gran = 1;
        Q = units;
        a.active = (int)(units < G ? units : G);
    } else {
        a.gran = a.KT;
        Q = RT;
        a.active = RT < G ? RT : G;
    }
    a.per = (int)(Q / a.active);
    a.rem = (int)(Q % a.active);
    a.d_rg = np / a.KT;
    a.d_kt = np % a.KT;
}

// shared-memory plan of a launch: fills x_bytes / part_bytes, returns the dynamic shared-memory size (0: does not fit)
inline size_t plan_launch(Args& c, size_t max_smem = 227 * 1024)
{
    c.x_bytes = 0;
    c.part_bytes = 0;
    for (int i = 0; i < c.n; i++) {
        const Stage& a = c.st[i];
        if (a.KT * 1024 > c.x_bytes) c.x_bytes = a.KT * 1024;
        const int tiles = (a.per + (a.rem ? 1 : 0)) * a.gran;
        if (tiles * kSub * 32 > c.part_bytes) c.part_bytes = tiles * kSub * 32;
    }
    c.part_bytes = (c.part_bytes + 127) & ~127;
    const size_t total = (size_t)kLutBytes + (size_t)kSlots * kSlotBytes + c.x_bytes + c.part_bytes + (1 + 2 * kSlots) * 8 /* barriers */ +
                         64 /* s_red */ + 32 /* s_misc */ + kMaxStages * sizeof(Plan);
    return total <= max_smem ? total : 0;
}

}  // namespace ring
}  // namespace q4
