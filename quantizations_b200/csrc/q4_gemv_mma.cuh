// q4_gemv_mma.cuh -- the batch-1 decode GEMV's fast path (fp16 / bf16 activations, blocksize 64, K % 128 == 0).
//
//   out[r] = sum_b absmax[r,b] * sum_{k in block b} x[k] * code[nib(r,k)]            (+ bias[r])
//
// Replaces reference csrc/kernels.cu:1061-1219 (kgemm_4bit_inference_naive) and the two launches core.py:467-468 issues
// before it.  HBM-bound: every packed byte is read once, straight into registers, and the work per byte is kept under the
// SM's issue budget at full HBM rate (B200: ~22 B/clk/SM of packed weight = 44 weights/clk/SM against 128 lane-instructions
// per clock -- the reference's ~5 instructions per weight cannot stream faster than ~1/4 of HBM):
//
//   * ONE shared-memory lookup per packed BYTE.  A 256-row table, row b = 32 lane-private copies of the pair
//     {code[b>>4], code[b&15]} already rounded to the activation type (what the reference's `T quant_map[16]` holds), followed
//     by 32 copies of code2[b] (fp32, the 8-bit absmax map).  Lane l reads word l of row b: conflict-free by construction.
//     The address is formed by a single PRMT that splices the weight byte into bits 8-15 of {window | lane*4}.
//   * The multiply-accumulate AND the sum over k run on the tensor pipe: the looked-up pair IS an A-fragment register of
//     mma.sync.m16n8k16 (two consecutive k of one row), so one HMMA per four lookups replaces four FMAs plus the per-row
//     shuffle reduction, accumulates in fp32 (no fp16 partial sums, no activation pre-scaling), and takes bf16 activations
//     natively.  This is not a GEMM reshape -- the HMMA's 8 columns are used as 8 quantisation blocks (below); the kernel
//     stays a byte stream bounded by HBM, the tensor pipe is ~4 % busy.
//   * Table and activations arrive without per-thread work where possible: the 64-KB table is one TMA bulk copy from a
//     prebuilt global image (q4_gemv_lut_build; L2-resident, shared by every layer), signalled on an mbarrier.
//
// Tile = 8 weight rows x 8 blocks (512 k) = 2 KB of packed bytes per warp step.  MMA row m < 8 is weight row m over the
// tile's blocks 0-3, MMA row m+8 is the SAME weight row over blocks 4-7.  The 16 k-slots of one MMA are 4 blocks x 4
// consecutive k; column n < 4 of the B operand carries x for block n (zero in the slots of other blocks), column n >= 4
// carries x for block n (= 4 + n - 4) of the second half.  So D[m][n] (m < 8, n < 4) and D[m+8][n] (n >= 4) are the 64
// per-(row, block) partial sums of the tile, two per lane, ready to be scaled by their block's absmax -- every lane decodes
// exactly the two absmax values it needs (one 16-bit load), nothing is redundant.  Lane (g = lane/4, t = lane%4) loads
// bytes [32t, 32t+32) and [128+32t, 160+32t) of weight row g's 256-byte tile row: the four lanes of a group read 256
// contiguous bytes, each with two 256-bit loads (whole DRAM sectors).
//
// A CTA owns a contiguous range of 8-row tiles (one contiguous byte range of the weight, queued HBM -> L2 with TMA bulk
// prefetches in its first instructions); its warps deal the (row tile, k tile) pairs round-robin, each keeps one tile in
// flight while computing another, and writes one partial per (row, k tile) to shared memory; after one barrier the partials
// are summed in a fixed order (deterministic) and stored with the bias / residual.
#pragma once

#include <type_traits>

#include "q4_common.cuh"

namespace q4 {

constexpr int kLutBytes = 65536;  // 256 rows x 256 B
constexpr int kMaxMats = 4;

struct MmaGemvArgs {
    const void* x;
    const float* code;   // 16-entry 4-bit code table (used only when lut == nullptr)
    const void* lut;     // prebuilt 64-KB table image for this (code, code2, dtype), or nullptr: build it in the kernel
    const uint8_t* Bq;   // packed weight [rows, K/2]
    AbsmaxView s;
    const float* offsets[kMaxMats];  // nested: per-matrix offset scalars (device pointers)
    int row_end[kMaxMats];           // exclusive end row of each matrix (INT_MAX for unused slots)
    void* out;
    const void* bias;        // [rows] or nullptr (may alias out: residual stream updated in place)
    const void* x_gate;      // optional: effective activation = silu(x_gate[k]) * x[k]
    const void* rms_weight;  // optional: effective activation = x * rsqrt(mean(x^2) + eps) * rms_weight
    float rms_eps;
    const uint8_t* next;     // optional L2 prefetch hint for the next launch
    int64_t next_bytes;
    int rows, K;
    int rt_total;  // ceil(rows / 8)
    int kt;        // ceil(K / 512): k tiles per row tile
    unsigned long long* trace;
    int debug;  // developer experiments (env Q4_GEMV_DEBUG)
    // fused one-shot all-reduce over tensor-parallel ranks (q4_allreduce_t), ar_world <= 1: off
    void* const* ar_peer_bases;
    int ar_world, ar_rank, ar_max_rows;
    // host-computed strides of the fast path's load cursor (tile index += warps per CTA)
    long long pw_step, pw_wrap;
    int ab_step, ab_wrap, d_rt, d_kt;
    int multi;  // grouped launch with per-matrix offsets (nested statistics)
    int rt_q, rt_r;  // rt_total / grid, rt_total % grid: CTA b owns row tiles [b*rt_q + min(b, rt_r), ...) -- no division on the device
    // exact prefetch hint (nx_grid > 0): the NEXT launch runs nx_grid CTAs over `next`, CTA j owning row tiles
    // [j*nx_rt_q + min(j, nx_rt_r), ...) of nx_tile_bytes each and loading the first nx_head of them before its activation exists
    int nx_grid, nx_rt_q, nx_rt_r, nx_head;
    long long nx_tile_bytes;
    int swiglu;  // Q4_GEMV_SWIGLU (host side only: selects the SWIGLU instantiation of the kernel)
};

// A launch runs `n` dependent GEMVs back to back ("chain": e.g. o_proj -> gate/up -> down_proj -> next layer's q/k/v): one
// persistent grid, the table copied once, a grid-wide barrier instead of a launch boundary between stages, and the first weight
// tiles of stage s+1 already in registers while the barrier is pending.  n == 1 is the plain single-GEMV launch.
constexpr int kMaxChain = 4;
template <int NST> struct MmaStagesArgs {
    MmaGemvArgs st[NST];
    int n;
    int x_bytes;          // shared-memory bytes reserved for the activation vector: max over the stages of kt * 1024
    unsigned* barrier;    // n > 1: grid-barrier counter in global memory (zero between launches; the kernel leaves it zero)
};
using MmaChainArgs = MmaStagesArgs<kMaxChain>;
using MmaSingleArgs = MmaStagesArgs<1>;  // the plain launch: a quarter of the parameter bytes

// Grid-wide barrier of a persistent launch whose CTAs are all co-resident (the dispatcher sizes the grid accordingly).
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v;
        for (long long spin = 0;; spin++) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v >= target) break;
            if (spin > (1ll << 27)) __trap();  // a CTA never arrived (grid not co-resident?): fail loudly
        }
    }
    __syncthreads();
}

constexpr int kArMaxCtas = 1024;       // Q4_AR_MAX_CTAS
constexpr int kArDataOffset = 65536;   // Q4_AR_DATA_OFFSET
constexpr int kArMaxWorld = 8;

template <typename T> struct Hmma;
template <> struct Hmma<__half> {
    static __device__ __forceinline__ void run(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
    {
        asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
};
template <> struct Hmma<__nv_bfloat16> {
    static __device__ __forceinline__ void run(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
    {
        asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
};

// table lookup: row = weight byte SEL of w, word = lane; IMM = where the table starts past the 64-KB boundary (+128: code2 half)
template <int SEL, int IMM> __device__ __forceinline__ uint32_t lut_lookup(uint32_t w, uint32_t lane_base)
{
    uint32_t v;
    const uint32_t addr = __byte_perm(w, lane_base, 0x7604 | (SEL << 4));
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}

__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes)
{
    bytes &= ~15u;
    if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// spread a byte range over the lanes of one warp in 8-KB pieces (TMA bulk prefetch: fire and forget, no registers)
__device__ __forceinline__ void bulk_prefetch_l2_range(const uint8_t* p, int64_t bytes, int lane)
{
    constexpr int64_t kPiece = 8192;
    const int64_t skew = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
    p += skew;
    bytes -= skew;
    for (int64_t o = (int64_t)lane * kPiece; o < bytes; o += 32 * kPiece)
        bulk_prefetch_l2(p + o, (uint32_t)(bytes - o < kPiece ? bytes - o : kPiece));
}

__device__ __forceinline__ void mma_trace(const MmaGemvArgs& a, int slot)
{
    if (a.trace && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.trace[blockIdx.x * 8 + slot] = t;
    }
}

// developer trace, CTA 0 only: extra marks after the per-CTA slots (tools/gemv_trace.py prints them relative to mark 0)
__device__ __forceinline__ void mma_trace_x(const MmaGemvArgs& a, int k)
{
    if (a.trace && threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.trace[1024 * 8 + k] = t;  // the trace buffer holds 1024 CTAs x 8 slots, then the extras
    }
}

static __device__ __noinline__ void prefetch_next_share(const uint8_t* next, int64_t next_bytes, int lane)
{
    const int64_t share = ((next_bytes / gridDim.x) + 15) & ~(int64_t)15;
    const int64_t lo = share * blockIdx.x;
    const int64_t n = lo + share <= next_bytes ? share : next_bytes - lo;
    if (n > 0) bulk_prefetch_l2_range(next + lo, n, lane);
}

// Exact form of the hint: only the head of every next-launch CTA's share -- the tiles it will load into registers during its cold
// start, ~14 MB over the grid -- goes HBM -> L2.  Measured on the Llama-3-8B stack (1.305 ms/step without a hint): issued in the
// CTA's first instructions 1.287 ms, after its loop 1.312 ms, right after griddepcontrol.wait 1.357 ms; two / four times the head
// 1.32 / 1.39 ms; additionally the CTA's OWN range beyond its first tiles 1.38 ms -- bulk L2 prefetch only pays for bytes that
// are on the critical path of a cold start.
static __device__ __noinline__ void prefetch_next_heads(const uint8_t* next, int nx_grid, int nx_rt_q, int nx_rt_r, int nx_head,
                                                 long long nx_tile_bytes, int lane)
{
    for (int j = blockIdx.x; j < nx_grid; j += gridDim.x) {
        const int rt0 = j * nx_rt_q + (j < nx_rt_r ? j : nx_rt_r);
        const int nrt = nx_rt_q + (j < nx_rt_r ? 1 : 0);
        const int head = nrt < nx_head ? nrt : nx_head;
        bulk_prefetch_l2_range(next + (int64_t)rt0 * nx_tile_bytes, (int64_t)head * nx_tile_bytes, lane);
    }
}

// table built in the kernel (callers without a prebuilt image): out of line for the same reason
template <typename T, bool NESTED>
__device__ __noinline__ void build_lut_in_kernel(uint8_t* lut, const float* code, const float* code2)
{
    for (int c = threadIdx.x; c < kLutBytes / 16; c += blockDim.x) {
        const int seg = c >> 3, b = seg >> 1;
        uint32_t word;
        if (seg & 1) word = NESTED ? __float_as_uint(__ldg(code2 + b)) : 0u;
        else word = pack2<T>(__ldg(code + (b >> 4)), __ldg(code + (b & 15)));
        *reinterpret_cast<uint4*>(lut + c * 16) = make_uint4(word, word, word, word);
    }
}

struct TileRegs {
    u32x8 wa, wb;  // 32 packed bytes of (row g, block t) and of (row g, block 4 + t)
    uint32_t q;    // nested: two 8-bit absmax codes (blocks 2t, 2t+1 of the tile)
    float s0, s1;  // nested: s0 = second-level absmax; else the two fp32 absmax values
};

// Activation staging with the decode glue fused in (rare path, kept out of line so the main kernel stays small):
//   x_gate:      x_eff = silu(gate) * x, F.silu rounded to T, then the product rounded to T (as the separate torch kernels round)
//   rms_weight:  x_eff = x * rsqrt(mean(x^2) + eps) * weight in fp32, rounded to T once
template <typename T, bool SWZ>
__device__ __noinline__ void stage_x_fused(const void* x, const void* x_gate, const void* rms_weight, float rms_eps, int K, uint4* s_x,
                                           float* s_red, int nchunk, int npad)
{
    struct { const void *x, *x_gate, *rms_weight; float rms_eps; int K; } a = {x, x_gate, rms_weight, rms_eps, K};
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    auto slot = [](int c) { return SWZ ? ((c & ~7) | ((c ^ (c >> 3)) & 7)) : c; };  // SWZ: the mma.sync kernel's bank swizzle
    float ss = 0.0f;
#pragma unroll 1
    for (int c = tid; c < npad; c += nthr) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (c < nchunk) {
            v = __ldcg(reinterpret_cast<const uint4*>(a.x) + c);
            uint32_t uw[4] = {v.x, v.y, v.z, v.w};
            if (a.x_gate) {
                const uint4 g4 = __ldcg(reinterpret_cast<const uint4*>(a.x_gate) + c);
                const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int q2 = 0; q2 < 4; q2++) {
                    const float2 gg = unpack2<T>(gw[q2]), u = unpack2<T>(uw[q2]);
                    // every CTA recomputes the whole vector: the fast exp / divide (~1e-6 relative, far below the rounding to T)
                    // instead of the IEEE ones take this from ~5 us to ~1.5 us per launch on a 14336-wide input
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
    if (!a.rms_weight) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) s_red[warp] = ss;
    __syncthreads();
    ss = 0.0f;
    for (int i = 0; i < (nthr >> 5); i++) ss += s_red[i];
    const float rs = rsqrtf(ss / (float)a.K + a.rms_eps);
#pragma unroll 1
    for (int c = tid; c < nchunk; c += nthr) {  // each thread rescales the chunks it wrote itself
        const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(a.rms_weight) + c);
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

constexpr int kDynBase = 1024;  // where dynamic shared memory starts in the CTA's shared window on sm_100 (probed on the host)
#ifdef Q4_MMA_THREADS
constexpr int kMmaThreads = Q4_MMA_THREADS;
#else
constexpr int kMmaThreads = 256;
#endif
constexpr int kBuffers = 3;     // weight tiles a warp holds in registers (one being consumed, the others in flight)

// SWIGLU (Q4_GEMV_SWIGLU): two matrices (gate, up) interleaved in chunks of 4 rows -- rows 8t..8t+3 are gate rows 4t..4t+3, rows
// 8t+4..8t+7 the up rows of the same index -- and the launch stores silu(gate) * up [rows / 2] instead of the rows.  A template
// parameter, not a run-time flag: as a flag it cost the plain kernel three registers and 5 % of its speed (1.341 vs 1.270 ms/step).
template <typename T, bool NESTED, bool COMPACT, bool TAIL, bool CHAIN, bool SWIGLU = false>
__global__ void __launch_bounds__(kMmaThreads, 512 / kMmaThreads)
gemv_mma_kernel(const __grid_constant__ typename std::conditional<CHAIN, MmaChainArgs, MmaSingleArgs>::type c)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    // Shared-memory plan.  The PRMT splice needs the table at (64-KB aligned window address) + (compile-time immediate).
    //   COMPACT: the table is the first thing in dynamic shared memory, which starts kDynBase into the window -> ~75-100 KB per
    //            CTA, two CTAs per SM: in a decode chain the NEXT launch's CTA is co-resident and runs its whole x-independent
    //            prologue (cold start, table copy, first weight tiles into registers) while this one computes.
    //   else:    64 KB of slack in front, table at the next 64-KB boundary (one CTA per SM).
    const uint32_t smem_saddr = (uint32_t)__cvta_generic_to_shared(smem);
    constexpr int kImm = COMPACT ? kDynBase : 0;
    const uint32_t lut_saddr = COMPACT ? (smem_saddr - kDynBase) : ((smem_saddr + 0xFFFFu) & 0xFFFF0000u);
    if (COMPACT && lut_saddr != 0) __trap();  // the host probe and the kernel disagree about the window layout
    uint8_t* lut = COMPACT ? smem : smem + (lut_saddr - smem_saddr);
    uint4* s_x = reinterpret_cast<uint4*>(lut + kLutBytes);                   // kt*64 chunks of 8 activations, swizzled
    float* s_red = reinterpret_cast<float*>(lut + kLutBytes + c.x_bytes);     // 32 floats
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_red + 32);                // mbarrier (+ pad)
    float* s_part = reinterpret_cast<float*>(s_bar + 2);                      // [row tile][k tile][8 rows]

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
    const int g = lane >> 2, t4 = lane & 3;

    pdl_launch_dependents();
    mma_trace(c.st[0], 0);

    // ---- table: one TMA bulk copy of the prebuilt image (no thread touches it), completion on an mbarrier
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(s_bar);
    if (c.st[0].lut && tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#ifdef Q4_GEMV_EXPERIMENT_SMALLTABLE  // developer experiment (wrong results): what the launch costs with a quarter of the table traffic
        constexpr int kCopy = kLutBytes / 4;
#else
        constexpr int kCopy = kLutBytes;
#endif
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kCopy) : "memory");
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             lut_saddr + kImm + i * (kCopy / 4)),
                         "l"(reinterpret_cast<const uint8_t*>(c.st[0].lut) + i * (kCopy / 4)), "r"(kCopy / 4), "r"(bar)
                         : "memory");
    }
    mma_trace(c.st[0], 7);
    // ---- optional hint: this CTA's share of what the NEXT launch will stream, HBM -> L2 (out of line: rarely used, and the cold
    //      start of a CTA is on the critical path of a launch chain -- the straight-line prologue should be short)
    if (warp == nw - 1 && c.st[0].next_bytes > 0) {
        if (c.st[0].nx_grid == 0) prefetch_next_share(c.st[0].next, c.st[0].next_bytes, lane);
        else prefetch_next_heads(c.st[0].next, c.st[0].nx_grid, c.st[0].nx_rt_q, c.st[0].nx_rt_r, c.st[0].nx_head, c.st[0].nx_tile_bytes, lane);
    }

  // CHAIN = false: exactly one stage, every argument a compile-time offset into the parameter bank (no indexed constant loads
  // in the loop); CHAIN = true: c.n stages, arguments indexed by the stage.
  for (int stage = 0; stage < (CHAIN ? c.n : 1); stage++) {  // (body indented as the single-stage kernel it grew from)
    const MmaGemvArgs& a = CHAIN ? c.st[stage] : c.st[0];
    const int K = a.K, R = a.rows, KT = a.kt;
    const int bpr = K >> 6;  // 64-wide blocks per row
    // this CTA's row tiles: a contiguous byte range of the packed weight and of the statistics
    const int bx = blockIdx.x;
    const int rt0 = bx * a.rt_q + (bx < a.rt_r ? bx : a.rt_r);
    const int rt1 = rt0 + a.rt_q + (bx < a.rt_r ? 1 : 0);
    const int ntiles = (rt1 - rt0) * KT;
    const int row_lo = rt0 * 8, row_hi = rt1 * 8 < R ? rt1 * 8 : R;
    const bool MULTI = a.multi != 0;
    float off[kMaxMats];
#pragma unroll
    for (int m = 0; m < kMaxMats; m++) off[m] = (NESTED && (MULTI || m == 0) && a.offsets[m]) ? __ldg(a.offsets[m]) : 0.0f;
    mma_trace_x(a, 0);

    // ---- tile bookkeeping: warp w takes tiles w, w + nw, ... of the CTA's (row tile, k tile) grid, k fastest
    // TAIL = false (K % 512 == 0 and rows % 8 == 0, every Llama shape): no bounds, no zero fill, and the load addresses advance
    // by constants (tile index += nw means kt += nw % KT, rt += nw / KT, one conditional wrap).
    struct Cursor { int t, rt, kt; };
    const int d_rt = a.d_rt, d_kt = a.d_kt;  // nw / KT, nw % KT
    auto after = [&](const Cursor& c) {  // the warp's next tile
        Cursor n = {c.t + nw, c.rt + d_rt, c.kt + d_kt};
        if (n.kt >= KT) {
            n.kt -= KT;
            n.rt++;
        }
        return n;
    };
    // load cursor of the fast path: byte pointer of (row g of the row tile, block t4 of the k tile) and the block index of the
    // lane's absmax pair
    const int64_t row_bytes = (int64_t)bpr * 32;
    int w_rt = 0, w_kt = warp;  // warp / KT, warp % KT (warp < 8: a short subtraction loop instead of a division)
    while (w_kt >= KT) {
        w_kt -= KT;
        w_rt++;
    }
    const uint8_t* pw = a.Bq + ((int64_t)(rt0 + w_rt) * 8 + g) * row_bytes + (int64_t)w_kt * 256 + t4 * 32;
    int ab = ((rt0 + w_rt) * 8 + g) * bpr + w_kt * 8 + 2 * t4;  // rows * bpr < 2^31: dispatcher
    int kt_ld = w_kt;
    auto issue = [&](TileRegs& r, const Cursor& c) {
        if (TAIL) {
            const int row = (rt0 + c.rt) * 8 + g;
            const int row_c = row < R ? row : R - 1;  // rows past the end load a valid row and are never stored
            const int blk0 = c.kt * 8;
            const int base = row_c * bpr + blk0;
#pragma unroll
            for (int j = 0; j < 8; j++) r.wa.v[j] = r.wb.v[j] = 0;
            if (blk0 + t4 < bpr) r.wa = ldg_stream_256(a.Bq + (int64_t)(base + t4) * 32);
            if (blk0 + 4 + t4 < bpr) r.wb = ldg_stream_256(a.Bq + (int64_t)(base + 4 + t4) * 32);
            r.q = 0;
            r.s0 = r.s1 = 0.0f;
            if (blk0 + 2 * t4 < bpr) {  // bpr is even: the pair is valid together
                const int sb = base + 2 * t4;
                if (NESTED) {
                    r.q = __ldg(reinterpret_cast<const unsigned short*>(a.s.qabsmax + sb));
                    r.s0 = __ldg(a.s.absmax2 + (sb >> a.s.shift2));
                } else {
                    const float2 f = __ldg(reinterpret_cast<const float2*>(a.s.absmax + sb));
                    r.s0 = f.x;
                    r.s1 = f.y;
                }
            }
        } else {
#ifdef Q4_GEMV_EXPERIMENT_NOLOAD  // developer experiment (wrong results): only the first tiles are loaded -> pure SM-side rate
            if (c.t < 3 * nw)
#endif
            {
                r.wa = ldg_stream_256(pw);
                r.wb = ldg_stream_256(pw + 128);
            }
            if (NESTED) {
                r.q = __ldg(reinterpret_cast<const unsigned short*>(a.s.qabsmax + ab));
                r.s0 = __ldg(a.s.absmax2 + (ab >> a.s.shift2));
            } else {
                const float2 f = __ldg(reinterpret_cast<const float2*>(a.s.absmax + ab));
                r.s0 = f.x;
                r.s1 = f.y;
            }
            pw += a.pw_step;
            ab += a.ab_step;
            kt_ld += d_kt;
            if (kt_ld >= KT) {
                kt_ld -= KT;
                pw += a.pw_wrap;
                ab += a.ab_wrap;
            }
        }
    };

    if (stage == 0 && !a.lut) build_lut_in_kernel<T, NESTED>(lut, a.code, a.s.code2);  // callers without a prebuilt image
    // ---- the first kBuffers tiles go into registers now: under programmatic dependent launch this happens while the previous
    //      kernel is still computing (weights do not depend on it), so a small matrix is entirely on chip before x exists
    TileRegs r0, r1, r2;
    Cursor c0 = {warp, w_rt, w_kt};
    Cursor c1 = after(c0), c2 = after(c1);
    mma_trace_x(a, 1);
    if (c0.t < ntiles) issue(r0, c0);
    mma_trace_x(a, 2);
    if (c1.t < ntiles) issue(r1, c1);
    if (c2.t < ntiles) issue(r2, c2);
    mma_trace_x(a, 3);
    mma_trace(a, 1);

    // ---- everything below may read the previous kernel's (stage 0) or the previous stage's output
    if (!CHAIN || stage == 0) pdl_wait();
    else grid_barrier(c.barrier, (unsigned)stage * gridDim.x);
    mma_trace(a, 2);

    {
        const int nchunk = K >> 3;  // 16-byte chunks of x
        const int npad = KT * 64;   // staged chunks (zero tail up to whole tiles)
        if (a.x_gate || a.rms_weight) {
            stage_x_fused<T, true>(a.x, a.x_gate, a.rms_weight, a.rms_eps, K, s_x, s_red, nchunk, npad);
        } else {
            // chunk c = (block c>>3, piece c&7) goes to piece (c&7) ^ (block&7): the eight lanes that later fetch eight
            // different blocks piece by piece hit eight different bank groups
            for (int cb = tid; cb < npad; cb += 4 * nthr) {
                uint4 v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = cb + j * nthr;
                    v[j] = make_uint4(0, 0, 0, 0);
                    if (c < nchunk) v[j] = __ldcg(reinterpret_cast<const uint4*>(a.x) + c);  // coherent: may be a previous stage's output
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = cb + j * nthr;
                    if (c < npad) s_x[(c & ~7) | ((c ^ (c >> 3)) & 7)] = v[j];
                }
            }
        }
    }
    mma_trace(a, 3);
    __syncthreads();
    if (stage == 0 && a.lut) {  // table landed?
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" ::"r"(bar)
            : "memory");
    }
    mma_trace(a, 6);

    // ---- main loop
    const uint32_t lane_base = lut_saddr | (uint32_t)(lane * 4);
    const bool xrole = t4 == (g & 3);  // this lane feeds column g of the B operand: x of the tile's block g
    const uint32_t x_saddr = (uint32_t)__cvta_generic_to_shared(s_x), xswz = (uint32_t)(g * 16);
    uint32_t xr[32];
#pragma unroll
    for (int i = 0; i < 32; i++) xr[i] = 0;
    int kt_loaded = -1;

    auto compute = [&](const TileRegs& r, const Cursor& c) {
        if (c.kt != kt_loaded) {  // warp-uniform
            if (xrole) {
                // chunk i of block kt*8+g sits at piece i ^ g: byte offset (kt*1024 + g*128 + i*16) ^ (g*16)
                const uint32_t base = x_saddr + (uint32_t)(c.kt * 1024 + g * 128);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"((base | (uint32_t)(i * 16)) ^ xswz));
                    xr[4 * i] = v.x; xr[4 * i + 1] = v.y; xr[4 * i + 2] = v.z; xr[4 * i + 3] = v.w;
                }
            }
            kt_loaded = c.kt;
        }
        // MMA j covers bytes 2j, 2j+1 of both chunks = k 4j .. 4j+3 of the lane's blocks.  The lookups run kAhead MMAs ahead of
        // the tensor pipe (software pipeline, everything unrolled): a warp then has 4*kAhead shared-memory loads in flight
        // instead of stalling on each group of four.
        float ce[4] = {0.0f, 0.0f, 0.0f, 0.0f}, co[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#ifdef Q4_KAHEAD
        constexpr int kAhead = Q4_KAHEAD;
#else
        constexpr int kAhead = 3;
#endif
        uint32_t f[kAhead + 1][4];
        auto fetch = [&](uint32_t (&d)[4], int j) {
            const uint32_t wa = r.wa.v[j >> 1], wb = r.wb.v[j >> 1];
            if (j & 1) {
                d[0] = lut_lookup<2, kImm>(wa, lane_base); d[1] = lut_lookup<2, kImm>(wb, lane_base);
                d[2] = lut_lookup<3, kImm>(wa, lane_base); d[3] = lut_lookup<3, kImm>(wb, lane_base);
            } else {
                d[0] = lut_lookup<0, kImm>(wa, lane_base); d[1] = lut_lookup<0, kImm>(wb, lane_base);
                d[2] = lut_lookup<1, kImm>(wa, lane_base); d[3] = lut_lookup<1, kImm>(wb, lane_base);
            }
        };
#pragma unroll
        for (int j = 0; j < kAhead; j++) fetch(f[j], j);
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (j + kAhead < 16) fetch(f[(j + kAhead) % (kAhead + 1)], j + kAhead);
            uint32_t(&a4)[4] = f[j % (kAhead + 1)];
#ifdef Q4_GEMV_EXPERIMENT_FMA  // developer experiment (wrong results): the same lookups feeding the FMA pipe instead of the tensor pipe
            {
                uint32_t* c32 = reinterpret_cast<uint32_t*>(j & 1 ? co : ce);
                asm("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(c32[0]) : "r"(a4[0]), "r"(xr[2 * j]));
                asm("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(c32[1]) : "r"(a4[1]), "r"(xr[2 * j]));
                asm("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(c32[2]) : "r"(a4[2]), "r"(xr[2 * j + 1]));
                asm("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(c32[3]) : "r"(a4[3]), "r"(xr[2 * j + 1]));
            }
#else
            if (j & 1) Hmma<T>::run(co, a4[0], a4[1], a4[2], a4[3], xr[2 * j], xr[2 * j + 1]);
            else Hmma<T>::run(ce, a4[0], a4[1], a4[2], a4[3], xr[2 * j], xr[2 * j + 1]);
#endif
        }
        // the lane's two useful sums: blocks 2t, 2t+1 of the tile (columns 2t, 2t+1; first half -> rows 0-7, second -> 8-15)
        const float u0 = t4 < 2 ? ce[0] + co[0] : ce[2] + co[2];
        const float u1 = t4 < 2 ? ce[1] + co[1] : ce[3] + co[3];
        float am0, am1;
        if (NESTED) {
            float o = off[0];
            if (MULTI) {
                const int row = (rt0 + c.rt) * 8 + g;
                if (SWIGLU) o = (g & 4) ? off[1] : off[0];  // 8-row tiles are aligned to the 4 + 4 interleave
                else o = row < a.row_end[0] ? off[0] : (row < a.row_end[1] ? off[1] : (row < a.row_end[2] ? off[2] : off[3]));
            }
            const float q0 = __uint_as_float(lut_lookup<0, kImm + 128>(r.q, lane_base));
            const float q1 = __uint_as_float(lut_lookup<1, kImm + 128>(r.q, lane_base));
            am0 = __fadd_rn(__fmul_rn(q0, r.s0), o);  // reference: kernels.cu:552 then core.py:468
            am1 = __fadd_rn(__fmul_rn(q1, r.s0), o);
        } else {
            am0 = r.s0;
            am1 = r.s1;
        }
        float part = fmaf(u0, am0, u1 * am1);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        if (t4 == 0) s_part[c.t * 8 + g] = part;  // tile index = rt * KT + kt
    };

    for (;;) {  // three register buffers in rotation: compute the oldest, refill it with the tile after the newest
        if (c0.t >= ntiles) break;
        compute(r0, c0);
        c0 = after(c2);
        if (c0.t < ntiles) issue(r0, c0);
        if (c1.t >= ntiles) break;
        compute(r1, c1);
        c1 = after(c0);
        if (c1.t < ntiles) issue(r1, c1);
        if (c2.t >= ntiles) break;
        compute(r2, c2);
        c2 = after(c1);
        if (c2.t < ntiles) issue(r2, c2);
    }
    mma_trace(a, 4);
    __syncthreads();

    // ---- fixed-order sum over the k tiles, [all-reduce over tensor-parallel ranks,] bias / residual, store
    auto row_total = [&](int i) {
        const float* p = s_part + (i >> 3) * KT * 8 + (i & 7);
        float total = 0.0f;
        for (int kt = 0; kt < KT; kt++) total += p[kt * 8];
        return total;
    };
    auto finish = [&](int i, float total) {
        const int r = row_lo + i;
        T y = Elem<T>::from_f32(total);
        const T* bias = reinterpret_cast<const T*>(a.bias);
        if (bias) y = Elem<T>::from_f32(Elem<T>::to_f32(y) + Elem<T>::to_f32(__ldcg(bias + r)));  // torch `out += bias`
        reinterpret_cast<T*>(a.out)[r] = y;
    };
    const int nrows = row_hi - row_lo;
    if (a.ar_world > 1) {
        // One-shot all-reduce in the epilogue (include/quantizations_b200.h: q4_allreduce_t), "low latency" style: every partial
        // travels as ONE 8-byte store {value, epoch} into every peer's exchange area, and the owner of a row polls the W slots of
        // that row until they carry the current epoch -- no flag, no fence, no barrier: one NVLink hop.  Exchange area of a rank:
        //   [epochs u32 [kArMaxCtas]] ... [slots {f32, u32} [2][world][max_rows]] at kArDataOffset
        // The same CTA index owns the same rows on every rank (identical launch), so the exchange is CTA-local; launches are
        // counted per CTA on the device, so replayed CUDA graphs stay in step across ranks; slots are double-buffered by epoch
        // parity because a fast peer may start its NEXT exchange while this rank still reads the current one.
        const int W = a.ar_world, me = a.ar_rank;
        uint8_t* mine = reinterpret_cast<uint8_t*>(a.ar_peer_bases[me]);
        uint32_t* s_epoch = reinterpret_cast<uint32_t*>(s_red);
        if (tid == 0) {
            uint32_t* ep = reinterpret_cast<uint32_t*>(mine) + blockIdx.x;
            const uint32_t e = *ep + 1;
            *ep = e;
            *s_epoch = e;
        }
        __syncthreads();
        const uint32_t epoch = *s_epoch;
        const size_t half = (size_t)(epoch & 1) * W * a.ar_max_rows;
        for (int i = tid; i < nrows; i += nthr) {
            const float total = row_total(i);
            for (int p = 0; p < W; p++) {
                uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(a.ar_peer_bases[p]) + kArDataOffset) + half +
                             (size_t)me * a.ar_max_rows + row_lo + i;
                asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(total)), "r"(epoch) : "memory");
            }
        }
        const uint2* slots = reinterpret_cast<const uint2*>(mine + kArDataOffset) + half;
        for (int i = tid; i < nrows; i += nthr) {
            float total = 0.0f;
            for (int p = 0; p < W; p++) {  // rank order: every rank computes the same sum
                const uint2* src = slots + (size_t)p * a.ar_max_rows + row_lo + i;
                uint32_t v, e;
                for (long long spin = 0;; spin++) {
                    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(e) : "l"(src) : "memory");
                    if (e == epoch) break;
                    if (spin > (1ll << 26)) __trap();  // a peer never arrived: fail loudly instead of hanging the GPU
                }
                total += __uint_as_float(v);
            }
            finish(i, total);
        }
    } else if (SWIGLU) {
        // SwiGLU in the epilogue: every 8-row tile holds four (gate, up) pairs, so this CTA owns both halves of each of its outputs.
        // Rounded exactly as the separate steps would: gate and up to T (what the plain launch stores), silu(gate) to T, the product
        // to T -- the same expressions as stage_x_fused, hence bit-identical to "grouped gate/up launch, then SwiGLU staging".
        T* hout = reinterpret_cast<T*>(a.out);
        for (int i = tid; i < (nrows >> 1); i += nthr) {
            const int t8 = (i >> 2) * 8, m = i & 3;
            const float gq = Elem<T>::to_f32(Elem<T>::from_f32(row_total(t8 + m)));
            const float uq = Elem<T>::to_f32(Elem<T>::from_f32(row_total(t8 + 4 + m)));
            const float sg = Elem<T>::to_f32(Elem<T>::from_f32(__fdividef(gq, 1.0f + __expf(-gq))));
            hout[(row_lo >> 1) + i] = Elem<T>::from_f32(sg * uq);
        }
    } else {
        for (int i = tid; i < nrows; i += nthr) finish(i, row_total(i));
    }
    mma_trace(a, 5);
  }  // stage loop

    if (CHAIN && c.n > 1) {  // leave the barrier counter at zero for the next launch: the last CTA to get here resets it
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const unsigned old = atomicAdd(c.barrier, 1u);
            if (old == (unsigned)c.n * gridDim.x - 1) *c.barrier = 0;
        }
    }
}

static __global__ void probe_dyn_smem_base_kernel(uint32_t* out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    out[0] = (uint32_t)__cvta_generic_to_shared(smem);
}

// 64-KB table image for (code, code2, T): row b = 32 x pair{code[b>>4], code[b&15]} as T, then 32 x code2[b] (fp32)
template <typename T>
__global__ void gemv_lut_build_kernel(const float* __restrict__ code, const float* __restrict__ code2, uint32_t* __restrict__ lut)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // word index
    if (i >= kLutBytes / 4) return;
    const int b = i >> 6;
    lut[i] = (i & 32) ? (code2 ? __float_as_uint(code2[b]) : 0u) : pack2<T>(code[b >> 4], code[b & 15]);
}

}  // namespace q4
