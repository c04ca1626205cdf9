// q4_gemv.cu -- batch-1 decode GEMV over a packed 4-bit weight:   out[r] = sum_k x[k] * code[nib(r,k)] * absmax[blk(r,k)]
//
// Replaces: reference csrc/kernels.cu:1061-1219 (kgemm_4bit_inference_naive) + launcher csrc/ops.cu:167-171, AND the
// two launches the reference issues before it on every call (dequantize_blockwise of the 8-bit absmax + the torch
// `absmax += offset`, core.py:467-468): the double-quant decode happens inside the kernel.
// Contract: within tolerance of the reference / fp64 truth (summation order differs by design); the decoded per-block
// absmax is bit-identical to the reference's (fp32 multiply then fp32 add).
// Roofline: HBM.  Algorithmic bytes per call = N*K/2 + N*K/bs (+ nested table, x, out) -- see DESIGN.md.
//
// Two kernels:
//
//  gemv_lut256_kernel  (fast path: blocksize 64, K % 64 == 0)
//    The reference spends one shared-memory lookup + one multiply per NIBBLE; at B200's bytes-per-clock that is
//    shared-memory- and issue-bound long before HBM.  Here one lookup serves a whole BYTE: a 256-entry table of
//    half2{code[b>>4], code[b&15]}, replicated once per lane (row stride 256 B, lane l reads word l of its row) so
//    that every access is bank-conflict free, and indexed by a single PRMT that splices the weight byte into the
//    address.  Per packed byte the inner loop is PRMT + LDS.32 + HFMA2 (two weights), i.e. 1.5 issue slots per
//    weight instead of ~4.  The other half of each 256-B row holds code2[b] (fp32, per lane) so the 8-bit absmax
//    decode is the same conflict-free lookup.
//    A thread owns a fixed 64-wide k-slice (exactly one quantisation block per row): its 64 activations stay in
//    registers as half2 (pre-scaled by a per-thread power of two so fp16 cannot overflow), it streams one 256-bit
//    load (one full DRAM sector) per row, keeps U rows in flight, accumulates 8 half2 products per chain before
//    widening to fp32, and multiplies by the block's absmax in fp32.  Rows are dealt round-robin to "groups"
//    (kw = ceil(K/2048) warps that together cover one row); U row-sums are reduced across lanes with a
//    transposing butterfly (9 shuffles for 4 rows instead of 20) and across the kw warps through shared memory
//    with a named barrier per group.
//    Programmatic dependent launch: weight loads and the table build are issued BEFORE griddepcontrol.wait, so in a
//    decode chain they overlap the previous layer's tail; only x is read after the wait.
//
//  gemv_generic_kernel (any even K, any valid blocksize, "exact" fp32 arithmetic: w = code*absmax, acc = fma(x,w,acc)
//    as the reference's T=float instance does) -- warp per row, x staged in shared memory as fp32.
#include <cstdlib>
#include <type_traits>

#include "q4_common.cuh"
#include "q4_launch.h"

namespace q4 {

// ------------------------------------------------------------------------------------------------ fast path

constexpr int kLutBytes = 65536;   // 256 rows x 256 B
constexpr int kRowsInFlight = 4;   // U

template <typename T> __device__ __forceinline__ void load_x64(const T* x, float (&xf)[64]);
template <> __device__ __forceinline__ void load_x64<float>(const float* x, float (&xf)[64])
{
#pragma unroll
    for (int j = 0; j < 16; j++) {
        float4 v = __ldg(reinterpret_cast<const float4*>(x) + j);
        xf[4 * j] = v.x; xf[4 * j + 1] = v.y; xf[4 * j + 2] = v.z; xf[4 * j + 3] = v.w;
    }
}
template <typename T> __device__ __forceinline__ void load_x64(const T* x, float (&xf)[64])
{
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + j);
        float2 a = unpack2<T>(v.x), b = unpack2<T>(v.y), c = unpack2<T>(v.z), d = unpack2<T>(v.w);
        xf[8 * j] = a.x; xf[8 * j + 1] = a.y; xf[8 * j + 2] = b.x; xf[8 * j + 3] = b.y;
        xf[8 * j + 4] = c.x; xf[8 * j + 5] = c.y; xf[8 * j + 6] = d.x; xf[8 * j + 7] = d.y;
    }
}

__device__ __forceinline__ uint32_t lut_at(const uint8_t* lut, uint32_t byte_offset)
{
    return *reinterpret_cast<const uint32_t*>(lut + byte_offset);
}
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ void named_barrier(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// A launch may cover up to kMaxMats weight matrices that share x and are stored back to back (packed bytes, 8-bit
// absmax, outputs): q/k/v or gate/up of a decoder layer.  To the kernel they are one [rows, K] matrix; only the
// double-quant offset (one scalar per quantize_4bit call) differs by row range.
constexpr int kMaxMats = 4;
struct GemvArgs {
    const void* x;
    const float* code;         // 16-entry 4-bit code table
    const uint8_t* Bq;         // packed weight [rows, K/2]
    AbsmaxView s;              // statistics (offset unused: see offsets[])
    const float* offsets[kMaxMats];  // nested: per-matrix offset scalars (device pointers)
    int row_end[kMaxMats];     // exclusive end row of each matrix (INT_MAX for unused slots)
    void* out;                 // [rows]
    const void* bias;          // [rows] or nullptr (may alias `out`: a residual stream updated in place)
    const void* x_gate;        // optional: the effective activation is silu(x_gate[k]) * x[k]   (SwiGLU, fused)
    const void* rms_weight;    // optional: the effective activation is x * rsqrt(mean(x^2) + rms_eps) * rms_weight  (RMSNorm, fused)
    float rms_eps;
    const uint8_t* next;       // optional: bytes the NEXT launch will stream (pulled into L2 while this one computes)
    int64_t next_bytes;
    int rows, K;
    int kw;      // warps that together cover one row: ceil(K / 2048)
    int groups;  // row groups per CTA: blockDim / (32 * kw)
    int rows_per_cta;
    int lut_iters;   // ceil(4096 / blockDim)
    int x_iters;     // ceil(K / 8 / blockDim): raw activation chunks per thread
    int debug_mode;             // developer experiments (env Q4_GEMV_DEBUG): 1 = skip the table lookups
    unsigned long long* trace;  // debug: per-CTA phase timestamps (globaltimer ns), 8 slots per CTA; nullptr = off
};

__device__ __forceinline__ void trace_mark(const GemvArgs& a, int slot)
{
    if (a.trace && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.trace[blockIdx.x * 8 + slot] = t;
    }
}

// Shared-memory plan.  ONE PRMT splices {window high half, weight byte, lane*4} into an LDS address, which requires the
// table to start at a constant offset from a 64-KB boundary of the CTA's shared window; the constant goes into the
// LDS immediate.  Two layouts:
//   COMPACT  the table sits at the very start of dynamic shared memory, which on sm_100 begins kDynBase = 1 KB into the
//            window (probed once on the host, see dyn_smem_base()); [table 64K | x K*2 | partials | words | scratch].
//            ~75-95 KB per CTA, so a 256-thread CTA leaves room for the NEXT kernel's CTA on the same SM: with
//            programmatic dependent launch its prologue (table build, L2 prefetch) overlaps this kernel's main loop.
//   ALIGNED  fallback if the probe disagrees: 128 KB requested, table at the 64-KB boundary inside it.
constexpr int kSmemAligned = 128 * 1024;
constexpr int kDynBase = 1024;

template <int IMM> __device__ __forceinline__ uint32_t lds_u32(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(saddr), "n"(IMM));
    return v;
}
template <int IMM> __device__ __forceinline__ float lds_f32(uint32_t saddr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(saddr), "n"(IMM));
    return v;
}
// TMA bulk prefetch of a byte range into L2 (UBLKPF.L2): fire-and-forget, no registers, no completion to wait for.
// `bytes` is rounded down to a multiple of 16; `p` must be 16-byte aligned.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes)
{
    bytes &= ~15u;
    if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// spread a range over the lanes of one warp in 8-KB pieces
__device__ __forceinline__ void prefetch_l2_range(const uint8_t* p, int64_t bytes, int lane)
{
    constexpr int64_t kPiece = 8192;
    const int64_t skew = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;  // start at the next 16-byte boundary
    p += skew;
    bytes -= skew;
    for (int64_t o = (int64_t)lane * kPiece; o < bytes; o += 32 * kPiece)
        prefetch_l2_bulk(p + o, (uint32_t)(bytes - o < kPiece ? bytes - o : kPiece));
}

__global__ void probe_dyn_smem_base(uint32_t* out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    out[0] = (uint32_t)__cvta_generic_to_shared(smem);
}

template <typename T, bool NESTED, bool MULTI, bool COMPACT>
__global__ void __launch_bounds__(512, 1)
gemv_lut256_kernel(const GemvArgs a)
{
    constexpr int U = kRowsInFlight;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int K = a.K, kw = a.kw, groups = a.groups, R = a.rows;
    const uint32_t smem_saddr = (uint32_t)__cvta_generic_to_shared(smem);
    // table address = lut_saddr + kImm + byte*256 + lane*4, with lut_saddr 64-KB aligned and kImm an LDS immediate
    constexpr int kImm = COMPACT ? kDynBase : 0;
    const uint32_t lut_saddr = COMPACT ? (smem_saddr & 0xFFFF0000u) : ((smem_saddr + 0xFFFFu) & 0xFFFF0000u);
    uint8_t* lut = COMPACT ? smem : smem + (lut_saddr - smem_saddr);
    uint8_t* rest = COMPACT ? smem + kLutBytes : smem;
    uint4* s_x = reinterpret_cast<uint4*>(rest);
    float* s_part = reinterpret_cast<float*>(rest + 2 * K);
    uint32_t* s_words = reinterpret_cast<uint32_t*>(s_part + 2 * groups * kw * U);
    float* s_red = reinterpret_cast<float*>(s_words + 512);

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int lgroup = warp / kw;        // row group inside the CTA
    const int kpos = warp - lgroup * kw; // which 2048-wide K slice this warp covers
    const int bpr = K >> 6;              // 64-wide blocks per row
    const int kblk = kpos * 32 + lane;   // this thread's block inside a row
    const bool active = kblk < bpr;
    const int kblk_c = active ? kblk : bpr - 1;  // inactive lanes load a valid address and multiply by x = 0

    // This CTA owns the contiguous rows [row_lo, row_hi): one contiguous byte range of the packed weight.
    const int row_lo = blockIdx.x * a.rows_per_cta;
    const int row_hi = row_lo + a.rows_per_cta < R ? row_lo + a.rows_per_cta : R;

    // Let the next kernel in the stream start its own prologue (its griddepcontrol.wait still orders it after us).
    pdl_launch_dependents();
    trace_mark(a, 0);

    // ---- 0b. the small latency-critical loads go out FIRST (requests are served in order: behind 64 KB of weight
    //          loads they would wait microseconds): code table, code2 table, offsets.  Word t < 256 of the staging area is
    //          half2{code[t>>4], code[t&15]}, word 256 + t is code2[t]; a thread carries up to 4 of the 512 words.
    float tab_a[4], tab_b[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int t = tid + q * nthr;
        tab_a[q] = tab_b[q] = 0.0f;
        if (t < 256) {
            tab_a[q] = __ldg(a.code + (t >> 4));
            tab_b[q] = __ldg(a.code + (t & 15));
        } else if (NESTED && t < 512) {
            tab_a[q] = __ldg(a.s.code2 + (t - 256));
        }
    }
    float off[kMaxMats];
#pragma unroll
    for (int m = 0; m < kMaxMats; m++) off[m] = (NESTED && (MULTI || m == 0) && a.offsets[m]) ? __ldg(a.offsets[m]) : 0.0f;

    // ---- 0. stream the whole slice HBM -> L2 now, decoupled from the SM's own progress: the demand loads below then
    //         see L2 latency, and HBM has the entire matrix queued from the first microsecond.  Optionally also this
    //         CTA's share of the bytes the NEXT launch will read (weights of the following Linear).
    const bool do_prefetch = a.debug_mode != 2 && a.debug_mode < 16;
    if (do_prefetch && warp == (nthr >> 5) - 1 && row_lo < row_hi) {
        prefetch_l2_range(a.Bq + (int64_t)row_lo * (K >> 1), (int64_t)(row_hi - row_lo) * (K >> 1), lane);
        if (NESTED) prefetch_l2_range(a.s.qabsmax + (int64_t)row_lo * bpr, (int64_t)(row_hi - row_lo) * bpr, lane);
    }
    if (do_prefetch && warp == (nthr >> 5) - 2 && a.next_bytes > 0) {
        const int64_t share = ((a.next_bytes / gridDim.x) + 15) & ~(int64_t)15;
        const int64_t lo = share * blockIdx.x;
        const int64_t n = lo + share <= a.next_bytes ? share : a.next_bytes - lo;
        if (n > 0) prefetch_l2_range(a.next + lo, n, lane);
    }

    // ---- 1. row bookkeeping.  Group g of the CTA takes rows row_lo + g, + groups, + 2*groups ...: every address advances
    //         by a constant stride.  The first U rows are only requested AFTER the prologue's shared-memory work (step 4):
    //         64 KB of outstanding LDG.256 per SM fill the load/store unit's queues and stall every LDS/STS behind them
    //         (measured: the table build took 2-3 us that way); the L2 prefetch of step 0 has the data on its way already.
    const int first = row_lo + lgroup;
    const int rows_mine = first < row_hi ? (row_hi - first + groups - 1) / groups : 0;
    int blk = first * bpr + kblk_c;      // block index of the next row to issue (rows * bpr < 2^31: dispatcher)
    const int blk_stride = groups * bpr;
    u32x8 w[U];
    uint32_t qa[U];
    float a2[U];
    const int pf_rows = a.debug_mode >= 16 ? a.debug_mode - 16 : 0;  // experiment: per-row L2 prefetch distance (rows)
    const int blk_end = (row_hi - 1) * bpr + kblk_c;                  // this thread's block in the slice's last row
    auto issue_row = [&](int i) {
        w[i] = ldg_stream_256(a.Bq + (int64_t)blk * 32);
        if (pf_rows) {
            const int pb = blk + pf_rows * blk_stride;
            if (pb <= blk_end) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Bq + (int64_t)pb * 32));
        }
        if (NESTED) {
            qa[i] = __ldg(a.s.qabsmax + blk);
            a2[i] = __ldg(a.s.absmax2 + (blk >> a.s.shift2));
        } else {
            a2[i] = __ldg(a.s.absmax + blk);
        }
        blk += blk_stride;
    };
    trace_mark(a, 1);
    // ---- 3. lookup table: stage the 512 distinct words (loaded in step 0b), then replicate each 32x:
    //         128-B segment 2b = half2{code[b>>4], code[b&15]}, segment 2b+1 = code2[b] (fp32).
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int t = tid + q * nthr;
        if (t < 256) {
            __half2 h = __halves2half2(__float2half_rn(tab_a[q]), __float2half_rn(tab_b[q]));
            s_words[t] = *reinterpret_cast<uint32_t*>(&h);
        } else if (t < 512) {
            s_words[t] = __float_as_uint(tab_a[q]);
        }
    }
    __syncthreads();
    trace_mark(a, 7);
    for (int it = 0, c = tid; it < a.lut_iters; it++, c += nthr) {  // trip count from the host: no integer division here
        if (c >= kLutBytes / 16) break;
        const int seg = c >> 3;
        if (NESTED || !(seg & 1)) {
            const uint32_t word = s_words[(seg >> 1) | ((seg & 1) << 8)];
            *reinterpret_cast<uint4*>(lut + c * 16) = make_uint4(word, word, word, word);
        }
    }
    // ---- 3b. first U rows in flight: after the table build (see step 1), before the dependency wait -- under programmatic
    //          dependent launch they stream in while the previous kernel is still running.
#pragma unroll
    for (int i = 0; i < U; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) w[i].v[j] = 0;
        qa[i] = 0;
        a2[i] = 0.0f;
        if (i < rows_mine) issue_row(i);
    }
    trace_mark(a, 2);

    // ---- everything below may read the previous kernel's output
    pdl_wait();
    trace_mark(a, 3);
    const T* xg = reinterpret_cast<const T*>(a.x);
    const int nchunk = K >> 3;  // 8-element (16-byte for 16-bit types) chunks of x
    constexpr int XR = 8;       // raw chunks a thread may hold: covers K <= 64 * blockDim (dispatcher)
    constexpr bool kRawX = sizeof(T) == 2;
    uint4 xraw[kRawX ? XR : 1];
    if constexpr (kRawX) {
#pragma unroll
        for (int j = 0; j < XR; j++) {
            const int c = tid + j * nthr;
            xraw[j] = make_uint4(0, 0, 0, 0);
            if (j < a.x_iters && c < nchunk) xraw[j] = __ldg(reinterpret_cast<const uint4*>(xg) + c);
        }
        // ---- fused input transforms (decode glue that would otherwise be separate launches and HBM round trips).  Both
        //      leave a "virtual x" in xraw, rounded to T exactly where the separate torch kernels round.
        using T16 = typename std::conditional<sizeof(T) == 2, T, __half>::type;
        if (a.x_gate) {  // SwiGLU: silu(gate) * up, F.silu then multiply, each rounded to T
#pragma unroll
            for (int j = 0; j < XR; j++) {
                const int c = tid + j * nthr;
                if (j < a.x_iters && c < nchunk) {
                    const uint4 g4 = __ldg(reinterpret_cast<const uint4*>(a.x_gate) + c);
                    const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
                    uint32_t uw[4] = {xraw[j].x, xraw[j].y, xraw[j].z, xraw[j].w};
#pragma unroll
                    for (int q2 = 0; q2 < 4; q2++) {
                        const float2 g = unpack2<T16>(gw[q2]), u = unpack2<T16>(uw[q2]);
                        const float2 sg = unpack2<T16>(pack2<T16>(g.x / (1.0f + expf(-g.x)), g.y / (1.0f + expf(-g.y))));
                        uw[q2] = pack2<T16>(sg.x * u.x, sg.y * u.y);
                    }
                    xraw[j] = make_uint4(uw[0], uw[1], uw[2], uw[3]);
                }
            }
        }
        if (a.rms_weight) {  // RMSNorm in fp32 over the whole vector (every CTA stages all of x), rounded to T once
            float ss = 0.0f;
#pragma unroll
            for (int j = 0; j < XR; j++) {
                if (j < a.x_iters) {
                    const uint32_t xw[4] = {xraw[j].x, xraw[j].y, xraw[j].z, xraw[j].w};
#pragma unroll
                    for (int q2 = 0; q2 < 4; q2++) {
                        const float2 f = unpack2<T16>(xw[q2]);
                        ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0) s_red[16 + warp] = ss;
            __syncthreads();
            ss = 0.0f;
            for (int i = 0; i < (nthr >> 5); i++) ss += s_red[16 + i];
            const float rs = rsqrtf(ss / (float)K + a.rms_eps);
#pragma unroll
            for (int j = 0; j < XR; j++) {
                const int c = tid + j * nthr;
                if (j < a.x_iters && c < nchunk) {
                    const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(a.rms_weight) + c);
                    const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
                    uint32_t xw[4] = {xraw[j].x, xraw[j].y, xraw[j].z, xraw[j].w};
#pragma unroll
                    for (int q2 = 0; q2 < 4; q2++) {
                        const float2 f = unpack2<T16>(xw[q2]), g = unpack2<T16>(ww[q2]);
                        xw[q2] = pack2<T16>(f.x * rs * g.x, f.y * rs * g.y);
                    }
                    xraw[j] = make_uint4(xw[0], xw[1], xw[2], xw[3]);
                }
            }
        }
    }

    // ---- 4. x -> half2, scaled by one power of two so that max|x| lands in [1,2) (fp16 cannot overflow; the products
    //         are accumulated 8 deep in fp16, then in fp32).  Done once per CTA from the raw chunks already in registers:
    //         CTA-wide max, convert, store swizzled; then each thread fetches the 64 values of its own k-slice.
    auto load_chunk = [&](int c, float (&v)[8]) {
        if constexpr (sizeof(T) == 2) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(xg) + c);
            const float2 f0 = unpack2<T>(q.x), f1 = unpack2<T>(q.y), f2 = unpack2<T>(q.z), f3 = unpack2<T>(q.w);
            v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y; v[4] = f2.x; v[5] = f2.y; v[6] = f3.x; v[7] = f3.y;
        } else {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(xg) + 2 * c);
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(xg) + 2 * c + 1);
            v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w; v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
        }
    };
    auto unpack_raw = [&](const uint4& q, float (&v)[8]) {
        using T16 = typename std::conditional<sizeof(T) == 2, T, __half>::type;
        const float2 f0 = unpack2<T16>(q.x), f1 = unpack2<T16>(q.y), f2 = unpack2<T16>(q.z), f3 = unpack2<T16>(q.w);
        v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y; v[4] = f2.x; v[5] = f2.y; v[6] = f3.x; v[7] = f3.y;
    };
    float m = 0.0f;
    if constexpr (kRawX) {
#pragma unroll
        for (int j = 0; j < XR; j++) {
            if (j < a.x_iters) {  // uniform: K = 4096 needs one chunk per thread, not eight
                float v[8];
                unpack_raw(xraw[j], v);  // chunks past the end are zeros
#pragma unroll
                for (int e8 = 0; e8 < 8; e8++) m = fmaxf(m, fabsf(v[e8]));
            }
        }
    } else {
        for (int c = tid; c < nchunk; c += nthr) {
            float v[8];
            load_chunk(c, v);
#pragma unroll
            for (int j = 0; j < 8; j++) m = fmaxf(m, fabsf(v[j]));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_red[warp] = m;
    __syncthreads();  // also: lookup table visible
    m = 0.0f;
    for (int i = 0; i < (nthr >> 5); i++) m = fmaxf(m, s_red[i]);
    int e = (int)((__float_as_uint(m) >> 23) & 0xFF);  // biased exponent of the largest |x|
    e = e < 1 ? 1 : (e > 253 ? 253 : e);
    const float scale = __uint_as_float((uint32_t)(254 - e) << 23);  // 2^(127-e)
    const float unscale = __uint_as_float((uint32_t)e << 23);        // 2^(e-127)
    auto store_chunk = [&](int c, const float (&v)[8]) {
        uint32_t h[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            __half2 hh = __floats2half2_rn(v[2 * j] * scale, v[2 * j + 1] * scale);
            h[j] = *reinterpret_cast<uint32_t*>(&hh);
        }
        const int kb = c >> 3, j = c & 7;
        s_x[kb * 8 + (j ^ (kb & 7))] = make_uint4(h[0], h[1], h[2], h[3]);  // swizzle: conflict-free both ways
    };
    if constexpr (kRawX) {
#pragma unroll
        for (int j = 0; j < XR; j++) {
            const int c = tid + j * nthr;
            if (j < a.x_iters && c < nchunk) {
                float v[8];
                unpack_raw(xraw[j], v);
                store_chunk(c, v);
            }
        }
    } else {
        for (int c = tid; c < nchunk; c += nthr) {  // second pass hits L1
            float v[8];
            load_chunk(c, v);
            store_chunk(c, v);
        }
    }

    __syncthreads();
    uint32_t xh[32];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint4 v = s_x[kblk_c * 8 + (j ^ (kblk_c & 7))];
        if (!active) v = make_uint4(0, 0, 0, 0);
        xh[4 * j] = v.x; xh[4 * j + 1] = v.y; xh[4 * j + 2] = v.z; xh[4 * j + 3] = v.w;
    }

    trace_mark(a, 4);
    // PRMT operand: {lane*4, 0, window bits 16-23, window bits 24-31}; selector 0x76i4 splices weight byte i into
    // byte 1 -> lut_saddr + byte*256 + lane*4, a complete shared address
    const uint32_t lane_base = lut_saddr | (uint32_t)(lane * 4);
    const int nbatch = (rows_mine + U - 1) / U;
    int buf = 0;

    for (int batch = 0; batch < nbatch; batch++) {
        float t[U];
        const int n0 = batch * U;
#pragma unroll
        for (int i = 0; i < U; i++) {
            t[i] = 0.0f;
            if (n0 + i < rows_mine) {  // warp-uniform: rows past the end cost nothing
                uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
                if (a.debug_mode == 1) {
#pragma unroll
                    for (int j = 0; j < 8; j++) acc0 ^= w[i].v[j];
                } else
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t wj = w[i].v[j];
                    acc0 = hfma2(lds_u32<kImm>(__byte_perm(wj, lane_base, 0x7604)), xh[4 * j + 0], acc0);
                    acc1 = hfma2(lds_u32<kImm>(__byte_perm(wj, lane_base, 0x7614)), xh[4 * j + 1], acc1);
                    acc2 = hfma2(lds_u32<kImm>(__byte_perm(wj, lane_base, 0x7624)), xh[4 * j + 2], acc2);
                    acc3 = hfma2(lds_u32<kImm>(__byte_perm(wj, lane_base, 0x7634)), xh[4 * j + 3], acc3);
                }
                const uint32_t sum = hadd2(hadd2(acc0, acc1), hadd2(acc2, acc3));
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&sum));
                float am;
                if (NESTED) {
                    float o = off[0];
                    if (MULTI) {
                        const int r = first + (n0 + i) * groups;
                        o = r < a.row_end[0] ? off[0] : (r < a.row_end[1] ? off[1] : (r < a.row_end[2] ? off[2] : off[3]));
                    }
                    const float c2 = lds_f32<kImm + 128>(__byte_perm(qa[i], lane_base, 0x7604));
                    am = __fadd_rn(__fmul_rn(c2, a2[i]), o);  // reference: kernels.cu:552 then core.py:468
                } else {
                    am = a2[i];
                }
                t[i] = (f.x + f.y) * (am * unscale);
                // register rotation: row i of the NEXT batch streams into the registers this row just vacated
                if (n0 + U + i < rows_mine) issue_row(i);
            }
        }

        // ---- lane reduction: U=4 row sums over 32 lanes with a transposing butterfly
        {
            const bool hi16 = lane & 16, hi8 = lane & 8;
            float keep0 = hi16 ? t[2] : t[0], keep1 = hi16 ? t[3] : t[1];
            float send0 = hi16 ? t[0] : t[2], send1 = hi16 ? t[1] : t[3];
            keep0 += __shfl_xor_sync(0xffffffffu, send0, 16);
            keep1 += __shfl_xor_sync(0xffffffffu, send1, 16);
            float keep = hi8 ? keep1 : keep0, send = hi8 ? keep0 : keep1;
            keep += __shfl_xor_sync(0xffffffffu, send, 8);
            keep += __shfl_xor_sync(0xffffffffu, keep, 4);
            keep += __shfl_xor_sync(0xffffffffu, keep, 2);
            keep += __shfl_xor_sync(0xffffffffu, keep, 1);
            // lanes 8i..8i+7 now hold the warp's sum for row i of the batch
            const int i = lane >> 3;
            float total = keep;
            if (kw > 1) {
                float* part = s_part + ((buf * groups + lgroup) * kw) * U;
                if ((lane & 7) == 0) part[kpos * U + i] = keep;
                named_barrier(1 + lgroup, kw * 32);
                if (kpos == 0) {
                    total = 0.0f;
                    for (int p = 0; p < kw; p++) total += part[p * U + i];
                }
                buf ^= 1;
            }
            if (kpos == 0 && (lane & 7) == 0 && n0 + i < rows_mine) {
                const int r = first + (n0 + i) * groups;
                T* out = reinterpret_cast<T*>(a.out);
                const T* bias = reinterpret_cast<const T*>(a.bias);
                T y = Elem<T>::from_f32(total);
                if (bias) y = Elem<T>::from_f32(Elem<T>::to_f32(y) + Elem<T>::to_f32(bias[r]));  // torch `out += bias`
                out[r] = y;
            }
        }
        if (batch == 0) trace_mark(a, 5);
    }
    trace_mark(a, 6);
}

unsigned long long* g_gemv_trace = nullptr;  // set by q4_debug_set_gemv_trace (developer tool, not part of the ABI)

// ------------------------------------------------------------------------------------------------ generic path

template <typename T, bool NESTED>
__global__ void __launch_bounds__(256)
gemv_generic_kernel(const T* __restrict__ x, const uint8_t* __restrict__ Bq, AbsmaxView s, const float* __restrict__ code,
                    const T* __restrict__ bias, T* __restrict__ out, int64_t N, int64_t K, int bs_shift)
{
    extern __shared__ __align__(16) uint8_t smem[];
    float* s_code = reinterpret_cast<float*>(smem);       // 16 entries, one per bank: conflict-free
    float* s_x = reinterpret_cast<float*>(smem) + 32;     // K floats
    if (threadIdx.x < 16) s_code[threadIdx.x] = __ldg(code + threadIdx.x);
    const float offset = NESTED ? __ldg(s.offset) : 0.0f;
    pdl_wait();
    pdl_launch_dependents();
    for (int64_t k = threadIdx.x; k < K; k += blockDim.x) s_x[k] = Elem<T>::to_f32(x[k]);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const bool vec = (K & 31) == 0 && ((reinterpret_cast<uintptr_t>(Bq) & 15) == 0);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < N; r += warps) {
        float acc = 0.0f;
        const uint8_t* row = Bq + ((r * K) >> 1);
        if (vec) {
            for (int64_t k0 = 32 * lane; k0 < K; k0 += 1024) {
                const uint4 pk = ldg_stream_128(row + (k0 >> 1));
                const float am = load_absmax<NESTED>(s, (r * K + k0) >> bs_shift, offset);
                const uint32_t wd[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float4 xa = *reinterpret_cast<const float4*>(s_x + k0 + 8 * j);
                    const float4 xb = *reinterpret_cast<const float4*>(s_x + k0 + 8 * j + 4);
                    const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const uint32_t byte = (wd[j] >> (8 * b)) & 0xFFu;
                        acc = fmaf(xv[2 * b], __fmul_rn(s_code[byte >> 4], am), acc);
                        acc = fmaf(xv[2 * b + 1], __fmul_rn(s_code[byte & 15], am), acc);
                    }
                }
            }
        } else {
            for (int64_t k = 2 * lane; k < K; k += 64) {  // one byte (two elements) per lane per step
                const int64_t e = r * K + k;
                const uint32_t byte = Bq[e >> 1];
                const float am = load_absmax<NESTED>(s, e >> bs_shift, offset);
                acc = fmaf(s_x[k], __fmul_rn(s_code[byte >> 4], am), acc);
                // element k+1 may fall into the next block only if blocksize were 1; blocksize >= 64 and e even
                acc = fmaf(s_x[k + 1], __fmul_rn(s_code[byte & 15], am), acc);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            T y = Elem<T>::from_f32(acc);
            if (bias) y = Elem<T>::from_f32(Elem<T>::to_f32(y) + Elem<T>::to_f32(bias[r]));
            out[r] = y;
        }
    }
}

// ------------------------------------------------------------------------------------------------ host dispatch

struct GemvPrologue {  // optional fused input transforms (16-bit activations only)
    const void* x_gate = nullptr;
    const void* rms_weight = nullptr;
    float rms_eps = 0.0f;
};

template <typename K, typename... Args>
static int launch_pdl(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    if (e != cudaSuccess) return (int)e;
    return finish_launch();
}

template <typename T>
static int gemv_dispatch(const T* x, const uint8_t* B, const q4_absmax_t* st, const float* code, const T* bias, T* out,
                         int64_t N, int64_t K, int blocksize, int flags, const void* next, int64_t next_bytes, cudaStream_t stream,
                         int nmat = 1, const float* const* offsets = nullptr, const int* row_end = nullptr,
                         const GemvPrologue* pro = nullptr)
{
    const AbsmaxView v = make_view(st);
    const bool nested = st->qabsmax != nullptr;
    const bool pdl = flags & Q4_GEMV_PDL;
    const int sms = sm_count();
    const int kw = (int)((K + 2047) / 2048);
    const bool fast = !(flags & Q4_GEMV_EXACT_F32) && blocksize == 64 && (K % 64) == 0 && kw <= 16 && K >= 64 && N * (K / 64) < (1ll << 31) && N < (1 << 30) &&
                      (reinterpret_cast<uintptr_t>(B) & 31) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                      (!nested || st->blocksize2 >= 64);
    if (fast) {
        // Warps per CTA.  16 (one CTA fills the SM) streams a large matrix fastest when the launch has the GPU to itself.
        // 8 (<= 256 threads x 128 registers, ~75-95 KB shared memory) leaves half of every SM free, so that a second launch
        // -- the next kernel's prologue under programmatic dependent launch, or an independent GEMV on another stream
        // (q/k/v, gate/up) -- is co-resident: chosen for small matrices and when the caller passes Q4_GEMV_SHARE_SM.
        static const int env_warps = getenv("Q4_GEMV_WARPS") ? atoi(getenv("Q4_GEMV_WARPS")) : 0;
        const int64_t warp_rows_per_sm = N * kw / sms;
        int target = env_warps ? env_warps : (((flags & Q4_GEMV_SHARE_SM) || warp_rows_per_sm <= 32) ? 8 : 16);
        const int warps = kw > target ? kw : target;
        int groups = warps / kw;
        if (groups * kw * 32 < 128) groups = (128 + kw * 32 - 1) / (kw * 32);  // the table staging covers 512 words with 4 per thread
        const int threads = groups * kw * 32;
        // shared-memory layout: COMPACT when dynamic shared memory starts kDynBase into the window (probed once)
        static int dyn_base = -1;
        if (dyn_base < 0) {
            uint32_t* d = nullptr;
            uint32_t h = 0;
            if (cudaMalloc(&d, 4) == cudaSuccess) {
                probe_dyn_smem_base<<<1, 32, 1024, stream>>>(d);
                if (cudaMemcpyAsync(&h, d, 4, cudaMemcpyDeviceToHost, stream) == cudaSuccess && cudaStreamSynchronize(stream) == cudaSuccess)
                    dyn_base = (int)(h & 0xFFFFu);
                cudaFree(d);
            }
            if (dyn_base < 0) dyn_base = 0;
        }
        static const int env_aligned = getenv("Q4_GEMV_ALIGNED") ? atoi(getenv("Q4_GEMV_ALIGNED")) : 0;
        const size_t rest = 2 * (size_t)K + sizeof(float) * (2 * groups * kw * kRowsInFlight + 512 + 32);
        const bool compact = dyn_base == kDynBase && !env_aligned && kLutBytes + rest <= 200 * 1024;
        if (!compact && rest > 63 * 1024) return Q4_ERR_SHAPE;
        const size_t smem = compact ? kLutBytes + rest : kSmemAligned;
        const bool multi = nmat > 1 && nested;  // without nested statistics a group is just a taller matrix
        auto kern = compact ? (nested ? (multi ? gemv_lut256_kernel<T, true, true, true> : gemv_lut256_kernel<T, true, false, true>)
                                      : gemv_lut256_kernel<T, false, false, true>)
                            : (nested ? (multi ? gemv_lut256_kernel<T, true, true, false> : gemv_lut256_kernel<T, true, false, false>)
                                      : gemv_lut256_kernel<T, false, false, false>);
        static bool attr_set[2][2][2] = {};
        if (!attr_set[compact][nested][multi]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, compact ? 200 * 1024 : kSmemAligned);
            if (e != cudaSuccess) return (int)e;
            attr_set[compact][nested][multi] = true;
        }
        GemvArgs a = {};
        a.x = x;
        a.code = code;
        a.Bq = B;
        a.s = v;
        for (int m = 0; m < kMaxMats; m++) {
            a.offsets[m] = nullptr;
            a.row_end[m] = 0x7fffffff;
        }
        a.offsets[0] = v.offset;
        if (multi) {
            for (int m = 0; m < nmat; m++) {
                a.offsets[m] = offsets[m];
                a.row_end[m] = row_end[m];
            }
            a.row_end[nmat - 1] = 0x7fffffff;
        }
        a.out = out;
        a.bias = bias;
        if (pro) {
            if (sizeof(T) != 2) return Q4_ERR_DTYPE;
            if ((reinterpret_cast<uintptr_t>(pro->x_gate) & 15) || (reinterpret_cast<uintptr_t>(pro->rms_weight) & 15)) return Q4_ERR_ALIGN;
            a.x_gate = pro->x_gate;
            a.rms_weight = pro->rms_weight;
            a.rms_eps = pro->rms_eps;
        }
        a.rows = (int)N;
        a.K = (int)K;
        a.kw = kw;
        a.groups = groups;
        a.lut_iters = (kLutBytes / 16 + threads - 1) / threads;
        a.x_iters = (int)((K / 8 + threads - 1) / threads);
        a.trace = g_gemv_trace;
        static const int env_debug = getenv("Q4_GEMV_DEBUG") ? atoi(getenv("Q4_GEMV_DEBUG")) : 0;
        a.debug_mode = env_debug;
        a.next = (reinterpret_cast<uintptr_t>(next) & 15) == 0 ? (const uint8_t*)next : nullptr;
        a.next_bytes = a.next ? next_bytes : 0;
        const int64_t want = (N + groups - 1) / groups;  // CTAs needed to give every group one row
        int grid = (int)(want < sms ? want : sms);
        a.rows_per_cta = (int)((N + grid - 1) / grid);
        grid = (int)((N + a.rows_per_cta - 1) / a.rows_per_cta);
        return launch_pdl(kern, dim3(grid), dim3(threads), smem, stream, pdl, a);
    }
    if (nmat > 1 || (pro && (pro->x_gate || pro->rms_weight))) return Q4_ERR_SHAPE;  // only the fast path groups / fuses
    // generic: x as fp32 in shared memory
    const size_t smem = 128 + sizeof(float) * (size_t)K;
    if (smem > 200 * 1024) return Q4_ERR_SHAPE;
    auto kern = nested ? gemv_generic_kernel<T, true> : gemv_generic_kernel<T, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    const int64_t want = (N + 7) / 8;
    const int64_t cap = (int64_t)sms * (smem > 100 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4));
    const int grid = (int)(want < cap ? want : cap);
    return launch_pdl(kern, dim3(grid), dim3(256), smem, stream, pdl, x, B, v, code, bias, out, N, K, ilog2(blocksize));
}

int gemv_4bit_fused(const q4_gemv_fused_t* f, cudaStream_t stream)
{
    if (!f) return Q4_ERR_NULL;
    if (!valid_blocksize(f->blocksize)) return Q4_ERR_BLOCKSIZE;
    const int nmat = f->nmat < 1 ? 1 : f->nmat;
    if (f->rows < 0 || f->K < 0 || (f->K & 1) || nmat > kMaxMats) return Q4_ERR_SHAPE;
    if (f->rows == 0) return 0;
    if (!f->x || !f->B || !f->code || !f->out) return Q4_ERR_NULL;
    if (int e = check_stats(f->stats)) return e;
    int one_end[1] = {(int)f->rows};
    const int* row_end = f->row_end ? f->row_end : one_end;
    const float* one_off[1] = {f->stats->offset};
    const float* const* offsets = f->offsets ? f->offsets : one_off;
    if (nmat > 1 && (!f->row_end || (f->stats->qabsmax && !f->offsets))) return Q4_ERR_NULL;
    for (int m = 0; m < nmat; m++)
        if (row_end[m] <= (m ? row_end[m - 1] : 0) || row_end[m] > f->rows) return Q4_ERR_SHAPE;
    if (row_end[nmat - 1] != f->rows) return Q4_ERR_SHAPE;
    GemvPrologue pro;
    pro.x_gate = f->x_gate;
    pro.rms_weight = f->rms_weight;
    pro.rms_eps = f->rms_eps;
    const int flags = f->flags & ~Q4_GEMV_EXACT_F32;
    switch (f->dtype) {
        case Q4_F16:
            return gemv_dispatch<__half>((const __half*)f->x, f->B, f->stats, f->code, (const __half*)f->bias, (__half*)f->out,
                                         f->rows, f->K, f->blocksize, flags, f->prefetch, f->prefetch_bytes, stream, nmat, offsets,
                                         row_end, &pro);
        case Q4_BF16:
            return gemv_dispatch<__nv_bfloat16>((const __nv_bfloat16*)f->x, f->B, f->stats, f->code, (const __nv_bfloat16*)f->bias,
                                                (__nv_bfloat16*)f->out, f->rows, f->K, f->blocksize, flags, f->prefetch,
                                                f->prefetch_bytes, stream, nmat, offsets, row_end, &pro);
        default: return Q4_ERR_DTYPE;
    }
}

int gemv_4bit_grouped(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* const* offsets, const int* row_end,
                      int nmat, const float* code, const void* bias, void* out, int64_t rows, int64_t K, int blocksize, int dtype,
                      int flags, const void* next, int64_t next_bytes, cudaStream_t stream)
{
    if (!valid_blocksize(blocksize)) return Q4_ERR_BLOCKSIZE;
    if (rows < 0 || K < 0 || (K & 1) || nmat < 1 || nmat > kMaxMats) return Q4_ERR_SHAPE;
    if (rows == 0) return 0;
    if (!x || !B || !code || !out || !row_end) return Q4_ERR_NULL;
    if (int e = check_stats(stats)) return e;
    if (stats->qabsmax) {
        if (!offsets) return Q4_ERR_NULL;
        for (int m = 0; m < nmat; m++)
            if (!offsets[m]) return Q4_ERR_NULL;
    }
    for (int m = 0; m < nmat; m++)
        if (row_end[m] <= (m ? row_end[m - 1] : 0) || row_end[m] > rows) return Q4_ERR_SHAPE;
    if (row_end[nmat - 1] != rows) return Q4_ERR_SHAPE;
    flags &= ~Q4_GEMV_EXACT_F32;
    switch (dtype) {
        case Q4_F16:
            return gemv_dispatch<__half>((const __half*)x, B, stats, code, (const __half*)bias, (__half*)out, rows, K, blocksize,
                                         flags, next, next_bytes, stream, nmat, offsets, row_end);
        case Q4_BF16:
            return gemv_dispatch<__nv_bfloat16>((const __nv_bfloat16*)x, B, stats, code, (const __nv_bfloat16*)bias,
                                                (__nv_bfloat16*)out, rows, K, blocksize, flags, next, next_bytes, stream, nmat,
                                                offsets, row_end);
        default: return Q4_ERR_DTYPE;
    }
}

int gemv_4bit(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias, void* out,
              int64_t N, int64_t K, int blocksize, int dtype, int flags, const void* next, int64_t next_bytes,
              cudaStream_t stream)
{
    if (!valid_blocksize(blocksize)) return Q4_ERR_BLOCKSIZE;
    if (N < 0 || K < 0 || (K & 1)) return Q4_ERR_SHAPE;
    if (N == 0) return 0;
    if (!x || !B || !code || !out) return Q4_ERR_NULL;
    if (int e = check_stats(stats)) return e;
    switch (dtype) {
        case Q4_F32:
            // fp32 activations ask for fp32 arithmetic (the reference's only wired instance): exact path
            return gemv_dispatch<float>((const float*)x, B, stats, code, (const float*)bias, (float*)out, N, K, blocksize,
                                        flags | Q4_GEMV_EXACT_F32, next, next_bytes, stream);
        case Q4_F16:
            return gemv_dispatch<__half>((const __half*)x, B, stats, code, (const __half*)bias, (__half*)out, N, K, blocksize,
                                         flags, next, next_bytes, stream);
        case Q4_BF16:
            return gemv_dispatch<__nv_bfloat16>((const __nv_bfloat16*)x, B, stats, code, (const __nv_bfloat16*)bias,
                                                (__nv_bfloat16*)out, N, K, blocksize, flags, next, next_bytes, stream);
        default: return Q4_ERR_DTYPE;
    }
}

}  // namespace q4
