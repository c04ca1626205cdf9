// q4_gemv_tc.cuh -- decode GEMV on the 5th-generation tensor cores (tcgen05, A operand from tensor memory).
//
//   out[r] = sum_b absmax[r,b] * sum_{k in block b} x[k] * code[nib(r,k)]            (+ bias[r])
//
// Same contract as q4_gemv_mma.cuh (replaces reference csrc/kernels.cu:1061-1219 + core.py:467-468); it exists because the
// legacy mma.sync path of sm_100 retires one m16n8k16 per 32 clocks per SM quarter = 16 packed bytes/clk/SM, i.e. it caps the
// stream at ~4.6 TB/s -- below HBM.  tcgen05.mma has no such limit and takes its A operand straight from tensor memory:
//
//   * a thread owns ONE weight row of a 128-row tile (TMEM lane = row).  Per quantisation block it loads the row's 32 packed
//     bytes (one 256-bit load = one DRAM sector), decodes them with 32 byte-table lookups (same conflict-free table as the
//     mma.sync kernel: word = pair {code[b>>4], code[b&15]} in the activation type) and writes the 32 pairs -- already in
//     A-operand order, two consecutive k per 32-bit column -- into tensor memory with two tcgen05.st (registers -> TMEM; no
//     shared-memory store, no conversion).
//   * one elected thread issues 4 x tcgen05.mma (M=128 rows, K=16, N=16) per block: D[128, 16] (+)= A[128, 16] . B[16, 16]^T
//     with B read from the activation vector in shared memory AS IT IS: the no-swizzle K-major descriptor is given a
//     16-byte leading offset and core-matrix rows that are 16 bytes apart, so row 0 of B is x[k0 .. k0+16) and rows 1-15
//     are shifted windows of x whose outputs (columns 1-15 of D) are simply never read.
//   * column 0 of D is the (row, block) partial sum: the owning thread reads it back (tcgen05.ld, one register), scales it by
//     the block's decoded absmax and accumulates its row in a register -- no cross-lane reduction anywhere.
//   * work = (row tile, block) units dealt in contiguous runs to the 2 x 148 (CTA, warpgroup) pairs, perfectly balanced; a row
//     tile shared by several runs is combined through a caller-provided workspace in a FIXED order by the last arriver
//     (deterministic; counters reset themselves).
//
// Warp roles (288 threads, <= half an SM so that the next launch's CTA is co-resident, see q4_gemv_mma.cuh):
//   warps 0-3 / 4-7   two producer warpgroups, each with its own run of units, 3 A buffers + 2 D buffers in tensor memory
//   warp 8            tensor-memory allocation, barrier setup, table copy, MMA issue for both warpgroups
#pragma once

#include <type_traits>

#include "q4_common.cuh"
#include "q4_gemv_mma.cuh"

namespace q4 {

constexpr int kTcGroups = 4;                 // producer warpgroups per CTA (each owns all 128 TMEM lanes, 128 columns)
constexpr int kTcThreads = 128 * kTcGroups;  // one CTA per SM
constexpr int kTcRows = 128;   // rows per tile = TMEM lanes = UMMA M
constexpr int kTcU = 3;        // packed blocks a thread keeps in flight in registers
constexpr int kTcUnroll = 6;   // lcm(kTcU, kTcA, kTcD): every buffer index is a constant after unrolling
constexpr int kTcA = 3;        // A buffers per warpgroup (32 TMEM columns each)
constexpr int kTcD = 2;        // D buffers per warpgroup (16 TMEM columns each)
constexpr int kTcCols = 512;   // TMEM columns per CTA: 4 warpgroups x (3 x 32 + 2 x 16)
constexpr int kTcXPad = 512;   // bytes of zeros after x: the shifted B windows of the last block read past the end
constexpr int kTcAhead = 16;   // L2 prefetch distance in units (blocks) ahead of the register loads

struct TcGemvArgs {
    const void* x;
    const void* lut;     // prebuilt table image (required)
    const uint8_t* Bq;   // packed weight [rows, K/2]
    AbsmaxView s;
    const float* offsets[kMaxMats];
    int row_end[kMaxMats];
    void* out;
    const void* bias;
    const void* x_gate;
    const void* rms_weight;
    float rms_eps;
    const uint8_t* next;
    int64_t next_bytes;
    float* ws_part;  // workspace: [row tile][max_seg][128] partial sums
    int* ws_count;   // workspace: [row tile] arrival counters (zero between launches)
    int max_seg;
    int rows, K;
    int tokens;      // rows of x / out: 1 (decode GEMV) .. 16 (small batch: the N columns of the MMA are the tokens)
    int rt_total;    // ceil(rows / 128)
    unsigned long long* trace;
    int debug;       // developer experiments (env Q4_GEMV_DEBUG)
};

__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tc_trace(const TcGemvArgs& a, int slot)
{
    if (a.trace && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.trace[blockIdx.x * 8 + slot] = t;
    }
}

// 16 registers -> 16 consecutive TMEM columns of the warp's 32 lanes
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// MT = 1: one activation vector, read in place through the overlapping descriptor (above).  MT = 16: up to 16 tokens -- x [tokens, K],
// out [tokens, rows]; every warpgroup stages the 64 k of its current block for all tokens into a 2-KB shared-memory slot laid out
// as the canonical K-major B tile ([8-element k chunk][token][16 B]: core-matrix rows = tokens), so column n of D is token n and one
// pass over the packed weight serves the whole batch.
template <typename T, bool NESTED, bool MULTI, int MT>
__global__ void __launch_bounds__(kTcThreads, 1)
gemv_tc_kernel(const TcGemvArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int K = a.K, R = a.rows;
    const int bpr = K >> 6;
    // shared memory: [table 64 KB at the start of dynamic shared memory = window offset kDynBase][x K*2 + pad][red][barriers]
    const uint32_t smem_saddr = (uint32_t)__cvta_generic_to_shared(smem);
    constexpr int kImm = kDynBase;
    if (smem_saddr != (uint32_t)kDynBase) __trap();  // the host probe and the kernel disagree about the window layout
    const uint32_t lut_saddr = 0;
    uint8_t* s_xb = smem + kLutBytes;
    uint4* s_x = reinterpret_cast<uint4*>(s_xb);
    const int xbytes = MT == 1 ? K * 2 + kTcXPad : kTcGroups * kTcA * 2048;  // MT > 1: per-warpgroup ring of B tiles
    float* s_red = reinterpret_cast<float*>(s_xb + xbytes);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_red + 32);  // [0] table, [1 + wg*6 + 3 + s] done (the slots before are unused)
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 1 + 6 * kTcGroups);
    int* s_flag = reinterpret_cast<int*>(s_tmem + 1);            // [kTcGroups]
    int* s_cnt = s_flag + kTcGroups;                             // [kTcGroups][kTcA] arrival counters of the A buffers

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wg = warp >> 2, wq4 = warp & 3;  // warpgroup, warp inside it (= TMEM lane quarter)
    const int t = tid & 127;                   // row inside the tile
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(s_bar);
    const uint32_t bar_done0 = bar0 + 8u * (1 + wg * 6 + 3);  // [1 + wg*6 + 3 + s]: MMA(unit) done, s = A buffer

    // unit runs: unit u = (row tile u / bpr, block u % bpr); run j = [j*U/G, (j+1)*U/G), j = 4*CTA + warpgroup
    const int64_t U = (int64_t)a.rt_total * bpr;
    const int G = kTcGroups * gridDim.x;
    auto run_lo = [&](int j) { return (int)(((int64_t)j * U) / G); };
    auto run_of = [&](int u) { return (int)((((int64_t)u + 1) * G - 1) / U); };  // the run that contains unit u

    pdl_launch_dependents();
    tc_trace(a, 0);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(s_tmem)),
                     "n"(kTcCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        if (lane < kTcGroups * kTcA) s_cnt[lane] = 0;
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
            for (int g = 0; g < kTcGroups; g++)
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * (1 + g * 6 + 3 + i)));  // done: tcgen05.commit
                }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0), "r"(kLutBytes) : "memory");
#pragma unroll
            for (int i = 0; i < 4; i++)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 lut_saddr + kImm + i * (kLutBytes / 4)),
                             "l"(reinterpret_cast<const uint8_t*>(a.lut) + i * (kLutBytes / 4)), "r"(kLutBytes / 4), "r"(bar0)
                             : "memory");
        }
    }
    if (warp == 1 && a.next_bytes > 0) {  // optional hint: this CTA's share of what the next launch streams, HBM -> L2
        const int64_t share = ((a.next_bytes / gridDim.x) + 15) & ~(int64_t)15;
        const int64_t lo = share * blockIdx.x;
        const int64_t n = lo + share <= a.next_bytes ? share : a.next_bytes - lo;
        if (n > 0) bulk_prefetch_l2_range(a.next + lo, n, lane);
    }

    // ---- this thread's run, the first kTcU blocks into registers, the next kTcAhead toward L2
    const int j_run = kTcGroups * blockIdx.x + wg;
    const int u_lo = run_lo(j_run), u_hi = run_lo(j_run + 1);
    const int n_units = u_hi - u_lo;
    u32x8 wq[kTcU];
    uint32_t aq[kTcU];   // nested: 8-bit absmax code; else the fp32 absmax bits
    int rt_ld = u_lo / bpr, b_ld = u_lo - rt_ld * bpr;  // load cursor
    float off[kMaxMats];
#pragma unroll
    for (int m = 0; m < kMaxMats; m++) off[m] = (NESTED && (MULTI || m == 0) && a.offsets[m]) ? __ldg(a.offsets[m]) : 0.0f;
    auto unit_ptr = [&](int rt, int b, int& blk) {
        const int row = rt * kTcRows + t;
        blk = (row < R ? row : R - 1) * bpr + b;  // rows * bpr < 2^31: dispatcher
        return a.Bq + (int64_t)blk * 32;
    };
    auto load_unit = [&](int k) {
        int blk;
        const uint8_t* p = unit_ptr(rt_ld, b_ld, blk);
        wq[k] = ldg_stream_256(p);
        if (NESTED) aq[k] = __ldg(a.s.qabsmax + blk);
        else aq[k] = __float_as_uint(__ldg(a.s.absmax + blk));
        if (++b_ld == bpr) {
            b_ld = 0;
            rt_ld++;
        }
    };
    // L2 prefetch cursor: one 128-byte line = 4 consecutive blocks of this thread's row
    int rt_pf = rt_ld, b_pf = b_ld, left_pf = n_units;
    auto prefetch_line = [&]() {  // prefetch the line holding unit (rt_pf, b_pf) and advance to the next line of the run
        int blk;
        const uint8_t* p = unit_ptr(rt_pf, b_pf, blk);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        const int step = 4 - (b_pf & 3);  // units to the next line boundary (rows are 128-byte multiples: bpr % 4 == 0)
        left_pf -= step;
        b_pf += step;
        if (b_pf >= bpr) {
            b_pf = 0;
            rt_pf++;
        }
    };
#pragma unroll
    for (int k = 0; k < kTcU; k++)
        if (k < n_units) load_unit(k);
    for (int i = 0; i < kTcAhead / 4 + 1 && left_pf > 0; i++) prefetch_line();
    tc_trace(a, 1);

    // ---- everything below may read the previous kernel's output
    pdl_wait();
    tc_trace(a, 2);
    if (MT == 1) {
        const int nchunk = K >> 3, npad = nchunk + kTcXPad / 16;
        if (a.x_gate || a.rms_weight) {
            stage_x_fused<T, false>(a.x, a.x_gate, a.rms_weight, a.rms_eps, K, s_x, s_red, nchunk, npad);
        } else {
            for (int cb = tid; cb < npad; cb += 4 * kTcThreads) {
                uint4 v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = cb + j * kTcThreads;
                    v[j] = make_uint4(0, 0, 0, 0);
                    if (c < nchunk) v[j] = __ldg(reinterpret_cast<const uint4*>(a.x) + c);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = cb + j * kTcThreads;
                    if (c < npad) s_x[c] = v[j];
                }
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // x (generic-proxy stores) -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;
    tc_trace(a, 3);
    tc_mbar_wait(bar0, 0);  // table landed
    tc_trace(a, 6);

    const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
    // instruction descriptor: D fp32, A/B fp16|bf16, K-major, N = 16, M = 128
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
    // B descriptor: K-major, no swizzle, leading (k-half) offset 16 B, stride (8-row group) offset 128 B, version 1
    // (MT > 1: the real tile -- k-chunk stride 256 B, 8-token group stride 128 B)
    const uint64_t bdesc0 = ((uint64_t)((MT == 1 ? 16 : 256) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
    const uint32_t x_saddr = (uint32_t)__cvta_generic_to_shared(s_xb);
    const uint32_t lane_base = lut_saddr | (uint32_t)(lane * 4);
    const uint32_t t_col = tmem_base + (uint32_t)(wg * 128);              // this warpgroup's columns, lane 0
    const uint32_t t_lane = t_col + ((uint32_t)(wq4 * 32) << 16);         // ... at this warp's lane quarter
    int rt_st = u_lo / bpr, b_st = u_lo - rt_st * bpr;  // stage cursor
    int rt_acc = rt_st;                                 // row tile the accumulator belongs to
    int rt_epi = rt_st, b_epi = b_st;                   // epilogue cursor
    float acc[MT];
#pragma unroll
    for (int n = 0; n < MT; n++) acc[n] = 0.0f;
    float am_pend[3];
    // second-level absmax: one value per 256 consecutive blocks of the flattened weight, i.e. it changes at most once per
    // ~256 units of this thread's row -- fetched when the index moves, not per unit
    int a2_idx = -1;
    float a2 = 0.0f;

    auto flush = [&](int rt) {
        const int first = run_of(rt * bpr);
        const int nseg = run_of(rt * bpr + bpr - 1) - first + 1;
        const int row = rt * kTcRows + t;
        bool fin = true;
        float* part = a.ws_part + ((int64_t)rt * a.max_seg) * (MT * kTcRows);  // [segment][token][128 rows]
        if (nseg > 1) {
            const int seg = j_run - first;
#pragma unroll
            for (int n = 0; n < MT; n++) __stcg(part + (seg * MT + n) * kTcRows + t, acc[n]);
            __threadfence();
            asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
            if (t == 0) s_flag[wg] = atomicAdd(a.ws_count + rt, 1);
            asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
            fin = s_flag[wg] == nseg - 1;
            if (fin) {  // last arriver: fixed-order sum of every run's partial
                __threadfence();
#pragma unroll
                for (int n = 0; n < MT; n++) {
                    float total = 0.0f;
                    for (int s = 0; s < nseg; s++) total += __ldcg(part + (s * MT + n) * kTcRows + t);
                    acc[n] = total;
                }
                if (t == 0) a.ws_count[rt] = 0;
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");  // s_flag reusable
        }
        if (fin && row < R) {
            const T* bias = reinterpret_cast<const T*>(a.bias);
            const float bv = bias ? Elem<T>::to_f32(bias[row]) : 0.0f;
#pragma unroll
            for (int n = 0; n < MT; n++) {
                if (n < a.tokens) {
                    T y = Elem<T>::from_f32(acc[n]);
                    if (bias) y = Elem<T>::from_f32(Elem<T>::to_f32(y) + bv);  // torch `out += bias`
                    reinterpret_cast<T*>(a.out)[(size_t)n * R + row] = y;
                }
            }
        }
    };

    // Per unit i (k = i % 6, all buffer indices constants after unrolling):
    //   1. finish the D read started at the end of the previous step: acc += D(i-2) * absmax(i-2)
    //   2. decode unit i: 32 lookups -> two tcgen05.st into A buffer i % 3; wait::st, fence, one arrive per warp on full(i)
    //   3. refill the registers just consumed, prefetch ahead
    //   4. the warp whose turn it is (i % 4) waits for all four arrivals and issues the 4 MMAs + commit -> done(i)
    //   5. wait for MMA(i-1) -- issued a whole step ago -- and start reading its D column (tcgen05.ld is asynchronous)
    // Nothing in a step waits for work issued in the same step.  Buffer reuse is safe by program order: every warp completed
    // the D read of unit i-2 (1.) before its arrival for unit i, and MMAs complete in issue order, so A(i-3) and D(i-2) are
    // free when the stores / MMAs of unit i start.
    uint32_t d_pend[MT];
#pragma unroll
    for (int n = 0; n < MT; n++) d_pend[n] = 0;
    for (int base = 0; base <= n_units + 1; base += kTcUnroll) {  // two extra steps drain the pipeline
#pragma unroll
        for (int k = 0; k < kTcUnroll; k++) {
            const int i = base + k;
            if (i < n_units) {
                // absmax of this unit (kept until its D column has been read, two steps later)
                float am;
                if (NESTED) {
                    float o = off[0];
                    const int row = rt_st * kTcRows + t;
                    if (MULTI) o = row < a.row_end[0] ? off[0] : (row < a.row_end[1] ? off[1] : (row < a.row_end[2] ? off[2] : off[3]));
                    const int idx = ((row < R ? row : R - 1) * bpr + b_st) >> a.s.shift2;
                    if (idx != a2_idx) {
                        a2 = __ldg(a.s.absmax2 + idx);
                        a2_idx = idx;
                    }
                    const float c2 = __uint_as_float(lut_lookup<0, kImm + 128>(aq[k % kTcU], lane_base));
                    am = __fadd_rn(__fmul_rn(c2, a2), o);  // reference: kernels.cu:552 then core.py:468
                } else {
                    am = __uint_as_float(aq[k % kTcU]);
                }
                am_pend[k % 3] = am;
                uint4 xpiece = make_uint4(0, 0, 0, 0);
                if (MT > 1) {  // this thread's 16 bytes of the block's B tile: k chunk t / 16 of token t % 16
                    const int n = t & 15;
                    if (n < a.tokens)
                        xpiece = __ldcg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.x) + (size_t)n * K + b_st * 64 + (t >> 4) * 8));
                }
                // 2. decode: 32 lookups -> 32 TMEM columns of this thread's lane (A buffer k % 3)
                const uint32_t a_t = t_lane + (uint32_t)((k % kTcA) * 32);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t v[16];
#pragma unroll
                    for (int w4 = 0; w4 < 4; w4++) {
                        const uint32_t wv = wq[k % kTcU].v[4 * h + w4];
                        v[4 * w4 + 0] = lut_lookup<0, kImm>(wv, lane_base);
                        v[4 * w4 + 1] = lut_lookup<1, kImm>(wv, lane_base);
                        v[4 * w4 + 2] = lut_lookup<2, kImm>(wv, lane_base);
                        v[4 * w4 + 3] = lut_lookup<3, kImm>(wv, lane_base);
                    }
                    tmem_st16(a_t + (uint32_t)(16 * h), v);
                }
                if (i >= 2 && i <= n_units + 1) {  // 1. unit i-2
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (rt_epi != rt_acc) {  // the run moved on to the next row tile
                        flush(rt_acc);
#pragma unroll
                        for (int n = 0; n < MT; n++) acc[n] = 0.0f;
                        rt_acc = rt_epi;
                    }
#pragma unroll
                    for (int n = 0; n < MT; n++) acc[n] = fmaf(__uint_as_float(d_pend[n]), am_pend[(k + kTcUnroll - 2) % kTcUnroll % 3], acc[n]);
                    if (++b_epi == bpr) {
                        b_epi = 0;
                        rt_epi++;
                    }
                }
                if (MT > 1) {
                    *reinterpret_cast<uint4*>(s_xb + (wg * kTcA + (k % kTcA)) * 2048 + t * 16) = xpiece;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy store -> visible to the tensor core
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                // the LAST of the group's four warps to get here issues the unit's MMAs at once (shared-memory counter instead of
                // a barrier: nobody waits)
                if (lane == 0) {
                    __threadfence_block();
                    const int old = atomicAdd(s_cnt + wg * kTcA + (k % kTcA), 1);
                    if (old == 3) {
                        s_cnt[wg * kTcA + (k % kTcA)] = 0;
                        __threadfence_block();
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t d_t = t_col + (uint32_t)(kTcA * 32 + (k % kTcD) * 16);
                        const uint32_t a_m = t_col + (uint32_t)((k % kTcA) * 32);
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint32_t b_saddr = MT == 1 ? x_saddr + (uint32_t)(b_st * 64 + q * 16) * 2u
                                                             : x_saddr + (uint32_t)((wg * kTcA + (k % kTcA)) * 2048 + q * 512);
                            const uint64_t bdesc = bdesc0 | (uint64_t)((b_saddr & 0x3FFFFu) >> 4);
                            asm volatile(
                                "{\n"
                                ".reg .pred p;\n"
                                "setp.ne.b32 p, %4, 0;\n"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
                                "}\n" ::"r"(d_t),
                                "r"(a_m + (uint32_t)(q * 8)), "l"(bdesc), "r"(idesc), "r"((uint32_t)(q != 0))
                                : "memory");
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_done0 + 8u * (k % kTcA))
                                     : "memory");
                    }
                }
                __syncwarp();
                // 3.
                if (i + kTcU < n_units) load_unit(k % kTcU);
                if ((b_st & 3) == 0 && left_pf > 0) prefetch_line();
                if (++b_st == bpr) {
                    b_st = 0;
                    rt_st++;
                }
            }
            else {  // drain steps: nothing to decode
                if (i >= 2 && i <= n_units + 1) {  // 1. unit i-2
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (rt_epi != rt_acc) {  // the run moved on to the next row tile
                        flush(rt_acc);
#pragma unroll
                        for (int n = 0; n < MT; n++) acc[n] = 0.0f;
                        rt_acc = rt_epi;
                    }
#pragma unroll
                    for (int n = 0; n < MT; n++) acc[n] = fmaf(__uint_as_float(d_pend[n]), am_pend[(k + kTcUnroll - 2) % kTcUnroll % 3], acc[n]);
                    if (++b_epi == bpr) {
                        b_epi = 0;
                        rt_epi++;
                    }
                }
            }
            if (i >= 1 && i <= n_units) {  // 5. unit i-1
                const int kp = (k + kTcUnroll - 1) % kTcUnroll;
                tc_mbar_wait(bar_done0 + 8u * (kp % kTcA), (uint32_t)(((i - 1) / kTcA) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_addr = t_lane + (uint32_t)(kTcA * 32 + (kp % kTcD) * 16);
                if constexpr (MT == 1) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(d_pend[0]) : "r"(d_addr) : "memory");
                } else {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(d_pend[0]), "=r"(d_pend[1]), "=r"(d_pend[2]), "=r"(d_pend[3]), "=r"(d_pend[4]), "=r"(d_pend[5]), "=r"(d_pend[6]),
                          "=r"(d_pend[7]), "=r"(d_pend[8]), "=r"(d_pend[9]), "=r"(d_pend[10]), "=r"(d_pend[11]), "=r"(d_pend[12]),
                          "=r"(d_pend[13]), "=r"(d_pend[14]), "=r"(d_pend[15])
                        : "r"(d_addr)
                        : "memory");
                }
            }
        }
    }
    if (n_units > 0) flush(rt_acc);
    tc_trace(a, 4);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTcCols) : "memory");
    }
    tc_trace(a, 5);
}

}  // namespace q4
