// q4_gemm.cu -- prefill / batched path: out[M, N] = X[M, K] . dequant(W[N, K])^T (+ bias) with the dequantisation fused
// into a tcgen05 (5th-gen tensor core) GEMM.
//
// Replaces: reference modules.py:63-64 -- dequantize_4bit (a full [N, K] fp16 weight written to HBM, core.py:619-631, plus
// two launches for the nested absmax), a full-weight dtype cast, then cuBLAS through F.linear.  Here the packed weight is
// the only thing read from HBM: 0.52 bytes per weight instead of 0.5 + 2 (write) + 2 (read) [+ 2 + 4 for the cast].
// Contract: tolerance-based (fp16/bf16 inputs, fp32 accumulation in TMEM); tests compare with an fp64 product of the
// oracle's dequantised weight.  Roofline: tensor pipe once M >= ~128 tokens, HBM (packed bytes) below that.
//
// Shape of the computation.  The 128-row MMA dimension is given to the WEIGHT rows and the flexible N dimension (16..256)
// to the tokens, so a 16-token batch wastes no tensor work:   D[128 weight rows, BN tokens] += A[128, 64] . B[BN, 64]^T
// per 64-wide k-block, A = dequantised weight tile, B = activation tile, both K-major fp16/bf16, SWIZZLE_128B.
//
// Warp roles (256 threads, one output tile per CTA):
//   warp 0      TMA producer, activations: per stage a 2-D TMA of the activation tile (swizzle 128B), completion on full_tma[stage].
//   warp 3      TMA producer, weights: the PACKED weight tile (128 rows x 32 bytes, no swizzle) of every k-block into its own, deeper
//               ring (kPackedStages x 4 KB) on full_p / empty_p -- the dequantise warps never wait for HBM.
//   warp 1      MMA issuer: one elected lane issues 4 x tcgen05.mma (K = 16 each) per stage into TMEM, then tcgen05.commit
//               releases the stage (empty[stage]); after the last k-block a commit signals the epilogue (tmem_full).
//   warp 2      TMEM allocation / deallocation.
//   warps 4-7   dequantise: thread t owns weight row t of the tile -- exactly one quantisation block (64 weights, 32 packed
//               bytes, one absmax) per k-block: 32 table lookups (the same per-lane-replicated byte table as the GEMV),
//               one half2 multiply by the block's absmax each, eight 16-byte stores into the swizzled A tile
//               (chunk c of row t lands at chunk c ^ (t & 7): conflict-free), fence.proxy.async, arrive on full_a[stage].
//               The same warps then run the epilogue: tcgen05.ld their 32-lane quarter of TMEM (lane = weight row,
//               column = token), add bias, convert, store out[token, row] (lanes = consecutive rows: coalesced).
#include <cuda.h>

#include <cstdlib>

#include "q4_common.cuh"
#include "q4_launch.h"
#include "q4_tma.h"

namespace q4 {

namespace gemm {

constexpr int kTileRows = 128;   // weight rows per CTA (UMMA M)
constexpr int kBK = 64;          // k-block = one quantisation block = 128 bytes of fp16 = one swizzle atom row
constexpr int kDequantWarp0 = 4; // warps 4..7

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, SWIZZLE_128B shared-memory operand descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14),
// leading byte offset (unused for swizzled K-major) 1 in [16,30), stride byte offset = 1024 B (one 8-row swizzle atom) >> 4
// in [32,46), descriptor version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

constexpr int kSets = 3;         // dequantise warp sets (4 warps each); set j handles k-blocks j, j+kSets, ...
constexpr int kPackedStages = 8; // ring of PACKED weight tiles (4 KB each), deeper than the A / activation stages: see the kernel
constexpr int kThreads = 128 + 128 * kSets;

struct Args {
    AbsmaxView s;
    const float* code;       // 16-entry code table
    const void* bias;        // [N] or nullptr
    void* out;               // [M, N]
    int M, N, K;
    int bn;                  // tokens per tile (16..256, multiple of 16)
    int stages;
    int splits;              // split-K factor = cluster size along z (1, 2, 4 or 8)
};

__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// shared memory plan (dynamic, 1024-byte aligned): [A tiles: stages x 16 KB][B tiles: stages x bn*128][packed: kPackedStages x 4 KB]
// [byte-pair table 256 x 4 B x 32 lanes = 32 KB][code2 table 1 KB][256 staged words][barriers][tmem base].
// With split-K the stage memory doubles as the leader's reduction buffer after the main loop: (splits-1) x [bn][128] fp32.
template <typename T, bool NESTED>
__global__ void __launch_bounds__(kThreads, 1)
gemm_dequant_tcgen05_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const Args a)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int stages = a.stages, bn = a.bn, splits = a.splits;
    uint8_t* s_a = smem;                                   // stages x [128 rows x 128 B], swizzled
    uint8_t* s_b = s_a + stages * (kTileRows * 128);       // stages x [bn rows x 128 B], swizzled by TMA
    uint8_t* s_p = s_b + stages * (bn * 128);              // stages x [128 rows x 32 B] packed
    uint32_t* s_lut = reinterpret_cast<uint32_t*>(s_p + kPackedStages * (kTileRows * 32));  // [256][32] pair words
    float* s_code2 = reinterpret_cast<float*>(s_lut + 256 * 32);                     // [256]
    uint32_t* s_words = reinterpret_cast<uint32_t*>(s_code2 + 256);                  // [256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_words + 256);
    // barrier order: full_tma[stages], full_a[stages], empty[stages], tmem_full, full_p[kPackedStages], empty_p[kPackedStages]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 3 * stages + 1 + 2 * kPackedStages);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * kTileRows;   // first weight row of this tile
    const int tok0 = blockIdx.y * bn;          // first token
    const int split = blockIdx.z;              // == rank in the (1,1,splits) cluster
    const int nkb_all = a.K / kBK;
    const int kb_lo = (int)((int64_t)nkb_all * split / splits), kb_hi = (int)((int64_t)nkb_all * (split + 1) / splits);
    const int nkb = kb_hi - kb_lo;             // k-blocks of this CTA (>= 1: dispatcher keeps splits <= nkb_all)
    auto full_tma = [&](int s) { return smem_u32(bars + s); };
    auto full_a = [&](int s) { return smem_u32(bars + stages + s); };
    auto empty = [&](int s) { return smem_u32(bars + 2 * stages + s); };
    const uint32_t tmem_full = smem_u32(bars + 3 * stages);
    // progress counters next to the TMEM base: [0] packed tiles issued by the weight producer, [1] k-blocks whose MMAs have completed.
    // The dequantise sets poll THESE instead of waiting on the ring barriers by parity: a set takes every kSets-th k-block, so it skips
    // phases of a barrier, and a parity wait two phases early passes at once (seen on B200 as tiles dequantised from the packed
    // bytes of k-block i + 8).  The counters are written by threads that see every phase in order.
    volatile int* s_flags = reinterpret_cast<volatile int*>(s_tmem + 2);
    auto full_p = [&](int s) { return smem_u32(bars + 3 * stages + 1 + s); };
    auto empty_p = [&](int s) { return smem_u32(bars + 3 * stages + 1 + kPackedStages + s); };

    // ---- one-time setup
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; s++) {
            mbar_init(full_tma(s), 1);
            mbar_init(full_a(s), 128);
            mbar_init(empty(s), 1);
        }
        mbar_init(tmem_full, 1);
        for (int s = 0; s < kPackedStages; s++) {
            mbar_init(full_p(s), 1);
            mbar_init(empty_p(s), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_flags[0] = 0;
        s_flags[1] = 0;
    }
    if (warp == 2) {
        uint32_t ncols = 32;  // power of two >= max(32, bn)
        while ((int)ncols < bn) ncols *= 2;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // tables: pair word b = half2{code[b>>4], code[b&15]} staged once, then replicated per lane (conflict-free lookups)
    if (tid < 256) {
        s_words[tid] = pack2<__half>(__ldg(a.code + (tid >> 4)), __ldg(a.code + (tid & 15)));  // |code| <= 1: fp16 for both types
        if (NESTED) s_code2[tid] = __ldg(a.s.code2 + tid);
    }
    const float offset = NESTED ? __ldg(a.s.offset) : 0.0f;
    __syncthreads();
    for (int c = tid; c < 256 * 32 / 4; c += kThreads) {
        const uint32_t wv = s_words[c >> 3];
        reinterpret_cast<uint4*>(s_lut)[c] = make_uint4(wv, wv, wv, wv);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;

    // Two independent producers.  A stage's life used to be TMA flight (~1000 clk) -> dequantise (~900) -> MMA (~540) -> commit, with
    // activation tile, packed tile and A tile tied to the same three stages: ~2450 clk / 3 = what a k-block took (1185 clk measured,
    // tensor pipe 47 % busy).  The packed bytes are 4 KB per k-block, so THEIR ring can be eight deep: the dequantise warps then find
    // their input long since landed, start on a k-block the moment its A stage is free (the lookups even before that), and run while
    // the activation tile of the same k-block is still in flight.
    if (warp == 0) {
        // ===== TMA producer, activation tiles
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(bn * 128);
            for (int i = 0; i < nkb; i++) {
                const int s = i % stages, kb = kb_lo + i;
                if (i >= stages) mbar_wait(empty(s), ((i / stages) - 1) & 1);
                mbar_expect_tx(full_tma(s), bytes);
                tma_load_2d(smem_u32(s_b + s * (bn * 128)), &map_x, kb * kBK, tok0, full_tma(s));
            }
        }
    } else if (warp == 3) {
        // ===== TMA producer, packed weight tiles
        if (lane == 0) {
            for (int i = 0; i < nkb; i++) {
                const int s = i % kPackedStages, kb = kb_lo + i;
                if (i >= kPackedStages) mbar_wait(empty_p(s), ((i / kPackedStages) - 1) & 1);
                mbar_expect_tx(full_p(s), kTileRows * 32);
                tma_load_2d(smem_u32(s_p + s * (kTileRows * 32)), &map_w, kb * 32, row0, full_p(s));
                s_flags[0] = i + 1;  // fill i is under way: full_p(s) is in the phase that this fill completes
            }
        }
    } else if (warp == 2) {
        // ===== MMA progress: sees every phase of the empty barriers in order and publishes the count
        if (lane == 0) {
            for (int i = 0; i < nkb; i++) {
                mbar_wait(empty(i % stages), (i / stages) & 1);
                s_flags[1] = i + 1;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer
        // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B fp16 or bf16, both K-major, N = bn, M = 128
        const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kTileRows >> 4) << 24);
        for (int i = 0; i < nkb; i++) {
            const int s = i % stages;
            const uint32_t ph = (i / stages) & 1;
            mbar_wait(full_tma(s), ph);
            mbar_wait(full_a(s), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint64_t adesc = make_desc_sw128(smem_u32(s_a + s * (kTileRows * 128)));
                const uint64_t bdesc = make_desc_sw128(smem_u32(s_b + s * (bn * 128)));
#pragma unroll
                for (int k = 0; k < kBK / 16; k++)  // 16 elements = 32 bytes = +2 in the (>>4) address field
                    umma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (i | k) != 0);
                umma_commit(empty(s));                      // stage reusable once these MMAs have read it
                if (i == nkb - 1) umma_commit(tmem_full);   // accumulator complete
            }
            __syncwarp();
        }
    } else if (warp >= kDequantWarp0) {
        // ===== dequantise: set j = (warp-4)/4 handles k-blocks j, j+kSets, ...; thread t of a set = weight row row0 + t
        const int set = (warp - kDequantWarp0) >> 2;
        const int t = (tid - kDequantWarp0 * 32) & 127;
        const int row = row0 + t;
        const bool row_ok = row < a.N;
        const int bpr = a.K >> 6;
        const uint32_t lut_lane = smem_u32(s_lut) + lane * 4;
        for (int i = set; i < nkb; i += kSets) {
            const int s = i % stages, kb = kb_lo + i;
            const uint32_t ph = (i / stages) & 1;
            // this block's absmax (independent of the pipeline)
            float am = 0.0f;
            if (row_ok) {
                const int64_t blk = (int64_t)row * bpr + kb;
                if (NESTED) am = __fadd_rn(__fmul_rn(s_code2[__ldg(a.s.qabsmax + blk)], __ldg(a.s.absmax2 + (blk >> a.s.shift2))), offset);
                else am = __ldg(a.s.absmax + blk);
            }
            const uint32_t am2 = pack2<__half>(am, am);
            const int sp = i % kPackedStages;
            for (long long spin = 0; s_flags[0] <= i; spin++) {  // the fill has been issued: the barrier is in its phase (as a rule long ago)
                __nanosleep(20);
                if (spin > (1ll << 26)) __trap();  // fail loudly instead of hanging the GPU
            }
            mbar_wait(full_p(sp), (i / kPackedStages) & 1);  // ... and the packed bytes have landed
            const uint4* pk = reinterpret_cast<const uint4*>(s_p + sp * (kTileRows * 32) + t * 32);
            const uint4 p0 = pk[0], p1 = pk[1];
            const uint32_t wd[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
            uint8_t* arow = s_a + s * (kTileRows * 128) + t * 128;
            uint32_t h[8][4];  // the row's 64 dequantised weights: computed before the A stage is known to be free
#pragma unroll
            for (int c = 0; c < 8; c++) {  // 16-byte chunk c = 8 weights = one packed word
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint32_t byte = (wd[c] >> (8 * b)) & 0xFFu;
                    uint32_t v;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(lut_lane + byte * 128));
                    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                        // bf16 has 8 mantissa bits: multiply in fp32 and round once, like the reference's dequantize
                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v));
                        h[c][b] = pack2<T>(f.x * am, f.y * am);
                    } else {
                        asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(h[c][b]) : "r"(v), "r"(am2));
                    }
                }
            }
            mbar_arrive(empty_p(sp));  // the thread's 32 packed bytes have been consumed (the lookups above depend on them)
            if (i >= stages) {  // A tile of this stage no longer read by the MMA of k-block i - stages
                for (long long spin = 0; s_flags[1] < i - stages + 1; spin++) {
                    __nanosleep(20);
                    if (spin > (1ll << 26)) __trap();
                }
                __threadfence_block();
            }
#pragma unroll
            for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(arow + ((c ^ (t & 7)) * 16)) = make_uint4(h[c][0], h[c][1], h[c][2], h[c][3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
            mbar_arrive(full_a(s));
        }
        mbar_wait(tmem_full, 0);  // every MMA of this CTA has completed: TMEM final, stage memory free
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    // ===== split-K: the CTAs of a cluster hold partial sums of the same output tile.  After every CTA's main loop is over
    // (first cluster barrier) the peers push their fp32 partials into the leader's freed stage memory over DSMEM, laid out
    // [peer][token][row] so that a warp's 32 rows are one 128-byte store; second barrier; the leader adds them.
    float* s_red = reinterpret_cast<float*>(smem);
    if (splits > 1) cluster_sync_all();

    if (warp >= kDequantWarp0) {
        // ===== epilogue: TMEM lane = weight row, column = token.  Warp w may only touch TMEM lanes [32 (w&3), +32); the
        // kSets warps sharing a quarter split the token columns between them.
        const int set = (warp - kDequantWarp0) >> 2;
        const int q = warp & 3;
        const int erow_t = q * 32 + lane;  // row inside the tile
        const int erow = row0 + erow_t;
        T* out = reinterpret_cast<T*>(a.out);
        const T* bias = reinterpret_cast<const T*>(a.bias);
        const float bv = (bias && erow < a.N) ? Elem<T>::to_f32(bias[erow]) : 0.0f;
        uint32_t red_remote = 0;
        if (splits > 1 && split > 0) {
            const uint32_t local = smem_u32(s_red + (size_t)(split - 1) * bn * kTileRows);
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(red_remote) : "r"(local), "r"(0));
        }
        for (int c0 = set * 16; c0 < bn; c0 += 16 * kSets) {
            uint32_t r[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (splits > 1 && split > 0) {
#pragma unroll
                for (int j = 0; j < 16; j++)
                    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(red_remote + (uint32_t)(((c0 + j) * kTileRows + erow_t) * 4)), "r"(r[j]) : "memory");
            } else if (splits == 1) {
                if (erow < a.N) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const int tok = tok0 + c0 + j;
                        if (tok < a.M) {
                            T y = Elem<T>::from_f32(__uint_as_float(r[j]));
                            if (bias) y = Elem<T>::from_f32(Elem<T>::to_f32(y) + bv);
                            out[(int64_t)tok * a.N + erow] = y;
                        }
                    }
                }
            }
        }
        if (splits > 1) {
            cluster_sync_all();  // peers' partials have landed in the leader's shared memory
            if (split == 0) {
                for (int c0 = set * 16; c0 < bn; c0 += 16 * kSets) {
                    uint32_t r[16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (erow < a.N) {
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const int tok = tok0 + c0 + j;
                            float acc = __uint_as_float(r[j]);
                            for (int p = 0; p < splits - 1; p++) acc += s_red[((size_t)p * bn + (c0 + j)) * kTileRows + erow_t];
                            if (tok < a.M) {
                                T y = Elem<T>::from_f32(acc);
                                if (bias) y = Elem<T>::from_f32(Elem<T>::to_f32(y) + bv);
                                out[(int64_t)tok * a.N + erow] = y;
                            }
                        }
                    }
                }
            }
        }
    } else if (splits > 1) {
        cluster_sync_all();  // warps 0-3 take part in the second cluster barrier too
    }

    // ---- teardown
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        uint32_t ncols = 32;
        while ((int)ncols < bn) ncols *= 2;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- host side

constexpr size_t kFixedSmem = 256 * 32 * 4 + 1024 + 1024 + (3 * 8 + 1 + 2 * kPackedStages) * 8 + 64 + (size_t)kPackedStages * kTileRows * 32;

// How a launch is cut: tokens per tile (bn), pipeline stages, split-K factor.
//  * One CTA fits per SM, so the launch runs in ceil(tiles x splits / SMs) waves, and a k-block costs a CTA about the same time
//    (~0.55 us measured) whatever bn is: the time of a launch is ~ waves x (k-blocks per CTA).  Hence: the widest token tile, and
//  * split-K over a thread-block cluster while tiles x splits stays within ONE wave (112 tiles x 2 would run as two waves and
//    measured twice the time of the unsplit launch), each CTA keeps >= 4 k-blocks and the reduction buffer fits the stage memory.
//  * Up to 128 tokens: the smallest power of two that holds them.  Beyond: any multiple of 16 up to 256 is a legal UMMA N; the
//    cost model picks among them.
struct Cut {
    int bn, stages, splits;
    int64_t cost;
};
static Cut plan_cut(int bn, int64_t M, int64_t N, int64_t K, int sms)
{
    Cut c;
    c.bn = bn;
    const size_t per_stage = kTileRows * 128 + (size_t)bn * 128;  // A tile + activation tile
    c.stages = (int)((225 * 1024 - kFixedSmem) / per_stage);
    if (c.stages > 8) c.stages = 8;
    const int64_t tiles = ((N + kTileRows - 1) / kTileRows) * ((M + bn - 1) / bn);
    c.splits = 1;
    while (c.splits < 8 && tiles * c.splits * 2 <= sms && (K / kBK) / (c.splits * 2) >= 4 &&
           (size_t)(c.splits * 2 - 1) * bn * kTileRows * 4 <= c.stages * per_stage)
        c.splits *= 2;
    const int64_t waves = (tiles * c.splits + sms - 1) / sms;
    // + prologue / epilogue of a tile, in k-blocks; ties go to the narrower tile (more stages fit, shorter ramp: 4096x4096 at 256
    // tokens measured 30 us with two 128-token tiles against 38 us with one 256-token tile, both split in two)
    c.cost = waves * ((K / kBK + c.splits - 1) / c.splits + 6) * 1024 + bn;
    return c;
}
static Cut pick_cut(int64_t M, int64_t N, int64_t K, int sms)
{
    if (M <= 128) return plan_cut(M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : 128, M, N, K, sms);
    Cut best = plan_cut(256, M, N, K, sms);
    for (int bn = 240; bn >= 128; bn -= 16) {
        const Cut c = plan_cut(bn, M, N, K, sms);
        if (c.cost < best.cost) best = c;
    }
    return best;
}

template <typename T>
static int launch(const T* X, const uint8_t* B, const q4_absmax_t* st, const float* code, const T* bias, T* out, int64_t M,
                  int64_t N, int64_t K, cudaStream_t stream)
{
    const AbsmaxView v = make_view(st);
    const bool nested = st->qabsmax != nullptr;
    static const int env_bn = getenv("Q4_GEMM_BN") ? atoi(getenv("Q4_GEMM_BN")) : 0;
    static const int env_splits = getenv("Q4_GEMM_SPLITS") ? atoi(getenv("Q4_GEMM_SPLITS")) : 0;
    const Cut cut = env_bn ? plan_cut(env_bn, M, N, K, sm_count()) : pick_cut(M, N, K, sm_count());
    const int bn = cut.bn, stages = cut.stages;
    const int splits = env_splits ? env_splits : cut.splits;
    if (stages < 2) return Q4_ERR_SHAPE;
    const size_t per_stage = kTileRows * 128 + (size_t)bn * 128;
    const size_t smem = stages * per_stage + kFixedSmem + 1024;

    CUtensorMap map_x, map_w;
    const CUtensorMapDataType dt = std::is_same<T, __nv_bfloat16>::value ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    if (!make_map_2d(&map_x, dt, X, (uint64_t)K, (uint64_t)M, (uint64_t)K * 2, kBK, (uint32_t)bn, CU_TENSOR_MAP_SWIZZLE_128B))
        return Q4_ERR_DEVICE;
    if (!make_map_2d(&map_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, B, (uint64_t)K / 2, (uint64_t)N, (uint64_t)K / 2, 32, kTileRows,
                     CU_TENSOR_MAP_SWIZZLE_NONE))
        return Q4_ERR_DEVICE;

    Args a;
    a.s = v;
    a.code = code;
    a.bias = bias;
    a.out = out;
    a.M = (int)M;
    a.N = (int)N;
    a.K = (int)K;
    a.bn = bn;
    a.stages = stages;
    a.splits = splits;
    auto kern = nested ? gemm_dequant_tcgen05_kernel<T, true> : gemm_dequant_tcgen05_kernel<T, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((N + kTileRows - 1) / kTileRows), (unsigned)((M + bn - 1) / bn), (unsigned)splits);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = (unsigned)splits;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, map_x, map_w, a);
    if (e != cudaSuccess) return (int)e;
    return finish_launch();
}

}  // namespace gemm

int gemm_4bit(const void* X, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias, void* out,
              int64_t M, int64_t N, int64_t K, int blocksize, int dtype, cudaStream_t stream)
{
    if (blocksize != 64) return Q4_ERR_BLOCKSIZE;  // the fused path needs one block per 64-wide k-block
    if (M < 0 || N < 0 || K <= 0 || (K % 64) != 0) return Q4_ERR_SHAPE;
    if (M == 0 || N == 0) return 0;
    if (!X || !B || !code || !out) return Q4_ERR_NULL;
    if (int e = check_stats(stats)) return e;
    if ((reinterpret_cast<uintptr_t>(X) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || ((K * 2) % 16) != 0) return Q4_ERR_ALIGN;
    if (N * (K / 64) >= (1ll << 31) || M >= (1ll << 31)) return Q4_ERR_SHAPE;
    switch (dtype) {
        case Q4_F16:
            return gemm::launch<__half>((const __half*)X, B, stats, code, (const __half*)bias, (__half*)out, M, N, K, stream);
        case Q4_BF16:
            return gemm::launch<__nv_bfloat16>((const __nv_bfloat16*)X, B, stats, code, (const __nv_bfloat16*)bias,
                                               (__nv_bfloat16*)out, M, N, K, stream);
        default: return Q4_ERR_DTYPE;
    }
}

}  // namespace q4
