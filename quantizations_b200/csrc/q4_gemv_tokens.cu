// q4_gemv_tokens.cu -- small-batch decode (2..16 tokens: speculative / multi-sequence decode) in ONE pass over the packed weight:
//     out[t, r] = sum_k x[t, k] * code[nib(B[r, k])] * absmax[(r*K + k) / 64]   (+ bias[r])
// The reference has no such path: more than one token sends it to a full dequantise + dense GEMM (modules.py:56-64), i.e. it writes
// and re-reads 4 x the packed bytes; one decode GEMV per token (what this repo did before) streams the packed weight `tokens` times.
//
// Same decode as the batch-1 kernels (one PRMT + one LDS per packed byte through the 64-KB byte table, mma.sync.m16n8k16), but the
// roles of the MMA's dimensions change.  Batch-1 uses the 8 B columns as 8 quantisation blocks (a block-diagonal activation), so the
// per-block absmax multiplies the accumulator once per tile; here the 8 columns are 8 TOKENS, the 16 A rows are 16 weight rows, and
// one quantisation block (64 k) is exactly four MMAs whose accumulator is scaled by the block's absmax and added to the running
// sums: 16 FFMA per 2 KB of packed weight on top of the batch-1 instruction mix.  The cost of a call is that of one pass whatever
// `tokens` (<= 8) is; 9..16 tokens are two passes.
//
// Work split: a CTA (16 warps, one per SM) owns row tiles (16 rows) bx, bx + G, ...; warp w owns the 256-k chunks w, w + 16, ... of
// every one of them, so its 8 x 256 activations stay in 32 registers per lane for the whole chunk (no shared-memory staging of x at
// all: each lane reads the 2 x 16 bytes per block it feeds to the B operand straight from global memory).  The packed bytes of an
// item (16 rows x 128 bytes) arrive in the warp's private 3-deep shared-memory ring by cp.async, whole 128-byte lines per
// instruction; a lane (g, t4) then reads the 8 bytes [8*t4, 8*t4 + 8) of every block of rows g and g + 8 from there (8 LDS.64).
// Partial sums of the 16 warps meet in shared memory in a fixed order: bit-reproducible, no atomics.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "q4_common.cuh"
#include "q4_gemv_mma.cuh"
#include "q4_launch.h"

namespace q4 {

extern int g_dyn_base_probed(cudaStream_t stream);  // q4_gemv.cu

namespace tok {

constexpr int kWarps = 16, kThreads = kWarps * 32;
constexpr int kTileRows = 16;    // weight rows per MMA
constexpr int kPassTiles = 8;    // row tiles whose partial sums are held in shared memory at once
constexpr int kMaxTokens = 8;    // B columns
constexpr int kAhead = 3;        // looked-up A fragments in flight ahead of the tensor pipe
constexpr int kStages = 3;       // depth of a warp's private ring of items (2 KB each)
constexpr int kItemBytes = kTileRows * 128;
constexpr int kRingBytes = kWarps * kStages * kItemBytes;
constexpr int kPartBytes = kPassTiles * kWarps * 128 * 4;
constexpr int kSmemBytes = kLutBytes + kRingBytes + kPartBytes + 16;
static_assert(kSmemBytes <= 232448 - 1024, "shared-memory plan");

struct Args {
    const void* x;       // [tokens, K]
    const uint8_t* B;    // [N, K/2]
    AbsmaxView s;
    const void* lut;     // 64-KB table image (q4_gemv_lut_build)
    const void* bias;    // [N] or null
    void* out;           // [tokens, N]
    int N, K, tokens;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint2 ldg_stream_64(const void* p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// statistics of one (row tile, chunk) item as a lane holds them: rows g and g + 8, the chunk's 4 blocks
template <bool NESTED> struct Stat {
    uint32_t st[NESTED ? 4 : 8];  // nested: {4 codes of row g, 4 codes of row g+8, absmax2 of either}; else the 8 fp32 absmax values
};

template <typename T, bool NESTED>
__global__ void __launch_bounds__(kThreads, 1) gemv_tokens_kernel(const Args c)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    // the PRMT splice of the lookups needs the table at (64-KB aligned window address) + (compile-time immediate): first thing in
    // dynamic shared memory, which starts kDynBase into the CTA's window (probed on the host, trapped here if violated)
    constexpr int kImm = kDynBase;
    const uint32_t smem_saddr = smem_u32(smem);
    if (smem_saddr != kDynBase) __trap();
    float* s_part = reinterpret_cast<float*>(smem + kLutBytes + kRingBytes);  // [kPassTiles][kWarps][4][32]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + kLutBytes + kRingBytes + kPartBytes);
    const uint32_t bar = smem_u32(s_bar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int bx = blockIdx.x, G = gridDim.x;
    const int N = c.N, K = c.K;
    const int nch = K >> 8, bpr = K >> 6, ntiles = N >> 4;
    const int halfK = K >> 1;
    const uint32_t lane_base = (uint32_t)(lane * 4);

    pdl_launch_dependents();
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kLutBytes) : "memory");
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_saddr + i * (kLutBytes / 4)),
                         "l"(reinterpret_cast<const uint8_t*>(c.lut) + i * (kLutBytes / 4)), "r"(kLutBytes / 4), "r"(bar)
                         : "memory");
    }
    __syncthreads();  // the barrier exists before anybody waits on it

    const int my_tiles = bx < ntiles ? (ntiles - bx + G - 1) / G : 0;
    const int nchw = warp < nch ? (nch - warp + kWarps - 1) / kWarps : 0;  // chunks of this warp
    const int nw = nch < kWarps ? nch : kWarps;                            // warps that produce partial sums
    const float offset = (NESTED && c.s.offset) ? __ldg(c.s.offset) : 0.0f;
    const uint8_t* const qabs = c.s.qabsmax;
    const float* const am2 = NESTED ? c.s.absmax2 : c.s.absmax;
    const int shift2 = c.s.shift2;

    // ---- the warp's private ring: kStages x (16 rows x 128 bytes).  An item arrives by 4 cp.async instructions of 16 bytes per lane,
    // each covering four whole 128-byte lines (lane -> row 4i + lane/8, 16-byte piece lane%8): fully coalesced, where reading the
    // lane's own 8 bytes per (row, block) straight from global memory touches 8 lines per instruction (measured: +30 % kernel time).
    // Piece p of row r sits at piece p ^ 2*(r & 3): the copies write whole rows, and the 64-bit reads below (row g, piece
    // 2b + t4/2 for the 16 lanes g = 0..3 of a half warp) meet 8 different pieces -- conflict-free either way.
    const uint32_t ring = smem_saddr + kLutBytes + (uint32_t)warp * (kStages * kItemBytes);
    const int crow = lane >> 3, cpiece = lane & 7;
    auto copy_item = [&](int stage, int tile, int ch) {
        const uint8_t* src = c.B + (size_t)(tile * kTileRows + crow) * halfK + ch * 128 + cpiece * 16;
        const uint32_t dst = ring + (uint32_t)stage * kItemBytes + (uint32_t)(crow * 128 + ((cpiece ^ (2 * (crow & 3))) << 4));
#pragma unroll
        for (int i = 0; i < 4; i++)  // rows crow + 4i: (row & 3) does not change
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 512), "l"(src + (size_t)(4 * i) * halfK) : "memory");
    };
    const uint32_t rd_off = (uint32_t)(g * 128 + (t4 & 1) * 8);
    const int rd_piece = t4 >> 1, rd_swz = 2 * (g & 3);
    auto load_stat = [&](Stat<NESTED>& it, int tile, int ch) {
        const int sa = (tile * kTileRows + g) * bpr + ch * 4, sb = sa + 8 * bpr;
        if (NESTED) {
            it.st[0] = __ldg(reinterpret_cast<const uint32_t*>(qabs + sa));
            it.st[1] = __ldg(reinterpret_cast<const uint32_t*>(qabs + sb));
            it.st[2] = __float_as_uint(__ldg(am2 + (sa >> shift2)));
            it.st[3] = __float_as_uint(__ldg(am2 + (sb >> shift2)));
        } else {
            const float4 fa = __ldg(reinterpret_cast<const float4*>(am2 + sa)), fb = __ldg(reinterpret_cast<const float4*>(am2 + sb));
            it.st[0] = __float_as_uint(fa.x); it.st[1] = __float_as_uint(fa.y); it.st[2] = __float_as_uint(fa.z); it.st[3] = __float_as_uint(fa.w);
            it.st[4] = __float_as_uint(fb.x); it.st[5] = __float_as_uint(fb.y); it.st[6] = __float_as_uint(fb.z); it.st[7] = __float_as_uint(fb.w);
        }
    };

    bool waited = false;
    for (int pass0 = 0; pass0 < my_tiles; pass0 += kPassTiles) {
        const int np = my_tiles - pass0 < kPassTiles ? my_tiles - pass0 : kPassTiles;
        const int nitems = nchw * np;  // item i: chunk warp + 16 * (i / np), tile bx + (pass0 + i % np) * G
        // running (tile-in-pass, chunk) of the item being copied (kStages - 1 ahead) and of the item whose statistics are loaded (1 ahead)
        int lt = 0, lch = warp, qt = 0, qch = warp;
        auto advance = [&](int& tt, int& cc) {
            if (++tt == np) {
                tt = 0;
                cc += kWarps;
            }
        };
        auto tile_of = [&](int t) { return bx + (pass0 + t) * G; };
#pragma unroll
        for (int i = 0; i < kStages - 1; i++) {
            if (i < nitems) {
                copy_item(i, tile_of(lt), lch);
                advance(lt, lch);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        Stat<NESTED> cur, nxt;
        if (nitems > 0) {
            load_stat(cur, tile_of(qt), qch);
            advance(qt, qch);
        }
        if (!waited) {
            // weights and statistics do not depend on the previous kernel; the activations do
            pdl_wait();
            uint32_t done;
            do {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(0u) : "memory");
            } while (!done);
            waited = true;
        }
        uint32_t xr[32];
        int t = 0, ch = warp, stage = 0;
        for (int i = 0; i < nitems; i++) {
            if (i + 1 < nitems) {
                load_stat(nxt, tile_of(qt), qch);
                advance(qt, qch);
            }
            if (t == 0) {
                // B operand of the chunk: column g = token g; the lane feeds k slots (2*t4, 2*t4+1, 2*t4+8, 2*t4+9) of MMA m of block b
                // with x[16*t4 + 4*m .. + 3] of the block: 32 contiguous bytes per block
                if (g < c.tokens) {
                    const uint4* px = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(c.x) + (size_t)g * K + ch * 256 + t4 * 16);
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const uint4 v0 = __ldcg(px + b * 8), v1 = __ldcg(px + b * 8 + 1);  // coherent: the previous kernel's output
                        xr[8 * b] = v0.x; xr[8 * b + 1] = v0.y; xr[8 * b + 2] = v0.z; xr[8 * b + 3] = v0.w;
                        xr[8 * b + 4] = v1.x; xr[8 * b + 5] = v1.y; xr[8 * b + 6] = v1.z; xr[8 * b + 7] = v1.w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 32; q++) xr[q] = 0u;
                }
            }
            // item i has landed (the kStages - 2 younger groups may be pending); every lane is past its reads of item i - 1, whose
            // stage the next copy overwrites
            asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 2) : "memory");
            __syncwarp();
            if (i + kStages - 1 < nitems) {
                int ws = stage + kStages - 1;
                if (ws >= kStages) ws -= kStages;
                copy_item(ws, tile_of(lt), lch);
                advance(lt, lch);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            // the lane's 8 bytes [8*t4, 8*t4 + 8) of block b of rows g (wa) and g + 8 (wb)
            uint2 wa[4], wb[4];
            {
                const uint32_t sa = ring + (uint32_t)stage * kItemBytes + rd_off;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint32_t ad = sa + (uint32_t)(((2 * b + rd_piece) ^ rd_swz) << 4);
                    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(wa[b].x), "=r"(wa[b].y) : "r"(ad));
                    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(wb[b].x), "=r"(wb[b].y) : "r"(ad + 1024));
                }
            }
            // ---- 16 rows x 256 k: MMA j = 4*b + m covers bytes 2m, 2m+1 of the lane's 8 bytes of block b
            uint32_t f[kAhead + 1][4];
            auto fetch = [&](uint32_t (&d)[4], int j) {
                const int b = j >> 2, m = j & 3;
                const uint32_t va = (m & 2) ? wa[b].y : wa[b].x, vb = (m & 2) ? wb[b].y : wb[b].x;
                if (m & 1) {
                    d[0] = lut_lookup<2, kImm>(va, lane_base); d[1] = lut_lookup<2, kImm>(vb, lane_base);
                    d[2] = lut_lookup<3, kImm>(va, lane_base); d[3] = lut_lookup<3, kImm>(vb, lane_base);
                } else {
                    d[0] = lut_lookup<0, kImm>(va, lane_base); d[1] = lut_lookup<0, kImm>(vb, lane_base);
                    d[2] = lut_lookup<1, kImm>(va, lane_base); d[3] = lut_lookup<1, kImm>(vb, lane_base);
                }
            };
#pragma unroll
            for (int j = 0; j < kAhead; j++) fetch(f[j], j);
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int b = 0; b < 4; b++) {
                float d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const int j = 4 * b + m;
                    if (j + kAhead < 16) fetch(f[(j + kAhead) % (kAhead + 1)], j + kAhead);
                    uint32_t(&a4)[4] = f[j % (kAhead + 1)];
                    Hmma<T>::run(d, a4[0], a4[1], a4[2], a4[3], xr[8 * b + 2 * m], xr[8 * b + 2 * m + 1]);
                }
                float ama, amb;
                if (NESTED) {
                    const float qa = __uint_as_float(b == 0   ? lut_lookup<0, kImm + 128>(cur.st[0], lane_base)
                                                     : b == 1 ? lut_lookup<1, kImm + 128>(cur.st[0], lane_base)
                                                     : b == 2 ? lut_lookup<2, kImm + 128>(cur.st[0], lane_base)
                                                              : lut_lookup<3, kImm + 128>(cur.st[0], lane_base));
                    const float qb = __uint_as_float(b == 0   ? lut_lookup<0, kImm + 128>(cur.st[1], lane_base)
                                                     : b == 1 ? lut_lookup<1, kImm + 128>(cur.st[1], lane_base)
                                                     : b == 2 ? lut_lookup<2, kImm + 128>(cur.st[1], lane_base)
                                                              : lut_lookup<3, kImm + 128>(cur.st[1], lane_base));
                    ama = __fadd_rn(__fmul_rn(qa, __uint_as_float(cur.st[2])), offset);  // reference: kernels.cu:552 then core.py:468
                    amb = __fadd_rn(__fmul_rn(qb, __uint_as_float(cur.st[3])), offset);
                } else {
                    ama = __uint_as_float(cur.st[b]);
                    amb = __uint_as_float(cur.st[4 + b]);
                }
                acc[0] = fmaf(d[0], ama, acc[0]);
                acc[1] = fmaf(d[1], ama, acc[1]);
                acc[2] = fmaf(d[2], amb, acc[2]);
                acc[3] = fmaf(d[3], amb, acc[3]);
            }
            // the warp's partial sums of this tile: first chunk stores, later chunks add (the slot is private to the warp)
            float* p = s_part + ((size_t)(t * kWarps + warp) * 4) * 32 + lane;
            if (ch == warp) {
#pragma unroll
                for (int q = 0; q < 4; q++) p[q * 32] = acc[q];
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++) p[q * 32] += acc[q];
            }
            if (++t == np) {
                t = 0;
                ch += kWarps;
            }
            if (++stage == kStages) stage = 0;
            cur = nxt;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // ---- the CTA's outputs of this pass: element e = (acc index q, lane) of a tile -> row g + 8*(q>>1), token 2*t4 + (q&1)
        for (int u = tid; u < np * 128; u += kThreads) {
            const int tl = u >> 7, e = u & 127;
            const int q = e >> 5, l = e & 31;
            const int row = tile_of(tl) * kTileRows + (l >> 2) + 8 * (q >> 1), token = 2 * (l & 3) + (q & 1);
            if (token < c.tokens) {
                float sum = 0.0f;
                for (int w = 0; w < nw; w++) sum += s_part[(size_t)(tl * kWarps + w) * 128 + e];
                if (c.bias) sum += Elem<T>::to_f32(reinterpret_cast<const T*>(c.bias)[row]);
                reinterpret_cast<T*>(c.out)[(size_t)token * N + row] = Elem<T>::from_f32(sum);
            }
        }
        __syncthreads();
    }
}

template <typename T>
static int launch_tokens(const Args& a, bool nested, bool pdl, cudaStream_t stream)
{
    auto kern = nested ? gemv_tokens_kernel<T, true> : gemv_tokens_kernel<T, false>;
    static bool attr_set_dev[kMaxDevices][2] = {};
    bool& attr_set = attr_set_dev[device_slot()][nested];
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) return Q4_ERR_SHAPE;
        attr_set = true;
    }
    const int ntiles = a.N >> 4;
    const int sms = sm_count();
    const int grid = ntiles < sms ? ntiles : sms;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
    if (e != cudaSuccess) return (int)e;
    return finish_launch();
}

}  // namespace tok

// 1..16 tokens; Q4_ERR_SHAPE (nothing launched) when the kernel does not cover the shape
int gemv_4bit_tokens(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias, void* out,
                     int tokens, int64_t N, int64_t K, int blocksize, int dtype, int flags, const void* lut, cudaStream_t stream)
{
    (void)code;  // the table image carries the code
    if (blocksize != 64 || K <= 0 || (K & 255) || (N & 15) || N > (1 << 24) || K > (1 << 20) || tokens < 1 || tokens > 16) return Q4_ERR_SHAPE;
    if (N * (K >> 6) > 0x7fffffffll) return Q4_ERR_SHAPE;  // 32-bit block indices
    if (!x || !B || !out || !lut) return Q4_ERR_NULL;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(lut)) & 15) return Q4_ERR_ALIGN;
    if (reinterpret_cast<uintptr_t>(B) & 7) return Q4_ERR_ALIGN;
    if (int e = check_stats(stats)) return e;
    const bool nested = stats->qabsmax != nullptr;
    if (nested ? (reinterpret_cast<uintptr_t>(stats->qabsmax) & 3) != 0 : (reinterpret_cast<uintptr_t>(stats->absmax) & 15) != 0) return Q4_ERR_ALIGN;
    if (dtype != Q4_F16 && dtype != Q4_BF16) return Q4_ERR_DTYPE;
    if (g_dyn_base_probed(stream) != kDynBase) return Q4_ERR_SHAPE;
    const size_t esz = 2;
    for (int t0 = 0; t0 < tokens; t0 += tok::kMaxTokens) {
        tok::Args a;
        a.x = reinterpret_cast<const uint8_t*>(x) + (size_t)t0 * K * esz;
        a.B = B;
        a.s = make_view(stats);
        a.lut = lut;
        a.bias = bias;
        a.out = reinterpret_cast<uint8_t*>(out) + (size_t)t0 * N * esz;
        a.N = (int)N;
        a.K = (int)K;
        a.tokens = tokens - t0 < tok::kMaxTokens ? tokens - t0 : tok::kMaxTokens;
        // only the first pass may start before its predecessor ends (the second reads nothing the first writes, but both are
        // ordinary stream work for whatever follows)
        const bool pdl = (flags & Q4_GEMV_PDL) != 0;
        const int rc = dtype == Q4_F16 ? tok::launch_tokens<__half>(a, nested, pdl, stream) : tok::launch_tokens<__nv_bfloat16>(a, nested, pdl, stream);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace q4
