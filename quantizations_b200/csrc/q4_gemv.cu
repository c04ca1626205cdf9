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
//  gemv_mma_kernel     (q4_gemv_mma.cuh; fp16 / bf16 activations, blocksize 64, K % 128 == 0): one shared-memory lookup per
//    packed byte, multiply-accumulate and k-reduction on the tensor pipe (mma.sync), table by TMA bulk copy, half-SM CTAs so
//    that under programmatic dependent launch the next launch's x-independent prologue overlaps this launch's compute.
//
//  gemv_generic_kernel (any even K, any valid blocksize, "exact" fp32 arithmetic: w = code*absmax, acc = fma(x,w,acc)
//    as the reference's T=float instance does) -- warp per row, x staged in shared memory as fp32.
#include <cstdlib>
#include <type_traits>

#include "q4_common.cuh"
#include "q4_gemv_mma.cuh"
#include "q4_gemv_tc.cuh"
#include "q4_launch.h"

namespace q4 {

// ------------------------------------------------------------------------------------------------ fast path

unsigned long long* g_gemv_trace = nullptr;  // set by q4_debug_set_gemv_trace (developer tool, not part of the ABI)
constexpr int kDefaultGridRule = 0;          // default of Q4_GEMV_GRID (see gemv_dispatch)

// ------------------------------------------------------------------------------------------------ generic path

template <typename T, bool NESTED>
__global__ void __launch_bounds__(256)
gemv_generic_kernel(const T* __restrict__ x, const uint8_t* __restrict__ Bq, AbsmaxView s, const float* __restrict__ code,
                    const T* __restrict__ bias, T* __restrict__ out, int64_t N, int64_t K, int bs_shift)
{
    extern __shared__ __align__(1024) uint8_t smem[];
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
    const void* lut = nullptr;  // prebuilt table image (q4_gemv_lut_build) for this code / code2 / dtype
    void* workspace = nullptr;  // split-K workspace of the tcgen05 kernel (zeroed once by the caller)
    int64_t workspace_bytes = 0;
    const q4_allreduce_t* ar = nullptr;  // fused all-reduce over tensor-parallel ranks
    int tokens = 1;                      // > 1: small-batch call (x [tokens, K], out [tokens, N]): tcgen05 kernel only
    int64_t next_K = 0;                  // in_features of the weight behind the prefetch hint (0: unknown)
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

// A prepared (not yet launched) stage of the mma.sync kernel: what q4_gemv_4bit_chain collects before its single launch.
struct MmaStage {
    MmaGemvArgs a;
    bool nested, ktail;
};
constexpr int kStagePrepared = 1, kStageNotChainable = 2;  // positive returns of gemv_dispatch(..., stage_out)

template <typename T, bool NESTED>
static void* mma_kernel_ptr(bool compact, bool ktail, bool chain, bool swiglu)
{
    using ChainT = void (*)(const MmaChainArgs);
    using KernT = void (*)(const MmaSingleArgs);
    if (swiglu) return (compact && !ktail && !chain) ? (void*)(KernT)gemv_mma_kernel<T, NESTED, true, false, false, true> : nullptr;
    if (chain) return (void*)(ChainT)gemv_mma_kernel<T, NESTED, true, false, true>;  // chains: compact layout, no ragged shapes
    KernT k = compact ? (ktail ? (KernT)gemv_mma_kernel<T, NESTED, true, true, false> : (KernT)gemv_mma_kernel<T, NESTED, true, false, false>)
                      : (ktail ? (KernT)gemv_mma_kernel<T, NESTED, false, true, false> : (KernT)gemv_mma_kernel<T, NESTED, false, false, false>);
    return (void*)k;
}

// one launch of the mma.sync kernel over `n` prepared stages (n == 1: the plain GEMV)
template <typename T>
static int launch_mma(const MmaChainArgs& c, bool nested, bool compact, bool ktail, int grid, size_t smem, bool pdl, cudaStream_t stream)
{
    const bool chain = c.n > 1;
    const bool swiglu = c.st[0].swiglu != 0;
    void* kern = nested ? mma_kernel_ptr<T, true>(compact, ktail, chain, swiglu) : mma_kernel_ptr<T, false>(compact, ktail, chain, swiglu);
    if (!kern) return Q4_ERR_SHAPE;  // the SwiGLU epilogue exists for the compact layout and whole tiles only
    static bool attr_set_dev[kMaxDevices][2][2][2][2][2] = {};
    auto& attr_set = attr_set_dev[device_slot()];
    if (!attr_set[nested][compact][ktail][chain][swiglu]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
        attr_set[nested][compact][ktail][chain][swiglu] = true;
    }
    if (chain) return launch_pdl((void (*)(const MmaChainArgs))kern, dim3(grid), dim3(kMmaThreads), smem, stream, pdl, c);
    MmaSingleArgs one = {};
    one.st[0] = c.st[0];
    one.n = 1;
    one.x_bytes = c.x_bytes;
    return launch_pdl((void (*)(const MmaSingleArgs))kern, dim3(grid), dim3(kMmaThreads), smem, stream, pdl, one);
}

static int g_dyn_base = -1;  // where dynamic shared memory starts in a CTA's window (probed once)

// A property of the architecture / driver, not of a device: probed once per process (outside stream capture) and shared by the
// kernels whose table address is a compile-time constant (q4_gemv_mma.cuh compact layout, q4_gemv_ring.cuh).
int g_dyn_base_probed(cudaStream_t stream)
{
    if (g_dyn_base < 0) {
        uint32_t* d = nullptr;
        uint32_t h = 0;
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(stream, &cap);
        if (cap == cudaStreamCaptureStatusNone && cudaMalloc(&d, 4) == cudaSuccess) {
            probe_dyn_smem_base_kernel<<<1, 32, 1024, stream>>>(d);
            if (cudaMemcpyAsync(&h, d, 4, cudaMemcpyDeviceToHost, stream) == cudaSuccess && cudaStreamSynchronize(stream) == cudaSuccess)
                g_dyn_base = (int)h;
            cudaFree(d);
        }
    }
    return g_dyn_base;
}

template <typename T>
static int gemv_dispatch(const T* x, const uint8_t* B, const q4_absmax_t* st, const float* code, const T* bias, T* out,
                         int64_t N, int64_t K, int blocksize, int flags, const void* next, int64_t next_bytes, cudaStream_t stream,
                         int nmat = 1, const float* const* offsets = nullptr, const int* row_end = nullptr,
                         const GemvPrologue* pro = nullptr, MmaStage* stage_out = nullptr)
{
    const AbsmaxView v = make_view(st);
    const bool nested = st->qabsmax != nullptr;
    const bool pdl = flags & Q4_GEMV_PDL;
    const int sms = sm_count();
    const int kw = (int)((K + 2047) / 2048);
    const bool fast = !(flags & Q4_GEMV_EXACT_F32) && blocksize == 64 && (K % 64) == 0 && kw <= 16 && K >= 64 && N * (K / 64) < (1ll << 31) && N < (1 << 30) &&
                      (reinterpret_cast<uintptr_t>(B) & 31) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                      (!nested || st->blocksize2 >= 64);
    if constexpr (sizeof(T) == 2) {
        // tensor-pipe streaming kernel (q4_gemv_mma.cuh): K % 128 == 0 so that a lane's two absmax entries are one aligned pair
        const bool mma_ok = fast && (K % 128) == 0 && K <= 32768 &&
                            (!nested || ((reinterpret_cast<uintptr_t>(st->qabsmax) & 1) == 0 && st->blocksize2 >= 128)) &&
                            (nested || (reinterpret_cast<uintptr_t>(st->absmax) & 7) == 0);
        // where does dynamic shared memory start in the CTA's window?  (probed once; the compact table layout depends on it)
        const int dyn_base = g_dyn_base_probed(stream);
        static const int env_impl = getenv("Q4_GEMV_IMPL") ? atoi(getenv("Q4_GEMV_IMPL")) : 0;  // 1: force the mma.sync kernel
        // tcgen05 kernel (q4_gemv_tc.cuh): needs the prebuilt table image and the split-K workspace
        if (fast && env_impl != 1 && !stage_out && !(flags & Q4_GEMV_SWIGLU) && pro && pro->lut && pro->workspace && !pro->ar && dyn_base == kDynBase && K <= 65536 && (K % 256) == 0 &&
            (reinterpret_cast<uintptr_t>(pro->workspace) & 15) == 0) {
            const int bpr = (int)(K / 64);
            const int rt_total = (int)((N + kTcRows - 1) / kTcRows);
            const int64_t U = (int64_t)rt_total * bpr;
            int grid = (int)(U / kTcGroups < sms ? U / kTcGroups : sms);
            const int tokens = pro->tokens;
            const int mt = tokens > 1 ? 16 : 1;
            const size_t smem = (size_t)kLutBytes + (mt == 1 ? (size_t)K * 2 + kTcXPad : (size_t)kTcGroups * kTcA * 2048) + 128 +
                                (1 + 6 * kTcGroups) * 8 + 128;
            if (grid >= 1 && smem <= 226 * 1024) {
                const int G2 = kTcGroups * grid;
                auto run_of = [&](int64_t u) { return (int)(((u + 1) * G2 - 1) / U); };
                int max_seg = 1;
                for (int rt = 0; rt < rt_total; rt++) {
                    const int n = run_of((int64_t)rt * bpr + bpr - 1) - run_of((int64_t)rt * bpr) + 1;
                    if (n > max_seg) max_seg = n;
                }
                // counters live in a FIXED region at the start (they must stay zero between launches of any shape), partials after it
                const size_t cnt_bytes = 64 * 1024;
                if ((size_t)rt_total * 4 > cnt_bytes) return Q4_ERR_SHAPE;
                const size_t need = cnt_bytes + (size_t)rt_total * max_seg * mt * kTcRows * 4;
                if ((int64_t)need <= pro->workspace_bytes) {
                    const bool multi = nmat > 1 && nested;
                    TcGemvArgs a = {};
                    a.x = x;
                    a.lut = pro->lut;
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
                    if ((reinterpret_cast<uintptr_t>(pro->x_gate) & 15) || (reinterpret_cast<uintptr_t>(pro->rms_weight) & 15) ||
                        (reinterpret_cast<uintptr_t>(pro->lut) & 15))
                        return Q4_ERR_ALIGN;
                    a.x_gate = pro->x_gate;
                    a.rms_weight = pro->rms_weight;
                    a.rms_eps = pro->rms_eps;
                    a.next = (reinterpret_cast<uintptr_t>(next) & 15) == 0 ? (const uint8_t*)next : nullptr;
                    a.next_bytes = a.next ? next_bytes : 0;
                    a.ws_count = reinterpret_cast<int*>(pro->workspace);
                    a.ws_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(pro->workspace) + cnt_bytes);
                    a.max_seg = max_seg;
                    a.rows = (int)N;
                    a.K = (int)K;
                    a.rt_total = rt_total;
                    a.trace = g_gemv_trace;
                    static const int env_debug = getenv("Q4_GEMV_DEBUG") ? atoi(getenv("Q4_GEMV_DEBUG")) : 0;
                    a.debug = env_debug;
                    a.tokens = tokens;
                    if (mt > 1 && (multi || a.x_gate || a.rms_weight)) return Q4_ERR_SHAPE;
                    auto kern = mt > 1 ? (nested ? gemv_tc_kernel<T, true, false, 16> : gemv_tc_kernel<T, false, false, 16>)
                                       : (nested ? (multi ? gemv_tc_kernel<T, true, true, 1> : gemv_tc_kernel<T, true, false, 1>)
                                                 : gemv_tc_kernel<T, false, false, 1>);
                    static bool attr_set_dev[kMaxDevices][2][2][2] = {};
                    auto& attr_set = attr_set_dev[device_slot()];
                    if (!attr_set[nested][multi][mt > 1]) {
                        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
                        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                        if (e != cudaSuccess) return (int)e;
                        attr_set[nested][multi][mt > 1] = true;
                    }
                    return launch_pdl(kern, dim3(grid), dim3(kTcThreads), smem, stream, pdl, a);
                }
            }
        }
        if (pro && pro->tokens > 1) return Q4_ERR_SHAPE;  // small batches exist only on the tcgen05 kernel
        if (mma_ok) {
            const bool multi = nmat > 1 && nested;
            MmaGemvArgs a = {};
            a.x = x;
            a.code = code;
            a.lut = pro ? pro->lut : nullptr;
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
                if ((reinterpret_cast<uintptr_t>(pro->x_gate) & 15) || (reinterpret_cast<uintptr_t>(pro->rms_weight) & 15) ||
                    (reinterpret_cast<uintptr_t>(pro->lut) & 15))
                    return Q4_ERR_ALIGN;
                a.x_gate = pro->x_gate;
                a.rms_weight = pro->rms_weight;
                a.rms_eps = pro->rms_eps;
            }
            a.next = (reinterpret_cast<uintptr_t>(next) & 15) == 0 ? (const uint8_t*)next : nullptr;
            a.next_bytes = a.next ? next_bytes : 0;
            if (pro && pro->ar && pro->ar->world > 1) {
                const q4_allreduce_t* ar = pro->ar;
                if (!ar->peer_bases) return Q4_ERR_NULL;
                // one row count per exchange area: the CTA -> rows mapping (and with it the double-buffer argument) must be identical in
                // every launch that shares it
                if (ar->world > kArMaxWorld || ar->rank < 0 || ar->rank >= ar->world || N != ar->max_rows) return Q4_ERR_SHAPE;
                a.ar_peer_bases = ar->peer_bases;
                a.ar_world = ar->world;
                a.ar_rank = ar->rank;
                a.ar_max_rows = ar->max_rows;
            }
            a.rows = (int)N;
            a.K = (int)K;
            a.rt_total = (int)((N + 7) / 8);
            a.kt = (int)((K + 511) / 512);
            constexpr int threads = kMmaThreads;
            {
                const int nw = threads / 32, bpr_ = (int)(K / 64);
                const long long row_bytes = (long long)bpr_ * 32;
                a.d_rt = nw / a.kt;
                a.d_kt = nw % a.kt;
                a.pw_step = (long long)a.d_rt * 8 * row_bytes + (long long)a.d_kt * 256;
                a.pw_wrap = 8 * row_bytes - (long long)a.kt * 256;
                a.ab_step = a.d_rt * 8 * bpr_ + a.d_kt * 8;
                a.ab_wrap = 8 * bpr_ - a.kt * 8;
            }
            a.trace = g_gemv_trace;
            static const int env_debug_mma = getenv("Q4_GEMV_DEBUG") ? atoi(getenv("Q4_GEMV_DEBUG")) : 0;
            a.debug = env_debug_mma;
            a.multi = multi ? 1 : 0;
            if (flags & Q4_GEMV_SWIGLU) {
                // interleaved gate/up pair: whole 8-row tiles, the two members equally long, nothing else in the epilogue
                if (nmat != 2 || (N % 8) != 0 || !row_end || row_end[0] * 2 != N || bias || (pro && pro->ar && pro->ar->world > 1) || stage_out)
                    return Q4_ERR_SHAPE;
                a.swiglu = 1;
            }
            if (stage_out) {  // q4_gemv_4bit_chain: hand the prepared stage back instead of launching it
                stage_out->a = a;
                stage_out->nested = nested;
                stage_out->ktail = (K % 512) != 0 || (N % 8) != 0;
                return kStagePrepared;
            }
            static const int env_aligned = getenv("Q4_GEMV_ALIGNED") ? atoi(getenv("Q4_GEMV_ALIGNED")) : 0;
            // Two half-SM CTAs per SM whenever there are enough row tiles: measured 6.5 % faster over the whole Llama-3-8B stack than
            // leaving half of every SM to the next launch's prologue (1.37 vs 1.46 ms/step) -- the loop wants all 16 warps.  The rule
            // depends on the row count only, so tensor-parallel launches that share an exchange area keep one CTA -> rows mapping.
            static const int env_mult_raw = getenv("Q4_GEMV_GRID_MULT") ? atoi(getenv("Q4_GEMV_GRID_MULT")) : 0;
            // Balanced grid (Q4_GEMV_GRID=1): the SMALLEST grid with the same ceil(rt_total / grid) leaves CTA slots to the next
            // launch of a decode chain.  Measured SLOWER than filling every slot (1.35 vs 1.30 ms/step: the loop wants all 16 warps of
            // every SM), so the plain rule stays the default.  Q4_GEMV_GRID: 0 = plain rule, 1 = balanced, n > 1 = exactly n CTAs.
            static const int env_grid = getenv("Q4_GEMV_GRID") ? atoi(getenv("Q4_GEMV_GRID")) : kDefaultGridRule;
            // The rule depends on the shape only (tensor parallel: one CTA -> rows mapping per exchange area; the prefetch hint: this
            // launch must know how the NEXT one will split its rows).
            auto plan_grid = [&](int rt_total, int kt) {
                const int mult = env_mult_raw > 0 ? env_mult_raw : ((flags & Q4_GEMV_SHARE_SM) ? 1 : (rt_total >= 2 * sms ? 2 : 1));
                // grid: two half-SM CTAs per SM (or one, leaving the other half to the next launch's prologue, see the kernel), or more
                // when the per-CTA partial-sum buffer would not fit
                const size_t tail = (size_t)kt * 1024 + 128 + 16;
                const int cap_ctas = sms * mult;
                int grid = rt_total < cap_ctas ? rt_total : cap_ctas;
                for (;;) {
                    const size_t part = (size_t)((rt_total + grid - 1) / grid) * kt * 32;
                    if (kLutBytes + tail + part <= 100 * 1024 || 2 * (size_t)kLutBytes + tail + part <= 200 * 1024 || grid >= rt_total) break;
                    grid += sms;
                }
                if (grid > rt_total) grid = rt_total;
                if (env_grid == 1) {
                    const int per = (rt_total + grid - 1) / grid;
                    grid = (rt_total + per - 1) / per;
                } else if (env_grid > 1 && env_grid <= grid) {
                    const size_t part = (size_t)((rt_total + env_grid - 1) / env_grid) * kt * 32;
                    if (kLutBytes + tail + part <= 100 * 1024) grid = env_grid;
                }
                return grid;
            };
            const size_t tail = (size_t)a.kt * 1024 + 128 + 16;
            const int grid = plan_grid(a.rt_total, a.kt);
            // exact prefetch hint: how the next launch will split ITS rows, and how many row tiles each of its CTAs loads before its
            // activation exists (kBuffers tiles per warp, k fastest: ceil(kBuffers * warps / kt) row tiles)
            if (a.next && pro && pro->next_K > 0 && !(pro->next_K >= 512 && (pro->next_K % 512) == 0 && a.next_bytes % (pro->next_K * 4) == 0)) {
                a.next = nullptr;  // a hint with in_features the head computation does not cover (ragged k tiles): no prefetch rather than
                a.next_bytes = 0;  // the whole-range one
            }
            if (a.next && pro && pro->next_K >= 512 && (pro->next_K % 512) == 0 && a.next_bytes % (pro->next_K * 4) == 0) {
                const int64_t nK = pro->next_K, nrows = a.next_bytes * 2 / nK;
                const int nkt = (int)(nK / 512), nrt = (int)(nrows / 8);
                if (nrt > 0 && nK <= 32768) {
                    const int ngrid = plan_grid(nrt, nkt);
                    a.nx_grid = ngrid;
                    a.nx_rt_q = nrt / ngrid;
                    a.nx_rt_r = nrt % ngrid;
                    static const int env_head = getenv("Q4_GEMV_HEAD_MULT") ? atoi(getenv("Q4_GEMV_HEAD_MULT")) : 1;  // developer switches
                    static const int env_tiles = getenv("Q4_GEMV_HEAD_TILES") ? atoi(getenv("Q4_GEMV_HEAD_TILES")) : kBuffers;
                    a.nx_head = env_head * ((env_tiles * (kMmaThreads / 32) + nkt - 1) / nkt);
                    a.nx_tile_bytes = 4 * nK;  // 8 rows x K/2 bytes
                }
            }
            if (a.ar_world > 1) {
                // The fused all-reduce addresses a peer's slots by ROW and tags them with per-CTA epochs: every launch that shares an
                // exchange area must map rows to CTAs identically on all ranks, launch after launch.  That holds when the grid follows
                // from the row count alone; a grid changed by the shared-memory fit loop (depends on K) or by a developer override
                // would let the epochs of two row-parallel layers with equal N but different K drift apart -- refuse it.
                const int mult0 = (flags & Q4_GEMV_SHARE_SM) ? 1 : (a.rt_total >= 2 * sms ? 2 : 1);
                const int rows_only = a.rt_total < sms * mult0 ? a.rt_total : sms * mult0;
                if (grid != rows_only || grid > kArMaxCtas) return Q4_ERR_SHAPE;
            }
            const size_t rest = tail + (size_t)((a.rt_total + grid - 1) / grid) * a.kt * 32;
            const bool compact = dyn_base == kDynBase && !env_aligned && kLutBytes + rest <= 220 * 1024;
            const size_t smem = (compact ? 1 : 2) * (size_t)kLutBytes + rest;
            if (smem > 226 * 1024) return Q4_ERR_SHAPE;
            const bool ktail = (K % 512) != 0 || (N % 8) != 0;
            a.rt_q = a.rt_total / grid;
            a.rt_r = a.rt_total % grid;
            MmaChainArgs c = {};
            c.st[0] = a;
            c.n = 1;
            c.x_bytes = a.kt * 1024;
            return launch_mma<T>(c, nested, compact, ktail, grid, smem, pdl, stream);
        }
    }
    if (stage_out) return kStageNotChainable;
    if (nmat > 1 || (flags & Q4_GEMV_SWIGLU) || (pro && (pro->x_gate || pro->rms_weight || (pro->ar && pro->ar->world > 1)))) return Q4_ERR_SHAPE;  // only the fast path groups / fuses
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

static int gemv_4bit_fused_impl(const q4_gemv_fused_t* f, cudaStream_t stream, MmaStage* stage_out)
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
    pro.lut = f->lut;
    pro.workspace = f->workspace;
    pro.workspace_bytes = f->workspace_bytes;
    pro.ar = f->allreduce;
    pro.next_K = f->prefetch_K;
    const int flags = f->flags & ~Q4_GEMV_EXACT_F32;
    switch (f->dtype) {
        case Q4_F16:
            return gemv_dispatch<__half>((const __half*)f->x, f->B, f->stats, f->code, (const __half*)f->bias, (__half*)f->out,
                                         f->rows, f->K, f->blocksize, flags, f->prefetch, f->prefetch_bytes, stream, nmat, offsets,
                                         row_end, &pro, stage_out);
        case Q4_BF16:
            return gemv_dispatch<__nv_bfloat16>((const __nv_bfloat16*)f->x, f->B, f->stats, f->code, (const __nv_bfloat16*)f->bias,
                                                (__nv_bfloat16*)f->out, f->rows, f->K, f->blocksize, flags, f->prefetch,
                                                f->prefetch_bytes, stream, nmat, offsets, row_end, &pro, stage_out);
        default: return Q4_ERR_DTYPE;
    }
}

int gemv_4bit_fused(const q4_gemv_fused_t* f, cudaStream_t stream) { return gemv_4bit_fused_impl(f, stream, nullptr); }

// `n` dependent decode GEMVs (stage i+1 may read what stage i wrote) in ONE persistent launch: see MmaChainArgs.  Falls back to n
// ordinary launches whenever a stage cannot run on the chained kernel, so the result never depends on the path taken.
int gemv_4bit_chain(const q4_gemv_fused_t* stages, int n, void* barrier_ws, cudaStream_t stream)
{
    if (n < 1) return 0;
    if (!stages) return Q4_ERR_NULL;
    static const int env_nochain = getenv("Q4_GEMV_NO_CHAIN") ? atoi(getenv("Q4_GEMV_NO_CHAIN")) : 0;
    bool chain = n > 1 && n <= kMaxChain && barrier_ws && !env_nochain && (reinterpret_cast<uintptr_t>(barrier_ws) & 3) == 0;
    MmaStage st[kMaxChain];
    for (int i = 0; chain && i < n; i++) {
        if (stages[i].rows == 0) chain = false;
        else {
            const int rc = gemv_4bit_fused_impl(&stages[i], stream, &st[i]);
            if (rc < 0 || rc > kStageNotChainable) return rc;  // argument error (or a CUDA error from the one-time probe)
            if (rc != kStagePrepared) chain = false;
        }
    }
    for (int i = 0; chain && i < n; i++) {
        // one kernel instance for the whole chain: same element type, nesting, no ragged shapes, the same table image; tensor-
        // parallel stages keep their fixed CTA -> rows mapping only at the plain grid, so they are not chained with wider grids
        if (stages[i].dtype != stages[0].dtype || st[i].nested != st[0].nested || st[i].ktail || !st[i].a.lut || st[i].a.lut != st[0].a.lut)
            chain = false;
    }
    if (chain && g_dyn_base != kDynBase) chain = false;
    if (chain) {
        MmaChainArgs c = {};
        c.n = n;
        c.barrier = reinterpret_cast<unsigned*>(barrier_ws);
        const int sms = sm_count();
        bool any_ar = false;
        for (int i = 0; i < n; i++) {
            c.st[i] = st[i].a;
            if (c.st[i].kt * 1024 > c.x_bytes) c.x_bytes = c.st[i].kt * 1024;
            any_ar = any_ar || st[i].a.ar_world > 1;
        }
        // grid: every CTA must be resident at once (grid barrier).  Two half-SM CTAs per SM when the shared-memory plan allows it
        // (all 16 warps per SM work on the chain), else one; tensor-parallel stages need the plain one-per-SM mapping.
        int grid = 0;
        size_t smem = 0;
        if (any_ar) chain = false;  // tensor-parallel stages keep the CTA -> rows mapping of their single launches: not chained
        for (int per_sm = 2; chain && per_sm >= 1; per_sm--) {
            grid = sms * per_sm;
            size_t part = 0;
            for (int i = 0; i < n; i++) {
                const size_t p = (size_t)((c.st[i].rt_total + grid - 1) / grid) * c.st[i].kt * 32;
                if (p > part) part = p;
            }
            smem = (size_t)kLutBytes + c.x_bytes + 128 + 16 + part;
            if (smem <= (per_sm == 2 ? 112 * 1024 : 226 * 1024)) break;
            if (per_sm == 1) chain = false;
        }
        if (chain && any_ar && grid > kArMaxCtas) chain = false;
        if (chain) {
            for (int i = 0; i < n; i++) {
                c.st[i].rt_q = c.st[i].rt_total / grid;
                c.st[i].rt_r = c.st[i].rt_total % grid;
            }
            const bool pdl = stages[0].flags & Q4_GEMV_PDL;
            switch (stages[0].dtype) {
                case Q4_F16: return launch_mma<__half>(c, st[0].nested, true, false, grid, smem, pdl, stream);
                case Q4_BF16: return launch_mma<__nv_bfloat16>(c, st[0].nested, true, false, grid, smem, pdl, stream);
                default: return Q4_ERR_DTYPE;
            }
        }
    }
    for (int i = 0; i < n; i++)
        if (int rc = gemv_4bit_fused_impl(&stages[i], stream, nullptr)) return rc;
    return 0;
}

// 2..16 tokens in one pass over the packed weight: x [tokens, K], out [tokens, N].  Default: q4_gemv_tokens.cu (mma.sync, the B columns
// are the tokens); Q4_GEMV_BATCH_TC5 or a shape that kernel refuses: the tcgen05 kernel (N columns of the MMA = tokens)
int gemv_4bit_batch(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias, void* out,
                    int tokens, int64_t N, int64_t K, int blocksize, int dtype, int flags, const void* lut, void* workspace,
                    int64_t workspace_bytes, cudaStream_t stream)
{
    if (!valid_blocksize(blocksize)) return Q4_ERR_BLOCKSIZE;
    if (N < 0 || K < 0 || (K & 1) || tokens < 1 || tokens > 16) return Q4_ERR_SHAPE;
    if (N == 0) return 0;
    if (!x || !B || !code || !out || !lut) return Q4_ERR_NULL;
    if (int e = check_stats(stats)) return e;
    if (!(flags & Q4_GEMV_BATCH_TC5)) {
        // default: the mma.sync kernel with the tokens on the B columns (q4_gemv_tokens.cu); shapes it does not cover go on
        const int rc = gemv_4bit_tokens(x, B, stats, code, bias, out, tokens, N, K, blocksize, dtype, flags, lut, stream);
        if (rc != Q4_ERR_SHAPE) return rc;
    }
    if (!workspace) return Q4_ERR_NULL;
    GemvPrologue pro;
    pro.lut = lut;
    pro.workspace = workspace;
    pro.workspace_bytes = workspace_bytes;
    pro.tokens = tokens;
    flags &= ~(Q4_GEMV_EXACT_F32 | Q4_GEMV_BATCH_TC5);
    switch (dtype) {
        case Q4_F16:
            return gemv_dispatch<__half>((const __half*)x, B, stats, code, (const __half*)bias, (__half*)out, N, K, blocksize, flags, nullptr,
                                         0, stream, 1, nullptr, nullptr, &pro);
        case Q4_BF16:
            return gemv_dispatch<__nv_bfloat16>((const __nv_bfloat16*)x, B, stats, code, (const __nv_bfloat16*)bias, (__nv_bfloat16*)out, N, K,
                                                blocksize, flags, nullptr, 0, stream, 1, nullptr, nullptr, &pro);
        default: return Q4_ERR_DTYPE;
    }
}

int gemv_lut_build(const float* code, const float* code2, int dtype, void* lut, cudaStream_t stream)
{
    if (!code || !lut) return Q4_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(lut) & 15) return Q4_ERR_ALIGN;
    const int threads = 256, grid = kLutBytes / 4 / threads;
    switch (dtype) {
        case Q4_F16: gemv_lut_build_kernel<__half><<<grid, threads, 0, stream>>>(code, code2, (uint32_t*)lut); break;
        case Q4_BF16: gemv_lut_build_kernel<__nv_bfloat16><<<grid, threads, 0, stream>>>(code, code2, (uint32_t*)lut); break;
        default: return Q4_ERR_DTYPE;
    }
    return finish_launch();
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
