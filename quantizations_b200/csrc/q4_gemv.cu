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

template <typename T, bool NESTED>
__global__ void __launch_bounds__(512, 1)
gemv_lut256_kernel(const T* __restrict__ x, const uint8_t* __restrict__ Bq, AbsmaxView s, const float* __restrict__ code,
                   const T* __restrict__ bias, T* __restrict__ out, int N, int K, int kw, int groups)
{
    constexpr int U = kRowsInFlight;
    extern __shared__ __align__(16) uint8_t smem[];
    // [0, 64K): lookup table.  [64K, ...): cross-warp partials, float[2][groups][kw][U]; then the staged source tables.
    float* s_part = reinterpret_cast<float*>(smem + kLutBytes);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int lgroup = warp / kw;        // row group inside the CTA
    const int kpos = warp - lgroup * kw; // which 2048-wide K slice this warp covers
    const int G = gridDim.x * groups;    // row groups in the grid
    const int gid = blockIdx.x + gridDim.x * lgroup;  // consecutive rows land on different SMs
    const int kblk = kpos * 32 + lane;   // this thread's 64-wide block inside a row
    const int bpr = K >> 6;              // blocks per row
    const bool active = kblk < bpr;

    // Let the next kernel in the stream start its own prologue (its griddepcontrol.wait still orders it after us).
    pdl_launch_dependents();

    // ---- 1. put the first U rows' loads in flight (weights and statistics do not depend on the previous kernel)
    u32x8 w[U];
    uint32_t qa[U];
    float a2[U];
    const int rows_mine = gid < N ? (N - gid + G - 1) / G : 0;
    auto issue = [&](int batch) {
#pragma unroll
        for (int i = 0; i < U; i++) {
            const int r = gid + (batch * U + i) * G;
            if (active && r < N) {
                const int64_t blk = (int64_t)r * bpr + kblk;
                w[i] = ldg_stream_256(Bq + blk * 32);
                if (NESTED) {
                    qa[i] = __ldg(s.qabsmax + blk);
                    a2[i] = __ldg(s.absmax2 + (blk >> s.shift2));
                } else {
                    a2[i] = __ldg(s.absmax + blk);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) w[i].v[j] = 0;
                qa[i] = 0;
                a2[i] = 0.0f;
            }
        }
    };
    issue(0);

    // ---- 2. build the per-lane-replicated table: 128-B segment 2b = half2{code[b>>4], code[b&15]} x32,
    //         segment 2b+1 = code2[b] x32.  The two source tables are staged in shared memory with ONE global round
    //         trip; threads then write consecutive 16-B chunks (conflict-free).
    float* s_src = s_part + 2 * groups * kw * U;  // [0,256): code2, [256,272): code
    if (NESTED && tid < 256) s_src[tid] = __ldg(s.code2 + tid);
    if (tid >= blockDim.x - 16) s_src[256 + (tid - (blockDim.x - 16))] = __ldg(code + (tid - (blockDim.x - 16)));
    __syncthreads();
    for (int c = tid; c < kLutBytes / 16; c += blockDim.x) {
        const int seg = c >> 3, b = seg >> 1;
        uint32_t word;
        if (seg & 1) {
            word = NESTED ? __float_as_uint(s_src[b]) : 0u;
        } else {
            __half2 h = __halves2half2(__float2half_rn(s_src[256 + (b >> 4)]), __float2half_rn(s_src[256 + (b & 15)]));
            word = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(smem + c * 16) = make_uint4(word, word, word, word);
    }
    const float offset = NESTED ? __ldg(s.offset) : 0.0f;

    // ---- 3. everything below may read the previous kernel's output
    pdl_wait();

    // ---- 4. this thread's 64 activations -> half2 registers, scaled by a power of two so |x| < 2
    uint32_t xh[32];
    float unscale = 1.0f;
    if (active) {
        float xf[64];
        load_x64<T>(x + (int64_t)kblk * 64, xf);
        float m = 0.0f;
#pragma unroll
        for (int j = 0; j < 64; j++) m = fmaxf(m, fabsf(xf[j]));
        int e = (int)((__float_as_uint(m) >> 23) & 0xFF);  // biased exponent of the largest |x|
        e = e < 1 ? 1 : (e > 253 ? 253 : e);
        const float scale = __uint_as_float((uint32_t)(254 - e) << 23);  // 2^(127-e): m*scale in [1,2)
        unscale = __uint_as_float((uint32_t)e << 23);                     // 2^(e-127)
#pragma unroll
        for (int j = 0; j < 32; j++) {
            __half2 h = __floats2half2_rn(xf[2 * j] * scale, xf[2 * j + 1] * scale);
            xh[j] = *reinterpret_cast<uint32_t*>(&h);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; j++) xh[j] = 0;
    }
    __syncthreads();  // table visible

    const uint32_t lane4 = lane * 4;  // < 256: the upper three bytes are the zeros PRMT index 5 picks up
    const int nbatch = (rows_mine + U - 1) / U;
    int buf = 0;

    for (int batch = 0; batch < nbatch; batch++) {
        float t[U];
#pragma unroll
        for (int i = 0; i < U; i++) {
            uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t wj = w[i].v[j];
                // table offset = byte*256 + lane*4: one PRMT splices weight byte b into byte 1 above lane*4
                acc0 = hfma2(lut_at(smem, __byte_perm(wj, lane4, 0x5504)), xh[4 * j + 0], acc0);
                acc1 = hfma2(lut_at(smem, __byte_perm(wj, lane4, 0x5514)), xh[4 * j + 1], acc1);
                acc2 = hfma2(lut_at(smem, __byte_perm(wj, lane4, 0x5524)), xh[4 * j + 2], acc2);
                acc3 = hfma2(lut_at(smem, __byte_perm(wj, lane4, 0x5534)), xh[4 * j + 3], acc3);
            }
            const uint32_t sum = hadd2(hadd2(acc0, acc1), hadd2(acc2, acc3));
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&sum));
            float am;
            if (NESTED) {
                const float c2 = __uint_as_float(lut_at(smem, __byte_perm(qa[i], lane4, 0x5504) + 128));
                am = __fadd_rn(__fmul_rn(c2, a2[i]), offset);  // reference: kernels.cu:552 then core.py:468
            } else {
                am = a2[i];
            }
            t[i] = (f.x + f.y) * (am * unscale);
        }
        // next batch's loads go out before the reduction so they overlap it
        if (batch + 1 < nbatch) issue(batch + 1);

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
            const int r = gid + (batch * U + i) * G;
            if (kpos == 0 && (lane & 7) == 0 && r < N) {
                T y = Elem<T>::from_f32(total);
                if (bias) y = Elem<T>::from_f32(Elem<T>::to_f32(y) + Elem<T>::to_f32(bias[r]));  // torch `out += bias`
                out[r] = y;
            }
        }
    }
}

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
                         int64_t N, int64_t K, int blocksize, int flags, cudaStream_t stream)
{
    const AbsmaxView v = make_view(st);
    const bool nested = st->qabsmax != nullptr;
    const bool pdl = flags & Q4_GEMV_PDL;
    const int sms = sm_count();
    const int kw = (int)((K + 2047) / 2048);
    const bool fast = !(flags & Q4_GEMV_EXACT_F32) && blocksize == 64 && (K % 64) == 0 && kw <= 16 && N < (1 << 30) &&
                      (reinterpret_cast<uintptr_t>(B) & 31) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                      (!nested || st->blocksize2 >= 64);
    if (fast) {
        const int groups = 16 / kw;
        const int threads = groups * kw * 32;
        const size_t smem = kLutBytes + sizeof(float) * (2 * groups * kw * kRowsInFlight + 272);
        auto kern = nested ? gemv_lut256_kernel<T, true> : gemv_lut256_kernel<T, false>;
        static bool attr_set[2] = {false, false};
        if (!attr_set[nested]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (e != cudaSuccess) return (int)e;
            attr_set[nested] = true;
        }
        const int64_t want = (N + groups - 1) / groups;  // CTAs needed to give every group one row
        const int grid = (int)(want < sms ? want : sms);
        return launch_pdl(kern, dim3(grid), dim3(threads), smem, stream, pdl, x, B, v, code, bias, out, (int)N, (int)K, kw,
                          groups);
    }
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

int gemv_4bit(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias, void* out,
              int64_t N, int64_t K, int blocksize, int dtype, int flags, cudaStream_t stream)
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
                                        flags | Q4_GEMV_EXACT_F32, stream);
        case Q4_F16:
            return gemv_dispatch<__half>((const __half*)x, B, stats, code, (const __half*)bias, (__half*)out, N, K, blocksize,
                                         flags, stream);
        case Q4_BF16:
            return gemv_dispatch<__nv_bfloat16>((const __nv_bfloat16*)x, B, stats, code, (const __nv_bfloat16*)bias,
                                                (__nv_bfloat16*)out, N, K, blocksize, flags, stream);
        default: return Q4_ERR_DTYPE;
    }
}

}  // namespace q4
