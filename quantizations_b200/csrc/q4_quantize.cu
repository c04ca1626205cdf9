// q4_quantize.cu -- blockwise quantize kernels (4-bit FP4/NF4 and the 8-bit codebook used for double-quant).
//
// Replaces: reference csrc/kernels.cu:340-478 (kQuantizeBlockwise) + launcher csrc/ops.cu:53-95.
// Contract: bit-exact packed bytes and absmax with the reference for FP4 (all input types) and for the 8-bit
// codebook.  Roofline: HBM -- reads sizeof(T)*n, writes n/2 + 4*n/bs (4-bit).
//
// Design (not the reference's): no CUB, no shared-memory transposes.  A thread owns 8 consecutive elements
// (one 128-bit load for 16-bit inputs), so a quantization block of BS elements is owned by BS/8 adjacent lanes and
// its absmax is a shuffle-only max reduction for BS <= 256 (the 4-bit default, 64, needs 3 shuffles); larger
// blocks add one shared-memory hop across warps.  Each thread emits its 8 nibbles as one 32-bit store (4-bit) or its
// 8 codes as one 64-bit store (8-bit), so loads and stores are fully coalesced.  max() is exact under any
// association, so the different reduction shape cannot change a bit.
#include "q4_common.cuh"
#include "q4_encode_lut.h"
#include "q4_launch.h"

namespace q4 {

template <typename T> __device__ __forceinline__ void load8(const T* A, int64_t e0, int64_t n, float (&v)[8]);

template <> __device__ __forceinline__ void load8<float>(const float* A, int64_t e0, int64_t n, float (&v)[8])
{
    if (e0 + 8 <= n) {
        uint4 a = ldg_stream_128(A + e0), b = ldg_stream_128(A + e0 + 4);
        v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
        v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = (e0 + j < n) ? A[e0 + j] : 0.0f;  // out-of-range slots read as 0 (kernels.cu:410)
    }
}

template <typename T> __device__ __forceinline__ void load8(const T* A, int64_t e0, int64_t n, float (&v)[8])
{
    if (e0 + 8 <= n) {
        uint4 a = ldg_stream_128(A + e0);
        float2 p0 = unpack2<T>(a.x), p1 = unpack2<T>(a.y), p2 = unpack2<T>(a.z), p3 = unpack2<T>(a.w);
        v[0] = p0.x; v[1] = p0.y; v[2] = p1.x; v[3] = p1.y; v[4] = p2.x; v[5] = p2.y; v[6] = p3.x; v[7] = p3.y;
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = (e0 + j < n) ? Elem<T>::to_f32(A[e0 + j]) : 0.0f;
    }
}

template <typename T, int BS, int QT>
__global__ void __launch_bounds__((BS / 8 > 256) ? BS / 8 : 256)
quantize_blockwise_kernel(const float* __restrict__ code, const T* __restrict__ A, float* __restrict__ absmax,
                          uint8_t* __restrict__ out, int64_t n, int64_t nblocks)
{
    constexpr int TPB = BS / 8;                    // threads that share one quantization block
    constexpr int CTA = TPB > 256 ? TPB : 256;     // threads per CTA
    constexpr int WPB = TPB > 32 ? TPB / 32 : 1;   // warps per quantization block (when TPB > 32)
    __shared__ float s_code[QT == Q4_GENERAL8BIT ? 256 : 1];
    __shared__ float s_warpmax[CTA / 32];

    if (QT == Q4_GENERAL8BIT) {
        for (int i = threadIdx.x; i < 256; i += CTA) s_code[i] = code[i];
        __syncthreads();
    }

    const int64_t t = (int64_t)blockIdx.x * CTA + threadIdx.x;  // global index of this thread's 8-element group
    const int64_t e0 = t * 8;
    float v[8];
    load8<T>(A, e0, n, v);

    // absmax: fold fmaxf(|v|) from -FLT_MAX (fmaxf drops NaN) exactly as kernels.cu:406,419
    float m = -FLT_MAX;
#pragma unroll
    for (int j = 0; j < 8; j++) m = fmaxf(m, fabsf(v[j]));
#pragma unroll
    for (int o = (TPB < 32 ? TPB : 32) / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if constexpr (TPB > 32) {
        const int warp = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) s_warpmax[warp] = m;
        __syncthreads();
        const int first = (warp / WPB) * WPB;
#pragma unroll
        for (int w = 0; w < WPB; w++) m = fmaxf(m, s_warpmax[first + w]);
    }

    const int64_t qb = t / TPB;  // quantization block index
    if (qb >= nblocks) return;
    if (t % TPB == 0) absmax[qb] = m;

    const float inv = __fdiv_rn(1.0f, m);  // IEEE divide, kernels.cu:438

    if (QT == Q4_GENERAL8BIT) {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) lo |= encode_8bit(s_code, __fmul_rn(v[j], inv)) << (8 * j);
#pragma unroll
        for (int j = 0; j < 4; j++) hi |= encode_8bit(s_code, __fmul_rn(v[4 + j], inv)) << (8 * j);
        if (e0 + 8 <= n) {
            *reinterpret_cast<uint2*>(out + e0) = make_uint2(lo, hi);
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (e0 + j < n) out[e0 + j] = (uint8_t)(((j < 4 ? lo : hi) >> (8 * (j & 3))) & 0xFF);
        }
    } else {
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float a = __fmul_rn(v[2 * j], inv), b = __fmul_rn(v[2 * j + 1], inv);
            uint32_t byte = QT == Q4_NF4 ? ((encode_nf4(a) << 4) | encode_nf4(b)) : ((encode_fp4(a) << 4) | encode_fp4(b));
            word |= byte << (8 * j);
        }
        if (BS >= 1024) {
            // Reference quirk kept for bit-parity (kernels.cu:450,465-470): with 4 values per thread the packing
            // accumulator is never cleared, so each odd byte of a block is OR-ed with the even byte before it.
            word |= (word & 0x00ff00ffu) << 8;
        }
        const int64_t byte0 = e0 >> 1;
        if (e0 + 8 <= n) {
            *reinterpret_cast<uint32_t*>(out + byte0) = word;
        } else {
            const int nbytes = (int)((n - e0 + 1) >> 1);  // (valid+1)/2 bytes are stored, kernels.cu:476
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (j < nbytes) out[byte0 + j] = (uint8_t)((word >> (8 * j)) & 0xFF);
        }
    }
}

// ---------------------------------------------------------------------------------------------- 4-bit fast path
//
// The kernel above spends ~50 instructions per element on the compare trees; at HBM rate the SM has ~11 (16-bit input).  This
// one encodes through the binned table of q4_encode_lut.h: one fused multiply-add (exact bin), one 8-byte shared-memory lookup
// {threshold inside the bin, code below it}, one compare.  The table is replicated 16 times (entry stride 128 B, replica =
// lane % 16), so the half-warp phases of a 64-bit lookup never conflict whatever the data; persistent CTAs (4 per SM) build
// it once and stream whole 2048-element chunks, four 128-bit loads in flight per thread.  Blocks whose absmax is not a
// normal finite number (all-zero, denormal, inf / NaN inputs) take the compare trees, so every special value encodes exactly
// as before.  Used for blocksize <= 256 (absmax by shuffles only) and n % 8 == 0.
constexpr int kEncReplicas = 16;
constexpr int kEncLutBytes = Q4_ENC_BINS * kEncReplicas * 8;

template <int QT> __device__ __forceinline__ uint32_t encode_lut(uint32_t lut_lane, float x)
{
    if (QT == Q4_NF4) {
        const float xc = fmaxf(x, -1.0f);  // NaN -> -1 -> code 0, like fifteen false compares
        const uint32_t bits = __float_as_uint(__fmaf_rn(xc, 128.0f, 8388736.0f));
        uint32_t thr, base;
        // (bits << 7) = 0x80000000 + bin * 128: the constant is folded into lut_lane (32-bit wrap-around)
        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(thr), "=r"(base) : "r"(bits * 128u + lut_lane));
        return base + (xc > __uint_as_float(thr) ? 1u : 0u);
    } else {
        const float a = fmaxf(fabsf(x), 0.0f);  // NaN -> 0 -> rank 0
        const uint32_t bits = __float_as_uint(__fmaf_rn(a, 256.0f, 8388608.0f));
        uint32_t thr, codes;
        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(thr), "=r"(codes) : "r"(bits * 128u + lut_lane));
        const uint32_t c = a > __uint_as_float(thr) ? (codes >> 4) : (codes & 0xFu);
        return c + (x < 0.0f ? 8u : 0u);
    }
}

template <typename T, int BS, int QT, int EPT>
__global__ void __launch_bounds__(256, 4)
quantize_4bit_lut_kernel(const T* __restrict__ A, float* __restrict__ absmax, uint8_t* __restrict__ out, int64_t ngroups)
{
    // A thread owns EPT consecutive elements (16 for 16-bit inputs: 32 bytes in, 8 bytes out, one reciprocal per 16 elements and
    // only log2(BS/16) shuffles for the block maximum; 8 for fp32 inputs or when n is not a multiple of 16).
    constexpr int TPB = BS / EPT;                         // lanes that share one quantization block (<= 32)
    constexpr int W = EPT * (int)sizeof(T) / 4;           // 32-bit words a thread loads per group (4 or 8)
    constexpr int UNROLL = W == 8 ? 2 : 4;                // 64 bytes in flight per thread
    extern __shared__ __align__(16) uint8_t enc_smem[];
    for (int b = threadIdx.x; b < Q4_ENC_BINS; b += 256) {  // replica 0 of every bin ...
        float thr;
        uint32_t word;
        q4_enc_entry(QT == Q4_NF4, b, &thr, &word);
        *reinterpret_cast<uint2*>(enc_smem + (size_t)b * kEncReplicas * 8) = make_uint2(__float_as_uint(thr), word);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Q4_ENC_BINS * kEncReplicas; i += 256)  // ... copied to the other fifteen
        if (i % kEncReplicas) *reinterpret_cast<uint2*>(enc_smem + (size_t)i * 8) = *reinterpret_cast<const uint2*>(enc_smem + (size_t)(i - i % kEncReplicas) * 8);
    __syncthreads();
    const uint32_t lut_lane = (uint32_t)__cvta_generic_to_shared(enc_smem) + (threadIdx.x & 15) * 8 - 0x80000000u;

    auto elem = [](const uint32_t (&raw)[W], int j) -> float {
        if constexpr (sizeof(T) == 4) return __uint_as_float(raw[j]);
        else {
            const float2 p = unpack2<T>(raw[j >> 1]);
            return (j & 1) ? p.y : p.x;
        }
    };
    const int64_t nchunks = (ngroups + 255) / 256;
    for (int64_t c0 = (int64_t)blockIdx.x * UNROLL; c0 < nchunks; c0 += (int64_t)gridDim.x * UNROLL) {
        uint32_t raw[UNROLL][W];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t g = (c0 + u) * 256 + threadIdx.x;
#pragma unroll
            for (int j = 0; j < W; j++) raw[u][j] = 0;  // out-of-range slots read as 0 (kernels.cu:410)
            if (g < ngroups) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(A) + g * (W * 4);
#pragma unroll
                for (int h = 0; h < W / 4; h++) {
                    const uint4 q4v = ldg_stream_128(src + 16 * h);
                    raw[u][4 * h] = q4v.x; raw[u][4 * h + 1] = q4v.y; raw[u][4 * h + 2] = q4v.z; raw[u][4 * h + 3] = q4v.w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int64_t g = (c0 + u) * 256 + threadIdx.x;
            float m = -FLT_MAX;
#pragma unroll
            for (int j = 0; j < EPT; j++) m = fmaxf(m, fabsf(elem(raw[u], j)));
#pragma unroll
            for (int o = TPB / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (g >= ngroups) continue;
            if ((threadIdx.x & (TPB - 1)) == 0) absmax[g / TPB] = m;
            const float inv = __fdiv_rn(1.0f, m);  // IEEE divide, kernels.cu:438
            uint32_t word[EPT / 8];
#pragma unroll
            for (int h = 0; h < EPT / 8; h++) word[h] = 0;
            if (m >= FLT_MIN && m <= FLT_MAX) {  // |v * inv| <= 1 + 2^-23 (or v is NaN): the binned encoder's domain
#pragma unroll
                for (int j = 0; j < EPT; j++)
                    word[j >> 3] |= encode_lut<QT>(lut_lane, __fmul_rn(elem(raw[u], j), inv)) << (8 * ((j & 7) >> 1) + ((j & 1) ? 0 : 4));
            } else {
#pragma unroll
                for (int j = 0; j < EPT; j++) {
                    const float xn = __fmul_rn(elem(raw[u], j), inv);
                    word[j >> 3] |= (QT == Q4_NF4 ? encode_nf4(xn) : encode_fp4(xn)) << (8 * ((j & 7) >> 1) + ((j & 1) ? 0 : 4));
                }
            }
            if constexpr (EPT == 16) *reinterpret_cast<uint2*>(out + g * 8) = make_uint2(word[0], word[1]);
            else *reinterpret_cast<uint32_t*>(out + g * 4) = word[0];
        }
    }
}

template <typename T, int QT>
static int launch_quantize(const float* code, const T* A, float* absmax, uint8_t* out, int blocksize, int64_t n,
                           cudaStream_t stream)
{
    if (n <= 0) return 0;
    const int64_t nblocks = (n + blocksize - 1) / blocksize;
    if constexpr (QT != Q4_GENERAL8BIT) {
        if ((n & 7) == 0 && blocksize <= 256) {
            // 16 elements per thread for 16-bit inputs (n % 16 == 0, 8-byte aligned output), else 8
            const bool wide = sizeof(T) == 2 && (n & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0;
            const int64_t ngroups = wide ? n / 16 : n / 8;
            const int unroll = (wide || sizeof(T) == 4) ? 2 : 4;
            const int64_t want = (ngroups + 256 * unroll - 1) / (256 * unroll);
            const int64_t cap = (int64_t)sm_count() * 4;
            const unsigned grid = (unsigned)(want < cap ? want : cap);
#define Q4_LUT_LAUNCH(BS, EPT)                                                                                            \
    {                                                                                                                     \
        static bool attr_dev[kMaxDevices] = {};                                                                           \
        bool& attr = attr_dev[device_slot()];                                                                             \
        if (!attr) {                                                                                                      \
            cudaError_t e = cudaFuncSetAttribute(quantize_4bit_lut_kernel<T, BS, QT, EPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEncLutBytes); \
            if (e != cudaSuccess) return (int)e;                                                                          \
            attr = true;                                                                                                  \
        }                                                                                                                 \
        quantize_4bit_lut_kernel<T, BS, QT, EPT><<<grid, 256, kEncLutBytes, stream>>>(A, absmax, out, ngroups);          \
        return finish_launch();                                                                                           \
    }
#define Q4_LUT_CASE(BS)                                                                                                   \
    case BS:                                                                                                              \
        if constexpr (sizeof(T) == 2) {                                                                                   \
            if (wide) Q4_LUT_LAUNCH(BS, 16)                                                                               \
        }                                                                                                                 \
        Q4_LUT_LAUNCH(BS, 8)
            switch (blocksize) {
                Q4_LUT_CASE(64)
                Q4_LUT_CASE(128)
                Q4_LUT_CASE(256)
                default: break;
            }
#undef Q4_LUT_CASE
#undef Q4_LUT_LAUNCH
        }
    }
#define Q4_CASE(BS)                                                                                              \
    case BS: {                                                                                                   \
        constexpr int CTA = (BS / 8 > 256) ? BS / 8 : 256;                                                       \
        const int64_t threads = nblocks * (BS / 8);                                                              \
        const int64_t grid = (threads + CTA - 1) / CTA;                                                          \
        quantize_blockwise_kernel<T, BS, QT><<<(unsigned)grid, CTA, 0, stream>>>(code, A, absmax, out, n, nblocks); \
        break;                                                                                                   \
    }
    switch (blocksize) {
        Q4_CASE(64)
        Q4_CASE(128)
        Q4_CASE(256)
        Q4_CASE(512)
        Q4_CASE(1024)
        Q4_CASE(2048)
        Q4_CASE(4096)
        default: return Q4_ERR_BLOCKSIZE;
    }
#undef Q4_CASE
    return finish_launch();
}

int quantize_4bit(const void* A, float* absmax, uint8_t* out, int blocksize, int64_t n, int quant_type, int in_dtype,
                  cudaStream_t stream)
{
    if (!valid_blocksize(blocksize)) return Q4_ERR_BLOCKSIZE;
    if (n < 0) return Q4_ERR_SHAPE;
    if (n > 0 && (!A || !absmax || !out)) return Q4_ERR_NULL;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(out) & 3)) return Q4_ERR_ALIGN;
    if (quant_type != Q4_FP4 && quant_type != Q4_NF4) return Q4_ERR_QUANT_TYPE;
#define Q4_DISPATCH(T)                                                                                     \
    return quant_type == Q4_FP4 ? launch_quantize<T, Q4_FP4>(nullptr, (const T*)A, absmax, out, blocksize, n, stream) \
                                : launch_quantize<T, Q4_NF4>(nullptr, (const T*)A, absmax, out, blocksize, n, stream)
    switch (in_dtype) {
        case Q4_F32: Q4_DISPATCH(float);
        case Q4_F16: Q4_DISPATCH(__half);
        case Q4_BF16: Q4_DISPATCH(__nv_bfloat16);
        default: return Q4_ERR_DTYPE;
    }
#undef Q4_DISPATCH
}

int quantize_8bit(const float* code, const float* A, float* absmax, uint8_t* out, int blocksize, int64_t n,
                  cudaStream_t stream)
{
    if (!valid_blocksize(blocksize)) return Q4_ERR_BLOCKSIZE;
    if (n < 0) return Q4_ERR_SHAPE;
    if (n > 0 && (!code || !A || !absmax || !out)) return Q4_ERR_NULL;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(out) & 7)) return Q4_ERR_ALIGN;
    return launch_quantize<float, Q4_GENERAL8BIT>(code, A, absmax, out, blocksize, n, stream);
}

}  // namespace q4
