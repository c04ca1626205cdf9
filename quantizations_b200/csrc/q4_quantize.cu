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

template <typename T, int QT>
static int launch_quantize(const float* code, const T* A, float* absmax, uint8_t* out, int blocksize, int64_t n,
                           cudaStream_t stream)
{
    if (n <= 0) return 0;
    const int64_t nblocks = (n + blocksize - 1) / blocksize;
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
