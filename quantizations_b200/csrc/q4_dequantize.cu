// q4_dequantize.cu -- blockwise dequantize kernels (4-bit FP4/NF4 -> fp16/bf16/fp32, 8-bit codebook -> fp32).
//
// Replaces: reference csrc/kernels.cu:480-568 (kDequantizeBlockwise) + launcher csrc/ops.cu:97-128, and for nested
// statistics also the two launches before it (dequantize_blockwise + `absmax += offset`, core.py:613-617), which
// are decoded inside this kernel with the reference's exact fp32 multiply-then-add.
// Contract: output bit-exact with the reference (FP4) / the oracle (NF4).  Roofline: HBM -- per element reads
// 1/2 + 1/bs bytes and writes sizeof(T).
//
// Design: a thread owns 8 consecutive outputs = one 32-bit packed word, so a warp reads 128 contiguous bytes and
// writes 512 (16-bit) or 1024 (fp32) contiguous bytes per step with 128-bit stores; eight independent steps are in
// flight per thread.  The 16-entry decode table and the 8-bit absmax map are staged in shared memory.
#include "q4_common.cuh"
#include "q4_launch.h"

namespace q4 {

template <typename T> __device__ __forceinline__ void store8(T* out, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* out, const float (&v)[8])
{
    stg_stream_128(out, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
    stg_stream_128(out + 4, make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
}
template <typename T> __device__ __forceinline__ void store8(T* out, const float (&v)[8])
{
    stg_stream_128(out, make_uint4(pack2<T>(v[0], v[1]), pack2<T>(v[2], v[3]), pack2<T>(v[4], v[5]), pack2<T>(v[6], v[7])));
}

// value of one nibble: signed table entry times absmax, one fp32 multiply (== (m*absmax)*sign of kernels.cu:92-110).
// `table` is the 16-entry decode table staged in shared memory: 16 words in 16 distinct banks, so any mix of
// indices across a warp is conflict-free (a constant-bank lookup would serialise on divergent indices).
__device__ __forceinline__ float decode_nibble(const float* table, uint32_t nib, float am)
{
    return __fmul_rn(table[nib], am);
}

template <typename T, int QT, bool NESTED>
__global__ void __launch_bounds__(256)
dequantize_4bit_kernel(const uint8_t* __restrict__ A, AbsmaxView s, T* __restrict__ out, int bs_shift, int64_t n)
{
    constexpr int STEPS = 8;  // independent 128-byte reads / 512-byte (16-bit) writes per warp in flight
    __shared__ float s_table[16];
    __shared__ float s_code2[NESTED ? 256 : 1];
    if (threadIdx.x < 16) s_table[threadIdx.x] = QT == Q4_NF4 ? kNf4Decode[threadIdx.x] : kFp4Decode[threadIdx.x];
    if (NESTED) s_code2[threadIdx.x] = __ldg(s.code2 + threadIdx.x);  // 8-bit absmax map: a lookup that depends on a loaded byte
    const float offset = NESTED ? __ldg(s.offset) : 0.0f;              // should not be a second trip to L2
    const int64_t ngroups = (n + 7) >> 3;  // 8-element groups
    const int64_t g0 = (int64_t)blockIdx.x * (256 * STEPS) + threadIdx.x;
    const bool whole = ((int64_t)(blockIdx.x + 1) * (256 * STEPS)) * 8 <= n;  // CTA-uniform: no bounds, no ragged group

    uint32_t word[STEPS];
    float am[STEPS];   // nested: second-level absmax until decoded
    uint32_t q[STEPS];
    if (whole) {
#pragma unroll
        for (int i = 0; i < STEPS; i++) {
            const int64_t e0 = (g0 + i * 256) << 3;
            word[i] = ldg_stream_32(A + (e0 >> 1));
            const int64_t b = e0 >> bs_shift;
            if (NESTED) {
                q[i] = __ldg(s.qabsmax + b);
                am[i] = __ldg(s.absmax2 + (b >> s.shift2));
            } else {
                am[i] = __ldg(s.absmax + b);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < STEPS; i++) {
            const int64_t g = g0 + i * 256;
            word[i] = 0;
            am[i] = 0.0f;
            q[i] = 0;
            if (g < ngroups) {
                const int64_t e0 = g << 3;
                if (e0 + 8 <= n) {
                    word[i] = ldg_stream_32(A + (e0 >> 1));
                } else {  // ragged tail: (n+1)/2 bytes exist
                    const int64_t nbytes = (n + 1) >> 1;
                    for (int j = 0; j < 4; j++)
                        if ((e0 >> 1) + j < nbytes) word[i] |= (uint32_t)A[(e0 >> 1) + j] << (8 * j);
                }
                const int64_t b = e0 >> bs_shift;
                if (NESTED) {
                    q[i] = __ldg(s.qabsmax + b);
                    am[i] = __ldg(s.absmax2 + (b >> s.shift2));
                } else {
                    am[i] = __ldg(s.absmax + b);
                }
            }
        }
    }
    __syncthreads();  // tables staged (the loads above are already in flight)
#pragma unroll
    for (int i = 0; i < STEPS; i++) {
        const int64_t g = g0 + i * 256;
        if (!whole && g >= ngroups) continue;
        const int64_t e0 = g << 3;
        // nested: one fp32 multiply (kernels.cu:552) then one fp32 add (core.py:615), never an FMA
        const float a = NESTED ? __fadd_rn(__fmul_rn(s_code2[q[i]], am[i]), offset) : am[i];
        float v[8];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t byte = (word[i] >> (8 * j)) & 0xFFu;
            v[2 * j] = decode_nibble(s_table, byte >> 4, a);
            v[2 * j + 1] = decode_nibble(s_table, byte & 0xFu, a);
        }
        if (whole || e0 + 8 <= n) {
            store8<T>(out + e0, v);
        } else {
            for (int j = 0; j < 8; j++)
                if (e0 + j < n) out[e0 + j] = Elem<T>::from_f32(v[j]);
        }
    }
}

__global__ void __launch_bounds__(256)
dequantize_8bit_kernel(const float* __restrict__ code, const uint8_t* __restrict__ A, const float* __restrict__ absmax,
                       float* __restrict__ out, int bs_shift, int64_t n)
{
    __shared__ float s_code[256];
    s_code[threadIdx.x] = code[threadIdx.x];
    __syncthreads();
    const int64_t e0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (e0 >= n) return;
    if (e0 + 4 <= n) {
        const uint32_t w = ldg_stream_32(A + e0);
        const float am = __ldg(absmax + (e0 >> bs_shift));  // 4 | blocksize: one absmax per thread
        float4 r;
        r.x = __fmul_rn(s_code[w & 0xFF], am);
        r.y = __fmul_rn(s_code[(w >> 8) & 0xFF], am);
        r.z = __fmul_rn(s_code[(w >> 16) & 0xFF], am);
        r.w = __fmul_rn(s_code[w >> 24], am);
        *reinterpret_cast<float4*>(out + e0) = r;
    } else {
        for (int64_t e = e0; e < n; e++) out[e] = __fmul_rn(s_code[A[e]], __ldg(absmax + (e >> bs_shift)));
    }
}

template <typename T>
static int launch_dequantize_4bit(const uint8_t* A, const q4_absmax_t* st, T* out, int blocksize, int64_t n, int quant_type,
                                  cudaStream_t stream)
{
    const AbsmaxView v = make_view(st);
    const int shift = ilog2(blocksize);
    const int64_t ngroups = (n + 7) / 8;
    const unsigned grid = (unsigned)((ngroups + 2047) / 2048);
    const bool nested = st->qabsmax != nullptr;
    if (quant_type == Q4_FP4) {
        if (nested) dequantize_4bit_kernel<T, Q4_FP4, true><<<grid, 256, 0, stream>>>(A, v, out, shift, n);
        else        dequantize_4bit_kernel<T, Q4_FP4, false><<<grid, 256, 0, stream>>>(A, v, out, shift, n);
    } else {
        if (nested) dequantize_4bit_kernel<T, Q4_NF4, true><<<grid, 256, 0, stream>>>(A, v, out, shift, n);
        else        dequantize_4bit_kernel<T, Q4_NF4, false><<<grid, 256, 0, stream>>>(A, v, out, shift, n);
    }
    return finish_launch();
}

int dequantize_4bit(const uint8_t* A, const q4_absmax_t* stats, void* out, int blocksize, int64_t n, int quant_type,
                    int out_dtype, cudaStream_t stream)
{
    if (!valid_blocksize(blocksize)) return Q4_ERR_BLOCKSIZE;
    if (n < 0) return Q4_ERR_SHAPE;
    if (quant_type != Q4_FP4 && quant_type != Q4_NF4) return Q4_ERR_QUANT_TYPE;
    if (n == 0) return 0;
    if (!A || !out) return Q4_ERR_NULL;
    if (int e = check_stats(stats)) return e;
    if ((reinterpret_cast<uintptr_t>(A) & 3) || (reinterpret_cast<uintptr_t>(out) & 15)) return Q4_ERR_ALIGN;
    switch (out_dtype) {
        case Q4_F32: return launch_dequantize_4bit<float>(A, stats, (float*)out, blocksize, n, quant_type, stream);
        case Q4_F16: return launch_dequantize_4bit<__half>(A, stats, (__half*)out, blocksize, n, quant_type, stream);
        case Q4_BF16: return launch_dequantize_4bit<__nv_bfloat16>(A, stats, (__nv_bfloat16*)out, blocksize, n, quant_type, stream);
        default: return Q4_ERR_DTYPE;
    }
}

int dequantize_8bit(const float* code, const uint8_t* A, const float* absmax, float* out, int blocksize, int64_t n,
                    cudaStream_t stream)
{
    if (!valid_blocksize(blocksize)) return Q4_ERR_BLOCKSIZE;
    if (n < 0) return Q4_ERR_SHAPE;
    if (n == 0) return 0;
    if (!code || !A || !absmax || !out) return Q4_ERR_NULL;
    if ((reinterpret_cast<uintptr_t>(A) & 3) || (reinterpret_cast<uintptr_t>(out) & 15)) return Q4_ERR_ALIGN;
    const unsigned grid = (unsigned)((n + 1023) / 1024);
    dequantize_8bit_kernel<<<grid, 256, 0, stream>>>(code, A, absmax, out, ilog2(blocksize), n);
    return finish_launch();
}

}  // namespace q4
