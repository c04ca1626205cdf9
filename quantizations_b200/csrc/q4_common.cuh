// q4_common.cuh -- shared device helpers for the sm_100a 4-bit Linear kernels.
//
// Bit-exactness contract: quantize and dequantize outputs are compared byte-for-byte with the
// reference's kernels (csrc/kernels.cu), so every place where the reference performs an fp32 operation is written
// with an explicit round-to-nearest intrinsic (__fmul_rn / __fadd_rn / __frcp_rn) that the compiler may not
// contract or reassociate.  The data structures here are ours; only the arithmetic is pinned.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/quantizations_b200.h"

namespace q4 {

// ---------------------------------------------------------------------------------------------- element traits

template <typename T> struct Elem;
template <> struct Elem<float> {
    static __device__ __forceinline__ float to_f32(float v) { return v; }
    static __device__ __forceinline__ float from_f32(float v) { return v; }
};
template <> struct Elem<__half> {
    static __device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f32(float v) { return __float2half_rn(v); }
};
template <> struct Elem<__nv_bfloat16> {
    static __device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_f32(float v) { return __float2bfloat16_rn(v); }
};

// 16-bit pair (as it sits in a 32-bit register, element 0 in the low half) -> two floats
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t w);
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t w)
{
    return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t w)
{
    // bf16 -> f32 is a 16-bit shift
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b)
{
    __half2 h = __halves2half2(__float2half_rn(a), __float2half_rn(b));
    return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b)
{
    __nv_bfloat162 h = __halves2bfloat162(__float2bfloat16_rn(a), __float2bfloat16_rn(b));
    return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------------------------------------- memory helpers

// streaming loads: weights are read exactly once per call -> keep them out of L1 (guide: Guideline 13/14)
__device__ __forceinline__ uint4 ldg_stream_128(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// 256-bit load: one full 32-byte DRAM sector per lane (sm_100+: LDG.E.256)
struct u32x8 { uint32_t v[8]; };
__device__ __forceinline__ u32x8 ldg_stream_256(const void* p)
{
    u32x8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                   "=r"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_32(const void* p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_128(void* p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

// Programmatic dependent launch (guide: Guideline 9).  wait(): block until the preceding kernel in the stream has
// completed and its writes are visible; launch_dependents(): let the next kernel's CTAs start their prologue.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- absmax decode

// Device-side view of q4_absmax_t (passed by value to kernels).
struct AbsmaxView {
    const float* absmax;
    const uint8_t* qabsmax;
    const float* code2;
    const float* absmax2;
    const float* offset;
    int shift2;  // log2(blocksize2)
};

// absmax of 4-bit block `b`.  Nested: one fp32 multiply (reference kernels.cu:552) then one fp32 add
// (reference core.py:468 / :615, a separate torch kernel) -- never fused into an FMA.
template <bool NESTED>
__device__ __forceinline__ float load_absmax(const AbsmaxView& s, int64_t b, float offset)
{
    if (NESTED) {
        float c = __ldg(s.code2 + __ldg(s.qabsmax + b));
        float a2 = __ldg(s.absmax2 + (b >> s.shift2));
        return __fadd_rn(__fmul_rn(c, a2), offset);
    }
    return __ldg(s.absmax + b);
}

// ---------------------------------------------------------------------------------------------- 4-bit codecs

// FP4 magnitude thresholds, ascending, as the reference's float literals (kernels.cu:141-159).  0.583333f is the
// reference's literal, not 7/12.
#define Q4_FP4_T0 0.00260417f
#define Q4_FP4_T1 0.0859375f
#define Q4_FP4_T2 0.20833333f
#define Q4_FP4_T3 0.29166667f
#define Q4_FP4_T4 0.4166667f
#define Q4_FP4_T5 0.583333f
#define Q4_FP4_T6 0.8333333f

// FP4 encode of a normalised value: the reference's comparison tree (kernels.cu:113-163) is a rank over seven sorted
// thresholds; the code of rank r is nibble r of 0x32547610.  Branch-free: 7 compares + shift.
__device__ __forceinline__ uint32_t encode_fp4(float x)
{
    uint32_t sign = x < 0.0f ? 8u : 0u;
    float a = fabsf(x);
    int rank = (a > Q4_FP4_T0) + (a > Q4_FP4_T1) + (a > Q4_FP4_T2) + (a > Q4_FP4_T3) + (a > Q4_FP4_T4) +
               (a > Q4_FP4_T5) + (a > Q4_FP4_T6);
    return ((0x32547610u >> (4 * rank)) & 0xFu) + sign;
}

// NF4 encode: index = number of the 15 midpoints of the NF4 table (reference kernels.cu:851) strictly below x
// (upstream bitsandbytes dQuantizeNF4; the reference has no NF4 quantiser -- parity unpinned, see oracle header).
__device__ __forceinline__ uint32_t encode_nf4(float x)
{
    return (x > -0.8480964004993439f) + (x > -0.6106329262256622f) + (x > -0.4599952697753906f) +
           (x > -0.33967943489551544f) + (x > -0.23460740596055984f) + (x > -0.13791173323988914f) +
           (x > -0.045525018125772476f) + (x > 0.03979014977812767f) + (x > 0.1202552504837513f) +
           (x > 0.2035212516784668f) + (x > 0.2920137718319893f) + (x > 0.3893125355243683f) +
           (x > 0.5016634166240692f) + (x > 0.6427869200706482f) + (x > 0.8614784181118011f);
}

// Signed decode tables.  FP4: the reference's dequant-tree constants (kernels.cu:92-110) with the sign folded in;
// (m * absmax) * (+-1) == (+-m) * absmax bit-for-bit, and entry 8 is -0.0 like the tree's 0*absmax*(-1).
static __device__ __constant__ float kFp4Decode[16] = {0.00000000f,  5.208333333e-03f,  0.66666667f,  1.00000000f,
                                                0.33333333f,  0.50000000f,       0.16666667f,  0.25000000f,
                                                -0.00000000f, -5.208333333e-03f, -0.66666667f, -1.00000000f,
                                                -0.33333333f, -0.50000000f,      -0.16666667f, -0.25000000f};
static __device__ __constant__ float kNf4Decode[16] = {-1.0f,
                                                -0.6961928009986877f,
                                                -0.5250730514526367f,
                                                -0.39491748809814453f,
                                                -0.28444138169288635f,
                                                -0.18477343022823334f,
                                                -0.09105003625154495f,
                                                0.0f,
                                                0.07958029955625534f,
                                                0.16093020141124725f,
                                                0.24611230194568634f,
                                                0.33791524171829224f,
                                                0.44070982933044434f,
                                                0.5626170039176941f,
                                                0.7229568362236023f,
                                                1.0f};

// 8-bit codebook encode, deterministic.  Same decisions as the reference's bisection (kernels.cu:183-237): seven
// halvings from pivot 127 with strict '>', remembering the last table entry seen on each side, then one midpoint test
// ((neighbour + val) * 0.5f, strict) toward the side x lies on.  `code` is a 256-float table in shared memory.
__device__ __forceinline__ uint32_t encode_8bit(const float* code, float x)
{
    int pivot = 127, above_idx = 255, below_idx = 0;
    float below = -1.0f, above = 1.0f;
    float val = code[pivot];
#pragma unroll
    for (int step = 64; step > 0; step >>= 1) {
        bool up = x > val;
        below_idx = up ? pivot : below_idx;
        below = up ? val : below;
        above_idx = up ? above_idx : pivot;
        above = up ? above : val;
        pivot += up ? step : -step;
        val = code[pivot];
    }
    if (above_idx == 255) above = code[255];
    if (below_idx == 0) below = code[0];
    if (x > val) {
        float mid = __fmul_rn(__fadd_rn(above, val), 0.5f);
        return x > mid ? above_idx : pivot;
    }
    float mid = __fmul_rn(__fadd_rn(below, val), 0.5f);
    return x < mid ? below_idx : pivot;
}

// host: q4_absmax_t (C ABI) -> AbsmaxView, and its validation
inline AbsmaxView make_view(const q4_absmax_t* st)
{
    AbsmaxView v;
    v.absmax = st->absmax;
    v.qabsmax = st->qabsmax;
    v.code2 = st->code2;
    v.absmax2 = st->absmax2;
    v.offset = st->offset;
    v.shift2 = 0;
    if (st->qabsmax)
        while ((1 << v.shift2) < st->blocksize2) v.shift2++;
    return v;
}

__host__ __device__ __forceinline__ bool valid_blocksize(int bs)
{
    return bs == 64 || bs == 128 || bs == 256 || bs == 512 || bs == 1024 || bs == 2048 || bs == 4096;
}

}  // namespace q4

namespace q4 {
inline int check_stats(const q4_absmax_t* st)
{
    if (!st) return Q4_ERR_NULL;
    if (st->qabsmax) {
        if (!st->code2 || !st->absmax2 || !st->offset) return Q4_ERR_NULL;
        if (!valid_blocksize(st->blocksize2)) return Q4_ERR_BLOCKSIZE;
    } else if (!st->absmax) {
        return Q4_ERR_NULL;
    }
    return 0;
}
}  // namespace q4
