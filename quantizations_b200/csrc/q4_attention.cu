// q4_attention.cu -- the non-Linear half of a batch-1 decode step of the Llama harness (quantizations_b200/llama.py) in ONE
// launch per layer: RoPE on the new token's q and k, KV-cache append, grouped-query attention of that one query over the cache.
//
// Not part of the reference (it has no model code: README.md:52-126 patches HF's LlamaForCausalLM); it exists because, once the
// seven Linear4bit GEMVs of a layer are four launches (q4_gemv_4bit_fused), the ~12 small torch kernels HF-style attention
// issues per layer (slice / mul / sub / cat for RoPE, two index_copy, SDPA with a mask over the whole static cache) are more
// than half of the decode step.  Same arithmetic as llama.py's torch path: HF "rotate_half" RoPE with every product and the
// sum rounded to the activation type, softmax(q.k / sqrt(hd)) in fp32, output rounded once.
//
// One CTA per query head, 8 warps; a lane owns 4 of the 128 head dimensions (one 8-byte load per cached key / value row, a
// warp reads 256 contiguous bytes); warps deal the cached positions round-robin with an online softmax each and are combined
// through shared memory.  The new token's k/v are used from registers (and written to the cache by the first head of each
// KV group), so there is no dependency between CTAs.
#include "q4_common.cuh"
#include "q4_launch.h"

namespace q4 {

struct DecodeAttnArgs {
    const void* qkv;      // [nh*hd | nkv*hd | nkv*hd]: q, k, v of the new token (output of the grouped q/k/v GEMV)
    const void* cos_tab;  // [max_len, hd/2]
    const void* sin_tab;
    void* k_cache;        // [nkv, max_len, hd]
    void* v_cache;
    const long long* pos; // device scalar: position of the new token = number of tokens already cached
    void* out;            // [nh*hd]
    int nh, nkv, max_len;
    float scale;
};

constexpr int kHd = 128;
constexpr int kAttnWarps = 8;

template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4])
{
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack2<T>(u.x), b = unpack2<T>(u.y);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T> __device__ __forceinline__ float rnd(float v) { return Elem<T>::to_f32(Elem<T>::from_f32(v)); }

// HF rotate_half RoPE on a 128-wide head held 4 dims per lane: dims < 64 pair with dims + 64 (lane ^ 16)
template <typename T> __device__ __forceinline__ void rope4(float (&x)[4], const float (&c)[4], const float (&s)[4], int lane)
{
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float other = __shfl_xor_sync(0xffffffffu, x[i], 16);
        const float a = rnd<T>(x[i] * c[i]), b = rnd<T>(other * s[i]);
        x[i] = rnd<T>(lane < 16 ? a - b : a + b);  // x1*c - x2*s  |  x2*c + x1*s
    }
}

constexpr int kAttnBatch = 8;  // cached rows a warp has in flight (k and v: 2 x 8 bytes per lane each)

// EARLY (Q4_ATTN_EARLY_CACHE): under programmatic dependent launch only q/k/v of the NEW token come from the preceding kernel;
// `pos`, the cos / sin rows and the cache rows below `pos` were written by earlier steps, so they are fetched BEFORE
// griddepcontrol.wait, while the q/k/v GEMV is still running -- the launch then costs one dependent memory round trip after the
// GEMV instead of 2 + ceil(pos / 8).
template <typename T, bool EARLY>
__global__ void __launch_bounds__(kAttnWarps * 32)
decode_attention_kernel(const DecodeAttnArgs a)
{
    __shared__ float s_m[kAttnWarps], s_l[kAttnWarps];
    __shared__ float s_acc[kAttnWarps][kHd];
    const int h = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int group = a.nh / a.nkv, kv = h / group;
    if (!EARLY) pdl_wait();  // qkv comes from the preceding GEMV
    pdl_launch_dependents();
    const long long pos = *a.pos;
    if (pos < 0 || pos >= a.max_len) __trap();  // the cache row / cos-sin row of `pos` would be out of bounds: fail loudly, never overwrite a neighbour
    const T* qkv = reinterpret_cast<const T*>(a.qkv);
    const int d0 = lane * 4;
    float q[4], kn[4], vn[4], c[4], s[4];
    {
        const int dc = d0 & 63;  // cos/sin index: dim mod 64
        load4(reinterpret_cast<const T*>(a.cos_tab) + pos * (kHd / 2) + dc, c);
        load4(reinterpret_cast<const T*>(a.sin_tab) + pos * (kHd / 2) + dc, s);
    }
    T* kc = reinterpret_cast<T*>(a.k_cache) + (size_t)kv * a.max_len * kHd;
    T* vc = reinterpret_cast<T*>(a.v_cache) + (size_t)kv * a.max_len * kHd;
    // cached positions are dealt to the warps round-robin and fetched kAttnBatch at a time: all loads of a batch are in flight
    // together (one memory round trip per batch, not per position)
    uint2 kr[kAttnBatch], vr[kAttnBatch];
    auto fetch = [&](long long base) {
#pragma unroll
        for (int b = 0; b < kAttnBatch; b++) {
            const long long j = base + warp + (long long)b * kAttnWarps;
            if (j < pos) {
                kr[b] = *reinterpret_cast<const uint2*>(kc + j * kHd + d0);
                vr[b] = *reinterpret_cast<const uint2*>(vc + j * kHd + d0);
            }
        }
    };
    fetch(0);
    if (EARLY) pdl_wait();
    load4(qkv + (size_t)h * kHd + d0, q);
    load4(qkv + (size_t)a.nh * kHd + (size_t)kv * kHd + d0, kn);
    load4(qkv + (size_t)(a.nh + a.nkv) * kHd + (size_t)kv * kHd + d0, vn);
    rope4<T>(q, c, s, lane);
    rope4<T>(kn, c, s, lane);
    if (h % group == 0 && warp == 0) {  // append the new token's (rotated) k and v
        uint2 pk, pv;
        pk.x = pack2<T>(kn[0], kn[1]); pk.y = pack2<T>(kn[2], kn[3]);
        pv.x = pack2<T>(vn[0], vn[1]); pv.y = pack2<T>(vn[2], vn[3]);
        *reinterpret_cast<uint2*>(kc + pos * kHd + d0) = pk;
        *reinterpret_cast<uint2*>(vc + pos * kHd + d0) = pv;
    }
    float m = -INFINITY, l = 0.0f, acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    auto visit = [&](const float (&kk)[4], const float (&vv)[4]) {
        float dot = q[0] * kk[0] + q[1] * kk[1] + q[2] * kk[2] + q[3] * kk[3];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        const float sc = dot * a.scale;
        const float mn = fmaxf(m, sc);
        const float corr = __expf(m - mn), p = __expf(sc - mn);
        l = l * corr + p;
#pragma unroll
        for (int i = 0; i < 4; i++) acc[i] = acc[i] * corr + p * vv[i];
        m = mn;
    };
    for (long long base = 0; base < pos; base += kAttnBatch * kAttnWarps) {
        if (base) fetch(base);
#pragma unroll
        for (int b = 0; b < kAttnBatch; b++) {
            if (base + warp + (long long)b * kAttnWarps < pos) {  // warp-uniform; same visiting order as one position at a time
                const float2 k0 = unpack2<T>(kr[b].x), k1 = unpack2<T>(kr[b].y), v0 = unpack2<T>(vr[b].x), v1 = unpack2<T>(vr[b].y);
                const float kk[4] = {k0.x, k0.y, k1.x, k1.y}, vv[4] = {v0.x, v0.y, v1.x, v1.y};
                visit(kk, vv);
            }
        }
    }
    if (warp == (int)(pos % kAttnWarps)) visit(kn, vn);  // the new token itself, from registers
    if (lane == 0) {
        s_m[warp] = m;
        s_l[warp] = l;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) s_acc[warp][d0 + i] = acc[i];
    __syncthreads();
    if (threadIdx.x < kHd) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < kAttnWarps; w++) M = fmaxf(M, s_m[w]);
        float L = 0.0f, o = 0.0f;
#pragma unroll
        for (int w = 0; w < kAttnWarps; w++) {
            const float f = s_m[w] == -INFINITY ? 0.0f : __expf(s_m[w] - M);
            L += s_l[w] * f;
            o += s_acc[w][threadIdx.x] * f;
        }
        reinterpret_cast<T*>(a.out)[(size_t)h * kHd + threadIdx.x] = Elem<T>::from_f32(o / L);
    }
}

int decode_attention(const void* qkv, const void* cos_tab, const void* sin_tab, void* k_cache, void* v_cache, const long long* pos,
                     void* out, int nh, int nkv, int hd, int max_len, int dtype, int flags, cudaStream_t stream)
{
    if (!qkv || !cos_tab || !sin_tab || !k_cache || !v_cache || !pos || !out) return Q4_ERR_NULL;
    if (hd != kHd || nh < 1 || nkv < 1 || nh % nkv || max_len < 1) return Q4_ERR_SHAPE;
    if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(cos_tab) | reinterpret_cast<uintptr_t>(sin_tab) |
         reinterpret_cast<uintptr_t>(k_cache) | reinterpret_cast<uintptr_t>(v_cache)) & 7)
        return Q4_ERR_ALIGN;
    DecodeAttnArgs a = {qkv, cos_tab, sin_tab, k_cache, v_cache, pos, out, nh, nkv, max_len, 1.0f / sqrtf((float)hd)};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nh);
    cfg.blockDim = dim3(kAttnWarps * 32);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (flags & Q4_GEMV_PDL) ? 1 : 0;
    cudaError_t e;
    const bool early = (flags & Q4_GEMV_PDL) && (flags & Q4_ATTN_EARLY_CACHE);
    switch (dtype) {
        case Q4_F16: e = early ? cudaLaunchKernelEx(&cfg, decode_attention_kernel<__half, true>, a)
                               : cudaLaunchKernelEx(&cfg, decode_attention_kernel<__half, false>, a); break;
        case Q4_BF16: e = early ? cudaLaunchKernelEx(&cfg, decode_attention_kernel<__nv_bfloat16, true>, a)
                                : cudaLaunchKernelEx(&cfg, decode_attention_kernel<__nv_bfloat16, false>, a); break;
        default: return Q4_ERR_DTYPE;
    }
    if (e != cudaSuccess) return (int)e;
    return finish_launch();
}

// ---------------------------------------------------------------------------------------------- greedy sampling glue
//
// argmax over the logits of a decode step in one short launch (torch's argmax over 128 256 bf16 values is a 42-us single-wave
// reduction; this is ~3 us).  A value and its index travel as one 64-bit key (order-preserving float bits, inverted index) so
// that max(key) = the largest value at the lowest index (NaN ranks above +inf, as in torch).  Blocks leave their best key in
// the workspace; the last block to finish reduces those and resets the counter, so the launch is replayable in a CUDA graph.
constexpr int kArgmaxThreads = 256, kArgmaxMaxBlocks = 256;

template <typename T> __device__ __forceinline__ unsigned long long argmax_key(T v, unsigned idx)
{
    const uint32_t u = __float_as_uint(Elem<T>::to_f32(v));
    const uint32_t o = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)o << 32) | (0xFFFFFFFFu - idx);
}

template <typename T>
__global__ void __launch_bounds__(kArgmaxThreads)
argmax_kernel(const T* __restrict__ x, long long n, long long* __restrict__ out, unsigned long long* __restrict__ ws)
{
    __shared__ unsigned long long s_best[kArgmaxThreads / 32];
    __shared__ bool s_last;
    unsigned long long best = 0;
    constexpr int V = 16 / sizeof(T);
    const long long nvec = n / V;
    for (long long i = (long long)blockIdx.x * kArgmaxThreads + threadIdx.x; i < nvec; i += (long long)gridDim.x * kArgmaxThreads) {
        const uint4 raw = *reinterpret_cast<const uint4*>(x + i * V);
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int j = 0; j < V; j++) {
            const unsigned long long k = argmax_key<T>(e[j], (unsigned)(i * V + j));
            best = k > best ? k : best;
        }
    }
    if (blockIdx.x == 0)
        for (long long i = nvec * V + threadIdx.x; i < n; i += kArgmaxThreads) {
            const unsigned long long k = argmax_key<T>(x[i], (unsigned)i);
            best = k > best ? k : best;
        }
    auto reduce_block = [&](unsigned long long v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
            v = w > v ? w : v;
        }
        if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = v;
        __syncthreads();
        v = s_best[0];
#pragma unroll
        for (int w = 1; w < kArgmaxThreads / 32; w++) v = s_best[w] > v ? s_best[w] : v;
        __syncthreads();
        return v;
    };
    best = reduce_block(best);
    if (threadIdx.x == 0) {
        ws[1 + blockIdx.x] = best;
        __threadfence();
        s_last = atomicAdd(reinterpret_cast<unsigned*>(ws), 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    unsigned long long v = 0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kArgmaxThreads) {
        const unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(ws + 1 + b);
        v = k > v ? k : v;
    }
    v = reduce_block(v);
    if (threadIdx.x == 0) {
        out[0] = (long long)(0xFFFFFFFFu - (uint32_t)(v & 0xFFFFFFFFu));
        *reinterpret_cast<unsigned*>(ws) = 0;  // ready for the next launch / graph replay
    }
}

int argmax(const void* x, int64_t n, int dtype, long long* out, void* workspace, cudaStream_t stream)
{
    if (!x || !out || !workspace) return Q4_ERR_NULL;
    if (n < 1 || n >= (1ll << 32)) return Q4_ERR_SHAPE;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 7)) return Q4_ERR_ALIGN;
    const int64_t want = (n / 8 + kArgmaxThreads - 1) / kArgmaxThreads;
    const int grid = (int)(want < 1 ? 1 : (want > kArgmaxMaxBlocks ? kArgmaxMaxBlocks : want));
    unsigned long long* ws = reinterpret_cast<unsigned long long*>(workspace);
    switch (dtype) {
        case Q4_F32: argmax_kernel<float><<<grid, kArgmaxThreads, 0, stream>>>((const float*)x, n, out, ws); break;
        case Q4_F16: argmax_kernel<__half><<<grid, kArgmaxThreads, 0, stream>>>((const __half*)x, n, out, ws); break;
        case Q4_BF16: argmax_kernel<__nv_bfloat16><<<grid, kArgmaxThreads, 0, stream>>>((const __nv_bfloat16*)x, n, out, ws); break;
        default: return Q4_ERR_DTYPE;
    }
    return finish_launch();
}

}  // namespace q4
