// pythonInterface.cpp -- the thin extern "C" layer of libquantizations_b200.so (declared in include/quantizations_b200.h).
//
// Same role and file name as the reference's pythonInterface.cpp, but a plain C ABI instead of a CPython module:
// the Python host side (quantizations_b200/_lib.py) binds it with ctypes, and any other host (C, C++, cgo, JNI ...)
// can bind the same symbols.  Section 1 keeps the reference's five entry points by name and argument order
// (reference pythonInterface.cpp:34-46); section 2 is the dtype-tagged / stream-aware / fused-statistics API.
// No arithmetic lives here: argument checks are in the launchers (q4_*.cu).
#include <atomic>
#include <cuda_runtime.h>

#include "../../include/quantizations_b200.h"
#include "q4_launch.h"

namespace q4 {

static std::atomic<int64_t> g_launches{0};

int finish_launch()
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();  // clears a non-sticky launch error so it is reported once, not by every later call
}

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

extern unsigned long long* g_gemv_trace;  // q4_gemv.cu

}  // namespace q4

extern "C" {

// ------------------------------------------------------------------------------------------------ 1. reference names

int cgemm_4bit_inference_naive_fp32(int m, int n, int k, float* A, unsigned char* B, float* absmax, float* datatype,
                                    float* out, int lda, int ldb, int ldc, int blocksize)
{
    (void)lda; (void)ldc;  // the reference ignores them too (kernels.cu:1062 uses only ldb)
    if (n != 1) return Q4_ERR_SHAPE;
    if (ldb != (k + 1) / 2) return Q4_ERR_SHAPE;  // packed rows are contiguous (core.py:482)
    q4_absmax_t st = {absmax, nullptr, nullptr, nullptr, nullptr, 0};
    return q4::gemv_4bit(A, B, &st, datatype, nullptr, out, m, k, blocksize, Q4_F32, Q4_GEMV_EXACT_F32, nullptr, 0, nullptr);
}

int cquantize_blockwise_fp16_fp4(float* code, void* A, float* absmax, unsigned char* out, int blocksize, const int n)
{
    (void)code;
    return q4::quantize_4bit(A, absmax, out, blocksize, n, Q4_FP4, Q4_F16, nullptr);
}

int cdequantize_blockwise_fp16_fp4(float* code, unsigned char* A, float* absmax, void* out, int blocksize, const int n)
{
    (void)code;
    q4_absmax_t st = {absmax, nullptr, nullptr, nullptr, nullptr, 0};
    return q4::dequantize_4bit(A, &st, out, blocksize, n, Q4_FP4, Q4_F16, nullptr);
}

int cquantize_blockwise_fp32(float* code, float* A, float* absmax, unsigned char* out, int blocksize, const int n)
{
    return q4::quantize_8bit(code, A, absmax, out, blocksize, n, nullptr);
}

int cdequantize_blockwise_fp32(float* code, unsigned char* A, float* absmax, float* out, int blocksize, const int n)
{
    return q4::dequantize_8bit(code, A, absmax, out, blocksize, n, nullptr);
}

// ------------------------------------------------------------------------------------------------ 2. generalised API

int q4_quantize_blockwise_4bit(const void* A, float* absmax, uint8_t* out, int blocksize, int64_t n, int quant_type,
                               int in_dtype, void* stream)
{
    return q4::quantize_4bit(A, absmax, out, blocksize, n, quant_type, in_dtype, (cudaStream_t)stream);
}

int q4_quantize_blockwise_8bit(const float* code, const float* A, float* absmax, uint8_t* out, int blocksize, int64_t n,
                               void* stream)
{
    return q4::quantize_8bit(code, A, absmax, out, blocksize, n, (cudaStream_t)stream);
}

int q4_dequantize_blockwise_8bit(const float* code, const uint8_t* A, const float* absmax, float* out, int blocksize,
                                 int64_t n, void* stream)
{
    return q4::dequantize_8bit(code, A, absmax, out, blocksize, n, (cudaStream_t)stream);
}

int q4_dequantize_blockwise_4bit(const uint8_t* A, const q4_absmax_t* stats, void* out, int blocksize, int64_t n,
                                 int quant_type, int out_dtype, void* stream)
{
    return q4::dequantize_4bit(A, stats, out, blocksize, n, quant_type, out_dtype, (cudaStream_t)stream);
}

int q4_gemv_4bit(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias, void* out,
                 int64_t N, int64_t K, int blocksize, int dtype, int flags, const void* prefetch, int64_t prefetch_bytes,
                 void* stream)
{
    return q4::gemv_4bit(x, B, stats, code, bias, out, N, K, blocksize, dtype, flags, prefetch, prefetch_bytes,
                         (cudaStream_t)stream);
}

int q4_gemv_4bit_grouped(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* const* offsets,
                         const int* row_end, int nmat, const float* code, const void* bias, void* out, int64_t rows, int64_t K,
                         int blocksize, int dtype, int flags, const void* prefetch, int64_t prefetch_bytes, void* stream)
{
    return q4::gemv_4bit_grouped(x, B, stats, offsets, row_end, nmat, code, bias, out, rows, K, blocksize, dtype, flags, prefetch,
                                 prefetch_bytes, (cudaStream_t)stream);
}

int q4_gemv_4bit_fused(const q4_gemv_fused_t* args, void* stream) { return q4::gemv_4bit_fused(args, (cudaStream_t)stream); }

int q4_gemv_4bit_chain(const q4_gemv_fused_t* stages, int n, void* barrier_ws, void* stream)
{
    return q4::gemv_4bit_chain(stages, n, barrier_ws, (cudaStream_t)stream);
}

int q4_gemv_4bit_ring(const q4_gemv_fused_t* stages, int n, void* workspace, int64_t workspace_bytes, void* stream)
{
    return q4::gemv_4bit_ring(stages, n, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int q4_gemv_4bit_batch(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias, void* out,
                       int tokens, int64_t N, int64_t K, int blocksize, int dtype, int flags, const void* lut, void* workspace,
                       int64_t workspace_bytes, void* stream)
{
    return q4::gemv_4bit_batch(x, B, stats, code, bias, out, tokens, N, K, blocksize, dtype, flags, lut, workspace, workspace_bytes,
                               (cudaStream_t)stream);
}

int q4_decode_attention(const void* qkv, const void* cos_tab, const void* sin_tab, void* k_cache, void* v_cache, const int64_t* pos,
                        void* out, int nh, int nkv, int hd, int max_len, int dtype, int flags, void* stream)
{
    return q4::decode_attention(qkv, cos_tab, sin_tab, k_cache, v_cache, (const long long*)pos, out, nh, nkv, hd, max_len, dtype, flags,
                                (cudaStream_t)stream);
}

int q4_argmax(const void* x, int64_t n, int dtype, int64_t* out, void* workspace, void* stream)
{
    return q4::argmax(x, n, dtype, (long long*)out, workspace, (cudaStream_t)stream);
}

int q4_gemv_lut_build(const float* code, const float* code2, int dtype, void* lut, void* stream)
{
    return q4::gemv_lut_build(code, code2, dtype, lut, (cudaStream_t)stream);
}

int q4_gemm_4bit(const void* X, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias, void* out,
                 int64_t M, int64_t N, int64_t K, int blocksize, int dtype, void* stream)
{
    return q4::gemm_4bit(X, B, stats, code, bias, out, M, N, K, blocksize, dtype, (cudaStream_t)stream);
}

// developer hook (not declared in the public header): per-CTA phase timestamps of the GEMV kernel, see tools/
void q4_debug_set_gemv_trace(unsigned long long* p) { q4::g_gemv_trace = p; }

// ------------------------------------------------------------------------------------------------ 3. introspection

int q4_abi_version(void) { return Q4_ABI_VERSION; }

const char* q4_error_string(int code)
{
    switch (code) {
        case 0: return "success";
        case Q4_ERR_BLOCKSIZE: return "blocksize must be one of 64,128,256,512,1024,2048,4096";
        case Q4_ERR_DTYPE: return "unsupported element type";
        case Q4_ERR_QUANT_TYPE: return "unsupported quantisation type";
        case Q4_ERR_SHAPE: return "invalid shape";
        case Q4_ERR_NULL: return "required pointer is NULL";
        case Q4_ERR_ALIGN: return "pointer is not sufficiently aligned";
        case Q4_ERR_DEVICE: return "device is not sm_100";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

int64_t q4_launch_count(void) { return q4::g_launches.load(std::memory_order_relaxed); }

int q4_device_info(int* sms, int* cc_major, int* cc_minor)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return (int)e;
    if (sms) *sms = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return 0;
}

}  // extern "C"
