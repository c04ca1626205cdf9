// q4_launch.h -- host-side declarations shared by the kernel translation units and the extern "C" layer.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/quantizations_b200.h"

namespace q4 {

// Called once after every kernel launch: bumps the launch counter (q4_launch_count) and returns the launch status
// (the reference never checks: ops.cu:28,50,82-94,125-127,170).  cudaGetLastError: a non-sticky launch error is reported
// once and cleared; sticky errors stay visible to the caller's own checks anyway.
int finish_launch();

int quantize_4bit(const void* A, float* absmax, uint8_t* out, int blocksize, int64_t n, int quant_type, int in_dtype,
                  cudaStream_t stream);
int quantize_8bit(const float* code, const float* A, float* absmax, uint8_t* out, int blocksize, int64_t n,
                  cudaStream_t stream);
int dequantize_8bit(const float* code, const uint8_t* A, const float* absmax, float* out, int blocksize, int64_t n,
                    cudaStream_t stream);
int dequantize_4bit(const uint8_t* A, const q4_absmax_t* stats, void* out, int blocksize, int64_t n, int quant_type,
                    int out_dtype, cudaStream_t stream);
int gemv_4bit(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias, void* out,
              int64_t N, int64_t K, int blocksize, int dtype, int flags, const void* next, int64_t next_bytes,
              cudaStream_t stream);

int gemv_4bit_grouped(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* const* offsets, const int* row_end,
                      int nmat, const float* code, const void* bias, void* out, int64_t rows, int64_t K, int blocksize, int dtype,
                      int flags, const void* next, int64_t next_bytes, cudaStream_t stream);
int gemv_4bit_fused(const q4_gemv_fused_t* f, cudaStream_t stream);
int gemv_4bit_chain(const q4_gemv_fused_t* stages, int n, void* barrier_ws, cudaStream_t stream);
int gemv_4bit_ring(const q4_gemv_fused_t* stages, int n, void* workspace, int64_t workspace_bytes, cudaStream_t stream);
int gemv_4bit_batch(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias, void* out,
                    int tokens, int64_t N, int64_t K, int blocksize, int dtype, int flags, const void* lut, void* workspace,
                    int64_t workspace_bytes, cudaStream_t stream);
int gemv_4bit_tokens(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias, void* out,
                     int tokens, int64_t N, int64_t K, int blocksize, int dtype, int flags, const void* lut, cudaStream_t stream);
int gemv_lut_build(const float* code, const float* code2, int dtype, void* lut, cudaStream_t stream);
int decode_attention(const void* qkv, const void* cos_tab, const void* sin_tab, void* k_cache, void* v_cache, const long long* pos,
                     void* out, int nh, int nkv, int hd, int max_len, int dtype, int flags, cudaStream_t stream);
int argmax(const void* x, int64_t n, int dtype, long long* out, void* workspace, cudaStream_t stream);
int gemm_4bit(const void* X, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias, void* out,
              int64_t M, int64_t N, int64_t K, int blocksize, int dtype, cudaStream_t stream);

int sm_count();  // cached multiprocessor count of the current device

// Index of the current device, clamped to [0, kMaxDevices): cudaFuncSetAttribute and the dynamic-shared-memory probe apply per
// DEVICE, so "already done" flags are kept per device, not per process (a process may drive several GPUs).
constexpr int kMaxDevices = 64;
inline int device_slot()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev < kMaxDevices ? dev : kMaxDevices - 1;
}

inline int ilog2(int v)
{
    int s = 0;
    while ((1 << s) < v) s++;
    return s;
}

}  // namespace q4
