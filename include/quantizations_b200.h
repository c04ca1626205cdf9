/*
 * include/quantizations_b200.h -- C ABI of libquantizations_b200.so
 *
 * B200-native (sm_100a) 4-bit weight-only Linear engine.  This header is the drop-in boundary for the hot path of
 * kkbwilldo/quantizations (reference at /root/reference): every entry point below replaces one function the
 * reference's FFI binds (its CPython module `kbkim_lib`, pythonInterface.cpp:154-161) or extends it with the
 * dtype / stream arguments the reference hard-codes.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  All pointers are DEVICE pointers unless stated otherwise.
 *  - The caller owns every buffer; nothing is allocated, freed or retained (as in the reference).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which is what the reference uses
 *    for every launch, ops.cu:82-94,125-127,170).  All calls are asynchronous and CUDA-graph capturable.
 *  - Every function returns 0 on success, a cudaError_t value (> 0) if the launch failed, or a negative
 *    Q4_ERR_* code for invalid arguments.  (The reference's functions return void and check nothing; a caller
 *    that ignores the int gets the reference behaviour.)  q4_error_string() decodes both ranges.
 *  - Packed 4-bit layout, absmax layout and the double-quant ("nested") statistics are exactly the reference's
 *    (core.py:536-576): byte i holds element 2i in its HIGH nibble and 2i+1 in its LOW nibble; absmax[b] covers
 *    elements [b*blocksize, (b+1)*blocksize) of the flattened [N, K] weight.
 */
#ifndef QUANTIZATIONS_B200_H
#define QUANTIZATIONS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define Q4_ABI_VERSION 1

/* element types of activations / dense weights */
enum { Q4_F32 = 0, Q4_F16 = 1, Q4_BF16 = 2 };
/* quantisation data types; 0/1 are the reference's DataType_t (ops.cuh:6-10), NF4 is new */
enum { Q4_GENERAL8BIT = 0, Q4_FP4 = 1, Q4_NF4 = 2 };
/* argument errors (negative so they never collide with cudaError_t) */
enum {
    Q4_ERR_BLOCKSIZE = -1, /* blocksize not in {64,128,256,512,1024,2048,4096} (core.py:350,408,549,603) */
    Q4_ERR_DTYPE = -2,     /* unknown element type */
    Q4_ERR_QUANT_TYPE = -3, /* unknown quant type (core.py:533,608 raise NotImplementedError) */
    Q4_ERR_SHAPE = -4,     /* negative size, odd K, n != 1 ... */
    Q4_ERR_NULL = -5,      /* required pointer is NULL */
    Q4_ERR_ALIGN = -6,     /* pointer alignment insufficient for the vectorised kernels */
    Q4_ERR_DEVICE = -7     /* not an sm_100 device */
};

/* ------------------------------------------------------------------------------------------------------------
 * 1. The reference's five entry points, same names, argument order and meaning.
 *    Replaces: pythonInterface.cpp:34-46 (c* shims) -> :15-27 -> ops.cu launchers.
 *    Differences: return int instead of void; launch on the legacy default stream like the reference.
 * ---------------------------------------------------------------------------------------------------------- */

/* out[m] = A[1,k] . dequant(B[m,k])^T, fp32 activations/outputs, `absmax` already fp32 per block, `datatype` the
 * 16-entry code table.  n must be 1.  Replaces pythonInterface.cpp:34 / ops.cu:167-171 / kernels.cu:1062. */
int cgemm_4bit_inference_naive_fp32(int m, int n, int k, float* A, unsigned char* B, float* absmax, float* datatype,
                                    float* out, int lda, int ldb, int ldc, int blocksize);
/* fp16 -> FP4 blockwise quantise.  `code` is ignored (NULL in the reference, core.py:553).
 * Replaces pythonInterface.cpp:37 / ops.cu:53-95 / kernels.cu:340 <half,BS,*,0,FP4>. */
int cquantize_blockwise_fp16_fp4(float* code, void* A, float* absmax, unsigned char* out, int blocksize, const int n);
/* FP4 -> fp16 blockwise dequantise with fp32 absmax.  Replaces pythonInterface.cpp:40 / ops.cu:97-128 / kernels.cu:480. */
int cdequantize_blockwise_fp16_fp4(float* code, unsigned char* A, float* absmax, void* out, int blocksize, const int n);
/* fp32 -> 8-bit codebook blockwise quantise (`code` = 256 sorted floats).  Replaces pythonInterface.cpp:43. */
int cquantize_blockwise_fp32(float* code, float* A, float* absmax, unsigned char* out, int blocksize, const int n);
/* 8-bit codebook -> fp32 blockwise dequantise.  Replaces pythonInterface.cpp:46. */
int cdequantize_blockwise_fp32(float* code, unsigned char* A, float* absmax, float* out, int blocksize, const int n);

/* ------------------------------------------------------------------------------------------------------------
 * 2. Generalised entry points (dtype-tagged, explicit stream, fused double-quant).
 * ---------------------------------------------------------------------------------------------------------- */

/* Double-quant ("nested") statistics of one quantised tensor, as the reference's QuantState holds them
 * (core.py:23-88, :563-576).  If qabsmax == NULL the tensor is not nested and `absmax` (fp32 per block) is used.
 * Otherwise absmax[b] = code2[qabsmax[b]] * absmax2[b / blocksize2] + *offset, evaluated as one fp32 multiply then
 * one fp32 add (kernels.cu:552 + core.py:468), inside the consuming kernel. */
typedef struct q4_absmax_t {
    const float* absmax;      /* fp32 [nblocks]              (non-nested) */
    const uint8_t* qabsmax;   /* uint8 [nblocks]             (nested)     */
    const float* code2;       /* fp32 [256] dynamic map      (nested)     */
    const float* absmax2;     /* fp32 [ceil(nblocks/blocksize2)]          */
    const float* offset;      /* fp32 [1], device memory     (nested)     */
    int blocksize2;           /* 256 in the reference (core.py:565)       */
} q4_absmax_t;

/* Blockwise 4-bit quantise: A (in_dtype, n elements, contiguous) -> out uint8[(n+1)/2], absmax fp32[ceil(n/bs)].
 * Bit-exact with kernels.cu:401-476 for Q4_FP4 (all three input types).  Replaces ops.cu:53-95 <T,0,FP4>. */
int q4_quantize_blockwise_4bit(const void* A, float* absmax, uint8_t* out, int blocksize, int64_t n, int quant_type,
                               int in_dtype, void* stream);
/* Blockwise 8-bit codebook quantise of fp32 data (used for the double-quant of absmax, core.py:565).
 * Bit-exact with kernels.cu:166-238 + :453-461.  Replaces ops.cu:53-95 <float,0,General8bit>. */
int q4_quantize_blockwise_8bit(const float* code, const float* A, float* absmax, uint8_t* out, int blocksize,
                               int64_t n, void* stream);
/* Blockwise 8-bit dequantise to fp32: out[i] = code[A[i]] * absmax[i / blocksize].  Replaces ops.cu:127. */
int q4_dequantize_blockwise_8bit(const float* code, const uint8_t* A, const float* absmax, float* out, int blocksize,
                                 int64_t n, void* stream);
/* Blockwise 4-bit dequantise, double-quant decode fused: A uint8[(n+1)/2] -> out (out_dtype, n elements).
 * Bit-exact with kernels.cu:528-567 (+ core.py:613-617 when nested).  Replaces ops.cu:121-125 + 2 torch launches. */
int q4_dequantize_blockwise_4bit(const uint8_t* A, const q4_absmax_t* stats, void* out, int blocksize, int64_t n,
                                 int quant_type, int out_dtype, void* stream);

/* Batch-1 decode GEMV with the double-quant decode fused:
 *     out[r] = sum_k x[k] * code[nib(B[r,k])] * absmax[(r*K + k) / blocksize]   (+ bias[r]),  r in [0, N)
 * x, out, bias are `dtype`; accumulation is fp32.  `code` = 16 fp32 entries (QuantState.code).
 * K must be even.  Replaces core.py:467-499 (3 launches) / ops.cu:167-171 / kernels.cu:1062.
 * flags: Q4_GEMV_EXACT_F32 forces the fp32-multiply path (reference arithmetic for T=float) for any dtype;
 *        Q4_GEMV_PDL launches with programmatic stream serialization (the kernel's prologue -- table build, first
 *        weight loads, L2 prefetch -- overlaps the previous kernel's tail; x is read only after it has completed);
 *        Q4_GEMV_SHARE_SM launches one half-SM CTA per SM instead of two, leaving half of every SM to an independent launch
 *        running concurrently on another stream (or to the next launch's prologue).
 * prefetch / prefetch_bytes (optional, NULL / 0): a byte range that the NEXT call will stream (typically the packed
 * weight of the following Linear4bit).  It is pulled into the 126 MB L2 with TMA bulk prefetches while this call
 * computes, so HBM never idles between dependent launches.  Purely a hint: results do not depend on it. */
enum { Q4_GEMV_DEFAULT = 0, Q4_GEMV_EXACT_F32 = 1, Q4_GEMV_PDL = 2, Q4_GEMV_SHARE_SM = 4, Q4_ATTN_EARLY_CACHE = 8, Q4_GEMV_SWIGLU = 16,
       Q4_GEMV_BATCH_TC5 = 32 /* q4_gemv_4bit_batch: take the tcgen05 kernel instead of the default mma.sync one */ };
int q4_gemv_4bit(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias,
                 void* out, int64_t N, int64_t K, int blocksize, int dtype, int flags, const void* prefetch,
                 int64_t prefetch_bytes, void* stream);

/* Grouped decode GEMV: up to 4 Linear4bit weights that share the activation vector (q/k/v, gate/up) in ONE launch.
 * The matrices must be stored back to back: B = packed bytes of all `rows` = sum N_i rows, `stats->qabsmax` / `stats->absmax2`
 * (or `stats->absmax`) likewise concatenated (each N_i*K/64 a multiple of blocksize2 so nested blocks do not straddle
 * matrices), `out` / `bias` are [rows].  Only the double-quant offset is per matrix: offsets[i] (device pointers, nested only;
 * stats->offset is ignored), row_end[i] = exclusive end row of matrix i (host array, row_end[nmat-1] == rows).
 * Same arithmetic as q4_gemv_4bit per matrix; blocksize 64, K % 64 == 0, fp16/bf16 only. */
int q4_gemv_4bit_grouped(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* const* offsets,
                         const int* row_end, int nmat, const float* code, const void* bias, void* out, int64_t rows, int64_t K,
                         int blocksize, int dtype, int flags, const void* prefetch, int64_t prefetch_bytes, void* stream);

/* Tensor-parallel row-parallel layers (o_proj, down_proj): every rank holds a partial [rows] vector that must be summed
 * over the ranks (one all-reduce of a hidden-sized vector per layer).  With this descriptor the decode GEMV does
 * that exchange itself, in its epilogue, over NVLink peer memory: each CTA stores the fp32 partial sums of ITS rows, tagged with
 * a launch epoch (one 8-byte store per value), into every peer's exchange area, polls its own area until the same rows of every
 * peer carry the epoch, and adds the partials in rank order (deterministic, bit-identical on every rank) before bias / residual
 * and the store -- a one-shot all-reduce with no extra launch, no flag, no fence and no grid-wide barrier.
 * peer_bases: DEVICE array [world] of the ranks' exchange areas (one symmetric allocation of Q4_AR_BYTES(max_rows) per rank,
 * zeroed once, e.g. torch.distributed._symmetric_memory: buffer_ptrs_dev); all ranks must issue the same sequence of calls. */
typedef struct q4_allreduce_t {
    void* const* peer_bases;
    int world, rank; /* world <= 8 */
    int max_rows;    /* rows of the row-parallel layers sharing this exchange area (the hidden size): every call must have rows == max_rows */
} q4_allreduce_t;
#define Q4_AR_MAX_CTAS 1024
#define Q4_AR_DATA_OFFSET 65536
#define Q4_AR_BYTES(max_rows) (Q4_AR_DATA_OFFSET + 2 * 8 * (int64_t)(max_rows) * 8)

/* Decode GEMV with the surrounding elementwise glue of a transformer block fused in (everything optional, NULL = off):
 *   rms_weight : the activation is RMS-normalised first, x * rsqrt(mean(x^2) + rms_eps) * rms_weight  (fp32, rounded to dtype)
 *   x_gate     : the activation is silu(x_gate[k]) * x[k]  (SwiGLU: x = up-projection output, x_gate = gate output)
 *   bias       : added to the output; it may alias `out`, which makes it a residual stream updated in place
 *   nmat > 1   : grouped launch, see q4_gemv_4bit_grouped (offsets / row_end as there; NULL for a single matrix)
 *   flags & Q4_GEMV_SWIGLU (q4_gemv_4bit_fused only): nmat == 2 and the two matrices (gate, up; equally long) are stored
 *                INTERLEAVED in chunks of 4 rows -- rows 8t..8t+3 of B / of the statistics are gate rows 4t..4t+3, rows 8t+4..8t+7
 *                the up rows 4t..4t+3 (4 rows x K/64 blocks must be a multiple of blocksize2, i.e. K % 4096 == 0 for the usual
 *                64 / 256) -- and the launch stores out[j] = silu(gate_j) * up_j for j in [0, rows/2): every CTA owns both halves of
 *                its outputs, so SwiGLU costs nothing and the down projection stages a plain vector.  No bias / allreduce.
 * so a decoder layer's Linear4bit work is four launches: norm+qkv, o+residual, norm+gate/up(+swiglu), (swiglu+)down+residual.
 * fp16/bf16 activations, blocksize 64, K % 64 == 0; with rms_weight K <= 16384. */
typedef struct q4_gemv_fused_t {
    const void* x;
    const void* x_gate;
    const void* rms_weight;
    float rms_eps;
    const uint8_t* B;
    const q4_absmax_t* stats;
    const float* const* offsets;
    const int* row_end;
    int nmat;
    const float* code;
    const void* bias;
    void* out;
    int64_t rows, K;
    int blocksize, dtype, flags;
    const void* prefetch;
    int64_t prefetch_bytes;
    const void* lut; /* optional: table image from q4_gemv_lut_build for this code / stats->code2 / dtype (NULL: built per launch) */
    void* workspace; /* optional: Q4_GEMV_WORKSPACE_BYTES of device memory, zeroed ONCE by the caller, owned by one stream at a time.
                      * With lut and workspace the tcgen05 kernel runs (row tiles split along K across CTAs, partial sums combined
                      * through the workspace in a fixed order; it leaves the workspace zeroed); without, the mma.sync kernel. */
    int64_t workspace_bytes;
    const q4_allreduce_t* allreduce; /* optional: sum the output over tensor-parallel ranks in the epilogue (see q4_allreduce_t) */
    int64_t prefetch_K; /* optional (0 = unknown): in_features of the weight behind `prefetch`.  With it the hint becomes exact: only
                         * the first tiles every CTA of the NEXT launch will ask for (what it keeps in registers before its
                         * activation exists, ~14 MB over the grid) are pulled into L2 instead of the whole range. */
} q4_gemv_fused_t;
int q4_gemv_4bit_fused(const q4_gemv_fused_t* args, void* stream);

/* Small batch (2 <= tokens <= 16: speculative / multi-sequence decode) in ONE pass over the packed weight:
 *     out[t, r] = sum_k x[t, k] * code[nib(B[r,k])] * absmax[(r*K + k) / 64]   (+ bias[r])
 * x [tokens, K], out [tokens, N] row-major contiguous, bias [N].  The reference sends more than one token to a full dequantise +
 * dense GEMM (modules.py:56-64).  Default kernel (q4_gemv_tokens.cu): mma.sync with the 8 B columns as tokens and the 16 A rows as
 * weight rows, each quantisation block = 4 MMAs scaled by its absmax; the cost is that of one GEMV pass for up to 8 tokens, two
 * passes for 9..16.  Needs the table image (`lut`), blocksize 64, K % 256 == 0, N % 16 == 0; `workspace` may be null.  Other
 * shapes, or flags & Q4_GEMV_BATCH_TC5: the tcgen05 decode kernel (needs the split-K workspace, see q4_gemv_fused_t);
 * Q4_ERR_SHAPE if neither covers the call (callers then use q4_gemm_4bit). */
int q4_gemv_4bit_batch(const void* x, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias, void* out,
                       int tokens, int64_t N, int64_t K, int blocksize, int dtype, int flags, const void* lut, void* workspace,
                       int64_t workspace_bytes, void* stream);

/* `n` (<= 4) DEPENDENT decode GEMVs in one persistent launch: stage i+1 may consume what stage i produced (x / x_gate / bias of a
 * later stage pointing at the `out` of an earlier one), e.g. o_proj -> gate/up -> down_proj -> the next layer's q/k/v.  Between
 * stages the grid synchronises on a counter in `barrier_ws` (4 bytes of device memory, zero before the first call, left zero;
 * owned by one stream at a time) instead of ending the kernel: the table image is fetched once, the cold start is paid once, and
 * the first weight tiles of the next stage are already in registers while the barrier is pending.  Same results as n calls of
 * q4_gemv_4bit_fused -- which is also what runs when a stage is not eligible (no table image, ragged shape, mixed types).
 * flags of stage 0 apply to the launch. */
int q4_gemv_4bit_chain(const q4_gemv_fused_t* stages, int n, void* barrier_ws, void* stream);

/* `n` (<= Q4_GEMV_RING_MAX_STAGES) DEPENDENT decode GEMVs in one launch of the persistent ring kernel (csrc/q4_gemv_ring.cuh): one
 * CTA per SM; a producer warp streams every stage's packed weight through a shared-memory ring with TMA and never stops at a stage
 * boundary (the weight does not depend on the activation); rows are split across CTAs at k-tile granularity (inter-CTA split-K,
 * fixed-order combine); between stages there is no grid barrier -- a stage's epilogue stores its outputs as tagged words that the next
 * stage's staging polls directly.  The dependencies must be expressed through the pointers:
 *   - stage i+1's x == stage i's out (the first K values are consumed), or
 *   - stage i is a grouped gate/up pair (nmat == 2, equal halves) and stage i+1 has x_gate == stage i's out and x == out + rows/2:
 *     the gate/up epilogue then publishes silu(gate) * up itself (stage i's out is still written in full);
 *   - a stage's bias may be memory the PREVIOUS kernel wrote, or exactly the out of an earlier stage at most 6 stages back (the
 *     residual stream);
 *   - stage 0's x / x_gate, and any stage's rms_weight, are memory the previous kernel wrote.
 * Row-parallel stages may carry q4_allreduce_t: the sum over the tensor-parallel ranks then runs in the stage's epilogue, before bias
 * and before the outputs are handed to the next stage.
 * Any other aliasing between stages, ragged shapes (K % 256, rows % 32, grouped members not ending on
 * 32-row boundaries), K > 16384, a missing table image or mixed types return Q4_ERR_SHAPE / Q4_ERR_ALIGN WITHOUT launching: the caller
 * falls back to q4_gemv_4bit_chain / q4_gemv_4bit_fused.  Results agree with those within the GEMV tolerance (same per-tile
 * arithmetic, same fixed summation order over k tiles).  `workspace`: Q4_GEMV_RING_WS_BYTES of 16-byte aligned device memory, zeroed
 * ONCE by the caller, owned by one stream at a time (it carries the exchange words and the per-CTA launch epochs, so replayed CUDA
 * graphs stay in step).  flags of stage 0 apply to the launch (Q4_GEMV_PDL).  The grid is one CTA per SM and its CTAs wait for each
 * other: it must not be launched while a kernel that waits for IT holds SMs. */
#define Q4_GEMV_RING_MAX_STAGES 32
#define Q4_GEMV_RING_WS_BYTES (73728 + 8 * 131072)
int q4_gemv_4bit_ring(const q4_gemv_fused_t* stages, int n, void* workspace, int64_t workspace_bytes, void* stream);

/* The decode GEMV decodes one packed BYTE per shared-memory lookup through a 64-KB table derived from the 16-entry 4-bit code
 * and the 256-entry absmax code (reference: `T quant_map[16]` in kernels.cu:1115-1121 and `code[q]` in kernels.cu:552).  The
 * table depends only on (code, code2, dtype), i.e. it is the same for every Linear4bit of a model: build it once into
 * Q4_GEMV_LUT_BYTES of 16-byte aligned device memory and pass it as q4_gemv_fused_t.lut; each launch then fetches it with one
 * TMA bulk copy instead of rebuilding it.  code2 may be NULL (no nested statistics).  dtype: Q4_F16 or Q4_BF16. */
#define Q4_GEMV_LUT_BYTES 65536
#define Q4_GEMV_WORKSPACE_BYTES (8 << 20)
int q4_gemv_lut_build(const float* code, const float* code2, int dtype, void* lut, void* stream);

/* Decode-step glue of the end-to-end harness (not a reference entry point: the reference has no model code): RoPE (HF
 * rotate_half convention) on the new token's q and k, KV-cache append at *pos, and grouped-query attention of that one query
 * over cache positions [0, *pos], in one launch.  qkv = [nh*hd | nkv*hd | nkv*hd] as the grouped q/k/v GEMV leaves it;
 * cos_tab / sin_tab [max_len, hd/2]; caches [nkv, max_len, hd]; out [nh*hd]; hd must be 128; dtype Q4_F16 / Q4_BF16;
 * flags: Q4_GEMV_PDL.  `pos` is a DEVICE scalar so the call can sit in a replayed CUDA graph.
 *        Q4_ATTN_EARLY_CACHE (with Q4_GEMV_PDL): the caller guarantees that only `qkv` is produced by the kernel launched just
 *        before this one -- `pos`, the tables and the cache rows below `pos` were complete before THAT kernel started (true inside a
 *        decode step: they are written by earlier steps) -- so they are read while the preceding kernel is still running. */
int q4_decode_attention(const void* qkv, const void* cos_tab, const void* sin_tab, void* k_cache, void* v_cache, const int64_t* pos,
                        void* out, int nh, int nkv, int hd, int max_len, int dtype, int flags, void* stream);

/* Greedy-sampling glue of the same harness: out[0] = index of the largest of the n values of x (lowest index on ties, NaN ranks
 * highest, like torch.argmax) in one short launch.  workspace: Q4_ARGMAX_WORKSPACE_BYTES of 8-byte aligned device memory, zeroed
 * once by the caller (the kernel leaves it ready for the next launch, so the call can sit in a replayed CUDA graph). */
#define Q4_ARGMAX_WORKSPACE_BYTES 4096
int q4_argmax(const void* x, int64_t n, int dtype, int64_t* out, void* workspace, void* stream);

/* Prefill / batched path with the dequantisation fused into a tcgen05 tensor-core GEMM:
 *     out[m, r] = sum_k X[m, k] * code[nib(B[r,k])] * absmax[(r*K + k) / 64]   (+ bias[r]),   m in [0, M), r in [0, N)
 * X [M, K], out [M, N], bias [N] are `dtype` (Q4_F16 or Q4_BF16), row-major contiguous; accumulation is fp32 (TMEM).
 * blocksize must be 64 and K a multiple of 64.  Replaces modules.py:63-64 (dequantize_4bit + cast + F.linear): the dense
 * weight is never written to memory. */
int q4_gemm_4bit(const void* X, const uint8_t* B, const q4_absmax_t* stats, const float* code, const void* bias, void* out,
                 int64_t M, int64_t N, int64_t K, int blocksize, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * 3. Introspection
 * ---------------------------------------------------------------------------------------------------------- */
int q4_abi_version(void);                 /* == Q4_ABI_VERSION */
const char* q4_error_string(int code);    /* static string; cudaGetErrorString for positive codes */
/* counts kernel launches issued through this library since process start (bench.py's gpu_launches) */
int64_t q4_launch_count(void);
/* fills sm_count / cc_major / cc_minor of the current device; returns 0 or a cudaError_t */
int q4_device_info(int* sm_count, int* cc_major, int* cc_minor);

#ifdef __cplusplus
}
#endif
#endif /* QUANTIZATIONS_B200_H */
