/*
 * oracle/q4_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product
 * (quantizations_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it, and only as the checker.
 *
 * Plain-C CPU restatement of the reference's 4-bit Linear hot path (kkbwilldo/quantizations).  Each function
 * cites the reference file:line whose arithmetic it restates.  The structure is deliberately NOT the
 * reference's (no thread blocks, no CUB): it is "one loop per quantization block", stating what each output
 * element is.  All float arithmetic is IEEE binary32 with contraction off (-ffp-contract=off), and fused
 * multiply-adds appear only where stated (fmaf), so the result is independent of the host compiler.
 *
 * Parity pin: tests/golden/ holds outputs of the reference's own kernels (oracle/_ref, built from
 * /root/reference by oracle/Makefile for sm_100a and executed on a B200 by tests/golden/make_golden.py);
 * tests/test_oracle_golden.py checks every function below against them bit-for-bit.
 *   NF4 quantize/dequantize is the exception: the reference has no NF4 enum (ops.cuh:6-10, core.py:533), its
 *   only NF4 artefact is the 16-entry table at kernels.cu:851.  Those two functions restate upstream
 *   bitsandbytes' published dQuantizeNF4 (nearest entry by the 15 midpoints, strict '>') on the FP4 kernels'
 *   block structure -> "NF4 quantize/dequantize: parity unpinned" (the NF4 *GEMV* is pinned: the reference
 *   GEMV takes the code table as an argument, kernels.cu:1119-1120).
 *
 * Element types: 16-bit inputs are passed as float arrays holding the exactly-widened values (half->float and
 * bf16->float are exact), so one entry point serves fp16/bf16/fp32 inputs.  Outputs that the reference rounds
 * to a 16-bit type are rounded here (round-to-nearest-even) and returned widened to float.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

enum { Q4O_F32 = 0, Q4O_F16 = 1, Q4O_BF16 = 2 };
enum { Q4O_FP4 = 1, Q4O_NF4 = 2 };


/* ---------------------------------------------------------------- tiny pthread parallel-for
 * (libgomp is not in this image).  Splits [0, n) into contiguous slices, one per thread; every loop body below
 * writes disjoint outputs, so results do not depend on the thread count.  Q4O_THREADS overrides the count. */
typedef void (*q4o_body)(long lo, long hi, void* ctx);
typedef struct { q4o_body fn; long lo, hi; void* ctx; } q4o_task;
static void* q4o_trampoline(void* p) { q4o_task* t = (q4o_task*)p; t->fn(t->lo, t->hi, t->ctx); return NULL; }

int q4o_num_threads(void)
{
    const char* e = getenv("Q4O_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return (int)n;
}

static void parallel_for(long n, long work_per_index, q4o_body fn, void* ctx)
{
    int nt = q4o_num_threads();
    if (n < 2 || n * work_per_index < 65536 || nt == 1) { fn(0, n, ctx); return; }
    if (nt > n) nt = (int)n;
    pthread_t th[256];
    q4o_task task[256];
    long per = (n + nt - 1) / nt;
    int started = 0;
    for (int i = 0; i < nt; i++) {
        long lo = i * per, hi = lo + per > n ? n : lo + per;
        if (lo >= hi) break;
        task[i] = (q4o_task){fn, lo, hi, ctx};
        if (pthread_create(&th[i], NULL, q4o_trampoline, &task[i]) != 0) { fn(lo, hi, ctx); th[i] = 0; }
        started = i + 1;
    }
    for (int i = 0; i < started; i++) if (th[i]) pthread_join(th[i], NULL);
}

/* ---------------------------------------------------------------- rounding helpers */

static float round_to_f16(float x) { return (float)(_Float16)x; } /* RNE, like __float2half_rn */

static float round_to_bf16(float x) /* RNE, like __float2bfloat16_rn */
{
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) { /* NaN */
        u = 0x7fff0000u;
    } else {
        u += 0x7fffu + ((u >> 16) & 1u);
        u &= 0xffff0000u;
    }
    memcpy(&x, &u, 4);
    return x;
}

static float round_to(float x, int dt)
{
    return dt == Q4O_F16 ? round_to_f16(x) : dt == Q4O_BF16 ? round_to_bf16(x) : x;
}

float q4o_round_to(float x, int dt) { return round_to(x, dt); }

/* ---------------------------------------------------------------- scalar codecs */

/* FP4 quantize of a value already normalised by 1/absmax.  reference: csrc/kernels.cu:113-163.
 * Sign from (x < 0) (so -0.0 and NaN give sign 0); magnitude by strict '>' against the reference's seven
 * float literals -- kept as the literals, 0.583333f is NOT 7/12. */
unsigned char q4o_quantize_fp4(float x)
{
    static const float thr[7] = {0.00260417f, 0.0859375f, 0.20833333f, 0.29166667f, 0.4166667f, 0.583333f, 0.8333333f};
    /* code emitted for "x exceeds exactly the first i thresholds" (kernels.cu:141-162) */
    static const unsigned char code_of_rank[8] = {0x0, 0x1, 0x6, 0x7, 0x4, 0x5, 0x2, 0x3};
    unsigned char sign = x < 0.0f ? 0x8 : 0x0;
    float a = fabsf(x);
    int rank = 0;
    for (int i = 0; i < 7; i++) rank += (a > thr[i]) ? 1 : 0; /* thresholds ascending: rank == position in the tree */
    return (unsigned char)(code_of_rank[rank] + sign);
}

/* FP4 dequantize: magnitude constant * absmax * sign, in that order.  reference: csrc/kernels.cu:70-111.
 * Nibble 0b1000 yields 0*absmax*(-1) = -0.0. */
float q4o_dequantize_fp4(unsigned char v, float absmax)
{
    static const float mag[8] = {0.00000000f, 5.208333333e-03f, 0.66666667f, 1.00000000f,
                                 0.33333333f, 0.50000000f,      0.16666667f, 0.25000000f};
    float sign = (v & 0x8) ? -1.0f : 1.0f;
    float t = mag[v & 0x7] * absmax;
    return t * sign;
}

/* NF4 table: reference csrc/kernels.cu:851 (its only NF4 artefact). */
static const float NF4_TABLE[16] = {-1.0f,
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

const float* q4o_nf4_table(void) { return NF4_TABLE; }

/* NF4 quantize (parity unpinned, see header): index = number of midpoints strictly below x. */
unsigned char q4o_quantize_nf4(float x)
{
    static const float mid[15] = {-0.8480964004993439f,  -0.6106329262256622f,  -0.4599952697753906f,
                                  -0.33967943489551544f, -0.23460740596055984f, -0.13791173323988914f,
                                  -0.045525018125772476f, 0.03979014977812767f, 0.1202552504837513f,
                                  0.2035212516784668f,    0.2920137718319893f,  0.3893125355243683f,
                                  0.5016634166240692f,    0.6427869200706482f,  0.8614784181118011f};
    int idx = 0;
    for (int i = 0; i < 15; i++) idx += (x > mid[i]) ? 1 : 0;
    return (unsigned char)idx;
}

float q4o_dequantize_nf4(unsigned char v, float absmax) { return NF4_TABLE[v & 0xF] * absmax; }

/* 8-bit codebook quantize, deterministic branch (STOCHASTIC == 0).  reference: csrc/kernels.cu:166-238.
 * NOT "nearest code": it is a 7-step bisection from pivot 127 followed by one midpoint test against the
 * neighbour remembered by the bisection, and it differs from nearest-code on exact ties. */
unsigned char q4o_quantize_8bit(const float* code, float x)
{
    int pivot = 127, hi_idx = 255, lo_idx = 0;
    float lo = -1.0f, hi = 1.0f;
    float val = code[pivot];
    for (int step = 64; step > 0; step >>= 1) {
        if (x > val) { lo_idx = pivot; lo = val; pivot += step; }
        else         { hi_idx = pivot; hi = val; pivot -= step; }
        val = code[pivot];
    }
    if (hi_idx == 255) hi = code[255];
    if (lo_idx == 0)   lo = code[0];
    if (x > val) {
        float midpoint = (hi + val) * 0.5f;
        return (unsigned char)(x > midpoint ? hi_idx : pivot);
    } else {
        float midpoint = (lo + val) * 0.5f;
        return (unsigned char)(x < midpoint ? lo_idx : pivot);
    }
}

/* ---------------------------------------------------------------- blockwise quantize */

/* Per-block absmax as the reference computes it.  reference: csrc/kernels.cu:404-424.
 * Every thread starts from -FLT_MAX and folds fmaxf(|v|) (fmaxf drops NaN); out-of-range slots of a partial
 * block are loaded as 0 (kernels.cu:410), so a partial block's absmax is >= 0 while an all-NaN full block
 * yields -FLT_MAX. */
static float block_absmax(const float* a, long valid, int blocksize)
{
    float m = -FLT_MAX;
    /* the device's fmaxf returns the other operand for ANY NaN; glibc's propagates signalling NaNs, so spell it out */
    for (long j = 0; j < valid; j++) {
        const float v = fabsf(a[j]);
        if (v == v) m = fmaxf(m, v);
    }
    if (valid < blocksize) m = fmaxf(m, 0.0f);
    return m;
}

typedef struct {
    const float* A; const float* code; float* absmax; uint8_t* out; int blocksize; long n; int quant_type;
} quant_ctx;

/* 4-bit blockwise quantize.  reference: csrc/kernels.cu:401-476 (FP4 branch :463-471), launcher ops.cu:76-94.
 *   absmax[b] = max|A[b*bs .. )|;  inv = 1.0f/absmax (IEEE divide);  byte t of the block =
 *   q(A[2t]*inv) << 4 | q(A[2t+1]*inv); slots past n are 0; (valid+1)/2 bytes are stored for a partial block.
 * Quirk kept on purpose (kernels.cu:450,465-470): for blocksize >= 1024 each thread packs 4 values into 2
 * bytes but never clears its accumulator, so every ODD byte is OR-ed with the byte before it. */
static void quant4_blocks(long b0, long b1, void* p)
{
    const quant_ctx* c = (const quant_ctx*)p;
    const int bs = c->blocksize;
    const int contaminate = bs >= 1024;
    for (long b = b0; b < b1; b++) {
        long base = b * bs;
        long valid = c->n - base > bs ? bs : c->n - base;
        float m = block_absmax(c->A + base, valid, bs);
        c->absmax[b] = m;
        float inv = 1.0f / m;
        long nbytes = (valid + 1) / 2;
        unsigned char prev = 0;
        for (long t = 0; t < nbytes; t++) {
            float v0 = c->A[base + 2 * t];
            float v1 = (2 * t + 1 < valid) ? c->A[base + 2 * t + 1] : 0.0f;
            unsigned char q0, q1;
            if (c->quant_type == Q4O_NF4) { q0 = q4o_quantize_nf4(v0 * inv); q1 = q4o_quantize_nf4(v1 * inv); }
            else                          { q0 = q4o_quantize_fp4(v0 * inv); q1 = q4o_quantize_fp4(v1 * inv); }
            unsigned char byte = (unsigned char)((q0 << 4) | q1);
            if (contaminate && (t & 1)) byte |= prev;
            prev = byte;
            c->out[base / 2 + t] = byte;
        }
    }
}

void q4o_quantize_blockwise_4bit(const float* A, float* absmax, uint8_t* out, int blocksize, long n, int quant_type)
{
    quant_ctx c = {A, NULL, absmax, out, blocksize, n, quant_type};
    parallel_for((n + blocksize - 1) / blocksize, blocksize, quant4_blocks, &c);
}

/* 8-bit blockwise quantize (the double-quant of absmax uses blocksize 256 on fp32 input).
 * reference: csrc/kernels.cu:396-398,401-461,476; launcher ops.cu:76-94. */
static void quant8_blocks(long b0, long b1, void* p)
{
    const quant_ctx* c = (const quant_ctx*)p;
    const int bs = c->blocksize;
    for (long b = b0; b < b1; b++) {
        long base = b * bs;
        long valid = c->n - base > bs ? bs : c->n - base;
        float m = block_absmax(c->A + base, valid, bs);
        c->absmax[b] = m;
        float inv = 1.0f / m;
        for (long j = 0; j < valid; j++) c->out[base + j] = q4o_quantize_8bit(c->code, c->A[base + j] * inv);
    }
}

void q4o_quantize_blockwise_8bit(const float* code, const float* A, float* absmax, uint8_t* out, int blocksize, long n)
{
    quant_ctx c = {A, code, absmax, out, blocksize, n, 0};
    parallel_for((n + blocksize - 1) / blocksize, 8L * blocksize, quant8_blocks, &c);
}

/* ---------------------------------------------------------------- blockwise dequantize */

typedef struct {
    const uint8_t* A; const float* code; const float* absmax; float* out; int blocksize; long n;
    int quant_type; int out_dtype; float offset;
} dequant_ctx;

/* 8-bit blockwise dequantize: out[i] = code[A[i]] * absmax[i / blocksize] (one fp32 multiply).
 * reference: csrc/kernels.cu:541,549-553; launcher ops.cu:127. */
static void dequant8_range(long lo, long hi, void* p)
{
    const dequant_ctx* c = (const dequant_ctx*)p;
    for (long i = lo; i < hi; i++) c->out[i] = c->code[c->A[i]] * c->absmax[i / c->blocksize];
}

void q4o_dequantize_blockwise_8bit(const float* code, const uint8_t* A, const float* absmax, float* out, int blocksize, long n)
{
    dequant_ctx c = {A, code, absmax, out, blocksize, n, 0, 0, 0.0f};
    parallel_for(n, 1, dequant8_range, &c);
}

/* Double-quant absmax decode as the reference's Python performs it: fp32 multiply in the kernel
 * (kernels.cu:552), then a separate fp32 add of the scalar offset in torch (core.py:467-468, :614-615). */
static void dequant_absmax_range(long lo, long hi, void* p)
{
    const dequant_ctx* c = (const dequant_ctx*)p;
    for (long i = lo; i < hi; i++) {
        float t = c->code[c->A[i]] * c->absmax[i / c->blocksize];
        c->out[i] = t + c->offset;
    }
}

void q4o_dequantize_absmax(const float* code2, const uint8_t* qabsmax, const float* absmax2, float offset,
                           float* out, int blocksize2, long nblocks)
{
    dequant_ctx c = {qabsmax, code2, absmax2, out, blocksize2, nblocks, 0, 0, offset};
    parallel_for(nblocks, 1, dequant_absmax_range, &c);
}

/* 4-bit blockwise dequantize.  reference: csrc/kernels.cu:528-567 (FP4 branch :554-561), launcher ops.cu:121-125.
 *   out[2i] = dq(A[i] >> 4, absmax[2i / bs]);  out[2i+1] = dq(A[i] & 15, absmax[2i / bs]);  float result rounded
 *   (RNE) to the output type on store; an odd n drops the last low nibble. */
static void dequant4_range(long lo, long hi, void* p)
{
    const dequant_ctx* c = (const dequant_ctx*)p;
    for (long i = lo; i < hi; i++) {
        float am = c->absmax[(2 * i) / c->blocksize];
        unsigned char h = c->A[i] >> 4, l = c->A[i] & 0xF;
        float v0 = c->quant_type == Q4O_NF4 ? q4o_dequantize_nf4(h, am) : q4o_dequantize_fp4(h, am);
        float v1 = c->quant_type == Q4O_NF4 ? q4o_dequantize_nf4(l, am) : q4o_dequantize_fp4(l, am);
        c->out[2 * i] = round_to(v0, c->out_dtype);
        if (2 * i + 1 < c->n) c->out[2 * i + 1] = round_to(v1, c->out_dtype);
    }
}

void q4o_dequantize_blockwise_4bit(const uint8_t* A, const float* absmax, float* out, int blocksize, long n,
                                   int quant_type, int out_dtype)
{
    dequant_ctx c = {A, NULL, absmax, out, blocksize, n, quant_type, out_dtype, 0.0f};
    parallel_for((n + 1) / 2, 2, dequant4_range, &c);
}

/* ---------------------------------------------------------------- GEMV */

typedef struct {
    const float* x; const uint8_t* B; const float* absmax; const float* code; float* out; double* out64;
    long N, K, M; int blocksize; int dt; int f32_fused; const float* W;
} gemv_ctx;

/* Batch-1 GEMV, restating the reference kernel's arithmetic AND summation order.
 * reference: csrc/kernels.cu:1126-1218 (launcher ops.cu:167-171; host core.py:477-499).
 *   out[r] = sum_k x[k] * (code[nib(r,k)] * absmax[(r*K + k) / blocksize])
 * Order: lane l of the row's warp walks k = 32*l + 1024*j (j = 0,1,...) 32 values at a time, sequentially,
 * adding into one fp32 accumulator; the 32 lane sums are then combined by a shuffle-down tree (offsets 1,2,4,8,16:
 * cub::WarpReduce::Sum, kernels.cu:1215).  One absmax is fetched per 32-value chunk (:1130-1131).
 *   dt == F32 : w = code*absmax (fp32 multiply), acc = fmaf(x, w, acc)   [nvcc contracts :1206 for T=float;
 *               f32_fused = 0 gives the uncontracted acc + x*w instead -- the golden vectors decide which is pinned]
 *   dt == F16/BF16 : code, absmax, code*absmax and x*w are each rounded to T, acc += (float)(x*w)  (:1120,1131,1169,1206)
 * The final sum is rounded to T on store (:1218).  x and out are passed widened to float.
 * Requires K even (the reference's ldb = (K+1)/2 addressing is only consistent for even K). */
static void gemv_rows(long r0, long r1, void* p)
{
    const gemv_ctx* c = (const gemv_ctx*)p;
    const long K = c->K, ldb = (K + 1) / 2;
    const int dt = c->dt;
    float qmap[16];
    for (int i = 0; i < 16; i++) qmap[i] = round_to(c->code[i], dt);
    for (long r = r0; r < r1; r++) {
        float lane_sum[32];
        for (int l = 0; l < 32; l++) {
            float acc = 0.0f;
            for (long k0 = 32L * l; k0 < K; k0 += 1024) {
                float am = round_to(c->absmax[(2 * ldb * r + k0) / c->blocksize], dt);
                for (int j = 0; j < 32; j++) {
                    long k = k0 + j;
                    unsigned char byte = (k / 2 < K / 2) ? c->B[ldb * r + k / 2] : 0x77; /* pad, kernels.cu:1150 */
                    unsigned char nib = (j & 1) ? (byte & 0xF) : (byte >> 4);
                    float xv = k < K ? c->x[k] : 0.0f;
                    if (dt == Q4O_F32) {
                        float w = qmap[nib] * am;
                        acc = c->f32_fused ? fmaf(xv, w, acc) : acc + xv * w;
                    } else {
                        float w = round_to(qmap[nib] * am, dt);
                        float pr = round_to(xv * w, dt);
                        acc = acc + pr;
                    }
                }
            }
            lane_sum[l] = acc;
        }
        for (int off = 1; off < 32; off <<= 1)
            for (int l = 0; l + off < 32; l++) lane_sum[l] = lane_sum[l] + lane_sum[l + off];
        c->out[r] = round_to(lane_sum[0], dt);
    }
}

void q4o_gemv_4bit(const float* x, const uint8_t* B, const float* absmax, const float* code, float* out,
                   long N, long K, int blocksize, int dt, int f32_fused)
{
    gemv_ctx c = {x, B, absmax, code, out, NULL, N, K, 1, blocksize, dt, f32_fused, NULL};
    parallel_for(N, K, gemv_rows, &c);
}

/* fp64 "truth" for the same contraction (tolerance tests grade both the product and the reference against it). */
static void gemv64_rows(long r0, long r1, void* p)
{
    const gemv_ctx* c = (const gemv_ctx*)p;
    for (long r = r0; r < r1; r++) {
        double acc = 0.0;
        for (long k = 0; k < c->K; k++) {
            long e = r * c->K + k;
            unsigned char byte = c->B[e / 2];
            unsigned char nib = (e & 1) ? (byte & 0xF) : (byte >> 4);
            acc += (double)c->x[k] * ((double)c->code[nib] * (double)c->absmax[e / c->blocksize]);
        }
        c->out64[r] = acc;
    }
}

void q4o_gemv_4bit_f64(const float* x, const uint8_t* B, const float* absmax, const float* code, double* out,
                       long N, long K, int blocksize)
{
    gemv_ctx c = {x, B, absmax, code, NULL, out, N, K, 1, blocksize, 0, 0, NULL};
    parallel_for(N, K, gemv64_rows, &c);
}

/* Prefill truth: out[m, r] = sum_k X[m,k] * W[r,k] in fp64, W = the (already rounded) dequantized weight.
 * reference: modules.py:63-64 (F.linear over dequantize_4bit). */
static void linear64_rows(long r0, long r1, void* p)
{
    const gemv_ctx* c = (const gemv_ctx*)p;
    for (long r = r0; r < r1; r++)
        for (long m = 0; m < c->M; m++) {
            double acc = 0.0;
            for (long k = 0; k < c->K; k++) acc += (double)c->x[m * c->K + k] * (double)c->W[r * c->K + k];
            c->out64[m * c->N + r] = acc;
        }
}

void q4o_linear_f64(const float* X, const float* Wdeq, double* out, long M, long N, long K)
{
    gemv_ctx c = {X, NULL, NULL, NULL, NULL, out, N, K, M, 0, 0, 0, Wdeq};
    parallel_for(N, M * K, linear64_rows, &c);
}
