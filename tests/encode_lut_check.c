/* Exhaustive host check of the binned 4-bit encoders (quantizations_b200/csrc/q4_encode_lut.h) against the compare trees
 * they replace (reference csrc/kernels.cu:113-163 for FP4; the 15 NF4 midpoints, strict '>').  Sweeps all 2^32 float bit
 * patterns; the binned encoder is only used on |x| <= 1 + 2^-23 or NaN (x = v * (1 / absmax) with a normal finite absmax),
 * NF4 additionally on any x below that.  Test infrastructure: built and run by tests/test_host.py.
 *
 *     gcc -O2 -fopenmp tests/encode_lut_check.c -lm -o /tmp/encode_lut_check && /tmp/encode_lut_check [stride]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../quantizations_b200/csrc/q4_encode_lut.h"

static float thr_tab[2][Q4_ENC_BINS];
static uint32_t word_tab[2][Q4_ENC_BINS];

static uint32_t ref_nf4(float x)
{
    uint32_t c = 0;
    for (int i = 0; i < 15; i++) c += x > q4_enc_threshold(1, i);
    return c;
}
static uint32_t ref_fp4(float x)
{
    const uint32_t sign = x < 0.0f ? 8u : 0u;
    const float a = fabsf(x);
    int rank = 0;
    for (int i = 0; i < 7; i++) rank += a > q4_enc_threshold(0, i);
    return q4_enc_fp4_rank_code(rank) + sign;
}
/* CUDA's fmaxf returns the other operand for ANY NaN; glibc's propagates signalling NaNs */
static float gpu_fmaxf(float a, float b) { return a != a ? b : (b != b ? a : fmaxf(a, b)); }

static uint32_t lut_nf4(float x)
{
    const float xc = gpu_fmaxf(x, -1.0f);
    const uint32_t bin = q4_enc_bin_bits(1, xc) & 0x1FFu;
    if (bin >= Q4_ENC_BINS) return 0xFFu;
    return word_tab[1][bin] + (xc > thr_tab[1][bin]);
}
static uint32_t lut_fp4(float x)
{
    const float a = gpu_fmaxf(fabsf(x), 0.0f);
    const uint32_t bin = q4_enc_bin_bits(0, a) & 0x1FFu;
    if (bin >= Q4_ENC_BINS) return 0xFFu;
    const uint32_t codes = word_tab[0][bin];
    return (a > thr_tab[0][bin] ? (codes >> 4) : (codes & 0xFu)) + (x < 0.0f ? 8u : 0u);
}

int main(int argc, char** argv)
{
    const uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
    for (int nf4 = 0; nf4 < 2; nf4++)
        for (int b = 0; b < Q4_ENC_BINS; b++)
            if (q4_enc_entry(nf4, b, &thr_tab[nf4][b], &word_tab[nf4][b]) > 1) {
                printf("FAIL: bin %d of %s holds more than one threshold\n", b, nf4 ? "nf4" : "fp4");
                return 1;
            }
    const float limit = 1.0f + 1.1920929e-7f;
    uint64_t bad = 0, checked = 0;
#pragma omp parallel for reduction(+ : bad, checked) schedule(static)
    for (int64_t i = 0; i < (int64_t)((1ull << 32) / stride); i++) {
        const uint32_t u = (uint32_t)((uint64_t)i * stride);
        float x;
        memcpy(&x, &u, 4);
        const int nan = x != x;
        if (nan || x <= limit) {
            checked++;
            if (ref_nf4(x) != lut_nf4(x)) bad++;
        }
        if (nan || fabsf(x) <= limit) {
            checked++;
            if (ref_fp4(x) != lut_fp4(x)) bad++;
        }
    }
    printf("%s: %llu comparisons, %llu mismatches\n", bad ? "FAIL" : "OK", (unsigned long long)checked, (unsigned long long)bad);
    return bad != 0;
}
