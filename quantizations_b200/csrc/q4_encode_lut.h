// q4_encode_lut.h -- binned form of the 4-bit encoders (plain C, shared by the CUDA kernels and the exhaustive host check
// tests/encode_lut_check.c).
//
// The reference encodes a normalised value x = v * (1 / absmax) with a tree of strict '>' compares against literal
// thresholds (FP4: csrc/kernels.cu:113-163, seven thresholds on |x|; NF4: upstream bitsandbytes' dQuantizeNF4, the 15
// midpoints of the table at kernels.cu:851).  Both are "number of thresholds strictly below x".  One fused multiply-add
// with a 2^23 magic constant rounds x * S to an integer bin exactly (single rounding, ties to even), and no closed bin
// [(b - 1/2) / S, (b + 1/2) / S] holds more than one threshold, so
//
//     code(x) = base[bin] + (x > thr[bin])          base[bin] = thresholds strictly below the bin, thr[bin] = the one inside
//
// is the same function as the compare tree for every x the fast path is used on (|x| <= 1 + 2^-23, or NaN): one shared-
// memory lookup and one compare instead of 7 / 15 compares.  tests/encode_lut_check.c sweeps all 2^32 bit patterns.
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define Q4_ENC_HD __host__ __device__ __forceinline__
#else
#define Q4_ENC_HD static inline
#endif

#define Q4_ENC_BINS 257  /* bins 0..256 */

// ascending thresholds; FP4: the reference's literals (kernels.cu:141-159; 0.583333f is the literal, not 7/12)
Q4_ENC_HD float q4_enc_threshold(int nf4, int i)
{
    const float fp4[7] = {0.00260417f, 0.0859375f, 0.20833333f, 0.29166667f, 0.4166667f, 0.583333f, 0.8333333f};
    const float nf4t[15] = {-0.8480964004993439f, -0.6106329262256622f, -0.4599952697753906f, -0.33967943489551544f,
                            -0.23460740596055984f, -0.13791173323988914f, -0.045525018125772476f, 0.03979014977812767f,
                            0.1202552504837513f, 0.2035212516784668f, 0.2920137718319893f, 0.3893125355243683f,
                            0.5016634166240692f, 0.6427869200706482f, 0.8614784181118011f};
    return nf4 ? nf4t[i] : fp4[i];
}

// FP4: the code of rank r (thresholds below |x|) is nibble r of 0x32547610 (kernels.cu:141-162)
Q4_ENC_HD uint32_t q4_enc_fp4_rank_code(int r) { return (0x32547610u >> (4 * r)) & 0xFu; }

// Table entry of bin b.  NF4: bins of x in [-1, 1], width 1/128, bin = rint(128 x) + 128; word = base.
// FP4: bins of |x| in [0, 1], width 1/256, bin = rint(256 |x|); word = code(base) | code(base + 1) << 4.
// Returns the number of thresholds inside the closed bin (must be <= 1: checked by the host test).
Q4_ENC_HD int q4_enc_entry(int nf4, int b, float* thr, uint32_t* word)
{
    const float scale = nf4 ? 128.0f : 256.0f;
    const float centre = nf4 ? (float)(b - 128) : (float)b;
    const float lo = (centre - 0.5f) / scale, hi = (centre + 0.5f) / scale;  // exact (dyadic)
    const int nthr = nf4 ? 15 : 7;
    int base = 0, inside = 0;
    float t_in = INFINITY;
    for (int i = 0; i < nthr; i++) {
        const float t = q4_enc_threshold(nf4, i);
        if (t < lo) base++;
        else if (t <= hi) {
            inside++;
            t_in = t;
        }
    }
    *thr = t_in;
    if (nf4) *word = (uint32_t)base;
    else *word = q4_enc_fp4_rank_code(base) | (q4_enc_fp4_rank_code(base < 7 ? base + 1 : 7) << 4);
    return inside;
}

// bin index of a (clamped) operand: low 9 bits of fma(x, S, 2^23 [+ 128])
Q4_ENC_HD uint32_t q4_enc_bin_bits(int nf4, float xc)
{
    const float r = nf4 ? fmaf(xc, 128.0f, 8388736.0f) : fmaf(xc, 256.0f, 8388608.0f);
    union { float f; uint32_t u; } c;
    c.f = r;
    return c.u;
}
