/*
 * cmsis_min.c - portable restatement of the CMSIS-DSP 1.6.0 float32 functions on the UA3REO
 * firmware's audio/FFT path (see cmsis/arm_math.h for provenance).  TEST INFRASTRUCTURE ONLY.
 * All arithmetic is IEEE-754 binary32 in the upstream operation order; build with
 * -ffp-contract=off so that no multiply-add is fused.
 */
#include "cmsis/arm_math.h"
#include "cmsis/arm_const_structs.h"
#include <stdio.h>
#include <stdlib.h>

/* ---------------------------------------------------------------- tables (arm_common_tables.c) */
static float32_t g_sin_table[FAST_MATH_TABLE_SIZE + 1];
static float32_t g_twiddle_512[1024];
static int g_tables_ready = 0;

static void tables_init(void)
{
    if (g_tables_ready) return;
    /* sinTable_f32: sin(2*pi*k/512), k = 0..512, published as float literals with 8 decimals */
    for (int k = 0; k <= FAST_MATH_TABLE_SIZE; k++) {
        char buf[32];
        snprintf(buf, sizeof buf, "%.8f", sin(2.0 * M_PI * (double)k / 512.0));
        g_sin_table[k] = strtof(buf, NULL);
    }
    /* twiddleCoef_512: {cos(2*pi*i/512), sin(2*pi*i/512)}, published with 9 significant digits */
    for (int i = 0; i < 512; i++) {
        char buf[32];
        snprintf(buf, sizeof buf, "%.9f", cos(2.0 * M_PI * (double)i / 512.0));
        g_twiddle_512[2 * i] = strtof(buf, NULL);
        snprintf(buf, sizeof buf, "%.9f", sin(2.0 * M_PI * (double)i / 512.0));
        g_twiddle_512[2 * i + 1] = strtof(buf, NULL);
    }
    g_tables_ready = 1;
}

const float32_t *ua3_cmsis_sin_table(void) { tables_init(); return g_sin_table; }
const float32_t *ua3_cmsis_twiddle_512(void) { tables_init(); return g_twiddle_512; }

/* fftLen tags the instance; twiddles/bit reversal are resolved inside arm_cfft_f32 */
const arm_cfft_instance_f32 arm_cfft_sR_f32_len512 = {512, NULL, NULL, 0};
const arm_cfft_instance_f32 arm_cfft_sR_f32_len256 = {256, NULL, NULL, 0};

/* ---------------------------------------------------------------- BasicMathFunctions / support */
void arm_scale_f32(float32_t *s, float32_t k, float32_t *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = s[i] * k; }
void arm_add_f32(float32_t *a, float32_t *b, float32_t *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = a[i] + b[i]; }
void arm_sub_f32(float32_t *a, float32_t *b, float32_t *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = a[i] - b[i]; }
void arm_mult_f32(float32_t *a, float32_t *b, float32_t *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = a[i] * b[i]; }
void arm_abs_f32(float32_t *s, float32_t *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = fabsf(s[i]); }
void arm_offset_f32(float32_t *s, float32_t o, float32_t *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = s[i] + o; }
void arm_negate_f32(float32_t *s, float32_t *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = -s[i]; }
void arm_copy_f32(float32_t *s, float32_t *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = s[i]; }
void arm_fill_f32(float32_t v, float32_t *d, uint32_t n) { for (uint32_t i = 0; i < n; i++) d[i] = v; }

/* arm_max_f32: maximum SIGNED value and the first index where it occurs */
void arm_max_f32(float32_t *s, uint32_t n, float32_t *res, uint32_t *idx)
{
    float32_t out = s[0];
    uint32_t oi = 0;
    for (uint32_t i = 1; i < n; i++)
        if (out < s[i]) { out = s[i]; oi = i; }
    *res = out; *idx = oi;
}
void arm_min_f32(float32_t *s, uint32_t n, float32_t *res, uint32_t *idx)
{
    float32_t out = s[0];
    uint32_t oi = 0;
    for (uint32_t i = 1; i < n; i++)
        if (out > s[i]) { out = s[i]; oi = i; }
    *res = out; *idx = oi;
}
void arm_mean_f32(float32_t *s, uint32_t n, float32_t *res)
{
    float32_t sum = 0.0f;
    for (uint32_t i = 0; i < n; i++) sum += s[i];
    *res = sum / (float32_t)n;
}
void arm_power_f32(float32_t *s, uint32_t n, float32_t *res)
{
    float32_t sum = 0.0f;
    for (uint32_t i = 0; i < n; i++) sum += s[i] * s[i];
    *res = sum;
}
void arm_rms_f32(float32_t *s, uint32_t n, float32_t *res)
{
    float32_t sum = 0.0f;
    for (uint32_t i = 0; i < n; i++) sum += s[i] * s[i];
    arm_sqrt_f32(sum / (float32_t)n, res);
}

/* ---------------------------------------------------------------- FastMathFunctions */
float32_t arm_sin_f32(float32_t x)
{
    tables_init();
    float32_t in = x * 0.159154943092f;
    int32_t n = (int32_t)in;
    if (x < 0.0f) n--;
    in = in - (float32_t)n;
    float32_t findex = (float32_t)FAST_MATH_TABLE_SIZE * in;
    uint16_t index = (uint16_t)findex;
    if (index >= FAST_MATH_TABLE_SIZE) { index = 0; findex -= (float32_t)FAST_MATH_TABLE_SIZE; }
    const float32_t fract = findex - (float32_t)index;
    const float32_t a = g_sin_table[index], b = g_sin_table[index + 1];
    return (1.0f - fract) * a + fract * b;
}

float32_t arm_cos_f32(float32_t x)
{
    tables_init();
    float32_t in = x * 0.159154943092f + 0.25f;
    int32_t n = (int32_t)in;
    if (in < 0.0f) n--;
    in = in - (float32_t)n;
    float32_t findex = (float32_t)FAST_MATH_TABLE_SIZE * in;
    uint16_t index = (uint16_t)findex;
    if (index >= FAST_MATH_TABLE_SIZE) { index = 0; findex -= (float32_t)FAST_MATH_TABLE_SIZE; }
    const float32_t fract = findex - (float32_t)index;
    const float32_t a = g_sin_table[index], b = g_sin_table[index + 1];
    return (1.0f - fract) * a + fract * b;
}

/* ---------------------------------------------------------------- IIR lattice */
void arm_iir_lattice_init_f32(arm_iir_lattice_instance_f32 *S, uint16_t numStages, float32_t *pk, float32_t *pv,
                              float32_t *pState, uint32_t blockSize)
{
    S->numStages = numStages;
    S->pkCoeffs = pk;      /* {kN, kN-1, ..., k1} */
    S->pvCoeffs = pv;      /* {vN, vN-1, ..., v0} */
    memset(pState, 0, (numStages + blockSize) * sizeof(float32_t));
    S->pState = pState;
}

void arm_iir_lattice_f32(const arm_iir_lattice_instance_f32 *S, float32_t *pSrc, float32_t *pDst, uint32_t blockSize)
{
    const uint32_t N = S->numStages;
    float32_t *pState = S->pState;
    for (uint32_t n = 0; n < blockSize; n++) {
        float32_t fcur = pSrc[n];                 /* fN(n) = x(n) */
        const float32_t *pk = S->pkCoeffs, *pv = S->pvCoeffs;
        float32_t *px1 = pState, *px2 = pState;
        float32_t acc = 0.0f;
        for (uint32_t i = 0; i < N; i++) {
            const float32_t gcurr = *px1++;                   /* g(i-1)(n-1) */
            const float32_t fnext = fcur - (pk[i] * gcurr);   /* f(i-1)(n) = f(i)(n) - k(i) g(i-1)(n-1) */
            const float32_t gnext = (fnext * pk[i]) + gcurr;  /* g(i)(n) = k(i) f(i-1)(n) + g(i-1)(n-1) */
            acc += gnext * pv[i];
            *px2++ = gnext;
            fcur = fnext;
        }
        acc += fcur * pv[N];                      /* y(n) += g0(n) * v0 */
        *px2++ = fcur;
        pDst[n] = acc;
        pState = pState + 1;                      /* the window slides by one: next sample reads what was written at +1 */
    }
    /* copy the last numStages values to the start of the state buffer */
    memmove(S->pState, S->pState + blockSize, N * sizeof(float32_t));
}

/* ---------------------------------------------------------------- biquads */
void arm_biquad_cascade_df2T_init_f32(arm_biquad_cascade_df2T_instance_f32 *S, uint8_t numStages, float32_t *pCoeffs, float32_t *pState)
{
    S->numStages = numStages; S->pCoeffs = pCoeffs; S->pState = pState;
    memset(pState, 0, 2u * numStages * sizeof(float32_t));
}

void arm_biquad_cascade_df2T_f32(const arm_biquad_cascade_df2T_instance_f32 *S, float32_t *pSrc, float32_t *pDst, uint32_t blockSize)
{
    float32_t *pIn = pSrc, *pOut = pDst, *pState = S->pState, *pCoeffs = S->pCoeffs;
    for (uint32_t stage = 0; stage < S->numStages; stage++) {
        const float32_t b0 = pCoeffs[0], b1 = pCoeffs[1], b2 = pCoeffs[2], a1 = pCoeffs[3], a2 = pCoeffs[4];
        pCoeffs += 5;
        float32_t d1 = pState[0], d2 = pState[1];
        for (uint32_t n = 0; n < blockSize; n++) {
            const float32_t Xn1 = pIn[n];
            const float32_t acc1 = b0 * Xn1 + d1;     /* y[n] = b0 x[n] + d1 */
            d1 = b1 * Xn1 + d2;                       /* d1 = b1 x[n] + a1 y[n] + d2 */
            d1 += a1 * acc1;
            d2 = b2 * Xn1;                            /* d2 = b2 x[n] + a2 y[n] */
            d2 += a2 * acc1;
            pOut[n] = acc1;
        }
        pState[0] = d1; pState[1] = d2;
        pState += 2;
        pIn = pDst;                                   /* later stages run in place on the output */
    }
}

void arm_biquad_cascade_df1_f32(const arm_biquad_casd_df1_inst_f32 *S, float32_t *pSrc, float32_t *pDst, uint32_t blockSize)
{
    float32_t *pIn = pSrc, *pState = S->pState, *pCoeffs = S->pCoeffs;
    for (uint32_t stage = 0; stage < S->numStages; stage++) {
        const float32_t b0 = pCoeffs[0], b1 = pCoeffs[1], b2 = pCoeffs[2], a1 = pCoeffs[3], a2 = pCoeffs[4];
        pCoeffs += 5;
        float32_t Xn1 = pState[0], Xn2 = pState[1], Yn1 = pState[2], Yn2 = pState[3];
        for (uint32_t n = 0; n < blockSize; n++) {
            const float32_t Xn = pIn[n];
            const float32_t acc = (b0 * Xn) + (b1 * Xn1) + (b2 * Xn2) + (a1 * Yn1) + (a2 * Yn2);
            pDst[n] = acc;
            Xn2 = Xn1; Xn1 = Xn; Yn2 = Yn1; Yn1 = acc;
        }
        pState[0] = Xn1; pState[1] = Xn2; pState[2] = Yn1; pState[3] = Yn2;
        pState += 4;
        pIn = pDst;
    }
}

/* ---------------------------------------------------------------- FIR / FIR decimator */
void arm_fir_init_f32(arm_fir_instance_f32 *S, uint16_t numTaps, float32_t *pCoeffs, float32_t *pState, uint32_t blockSize)
{
    S->numTaps = numTaps; S->pCoeffs = pCoeffs; S->pState = pState;
    memset(pState, 0, (numTaps + (blockSize - 1u)) * sizeof(float32_t));
}

void arm_fir_f32(const arm_fir_instance_f32 *S, float32_t *pSrc, float32_t *pDst, uint32_t blockSize)
{
    const uint32_t T = S->numTaps;
    float32_t *pState = S->pState;
    float32_t *pStateCurnt = &S->pState[T - 1u];
    for (uint32_t n = 0; n < blockSize; n++) {
        *pStateCurnt++ = pSrc[n];
        const float32_t *px = pState, *pb = S->pCoeffs;
        float32_t acc = 0.0f;
        for (uint32_t i = 0; i < T; i++) acc += px[i] * pb[i];   /* pCoeffs[i] meets the sample (T-1-i) old */
        pDst[n] = acc;
        pState++;
    }
    memmove(S->pState, S->pState + blockSize, (T - 1u) * sizeof(float32_t));
}

arm_status arm_fir_decimate_init_f32(arm_fir_decimate_instance_f32 *S, uint16_t numTaps, uint8_t M, float32_t *pCoeffs,
                                     float32_t *pState, uint32_t blockSize)
{
    if ((blockSize % M) != 0u) return ARM_MATH_LENGTH_ERROR;
    S->numTaps = numTaps; S->pCoeffs = pCoeffs; S->M = M; S->pState = pState;
    memset(pState, 0, (numTaps + (blockSize - 1u)) * sizeof(float32_t));
    return ARM_MATH_SUCCESS;
}

void arm_fir_decimate_f32(const arm_fir_decimate_instance_f32 *S, float32_t *pSrc, float32_t *pDst, uint32_t blockSize)
{
    const uint32_t T = S->numTaps, M = S->M;
    float32_t *pState = S->pState;
    float32_t *pStateCurnt = S->pState + (T - 1u);
    const uint32_t outBlock = blockSize / M;
    for (uint32_t o = 0; o < outBlock; o++) {
        for (uint32_t i = 0; i < M; i++) *pStateCurnt++ = *pSrc++;
        const float32_t *px = pState, *pb = S->pCoeffs;
        float32_t sum0 = 0.0f;
        for (uint32_t i = 0; i < T; i++) sum0 += px[i] * pb[i];
        pState = pState + M;
        *pDst++ = sum0;
    }
    memmove(S->pState, S->pState + blockSize, (T - 1u) * sizeof(float32_t));
}

/* ---------------------------------------------------------------- normalised LMS */
void arm_lms_norm_init_f32(arm_lms_norm_instance_f32 *S, uint16_t numTaps, float32_t *pCoeffs, float32_t *pState,
                           float32_t mu, uint32_t blockSize)
{
    S->numTaps = numTaps; S->pCoeffs = pCoeffs;
    memset(pState, 0, (numTaps + (blockSize - 1u)) * sizeof(float32_t));
    S->pState = pState; S->mu = mu; S->energy = 0.0f; S->x0 = 0.0f;
}

void arm_lms_norm_f32(arm_lms_norm_instance_f32 *S, float32_t *pSrc, float32_t *pRef, float32_t *pOut, float32_t *pErr,
                      uint32_t blockSize)
{
    const uint32_t T = S->numTaps;
    float32_t *pState = S->pState, *pCoeffs = S->pCoeffs;
    float32_t *pStateCurnt = &S->pState[T - 1u];
    const float32_t mu = S->mu;
    float32_t energy = S->energy, x0 = S->x0;
    for (uint32_t n = 0; n < blockSize; n++) {
        *pStateCurnt++ = pSrc[n];
        const float32_t in = pSrc[n];              /* read before pOut[n] is written: in-place safe */
        energy -= x0 * x0;
        energy += in * in;
        float32_t acc = 0.0f;
        for (uint32_t i = 0; i < T; i++) acc += pState[i] * pCoeffs[i];
        pOut[n] = acc;
        const float32_t e = pRef[n] - acc;
        pErr[n] = e;
        const float32_t w = (e * mu) / (energy + 0.000000119209289f);
        for (uint32_t i = 0; i < T; i++) pCoeffs[i] += w * pState[i];
        x0 = *pState;
        pState = pState + 1;
    }
    S->energy = energy; S->x0 = x0;
    memmove(S->pState, S->pState + blockSize, (T - 1u) * sizeof(float32_t));
}

/* ---------------------------------------------------------------- complex FFT (radix-8 path, len 512) */
static void radix8_butterfly_f32(float32_t *pSrc, uint16_t fftLen, const float32_t *pCoef, uint16_t twidCoefModifier)
{
    uint32_t ia1, ia2, ia3, ia4, ia5, ia6, ia7;
    uint32_t i1, i2, i3, i4, i5, i6, i7, i8;
    uint32_t id, n1, n2, j;
    float32_t r1, r2, r3, r4, r5, r6, r7, r8, t1, t2;
    float32_t s1, s2, s3, s4, s5, s6, s7, s8, p1, p2, p3, p4;
    float32_t co2, co3, co4, co5, co6, co7, co8, si2, si3, si4, si5, si6, si7, si8;
    const float32_t C81 = 0.70710678118f;

    n2 = fftLen;
    do {
        n1 = n2;
        n2 = n2 >> 3;
        i1 = 0;
        do {
            i2 = i1 + n2; i3 = i2 + n2; i4 = i3 + n2; i5 = i4 + n2; i6 = i5 + n2; i7 = i6 + n2; i8 = i7 + n2;
            r1 = pSrc[2 * i1] + pSrc[2 * i5];
            r5 = pSrc[2 * i1] - pSrc[2 * i5];
            r2 = pSrc[2 * i2] + pSrc[2 * i6];
            r6 = pSrc[2 * i2] - pSrc[2 * i6];
            r3 = pSrc[2 * i3] + pSrc[2 * i7];
            r7 = pSrc[2 * i3] - pSrc[2 * i7];
            r4 = pSrc[2 * i4] + pSrc[2 * i8];
            r8 = pSrc[2 * i4] - pSrc[2 * i8];
            t1 = r1 - r3; r1 = r1 + r3; r3 = r2 - r4; r2 = r2 + r4;
            pSrc[2 * i1] = r1 + r2;
            pSrc[2 * i5] = r1 - r2;
            r1 = pSrc[2 * i1 + 1] + pSrc[2 * i5 + 1];
            s5 = pSrc[2 * i1 + 1] - pSrc[2 * i5 + 1];
            r2 = pSrc[2 * i2 + 1] + pSrc[2 * i6 + 1];
            s6 = pSrc[2 * i2 + 1] - pSrc[2 * i6 + 1];
            s3 = pSrc[2 * i3 + 1] + pSrc[2 * i7 + 1];
            s7 = pSrc[2 * i3 + 1] - pSrc[2 * i7 + 1];
            r4 = pSrc[2 * i4 + 1] + pSrc[2 * i8 + 1];
            s8 = pSrc[2 * i4 + 1] - pSrc[2 * i8 + 1];
            t2 = r1 - s3; r1 = r1 + s3; s3 = r2 - r4; r2 = r2 + r4;
            pSrc[2 * i1 + 1] = r1 + r2;
            pSrc[2 * i5 + 1] = r1 - r2;
            pSrc[2 * i3] = t1 + s3;
            pSrc[2 * i7] = t1 - s3;
            pSrc[2 * i3 + 1] = t2 - r3;
            pSrc[2 * i7 + 1] = t2 + r3;
            r1 = (r6 - r8) * C81; r6 = (r6 + r8) * C81;
            r2 = (s6 - s8) * C81; s6 = (s6 + s8) * C81;
            t1 = r5 - r1; r5 = r5 + r1; r8 = r7 - r6; r7 = r7 + r6;
            t2 = s5 - r2; s5 = s5 + r2; s8 = s7 - s6; s7 = s7 + s6;
            pSrc[2 * i2] = r5 + s7;
            pSrc[2 * i8] = r5 - s7;
            pSrc[2 * i6] = t1 + s8;
            pSrc[2 * i4] = t1 - s8;
            pSrc[2 * i2 + 1] = s5 - r7;
            pSrc[2 * i8 + 1] = s5 + r7;
            pSrc[2 * i6 + 1] = t2 - r8;
            pSrc[2 * i4 + 1] = t2 + r8;
            i1 += n1;
        } while (i1 < fftLen);

        if (n2 < 8) break;

        ia1 = 0;
        j = 1;
        do {
            id = ia1 + twidCoefModifier;
            ia1 = id; ia2 = ia1 + id; ia3 = ia2 + id; ia4 = ia3 + id; ia5 = ia4 + id; ia6 = ia5 + id; ia7 = ia6 + id;
            co2 = pCoef[2 * ia1]; co3 = pCoef[2 * ia2]; co4 = pCoef[2 * ia3]; co5 = pCoef[2 * ia4];
            co6 = pCoef[2 * ia5]; co7 = pCoef[2 * ia6]; co8 = pCoef[2 * ia7];
            si2 = pCoef[2 * ia1 + 1]; si3 = pCoef[2 * ia2 + 1]; si4 = pCoef[2 * ia3 + 1]; si5 = pCoef[2 * ia4 + 1];
            si6 = pCoef[2 * ia5 + 1]; si7 = pCoef[2 * ia6 + 1]; si8 = pCoef[2 * ia7 + 1];
            i1 = j;
            do {
                i2 = i1 + n2; i3 = i2 + n2; i4 = i3 + n2; i5 = i4 + n2; i6 = i5 + n2; i7 = i6 + n2; i8 = i7 + n2;
                r1 = pSrc[2 * i1] + pSrc[2 * i5];
                r5 = pSrc[2 * i1] - pSrc[2 * i5];
                r2 = pSrc[2 * i2] + pSrc[2 * i6];
                r6 = pSrc[2 * i2] - pSrc[2 * i6];
                r3 = pSrc[2 * i3] + pSrc[2 * i7];
                r7 = pSrc[2 * i3] - pSrc[2 * i7];
                r4 = pSrc[2 * i4] + pSrc[2 * i8];
                r8 = pSrc[2 * i4] - pSrc[2 * i8];
                t1 = r1 - r3; r1 = r1 + r3; r3 = r2 - r4; r2 = r2 + r4;
                pSrc[2 * i1] = r1 + r2;
                r2 = r1 - r2;
                s1 = pSrc[2 * i1 + 1] + pSrc[2 * i5 + 1];
                s5 = pSrc[2 * i1 + 1] - pSrc[2 * i5 + 1];
                s2 = pSrc[2 * i2 + 1] + pSrc[2 * i6 + 1];
                s6 = pSrc[2 * i2 + 1] - pSrc[2 * i6 + 1];
                s3 = pSrc[2 * i3 + 1] + pSrc[2 * i7 + 1];
                s7 = pSrc[2 * i3 + 1] - pSrc[2 * i7 + 1];
                s4 = pSrc[2 * i4 + 1] + pSrc[2 * i8 + 1];
                s8 = pSrc[2 * i4 + 1] - pSrc[2 * i8 + 1];
                t2 = s1 - s3; s1 = s1 + s3; s3 = s2 - s4; s2 = s2 + s4;
                r1 = t1 + s3; t1 = t1 - s3;
                pSrc[2 * i1 + 1] = s1 + s2;
                s2 = s1 - s2;
                s1 = t2 - r3; t2 = t2 + r3;
                p1 = co5 * r2; p2 = si5 * s2; p3 = co5 * s2; p4 = si5 * r2;
                pSrc[2 * i5] = p1 + p2; pSrc[2 * i5 + 1] = p3 - p4;
                p1 = co3 * r1; p2 = si3 * s1; p3 = co3 * s1; p4 = si3 * r1;
                pSrc[2 * i3] = p1 + p2; pSrc[2 * i3 + 1] = p3 - p4;
                p1 = co7 * t1; p2 = si7 * t2; p3 = co7 * t2; p4 = si7 * t1;
                pSrc[2 * i7] = p1 + p2; pSrc[2 * i7 + 1] = p3 - p4;
                r1 = (r6 - r8) * C81; r6 = (r6 + r8) * C81;
                s1 = (s6 - s8) * C81; s6 = (s6 + s8) * C81;
                t1 = r5 - r1; r5 = r5 + r1; r8 = r7 - r6; r7 = r7 + r6;
                t2 = s5 - s1; s5 = s5 + s1; s8 = s7 - s6; s7 = s7 + s6;
                r1 = r5 + s7; r5 = r5 - s7; r6 = t1 + s8; t1 = t1 - s8;
                s1 = s5 - r7; s5 = s5 + r7; s6 = t2 - r8; t2 = t2 + r8;
                p1 = co2 * r1; p2 = si2 * s1; p3 = co2 * s1; p4 = si2 * r1;
                pSrc[2 * i2] = p1 + p2; pSrc[2 * i2 + 1] = p3 - p4;
                p1 = co8 * r5; p2 = si8 * s5; p3 = co8 * s5; p4 = si8 * r5;
                pSrc[2 * i8] = p1 + p2; pSrc[2 * i8 + 1] = p3 - p4;
                p1 = co6 * r6; p2 = si6 * s6; p3 = co6 * s6; p4 = si6 * r6;
                pSrc[2 * i6] = p1 + p2; pSrc[2 * i6 + 1] = p3 - p4;
                p1 = co4 * t1; p2 = si4 * t2; p3 = co4 * t2; p4 = si4 * t1;
                pSrc[2 * i4] = p1 + p2; pSrc[2 * i4 + 1] = p3 - p4;
                i1 += n1;
            } while (i1 < fftLen);
            j++;
        } while (j < n2);
        twidCoefModifier <<= 3;
    } while (n2 > 7);
}

/* arm_bitreversal_32 with armBitRevIndexTable512: the net effect is the base-8 digit-reversal
 * permutation of the complex samples, restated directly. */
static void digit_reverse_512(float32_t *p)
{
    for (uint32_t i = 0; i < 512; i++) {
        const uint32_t r = ((i & 7u) << 6) | (i & 0x38u) | (i >> 6);
        if (r > i) {
            float32_t tr = p[2 * i], ti = p[2 * i + 1];
            p[2 * i] = p[2 * r]; p[2 * i + 1] = p[2 * r + 1];
            p[2 * r] = tr; p[2 * r + 1] = ti;
        }
    }
}

void arm_cfft_f32(const arm_cfft_instance_f32 *S, float32_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag)
{
    tables_init();
    if (S->fftLen != 512) abort();      /* only the firmware's FFT_SIZE == 512 configuration is restated */
    if (ifftFlag) for (uint32_t l = 0; l < 512; l++) p1[2 * l + 1] = -p1[2 * l + 1];
    radix8_butterfly_f32(p1, 512, g_twiddle_512, 1);
    if (bitReverseFlag) digit_reverse_512(p1);
    if (ifftFlag) for (uint32_t l = 0; l < 512; l++) { p1[2 * l] *= (1.0f / 512.0f); p1[2 * l + 1] = -p1[2 * l + 1] * (1.0f / 512.0f); }
}

void arm_cmplx_mag_f32(float32_t *pSrc, float32_t *pDst, uint32_t numSamples)
{
    for (uint32_t i = 0; i < numSamples; i++) {
        const float32_t re = pSrc[2 * i], im = pSrc[2 * i + 1];
        arm_sqrt_f32((re * re) + (im * im), &pDst[i]);
    }
}
