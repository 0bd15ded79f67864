/*
 * arm_math.h - portable restatement of the CMSIS-DSP 1.6.0 float32 API subset that the UA3REO
 * firmware's receive/transmit audio path links against.  TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * CMSIS-DSP is a third-party dependency of the reference that is NOT vendored in its tree: the
 * firmware links the prebuilt arm_cortexM4lf_math.lib of Keil pack ARM.CMSIS 5.5.1 / CMSIS-DSP 1.6.0
 * (STM32/MDK-ARM/UA3REO.uvprojx:339,943-944).  The pack is not in the build image either, so the
 * functions below restate the published algorithms of that version (operation order as in the
 * upstream sources, IEEE-754 binary32, no FMA contraction).  Call sites: SURVEY.md 2.4.
 * Provenance: restated from the public CMSIS 5.5.1 sources as recalled - parity with the binary
 * library is unpinned (no copy of it is available to run).
 */
#ifndef UA3_ORACLE_ARM_MATH_H
#define UA3_ORACLE_ARM_MATH_H
#include <stdint.h>
#include <string.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float float32_t;
typedef double float64_t;
typedef int16_t q15_t;
typedef int32_t q31_t;

#ifndef PI
#define PI 3.14159265358979f
#endif
#define FAST_MATH_TABLE_SIZE 512

typedef struct { uint16_t numStages; float32_t *pState; float32_t *pkCoeffs; float32_t *pvCoeffs; } arm_iir_lattice_instance_f32;
typedef struct { uint8_t numStages; float32_t *pState; float32_t *pCoeffs; } arm_biquad_cascade_df2T_instance_f32;
typedef struct { uint32_t numStages; float32_t *pState; float32_t *pCoeffs; } arm_biquad_casd_df1_inst_f32;
typedef struct { uint16_t numTaps; float32_t *pState; float32_t *pCoeffs; } arm_fir_instance_f32;
typedef struct { uint8_t M; uint16_t numTaps; float32_t *pCoeffs; float32_t *pState; } arm_fir_decimate_instance_f32;
typedef struct { uint16_t numTaps; float32_t *pState; float32_t *pCoeffs; float32_t mu; float32_t energy; float32_t x0; } arm_lms_norm_instance_f32;
typedef struct { uint16_t fftLen; const float32_t *pTwiddle; const uint16_t *pBitRevTable; uint16_t bitRevLength; } arm_cfft_instance_f32;
typedef enum { ARM_MATH_SUCCESS = 0, ARM_MATH_ARGUMENT_ERROR = -1, ARM_MATH_LENGTH_ERROR = -2 } arm_status;

void arm_iir_lattice_init_f32(arm_iir_lattice_instance_f32 *S, uint16_t numStages, float32_t *pkCoeffs, float32_t *pvCoeffs, float32_t *pState, uint32_t blockSize);
void arm_iir_lattice_f32(const arm_iir_lattice_instance_f32 *S, float32_t *pSrc, float32_t *pDst, uint32_t blockSize);
void arm_biquad_cascade_df2T_init_f32(arm_biquad_cascade_df2T_instance_f32 *S, uint8_t numStages, float32_t *pCoeffs, float32_t *pState);
void arm_biquad_cascade_df2T_f32(const arm_biquad_cascade_df2T_instance_f32 *S, float32_t *pSrc, float32_t *pDst, uint32_t blockSize);
void arm_biquad_cascade_df1_f32(const arm_biquad_casd_df1_inst_f32 *S, float32_t *pSrc, float32_t *pDst, uint32_t blockSize);
void arm_fir_init_f32(arm_fir_instance_f32 *S, uint16_t numTaps, float32_t *pCoeffs, float32_t *pState, uint32_t blockSize);
void arm_fir_f32(const arm_fir_instance_f32 *S, float32_t *pSrc, float32_t *pDst, uint32_t blockSize);
arm_status arm_fir_decimate_init_f32(arm_fir_decimate_instance_f32 *S, uint16_t numTaps, uint8_t M, float32_t *pCoeffs, float32_t *pState, uint32_t blockSize);
void arm_fir_decimate_f32(const arm_fir_decimate_instance_f32 *S, float32_t *pSrc, float32_t *pDst, uint32_t blockSize);
void arm_lms_norm_init_f32(arm_lms_norm_instance_f32 *S, uint16_t numTaps, float32_t *pCoeffs, float32_t *pState, float32_t mu, uint32_t blockSize);
void arm_lms_norm_f32(arm_lms_norm_instance_f32 *S, float32_t *pSrc, float32_t *pRef, float32_t *pOut, float32_t *pErr, uint32_t blockSize);
void arm_cfft_f32(const arm_cfft_instance_f32 *S, float32_t *p1, uint8_t ifftFlag, uint8_t bitReverseFlag);
void arm_cmplx_mag_f32(float32_t *pSrc, float32_t *pDst, uint32_t numSamples);
float32_t arm_cos_f32(float32_t x);
float32_t arm_sin_f32(float32_t x);
void arm_max_f32(float32_t *pSrc, uint32_t blockSize, float32_t *pResult, uint32_t *pIndex);
void arm_min_f32(float32_t *pSrc, uint32_t blockSize, float32_t *pResult, uint32_t *pIndex);
void arm_mean_f32(float32_t *pSrc, uint32_t blockSize, float32_t *pResult);
void arm_scale_f32(float32_t *pSrc, float32_t scale, float32_t *pDst, uint32_t blockSize);
void arm_add_f32(float32_t *pSrcA, float32_t *pSrcB, float32_t *pDst, uint32_t blockSize);
void arm_sub_f32(float32_t *pSrcA, float32_t *pSrcB, float32_t *pDst, uint32_t blockSize);
void arm_mult_f32(float32_t *pSrcA, float32_t *pSrcB, float32_t *pDst, uint32_t blockSize);
void arm_abs_f32(float32_t *pSrc, float32_t *pDst, uint32_t blockSize);
void arm_copy_f32(float32_t *pSrc, float32_t *pDst, uint32_t blockSize);
void arm_fill_f32(float32_t value, float32_t *pDst, uint32_t blockSize);
void arm_offset_f32(float32_t *pSrc, float32_t offset, float32_t *pDst, uint32_t blockSize);
void arm_negate_f32(float32_t *pSrc, float32_t *pDst, uint32_t blockSize);
void arm_power_f32(float32_t *pSrc, uint32_t blockSize, float32_t *pResult);
void arm_rms_f32(float32_t *pSrc, uint32_t blockSize, float32_t *pResult);

static inline arm_status arm_sqrt_f32(float32_t in, float32_t *pOut)
{
    if (in >= 0.0f) { *pOut = sqrtf(in); return ARM_MATH_SUCCESS; }
    *pOut = 0.0f;
    return ARM_MATH_ARGUMENT_ERROR;
}

/* the sine table of arm_common_tables.c (513 entries, 8 decimals), exposed for the tests */
const float32_t *ua3_cmsis_sin_table(void);
const float32_t *ua3_cmsis_twiddle_512(void);

#ifdef __cplusplus
}
#endif
#endif
