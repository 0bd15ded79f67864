/* arm_const_structs.h - CMSIS-DSP 1.6.0 CFFT instance constants used by fft.c:12-17 (oracle/ only). */
#ifndef UA3_ORACLE_ARM_CONST_STRUCTS_H
#define UA3_ORACLE_ARM_CONST_STRUCTS_H
#include "arm_math.h"
extern const arm_cfft_instance_f32 arm_cfft_sR_f32_len512;
extern const arm_cfft_instance_f32 arm_cfft_sR_f32_len256;
#endif
