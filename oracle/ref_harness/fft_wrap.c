/* fft_wrap.c - compiles the reference's fft.c where it lies (unmodified) and adds read accessors for
 * its file-static results.  TEST INFRASTRUCTURE ONLY. */
#include UA3_REF_FFT_C
const float *ua3_fft_output_mean(void) { return FFTOutput_mean; }
const uint16_t *ua3_fft_wtf_row0(void) { return &wtf_buffer[0][0]; }
float ua3_fft_max_value(void) { return maxValueFFT; }
uint16_t ua3_fft_max_value_errors(void) { return maxValueErrors; }
const uint16_t *ua3_fft_wtf_all(void) { return &wtf_buffer[0][0]; }    /* FFT_WTF_HEIGHT x FFT_PRINT_SIZE, row 0 newest */
