/* shim_access.c - the accessors fw_rx.c reads its results through, for the GPU-backed build (fw_rx_b200): the same
 * names fft_wrap.c / cw_wrap.c give the all-CPU build, served from the buffers ua3reo_fw_shim.c publishes.
 * TEST INFRASTRUCTURE ONLY. */
#include <stdint.h>
extern float ua3reo_shim_fft_mean[];
extern uint16_t ua3reo_shim_wtf_row0[];
extern float ua3reo_shim_cw_magnitude;
extern uint16_t ua3reo_shim_wtf_buffer[];
const float *ua3_fft_output_mean(void) { return ua3reo_shim_fft_mean; }
const uint16_t *ua3_fft_wtf_row0(void) { return ua3reo_shim_wtf_row0; }
float ua3_fft_max_value(void) { return 0.0f; }     /* maxValueFFT stays on the device; the tests skip this column */
float ua3_cw_magnitude(void) { return ua3reo_shim_cw_magnitude; }
const uint16_t *ua3_fft_wtf_all(void) { return ua3reo_shim_wtf_buffer; }
