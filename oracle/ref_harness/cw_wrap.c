/* cw_wrap.c - compiles the reference's cw_decoder.c where it lies (unmodified) and exposes the Goertzel
 * magnitude it keeps in a file-static.  TEST INFRASTRUCTURE ONLY. */
#include UA3_REF_CW_C
float ua3_cw_magnitude(void) { return magnitude; }
