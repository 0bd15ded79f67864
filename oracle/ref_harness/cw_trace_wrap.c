/* cw_trace_wrap.c - drives the reference's CW decoder, CWDecoder_Process() (cw_decoder.c:56-170, compiled unmodified where
 * it lies), over 192-sample audio blocks with a 4 ms tick per block and records what it appends to its element buffer
 * (dots/dashes) and to CW_Decoder_Text (decoded characters).  TEST INFRASTRUCTURE ONLY (oracle/_ref/fw_cw).
 *
 * The reference appends with unbounded strcat() into code[20] and into CW_Decoder_Text[15], an array initialised
 * with exactly 15 characters and no terminator (cw_decoder.c:14,37,252-257) - undefined behaviour on a host with
 * fortified libc.  The wrapper therefore routes the file's strcat/shiftTextLeft calls to a tracing version: appends to
 * `code` are kept (bounded to 19 characters, which is also all the decode table can match), appends to the text bar are
 * recorded and not performed.  Nothing else of the file is touched.
 *
 *   fw_cw <audio.f32> : blocks of 192 float32 samples; per block one output line
 *                       "<magnitude as %a> <CW_Decoder_WPM> <code after the block or => <characters appended, as hex bytes, or =>"
 */
#include <string.h>
#include <stdio.h>
#include <stdint.h>
static char ua3_text_trace[64];
static char *ua3_trace_strcat(char *dst, const char *src);
#define strcat ua3_trace_strcat
#define shiftTextLeft ua3_noop_shift
#include UA3_REF_CW_C
#undef strcat
#undef shiftTextLeft

void ua3_noop_shift(char *string, int16_t n) { (void)string; (void)n; }

static char *ua3_trace_strcat(char *dst, const char *src)
{
    if (dst == code) {
        size_t n = strlen(code);
        while (*src && n < sizeof code - 1) code[n++] = *src++;
        code[n] = '\0';
    } else {
        strncat(ua3_text_trace, src, sizeof ua3_text_trace - strlen(ua3_text_trace) - 1);
    }
    return dst;
}

struct TRX_SETTINGS TRX;
volatile DEF_LCD_UpdateQuery LCD_UpdateQuery;
static uint32_t g_tick;
uint32_t HAL_GetTick(void) { return g_tick; }

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: fw_cw audio.f32\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    CWDecoder_Init();
    float blk[CWDECODER_SAMPLES];
    while (fread(blk, sizeof(float), CWDECODER_SAMPLES, f) == CWDECODER_SAMPLES) {
        g_tick += 4;                                   /* 192 samples at 48 kHz */
        ua3_text_trace[0] = '\0';
        CWDecoder_Process(blk);
        printf("%a %u %s ", (double)magnitude, (unsigned)CW_Decoder_WPM, code[0] ? code : "=");
        if (!ua3_text_trace[0]) printf("=");
        for (const char *p = ua3_text_trace; *p; ++p) printf("%02x", (unsigned)(unsigned char)*p);     /* hex: ' ', '_', '-' are all decodable characters */
        printf("\n");
    }
    fclose(f);
    return 0;
}
