/* autogain_wrap.c - compiles the reference's trx_manager.c where it lies (unmodified) and drives its
 * TRX_DoAutoGain() state machine (trx_manager.c:268-356) over a sequence of TRX_ADC_MAXAMPLITUDE values read from
 * stdin; prints stage, wait counter, Preamp, ATT, LPF, BPF after every call.  TEST INFRASTRUCTURE ONLY
 * (oracle/_ref/fw_autogain; linked with unresolved symbols ignored - nothing else of trx_manager.c is called). */
#include UA3_REF_TRX_C
#include <stdio.h>
struct TRX_SETTINGS TRX;
volatile DEF_LCD_UpdateQuery LCD_UpdateQuery;
int main(void)
{
    TRX.AutoGain = true;
    int v;
    while (scanf("%d", &v) == 1) {
        TRX_ADC_MAXAMPLITUDE = v;
        TRX_DoAutoGain();
        printf("%d %d %d %d %d %d\n", autogain_stage, autogain_wait_reaction, TRX.Preamp, TRX.ATT, TRX.LPF, TRX.BPF);
    }
    return 0;
}
