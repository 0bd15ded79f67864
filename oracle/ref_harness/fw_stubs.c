/*
 * fw_stubs.c - definitions the firmware's signal-path sources expect from the parts of the firmware
 * that are NOT on the path (HAL, LCD, codec, USB, TRX manager, settings flash).  TEST INFRASTRUCTURE.
 * Behaviour-relevant stubs:
 *   HAL_DMA_Start/_IT     memcpy of len 32-bit words (what DMA2 mem-to-mem does; functions.c:15-19)
 *   TRX_getMode/CurrentVFO/TRX_on_TX  as trx_manager.c:57-61,222-225 and settings.c:98-104
 *   GPIOA->IDR            next byte of the frame queued by the harness (bus model of stm32_interface.v:228-271)
 */
#include "stm32f4xx_hal.h"
#include "arm_math.h"
#include "settings.h"
#include "trx_manager.h"
#include "fpga.h"
#include "wm8731.h"
#include "lcd.h"
#include "usbd_audio_if.h"
#include "usbd_debug_if.h"
#include "cw_decoder.h"
#include <stdio.h>
#include <stdlib.h>

/* ---- GPIO / bus model ---- */
uint8_t ua3_bus_frame[8];
int ua3_bus_pos = 0;
static uint32_t bus_read(void) { const uint32_t v = ua3_bus_frame[ua3_bus_pos & 7]; ua3_bus_pos++; return v; }
static uint32_t zero_read(void) { return 0; }
GPIO_TypeDef ua3_gpio_a = {.idr_fn = bus_read}, ua3_gpio_b = {.idr_fn = zero_read}, ua3_gpio_c = {.idr_fn = zero_read},
             ua3_gpio_d = {.idr_fn = zero_read}, ua3_gpio_e = {.idr_fn = zero_read};
void HAL_GPIO_Init(GPIO_TypeDef *g, GPIO_InitTypeDef *i)            /* the direction bits only: bus_hdl.c resolves the bus with them */
{
    for (uint32_t p = 0; p < 16; p++)
        if (i->Pin & (1u << p)) g->MODER = (g->MODER & ~(3u << (2 * p))) | ((i->Mode & 3u) << (2 * p));
}
void HAL_GPIO_WritePin(GPIO_TypeDef *g, uint16_t p, GPIO_PinState s) { (void)g; (void)p; (void)s; }
GPIO_PinState HAL_GPIO_ReadPin(GPIO_TypeDef *g, uint16_t p) { (void)g; (void)p; return GPIO_PIN_RESET; }

/* ---- DMA: memory-to-memory copy of len words; addresses are 32-bit (binary is linked -no-pie) ---- */
HAL_StatusTypeDef HAL_DMA_Start(DMA_HandleTypeDef *h, uint32_t src, uint32_t dst, uint32_t len)
{
    (void)h;
    memmove((void *)(uintptr_t)dst, (const void *)(uintptr_t)src, (size_t)len * 4u);
    return HAL_OK;
}
HAL_StatusTypeDef HAL_DMA_Start_IT(DMA_HandleTypeDef *h, uint32_t src, uint32_t dst, uint32_t len) { return HAL_DMA_Start(h, src, dst, len); }
HAL_StatusTypeDef HAL_DMA_PollForTransfer(DMA_HandleTypeDef *h, HAL_DMA_LevelCompleteTypeDef l, uint32_t t) { (void)h; (void)l; (void)t; return HAL_OK; }
DMA_HandleTypeDef hdma_memtomem_dma2_stream0, hdma_memtomem_dma2_stream1, hdma_memtomem_dma2_stream3,
                  hdma_memtomem_dma2_stream6, hdma_memtomem_dma2_stream7, hdma_i2s3_ext_rx, hdma_spi3_tx;

/* ---- misc HAL ---- */
ua3_dwt_t ua3_dwt; ua3_coredebug_t ua3_coredebug; uint32_t SystemCoreClock = 168000000u;
I2S_HandleTypeDef hi2s3; IWDG_HandleTypeDef hiwdg; UART_HandleTypeDef huart1; SPI_HandleTypeDef hspi1;
static uint32_t g_tick = 0;
uint32_t HAL_GetTick(void) { return g_tick; }
void ua3_set_tick(uint32_t t) { g_tick = t; }
void HAL_Delay(uint32_t ms) { g_tick += ms; }
HAL_StatusTypeDef HAL_UART_Transmit(UART_HandleTypeDef *h, uint8_t *d, uint16_t n, uint32_t t) { (void)h; (void)d; (void)n; (void)t; return HAL_OK; }
HAL_StatusTypeDef HAL_IWDG_Refresh(IWDG_HandleTypeDef *h) { (void)h; return HAL_OK; }
void DEBUG_Transmit_FIFO(uint8_t *d, uint16_t n) { (void)d; (void)n; }
bool DEBUG_Transmit_FIFO_Events(void) { return true; }

/* ---- codec / USB audio buffers (wm8731.c, usbd_audio_if.c) ---- */
int32_t CODEC_Audio_Buffer_RX[CODEC_AUDIO_BUFFER_SIZE];
int32_t CODEC_Audio_Buffer_TX[CODEC_AUDIO_BUFFER_SIZE];
volatile bool WM8731_DMA_state = true;
volatile bool WM8731_Buffer_underrun = false;
volatile uint32_t WM8731_DMA_samples = 0;
int16_t USB_AUDIO_rx_buffer_a[(USB_AUDIO_RX_BUFFER_SIZE / 2)];
int16_t USB_AUDIO_rx_buffer_b[(USB_AUDIO_RX_BUFFER_SIZE / 2)];
int16_t USB_AUDIO_tx_buffer[(USB_AUDIO_TX_BUFFER_SIZE / 2)];
volatile bool USB_AUDIO_current_rx_buffer = false;
volatile bool USB_AUDIO_need_rx_buffer = false;
int16_t USB_AUDIO_GetTXBufferIndex_FS(void) { return 0; }
volatile uint32_t RX_USB_AUDIO_SAMPLES = 0, TX_USB_AUDIO_SAMPLES = 0;
volatile bool RX_USB_AUDIO_underrun = false;

/* ---- LCD (lcd.c, LCD/lcd_driver.c): the FFT display code draws through these ---- */
volatile bool LCD_busy = false, LCD_bandMenuOpened = false, LCD_timeMenuOpened = false, LCD_modeMenuOpened = false,
              LCD_systemMenuOpened = false, LCD_mainMenuOpened = false, LCD_NotchEdit = false;
volatile DEF_LCD_UpdateQuery LCD_UpdateQuery;
static uint32_t lcd_scratch[1024];
uint32_t LCD_FSMC_COMM_ADDR = 0, LCD_FSMC_DATA_ADDR = 0;
void ua3_lcd_stub_init(void) { LCD_FSMC_DATA_ADDR = (uint32_t)(uintptr_t)lcd_scratch; LCD_FSMC_COMM_ADDR = LCD_FSMC_DATA_ADDR; }
void LCD_showError(char text[], bool redraw) { (void)redraw; fprintf(stderr, "LCD_showError: %s\n", text); }
void LCDDriver_SetCursorAreaPosition(uint16_t x1, uint16_t y1, uint16_t x2, uint16_t y2) { (void)x1; (void)y1; (void)x2; (void)y2; }
void LCDDriver_SendData(uint16_t d) { (void)d; }
void LCDDriver_drawFastHLine(int16_t x, int16_t y, int16_t w, uint16_t c) { (void)x; (void)y; (void)w; (void)c; }
void LCDDriver_drawFastVLine(int16_t x, int16_t y, int16_t h, uint16_t c) { (void)x; (void)y; (void)h; (void)c; }
/* RGB565 packing used by the waterfall colour map (prototype LCD/lcd_driver.h:487): 5-6-5 bits, truncating */
uint16_t rgb888torgb565(uint8_t red, uint8_t green, uint8_t blue)
{
    return (uint16_t)(((unsigned)(red & 0xF8u) << 8) | ((unsigned)(green & 0xFCu) << 3) | ((unsigned)blue >> 3));
}

/* ---- TRX manager / settings subset ---- */
struct TRX_SETTINGS TRX;
volatile bool NeedSaveSettings = false;
volatile bool TRX_ptt_hard = false, TRX_ptt_cat = false, TRX_old_ptt_cat = false, TRX_key_serial = false,
              TRX_old_key_serial = false, TRX_key_hard = false, TRX_IQ_swap = false, TRX_squelched = false,
              TRX_tune = false, TRX_inited = true, TRX_ADC_OTR = false, TRX_DAC_OTR = false;
volatile uint16_t TRX_Key_Timeout_est = 0;
volatile int16_t TRX_RX_dBm = -100, TRX_ADC_MINAMPLITUDE = 0, TRX_ADC_MAXAMPLITUDE = 0;
VFO *CurrentVFO(void) { return !TRX.current_vfo ? &TRX.VFO_A : &TRX.VFO_B; }          /* settings.c:98-104 */
uint8_t TRX_getMode(void) { return CurrentVFO()->Mode; }                               /* trx_manager.c:222-225 */
bool TRX_on_TX(void)                                                                    /* trx_manager.c:57-61 */
{
    if (TRX_ptt_hard || TRX_ptt_cat || TRX_tune || TRX_getMode() == TRX_MODE_LOOPBACK || TRX_Key_Timeout_est > 0) return true;
    return false;
}
volatile TRX_FrontPanel_Type TRX_FrontPanel;
volatile uint8_t TRX_Time_InActive = 0, TRX_Fan_Timeout = 0;
uint32_t TRX_getFrequency(void) { return CurrentVFO()->Freq; }

/* the CW decoder itself (cw_decoder.c) is compiled from the reference through cw_wrap.c */

/* profiler.c */
void StartProfiler(uint8_t pid) { (void)pid; }
void EndProfiler(uint8_t pid, bool s) { (void)pid; (void)s; }
