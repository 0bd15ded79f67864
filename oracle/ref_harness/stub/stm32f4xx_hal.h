/*
 * stm32f4xx_hal.h - host stub of the STM32 HAL, just enough for the UA3REO firmware's signal-path
 * sources (audio_processor.c, audio_filters.c, agc.c, noise_reduction.c, fft.c, functions.c,
 * fpga.c) to compile UNMODIFIED on x86-64.  TEST INFRASTRUCTURE ONLY (oracle/ref_harness).
 *
 *  - DMA memory-to-memory transfers become memcpy of `len` 32-bit words (functions.c:15-19).
 *    The firmware passes addresses as uint32_t, so the harness is linked non-PIE (-no-pie) to keep
 *    every static buffer below 4 GiB.
 *  - GPIOx->IDR is routed through a function pointer so that the unmodified bus driver
 *    (fpga.c:286-401) reads the 8 frame bytes the harness queues, one per FPGA_readPacket().
 */
#ifndef UA3_STUB_STM32F4XX_HAL_H
#define UA3_STUB_STM32F4XX_HAL_H
#include <stdint.h>
#include <stdbool.h>
#include <string.h>

#define __IO volatile
#define __weak __attribute__((weak))
#define UNUSED(x) ((void)(x))

typedef enum { HAL_OK = 0, HAL_ERROR = 1, HAL_BUSY = 2, HAL_TIMEOUT = 3 } HAL_StatusTypeDef;
typedef enum { GPIO_PIN_RESET = 0, GPIO_PIN_SET } GPIO_PinState;
#define HAL_MAX_DELAY 0xFFFFFFFFU

typedef struct ua3_gpio {
    __IO uint32_t moder_slot[1], OTYPER, OSPEEDR, PUPDR;   /* MODER: see BSRR below */
    uint32_t (*idr_fn)(void);            /* stands in for the IDR register: see IDR below */
    __IO uint32_t ODR, bsrr_slot[1], LCKR;   /* BSRR: see below */
    __IO uint32_t AFR[2];
} GPIO_TypeDef;
#define IDR idr_fn()
/* BSRR is write-only set/reset.  Plain builds keep the last word written.  With -DUA3_BUS_HDL (fpga.c linked against the
 * reference's own stm32_interface.v, oracle/ref_harness/bus_hdl.c) every store `GPIOx->BSRR = v` first calls the hook, which
 * applies the PREVIOUS store to the pins (and clocks the Verilog model on a rising FPGA_CLK edge) - the index expression is
 * evaluated before the store, so the pin writes reach the model one by one, in program order. */
#ifdef UA3_BUS_HDL
int ua3_bsrr_hook(void);
#define BSRR bsrr_slot[ua3_bsrr_hook()]
#define MODER moder_slot[ua3_bsrr_hook()]   /* a direction change must not overtake the pin write before it */
#else
#define BSRR bsrr_slot[0]
#define MODER moder_slot[0]
#endif
typedef struct { uint32_t Pin, Mode, Pull, Speed, Alternate; } GPIO_InitTypeDef;
extern GPIO_TypeDef ua3_gpio_a, ua3_gpio_b, ua3_gpio_c, ua3_gpio_d, ua3_gpio_e;
#define GPIOA (&ua3_gpio_a)
#define GPIOB (&ua3_gpio_b)
#define GPIOC (&ua3_gpio_c)
#define GPIOD (&ua3_gpio_d)
#define GPIOE (&ua3_gpio_e)
#define GPIO_PIN_0 0x0001U
#define GPIO_PIN_1 0x0002U
#define GPIO_PIN_2 0x0004U
#define GPIO_PIN_3 0x0008U
#define GPIO_PIN_4 0x0010U
#define GPIO_PIN_5 0x0020U
#define GPIO_PIN_6 0x0040U
#define GPIO_PIN_7 0x0080U
#define GPIO_PIN_8 0x0100U
#define GPIO_PIN_9 0x0200U
#define GPIO_PIN_10 0x0400U
#define GPIO_PIN_11 0x0800U
#define GPIO_PIN_12 0x1000U
#define GPIO_PIN_13 0x2000U
#define GPIO_PIN_14 0x4000U
#define GPIO_PIN_15 0x8000U
#define GPIO_MODE_INPUT 0x00000000U
#define GPIO_MODE_OUTPUT_PP 0x00000001U
#define GPIO_NOPULL 0x00000000U
#define GPIO_PULLUP 0x00000001U
#define GPIO_PULLDOWN 0x00000002U
#define GPIO_SPEED_FREQ_VERY_HIGH 0x00000003U
#define GPIO_MODER_MODER0 0x00000003U
void HAL_GPIO_Init(GPIO_TypeDef *, GPIO_InitTypeDef *);
void HAL_GPIO_WritePin(GPIO_TypeDef *, uint16_t, GPIO_PinState);
GPIO_PinState HAL_GPIO_ReadPin(GPIO_TypeDef *, uint16_t);

typedef struct { uint32_t dummy; } DMA_Stream_TypeDef;
typedef struct { DMA_Stream_TypeDef *Instance; uint32_t dummy; } DMA_HandleTypeDef;
typedef enum { HAL_DMA_FULL_TRANSFER = 0, HAL_DMA_HALF_TRANSFER = 1 } HAL_DMA_LevelCompleteTypeDef;
HAL_StatusTypeDef HAL_DMA_Start(DMA_HandleTypeDef *, uint32_t src, uint32_t dst, uint32_t len_words);
HAL_StatusTypeDef HAL_DMA_Start_IT(DMA_HandleTypeDef *, uint32_t src, uint32_t dst, uint32_t len_words);
HAL_StatusTypeDef HAL_DMA_PollForTransfer(DMA_HandleTypeDef *, HAL_DMA_LevelCompleteTypeDef, uint32_t timeout);
#define __HAL_DMA_GET_COUNTER(h) (0U)

typedef struct { uint32_t dummy; } I2S_HandleTypeDef;
typedef struct { uint32_t dummy; } IWDG_HandleTypeDef;
typedef struct { uint32_t dummy; } SPI_HandleTypeDef;
typedef struct { uint32_t dummy; } UART_HandleTypeDef;
typedef struct { uint32_t dummy; } TIM_HandleTypeDef;
typedef struct { uint32_t dummy; } RTC_HandleTypeDef;
typedef struct { uint32_t dummy; } ADC_HandleTypeDef;
typedef struct { uint32_t dummy; } PCD_HandleTypeDef;
typedef struct { uint32_t dummy; } SRAM_HandleTypeDef;
typedef struct { uint32_t dummy; } I2C_HandleTypeDef;
extern IWDG_HandleTypeDef hiwdg;   /* main.c global the firmware sources use without a local extern */
HAL_StatusTypeDef HAL_UART_Transmit(UART_HandleTypeDef *, uint8_t *, uint16_t, uint32_t);
HAL_StatusTypeDef HAL_UART_Transmit_IT(UART_HandleTypeDef *, uint8_t *, uint16_t);
HAL_StatusTypeDef HAL_IWDG_Refresh(IWDG_HandleTypeDef *);
uint32_t HAL_GetTick(void);
void HAL_Delay(uint32_t);
uint32_t HAL_RCC_GetHCLKFreq(void);

typedef struct { __IO uint32_t CTRL, CYCCNT; } ua3_dwt_t;
typedef struct { __IO uint32_t DEMCR; } ua3_coredebug_t;
extern ua3_dwt_t ua3_dwt;
extern ua3_coredebug_t ua3_coredebug;
#define DWT (&ua3_dwt)
#define CoreDebug (&ua3_coredebug)
#define DWT_CTRL_CYCCNTENA_Pos 0U
#define DWT_CTRL_CYCCNTENA_Msk 1U
#define CoreDebug_DEMCR_TRCENA_Msk (1U << 24)
extern uint32_t SystemCoreClock;

#endif
