/* usbd_ioreq.h - host stub of the ST USB device library types named by usbd_ua3reo.h (oracle only). */
#ifndef UA3_STUB_USBD_IOREQ_H
#define UA3_STUB_USBD_IOREQ_H
#include "stm32f4xx_hal.h"
#define USB_MAX_EP0_SIZE 64U
typedef struct { uint32_t dummy; } USBD_HandleTypeDef;
typedef struct { uint32_t dummy; } USBD_ClassTypeDef;
#endif
