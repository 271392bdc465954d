/*
  tmr_b200_ext.cpp -- B200-only extensions next to the flat C binding: access
  to the device forest behind a TMROctForest so that measurement code can drive
  the CUDA layer (include/tmrgpu.h) on device-resident inputs.
*/
#include "TMROctForest.h"
#include "tmrgpu.h"

extern "C" {

/* tmrgpu_forest behind a tmrc_forest handle (TMROctForest*) */
tmrgpu_forest *tmr_b200_device_forest(void *forest) {
  return static_cast<TMROctForest *>(forest)->getDeviceForest();
}

}  // extern "C"
