// todo_stubs.cu — TEMPORARY: entry points declared in include/msb200dsp.h whose kernels are not written yet.
// They fail loudly (MSB200_ESTATE); each block is deleted when the real implementation lands.
#include "msb200_internal.h"
#define NOTYET(name) msb200_set_error(name ": not implemented yet"); return MSB200_ESTATE
extern "C" {
int msb200_scaler_create(msb200_ctx *, int, int, int, int, int, int, msb200_scaler **) { NOTYET("msb200_scaler_create"); }
void msb200_scaler_destroy(msb200_scaler *) {}
size_t msb200_scaler_src_frame_bytes(msb200_scaler *) { return 0; }
size_t msb200_scaler_dst_frame_bytes(msb200_scaler *) { return 0; }
int msb200_scaler_process(msb200_scaler *, int, const uint8_t *, uint8_t *) { NOTYET("msb200_scaler_process"); }
int msb200_scaler_process_dev(msb200_scaler *, int, const void *, void *) { NOTYET("msb200_scaler_process_dev"); }
}
