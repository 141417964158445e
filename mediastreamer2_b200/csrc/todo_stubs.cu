// todo_stubs.cu — TEMPORARY: entry points declared in include/msb200dsp.h whose kernels are not written yet.
// They fail loudly (MSB200_ESTATE); each block is deleted when the real implementation lands.
#include "msb200_internal.h"
#define NOTYET(name) msb200_set_error(name ": not implemented yet"); return MSB200_ESTATE
extern "C" {
int msb200_aec_frame_size_for_rate(int sample_rate, int framesize_at_8000) { // adjust_framesize, speexec.c:171-180
	int newsize = (framesize_at_8000 * sample_rate) / 8000, n = 1, next;
	while ((next = n << 1) <= newsize) n = next;
	return n;
}
int msb200_aec_create(msb200_ctx *, int, int, int, int, msb200_aec **) { NOTYET("msb200_aec_create"); }
void msb200_aec_destroy(msb200_aec *) {}
int msb200_aec_get_info(msb200_aec *, msb200_aec_info *) { NOTYET("msb200_aec_get_info"); }
int msb200_aec_reset(msb200_aec *, int) { NOTYET("msb200_aec_reset"); }
int msb200_aec_process(msb200_aec *, const int16_t *, const int16_t *, int16_t *, int) { NOTYET("msb200_aec_process"); }
int msb200_aec_process_dev(msb200_aec *, const void *, const void *, void *, int, int) { NOTYET("msb200_aec_process_dev"); }
size_t msb200_aec_state_blob_size(msb200_aec *) { return 0; }
int msb200_aec_get_state_blob(msb200_aec *, int, void *, size_t) { NOTYET("msb200_aec_get_state_blob"); }
int msb200_aec_set_state_blob(msb200_aec *, int, const void *, size_t) { NOTYET("msb200_aec_set_state_blob"); }
int msb200_aec_probe(msb200_aec *, int, const char *, float *, int) { NOTYET("msb200_aec_probe"); }
int msb200_chain_create(msb200_ctx *, const msb200_chain_params *, msb200_chain **) { NOTYET("msb200_chain_create"); }
void msb200_chain_destroy(msb200_chain *) {}
int msb200_chain_next_out_samples(msb200_chain *) { NOTYET("msb200_chain_next_out_samples"); }
int msb200_chain_max_out_samples(msb200_chain *) { NOTYET("msb200_chain_max_out_samples"); }
int msb200_chain_tick(msb200_chain *, const int16_t *, const int16_t *, int16_t *, int *) { NOTYET("msb200_chain_tick"); }
int msb200_chain_tick_dev(msb200_chain *, const void *, const void *, void *, int *) { NOTYET("msb200_chain_tick_dev"); }
int msb200_chain_launches_per_tick(msb200_chain *) { NOTYET("msb200_chain_launches_per_tick"); }
msb200_aec *msb200_chain_aec(msb200_chain *) { return nullptr; }
int msb200_scaler_create(msb200_ctx *, int, int, int, int, int, int, msb200_scaler **) { NOTYET("msb200_scaler_create"); }
void msb200_scaler_destroy(msb200_scaler *) {}
size_t msb200_scaler_src_frame_bytes(msb200_scaler *) { return 0; }
size_t msb200_scaler_dst_frame_bytes(msb200_scaler *) { return 0; }
int msb200_scaler_process(msb200_scaler *, int, const uint8_t *, uint8_t *) { NOTYET("msb200_scaler_process"); }
int msb200_scaler_process_dev(msb200_scaler *, int, const void *, void *) { NOTYET("msb200_scaler_process_dev"); }
}
