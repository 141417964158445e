/* compat shim (our own code): OrtpExtremum (sliding min/max used by MSVolume) and ortp_log10f. */
#ifndef MSB200_COMPAT_ORTP_UTILS_H
#define MSB200_COMPAT_ORTP_UTILS_H
#include "ortp/port.h"
#include <math.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct _OrtpExtremum {
	float current_extremum;
	float last_stable;
	uint64_t extremum_time;
	int period;
} OrtpExtremum;
void ortp_extremum_reset(OrtpExtremum *obj);
void ortp_extremum_init(OrtpExtremum *obj, int period);
bool_t ortp_extremum_record_min(OrtpExtremum *obj, uint64_t curtime, float value);
bool_t ortp_extremum_record_max(OrtpExtremum *obj, uint64_t curtime, float value);
float ortp_extremum_get_current(OrtpExtremum *obj);
float ortp_extremum_get_previous(OrtpExtremum *obj);
#define ortp_log10f(x) log10f(x)
#ifdef __cplusplus
}
#endif
#endif
