/* compat shim (our own code): opaque PayloadType/RtpProfile so mscommon.h and msfactory.h parse. */
#ifndef MSB200_COMPAT_ORTP_PAYLOADTYPE_H
#define MSB200_COMPAT_ORTP_PAYLOADTYPE_H
#include "ortp/port.h"
typedef struct _PayloadType {
	int type;
	int clock_rate;
	char bits_per_sample;
	char *zero_pattern;
	int pattern_length;
	int normal_bitrate;
	char *mime_type;
	int channels;
	char *recv_fmtp;
	char *send_fmtp;
	int flags;
	void *user_data;
} PayloadType;
typedef PayloadType OrtpPayloadType;
typedef struct _RtpProfile RtpProfile;
/* "a=fmtp" parameter lookup (oRTP payloadtype.c): value of `param_name` in a "k1=v1; k2=v2" list */
bool_t fmtp_get_value(const char *fmtp, const char *param_name, char *result, size_t result_len);
#endif
