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
#endif
