/* compat shim (our own code): the few oRTP port macros the mediastreamer2 headers reference. */
#ifndef MSB200_COMPAT_ORTP_PORT_H
#define MSB200_COMPAT_ORTP_PORT_H
#include "bctoolbox/port.h"
#define ORTP_INLINE inline
#define ORTP_PUBLIC
#define ORTP_VAR_PUBLIC extern
#define ORTP_DEPRECATED __attribute__((deprecated))
#define ortp_malloc bctbx_malloc
#define ortp_malloc0 bctbx_malloc0
#define ortp_free bctbx_free
#define ortp_new bctbx_new
#define ortp_new0 bctbx_new0
#endif
