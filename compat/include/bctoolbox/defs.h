/* compat shim (our own code): the minimum of bctoolbox/defs.h the mediastreamer2 headers use.
 * Test/host-harness infrastructure only: lets the UNMODIFIED reference runtime under /root/reference
 * compile in a container that has no bctoolbox. A production drop-in builds against real bctoolbox. */
#ifndef MSB200_COMPAT_BCTBX_DEFS_H
#define MSB200_COMPAT_BCTBX_DEFS_H
#ifndef BCTBX_UNUSED
#ifdef __cplusplus
#define BCTBX_UNUSED(x)
#else
#define BCTBX_UNUSED(x) x __attribute__((unused))
#endif
#endif
#define BCTBX_PUBLIC
#define BCTBX_DEPRECATED __attribute__((deprecated))
#define BCTBX_NO_BREAK __attribute__((fallthrough))
#ifndef MIN
#define MIN(a, b) (((a) > (b)) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#endif
#endif
