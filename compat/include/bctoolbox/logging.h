/* compat shim (our own code): printf-style logging under the bctbx_* names. */
#ifndef MSB200_COMPAT_BCTBX_LOGGING_H
#define MSB200_COMPAT_BCTBX_LOGGING_H
#include "bctoolbox/defs.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef enum {
	BCTBX_LOG_DEBUG = 1, BCTBX_LOG_TRACE = 1 << 1, BCTBX_LOG_MESSAGE = 1 << 2, BCTBX_LOG_WARNING = 1 << 3,
	BCTBX_LOG_ERROR = 1 << 4, BCTBX_LOG_FATAL = 1 << 5, BCTBX_LOG_LOGLEV_END = 1 << 6
} BctbxLogLevel;
#ifndef BCTBX_LOG_DOMAIN
#define BCTBX_LOG_DOMAIN "mediastreamer"
#endif
typedef struct _bctbx_log_tags bctbx_log_tags_t;
void bctbx_set_log_level(const char *domain, BctbxLogLevel level);
void bctbx_set_log_level_mask(const char *domain, int mask);
unsigned int bctbx_get_log_level_mask(const char *domain);
void bctbx_debug(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void bctbx_message(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void bctbx_warning(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void bctbx_error(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void bctbx_fatal(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
bctbx_log_tags_t *bctbx_create_log_tags_copy(void);
void bctbx_paste_log_tags(const bctbx_log_tags_t *tags);
void bctbx_log_tags_destroy(bctbx_log_tags_t *tags);
#ifdef __cplusplus
}
#endif
#endif
