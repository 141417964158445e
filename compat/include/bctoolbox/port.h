/* compat shim (our own code): memory, threads (pthreads), time — the bctbx_* names the reference uses. */
#ifndef MSB200_COMPAT_BCTBX_PORT_H
#define MSB200_COMPAT_BCTBX_PORT_H
#include "bctoolbox/defs.h"
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef unsigned char bool_t;
#undef TRUE
#undef FALSE
#define TRUE 1
#define FALSE 0
typedef pthread_mutex_t bctbx_mutex_t;
typedef pthread_cond_t bctbx_cond_t;
typedef pthread_t bctbx_thread_t;
#define bctbx_mutex_init pthread_mutex_init
#define bctbx_mutex_lock pthread_mutex_lock
#define bctbx_mutex_unlock pthread_mutex_unlock
#define bctbx_mutex_destroy pthread_mutex_destroy
#define bctbx_cond_init pthread_cond_init
#define bctbx_cond_wait pthread_cond_wait
#define bctbx_cond_signal pthread_cond_signal
#define bctbx_cond_broadcast pthread_cond_broadcast
#define bctbx_cond_destroy pthread_cond_destroy
#define bctbx_thread_create pthread_create
#define bctbx_thread_join pthread_join
#define bctbx_thread_self pthread_self
#define bctbx_thread_exit pthread_exit
typedef struct {
	int64_t tv_sec;
	int64_t tv_nsec;
} bctoolboxTimeSpec;
void *bctbx_malloc(size_t sz);
void *bctbx_malloc0(size_t sz);
void *bctbx_realloc(void *p, size_t sz);
void bctbx_free(void *p);
char *bctbx_strdup(const char *s);
char *bctbx_strndup(const char *s, int n);
char *bctbx_strdup_printf(const char *fmt, ...);
char *bctbx_strdup_vprintf(const char *fmt, va_list ap);
char *bctbx_strcat_printf(char *dst, const char *fmt, ...);
#define bctbx_new(type, count) ((type *)bctbx_malloc(sizeof(type) * (count)))
#define bctbx_new0(type, count) ((type *)bctbx_malloc0(sizeof(type) * (count)))
void bctbx_get_cur_time(bctoolboxTimeSpec *ts);
uint64_t bctbx_get_cur_time_ms(void);
void bctbx_sleep_ms(int ms);
void bctbx_set_self_thread_name(const char *name);
bool_t bctbx_is_matching_regex_log(const char *entry, const char *regex, bool_t show_log);
#ifdef __cplusplus
}
#endif
#endif
