/* compat shim (our own code, test/host-harness infrastructure): implements the bctoolbox / oRTP symbols that the
 * UNMODIFIED mediastreamer2 base runtime (src/base/{msfilter,msqueue,msticker,eventqueue,msfactory}.c) and the in-tree
 * filters call, so that they compile and run in a container that has neither library.
 * Nothing here is on the product's data path. */
#define _GNU_SOURCE
#include "bctoolbox/list.h"
#include "bctoolbox/logging.h"
#include "bctoolbox/port.h"
#include "ortp/str_utils.h"
#include "ortp/utils.h"
#include "ortp/payloadtype.h"

/* ------------------------------------------------------------------ memory / strings */
void *bctbx_malloc(size_t sz) {
	return malloc(sz ? sz : 1);
}
void *bctbx_malloc0(size_t sz) {
	return calloc(1, sz ? sz : 1);
}
void *bctbx_realloc(void *p, size_t sz) {
	return realloc(p, sz);
}
void bctbx_free(void *p) {
	free(p);
}
char *bctbx_strdup(const char *s) {
	return s ? strdup(s) : NULL;
}
char *bctbx_strndup(const char *s, int n) {
	return s ? strndup(s, (size_t)n) : NULL;
}
char *bctbx_strdup_vprintf(const char *fmt, va_list ap) {
	char *out = NULL;
	if (vasprintf(&out, fmt, ap) < 0) return NULL;
	return out;
}
char *bctbx_strdup_printf(const char *fmt, ...) {
	va_list ap;
	char *out;
	va_start(ap, fmt);
	out = bctbx_strdup_vprintf(fmt, ap);
	va_end(ap);
	return out;
}
char *bctbx_strcat_printf(char *dst, const char *fmt, ...) {
	va_list ap;
	char *tail, *out;
	va_start(ap, fmt);
	tail = bctbx_strdup_vprintf(fmt, ap);
	va_end(ap);
	if (!dst) return tail;
	out = (char *)realloc(dst, strlen(dst) + strlen(tail) + 1);
	strcat(out, tail);
	free(tail);
	return out;
}

/* ------------------------------------------------------------------ time / threads */
void bctbx_get_cur_time(bctoolboxTimeSpec *ts) {
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	ts->tv_sec = t.tv_sec;
	ts->tv_nsec = t.tv_nsec;
}
uint64_t bctbx_get_cur_time_ms(void) {
	bctoolboxTimeSpec ts;
	bctbx_get_cur_time(&ts);
	return (uint64_t)ts.tv_sec * 1000ULL + (uint64_t)((ts.tv_nsec + 500000LL) / 1000000LL);
}
void bctbx_sleep_ms(int ms) {
	struct timespec t;
	t.tv_sec = ms / 1000;
	t.tv_nsec = (long)(ms % 1000) * 1000000L;
	nanosleep(&t, NULL);
}
void bctbx_set_self_thread_name(const char *name) {
	char buf[16];
	strncpy(buf, name ? name : "ms", sizeof(buf) - 1);
	buf[sizeof(buf) - 1] = 0;
	pthread_setname_np(pthread_self(), buf);
}
bool_t bctbx_is_matching_regex_log(const char *entry, const char *regex, bool_t show_log) {
	(void)show_log;
	return entry && regex && strstr(entry, regex) != NULL;
}

/* ------------------------------------------------------------------ logging */
static int log_mask = BCTBX_LOG_WARNING | BCTBX_LOG_ERROR | BCTBX_LOG_FATAL;
void bctbx_set_log_level_mask(const char *domain, int mask) {
	(void)domain;
	log_mask = mask;
}
unsigned int bctbx_get_log_level_mask(const char *domain) {
	(void)domain;
	return (unsigned)log_mask;
}
void bctbx_set_log_level(const char *domain, BctbxLogLevel level) {
	int mask = BCTBX_LOG_FATAL;
	(void)domain;
	if (level <= BCTBX_LOG_ERROR) mask |= BCTBX_LOG_ERROR;
	if (level <= BCTBX_LOG_WARNING) mask |= BCTBX_LOG_WARNING;
	if (level <= BCTBX_LOG_MESSAGE) mask |= BCTBX_LOG_MESSAGE;
	if (level <= BCTBX_LOG_DEBUG) mask |= BCTBX_LOG_DEBUG | BCTBX_LOG_TRACE;
	log_mask = mask;
}
static void vlog(int lev, const char *tag, const char *fmt, va_list ap) {
	static int env_checked = 0;
	if (!env_checked) {
		const char *e = getenv("MSB200_COMPAT_LOG");
		env_checked = 1;
		if (e && *e == '1') log_mask |= BCTBX_LOG_MESSAGE;
		if (e && *e == '0') log_mask = BCTBX_LOG_FATAL;
	}
	if (!(log_mask & lev)) return;
	fprintf(stderr, "ms2-%s: ", tag);
	vfprintf(stderr, fmt, ap);
	fputc('\n', stderr);
}
#define DEFLOG(name, lev, tag)                                                                                         \
	void name(const char *fmt, ...) {                                                                                  \
		va_list ap;                                                                                                    \
		va_start(ap, fmt);                                                                                             \
		vlog(lev, tag, fmt, ap);                                                                                       \
		va_end(ap);                                                                                                    \
	}
DEFLOG(bctbx_debug, BCTBX_LOG_DEBUG, "debug")
DEFLOG(bctbx_message, BCTBX_LOG_MESSAGE, "message")
DEFLOG(bctbx_warning, BCTBX_LOG_WARNING, "warning")
DEFLOG(bctbx_error, BCTBX_LOG_ERROR, "error")
void bctbx_fatal(const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vlog(BCTBX_LOG_FATAL, "fatal", fmt, ap);
	va_end(ap);
	abort();
}
bctbx_log_tags_t *bctbx_create_log_tags_copy(void) {
	return NULL;
}
void bctbx_paste_log_tags(const bctbx_log_tags_t *tags) {
	(void)tags;
}
void bctbx_log_tags_destroy(bctbx_log_tags_t *tags) {
	(void)tags;
}

/* ------------------------------------------------------------------ lists */
bctbx_list_t *bctbx_list_new(void *data) {
	bctbx_list_t *e = (bctbx_list_t *)calloc(1, sizeof(*e));
	e->data = data;
	return e;
}
bctbx_list_t *bctbx_list_append_link(bctbx_list_t *l, bctbx_list_t *n) {
	bctbx_list_t *it = l;
	if (!n) return l;
	if (!l) return n;
	while (it->next) it = it->next;
	it->next = n;
	n->prev = it;
	return l;
}
bctbx_list_t *bctbx_list_append(bctbx_list_t *l, void *data) {
	return bctbx_list_append_link(l, bctbx_list_new(data));
}
bctbx_list_t *bctbx_list_prepend_link(bctbx_list_t *l, bctbx_list_t *n) {
	if (l) {
		n->next = l;
		l->prev = n;
	}
	return n;
}
bctbx_list_t *bctbx_list_prepend(bctbx_list_t *l, void *data) {
	return bctbx_list_prepend_link(l, bctbx_list_new(data));
}
bctbx_list_t *bctbx_list_concat(bctbx_list_t *a, bctbx_list_t *b) {
	return bctbx_list_append_link(a, b);
}
bctbx_list_t *bctbx_list_free_with_data(bctbx_list_t *l, bctbx_list_free_func fn) {
	while (l) {
		bctbx_list_t *n = l->next;
		if (fn) fn(l->data);
		free(l);
		l = n;
	}
	return NULL;
}
bctbx_list_t *bctbx_list_free(bctbx_list_t *l) {
	return bctbx_list_free_with_data(l, NULL);
}
bctbx_list_t *bctbx_list_unlink(bctbx_list_t *l, bctbx_list_t *e) {
	if (e->prev) e->prev->next = e->next;
	else l = e->next;
	if (e->next) e->next->prev = e->prev;
	e->next = e->prev = NULL;
	return l;
}
bctbx_list_t *bctbx_list_erase_link(bctbx_list_t *l, bctbx_list_t *e) {
	l = bctbx_list_unlink(l, e);
	free(e);
	return l;
}
bctbx_list_t *bctbx_list_remove_link(bctbx_list_t *l, bctbx_list_t *e) {
	return bctbx_list_unlink(l, e);
}
bctbx_list_t *bctbx_list_find(bctbx_list_t *l, const void *data) {
	for (; l; l = l->next)
		if (l->data == data) return l;
	return NULL;
}
bctbx_list_t *bctbx_list_remove(bctbx_list_t *l, void *data) {
	bctbx_list_t *e = bctbx_list_find(l, data);
	return e ? bctbx_list_erase_link(l, e) : l;
}
bctbx_list_t *bctbx_list_find_custom(const bctbx_list_t *l, bctbx_compare_func cmp, const void *user) {
	for (; l; l = l->next)
		if (cmp(l->data, user) == 0) return (bctbx_list_t *)l;
	return NULL;
}
bctbx_list_t *bctbx_list_remove_custom(bctbx_list_t *l, bctbx_compare_func cmp, const void *user) {
	bctbx_list_t *it = l, *n;
	while (it) {
		n = it->next;
		if (cmp(it->data, user) == 0) l = bctbx_list_erase_link(l, it);
		it = n;
	}
	return l;
}
bctbx_list_t *bctbx_list_insert(bctbx_list_t *l, bctbx_list_t *before, void *data) {
	bctbx_list_t *n;
	if (!before) return bctbx_list_append(l, data);
	n = bctbx_list_new(data);
	n->next = before;
	n->prev = before->prev;
	if (before->prev) before->prev->next = n;
	else l = n;
	before->prev = n;
	return l;
}
bctbx_list_t *bctbx_list_insert_sorted(bctbx_list_t *l, void *data, bctbx_compare_func cmp) {
	bctbx_list_t *it;
	for (it = l; it; it = it->next)
		if (cmp(data, it->data) <= 0) return bctbx_list_insert(l, it, data);
	return bctbx_list_append(l, data);
}
bctbx_list_t *bctbx_list_copy(const bctbx_list_t *l) {
	bctbx_list_t *out = NULL;
	for (; l; l = l->next)
		out = bctbx_list_append(out, l->data);
	return out;
}
bctbx_list_t *bctbx_list_copy_with_data(const bctbx_list_t *l, bctbx_list_copy_func fn) {
	bctbx_list_t *out = NULL;
	for (; l; l = l->next)
		out = bctbx_list_append(out, fn(l->data));
	return out;
}
void bctbx_list_for_each(const bctbx_list_t *l, bctbx_list_iterate_func fn) {
	while (l) {
		const bctbx_list_t *n = l->next;
		fn(l->data);
		l = n;
	}
}
void bctbx_list_for_each2(const bctbx_list_t *l, bctbx_list_iterate2_func fn, void *user) {
	while (l) {
		const bctbx_list_t *n = l->next;
		fn(l->data, user);
		l = n;
	}
}
size_t bctbx_list_size(const bctbx_list_t *l) {
	size_t n = 0;
	for (; l; l = l->next)
		n++;
	return n;
}
void *bctbx_list_nth_data(const bctbx_list_t *l, int n) {
	for (; l && n > 0; l = l->next)
		n--;
	return l ? l->data : NULL;
}
int bctbx_list_position(const bctbx_list_t *l, bctbx_list_t *e) {
	int i = 0;
	for (; l; l = l->next, i++)
		if (l == e) return i;
	return -1;
}
int bctbx_list_index(const bctbx_list_t *l, void *data) {
	int i = 0;
	for (; l; l = l->next, i++)
		if (l->data == data) return i;
	return -1;
}

/* ------------------------------------------------------------------ message blocks */
struct datab {
	unsigned char *db_base;
	unsigned char *db_lim;
	void (*db_freefn)(void *);
	int db_ref; /* atomically updated */
};
static dblk_t *datab_alloc(size_t size) {
	dblk_t *db = (dblk_t *)malloc(sizeof(dblk_t) + size + 16);
	db->db_base = (unsigned char *)(db + 1);
	db->db_lim = db->db_base + size;
	db->db_freefn = NULL;
	db->db_ref = 1;
	return db;
}
void dblk_ref(dblk_t *d) {
	__atomic_add_fetch(&d->db_ref, 1, __ATOMIC_SEQ_CST);
}
void dblk_unref(dblk_t *d) {
	if (__atomic_sub_fetch(&d->db_ref, 1, __ATOMIC_SEQ_CST) == 0) {
		if (d->db_freefn) d->db_freefn(d->db_base);
		free(d);
	}
}
unsigned char *dblk_base(dblk_t *db) {
	return db->db_base;
}
unsigned char *dblk_lim(dblk_t *db) {
	return db->db_lim;
}
int dblk_ref_value(dblk_t *db) {
	return __atomic_load_n(&db->db_ref, __ATOMIC_SEQ_CST);
}
void mblk_init(mblk_t *mp) {
	memset(mp, 0, sizeof(*mp));
}
void mblk_meta_copy(const mblk_t *source, mblk_t *dest) {
	dest->reserved1 = source->reserved1;
	dest->reserved2 = source->reserved2;
	dest->timestamp = source->timestamp;
	dest->ttl_or_hl = source->ttl_or_hl;
}
mblk_t *allocb(size_t size, int unused) {
	mblk_t *mp = (mblk_t *)calloc(1, sizeof(mblk_t));
	(void)unused;
	mp->b_datap = datab_alloc(size);
	mp->b_rptr = mp->b_wptr = mp->b_datap->db_base;
	return mp;
}
mblk_t *esballoc(uint8_t *buf, size_t size, int pri, void (*freefn)(void *)) {
	mblk_t *mp = (mblk_t *)calloc(1, sizeof(mblk_t));
	dblk_t *db = (dblk_t *)malloc(sizeof(dblk_t));
	(void)pri;
	db->db_base = buf;
	db->db_lim = buf + size;
	db->db_freefn = freefn;
	db->db_ref = 1;
	mp->b_datap = db;
	mp->b_rptr = mp->b_wptr = buf;
	return mp;
}
void freeb(mblk_t *m) {
	if (m->b_datap) dblk_unref(m->b_datap);
	free(m);
}
void freemsg(mblk_t *mp) {
	while (mp) {
		mblk_t *n = mp->b_cont;
		freeb(mp);
		mp = n;
	}
}
mblk_t *dupb(mblk_t *m) {
	mblk_t *n = (mblk_t *)calloc(1, sizeof(mblk_t));
	dblk_ref(m->b_datap);
	mblk_meta_copy(m, n);
	n->b_datap = m->b_datap;
	n->b_rptr = m->b_rptr;
	n->b_wptr = m->b_wptr;
	return n;
}
mblk_t *dupmsg(mblk_t *m) {
	mblk_t *head = dupb(m), *tail = head;
	for (m = m->b_cont; m; m = m->b_cont) {
		tail->b_cont = dupb(m);
		tail = tail->b_cont;
	}
	return head;
}
mblk_t *copyb(const mblk_t *mp) {
	size_t len = (size_t)(mp->b_wptr - mp->b_rptr);
	mblk_t *n = allocb(len, 0);
	memcpy(n->b_wptr, mp->b_rptr, len);
	n->b_wptr += len;
	mblk_meta_copy(mp, n);
	return n;
}
mblk_t *copymsg(const mblk_t *mp) {
	mblk_t *head = copyb(mp), *tail = head;
	for (mp = mp->b_cont; mp; mp = mp->b_cont) {
		tail->b_cont = copyb(mp);
		tail = tail->b_cont;
	}
	return head;
}
size_t msgdsize(const mblk_t *mp) {
	size_t n = 0;
	for (; mp; mp = mp->b_cont)
		n += (size_t)(mp->b_wptr - mp->b_rptr);
	return n;
}
void msgpullup(mblk_t *mp, size_t len) {
	size_t total = msgdsize(mp), wlen = 0;
	dblk_t *db;
	mblk_t *it;
	if (mp->b_cont == NULL && len == (size_t)-1) return;
	if (len == (size_t)-1 || len > total) len = total;
	db = datab_alloc(len);
	for (it = mp; it && wlen < len; it = it->b_cont) {
		size_t n = (size_t)(it->b_wptr - it->b_rptr);
		if (n > len - wlen) n = len - wlen;
		memcpy(db->db_base + wlen, it->b_rptr, n);
		wlen += n;
	}
	freemsg(mp->b_cont);
	mp->b_cont = NULL;
	dblk_unref(mp->b_datap);
	mp->b_datap = db;
	mp->b_rptr = db->db_base;
	mp->b_wptr = db->db_base + wlen;
}
mblk_t *concatb(mblk_t *mp, mblk_t *newm) {
	while (mp->b_cont)
		mp = mp->b_cont;
	mp->b_cont = newm;
	while (newm->b_cont)
		newm = newm->b_cont;
	return newm;
}
void qinit(queue_t *q) {
	mblk_init(&q->_q_stopper);
	q->_q_stopper.b_next = &q->_q_stopper;
	q->_q_stopper.b_prev = &q->_q_stopper;
	q->q_mcount = 0;
}
void insq(queue_t *q, mblk_t *emp, mblk_t *mp) {
	if (emp == NULL) emp = &q->_q_stopper;
	q->q_mcount++;
	mp->b_next = emp;
	mp->b_prev = emp->b_prev;
	emp->b_prev->b_next = mp;
	emp->b_prev = mp;
}
void putq(queue_t *q, mblk_t *m) {
	insq(q, NULL, m);
}
void remq(queue_t *q, mblk_t *mp) {
	q->q_mcount--;
	mp->b_prev->b_next = mp->b_next;
	mp->b_next->b_prev = mp->b_prev;
	mp->b_next = mp->b_prev = NULL;
}
mblk_t *getq(queue_t *q) {
	mblk_t *m = q->_q_stopper.b_next;
	if (m == &q->_q_stopper) return NULL;
	remq(q, m);
	return m;
}
mblk_t *peekq(queue_t *q) {
	mblk_t *m = q->_q_stopper.b_next;
	return m == &q->_q_stopper ? NULL : m;
}
void flushq(queue_t *q, int how) {
	mblk_t *m;
	(void)how;
	while ((m = getq(q)) != NULL)
		freemsg(m);
}
void msgb_allocator_init(msgb_allocator_t *pa) {
	qinit(&pa->q);
	pa->max_blocks = 0;
}
void msgb_allocator_set_max_blocks(msgb_allocator_t *pa, int max_blocks) {
	pa->max_blocks = max_blocks;
}
mblk_t *msgb_allocator_alloc(msgb_allocator_t *pa, size_t size) {
	queue_t *q = &pa->q;
	mblk_t *m, *found = NULL;
	int busy = 0;
	for (m = qbegin(q); !qend(q, m); m = qnext(q, m)) {
		if ((size_t)(m->b_datap->db_lim - m->b_datap->db_base) >= size) {
			if (dblk_ref_value(m->b_datap) == 1) {
				found = m;
				break;
			}
			busy++;
		}
	}
	if (pa->max_blocks != 0 && busy >= pa->max_blocks) return NULL;
	if (!found) {
		found = allocb(size, 0);
		putq(q, found);
	}
	return dupb(found);
}
void msgb_allocator_uninit(msgb_allocator_t *pa) {
	flushq(&pa->q, -1);
}

/* ------------------------------------------------------------------ OrtpExtremum */
void ortp_extremum_reset(OrtpExtremum *obj) {
	obj->current_extremum = 0;
	obj->extremum_time = (uint64_t)-1;
	obj->last_stable = 0;
}
void ortp_extremum_init(OrtpExtremum *obj, int period) {
	ortp_extremum_reset(obj);
	obj->period = period;
}
static bool_t extremum_roll(OrtpExtremum *obj, uint64_t curtime, float value) {
	if (obj->extremum_time != (uint64_t)-1) {
		if ((int)(curtime - obj->extremum_time) > obj->period) {
			obj->last_stable = obj->current_extremum;
			obj->extremum_time = curtime;
			obj->current_extremum = value;
			return TRUE;
		}
		return FALSE;
	}
	obj->last_stable = value;
	obj->current_extremum = value;
	obj->extremum_time = curtime;
	return TRUE;
}
bool_t ortp_extremum_record_min(OrtpExtremum *obj, uint64_t curtime, float value) {
	bool_t ret = extremum_roll(obj, curtime, value);
	if (value < obj->current_extremum) {
		obj->current_extremum = value;
		obj->extremum_time = curtime;
		ret = TRUE;
	}
	return ret;
}
bool_t ortp_extremum_record_max(OrtpExtremum *obj, uint64_t curtime, float value) {
	bool_t ret = extremum_roll(obj, curtime, value);
	if (value > obj->current_extremum) {
		obj->current_extremum = value;
		obj->extremum_time = curtime;
		ret = TRUE;
	}
	return ret;
}
float ortp_extremum_get_current(OrtpExtremum *obj) {
	return obj->current_extremum;
}
float ortp_extremum_get_previous(OrtpExtremum *obj) {
	return obj->last_stable;
}

/* ---------------------------------------------------------------- fmtp parameters (used by the G.711 encoders' ptime) */
bool_t fmtp_get_value(const char *fmtp, const char *param_name, char *result, size_t result_len) {
	const size_t klen = strlen(param_name);
	const char *p = fmtp;
	if (result_len == 0) return FALSE;
	while (p && *p) {
		const char *end = strchr(p, ';');
		const size_t len = end ? (size_t)(end - p) : strlen(p);
		const char *k = p;
		size_t n = len;
		while (n && (*k == ' ' || *k == '\t')) ++k, --n;
		if (n > klen && strncasecmp(k, param_name, klen) == 0 && k[klen] == '=') {
			size_t vlen = n - klen - 1;
			if (vlen >= result_len) vlen = result_len - 1;
			memcpy(result, k + klen + 1, vlen);
			result[vlen] = 0;
			return TRUE;
		}
		p = end ? end + 1 : NULL;
	}
	return FALSE;
}

/* ---------------------------------------------------------------- bctoolbox/vfs.h over POSIX descriptors
 * (the reference's WAV reader msfileplayer.c:98-150 and utils/audiodiff.c read their files through it) */
#include "bctoolbox/vfs.h"
#include <sys/stat.h>
#include <unistd.h>
struct bctbx_vfs_t {
	int unused;
};
static bctbx_vfs_t g_default_vfs;
bctbx_vfs_t *bctbx_vfs_get_default(void) {
	return &g_default_vfs;
}
bctbx_vfs_file_t *bctbx_file_open2(bctbx_vfs_t *vfs, const char *path, int openflags) {
	(void)vfs;
	int fd = open(path, openflags, 0644);
	if (fd < 0) return NULL;
	bctbx_vfs_file_t *f = (bctbx_vfs_file_t *)calloc(1, sizeof(*f));
	f->fd = fd;
	f->offset = 0;
	return f;
}
bctbx_vfs_file_t *bctbx_file_open(bctbx_vfs_t *vfs, const char *path, const char *mode) {
	int flags = O_RDONLY;
	if (mode && strchr(mode, 'w')) flags = O_WRONLY | O_CREAT | O_TRUNC;
	if (mode && strchr(mode, '+')) flags = O_RDWR | O_CREAT;
	return bctbx_file_open2(vfs, path, flags);
}
int64_t bctbx_file_size(bctbx_vfs_file_t *f) {
	struct stat st;
	if (!f || fstat(f->fd, &st) != 0) return BCTBX_VFS_ERROR;
	return (int64_t)st.st_size;
}
int bctbx_file_close(bctbx_vfs_file_t *f) {
	if (!f) return BCTBX_VFS_ERROR;
	close(f->fd);
	free(f);
	return BCTBX_VFS_OK;
}
ssize_t bctbx_file_read(bctbx_vfs_file_t *f, void *buf, size_t count, off_t offset) {
	if (!f) return BCTBX_VFS_ERROR;
	ssize_t r = pread(f->fd, buf, count, offset);
	return r < 0 ? BCTBX_VFS_ERROR : r;
}
ssize_t bctbx_file_read2(bctbx_vfs_file_t *f, void *buf, size_t count) {
	ssize_t r = bctbx_file_read(f, buf, count, f ? f->offset : 0);
	if (r > 0) f->offset += r;
	return r;
}
ssize_t bctbx_file_write(bctbx_vfs_file_t *f, const void *buf, size_t count, off_t offset) {
	if (!f) return BCTBX_VFS_ERROR;
	ssize_t r = pwrite(f->fd, buf, count, offset);
	return r < 0 ? BCTBX_VFS_ERROR : r;
}
ssize_t bctbx_file_write2(bctbx_vfs_file_t *f, const void *buf, size_t count) {
	ssize_t r = bctbx_file_write(f, buf, count, f ? f->offset : 0);
	if (r > 0) f->offset += r;
	return r;
}
off_t bctbx_file_seek(bctbx_vfs_file_t *f, off_t offset, int whence) {
	if (!f) return BCTBX_VFS_ERROR;
	if (whence == SEEK_SET) f->offset = offset;
	else if (whence == SEEK_CUR) f->offset += offset;
	else if (whence == SEEK_END) f->offset = (off_t)bctbx_file_size(f) + offset;
	return f->offset;
}
int bctbx_file_truncate(bctbx_vfs_file_t *f, int64_t size) {
	return (f && ftruncate(f->fd, size) == 0) ? BCTBX_VFS_OK : BCTBX_VFS_ERROR;
}
int bctbx_file_sync(bctbx_vfs_file_t *f) {
	return (f && fsync(f->fd) == 0) ? BCTBX_VFS_OK : BCTBX_VFS_ERROR;
}
