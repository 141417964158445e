/* compat shim (our own code): STREAMS-like message blocks with the names and semantics oRTP exposes
 * (allocb/dupb/dupmsg/freemsg with a ref-counted data block; queue_t with a stopper node).
 * The field layout here is OURS; product code never hard-codes it (it is compiled against whichever
 * <ortp/str_utils.h> the host provides). */
#ifndef MSB200_COMPAT_ORTP_STR_UTILS_H
#define MSB200_COMPAT_ORTP_STR_UTILS_H
#include "ortp/port.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct datab dblk_t;
typedef struct msgb {
	struct msgb *b_prev;
	struct msgb *b_next;
	struct msgb *b_cont;
	dblk_t *b_datap;
	unsigned char *b_rptr;
	unsigned char *b_wptr;
	uint32_t reserved1;
	uint32_t reserved2;
	struct timeval timestamp;
	uint8_t ttl_or_hl;
} mblk_t;
typedef struct _queue {
	mblk_t _q_stopper;
	int q_mcount;
} queue_t;
#define BPRI_MED 0
void dblk_ref(dblk_t *d);
void dblk_unref(dblk_t *d);
unsigned char *dblk_base(dblk_t *db);
unsigned char *dblk_lim(dblk_t *db);
int dblk_ref_value(dblk_t *db);
void qinit(queue_t *q);
void putq(queue_t *q, mblk_t *m);
mblk_t *getq(queue_t *q);
void insq(queue_t *q, mblk_t *emp, mblk_t *mp);
void remq(queue_t *q, mblk_t *mp);
mblk_t *peekq(queue_t *q);
void flushq(queue_t *q, int how);
#define FLUSHALL 0
void mblk_init(mblk_t *mp);
void mblk_meta_copy(const mblk_t *source, mblk_t *dest);
mblk_t *allocb(size_t size, int unused);
mblk_t *esballoc(uint8_t *buf, size_t size, int pri, void (*freefn)(void *));
void freeb(mblk_t *m);
void freemsg(mblk_t *mp);
mblk_t *dupb(mblk_t *m);
mblk_t *dupmsg(mblk_t *m);
mblk_t *copyb(const mblk_t *mp);
mblk_t *copymsg(const mblk_t *mp);
size_t msgdsize(const mblk_t *mp);
void msgpullup(mblk_t *mp, size_t len);
mblk_t *concatb(mblk_t *mp, mblk_t *newm);
#define qempty(q) (&(q)->_q_stopper == (q)->_q_stopper.b_next)
#define qfirst(q) ((q)->_q_stopper.b_next != &(q)->_q_stopper ? (q)->_q_stopper.b_next : NULL)
#define qbegin(q) ((q)->_q_stopper.b_next)
#define qlast(q) ((q)->_q_stopper.b_prev != &(q)->_q_stopper ? (q)->_q_stopper.b_prev : NULL)
#define qend(q, mp) ((mp) == &(q)->_q_stopper)
#define qnext(q, mp) ((mp)->b_next)
typedef struct _msgb_allocator {
	queue_t q;
	int max_blocks;
} msgb_allocator_t;
void msgb_allocator_init(msgb_allocator_t *pa);
void msgb_allocator_set_max_blocks(msgb_allocator_t *pa, int max_blocks);
mblk_t *msgb_allocator_alloc(msgb_allocator_t *pa, size_t size);
void msgb_allocator_uninit(msgb_allocator_t *pa);
#ifdef __cplusplus
}
#endif
#endif
