/* compat shim (our own code): doubly linked list with the bctbx_list_* names the reference runtime calls. */
#ifndef MSB200_COMPAT_BCTBX_LIST_H
#define MSB200_COMPAT_BCTBX_LIST_H
#include "bctoolbox/defs.h"
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct _bctbx_list {
	struct _bctbx_list *next;
	struct _bctbx_list *prev;
	void *data;
} bctbx_list_t;
typedef int (*bctbx_compare_func)(const void *, const void *);
typedef void (*bctbx_list_iterate_func)(void *);
typedef void (*bctbx_list_iterate2_func)(void *, void *);
typedef void (*bctbx_list_free_func)(void *);
typedef void *(*bctbx_list_copy_func)(void *);

bctbx_list_t *bctbx_list_new(void *data);
bctbx_list_t *bctbx_list_append(bctbx_list_t *l, void *data);
bctbx_list_t *bctbx_list_append_link(bctbx_list_t *l, bctbx_list_t *n);
bctbx_list_t *bctbx_list_prepend(bctbx_list_t *l, void *data);
bctbx_list_t *bctbx_list_prepend_link(bctbx_list_t *l, bctbx_list_t *n);
bctbx_list_t *bctbx_list_concat(bctbx_list_t *a, bctbx_list_t *b);
bctbx_list_t *bctbx_list_free(bctbx_list_t *l);
bctbx_list_t *bctbx_list_free_with_data(bctbx_list_t *l, bctbx_list_free_func fn);
bctbx_list_t *bctbx_list_remove(bctbx_list_t *l, void *data);
bctbx_list_t *bctbx_list_remove_custom(bctbx_list_t *l, bctbx_compare_func cmp, const void *user);
bctbx_list_t *bctbx_list_unlink(bctbx_list_t *l, bctbx_list_t *e);
bctbx_list_t *bctbx_list_erase_link(bctbx_list_t *l, bctbx_list_t *e);
bctbx_list_t *bctbx_list_remove_link(bctbx_list_t *l, bctbx_list_t *e);
bctbx_list_t *bctbx_list_find(bctbx_list_t *l, const void *data);
bctbx_list_t *bctbx_list_find_custom(const bctbx_list_t *l, bctbx_compare_func cmp, const void *user);
bctbx_list_t *bctbx_list_insert_sorted(bctbx_list_t *l, void *data, bctbx_compare_func cmp);
bctbx_list_t *bctbx_list_insert(bctbx_list_t *l, bctbx_list_t *before, void *data);
bctbx_list_t *bctbx_list_copy(const bctbx_list_t *l);
bctbx_list_t *bctbx_list_copy_with_data(const bctbx_list_t *l, bctbx_list_copy_func fn);
void bctbx_list_for_each(const bctbx_list_t *l, bctbx_list_iterate_func fn);
void bctbx_list_for_each2(const bctbx_list_t *l, bctbx_list_iterate2_func fn, void *user);
size_t bctbx_list_size(const bctbx_list_t *l);
void *bctbx_list_nth_data(const bctbx_list_t *l, int n);
int bctbx_list_position(const bctbx_list_t *l, bctbx_list_t *e);
int bctbx_list_index(const bctbx_list_t *l, void *data);
static inline bctbx_list_t *bctbx_list_next(const bctbx_list_t *e) { return e->next; }
static inline void *bctbx_list_get_data(const bctbx_list_t *e) { return e->data; }
#ifdef __cplusplus
}
#endif
#endif
