/* uthash.h — intrusive hash tables with the macro interface of Troy Hanson's uthash as the behaviour modules use it
 * (HASH_FIND_INT / HASH_ADD_INT / HASH_DEL / HASH_ITER / HASH_COUNT, plus the generic HASH_FIND / HASH_ADD with a named
 * handle).  Own implementation for the host core: chained buckets for lookup and a doubly linked list in INSERTION ORDER
 * for iteration — the order HASH_ITER walks in is part of the contract (server_picker_opt2 takes the first minimum over
 * the holders in that order, client_behaviour.h:90-124). */
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

struct UT_hash_table;
typedef struct UT_hash_handle {
	struct UT_hash_table *tbl;
	void *prev, *next;                 /* neighbours in insertion order (element pointers) */
	struct UT_hash_handle *chain;      /* next handle in the same bucket */
	const void *key; unsigned keylen, hashv;
} UT_hash_handle;
typedef struct UT_hash_table {
	UT_hash_handle **buckets; unsigned n_buckets, count;
	void *tail;                        /* last element in insertion order */
	ptrdiff_t hho;                     /* offset of the handle inside an element */
} UT_hash_table;

static inline unsigned ns_hash_bytes(const void *key, unsigned len){   /* FNV-1a */
	const unsigned char *p = (const unsigned char *)key; unsigned h = 2166136261u;
	for (unsigned i = 0; i < len; i++) { h ^= p[i]; h *= 16777619u; }
	return h;
}
#define NS_HH(tbl, elt) ((UT_hash_handle *)((char *)(elt)+(tbl)->hho))
static inline void *ns_hash_find(UT_hash_table *t, const void *key, unsigned len){
	if (!t) return NULL;
	unsigned hv = ns_hash_bytes(key, len);
	for (UT_hash_handle *h = t->buckets[hv & (t->n_buckets-1)]; h; h = h->chain)
		if (h->hashv == hv && h->keylen == len && memcmp(h->key, key, len) == 0) return (char *)h-t->hho;
	return NULL;
}
static inline void ns_hash_grow(UT_hash_table *t){
	unsigned nb = t->n_buckets*2;
	UT_hash_handle **b = (UT_hash_handle **)calloc(nb, sizeof *b);
	for (unsigned i = 0; i < t->n_buckets; i++)
		for (UT_hash_handle *h = t->buckets[i], *nx; h; h = nx) { nx = h->chain; h->chain = b[h->hashv & (nb-1)]; b[h->hashv & (nb-1)] = h; }
	free(t->buckets); t->buckets = b; t->n_buckets = nb;
}
/* returns the table (created on the first add); *head keeps pointing at the first element */
static inline void ns_hash_add(void **head, void *elt, UT_hash_handle *hh, const void *key, unsigned len){
	UT_hash_table *t;
	if (!*head) {
		t = (UT_hash_table *)calloc(1, sizeof *t);
		t->n_buckets = 32; t->buckets = (UT_hash_handle **)calloc(t->n_buckets, sizeof *t->buckets);
		t->hho = (char *)hh-(char *)elt;
		*head = elt; hh->prev = NULL;
	} else {
		t = ((UT_hash_handle *)((char *)*head+((char *)hh-(char *)elt)))->tbl;
		hh->prev = t->tail; NS_HH(t, t->tail)->next = elt;
	}
	hh->tbl = t; hh->next = NULL; t->tail = elt;
	hh->key = key; hh->keylen = len; hh->hashv = ns_hash_bytes(key, len);
	hh->chain = t->buckets[hh->hashv & (t->n_buckets-1)]; t->buckets[hh->hashv & (t->n_buckets-1)] = hh;
	if (++t->count > 4*t->n_buckets) ns_hash_grow(t);
}
static inline void ns_hash_del(void **head, void *elt, UT_hash_handle *hh){
	UT_hash_table *t = hh->tbl;
	UT_hash_handle **pp = &t->buckets[hh->hashv & (t->n_buckets-1)];
	while (*pp && *pp != hh) pp = &(*pp)->chain;
	if (*pp) *pp = hh->chain;
	if (hh->prev) NS_HH(t, hh->prev)->next = hh->next; else *head = hh->next;
	if (hh->next) NS_HH(t, hh->next)->prev = hh->prev; else t->tail = hh->prev;
	if (--t->count == 0) { free(t->buckets); free(t); *head = NULL; }
	hh->tbl = NULL;
}

#define HASH_FIND(hh, head, keyptr, keylen_in, out) \
	do { (out) = (head) ? (__typeof__(out))ns_hash_find((head)->hh.tbl, (keyptr), (unsigned)(keylen_in)) : NULL; } while (0)
#define HASH_ADD_KEYPTR(hh, head, keyptr, keylen_in, add) \
	do { void *ns_h_ = (void *)(head); ns_hash_add(&ns_h_, (add), &(add)->hh, (keyptr), (unsigned)(keylen_in)); (head) = (__typeof__(head))ns_h_; } while (0)
#define HASH_ADD(hh, head, fieldname, keylen_in, add) HASH_ADD_KEYPTR(hh, head, &((add)->fieldname), keylen_in, add)
#define HASH_DELETE(hh, head, delptr) \
	do { void *ns_h_ = (void *)(head); ns_hash_del(&ns_h_, (delptr), &(delptr)->hh); (head) = (__typeof__(head))ns_h_; } while (0)
#define HASH_FIND_INT(head, findint, out) HASH_FIND(hh, head, findint, sizeof(int), out)
#define HASH_ADD_INT(head, intfield, add) HASH_ADD(hh, head, intfield, sizeof(int), add)
#define HASH_DEL(head, delptr) HASH_DELETE(hh, head, delptr)
#define HASH_CNT(hh, head) ((head) ? (head)->hh.tbl->count : 0u)
#define HASH_COUNT(head) HASH_CNT(hh, head)
/* el walks the elements in insertion order; tmp always holds the successor, so el may be deleted inside the loop */
#define HASH_ITER(hh, head, el, tmp) \
	for ((el) = (head), (tmp) = (el) ? (__typeof__(el))(el)->hh.next : NULL; (el); (el) = (tmp), (tmp) = (el) ? (__typeof__(el))(el)->hh.next : NULL)
