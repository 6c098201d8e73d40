/* skiplist.h — ordered lists with express lanes, the API of the reference header of this name (skiplist.h:8-158): a node
 * carries next[]/prev[] arrays of its own height, level 0 links every node (modules read ->next[0] / ->prev[0] directly),
 * a new node's height comes from random() (one draw per insert: skiplist.h:34-40 — the simulation's random() stream is
 * shared with the module, so the draw is part of the contract).  Own implementation.  The host core's range mirrors use
 * ns_skip_link_after() instead of skip_list_insert(): their draws are burned on the device. */
#pragma once
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include "common.h"

#define MAX_HEIGHT 32

struct skip_list_head { struct skip_list_head **next, **prev; int h; };
typedef int (*skip_list_cmp)(struct skip_list_head *a, void *key);

#define skip_list_entry(ptr, type, member) container_of(ptr, type, member)
#define skip_list_empty(ptr) ((ptr)->next == NULL || (ptr)->next[0] == NULL)

static inline void skip_list_init_head(struct skip_list_head *h){
	h->h = MAX_HEIGHT; h->next = (struct skip_list_head **)calloc(MAX_HEIGHT, sizeof(void *)); h->prev = NULL;
}
static inline void skip_list_deinit_head(struct skip_list_head *h){ h->h = 0; free(h->next); h->next = NULL; }
/* geometric height: one more level per trailing 1 bit of ONE random() draw */
static inline int skip_list_gen_height(void){
	int r = (int)random(), h = 1;
	while (r & 1) { h++; r >>= 1; }
	return h;
}
static inline void ns_skip_node_alloc(struct skip_list_head *n, int height){
	n->next = (struct skip_list_head **)calloc((size_t)height, sizeof(void *));
	n->prev = (struct skip_list_head **)calloc((size_t)height, sizeof(void *));
	n->h = height;
}
static inline void skip_list_init_node(struct skip_list_head *n){ ns_skip_node_alloc(n, skip_list_gen_height()); }
static inline void skip_list_deinit_node(struct skip_list_head *n){ free(n->next); free(n->prev); n->next = n->prev = NULL; n->h = 0; }

/* per level, the last node that orders before key (the head if none) */
static inline void skip_list_previous(struct skip_list_head *head, void *key, skip_list_cmp cmp, struct skip_list_head **res){
	struct skip_list_head *n = head;
	for (int lv = head->h-1; lv >= 0; lv--) {
		while (n->next[lv] && cmp(n->next[lv], key) < 0) n = n->next[lv];
		res[lv] = n;
	}
}
static inline void ns_skip_splice(struct skip_list_head *n, struct skip_list_head **before){
	for (int lv = 0; lv < n->h; lv++) {
		n->next[lv] = before[lv]->next[lv]; n->prev[lv] = before[lv];
		before[lv]->next[lv] = n;
		if (n->next[lv]) n->next[lv]->prev[lv] = n;
	}
}
static inline void skip_list_insert(struct skip_list_head *h, struct skip_list_head *n, void *key, skip_list_cmp cmp){
	struct skip_list_head *before[MAX_HEIGHT];
	skip_list_init_node(n);
	skip_list_previous(h, key, cmp, before);
	ns_skip_splice(n, before);
}
/* a height-1 node right after `after` (a node or the head): no random() draw */
static inline void ns_skip_link_after(struct skip_list_head *n, struct skip_list_head *after){
	ns_skip_node_alloc(n, 1);
	n->next[0] = after->next[0]; n->prev[0] = after;
	after->next[0] = n;
	if (n->next[0]) n->next[0]->prev[0] = n;
}
/* smallest element >= key */
static inline struct skip_list_head *skip_list_find_ge(struct skip_list_head *h, void *key, skip_list_cmp cmp){
	struct skip_list_head *before[MAX_HEIGHT];
	before[0] = h;
	skip_list_previous(h, key, cmp, before);
	return before[0]->next[0];
}
/* largest element <= key */
static inline struct skip_list_head *skip_list_find_le(struct skip_list_head *h, void *key, skip_list_cmp cmp){
	struct skip_list_head *before[MAX_HEIGHT];
	before[0] = h;
	skip_list_previous(h, key, cmp, before);
	if (before[0]->next[0] && cmp(before[0]->next[0], key) == 0) return before[0]->next[0];
	return before[0] == h ? NULL : before[0];
}
static inline void skip_list_delete(struct skip_list_head *n){
	for (int lv = 0; lv < n->h; lv++) {
		n->prev[lv]->next[lv] = n->next[lv];
		if (n->next[lv]) n->next[lv]->prev[lv] = n->prev[lv];
	}
	skip_list_deinit_node(n);
}
static inline struct skip_list_head *skip_list_extract_by_key(struct skip_list_head *h, void *key, skip_list_cmp cmp){
	struct skip_list_head *n = skip_list_find_ge(h, key, cmp);
	if (!n || cmp(n, key) != 0) return NULL;
	skip_list_delete(n);
	return n;
}
static inline void skip_list_clear(struct skip_list_head *h, void (*freep)(struct skip_list_head *)){
	struct skip_list_head *n = h->next[0];
	while (n) { struct skip_list_head *nx = n->next[0]; freep(n); n = nx; }
}
