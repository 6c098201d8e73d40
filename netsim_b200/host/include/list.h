/* list.h — circular doubly linked lists embedded in their elements, the subset of the kernel-style API the
 * behaviour modules use (list.h there: INIT_LIST_HEAD, list_add = insert right after the head, list_del, list_empty,
 * list_entry, list_for_each_entry[_reverse]).  Own implementation for the host core's mirrors. */
#pragma once
#include "common.h"

struct list_head { struct list_head *next, *prev; };

#define LIST_HEAD_INIT(name) { &(name), &(name) }
#define LIST_HEAD(name) struct list_head name = LIST_HEAD_INIT(name)

static inline void INIT_LIST_HEAD(struct list_head *h){ h->next = h; h->prev = h; }
static inline void ns_list_link(struct list_head *n, struct list_head *before, struct list_head *after){
	n->prev = before; n->next = after; before->next = n; after->prev = n;
}
static inline void list_add(struct list_head *n, struct list_head *head){ ns_list_link(n, head, head->next); }        /* a stack: newest first */
static inline void list_add_tail(struct list_head *n, struct list_head *head){ ns_list_link(n, head->prev, head); }
static inline void list_del(struct list_head *e){
	e->prev->next = e->next; e->next->prev = e->prev;
	e->next = (struct list_head *)0x00100100; e->prev = (struct list_head *)0x00200200;   /* non-NULL after removal, like the reference's poison values */
}
static inline int list_empty(const struct list_head *h){ return h->next == h; }

#define list_entry(ptr, type, member) container_of(ptr, type, member)
#define list_first_entry(head, type, member) list_entry((head)->next, type, member)
#define list_for_each(pos, head) for (pos = (head)->next; pos != (head); pos = pos->next)
#define list_for_each_entry(pos, head, member) \
	for (pos = list_entry((head)->next, __typeof__(*pos), member); &pos->member != (head); \
	     pos = list_entry(pos->member.next, __typeof__(*pos), member))
#define list_for_each_entry_reverse(pos, head, member) \
	for (pos = list_entry((head)->prev, __typeof__(*pos), member); &pos->member != (head); \
	     pos = list_entry(pos->member.prev, __typeof__(*pos), member))
#define list_for_each_entry_safe(pos, n, head, member) \
	for (pos = list_entry((head)->next, __typeof__(*pos), member), n = list_entry(pos->member.next, __typeof__(*pos), member); \
	     &pos->member != (head); pos = n, n = list_entry(n->member.next, __typeof__(*n), member))
