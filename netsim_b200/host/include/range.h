/* range.h — range lookups and the lazy progress update on the mirrors (range.h:10-125 there).  The ranges are advanced
 * on the device; a module that advances one itself (_range_update from a handler) changes the same fp64 state the
 * device would have produced later, and the host core writes it back (NS_OP_RANGE_SET). */
#pragma once
#include "data.h"
#include "store.h"
#include "event.h"

void range_calc_and_requeue_events(struct flow *f, struct sim_state *s);
void range_merge_with_next(struct range *rng, struct sim_state *s);

/* the key is an int in every caller (range.h:10-27) */
static inline int range_include_cmp(struct skip_list_head *a, void *_key){
	struct range *rng = skip_list_entry(a, struct range, ranges);
	int key = *(int *)_key;
	if (rng->start <= key) return rng->start+rng->len >= key ? 0 : -1;
	return 1;
}
static inline int range_start_cmp(struct skip_list_head *a, void *_key){
	struct range *rng = skip_list_entry(a, struct range, ranges);
	return rng->start-*(int *)_key;
}
/* the range whose [start, start+len] holds `start` */
static inline struct range *range_get(struct resource *rsrc, int start){
	struct skip_list_head *r = skip_list_find_ge(&rsrc->ranges, &start, range_include_cmp);
	struct range *rng = skip_list_entry(r, struct range, ranges);
	if (!r || rng->start > start) return NULL;
	return rng;
}
static inline struct range *range_get_by_start(struct resource *rsrc, int start){
	struct skip_list_head *r = skip_list_find_le(&rsrc->ranges, &start, range_start_cmp);
	return r ? skip_list_entry(r, struct range, ranges) : NULL;
}
static inline struct range *range_get_next(struct resource *rsrc, int start){
	struct skip_list_head *r = skip_list_find_ge(&rsrc->ranges, &start, range_start_cmp);
	return r ? skip_list_entry(r, struct range, ranges) : NULL;
}
/* Kahan-compensated len += speed*dt, range.h:91-106 */
static inline void _range_update(struct range *r, struct sim_state *s){
	if (s->now == r->last_update) return;
	if (r->producer == NULL) { r->last_update = s->now; return; }
	double delta = r->producer->speed[1]*(s->now-r->last_update), t;
	delta = delta-r->lenc;
	t = delta+r->len;
	r->lenc = (t-r->len)-delta;
	r->len = t;
	r->last_update = s->now;
}
static inline void flow_range_update(struct flow *f, struct sim_state *s){
	if (!f) return;
	_range_update(f->srng, s);
	_range_update(f->drng, s);
}
