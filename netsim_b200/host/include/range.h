/* range.h — range lookups on the mirrors (range.h:30-63 there); the ranges themselves are advanced
 * on the device. */
#pragma once
#include "data.h"
#include "store.h"
#include "event.h"

void ns_host_refresh_ranges(struct resource *r);
static inline struct range *range_get(struct resource *rsrc, int start){
	ns_host_refresh_ranges(rsrc);
	for (struct range *g = rsrc->first; g; g = g->next)
		if ((double)g->start+g->len >= (double)start) return g->start <= (size_t)start ? g : NULL;
	return NULL;
}
