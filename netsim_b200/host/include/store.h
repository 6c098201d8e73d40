/* store.h — per-node resource lookup (store.h:5-10 there). */
#pragma once
#include "data.h"

static inline struct resource *store_get(struct store *s, id_t resource_id){
	for (struct resource *r = s ? s->head : NULL; r; r = r->next_in_store)
		if (r->resource_id == resource_id) return r;
	return NULL;
}
