/* flow.h — what modules see of the rate model (flow.h:5-41 there). */
#pragma once
#include <assert.h>
#include "data.h"

struct spd_event {
	struct event *e;
	struct flow *f;
	double speed;
	enum peer_type type;
	struct list_head spd_evs;
};

void flow_close(struct flow *f, struct sim_state *s);
/* flow.h:29-35: a probe of src->outs, keyed by the destination id */
static inline bool is_connected(struct node *src, struct node *dst){
	struct flow *tf = NULL;
	HASH_FIND(hh2, src->outs, &dst->node_id, sizeof(dst->node_id), tf);
	return tf != NULL;
}
static inline double get_share(struct flow *f, int dir){
	struct node *n = f->peer[dir];
	double total = n->total_bwupbound[dir], max = n->maximum_bandwidth[dir];
	return total > max ? f->bwupbound*max/total : f->bwupbound;
}
