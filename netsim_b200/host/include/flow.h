/* flow.h — what modules see of the rate model (flow.h:5-41 there). */
#pragma once
#include "data.h"

struct spd_event {
	struct event *e;
	struct flow *f;
	double speed;
	enum peer_type type;
};

void flow_close(struct flow *f, struct sim_state *s);
bool is_connected(struct node *src, struct node *dst);
static inline double get_share(struct flow *f, int dir){
	struct node *n = f->peer[dir];
	double total = n->total_bwupbound[dir], max = n->maximum_bandwidth[dir];
	return total > max ? f->bwupbound*max/total : f->bwupbound;
}
