/* sim.h — the module-facing API (sim.h:14-53 there), implemented over include/netsim_b200.h by
 * netsim_b200/host/core.c. */
#pragma once
#include "data.h"
#include "log.h"
#include "record.h"
#include "range.h"

struct packet { struct node *src, *dst; void *data; int len; };

struct node *sim_create_node(struct sim_state *s);
void sim_register_handler(int type, int priority, event_handler_func f, struct sim_state *s);
struct flow *sim_establish_flow(id_t rid, size_t start, struct node *src, struct node *dst, struct sim_state *s);
void sim_send_packet(void *data, int len, struct node *src, struct node *dst, struct sim_state *s);
struct resource *sim_node_new_resource(struct node *n, size_t len);
struct resource *sim_node_add_resource(struct node *n, struct resource *r);
void ns_host_refresh(struct sim_state *s);

static inline void sim_node_change_state(struct node *n, enum node_state state, struct sim_state *s){   /* sim.h:25-28 */
	n->state = state;
	write_record(0, R_NODE_STATE, n->node_id, 1, &n->state, s);
}
/* print_range, sim.h:29-40 there: the log line carries the including module's LOG_DOMAIN */
static inline void print_range(struct resource *r){
	for (struct skip_list_head *h = r->ranges.next[0]; h; h = h->next[0]) {
		struct range *rng = skip_list_entry(h, struct range, ranges);
		log_info("Range [%zu, %lf)\n", rng->start, rng->start+rng->len);
	}
}
static inline void sim_end(struct sim_state *s){ s->exit = true; }
void node_del_resource(struct resource *r);      /* sim.h:47-53 there */
