/* modx.c — a behaviour module written against netsim's module API only (sim.h, event.h, record.h, flow.h), compiled twice by
 * tests/test_host_dropin.py: against the reference's headers (run by oracle/_ref/netsim) and against
 * netsim_b200/host/include (run by netsim_dev on the GPU).  Both runs must leave the same stderr and the same record file.
 * It exercises what the reference's own test modules do not: more than 16 flows, sim_node_change_state (R_NODE_STATE
 * records interleaved with the core's), event_remove of a module-owned event, a module's own write_record, is_connected,
 * flow_close from a handler, handlers on three event types. */
#define LOG_DOMAIN "modx"
#include <stdio.h>
#include "log.h"
#include "data.h"
#include "sim.h"
#include "event.h"
#include "flow.h"
#include "record.h"
#include "store.h"

#define NCL 24
static struct node *srv[2], *cl[NCL];
static id_t rid;
static struct event *pending[NCL];
static struct flow *fl[NCL];
static int n_done;

static double bw(void *a, void *b){ return 600; }
static double dl(void *a, void *b){ return 0.25; }

struct ue { int kind, i; };

static void on_user(struct event *e, struct sim_state *s){
	struct ue *u = e->data;
	if (u->kind == 0) {                         /* client i starts: from server i%2, state record, a cancellable follow-up */
		int i = u->i;
		if (!is_connected(srv[i%2], cl[i])) {
			fl[i] = sim_establish_flow(rid, 0, srv[i%2], cl[i], s);
			sim_node_change_state(cl[i], N_PLAYING, s);
			uint32_t mark = 1000+i;
			write_record(1, 3, cl[i]->node_id, 4, &mark, s);
		}
		struct ue *v = talloc(1, struct ue);
		v->kind = 1; v->i = i;
		pending[i] = event_new(s->now+40+i, USER, v);
		event_add(pending[i], s);
		if (i%3 == 0) { event_remove(pending[i]); log_info("[%.03lf] follow-up of %d cancelled, active %d\n", s->now, i, (int)is_active(pending[i])); }
	} else if (u->kind == 1) {                  /* follow-up: every 4th client gives up */
		int i = u->i;
		log_info("[%.03lf] follow-up %d connected %d\n", s->now, i, (int)is_connected(srv[i%2], cl[i]));
		if (i%4 == 1 && fl[i]) { flow_close(fl[i], s); fl[i] = NULL; sim_node_change_state(cl[i], N_IDLE, s); }
	}
}
static void on_done(struct event *e, struct sim_state *s){
	struct flow *f = e->data;
	n_done++;
	log_info("[%.03lf] done %d: node %u speed %.3lf\n", s->now, n_done, f->peer[1]->node_id, f->speed[1]);
	for (int i = 0; i < NCL; i++) if (fl[i] == f) { fl[i] = NULL; sim_node_change_state(cl[i], N_DONE, s); }
}
static void on_speed(struct event *e, struct sim_state *s){
	struct spd_event *se = e->data;
	if (se->type == P_RCV && se->speed < 100) log_info("[%.03lf] slow %u %.3lf\n", s->now, se->f->peer[1]->node_id, se->speed);
}

void modx_init(struct sim_state *s){
	open_record("RX", 1, s);
	s->bwcalc = bw; s->dlycalc = dl;
	for (int i = 0; i < 2; i++) {
		srv[i] = sim_create_node(s);
		srv[i]->maximum_bandwidth[0] = 3000; srv[i]->maximum_bandwidth[1] = 3000;
		sim_node_change_state(srv[i], N_SERVER, s);
	}
	for (int i = 0; i < NCL; i++) {
		cl[i] = sim_create_node(s);
		cl[i]->maximum_bandwidth[0] = 500; cl[i]->maximum_bandwidth[1] = 1000;
	}
	struct resource *r = sim_node_new_resource(srv[0], 60000);
	sim_node_add_resource(srv[1], r);
	rid = r->resource_id;
	sim_register_handler(USER, HNDR_USER, on_user, s);
	sim_register_handler(FLOW_DONE, HNDR_USER, on_done, s);
	sim_register_handler(SPEED_CHANGE, HNDR_USER, on_speed, s);
	for (int i = 0; i < NCL; i++) {
		struct ue *u = talloc(1, struct ue);
		u->kind = 0; u->i = i;
		event_add(event_new(1.0+0.5*(i/3), USER, u), s);
	}
}
