/* core.c — netsim's host core re-implemented over the B200 device layer (include/netsim_b200.h).
 *
 * This is the host side of the drop-in boundary: the reference's sim.c / user.c / record.c / main.c
 * API (sim.h:14-53, record.h:19-22, event.h:13-74, flow.h:13-41) with the same names, argument meaning
 * and error behaviour, so that UNMODIFIED module sources (modules/test00.c ... test05.c, test_client1-3.c with
 * client_behaviour.c / resource.c / gaussian.c, p2p.c with cloud_behaviour.c) compile against
 * netsim_b200/host/include and run with every rate solve, byte-progress update, completion re-time and event pop
 * on the GPU.  A module handler registered with sim_register_handler is called on the host between the device's
 * DEFAULT and CLEANER work for that event (ns_dev_run_host).
 *
 * What modules read directly are MIRRORS (host/include/data.h): the same structs, lists, hashes and per-resource range
 * skip lists as the reference's, whose structure this file keeps (a flow is linked when it is established, unlinked when
 * it closes, a merged range leaves its list) and whose numbers — speeds, usage, range progress, completion times — are
 * refreshed from the device (ns_dev_host_view) before every handler and after every call that changes them.  Three
 * things flow the other way: the module's API calls (ns_dev_exec operations), a range a handler advanced itself
 * (_range_update from is_resource_usable / client_lowwm_event ...: NS_OP_RANGE_SET, the same fp64 state the device would
 * have produced), and random(): the module and the core share ONE glibc random() stream (node / resource / flow ids, one
 * draw per skip-list insert, the module's own gaussian draws), so random() / srandom() are defined here and draw from
 * the simulation's generator — on the host until the device tables exist, from the device's state afterwards.
 *
 * Module-side records (sim_create_node, sim_node_change_state, a module's own write_record) are merged into the
 * device's record stream in call order; event_remove of a module-owned event cancels it on the device.
 *
 * Limits (documented in INTEGRATION.md): sim_send_packet is not forwarded (dead in the reference too).
 */
#define _GNU_SOURCE
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define LOG_DOMAIN "sim_api"
#include "log.h"
#include "data.h"
#include "sim.h"
#include "flow.h"
#include "event.h"
#include "store.h"
#include "user.h"
#include "record.h"
#include "core.h"
#include "recfile.h"

#include "../../include/ns_scenario.h"
#include "../../include/netsim_b200.h"

const double eps = 1e-4;            /* sim.c:11 */
#define NS_QMAX 192

struct ns_host_core {
	ns_dev *dev;
	struct ns_rand_state rng;                              /* the random() stream until the device tables exist */
	struct node **nodes; int n_nodes, cap_nodes, nodes_pushed;
	struct resource **rsrc; int n_rsrc, cap_rsrc, rsrc_pushed;   /* every resource copy, in creation order */
	struct flow **flows; int n_flows, cap_flows;           /* by creation ordinal */
	struct flow **live; int n_live, cap_live;              /* the flows that are still open */
	struct range **rng_by_slot; int cap_rng;               /* device range slot -> mirror */
	struct resource **res_by_slot; int cap_res;            /* device resource slot -> mirror */
	struct event **uev; int n_uev, cap_uev;                /* module-owned USER events by token */
	char *record_path; int want_records;
	uint8_t *rec; size_t rec_len, rec_cap;                 /* record stream pulled from the device so far (+ module records) */
	size_t dev_off;                                        /* device stream offset already pulled */
	unsigned epoch;
	/* operations that only touch the event queue or a range's progress state (event_add / event_remove of module-owned events,
	 * range write-backs) wait here and travel in ONE ns_dev_exec, in call order, before anything else talks to the device */
	struct ns_sc_op q[NS_QMAX]; struct event *q_ev[NS_QMAX]; int nq;
	struct sim_state *s;
	int device;
};

static struct ns_host_core *core_of(struct sim_state *s){ return s->core; }
static struct ns_host_core *the_core;     /* sim_node_new_resource(), event_remove() and random() have no sim_state argument */

void ns_host_log(int level, const char *domain, const char *fmt, ...){
	if (level > global_log_level) return;
	va_list ap;
	va_start(ap, fmt);
	fprintf(stderr, "%s: ", domain);
	vfprintf(stderr, fmt, ap);
	va_end(ap);
}
/* NS_HOST_PROF=1: where the wall time of a run goes (printed to stdout at the end) */
#include <time.h>
static double prof_t[8]; static unsigned long prof_n[8]; static int prof_on = -1;
static inline double prof_now(void){ struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec+1e-9*ts.tv_nsec; }
#define PROF_BEGIN double pt0_ = prof_on > 0 ? prof_now() : 0
#define PROF_END(k) do { if (prof_on > 0) { prof_t[k] += prof_now()-pt0_; prof_n[k]++; } } while (0)
static void die(int rc, const char *what){ log_err("device core: %s: %s\n", what, ns_strerror(rc)); exit(3); }

#define GROW(arr, n, cap) do { if ((n) >= (cap)) { int c_ = (cap) ? (cap) : 16; while (c_ <= (n)) c_ *= 2; \
	(arr) = realloc((arr), (size_t)c_*sizeof *(arr)); memset((arr)+(cap), 0, (size_t)(c_-(cap))*sizeof *(arr)); (cap) = c_; } } while (0)

/* ---------------------------------------------------------------- the shared random() stream */
static void q_flush(struct ns_host_core *c);
static uint32_t core_random(struct ns_host_core *c){
	if (!c->dev) return ns_rand_next_host(&c->rng);
	q_flush(c);                                             /* a queued event_add burns its draw first */
	struct ns_rand_state g;
	int rc = ns_dev_rng_get(c->dev, &g);
	if (rc) die(rc, "random");
	uint32_t v = ns_rand_next_host(&g);
	if ((rc = ns_dev_rng_set(c->dev, &g))) die(rc, "random");
	return v;
}
/* modules (and the containers of host/include: one draw per skip_list_insert) call these by name: they resolve to the
 * executable's definitions, i.e. to the simulation's generator, before libc's */
long random(void){
	if (!the_core) { static struct ns_rand_state g; static int seeded; if (!seeded) { ns_rand_seed_host(&g, 1); seeded = 1; } return (long)ns_rand_next_host(&g); }
	return (long)core_random(the_core);
}
void srandom(unsigned seed){
	if (!the_core) return;
	if (the_core->dev) { struct ns_rand_state g; ns_rand_seed_host(&g, seed); ns_dev_rng_set(the_core->dev, &g); }
	else ns_rand_seed_host(&the_core->rng, seed);
}

struct sim_state *ns_host_state_new(int device){
	struct sim_state *s = talloc(1, struct sim_state);
	struct ns_host_core *c = talloc(1, struct ns_host_core);
	s->core = c; c->s = s; c->device = device;
	for (int i = 0; i < LAST_EVENT; i++) INIT_LIST_HEAD(&s->handlers[i]);
	ns_rand_seed_host(&c->rng, 1);            /* glibc default state == srandom(1) */
	the_core = c;
	return s;
}

/* ---------------------------------------------------------------- device calls */
static void dev_exec(struct ns_host_core *c, struct ns_sc_op *ops, int n, const char *what){
	if (ops != c->q) q_flush(c);
	PROF_BEGIN;
	int rc = ns_dev_exec(c->dev, ops, n);
	PROF_END(2);
	if (rc == NS_E_BACK_IN_TIME) { log_err("Add event back in time\n"); abort(); }
	if (rc == NS_E_NAN_TIME) { log_err("Add event at nan\n"); abort(); }
	if (rc) die(rc, what);
}
static void q_flush(struct ns_host_core *c){
	if (!c->nq || !c->dev) return;
	const int n = c->nq;
	c->nq = 0;
	dev_exec(c, c->q, n, "event_add / event_remove");
	for (int i = 0; i < n; i++)
		if (c->q[i].kind == NS_OP_SCHED && c->q_ev[i] && c->q_ev[i]->q_idx == i) { c->q_ev[i]->dev_slot = c->q[i].b; c->q_ev[i]->q_idx = -1; }
}
static struct ns_sc_op *q_push(struct ns_host_core *c, struct event *e){
	if (c->nq == NS_QMAX) q_flush(c);
	struct ns_sc_op *op = &c->q[c->nq];
	memset(op, 0, sizeof *op);
	c->q_ev[c->nq++] = e;
	return op;
}
/* nodes and resources the module has created since the last call, in creation order.  Pushed lazily: a module sets
 * maximum_bandwidth / bit_rate on the struct AFTER sim_create_node / sim_node_new_resource returned (p2p.c:144-170). */
static void flush_pending(struct ns_host_core *c){
	if (!c->dev) return;
	struct ns_sc_op ops[64]; int n = 0;
	while (c->nodes_pushed < c->n_nodes) {
		struct node *nd = c->nodes[c->nodes_pushed++];
		struct ns_sc_op *op = &ops[n++];
		memset(op, 0, sizeof *op);
		op->kind = NS_OP_NEWNODE; op->a = nd->index; op->c = (int32_t)nd->node_id; op->x = nd->maximum_bandwidth[0]; op->y = nd->maximum_bandwidth[1];
		nd->on_device = true;
		if (n == 64) { dev_exec(c, ops, n, "sim_create_node"); n = 0; }
	}
	if (n) { dev_exec(c, ops, n, "sim_create_node"); n = 0; }
	for (int k = 0; k < c->n_rsrc; k++) {                   /* resource->auto_delete is set by the module on the struct (resource.c:63) */
		struct resource *r = c->rsrc[k];
		if (!r || !r->auto_delete || r->autodel_pushed || r->dev_slot < 0) continue;
		struct ns_sc_op op;
		memset(&op, 0, sizeof op);
		op.kind = NS_OP_AUTODEL; op.a = r->dev_slot;
		dev_exec(c, &op, 1, "auto_delete");
		r->autodel_pushed = true;
	}
	while (c->rsrc_pushed < c->n_rsrc) {
		struct resource *who[64];
		while (c->rsrc_pushed < c->n_rsrc && n < 64) {
			struct resource *r = c->rsrc[c->rsrc_pushed++];
			if (!r) continue;                               /* deleted again before it reached the device */
			struct ns_sc_op *op = &ops[n];
			who[n++] = r;
			memset(op, 0, sizeof *op);
			op->kind = NS_OP_NEWRES; op->a = r->owner->index; op->b = r->bit_rate; op->c = (int32_t)r->resource_id; op->start = (int64_t)r->len;
		}
		if (!n) break;
		dev_exec(c, ops, n, "sim_node_new_resource");
		for (int i = 0; i < n; i++) {
			struct resource *r = who[i];
			if (ops[i].b < 0) continue;
			r->dev_slot = ops[i].b;
			GROW(c->res_by_slot, r->dev_slot, c->cap_res); c->res_by_slot[r->dev_slot] = r;
			struct range *g = r->ranges.next[0] ? skip_list_entry(r->ranges.next[0], struct range, ranges) : NULL;
			if (g && ops[i].c >= 0) { g->dev_slot = ops[i].c; GROW(c->rng_by_slot, g->dev_slot, c->cap_rng); c->rng_by_slot[g->dev_slot] = g; }
		}
		n = 0;
	}
}

/* the device tables are laid out when the first flow (or the event loop) needs them */
static void materialize(struct ns_host_core *c){
	if (c->dev) return;
	struct ns_sc_header H;
	memset(&H, 0, sizeof H);
	H.magic = NS_SC_MAGIC; H.version = NS_SC_VERSION; H.seed = 1; H.n_groups = 1;
	char buf[sizeof H+8];
	memset(buf, 0, sizeof buf);
	memcpy(buf, &H, sizeof H);                              /* no nodes, no resources, group_off = {0, 0}: the module itself is the script */
	struct ns_dev_opts o;
	memset(&o, 0, sizeof o);
	/* nothing is known about how many nodes / flows the module will create: the tables are sized by the environment
	 * (the defaults fit a p2p.c run of a few thousand clients) */
	const char *ev;
	o.device = c->device; o.mode = NS_MODE_SEQ; o.want_records = c->want_records;
	o.features_off = 8;                                     /* NS_FEAT_INIT_REC: sim_create_node writes the R_NODE_CREATE records itself */
	o.cap_nodes = (ev = getenv("NS_CAP_NODES")) ? atoi(ev) : (c->n_nodes*2 > 1 << 14 ? c->n_nodes*2 : 1 << 14);
	o.cap_flows = (ev = getenv("NS_CAP_FLOWS")) ? atoi(ev) : 1 << 16;      /* concurrent flows */
	o.cap_ord = (ev = getenv("NS_CAP_ORD")) ? atoi(ev) : 1 << 21;          /* flows ever created */
	o.cap_ranges = 2*o.cap_flows+4*o.cap_nodes+16;
	o.cap_users = (ev = getenv("NS_CAP_USERS")) ? atoi(ev) : 1 << 16;      /* pending module-owned events */
	o.ring_cap = 1u << 22; o.record_cap = 64u << 20;
	int rc = ns_dev_create(&c->dev, buf, sizeof buf, &o);
	if (rc) { c->dev = NULL; die(rc, "ns_dev_create"); }
	flush_pending(c);
	if ((rc = ns_dev_rng_set(c->dev, &c->rng))) die(rc, "ns_dev_rng_set");   /* the stream goes on where the host left it */
}

/* ---------------------------------------------------------------- mirrors */
static void resource_forget(struct ns_host_core *c, struct resource *r);
static void flow_unlink(struct ns_host_core *c, struct flow *f){
	if (f->closed) return;
	f->closed = true;
	c->live[f->live_idx] = c->live[--c->n_live]; c->live[f->live_idx]->live_idx = f->live_idx;
	list_del(&f->conns[0]); list_del(&f->conns[1]);
	HASH_DELETE(hh2, f->peer[0]->outs, f);
	if (f->hh.tbl) HASH_DEL(c->s->flows, f);
	struct resource *sr = f->srng ? f->srng->owner : NULL;
	if (f->srng) { list_del(&f->consumers); sr->nconsumer--; }
	if (f->drng) { if (f->drng->producer == f) f->drng->producer = NULL; f->drng->owner->nproducer--; }
	f->done->active = false; f->drain->active = false;
	if (sr && sr->nconsumer == 0 && sr->nproducer == 0 && sr->auto_delete) resource_forget(c, sr);   /* flow.c:278-281; the device's flow_close does the same */
}
static struct range *range_mirror_new(struct ns_host_core *c, struct resource *r, int slot, size_t start, double len){
	struct range *g = range_new(start, 0);
	g->len = len; g->total_len = r->len; g->owner = r; g->dev_slot = slot;
	int key = (int)start;
	struct skip_list_head *after = skip_list_find_le(&r->ranges, &key, range_start_cmp);   /* sorted by start; no random() draw: the device burns it */
	ns_skip_link_after(&g->ranges, after ? after : &r->ranges);
	if (slot >= 0) { GROW(c->rng_by_slot, slot, c->cap_rng); c->rng_by_slot[slot] = g; }
	return g;
}
/* every number a handler may read, as the device has it now */
static void refresh(struct ns_host_core *c){
	if (!c->dev) return;
	q_flush(c);
	struct sim_state *s = c->s;
	const struct ns_view_hdr *vh; const double *nd; const struct ns_view_flow *vf; const struct ns_view_range *vr;
	PROF_BEGIN;
	int rc = ns_dev_host_view(c->dev, &vh, &nd, &vf, &vr);
	PROF_END(1);
	if (rc) die(rc, "ns_dev_host_view");
	double pt1_ = prof_on > 0 ? prof_now() : 0;
	for (int i = 0; i < vh->n_nodes && i < c->n_nodes; i++) {
		c->nodes[i]->bandwidth_usage[0] = nd[4*i]; c->nodes[i]->bandwidth_usage[1] = nd[4*i+1];
		c->nodes[i]->total_bwupbound[0] = nd[4*i+2]; c->nodes[i]->total_bwupbound[1] = nd[4*i+3];
	}
	const unsigned ep = ++c->epoch;
	for (int i = 0; i < vh->hi_range; i++) {
		const struct ns_view_range *v = &vr[i];
		if (v->slot < 0) continue;
		struct resource *r = v->res < c->cap_res ? c->res_by_slot[v->res] : NULL;
		if (!r) continue;                                   /* a resource the host never saw cannot exist: every resource is created here */
		struct range *g = v->slot < c->cap_rng ? c->rng_by_slot[v->slot] : NULL;
		if (!g || g->owner != r || g->start != (size_t)v->start) g = range_mirror_new(c, r, v->slot, (size_t)v->start, v->len);
		g->total_len = (size_t)v->total_len;
		g->len = g->dev_len = v->len; g->lenc = g->dev_lenc = v->lenc; g->last_update = g->dev_last = v->last_update;
		g->producer = (v->producer_ord >= 0 && v->producer_ord < c->n_flows) ? c->flows[v->producer_ord] : NULL;
		g->seen = ep;
	}
	for (int k = 0; k < c->cap_rng; k++) {                  /* merged into its predecessor (range_merge_with_next, range.c:83-124): out of the list */
		struct range *g = c->rng_by_slot[k];
		if (!g || g->seen == ep) continue;
		struct flow *f;
		while (!list_empty(&g->consumers)) { f = list_entry(g->consumers.next, struct flow, consumers); list_del(&f->consumers); INIT_LIST_HEAD(&f->consumers); f->srng = NULL; }
		if (g->ranges.next) skip_list_delete(&g->ranges);
		c->rng_by_slot[k] = NULL;
	}
	for (int i = 0; i < vh->hi_flow; i++) {
		const struct ns_view_flow *v = &vf[i];
		if (v->slot < 0) continue;
		struct flow *f = (v->ord >= 0 && v->ord < c->n_flows) ? c->flows[v->ord] : NULL;
		if (!f || f->closed) continue;
		f->seen = ep;
		f->speed[0] = v->speed[0]; f->speed[1] = v->speed[1];
		if (!f->hh.tbl) { f->flow_id = v->flow_id; HASH_ADD_INT(s->flows, flow_id, f); }
		f->done->active = !isinf(v->done_t); if (f->done->active) f->done->time = v->done_t;
		f->drain->active = !isinf(v->drain_t); if (f->drain->active) f->drain->time = v->drain_t;
		f->drain->type = v->drain_type == FLOW_SPEED_THROTTLE ? FLOW_SPEED_THROTTLE : FLOW_DRAIN;
		struct range *dr = (v->drng >= 0 && v->drng < c->cap_rng) ? c->rng_by_slot[v->drng] : NULL;
		struct range *sr = (v->srng >= 0 && v->srng < c->cap_rng) ? c->rng_by_slot[v->srng] : NULL;
		if (dr) f->drng = dr;
		if (sr && sr != f->srng) {                          /* its source range was merged into the one before it */
			if (f->srng) list_del(&f->consumers);
			f->srng = sr; list_add(&f->consumers, &sr->consumers);
		}
	}
	for (int k = c->n_live-1; k >= 0; k--) {                /* closed by the device (flow_done_cleaner, flow.c:403-406) */
		struct flow *f = c->live[k];
		if (f->seen != ep) flow_unlink(c, f);
	}
	if (prof_on > 0) { prof_t[3] += prof_now()-pt1_; prof_n[3]++; }
}
/* a handler advanced a range itself (_range_update): the device takes the new state over */
static void push_range_updates(struct ns_host_core *c){
	if (!c->dev) return;
	for (int k = 0; k < c->cap_rng; k++) {
		struct range *g = c->rng_by_slot[k];
		if (!g || (g->len == g->dev_len && g->lenc == g->dev_lenc && g->last_update == g->dev_last)) continue;
		struct ns_sc_op *op = q_push(c, NULL);
		op->kind = NS_OP_RANGE_SET; op->a = k; op->x = g->len; op->y = g->lenc; memcpy(&op->start, &g->last_update, 8);
		g->dev_len = g->len; g->dev_lenc = g->lenc; g->dev_last = g->last_update;
	}
}

/* ---------------------------------------------------------------- the module API */
/* sim_create_node, sim.c:113-125: id = random() until unused */
struct node *sim_create_node(struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	struct node *n = talloc(1, struct node), *on;
	n->state = N_OFFLINE;
	n->store = store_new();
	INIT_LIST_HEAD(&n->conns[0]); INIT_LIST_HEAD(&n->conns[1]);
	do {
		n->node_id = core_random(c);
		HASH_FIND_INT(s->nodes, &n->node_id, on);
	} while (on);
	HASH_ADD_INT(s->nodes, node_id, n);
	n->index = c->n_nodes;
	GROW(c->nodes, c->n_nodes, c->cap_nodes);
	c->nodes[c->n_nodes++] = n;
	uint8_t padding = 0;
	write_record(0, R_NODE_CREATE, n->node_id, 1, &padding, s);        /* sim.c:122-123 */
	return n;
}

/* a complete copy of a resource on a node: one range [0, len), whose skip-list insert takes one random() draw */
static void resource_attach(struct ns_host_core *c, struct node *n, struct resource *r){
	r->owner = n;
	struct range *g = range_new(0, r->len);
	g->total_len = r->len; g->owner = r;
	ns_skip_link_after(&g->ranges, &r->ranges);
	core_random(c);                                         /* resource_new_range -> skip_list_insert (range.h:71-82, skiplist.h:34-40) */
	GROW(c->rsrc, c->n_rsrc, c->cap_rsrc);
	r->all_idx = c->n_rsrc;
	c->rsrc[c->n_rsrc++] = r;
}
/* sim_node_new_resource, sim.c:129-140 */
struct resource *sim_node_new_resource(struct node *n, size_t len){
	struct ns_host_core *c = the_core;
	struct resource *r = resource_new(core_random(c), len);
	r->owner = n;
	while (store_set(n->store, r) == -1) r->resource_id = core_random(c);
	resource_attach(c, n, r);
	return r;
}
/* sim_node_add_resource, sim.c:141-154 */
struct resource *sim_node_add_resource(struct node *n, struct resource *r){
	struct ns_host_core *c = the_core;
	if (store_get(n->store, r->resource_id)) return NULL;
	struct resource *nr = resource_new(r->resource_id, r->len);
	nr->bit_rate = r->bit_rate;
	nr->owner = n;
	store_set(n->store, nr);
	resource_attach(c, n, nr);
	return nr;
}
/* range_clear + store_delete of the mirror (sim.h:47-53); the structs stay allocated: a module may still hold a pointer */
static void resource_forget(struct ns_host_core *c, struct resource *r){
	for (struct skip_list_head *h = r->ranges.next ? r->ranges.next[0] : NULL, *nx; h; h = nx) {
		nx = h->next[0];
		struct range *g = skip_list_entry(h, struct range, ranges);
		if (g->dev_slot >= 0 && g->dev_slot < c->cap_rng && c->rng_by_slot[g->dev_slot] == g) c->rng_by_slot[g->dev_slot] = NULL;
		skip_list_deinit_node(&g->ranges);
	}
	skip_list_deinit_head(&r->ranges);
	store_delete(r->owner->store, r);
	if (r->dev_slot >= 0 && r->dev_slot < c->cap_res) c->res_by_slot[r->dev_slot] = NULL;
	c->rsrc[r->all_idx] = NULL;
}
void node_del_resource(struct resource *r){
	struct ns_host_core *c = the_core;
	if (c->dev) {
		flush_pending(c);                                   /* (it may have been created a moment ago) */
		struct ns_sc_op op;
		memset(&op, 0, sizeof op);
		op.kind = NS_OP_DELRES; op.a = r->dev_slot;
		dev_exec(c, &op, 1, "node_del_resource");
	}                                                       /* (not on the device yet: resource_forget empties its slot in the pending list) */
	resource_forget(c, r);
}

/* sim_register_handler, sim.c:96-111: sorted by priority, equal priority in registration order */
void sim_register_handler(int type, int priority, event_handler_func f, struct sim_state *s){
	struct event_handler *neh = talloc(1, struct event_handler), *eh;
	neh->f = f; neh->pri = priority;
	struct list_head *pos = &s->handlers[type];
	list_for_each_entry(eh, &s->handlers[type], handlers) if ((int)eh->pri > priority) { pos = &eh->handlers; break; }
	list_add_tail(&neh->handlers, pos);                     /* before the first handler of a higher priority */
}

/* sim_establish_flow, sim.c:42-94 */
struct flow *sim_establish_flow(id_t rid, size_t start, struct node *src, struct node *dst, struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	struct resource *sr = store_get(src->store, rid);
	if (!sr) { log_err("The resource %d doesn't exist on source node %d\n", rid, src->node_id); return NULL; }
	struct resource *dr = store_get(dst->store, rid);
	if (dr && range_get(dr, start)) { log_err("Trying to create a flow to a existing range.\n"); return NULL; }   /* sim.c:62-67 */
	materialize(c);
	flush_pending(c);
	push_range_updates(c);
	struct ns_sc_op op;
	memset(&op, 0, sizeof op);
	op.kind = NS_OP_CREATE_RID; op.a = src->index; op.b = dst->index; op.c = (int32_t)rid; op.start = (int64_t)start;
	op.x = s->bwcalc(src->user_data, dst->user_data);      /* flow.c:303-304 */
	op.y = s->dlycalc(src->user_data, dst->user_data);     /* flow.c:309 */
	dev_exec(c, &op, 1, "sim_establish_flow");
	if (op.a < 0) die(NS_E_INTERNAL, "sim_establish_flow refused by the device");
	if (!dr) {                                              /* sim.c:54-61: the destination learns of the resource */
		dr = resource_new(sr->resource_id, sr->len);
		dr->bit_rate = sr->bit_rate; dr->owner = dst;
		store_set(dst->store, dr);
		dr->dev_slot = op.b;
		GROW(c->res_by_slot, dr->dev_slot, c->cap_res); c->res_by_slot[dr->dev_slot] = dr;
		GROW(c->rsrc, c->n_rsrc, c->cap_rsrc);              /* (already on the device: flush_pending is up to date before every create) */
		dr->all_idx = c->n_rsrc; c->rsrc[c->n_rsrc++] = dr; c->rsrc_pushed = c->n_rsrc;
	}
	struct flow *f = talloc(1, struct flow);
	f->ord = op.a;
	GROW(c->flows, f->ord, c->cap_flows);
	if (c->n_flows <= f->ord) c->n_flows = f->ord+1;
	c->flows[f->ord] = f;
	GROW(c->live, c->n_live, c->cap_live);
	f->live_idx = c->n_live; c->live[c->n_live++] = f;
	f->peer[0] = src; f->peer[1] = dst; f->dst_id = dst->node_id;
	f->bwupbound = op.x; f->delay = op.y; f->resource_id = rid;
	f->start = start; f->begin_time = s->now;
	INIT_LIST_HEAD(&f->spd_evs);
	list_add(&f->conns[0], &src->conns[0]);                 /* flow.c:311-312: newest first */
	list_add(&f->conns[1], &dst->conns[1]);
	HASH_ADD(hh2, src->outs, dst_id, sizeof(f->dst_id), f);
	f->done = event_new(0, FLOW_DONE, f); f->drain = event_new(0, FLOW_DRAIN, f);
	f->drng = range_mirror_new(c, dr, (int)op.c, start, 0);
	f->drng->producer = f; f->drng->last_update = s->now;
	f->srng = ((int)op.start >= 0 && (int)op.start < c->cap_rng) ? c->rng_by_slot[(int)op.start] : NULL;
	if (f->srng) list_add(&f->consumers, &f->srng->consumers); else INIT_LIST_HEAD(&f->consumers);
	sr->nconsumer++; dr->nproducer++;
	refresh(c);
	return f;
}

/* flow_close (flow.c:241-292) called by a module */
void flow_close(struct flow *f, struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	if (!f || f->closed) return;
	push_range_updates(c);
	struct ns_sc_op op;
	memset(&op, 0, sizeof op);
	op.kind = NS_OP_CLOSE; op.a = f->ord;
	dev_exec(c, &op, 1, "flow_close");
	flow_unlink(c, f);
	refresh(c);
}

/* event_add, event.h:56-74, for module-owned events */
void event_add(struct event *e, struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	if (!e || e->active) return;
	if (e->time < s->now) {                                 /* event.h:58-65 */
		if (e->time < s->now-eps) { log_err("Add event back in time, e: %lf now: %lf\n", e->time, s->now); abort(); }
		e->time = s->now;
	}
	if (isnan(e->time)) { log_err("Add event at nan\n"); abort(); }
	materialize(c);
	if (e->token < 0) { GROW(c->uev, c->n_uev, c->cap_uev); e->token = c->n_uev; c->uev[c->n_uev++] = e; }
	struct ns_sc_op *op = q_push(c, e);
	op->kind = NS_OP_SCHED; op->a = e->token; op->c = 0; op->x = e->time;
	e->q_idx = c->nq-1;
	e->active = true;
	e->dev_slot = -1;                                       /* known when the queue has been executed */
}
/* event_remove, event.h:13-18, of a module-owned event */
void event_remove(struct event *e){
	struct ns_host_core *c = the_core;
	if (!e || !e->active) return;
	e->active = false;
	if (!c || !c->dev) return;
	if (c->nq == NS_QMAX) q_flush(c);                       /* (room first: afterwards q_idx / dev_slot are final) */
	struct ns_sc_op *op;
	if (e->q_idx >= 0) {                                    /* its event_add is still in the queue: cancel the slot that one will get */
		op = q_push(c, NULL);
		op->kind = NS_OP_CANCEL; op->a = -(e->q_idx+1);
		e->q_idx = -1;
		return;
	}
	if (e->dev_slot < 0) return;
	op = q_push(c, NULL);
	op->kind = NS_OP_CANCEL; op->a = e->dev_slot;
	e->dev_slot = -1;
}

void sim_send_packet(void *data, int len, struct node *src, struct node *dst, struct sim_state *s){
	(void)data; (void)len; (void)src; (void)dst; (void)s;   /* dead in the reference too: no PACKET_DONE handler */
}
void ns_host_refresh(struct sim_state *s){ refresh(core_of(s)); }

void user_speed_change(struct event *e, struct sim_state *s){ (void)e; (void)s; }   /* user.c:7-13 */
void user_done(struct event *e, struct sim_state *s){ (void)e; (void)s; }

/* ---------------------------------------------------------------- records */
static void rec_reserve(struct ns_host_core *c, size_t more){
	if (c->rec_len+more <= c->rec_cap) return;
	while (c->rec_len+more > c->rec_cap) c->rec_cap = c->rec_cap ? c->rec_cap*2 : 1 << 20;
	c->rec = realloc(c->rec, c->rec_cap);
}
/* append what the device has produced since the last pull to the host copy of the record stream */
static void pull_records(struct ns_host_core *c){
	if (!c->want_records || !c->dev) return;
	for (;;) {
		size_t got = 0, total = 0;
		ns_dev_records(c->dev, NULL, 0, c->dev_off, &got, &total);
		if (total <= c->dev_off) return;
		rec_reserve(c, total-c->dev_off);
		ns_dev_records(c->dev, c->rec+c->rec_len, total-c->dev_off, c->dev_off, &got, &total);
		if (!got) return;
		c->rec_len += got; c->dev_off += got;
	}
}
/* write_record, record.c:27-117, for records the MODULE side writes (sim_create_node, sim_node_change_state, a module's own):
 * what the device has produced so far is pulled first, so the merged stream keeps the reference's order */
void write_record(uint8_t major, uint8_t minor, uint32_t id, int8_t bytes, void *value, struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	if (!c->record_path) return;                            /* record.c:29-32: no file, no record */
	pull_records(c);
	rec_reserve(c, 19);
	uint8_t *b = c->rec+c->rec_len; int n = 0;
	b[n++] = major; b[n++] = minor;
	b[n++] = id >> 24; b[n++] = id >> 16; b[n++] = id >> 8; b[n++] = id;
	uint32_t t = (uint32_t)(int)s->now;                      /* record.c:66-71 */
	b[n++] = t >> 24; b[n++] = t >> 16; b[n++] = t >> 8; b[n++] = t;
	uint16_t ut = (uint16_t)(int)((s->now-(double)((int)s->now))*1000);
	b[n++] = ut >> 8; b[n++] = ut;
	b[n++] = (uint8_t)bytes;
	if (bytes == 1) b[n++] = *(uint8_t *)value;
	else if (bytes == 2) { uint16_t v = *(uint16_t *)value; b[n++] = v >> 8; b[n++] = v; }
	else if (bytes == 4) { uint32_t v = *(uint32_t *)value; b[n++] = v >> 24; b[n++] = v >> 16; b[n++] = v >> 8; b[n++] = v; }
	else {                                                  /* -1: a double as integer part + milli part, record.c:98-106 */
		double tmp = *(double *)value;
		uint32_t ip = (uint32_t)(int)tmp;
		b[n++] = ip >> 24; b[n++] = ip >> 16; b[n++] = ip >> 8; b[n++] = ip;
		tmp = tmp-(double)((int)tmp);
		uint16_t fp = (uint16_t)(int)(tmp*1000);
		b[n++] = fp >> 8; b[n++] = fp;
	}
	c->rec_len += (size_t)n;
}
/* open_record, record.c:119-153: remember the file; the record stream is saved there */
void open_record(const char *filename, int create, struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	(void)create;
	if (c->record_path) return;                             /* no-op if already open, record.c:120-123 */
	c->record_path = strdup(filename);
	c->want_records = 1;
}
static void save_records(struct ns_host_core *c){
	if (!c->record_path) return;
	pull_records(c);
	int rc = ns_recfile_save(c->record_path, c->rec, c->rec_len, ns_recfile_split_from_env(), ns_recfile_flags_from_env(), NULL);   /* recfile.c */
	if (rc < 0) log_err("Can't write record file %s (%d)\n", c->record_path, rc);
}

/* ---------------------------------------------------------------- the event loop, main.c:50-67, with the default handlers (main.c:44-48) on the device */
int ns_host_run(struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	if (prof_on < 0) { const char *pe = getenv("NS_HOST_PROF"); prof_on = pe && *pe && *pe != '0'; }
	materialize(c);
	flush_pending(c);
	push_range_updates(c);
	uint32_t mask = 1u << USER;
	for (int t = 0; t < LAST_EVENT; t++) if (!list_empty(&s->handlers[t])) mask |= 1u << t;
	ns_dev_set_host_mask(c->dev, mask);
	int rc = 0;
	while (!s->exit) {
		struct ns_host_event ev;
		q_flush(c);
		{ PROF_BEGIN; rc = ns_dev_run_host(c->dev, &ev); PROF_END(0); }
		if (rc == NS_E_BACK_IN_TIME) { log_err("Add event back in time\n"); abort(); }
		if (rc) { log_err("device core: %s\n", ns_strerror(rc)); break; }
		if (!ev.stopped) break;                             /* queue empty */
		if (ev.stopped == 2) {                              /* record buffer nearly full BEFORE a batch: nothing was handled, */
			size_t before = c->dev_off;                     /* no event to dispatch — copy the records out and go on */
			pull_records(c);
			if (ns_dev_records_drain(c->dev) || (c->dev_off == before && c->want_records)) {
				log_err("device core: record buffer too small for one batch\n");
				rc = NS_E_RECORDS_LOST;
				break;
			}
			continue;
		}
		s->now = ev.time;                                   /* main.c:60 */
		struct event tmp, *e = &tmp;
		struct spd_event se;
		struct flow *f = (ev.flow_ord >= 0 && ev.flow_ord < c->n_flows) ? c->flows[ev.flow_ord] : NULL;
		memset(&tmp, 0, sizeof tmp);
		tmp.type = ev.type; tmp.time = ev.time; tmp.token = -1; tmp.dev_slot = -1; tmp.q_idx = -1;
		bool need_free = false;
		refresh(c);
		switch (ev.type) {
		case SPEED_CHANGE:
			memset(&se, 0, sizeof se);
			se.e = e; se.f = f; se.speed = ev.speed; se.type = ev.dir;
			tmp.data = &se;
			break;
		case USER:
			e = (ev.token >= 0 && ev.token < c->n_uev) ? c->uev[ev.token] : NULL;
			if (!e) continue;
			e->active = false;                              /* event_pop, event.h:20-25 */
			e->dev_slot = -1;
			need_free = e->auto_free;
			break;
		case FLOW_DONE:
			if (f) { e = f->done; e->active = false; }
			tmp.data = f;
			break;
		default:
			if (f) { e = f->drain; e->active = false; }
			tmp.data = f;
		}
		struct event_handler *eh;
		{ PROF_BEGIN; list_for_each_entry(eh, &s->handlers[ev.type], handlers) eh->f(e, s); PROF_END(4); }
		if (ev.type == FLOW_DONE && f) flow_unlink(c, f);   /* flow_done_cleaner (flow.c:403-406) runs on the device next */
		{ PROF_BEGIN; flush_pending(c); push_range_updates(c); PROF_END(5); }
		q_flush(c);                                         /* (nothing queued refers to the event freed next) */
		if (need_free) { c->uev[ev.token] = NULL; free(e); }   /* main.c:65-66 */
		if (s->exit) {
			struct ns_sc_op op;
			memset(&op, 0, sizeof op);
			op.kind = NS_OP_END;
			ns_dev_exec(c->dev, &op, 1);
		}
	}
	save_records(c);
	if (prof_on > 0) {
		static const char *nm[] = { "ns_dev_run_host", "ns_dev_host_view", "ns_dev_exec", "mirror update", "module handlers (incl. their device calls)", "pending scans" };
		for (int k = 0; k < 6; k++) printf("host core: %-44s %9.3f s  %9lu calls  %7.1f us/call\n", nm[k], prof_t[k], prof_n[k], prof_n[k] ? 1e6*prof_t[k]/prof_n[k] : 0.0);
	}
	return rc;
}

void ns_host_finish(struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	if (c->dev) ns_dev_destroy(c->dev);
	c->dev = NULL;
}
