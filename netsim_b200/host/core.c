/* core.c — netsim's host core re-implemented over the B200 device layer (include/netsim_b200.h).
 *
 * This is the host side of the drop-in boundary: the reference's sim.c / user.c / record.c / main.c
 * API (sim.h:14-53, record.h:19-22, event.h:13-74, flow.h:13-41) with the same names, argument meaning
 * and error behaviour, so that UNMODIFIED module sources (modules/test00.c ... test05.c) compile
 * against netsim_b200/host/include and run with every rate solve, byte-progress update, completion
 * re-time and event pop on the GPU.  A module handler registered with sim_register_handler is called
 * on the host between the device's DEFAULT and CLEANER work for that event (ns_dev_run_host), with the
 * mirror structs it may read refreshed first.
 *
 * Module-side records (sim_create_node, sim_node_change_state, a module's own write_record) are merged into the
 * device's record stream in call order; event_remove of a module-owned event cancels it on the device.
 *
 * Limits (documented in INTEGRATION.md): nodes and resources must be created before the first flow;
 * sim_send_packet is not forwarded (dead in the reference too); the mirrors carry no list.h / uthash.h /
 * skip-list containers, so modules that walk those (p2p.c's behaviour library) do not compile against
 * them — their client side runs on the device through the script form (NS_SC_F_CLIENTS).
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

#include "../../include/ns_scenario.h"
#include "../../include/netsim_b200.h"

const double eps = 1e-4;            /* sim.c:11 */

struct handler { event_handler_func f; int pri; struct handler *next; };

struct ns_host_core {
	ns_dev *dev;
	struct ns_rand_state rng;
	struct node **nodes; int n_nodes, cap_nodes;
	struct resource **rsrc; int n_rsrc, cap_rsrc;          /* catalogue (sim_node_new_resource order) */
	struct { int node, res; } *copies; int n_copies, cap_copies;
	struct flow **flows; int n_flows, cap_flows;           /* by creation ordinal */
	struct event **uev; int n_uev, cap_uev;                /* module-owned USER events by token */
	struct handler *handlers[LAST_EVENT];
	char *record_path; int want_records;
	uint8_t *rec; size_t rec_len, rec_cap;                 /* record stream pulled from the device so far (+ module records) */
	size_t dev_off;                                        /* device stream offset already pulled */
	struct sim_state *s;
	int device, mode;
};

static struct ns_host_core *core_of(struct sim_state *s){ return s->core; }
static struct ns_host_core *the_core;     /* sim_node_new_resource() has no sim_state argument */

void ns_host_log(int level, const char *domain, const char *fmt, ...){
	if (level > global_log_level) return;
	va_list ap;
	va_start(ap, fmt);
	fprintf(stderr, "%s: ", domain);
	vfprintf(stderr, fmt, ap);
	va_end(ap);
}

#define GROW(arr, n, cap) do { if ((n) == (cap)) { (cap) = (cap) ? (cap)*2 : 16; (arr) = realloc((arr), (size_t)(cap)*sizeof *(arr)); } } while (0)

struct sim_state *ns_host_state_new(int device){
	struct sim_state *s = talloc(1, struct sim_state);
	struct ns_host_core *c = talloc(1, struct ns_host_core);
	s->core = c; c->s = s; c->device = device;
	ns_rand_seed_host(&c->rng, 1);            /* glibc default state == srandom(1) */
	the_core = c;
	return s;
}

/* sim_create_node, sim.c:113-125: id = random() until unused */
struct node *sim_create_node(struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	if (c->dev) { log_err("sim_create_node after the first flow is not supported by the device core\n"); return NULL; }
	struct node *n = talloc(1, struct node);
	n->state = N_OFFLINE;
	n->store = talloc(1, struct store);
	for (;;) {
		n->node_id = ns_rand_next_host(&c->rng);
		int dup = 0;
		for (int i = 0; i < c->n_nodes; i++) if (c->nodes[i]->node_id == n->node_id) dup = 1;
		if (!dup) break;
	}
	n->index = c->n_nodes;
	GROW(c->nodes, c->n_nodes, c->cap_nodes);
	c->nodes[c->n_nodes++] = n;
	uint8_t padding = 0;
	write_record(0, R_NODE_CREATE, n->node_id, 1, &padding, s);        /* sim.c:122-123 */
	return n;
}

static struct resource *store_add(struct node *n, id_t rid, size_t len, int script_index){
	struct resource *r = talloc(1, struct resource);
	r->resource_id = rid; r->len = len; r->owner = n; r->script_index = script_index;
	r->next_in_store = n->store->head; n->store->head = r;
	n->store->total_size += len;
	return r;
}

/* sim_node_new_resource, sim.c:129-140 */
struct resource *sim_node_new_resource(struct node *n, size_t len){
	struct ns_host_core *c = the_core;
	if (c->dev) { log_err("sim_node_new_resource after the first flow is not supported by the device core\n"); return NULL; }
	id_t rid = ns_rand_next_host(&c->rng);
	while (store_get(n->store, rid)) rid = ns_rand_next_host(&c->rng);
	struct resource *r = store_add(n, rid, len, c->n_rsrc);
	ns_rand_next_host(&c->rng);               /* resource_new_range -> skip_list_insert burns one draw */
	GROW(c->rsrc, c->n_rsrc, c->cap_rsrc);
	c->rsrc[c->n_rsrc++] = r;
	return r;
}

/* sim_node_add_resource, sim.c:141-154 */
struct resource *sim_node_add_resource(struct node *n, struct resource *r){
	struct ns_host_core *c = the_core;
	if (c->dev) { log_err("sim_node_add_resource after the first flow is not supported by the device core\n"); return NULL; }
	if (store_get(n->store, r->resource_id)) return NULL;
	int idx = -1;
	for (int i = 0; i < c->n_rsrc; i++) if (c->rsrc[i]->resource_id == r->resource_id) { idx = i; break; }
	if (idx < 0) return NULL;
	struct resource *nr = store_add(n, r->resource_id, r->len, idx);
	nr->bit_rate = r->bit_rate;
	ns_rand_next_host(&c->rng);
	GROW(c->copies, c->n_copies, c->cap_copies);
	c->copies[c->n_copies].node = n->index; c->copies[c->n_copies].res = idx; c->n_copies++;
	return nr;
}

/* sim_register_handler, sim.c:96-111: sorted by priority, equal priority in registration order */
void sim_register_handler(int type, int priority, event_handler_func f, struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	struct handler *h = talloc(1, struct handler), **pp = &c->handlers[type];
	h->f = f; h->pri = priority;
	while (*pp && (*pp)->pri <= priority) pp = &(*pp)->next;
	h->next = *pp; *pp = h;
}

/* the device tables are laid out when the first flow (or the event loop) needs them */
static int materialize(struct ns_host_core *c){
	if (c->dev) return 0;
	struct ns_sc_header H;
	memset(&H, 0, sizeof H);
	H.magic = NS_SC_MAGIC; H.version = NS_SC_VERSION; H.seed = 1;
	H.n_nodes = c->n_nodes; H.n_resources = c->n_rsrc; H.n_copies = c->n_copies; H.n_groups = 1;
	size_t len = sizeof H+(size_t)c->n_nodes*16+(size_t)c->n_rsrc*sizeof(struct ns_sc_resource)+
		(size_t)c->n_copies*sizeof(struct ns_sc_copy)+8;
	char *buf = calloc(1, len), *p = buf;
	memcpy(p, &H, sizeof H); p += sizeof H;
	for (int i = 0; i < c->n_nodes; i++) { memcpy(p, &c->nodes[i]->maximum_bandwidth[0], 8); p += 8; }
	for (int i = 0; i < c->n_nodes; i++) { memcpy(p, &c->nodes[i]->maximum_bandwidth[1], 8); p += 8; }
	for (int i = 0; i < c->n_rsrc; i++) {
		struct ns_sc_resource r = { c->rsrc[i]->owner->index, c->rsrc[i]->bit_rate, (int64_t)c->rsrc[i]->len };
		memcpy(p, &r, sizeof r); p += sizeof r;
	}
	for (int i = 0; i < c->n_copies; i++) {
		struct ns_sc_copy cp = { c->copies[i].node, c->copies[i].res };
		memcpy(p, &cp, sizeof cp); p += sizeof cp;
	}
	/* group_off = {0, 0}: the module itself is the script */
	struct ns_dev_opts o;
	memset(&o, 0, sizeof o);
	/* the module is the script: nothing is known about how many flows it will establish, so the tables are
	 * sized by the environment (defaults fit a p2p.c run of a few thousand clients) */
	const char *ev;
	o.device = c->device; o.mode = NS_MODE_SEQ; o.want_records = c->want_records;
	o.features_off = 8;                                     /* NS_FEAT_INIT_REC: sim_create_node wrote the R_NODE_CREATE records itself, interleaved with the module's own */
	o.cap_flows = (ev = getenv("NS_CAP_FLOWS")) ? atoi(ev) : 1 << 17;      /* concurrent flows */
	o.cap_ord = (ev = getenv("NS_CAP_ORD")) ? atoi(ev) : 1 << 21;          /* flows ever created */
	o.cap_ranges = 2*o.cap_flows+c->n_rsrc+c->n_copies+16;
	o.cap_users = (ev = getenv("NS_CAP_USERS")) ? atoi(ev) : 1 << 17;      /* pending module-owned events */
	o.ring_cap = 1u << 22; o.record_cap = 64u << 20;
	int rc = ns_dev_create(&c->dev, buf, len, &o);
	free(buf);
	if (rc) { log_err("device core: %s\n", ns_strerror(rc)); exit(3); }
	for (int i = 0; i < c->n_nodes; i++) {              /* the device drew the same ids in the same order */
		uint32_t id; ns_dev_node_id(c->dev, i, &id);
		if (id != c->nodes[i]->node_id) { log_err("node id stream diverged\n"); exit(3); }
	}
	for (int i = 0; i < c->n_rsrc; i++) {
		uint32_t id; ns_dev_resource_id(c->dev, i, &id);
		if (id != c->rsrc[i]->resource_id) { log_err("resource id stream diverged\n"); exit(3); }
	}
	return 0;
}

static void refresh_flow(struct ns_host_core *c, struct flow *f){
	struct ns_flow_info fi;
	if (ns_dev_flow_info(c->dev, f->ord, &fi) || !fi.alive) return;
	f->speed[0] = fi.speed[0]; f->speed[1] = fi.speed[1];
	f->flow_id = fi.flow_id;
	if (f->drng) { f->drng->start = (size_t)fi.drng_start; f->drng->len = fi.drng_len; f->drng->last_update = c->s->now; }
}
static void refresh_nodes(struct ns_host_core *c){
	double *st = malloc((size_t)c->n_nodes*32); size_t k = 0;
	if (!ns_dev_node_state(c->dev, st, c->n_nodes, &k))
		for (size_t i = 0; i < k; i++) {
			c->nodes[i]->bandwidth_usage[0] = st[4*i]; c->nodes[i]->bandwidth_usage[1] = st[4*i+1];
			c->nodes[i]->total_bwupbound[0] = st[4*i+2]; c->nodes[i]->total_bwupbound[1] = st[4*i+3];
		}
	free(st);
}

void ns_host_refresh_ranges(struct resource *r){
	struct ns_host_core *c = the_core;
	if (!c || !c->dev) return;
	int cap = 64, n = 0;
	struct ns_range_info *ri = malloc((size_t)cap*sizeof *ri);
	for (;;) {                                              /* as many ranges as the resource has */
		if (ns_dev_node_ranges(c->dev, r->owner->index, r->resource_id, ri, cap, &n)) { free(ri); return; }
		if (n < cap) break;
		cap *= 2;
		ri = realloc(ri, (size_t)cap*sizeof *ri);
	}
	for (struct range *g = r->first, *nx; g; g = nx) { nx = g->next; free(g); }
	r->first = NULL;
	struct range **pp = &r->first;
	for (int i = 0; i < n; i++) {
		struct range *g = talloc(1, struct range);
		g->start = (size_t)ri[i].start; g->total_len = (size_t)ri[i].total_len; g->len = ri[i].len;
		g->owner = r; g->last_update = c->s->now;
		g->producer = (ri[i].producer_ord >= 0 && ri[i].producer_ord < c->n_flows) ? c->flows[ri[i].producer_ord] : NULL;
		*pp = g; pp = &g->next;
	}
	free(ri);
}

/* sim_establish_flow, sim.c:42-94 */
struct flow *sim_establish_flow(id_t rid, size_t start, struct node *src, struct node *dst, struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	materialize(c);
	struct resource *sr = store_get(src->store, rid);
	if (!sr) { log_err("The resource %d doesn't exist on source node %d\n", rid, src->node_id); return NULL; }
	struct resource *dr = store_get(dst->store, rid);
	if (!dr) { dr = store_add(dst, rid, sr->len, sr->script_index); dr->bit_rate = sr->bit_rate; }
	struct ns_sc_op op;
	memset(&op, 0, sizeof op);
	op.kind = NS_OP_CREATE; op.a = src->index; op.b = dst->index; op.c = sr->script_index; op.start = (int64_t)start;
	op.x = s->bwcalc(src->user_data, dst->user_data);      /* flow.c:303-304 */
	op.y = s->dlycalc(src->user_data, dst->user_data);     /* flow.c:309 */
	int rc = ns_dev_exec(c->dev, &op, 1);
	if (rc) { log_err("device core: %s\n", ns_strerror(rc)); exit(3); }
	GROW(c->flows, c->n_flows, c->cap_flows);
	if (op.a < 0) {                                        /* sim.c:62-67 */
		log_err("Trying to create a flow to a existing range.\n");
		c->flows[c->n_flows++] = NULL;
		return NULL;
	}
	struct flow *f = talloc(1, struct flow);
	f->ord = op.a;
	while (c->n_flows < f->ord) { GROW(c->flows, c->n_flows, c->cap_flows); c->flows[c->n_flows++] = NULL; }
	GROW(c->flows, c->n_flows, c->cap_flows);
	c->flows[c->n_flows++] = f;
	f->peer[0] = src; f->peer[1] = dst; f->dst_id = dst->node_id;
	f->out_next = src->outs; src->outs = f;
	f->bwupbound = op.x; f->delay = op.y; f->resource_id = rid;
	f->start = start; f->begin_time = s->now;
	f->drng = talloc(1, struct range);
	f->drng->owner = dr; f->drng->total_len = dr->len; f->drng->producer = f; f->drng->start = start;
	sr->nconsumer++; dr->nproducer++;
	refresh_flow(c, f);
	return f;
}

/* flow_close (flow.c:241-292) called by a module */
void flow_close(struct flow *f, struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	if (!f || f->closed) return;
	struct ns_sc_op op;
	memset(&op, 0, sizeof op);
	op.kind = NS_OP_CLOSE; op.a = f->ord;
	ns_dev_exec(c->dev, &op, 1);
	f->closed = true;
}

bool is_connected(struct node *src, struct node *dst){    /* flow.h:29-35: src->outs, keyed by the destination */
	struct flow **pp = &src->outs;
	while (*pp) {
		if ((*pp)->closed) { *pp = (*pp)->out_next; continue; }       /* unlink what has been closed since */
		if ((*pp)->peer[1] == dst) return true;
		pp = &(*pp)->out_next;
	}
	return false;
}

/* event_add, event.h:56-74, for module-owned events */
void event_add(struct event *e, struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	if (!e || e->active) return;
	materialize(c);
	if (e->token < 0) { GROW(c->uev, c->n_uev, c->cap_uev); e->token = c->n_uev; c->uev[c->n_uev++] = e; }
	struct ns_sc_op op;
	memset(&op, 0, sizeof op);
	op.kind = NS_OP_SCHED; op.a = e->token; op.c = 0; op.x = e->time;
	int rc = ns_dev_exec(c->dev, &op, 1);
	if (rc == NS_E_BACK_IN_TIME) { log_err("Add event back in time, e: %lf now: %lf\n", e->time, s->now); abort(); }
	if (rc == NS_E_NAN_TIME) { log_err("Add event at nan\n"); abort(); }
	e->active = true;
	e->dev_slot = op.b;
}
/* event_remove, event.h:13-18, of a module-owned event */
void event_remove(struct event *e){
	struct ns_host_core *c = the_core;
	if (!e || !e->active) return;
	e->active = false;
	if (!c || !c->dev || e->dev_slot < 0) return;
	struct ns_sc_op op;
	memset(&op, 0, sizeof op);
	op.kind = NS_OP_CANCEL; op.a = e->dev_slot;
	ns_dev_exec(c->dev, &op, 1);
	e->dev_slot = -1;
}

void sim_send_packet(void *data, int len, struct node *src, struct node *dst, struct sim_state *s){
	(void)data; (void)len; (void)src; (void)dst; (void)s;   /* dead in the reference too: no PACKET_DONE handler */
}
void sim_node_change_state(struct node *n, enum node_state state, struct sim_state *s){   /* sim.h:25-28 */
	n->state = state;
	uint8_t v = (uint8_t)state;
	write_record(0, R_NODE_STATE, n->node_id, 1, &v, s);
}

void user_speed_change(struct event *e, struct sim_state *s){ (void)e; (void)s; }   /* user.c:7-13 */
void user_done(struct event *e, struct sim_state *s){ (void)e; (void)s; }

static void rec_reserve(struct ns_host_core *c, size_t more);
static void pull_records(struct ns_host_core *c);
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
/* open_record, record.c:119-153: remember the file; the device's record stream is saved there */
void open_record(const char *filename, int create, struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	(void)create;
	if (c->record_path) return;                             /* no-op if already open, record.c:120-123 */
	c->record_path = strdup(filename);
	c->want_records = 1;
}

/* append what the device has produced since the last pull to the host copy of the record stream */
static void rec_reserve(struct ns_host_core *c, size_t more){
	if (c->rec_len+more <= c->rec_cap) return;
	while (c->rec_len+more > c->rec_cap) c->rec_cap = c->rec_cap ? c->rec_cap*2 : 1 << 20;
	c->rec = realloc(c->rec, c->rec_cap);
}
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
static void save_records(struct ns_host_core *c){
	if (!c->record_path || !c->dev) return;
	pull_records(c);
	size_t total = c->rec_len;
	uint8_t *buf = malloc(total+4);
	memcpy(buf+4, c->rec, total);
	buf[0] = total >> 24; buf[1] = total >> 16; buf[2] = total >> 8; buf[3] = total;   /* record.c:111-115 */
	size_t size = 1024*1024;                                /* record.c:133 initial size, doubled at :38-51 */
	while (4+total+19 >= size) size <<= 1;
	FILE *f = fopen(c->record_path, "wb");
	if (f) {
		fwrite(buf, 1, total+4, f);
		if (ftruncate(fileno(f), (off_t)size)) { }
		fclose(f);
	}
	free(buf);
}

/* the event loop, main.c:50-67, with the default handlers (main.c:44-48) on the device */
int ns_host_run(struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	materialize(c);
	uint32_t mask = 1u << USER;
	for (int t = 0; t < LAST_EVENT; t++) if (c->handlers[t]) mask |= 1u << t;
	ns_dev_set_host_mask(c->dev, mask);
	int rc = 0;
	while (!s->exit) {
		struct ns_host_event ev;
		rc = ns_dev_run_host(c->dev, &ev);
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
		tmp.type = ev.type; tmp.time = ev.time; tmp.token = -1;
		bool need_free = false;
		refresh_nodes(c);
		if (f) refresh_flow(c, f);
		switch (ev.type) {
		case SPEED_CHANGE:
			se.e = e; se.f = f; se.speed = ev.speed; se.type = ev.dir;
			tmp.data = &se;
			break;
		case USER:
			e = (ev.token >= 0 && ev.token < c->n_uev) ? c->uev[ev.token] : NULL;
			if (!e) continue;
			e->active = false;                              /* event_pop, event.h:20-25 */
			need_free = e->auto_free;
			break;
		default:
			tmp.data = f;
		}
		for (struct handler *h = c->handlers[ev.type]; h; h = h->next) h->f(e, s);
		if (ev.type == FLOW_DONE && f) f->closed = true;    /* flow_done_cleaner runs on the device next */
		if (need_free) { c->uev[ev.token] = NULL; free(e); }   /* main.c:65-66 */
		if (s->exit) {
			struct ns_sc_op op;
			memset(&op, 0, sizeof op);
			op.kind = NS_OP_END;
			ns_dev_exec(c->dev, &op, 1);
		}
	}
	/* let a pending flow_done_cleaner finish so the record stream is complete */
	save_records(c);
	return rc;
}

void ns_host_finish(struct sim_state *s){
	struct ns_host_core *c = core_of(s);
	if (c->dev) ns_dev_destroy(c->dev);
	c->dev = NULL;
}
