/* ref_scenario.c — TEST INFRASTRUCTURE, not product code.
 *
 * A behaviour module for the UNMODIFIED reference core (built by oracle/Makefile into
 * oracle/_ref/modules/scenario.so and loaded by oracle/_ref/netsim exactly like modules/test05:
 * main.c:25-38 dlopen()s "<path>.so" and calls "<basename>_init").  It replays a scenario script
 * (include/ns_scenario.h) through the reference's public module API only:
 *   sim_create_node / sim_node_new_resource / sim_node_add_resource / sim_establish_flow (sim.h:14-23),
 *   sim_register_handler (sim.c:96-111), event_new + event_add (data.h:188-198, event.h:56-74),
 *   flow_close (flow.h:13), open_record (record.h:22), sim_end (sim.h:42-45).
 * The record file it leaves behind (plus an optional fp64 event trace) is the ground truth that
 * pins oracle/ns_oracle.c and the CUDA engine.
 *
 * Environment: NS_SCENARIO (input file), NS_RECORD (record file, default "ns_record"),
 * NS_TRACE (optional binary trace: 32-byte rows {u32 type,u32 ord,f64 time,f64 v0,f64 v1}),
 * NS_MAX_EVENTS (stop after that many popped events), NS_STEP_EVENTS (wall clock every that many
 * events, for bench.py's reference arm), NS_STATS (json with counts and timing).
 */
#define _GNU_SOURCE
#include <unistd.h>
#include <time.h>
#include <stdint.h>

#define LOG_DOMAIN "scenario"

#include "log.h"
#include "data.h"
#include "sim.h"
#include "store.h"
#include "user.h"
#include "event.h"
#include "flow.h"
#include "record.h"
#include "uthash.h"
#include "client_behaviour.h"   /* the reference's behaviour library, client side: modules/client_behaviour.{h,c}, p2p_common.h */

#include "ns_scenario.h"

struct fl_ent { struct flow *f; int ord; UT_hash_handle hh; };

static struct ns_sc_header H;
static double *max_up, *max_down;
static struct ns_sc_resource *rsrc;
static struct ns_sc_copy *copies;
static int32_t *group_off;
static struct ns_sc_op *ops;
static struct ns_sc_ondone *ondone;

static struct node **nodes;
static id_t *rids;
static struct flow **flows;   /* by creation ordinal; NULL once closed / never created */
static int nflows, cap_flows;
static struct fl_ent *by_ptr;
static double cur_ub, cur_delay;
static FILE *trace;
static uint64_t n_events, max_events, n_by_type[LAST_EVENT], n_create, n_close, n_throttle_taken;
static uint64_t step_events; static double *step_t; static int n_steps, cap_steps;
static struct timespec t_init_done;
static const char *stats_path;
static struct sim_state *S;

/* scripts with NS_SC_F_CLIENTS: nodes carry the reference's struct def_user (first member, so that the library's casts work) */
struct cl_user { struct def_user d; int index; };
static struct ns_sc_client *clients;
static int has_clients, in_play;

/* s->bwcalc / s->dlycalc (flow.c:303-309): explicit per-op values for NS_OP_CREATE, the reference's own distance rules
 * (p2p_common.h:181-189,214-221) for the flows client_new_play1 establishes */
static double sc_bw(void *a, void *b){ return in_play ? distance_based_bw(a, b) : cur_ub; }
static double sc_dly(void *a, void *b){ return in_play ? distance_based_delay(a, b) : cur_delay; }
static int is_client(struct node *n){ return has_clients && n->user_data && ((struct def_user *)n->user_data)->nt == N_IDLE; }

static void *xread(FILE *f, size_t n, size_t sz){
	void *p = calloc(n ? n : 1, sz);
	if (n && fread(p, sz, n, f) != n) { fprintf(stderr, "scenario: short read\n"); exit(2); }
	return p;
}

static int ord_of(struct flow *f){
	struct fl_ent *e = NULL;
	HASH_FIND_PTR(by_ptr, &f, e);
	return e ? e->ord : -1;
}

static void forget(struct flow *f){
	struct fl_ent *e = NULL;
	HASH_FIND_PTR(by_ptr, &f, e);
	if (e) { flows[e->ord] = NULL; HASH_DEL(by_ptr, e); free(e); }
}

static void trace_row(uint32_t type, uint32_t ord, double t, double v0, double v1){
	if (!trace) return;
	struct { uint32_t type, ord; double t, v0, v1; } r = { type, ord, t, v0, v1 };
	fwrite(&r, sizeof r, 1, trace);
}

static void note_flow(struct flow *f){
	if (nflows == cap_flows) {
		cap_flows = cap_flows ? cap_flows*2 : 1024;
		flows = realloc(flows, cap_flows*sizeof *flows);
	}
	int ord = nflows++;
	flows[ord] = f;
	if (f) {
		struct fl_ent *e = calloc(1, sizeof *e);
		e->f = f; e->ord = ord;
		HASH_ADD_PTR(by_ptr, f, e);
		n_create++;
	}
}

static void run_group(int g, struct sim_state *s){
	for (int i = group_off[g]; i < group_off[g+1]; i++) {
		struct ns_sc_op *o = &ops[i];
		switch (o->kind) {
		case NS_OP_PLAY1: {
			/* the reference's own client_new_play1 (client_behaviour.c:364-387): server_picker_opt1 + qselect_eval,
			 * client_new_connection per kept server, client_start_play */
			struct def_sim *ds = s->user_data;
			struct node *cl = nodes[o->a];
			struct list_head *old = cl->conns[1].next;
			ds->fetch_metric = o->b ? share_metric : distance_metric;
			in_play = 1;
			client_new_play1(rids[o->c], cl, s);
			in_play = 0;
			/* the flows it established are at the head of the client's receive list, newest first: they take the next
			 * creation ordinals in creation order */
			int k = 0;
			for (struct list_head *h = cl->conns[1].next; h != &cl->conns[1] && h != old; h = h->next) k++;
			struct flow **nf = calloc(k ? k : 1, sizeof *nf);
			int j = k;
			for (struct list_head *h = cl->conns[1].next; h != &cl->conns[1] && h != old; h = h->next)
				nf[--j] = list_entry(h, struct flow, conns[1]);
			for (j = 0; j < k; j++) note_flow(nf[j]);
			free(nf);
			break; }
		case NS_OP_CREATE: {
			cur_ub = o->x; cur_delay = o->y;
			struct flow *f = sim_establish_flow(rids[o->c], (size_t)o->start, nodes[o->a], nodes[o->b], s);
			note_flow(f);
			break; }
		case NS_OP_SCHED: {
			struct event *e = event_new(o->c ? s->now+o->x : o->x, USER, (void *)(intptr_t)o->a);
			e->auto_free = true;
			event_add(e, s);
			break; }
		case NS_OP_CLOSE:
			if (o->a < nflows && flows[o->a]) {
				struct flow *f = flows[o->a];
				forget(f);
				flow_close(f, s);
				n_close++;
			}
			break;
		case NS_OP_END:
			sim_end(s);
			break;
		}
	}
}

static double since_init(void);
static void count_event(struct event *e, struct sim_state *s){
	n_events++;
	n_by_type[e->type]++;
	if (step_events && n_events % step_events == 0) {         /* NS_STEP_EVENTS: wall clock per window */
		if (n_steps == cap_steps) { cap_steps = cap_steps ? cap_steps*2 : 64; step_t = realloc(step_t, cap_steps*sizeof *step_t); }
		step_t[n_steps++] = since_init();
	}
	if (max_events && n_events >= max_events)
		sim_end(s);
}

static void on_user(struct event *e, struct sim_state *s){
	if (has_clients && (uintptr_t)e->data >= (uintptr_t)H.n_groups+4096) {
		/* a client's own event d->e (client_behaviour.c:31-34): what modules/p2p.c:59-67 does with it */
		struct user_event *ue = e->data;
		struct cl_user *cu = (struct cl_user *)ue->d;
		trace_row(USER, 0x40000000u|(uint32_t)cu->index, s->now, (double)ue->type, 0);
		client_next_state_from_event(e, s);
		client_handle_next_state(cu->d.n, s);
		count_event(e, s);
		return;
	}
	int g = (int)(intptr_t)e->data;
	trace_row(USER, g, s->now, 0, 0);
	run_group(g, s);
	count_event(e, s);
}

static void on_speed(struct event *e, struct sim_state *s){
	struct spd_event *se = e->data;
	if (trace)
		trace_row(SPEED_CHANGE | (se->type << 8), ord_of(se->f), s->now, se->speed, se->f->speed[se->type]);
	if (is_client(se->f->peer[1])) client_speed_change(e, s);          /* modules/p2p.c:184 */
	count_event(e, s);
}

static int ondone_group(int ord){
	int lo = 0, hi = H.n_ondone-1;
	while (lo <= hi) {
		int m = (lo+hi)/2;
		if (ondone[m].flow_ord == ord) return ondone[m].group;
		if (ondone[m].flow_ord < ord) lo = m+1; else hi = m-1;
	}
	return -1;
}

static void on_done(struct event *e, struct sim_state *s){
	struct flow *f = e->data;
	int ord = ord_of(f);
	trace_row(FLOW_DONE, ord, s->now, f->drng->len, f->speed[1]);
	if (is_client(f->peer[1])) client_done(e, s);                      /* modules/p2p.c:113-120 */
	int g = H.n_ondone ? ondone_group(ord) : -1;
	if (g >= 0)
		run_group(g, s);
	forget(f);   /* flow_done_cleaner (flow.c:403-406) closes it right after us */
	n_close++;
	count_event(e, s);
}

static void on_throttle(struct event *e, struct sim_state *s){
	struct flow *f = e->data;
	trace_row(FLOW_SPEED_THROTTLE, ord_of(f), s->now, f->speed[0], f->speed[1]);
	if (is_client(f->peer[1])) client_speed_throttle(e, s);            /* modules/p2p.c:182 */
	count_event(e, s);
}

static void on_drain(struct event *e, struct sim_state *s){
	struct flow *f = e->data;
	trace_row(FLOW_DRAIN, ord_of(f), s->now, 0, 0);
	count_event(e, s);
}

static double since_init(void){
	struct timespec t1;
	clock_gettime(CLOCK_MONOTONIC, &t1);
	return (t1.tv_sec-t_init_done.tv_sec)+1e-9*(t1.tv_nsec-t_init_done.tv_nsec);
}

static void finish(void){
	double el = since_init();
	if (trace) fclose(trace);
	if (!stats_path) return;
	FILE *f = fopen(stats_path, "w");
	if (!f) return;
	/* solves = bwspread invocations (flow.c:35): 1 per SPEED_CHANGE (flow.c:361), 1 per create
	 * (flow.c:318), 2 per close (flow.c:249-250); throttle solves are counted by the oracle only. */
	fprintf(f, "{\"events\": %llu, \"speed_change\": %llu, \"flow_done\": %llu, \"drain\": %llu, "
		   "\"throttle\": %llu, \"user\": %llu, \"creates\": %llu, \"closes\": %llu, "
		   "\"loop_seconds\": %.9f, \"now\": %.17g, \"step_times\": [",
		(unsigned long long)n_events, (unsigned long long)n_by_type[SPEED_CHANGE],
		(unsigned long long)n_by_type[FLOW_DONE], (unsigned long long)n_by_type[FLOW_DRAIN],
		(unsigned long long)n_by_type[FLOW_SPEED_THROTTLE], (unsigned long long)n_by_type[USER],
		(unsigned long long)n_create, (unsigned long long)n_close, el, S ? S->now : 0.0);
	for (int i = 0; i < n_steps; i++) fprintf(f, "%s%.9f", i ? ", " : "", step_t[i]);
	fprintf(f, "]}\n");
	fclose(f);
}

void scenario_init(struct sim_state *s){
	const char *path = getenv("NS_SCENARIO");
	const char *rec = getenv("NS_RECORD");
	const char *tr = getenv("NS_TRACE");
	const char *me = getenv("NS_MAX_EVENTS");
	stats_path = getenv("NS_STATS");
	if (!path) { fprintf(stderr, "scenario: NS_SCENARIO not set\n"); exit(2); }
	if (!rec) rec = "ns_record";
	if (me) max_events = strtoull(me, NULL, 10);
	if (getenv("NS_STEP_EVENTS")) step_events = strtoull(getenv("NS_STEP_EVENTS"), NULL, 10);
	FILE *f = fopen(path, "rb");
	if (!f) { perror(path); exit(2); }
	if (fread(&H, sizeof H, 1, f) != 1 || H.magic != NS_SC_MAGIC || H.version != NS_SC_VERSION) {
		fprintf(stderr, "scenario: bad header\n"); exit(2);
	}
	max_up = xread(f, H.n_nodes, sizeof(double));
	max_down = xread(f, H.n_nodes, sizeof(double));
	rsrc = xread(f, H.n_resources, sizeof *rsrc);
	copies = xread(f, H.n_copies, sizeof *copies);
	group_off = xread(f, H.n_groups+1, sizeof(int32_t));
	ops = xread(f, H.n_ops, sizeof *ops);
	ondone = xread(f, H.n_ondone, sizeof *ondone);
	has_clients = (H.flags & NS_SC_F_CLIENTS) != 0;
	if (has_clients) clients = xread(f, H.n_nodes, sizeof *clients);
	fclose(f);
	if (tr && *tr) trace = fopen(tr, "wb");

	S = s;
	srandom(H.seed);
	if (strcmp(rec, "-") != 0) {
		unlink(rec);
		open_record(rec, 1, s);
	}
	s->bwcalc = sc_bw;
	s->dlycalc = sc_dly;

	struct def_sim *ds = NULL;
	if (has_clients) {                                                 /* modules/p2p.c:124-141 */
		init_sim_size(s, 20, sizeof(struct def_sim)+16);
		ds = s->user_data;
		ds->eval_size = H.n_nodes+1;
		ds->eval_table = realloc(ds->eval_table, ds->eval_size*sizeof(struct nv_pair));
		ds->fetch_metric = distance_metric; ds->push_metric = distance_metric; ds->new_conn_cb = NULL;
	}
	nodes = calloc(H.n_nodes, sizeof *nodes);
	for (int i = 0; i < H.n_nodes; i++) {
		nodes[i] = sim_create_node(s);
		nodes[i]->user_data = NULL;
		nodes[i]->maximum_bandwidth[0] = max_up[i];
		nodes[i]->maximum_bandwidth[1] = max_down[i];
		if (has_clients && clients[i].kind != NS_CK_PLAIN) {
			struct cl_user *cu = calloc(1, sizeof *cu);
			cu->index = i;
			cu->d.n = nodes[i];
			cu->d.time_zone = clients[i].time_zone;
			cu->d.lowwm = clients[i].lowwm; cu->d.highwm = clients[i].highwm;
			nodes[i]->user_data = &cu->d;
			if (clients[i].kind == NS_CK_SERVER) {                     /* p2p_new_server, p2p_common.h:101-115 */
				sim_node_change_state(nodes[i], N_SERVER, s);
				cu->d.next_state = nodes[i]->state;
				cu->d.nt = N_SERVER;
				struct server *ss = talloc(1, struct server);
				ss->n = nodes[i];
				list_add(&ss->servers, &ds->servers);
				ds->nsvr++;
			} else {                                                   /* modules/p2p.c:166-174 */
				cu->d.nt = N_IDLE;
				nodes[i]->state = N_IDLE;
			}
		}
	}
	rids = calloc(H.n_resources ? H.n_resources : 1, sizeof *rids);
	struct resource **rp = calloc(H.n_resources ? H.n_resources : 1, sizeof *rp);
	for (int i = 0; i < H.n_resources; i++) {
		rp[i] = sim_node_new_resource(nodes[rsrc[i].owner], (size_t)rsrc[i].len);
		rp[i]->bit_rate = rsrc[i].bit_rate;
		rids[i] = rp[i]->resource_id;
	}
	for (int i = 0; i < H.n_copies; i++)
		sim_node_add_resource(nodes[copies[i].node], rp[copies[i].res]);
	free(rp);

	sim_register_handler(FLOW_DONE, HNDR_USER, on_done, s);
	sim_register_handler(USER, HNDR_USER, on_user, s);
	sim_register_handler(SPEED_CHANGE, HNDR_USER, on_speed, s);
	sim_register_handler(FLOW_SPEED_THROTTLE, HNDR_USER, on_throttle, s);
	sim_register_handler(FLOW_DRAIN, HNDR_USER, on_drain, s);

	run_group(0, s);

	atexit(finish);
	clock_gettime(CLOCK_MONOTONIC, &t_init_done);
}
