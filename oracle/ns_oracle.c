/* ns_oracle.c — TEST INFRASTRUCTURE, not product code.
 *
 * Sequential CPU restatement of the reference's flow/event hot path, over index-linked tables
 * instead of the reference's calloc'd structs.  Every function cites the reference lines it
 * follows (paths relative to /root/reference).  It executes scenario scripts
 * (include/ns_scenario.h) exactly as oracle/ref_scenario.c drives the real core, and produces the
 * same record byte stream (record.c:27-117).
 *
 * PARITY PINNED: tests/test_oracle_golden.py runs the compiled reference (oracle/_ref) on the
 * restated test00..test05 modules and on randomized scripts and requires byte-identical record
 * payloads and bit-identical fp64 event traces; tests/golden/ holds the vectors for boxes
 * without /root/reference.
 *
 * Floating point: compile with -ffp-contract=off (no FMA), same as the reference build recipe.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ns_scenario.h"
#include "ns_oracle.h"

static const double eps = 1e-4;               /* sim.c:11 */
#define fequ(a, b) ((a)-(b) > -eps && (a)-(b) < eps)   /* common.h:18 */

enum { SND = 0, RCV = 1 };
/* data.h:114-122 */
enum { FLOW_DRAIN = 0, FLOW_DONE = 1, FLOW_SPEED_THROTTLE = 2, PACKET_DONE = 3, SPEED_CHANGE = 4, USER = 5 };
/* record.h:7-17 */
enum { R_USAGE = 0, R_SPD = 2, R_CONN_CREATE = 4, R_CONN_DST = 5, R_CONN_CLOSE = 6, R_NODE_STATE = 7, R_NODE_CREATE = 8 };
enum { K_DONE = 0, K_DRAIN = 1, K_SPD = 2, K_USER = 3 };

struct o_node { double max[2], usage[2], total[2]; uint32_t id; int conn_head[2]; int res_head; };
struct o_res { uint32_t id; int64_t len; int node, first, nprod, ncons, next_in_node; };
struct o_range { int64_t start, total; double len, lenc, last; int prod, res, next, prev, cons_head, next_free; };
struct o_flow {
	double ub, speed[2], delay;
	int peer[2]; uint32_t id; int cnext[2], cprev[2];
	int srng, drng, cons_next, cons_prev;
	double done_t, drain_t; uint64_t done_seq, drain_seq;
	uint8_t done_act, drain_act, drain_type, alive;
	int spd_head, ord, next_free;
};
struct o_spd { double time, speed; uint64_t seq; int flow; uint8_t dir, active; int next, prev, next_free; };
struct o_user { double time; uint64_t seq; int group; uint8_t active; int next_free; };
struct o_hent { double time; uint64_t seq; int kind, idx; };

#define VEC(T) struct { T *v; int n, cap; }
#define VPUSH(vec) (((vec).n == (vec).cap ? ((vec).cap = (vec).cap ? (vec).cap*2 : 64, \
	(vec).v = realloc((vec).v, (size_t)(vec).cap*sizeof *(vec).v)) : 0), &(vec).v[(vec).n++])

struct ns_oracle {
	struct ns_sc_header H;
	const int32_t *group_off; const struct ns_sc_op *ops; const struct ns_sc_ondone *ondone;
	void *blob;
	double now; uint64_t seq; struct ns_rand rng; int exit_flag, status;
	VEC(struct o_node) nodes; VEC(struct o_res) res; VEC(struct o_range) ranges; VEC(struct o_flow) flows;
	VEC(struct o_spd) spd; VEC(struct o_user) users; VEC(struct o_hent) heap;
	VEC(int) ord2flow;             /* creation ordinal -> flow slot or -1 */
	VEC(uint32_t) rids;
	int free_flow, free_range, free_spd, free_user;
	uint8_t *rec; size_t rec_n, rec_cap; int want_rec;
	struct ns_trace_row *tr; size_t tr_n, tr_cap; int want_tr;
	uint64_t c[NS_C_COUNT];
};

/* ---------------------------------------------------------------- glibc random() TYPE_3 */
void ns_rand_seed(struct ns_rand *g, uint32_t seed){
	int32_t word = seed ? (int32_t)seed : 1;
	g->st[0] = word;
	for (int i = 1; i < 31; i++) {
		long hi = word/127773, lo = word%127773;
		long w = 16807*lo-2836*hi;
		if (w < 0) w += 2147483647;
		word = (int32_t)w;
		g->st[i] = word;
	}
	g->f = 3; g->r = 0;
	for (int i = 0; i < 310; i++) ns_rand_next(g);
}
uint32_t ns_rand_next(struct ns_rand *g){
	uint32_t v = (uint32_t)g->st[g->f]+(uint32_t)g->st[g->r];
	g->st[g->f] = (int32_t)v;
	if (++g->f >= 31) g->f = 0;
	if (++g->r >= 31) g->r = 0;
	return v >> 1;
}

/* ---------------------------------------------------------------- records (record.c:27-117) */
static void rec_put(ns_oracle *o, const void *p, size_t n){
	if (o->rec_n+n > o->rec_cap) {
		o->rec_cap = o->rec_cap ? o->rec_cap*2 : 4096;
		while (o->rec_cap < o->rec_n+n) o->rec_cap *= 2;
		o->rec = realloc(o->rec, o->rec_cap);
	}
	memcpy(o->rec+o->rec_n, p, n);
	o->rec_n += n;
}
static void write_record(ns_oracle *o, uint8_t minor, uint32_t id, int bytes, uint32_t ival, double dval){
	if (!o->want_rec) return;
	uint8_t b[19]; int n = 0;
	b[n++] = 0; b[n++] = minor;                                          /* record.c:54-58 */
	b[n++] = id>>24; b[n++] = id>>16; b[n++] = id>>8; b[n++] = id;       /* :60 */
	uint32_t t = (uint32_t)(int)o->now;                                  /* :66 */
	b[n++] = t>>24; b[n++] = t>>16; b[n++] = t>>8; b[n++] = t;
	uint16_t ut = (uint16_t)(int)((o->now-((int)o->now))*1000);          /* :69 */
	b[n++] = ut>>8; b[n++] = ut;
	b[n++] = (uint8_t)(int8_t)bytes;                                     /* :74 */
	if (bytes == 1) b[n++] = (uint8_t)ival;                              /* :86-89 */
	else if (bytes == 4) { b[n++] = ival>>24; b[n++] = ival>>16; b[n++] = ival>>8; b[n++] = ival; }
	else {                                                               /* :98-106 */
		double tmp = dval;
		uint32_t ip = (uint32_t)(int)tmp;
		b[n++] = ip>>24; b[n++] = ip>>16; b[n++] = ip>>8; b[n++] = ip;
		tmp = tmp-((int)tmp);
		uint16_t fp = (uint16_t)(int)(tmp*1000);
		b[n++] = fp>>8; b[n++] = fp;
	}
	rec_put(o, b, n);
}
static void write_usage(ns_oracle *o, int dir, int n){                   /* record_wrapper.h:6-9 */
	write_record(o, R_USAGE|dir, o->nodes.v[n].id, -1, 0, o->nodes.v[n].usage[dir]);
}
static void trace_row(ns_oracle *o, uint32_t type, uint32_t ord, double v0, double v1){
	if (!o->want_tr) return;
	if (o->tr_n == o->tr_cap) {
		o->tr_cap = o->tr_cap ? o->tr_cap*2 : 1024;
		o->tr = realloc(o->tr, o->tr_cap*sizeof *o->tr);
	}
	struct ns_trace_row r = { type, ord, o->now, v0, v1 };
	o->tr[o->tr_n++] = r;
}

/* ---------------------------------------------------------------- event queue
 * event.h:13-88 over skiplist.h: ordered by time, FIFO among equal times (event_cmp never
 * returns 0, so an insert walks past every element with time <= key).  Restated as a binary
 * heap on (time, insertion seq) with lazy deletion; each successful insert burns one random()
 * (skiplist.h:34-40 via :83). */
static int hless(const struct o_hent *a, const struct o_hent *b){
	return a->time < b->time || (a->time == b->time && a->seq < b->seq);
}
static void heap_push(ns_oracle *o, double time, uint64_t seq, int kind, int idx){
	struct o_hent *e = VPUSH(o->heap);
	int i = o->heap.n-1;
	struct o_hent x = { time, seq, kind, idx };
	(void)e;
	while (i > 0) {
		int p = (i-1)/2;
		if (!hless(&x, &o->heap.v[p])) break;
		o->heap.v[i] = o->heap.v[p];
		i = p;
	}
	o->heap.v[i] = x;
}
static struct o_hent heap_pop(ns_oracle *o){
	struct o_hent top = o->heap.v[0], x = o->heap.v[--o->heap.n];
	int i = 0, n = o->heap.n;
	while (1) {
		int l = 2*i+1, r = l+1, m;
		if (l >= n) break;
		m = (r < n && hless(&o->heap.v[r], &o->heap.v[l])) ? r : l;
		if (!hless(&o->heap.v[m], &x)) break;
		o->heap.v[i] = o->heap.v[m];
		i = m;
	}
	if (n) o->heap.v[i] = x;
	return top;
}
/* event_add, event.h:56-74 */
static void ev_add(ns_oracle *o, double *time, uint8_t *active, uint64_t *seq, int kind, int idx){
	if (*active) return;
	if (*time < o->now) {
		if (*time < o->now-eps) { if (!o->status) o->status = NS_O_BACK_IN_TIME; return; }
		*time = o->now;
	}
	if (isnan(*time)) { if (!o->status) o->status = NS_O_NAN_TIME; return; }
	*active = 1;
	*seq = o->seq++;
	ns_rand_next(&o->rng); o->c[NS_C_DRAWS]++;
	o->c[NS_C_PENDING]++;
	heap_push(o, *time, *seq, kind, idx);
}
static void ev_remove(ns_oracle *o, uint8_t *active){                    /* event.h:13-18 */
	if (!*active) return;
	*active = 0;
	o->c[NS_C_PENDING]--;
}
static int is_later_than(double time, uint8_t active, double etime){     /* event.h:82-88 */
	if (!active) return 0;
	return time > etime+eps;
}

/* ---------------------------------------------------------------- rate model */
static double get_share(ns_oracle *o, int f, int dir){                   /* flow.h:37-41 */
	struct o_flow *fl = &o->flows.v[f];
	struct o_node *n = &o->nodes.v[fl->peer[dir]];
	double total = n->total[dir], max = n->max[dir];
	return total > max ? fl->ub*max/total : fl->ub;
}

static void queue_speed_event(ns_oracle *o, int f, int dir, double speed){ /* flow.c:16-29 */
	int i;
	if (o->free_spd >= 0) { i = o->free_spd; o->free_spd = o->spd.v[i].next_free; }
	else { VPUSH(o->spd); i = o->spd.n-1; }
	struct o_spd *se = &o->spd.v[i];
	struct o_flow *fl = &o->flows.v[f];
	se->dir = dir; se->speed = speed; se->flow = f; se->active = 0;
	se->time = o->now+fl->delay;
	se->prev = -1; se->next = fl->spd_head;                              /* list_add head */
	if (fl->spd_head >= 0) o->spd.v[fl->spd_head].prev = i;
	fl->spd_head = i;
	ev_add(o, &se->time, &se->active, &se->seq, K_SPD, i);
}
static void spd_unlink(ns_oracle *o, int i){
	struct o_spd *se = &o->spd.v[i];
	struct o_flow *fl = &o->flows.v[se->flow];
	if (se->prev >= 0) o->spd.v[se->prev].next = se->next; else fl->spd_head = se->next;
	if (se->next >= 0) o->spd.v[se->next].prev = se->prev;
}
static void spd_free(ns_oracle *o, int i){
	o->spd.v[i].next_free = o->free_spd; o->free_spd = i;
}

static void range_calc_and_requeue_events(ns_oracle *o, int f);

/* flow.c:35-205 */
static void bwspread(ns_oracle *o, int f, double amount, int dir, int close){
	struct o_flow *F = o->flows.v;
	struct o_flow *fl = &F[f];
	o->c[NS_C_SOLVES]++;
	if (close) amount = -fl->speed[dir];                                 /* :51-54 */
	int ni = fl->peer[dir];
	struct o_node *n = &o->nodes.v[ni];
	double total = n->total[dir], max = n->max[dir];                     /* :59 */
	double used = n->usage[dir];
	double share = total > max ? fl->ub*max/total : fl->ub;              /* :61 */
	if (amount > eps && fl->speed[dir]+amount > share) {                 /* :64 */
		if (fl->speed[dir] > share) {
			queue_speed_event(o, f, !dir, fl->speed[dir]);               /* :72 */
			return;
		}
		queue_speed_event(o, f, !dir, share);                            /* :75 */
		amount = share-fl->speed[dir];
	}
	if (amount > -eps && amount < eps) return;                           /* :80-83 */
	fl->speed[dir] += amount;                                            /* :85 */
	if (amount < eps && amount > -64) {                                  /* :86-91 */
		n->usage[dir] += amount;
		write_usage(o, dir, ni);
		return;
	}
	if (total < max+eps) {                                               /* :94-109 */
		if (!close || total+fl->ub < max+eps) {
			n->usage[dir] += amount;
			write_usage(o, dir, ni);
			return;
		}
	}
	double spread_amount = amount;                                       /* :112 */
	if (amount > -eps) {
		if (used+amount < max+eps) {                                     /* :114-119 */
			n->usage[dir] += amount;
			write_usage(o, dir, ni);
			return;
		} else
			spread_amount = amount-max+used;                             /* :122 */
	}
	o->c[NS_C_FULL_SPREADS]++;
	double e = 0;                                                        /* :128-145 */
	for (int nf = n->conn_head[dir]; nf >= 0; nf = F[nf].cnext[dir]) {
		o->c[NS_C_FLOW_VISITS]++;
		if (nf == f) continue;
		double lshare = F[nf].ub*max/total;
		if (amount < eps) {
			lshare = get_share(o, nf, !dir);
			if (F[nf].speed[dir] < lshare)
				e += lshare-F[nf].speed[dir];
		} else if (amount > eps && F[nf].speed[dir] > lshare)
			e += F[nf].speed[dir]-lshare;
	}
	if (amount < eps && -amount > e)                                     /* :149-154 */
		amount = -e;
	double new_use = 0;                                                  /* :156-200 */
	for (int nf = n->conn_head[dir]; nf >= 0; nf = F[nf].cnext[dir]) {
		o->c[NS_C_FLOW_VISITS]++;
		if (nf == f) { new_use += fl->speed[dir]; continue; }
		double lshare = F[nf].ub*max/total;
		double delta;
		if (amount < eps) {
			lshare = get_share(o, nf, !dir);
			if (F[nf].speed[dir] < lshare) {
				delta = lshare-F[nf].speed[dir];
				double new_speed = F[nf].speed[dir]-amount*delta/e;      /* :171 */
				if (dir == SND)
					F[nf].speed[dir] = new_speed;                        /* :172-176 */
				queue_speed_event(o, nf, !dir, new_speed);               /* :178 */
			}
		} else if (amount > eps) {
			if (F[nf].speed[dir] > lshare) {
				delta = F[nf].speed[dir]-lshare;
				F[nf].speed[dir] -= spread_amount*delta/e;               /* :187 */
				queue_speed_event(o, nf, !dir, F[nf].speed[dir]);        /* :189 */
				if (dir == RCV)
					range_calc_and_requeue_events(o, nf);                /* :193-194 */
			}
		}
		new_use += F[nf].speed[dir];                                     /* :199 */
	}
	n->usage[dir] = new_use;                                             /* :201-202 */
	write_usage(o, dir, ni);
}

/* ---------------------------------------------------------------- ranges */
static void range_update(ns_oracle *o, int r){                           /* range.h:91-118 */
	struct o_range *R = &o->ranges.v[r];
	if (o->now == R->last) return;
	if (R->prod < 0) { R->last = o->now; return; }
	double delta = o->flows.v[R->prod].speed[1]*(o->now-R->last), t;
	delta = delta-R->lenc;
	t = delta+R->len;
	R->lenc = (t-R->len)-delta;
	R->len = t;
	R->last = o->now;
}
static void flow_range_update(ns_oracle *o, int f){                      /* range.h:120-125 */
	if (f < 0) return;
	struct o_flow *fl = &o->flows.v[f];
	if (fl->srng < 0 || fl->drng < 0) { if (!o->status) o->status = NS_O_NULL_RANGE; return; }
	range_update(o, fl->srng);
	range_update(o, fl->drng);
}

/* range.c:16-79 */
static void range_calc_and_requeue_events(ns_oracle *o, int f){
	if (f < 0) return;
	struct o_flow *fl = &o->flows.v[f];
	ev_remove(o, &fl->done_act);                                         /* :19-20 */
	ev_remove(o, &fl->drain_act);
	flow_range_update(o, f);                                             /* :21 */
	if (o->status) return;
	if (fl->speed[1] < eps) return;                                      /* :26-27 */
	struct o_range *srng = &o->ranges.v[fl->srng], *drng = &o->ranges.v[fl->drng];
	int prod = srng->prod;
	double sgrow = prod >= 0 ? o->flows.v[prod].speed[1] : 0;            /* :33 */
	double fbw = fl->speed[1];
	double drng_start = (double)drng->start+drng->len-(double)srng->start; /* :36 */
	double drain_time = (srng->len-drng_start)/(fbw-sgrow);              /* :38 */
	if (prod < 0) {                                                      /* :40-44 */
		fl->drain_t = o->now+(srng->len-drng_start)/fbw;
		fl->drain_type = FLOW_DRAIN;
		ev_add(o, &fl->drain_t, &fl->drain_act, &fl->drain_seq, K_DRAIN, f);
	} else if (!is_later_than(drain_time+o->now, o->flows.v[prod].done_act, o->flows.v[prod].done_t) &&
		   sgrow < fbw) {                                            /* :45-61 */
		fl->drain_type = FLOW_SPEED_THROTTLE;
		if (fequ(srng->len, drng_start)) {
			if (!fequ(sgrow, fbw)) {
				fl->drain_t = o->now;
				ev_add(o, &fl->drain_t, &fl->drain_act, &fl->drain_seq, K_DRAIN, f);
			}
		} else {
			fl->drain_t = o->now+drain_time;
			ev_add(o, &fl->drain_t, &fl->drain_act, &fl->drain_seq, K_DRAIN, f);
		}
	}
	int npos;                                                            /* :64-69, int npos :32 */
	if (drng->next >= 0) npos = (int)o->ranges.v[drng->next].start;
	else npos = (int)drng->total;
	double done_time = ((double)(uint64_t)((uint64_t)(int64_t)npos-(uint64_t)drng->start)-drng->len)/fbw; /* :70 */
	if (!is_later_than(o->now+done_time, fl->drain_act, fl->drain_t)) {  /* :73-78 */
		ev_remove(o, &fl->drain_act);
		fl->done_t = o->now+done_time;
		ev_add(o, &fl->done_t, &fl->done_act, &fl->done_seq, K_DONE, f);
	}
}
static void range_update_consumer_events(ns_oracle *o, int r){           /* range.h:65-69 */
	for (int f = o->ranges.v[r].cons_head; f >= 0; f = o->flows.v[f].cons_next)
		range_calc_and_requeue_events(o, f);
}

static int new_range(ns_oracle *o){
	int i;
	if (o->free_range >= 0) { i = o->free_range; o->free_range = o->ranges.v[i].next_free; }
	else { VPUSH(o->ranges); i = o->ranges.n-1; }
	memset(&o->ranges.v[i], 0, sizeof o->ranges.v[i]);
	o->ranges.v[i].prod = o->ranges.v[i].next = o->ranges.v[i].prev = o->ranges.v[i].cons_head = -1;
	return i;
}
/* range_get, range.h:30-41 with range_include_cmp :10-20 (key is an int) */
static int range_get(ns_oracle *o, int res, int start){
	int r = o->res.v[res].first;
	for (; r >= 0; r = o->ranges.v[r].next) {
		struct o_range *R = &o->ranges.v[r];
		int c;
		if (R->start <= (int64_t)start) c = ((double)R->start+R->len >= (double)start) ? 0 : -1;
		else c = 1;
		if (c >= 0) break;
	}
	if (r < 0) return -1;
	if (o->ranges.v[r].start > (int64_t)start) return -1;
	return r;
}
/* resource_new_range, range.h:71-82: sorted insert by start, burns one random() */
static int resource_new_range(ns_oracle *o, int res, int64_t start, double len){
	int i = new_range(o);
	struct o_range *R = &o->ranges.v[i];
	R->start = start; R->len = len; R->total = o->res.v[res].len; R->res = res;
	ns_rand_next(&o->rng); o->c[NS_C_DRAWS]++;
	int prev = -1, cur = o->res.v[res].first;
	while (cur >= 0 && (int)((uint64_t)o->ranges.v[cur].start-(uint64_t)(int64_t)(int)start) < 0) {  /* range_start_cmp :22-26 */
		prev = cur; cur = o->ranges.v[cur].next;
	}
	R->prev = prev; R->next = cur;
	if (prev >= 0) o->ranges.v[prev].next = i; else o->res.v[res].first = i;
	if (cur >= 0) o->ranges.v[cur].prev = i;
	return i;
}
static int store_get(ns_oracle *o, int node, uint32_t rid){              /* store.h:5-10 */
	for (int r = o->nodes.v[node].res_head; r >= 0; r = o->res.v[r].next_in_node)
		if (o->res.v[r].id == rid) return r;
	return -1;
}
static int resource_new(ns_oracle *o, int node, uint32_t rid, int64_t len){  /* data.h:241-249 + store_set */
	struct o_res *R = VPUSH(o->res);
	int i = o->res.n-1;
	R->id = rid; R->len = len; R->node = node; R->first = -1; R->nprod = R->ncons = 0;
	R->next_in_node = o->nodes.v[node].res_head;
	o->nodes.v[node].res_head = i;
	return i;
}

/* range_merge_with_next, range.c:83-124 */
static void range_merge_with_next(ns_oracle *o, int r){
	struct o_range *R = &o->ranges.v[r];
	int nr = R->next;
	if (nr < 0) return;
	struct o_range *N = &o->ranges.v[nr];
	R->len = (double)(uint64_t)((uint64_t)N->start-(uint64_t)R->start)+N->len;  /* :94 */
	R->prod = N->prod;                                                   /* :96-98 */
	if (R->prod >= 0) o->flows.v[R->prod].drng = r;
	while (N->cons_head >= 0) {                                          /* :100-107 */
		int f = N->cons_head;
		struct o_flow *fl = &o->flows.v[f];
		N->cons_head = fl->cons_next;
		if (fl->cons_next >= 0) o->flows.v[fl->cons_next].cons_prev = -1;
		fl->cons_prev = -1; fl->cons_next = R->cons_head;
		if (R->cons_head >= 0) o->flows.v[R->cons_head].cons_prev = f;
		R->cons_head = f;
		fl->srng = r;
	}
	R->next = N->next;                                                   /* :109-110 */
	if (N->next >= 0) o->ranges.v[N->next].prev = r;
	N->next_free = o->free_range; o->free_range = nr;
	range_update(o, r);                                                  /* :113 */
	for (int f = R->cons_head; f >= 0; f = o->flows.v[f].cons_next) {    /* :114-117 */
		range_update(o, o->flows.v[f].drng);
		range_calc_and_requeue_events(o, f);
	}
	if (R->prod >= 0) {                                                  /* :119-123 */
		range_update(o, o->flows.v[R->prod].srng);
		range_calc_and_requeue_events(o, R->prod);
	}
}

/* ---------------------------------------------------------------- flows */
static int flow_create(ns_oracle *o, int src, int dst, double ub, double delay){ /* flow.c:296-337 */
	int f;
	if (o->free_flow >= 0) { f = o->free_flow; o->free_flow = o->flows.v[f].next_free; }
	else { VPUSH(o->flows); f = o->flows.n-1; }
	struct o_flow *fl = &o->flows.v[f];
	memset(fl, 0, sizeof *fl);
	fl->ub = ub; fl->delay = delay; fl->peer[0] = src; fl->peer[1] = dst;
	fl->spd_head = -1; fl->srng = fl->drng = -1; fl->cons_next = fl->cons_prev = -1; fl->alive = 1;
	struct o_node *S = &o->nodes.v[src], *D = &o->nodes.v[dst];
	fl->cprev[0] = -1; fl->cnext[0] = S->conn_head[0];                   /* :311 list_add head */
	if (S->conn_head[0] >= 0) o->flows.v[S->conn_head[0]].cprev[0] = f;
	S->conn_head[0] = f;
	fl->cprev[1] = -1; fl->cnext[1] = D->conn_head[1];                   /* :312 */
	if (D->conn_head[1] >= 0) o->flows.v[D->conn_head[1]].cprev[1] = f;
	D->conn_head[1] = f;
	S->total[0] += ub;                                                   /* :314-315 */
	D->total[1] += ub;
	double spd0 = get_share(o, f, 0);                                    /* :317-318 */
	bwspread(o, f, spd0, 0, 0);
	fl = &o->flows.v[f];
	queue_speed_event(o, f, RCV, fl->speed[0]);                          /* :322 */
	fl->id = ns_rand_next(&o->rng); o->c[NS_C_DRAWS]++;                  /* :326-330 */
	write_record(o, R_CONN_CREATE, S->id, 4, fl->id, 0);                 /* :333-334 */
	write_record(o, R_CONN_DST, fl->id, 4, D->id, 0);
	return f;
}

static void flow_close(ns_oracle *o, int f){                             /* flow.c:241-292 */
	struct o_flow *fl = &o->flows.v[f];
	struct o_node *S = &o->nodes.v[fl->peer[0]], *D = &o->nodes.v[fl->peer[1]];
	S->total[0] -= fl->ub;                                               /* :246-247 */
	D->total[1] -= fl->ub;
	flow_range_update(o, f);                                             /* :248 */
	bwspread(o, f, fl->speed[0], 0, 1);                                  /* :249-250 */
	bwspread(o, f, fl->speed[1], 1, 1);
	for (int d = 0; d < 2; d++) {                                        /* :251-252 */
		struct o_node *n = &o->nodes.v[fl->peer[d]];
		if (fl->cprev[d] >= 0) o->flows.v[fl->cprev[d]].cnext[d] = fl->cnext[d]; else n->conn_head[d] = fl->cnext[d];
		if (fl->cnext[d] >= 0) o->flows.v[fl->cnext[d]].cprev[d] = fl->cprev[d];
	}
	while (fl->spd_head >= 0) {                                          /* :260-270 */
		int i = fl->spd_head;
		spd_unlink(o, i);
		ev_remove(o, &o->spd.v[i].active);
		spd_free(o, i);
	}
	if (fl->drng >= 0 && o->ranges.v[fl->drng].prod == f)                /* :273-274 */
		o->ranges.v[fl->drng].prod = -1;
	o->res.v[o->ranges.v[fl->srng].res].ncons--;                         /* :275-276 */
	o->res.v[o->ranges.v[fl->drng].res].nprod--;
	/* :279-281 auto_delete is only ever set by modules/resource.c:63; scripts never set it */
	if (fl->cons_prev >= 0) o->flows.v[fl->cons_prev].cons_next = fl->cons_next;   /* :282 */
	else o->ranges.v[fl->srng].cons_head = fl->cons_next;
	if (fl->cons_next >= 0) o->flows.v[fl->cons_next].cons_prev = fl->cons_prev;
	ev_remove(o, &fl->done_act);                                         /* :283-284 */
	ev_remove(o, &fl->drain_act);
	write_record(o, R_CONN_CLOSE, fl->id, 1, 0, 0);                      /* :289-290 */
	fl->alive = 0;
	if (fl->ord >= 0) o->ord2flow.v[fl->ord] = -1;
	fl->next_free = o->free_flow; o->free_flow = f;
	o->c[NS_C_CLOSES]++; o->c[NS_C_LIVE_FLOWS]--;
}

/* sim_establish_flow, sim.c:42-94 */
static int sim_establish_flow(ns_oracle *o, uint32_t rid, int64_t start, int src, int dst, double ub, double delay){
	int sr = store_get(o, src, rid);
	if (sr < 0) return -1;                                               /* :49-53 */
	int dr = store_get(o, dst, rid);
	if (dr < 0) dr = resource_new(o, dst, rid, o->res.v[sr].len);        /* :54-61 */
	if (range_get(o, dr, (int)start) >= 0) return -1;                    /* :62-67 */
	int rng = range_get(o, sr, (int)start);                              /* :69 */
	int f = flow_create(o, src, dst, ub, delay);                         /* :70 */
	int dg = resource_new_range(o, dr, start, 0);                        /* :74 */
	struct o_flow *fl = &o->flows.v[f];
	fl->drng = dg;
	o->ranges.v[dg].prod = f;                                            /* :75-76 */
	o->ranges.v[dg].last = o->now;
	fl->srng = rng;                                                      /* :77 */
	if (rng < 0) { if (!o->status) o->status = NS_O_NULL_RANGE; return f; }
	range_calc_and_requeue_events(o, f);                                 /* :79 */
	fl = &o->flows.v[f];
	fl->cons_prev = -1; fl->cons_next = o->ranges.v[rng].cons_head;      /* :80 list_add head */
	if (fl->cons_next >= 0) o->flows.v[fl->cons_next].cons_prev = f;
	o->ranges.v[rng].cons_head = f;
	o->res.v[sr].ncons++;                                                /* :82-83 */
	o->res.v[dr].nprod++;
	int ph = o->ranges.v[dg].prev;                                       /* :86-91 */
	if (ph >= 0)
		range_calc_and_requeue_events(o, o->ranges.v[ph].prod);
	return f;
}

/* ---------------------------------------------------------------- handlers */
static void handle_speed_change(ns_oracle *o, int si){                   /* flow.c:341-384 */
	struct o_spd *se = &o->spd.v[si];
	int f = se->flow, type = se->dir;
	double speed = se->speed;
	struct o_flow *fl = &o->flows.v[f];
	if (type == RCV)
		flow_range_update(o, f);                                         /* :355-357 */
	double delta = speed-fl->speed[type];                                /* :359 */
	bwspread(o, f, delta, type, 0);                                      /* :361 */
	spd_unlink(o, si);                                                   /* :365 */
	if (type == RCV) {                                                   /* :367-374 */
		range_calc_and_requeue_events(o, f);
		range_update_consumer_events(o, o->flows.v[f].drng);
	} else if (delta > 0)
		queue_speed_event(o, f, RCV, o->flows.v[f].speed[0]);
	write_record(o, R_SPD|type, o->flows.v[f].id, -1, 0, speed);         /* :383 */
	trace_row(o, SPEED_CHANGE|(type<<8), o->flows.v[f].ord, speed, o->flows.v[f].speed[type]);
	spd_free(o, si);                                                     /* speed_change_free :386-389 */
}

static void run_group(ns_oracle *o, int g);

static void handle_flow_done(ns_oracle *o, int f){
	/* flow_done_handler, flow.c:391-401 */
	flow_range_update(o, f);
	int drng = o->flows.v[f].drng;
	int next = o->ranges.v[drng].next;
	if (next >= 0) {
		flow_range_update(o, o->ranges.v[next].prod);
		range_merge_with_next(o, drng);
	}
	/* HNDR_USER slot: oracle/ref_scenario.c on_done() */
	int ord = o->flows.v[f].ord;
	trace_row(o, FLOW_DONE, ord, o->ranges.v[o->flows.v[f].drng].len, o->flows.v[f].speed[1]);
	if (o->H.n_ondone) {
		int lo = 0, hi = o->H.n_ondone-1, g = -1;
		while (lo <= hi) {
			int m = (lo+hi)/2;
			if (o->ondone[m].flow_ord == ord) { g = o->ondone[m].group; break; }
			if (o->ondone[m].flow_ord < ord) lo = m+1; else hi = m-1;
		}
		if (g >= 0) run_group(o, g);
	}
	/* flow_done_cleaner, flow.c:403-406 */
	flow_close(o, f);
}

static void handle_throttle(ns_oracle *o, int f){                        /* flow.c:408-423 */
	struct o_flow *fl = &o->flows.v[f];
	int prod = o->ranges.v[fl->srng].prod;
	if (prod < 0) { if (!o->status) o->status = NS_O_NO_PRODUCER; return; }
	double delta = o->flows.v[prod].speed[1]-fl->speed[0];               /* :412 */
	if (!(delta >= eps)) {
		flow_range_update(o, f);                                         /* :416 */
		bwspread(o, f, delta, SND, 0);                                   /* :418 */
		delta = o->flows.v[prod].speed[1]-o->flows.v[f].speed[1];        /* :419 */
		bwspread(o, f, delta, RCV, 0);
		range_calc_and_requeue_events(o, f);                             /* :421-422 */
		range_update_consumer_events(o, o->flows.v[f].drng);
	}
	trace_row(o, FLOW_SPEED_THROTTLE, o->flows.v[f].ord, o->flows.v[f].speed[0], o->flows.v[f].speed[1]);
}

static void run_group(ns_oracle *o, int g){                              /* oracle/ref_scenario.c run_group() */
	for (int i = o->group_off[g]; i < o->group_off[g+1] && !o->status; i++) {
		const struct ns_sc_op *op = &o->ops[i];
		switch (op->kind) {
		case NS_OP_CREATE: {
			int f = sim_establish_flow(o, o->rids.v[op->c], op->start, op->a, op->b, op->x, op->y);
			int ord = o->ord2flow.n;
			*VPUSH(o->ord2flow) = f;
			if (f >= 0) {
				o->flows.v[f].ord = ord;
				o->c[NS_C_CREATES]++;
				if (++o->c[NS_C_LIVE_FLOWS] > o->c[NS_C_PEAK_FLOWS]) o->c[NS_C_PEAK_FLOWS] = o->c[NS_C_LIVE_FLOWS];
			}
			break; }
		case NS_OP_SCHED: {
			int i2;
			if (o->free_user >= 0) { i2 = o->free_user; o->free_user = o->users.v[i2].next_free; }
			else { VPUSH(o->users); i2 = o->users.n-1; }
			struct o_user *u = &o->users.v[i2];
			u->time = op->c ? o->now+op->x : op->x; u->group = op->a; u->active = 0;
			ev_add(o, &u->time, &u->active, &u->seq, K_USER, i2);
			break; }
		case NS_OP_CLOSE:
			if (op->a < o->ord2flow.n && o->ord2flow.v[op->a] >= 0)
				flow_close(o, o->ord2flow.v[op->a]);
			break;
		case NS_OP_END:
			o->exit_flag = 1;                                            /* sim.h:42-45 */
			break;
		}
	}
}

/* ---------------------------------------------------------------- public API */
ns_oracle *ns_oracle_create(const void *scenario, size_t len, int want_trace, int want_records){
	ns_oracle *o = calloc(1, sizeof *o);
	o->want_tr = want_trace; o->want_rec = want_records;
	o->free_flow = o->free_range = o->free_spd = o->free_user = -1;
	if (len < sizeof o->H) { o->status = NS_O_BAD_SCENARIO; return o; }
	o->blob = malloc(len);
	memcpy(o->blob, scenario, len);
	memcpy(&o->H, o->blob, sizeof o->H);
	struct ns_sc_header *H = &o->H;
	if (H->magic != NS_SC_MAGIC || H->version != NS_SC_VERSION) { o->status = NS_O_BAD_SCENARIO; return o; }
	if (H->flags) { o->status = NS_O_BAD_SCENARIO; return o; }   /* scripts with client behaviour (NS_SC_F_CLIENTS) are checked against the compiled reference itself */
	size_t need = sizeof *H+(size_t)H->n_nodes*16+(size_t)H->n_resources*sizeof(struct ns_sc_resource)+
		(size_t)H->n_copies*sizeof(struct ns_sc_copy)+(size_t)(H->n_groups+1)*4+
		(size_t)H->n_ops*sizeof(struct ns_sc_op)+(size_t)H->n_ondone*sizeof(struct ns_sc_ondone);
	if (len < need || H->n_groups < 1) { o->status = NS_O_BAD_SCENARIO; return o; }
	const char *p = (const char *)o->blob+sizeof *H;
	const double *max_up = (const double *)p; p += (size_t)H->n_nodes*8;
	const double *max_down = (const double *)p; p += (size_t)H->n_nodes*8;
	const struct ns_sc_resource *rs = (const void *)p; p += (size_t)H->n_resources*sizeof *rs;
	const struct ns_sc_copy *cp = (const void *)p; p += (size_t)H->n_copies*sizeof *cp;
	o->group_off = (const void *)p; p += (size_t)(H->n_groups+1)*4;
	o->ops = (const void *)p; p += (size_t)H->n_ops*sizeof(struct ns_sc_op);
	o->ondone = (const void *)p;

	ns_rand_seed(&o->rng, H->seed);
	/* sim_create_node, sim.c:113-125: id = random() until unused; R_NODE_CREATE record */
	uint32_t hcap = 16; while (hcap < (uint32_t)H->n_nodes*2u+16u) hcap <<= 1;
	uint32_t *hset = calloc(hcap, sizeof *hset); uint8_t *hocc = calloc(hcap, 1);
	for (int i = 0; i < H->n_nodes; i++) {
		struct o_node *n = VPUSH(o->nodes);
		memset(n, 0, sizeof *n);
		n->conn_head[0] = n->conn_head[1] = n->res_head = -1;
		while (1) {
			uint32_t id = ns_rand_next(&o->rng); o->c[NS_C_DRAWS]++;
			uint32_t h = (id*2654435761u)&(hcap-1);
			int dup = 0;
			while (hocc[h]) { if (hset[h] == id) { dup = 1; break; } h = (h+1)&(hcap-1); }
			if (dup) continue;
			hocc[h] = 1; hset[h] = id; n->id = id;
			break;
		}
		write_record(o, R_NODE_CREATE, n->id, 1, 0, 0);
		n->max[0] = max_up[i]; n->max[1] = max_down[i];
	}
	free(hset); free(hocc);
	/* sim_node_new_resource, sim.c:129-140 */
	for (int i = 0; i < H->n_resources; i++) {
		uint32_t rid = ns_rand_next(&o->rng); o->c[NS_C_DRAWS]++;
		while (store_get(o, rs[i].owner, rid) >= 0) { rid = ns_rand_next(&o->rng); o->c[NS_C_DRAWS]++; }
		int r = resource_new(o, rs[i].owner, rid, rs[i].len);
		resource_new_range(o, r, 0, (double)rs[i].len);
		*VPUSH(o->rids) = rid;
	}
	/* sim_node_add_resource, sim.c:141-154 */
	for (int i = 0; i < H->n_copies; i++) {
		uint32_t rid = o->rids.v[cp[i].res];
		if (store_get(o, cp[i].node, rid) >= 0) continue;
		int r = resource_new(o, cp[i].node, rid, rs[cp[i].res].len);
		resource_new_range(o, r, 0, (double)rs[cp[i].res].len);
	}
	run_group(o, 0);
	return o;
}

int ns_oracle_run(ns_oracle *o, uint64_t max_events){
	while (o->heap.n && !o->exit_flag && !o->status) {                   /* main.c:50-67 */
		if (max_events && o->c[NS_C_EVENTS] >= max_events) break;
		struct o_hent e = heap_pop(o);
		uint8_t *act; uint64_t seq;
		switch (e.kind) {
		case K_DONE: act = &o->flows.v[e.idx].done_act; seq = o->flows.v[e.idx].done_seq; break;
		case K_DRAIN: act = &o->flows.v[e.idx].drain_act; seq = o->flows.v[e.idx].drain_seq; break;
		case K_SPD: act = &o->spd.v[e.idx].active; seq = o->spd.v[e.idx].seq; break;
		default: act = &o->users.v[e.idx].active; seq = o->users.v[e.idx].seq; break;
		}
		if (!*act || seq != e.seq) continue;                             /* cancelled (event_remove) */
		*act = 0; o->c[NS_C_PENDING]--;                                  /* event_pop, event.h:20-25 */
		o->now = e.time;                                                 /* main.c:60 */
		o->c[NS_C_EVENTS]++;
		switch (e.kind) {
		case K_SPD: o->c[NS_C_SPEED_CHANGE]++; handle_speed_change(o, e.idx); break;
		case K_DONE: o->c[NS_C_FLOW_DONE]++; handle_flow_done(o, e.idx); break;
		case K_DRAIN:
			if (o->flows.v[e.idx].drain_type == FLOW_SPEED_THROTTLE) { o->c[NS_C_THROTTLE]++; handle_throttle(o, e.idx); }
			else { o->c[NS_C_DRAIN]++; trace_row(o, FLOW_DRAIN, o->flows.v[e.idx].ord, 0, 0); }
			break;
		default: {
			int g = o->users.v[e.idx].group;
			o->c[NS_C_USER]++;
			o->users.v[e.idx].next_free = o->free_user; o->free_user = e.idx;
			trace_row(o, USER, g, 0, 0);
			run_group(o, g);
			break; }
		}
	}
	o->c[NS_C_SEQ] = o->seq;
	return o->status;
}

int ns_oracle_status(const ns_oracle *o){ return o->status; }
double ns_oracle_now(const ns_oracle *o){ return o->now; }
const uint8_t *ns_oracle_records(const ns_oracle *o, size_t *len){ *len = o->rec_n; return o->rec; }
const struct ns_trace_row *ns_oracle_trace(const ns_oracle *o, size_t *rows){ *rows = o->tr_n; return o->tr; }
void ns_oracle_counters(const ns_oracle *o, uint64_t out[NS_C_COUNT]){
	memcpy(out, o->c, sizeof o->c);
	out[NS_C_SEQ] = o->seq;
}
size_t ns_oracle_flow_state(const ns_oracle *o, double *out4, size_t max_flows){
	size_t n = (size_t)o->ord2flow.n < max_flows ? (size_t)o->ord2flow.n : max_flows;
	for (size_t i = 0; i < n; i++) {
		int f = o->ord2flow.v[i];
		if (f < 0) { out4[4*i] = 0; out4[4*i+1] = out4[4*i+2] = out4[4*i+3] = 0; continue; }
		out4[4*i] = 1; out4[4*i+1] = o->flows.v[f].speed[0]; out4[4*i+2] = o->flows.v[f].speed[1];
		out4[4*i+3] = o->flows.v[f].drng >= 0 ? o->ranges.v[o->flows.v[f].drng].len : 0;
	}
	return n;
}
size_t ns_oracle_node_state(const ns_oracle *o, double *out4, size_t max_nodes){
	size_t n = (size_t)o->nodes.n < max_nodes ? (size_t)o->nodes.n : max_nodes;
	for (size_t i = 0; i < n; i++) {
		out4[4*i] = o->nodes.v[i].usage[0]; out4[4*i+1] = o->nodes.v[i].usage[1];
		out4[4*i+2] = o->nodes.v[i].total[0]; out4[4*i+3] = o->nodes.v[i].total[1];
	}
	return n;
}
void ns_oracle_destroy(ns_oracle *o){
	if (!o) return;
	free(o->blob); free(o->nodes.v); free(o->res.v); free(o->ranges.v); free(o->flows.v); free(o->spd.v);
	free(o->users.v); free(o->heap.v); free(o->ord2flow.v); free(o->rids.v); free(o->rec); free(o->tr);
	free(o);
}

/* ---------------------------------------------------------------- pickers */
static void qsel(int32_t *val, int32_t *node, size_t nmemb, size_t k){   /* modules/qselect.h:5-27 */
	size_t i, st;
	int32_t tv, tn;
	for (st = i = 0; i < nmemb-1; i++) {
		if ((int32_t)((uint32_t)val[i]-(uint32_t)val[nmemb-1]) > 0) continue;   /* cmp(v+i, v+nmemb-1) > 0 */
		tv = val[i]; val[i] = val[st]; val[st] = tv;
		tn = node[i]; node[i] = node[st]; node[st] = tn;
		st++;
	}
	tv = val[st]; val[st] = val[nmemb-1]; val[nmemb-1] = tv;
	tn = node[st]; node[st] = node[nmemb-1]; node[nmemb-1] = tn;
	if (k-1 == st) return;
	if (st > k-1) qsel(val, node, st, k);
	else qsel(val+st, node+st, nmemb-st, k-st);
}
void ns_oracle_qselect(int32_t *val, int32_t *node, size_t nmemb, size_t k){
	if (nmemb == 0 || k == 0 || k > nmemb) return;
	qsel(val, node, nmemb, k);
}
int ns_oracle_pick_opt1(int32_t *val, int32_t *node, int c, int k){      /* client_behaviour.h:77-87 */
	if (c <= k) return c;
	ns_oracle_qselect(val, node, (size_t)c, (size_t)k);
	return k;
}
int ns_oracle_pick_opt2(const int32_t *val, const uint8_t *usable, int c){ /* client_behaviour.h:94-117 */
	int32_t min = INT32_MAX; int res = -1;
	for (int i = 0; i < c; i++) {
		if (!usable[i]) continue;
		if (val[i] < min) { res = i; min = val[i]; }
	}
	return res;
}

/* distance_metric / share_metric, modules/p2p_common.h:171-198 (distance_based_bw :181-189) */
int ns_oracle_metric(int metric, int tz_node, int tz_client, double total_up, double max_up){
	int d = tz_node-tz_client;
	if (d < 0) d = -d;
	if (metric == 0) return d;
	double bw = 4000-100*d;
	double total = total_up+bw;
	if (total > max_up) return -bw*max_up/total;
	return -bw;
}
