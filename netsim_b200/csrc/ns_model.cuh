// ns_model.cuh — the flow-level bandwidth model as warp-cooperative device functions.
//
// Execution model: every function here is called by ALL 32 lanes of one warp with identical
// (warp-uniform) arguments.  Scalar state is loaded redundantly (one broadcast transaction), control
// flow is uniform, stores are issued by lane 0, and the O(degree) loops of a full spread
// (flow.c:127-200) are strided over the lanes.  Floating-point results are bit-identical to the
// sequential reference: every sum that the reference accumulates in list order (e, new_use) is
// accumulated in the same order from lane-shuffled terms; nvcc is run with -fmad=false.
//
// Each function cites the reference lines it re-implements (paths relative to yshui/netsim).
#pragma once
#include "ns_sim.cuh"

namespace {   // internal linkage: every engine translation unit carries its own copy of the device code

struct Ctx {
	const Sim *S; SimState *st; int lane;
	double now;
	u64 head, tail;        // ring window, cached in registers (flushed by ctx_flush)
	u64 burns, rec_len, tr_len;
	int par;               // 1 inside the grid-parallel path: emissions are staged
	int scalar;            // 1 when ONE thread runs the handler alone (O(1) solves, ns_par.cuh phase E1)
	int pre_slot;          // >= 0 while a CREATE op runs in the parallel init: its flow slot / ordinal (ns_par.cuh)
	unsigned id_pos;       // ... and where in the op's draws the flow id falls
	const ParStage *ps; unsigned ev, lseq, rec_n, tr_n; long long ovf_base;
	u64 lc[LC_N];
};

#define L0 if (c.scalar || c.lane == 0)
#define WSYNC() do { if (!c.scalar) __syncwarp(); } while (0)

__device__ __forceinline__ void ctx_load(Ctx &c, const Sim *S, int lane){
	c.S = S; c.st = S->st; c.lane = lane; c.par = 0; c.scalar = 0; c.ps = nullptr; c.pre_slot = -1; c.id_pos = 0;
	c.now = c.st->now; c.head = c.st->head; c.tail = c.st->tail; c.burns = c.st->burns;
	c.rec_len = c.st->rec_len; c.tr_len = c.st->tr_len;
}
__device__ __forceinline__ void ctx_flush(Ctx &c){
	L0 {
		c.st->now = c.now; c.st->head = c.head; c.st->tail = c.tail; c.st->burns = c.burns;
		c.st->rec_len = c.rec_len; c.st->tr_len = c.tr_len;
	}
	WSYNC();
}
__device__ __forceinline__ void set_status(Ctx &c, int s){
	L0 { atomicCAS(&c.st->status, 0, s); }
}
__device__ __forceinline__ int lc_slot(int idx){
	return idx == NS_CNT_EVENTS ? LC_EVENTS : idx == NS_CNT_SPEED_CHANGE ? LC_SPEED_CHANGE : idx == NS_CNT_SOLVES ? LC_SOLVES :
	       idx == NS_CNT_FULL_SPREADS ? LC_FULL_SPREADS : idx == NS_CNT_RETIMES ? LC_RETIMES : LC_FLOW_VISITS;
}
__device__ __forceinline__ void cnt_add(Ctx &c, int idx, u64 n){
	if (c.par) { c.lc[lc_slot(idx)] += n; return; }
	L0 { c.st->cnt[idx] += n; }
}
__device__ __forceinline__ void tmin_lower(double *p, double t){
	// times are >= 0, so the IEEE bit pattern orders like the value
	atomicMin((u64 *)p, (u64)__double_as_longlong(t));
}

// ---------------------------------------------------------------- records (record.c:27-117)
__device__ __forceinline__ int rec_encode(uint8_t *b, double now, uint8_t minor, uint32_t id, int bytes, uint32_t ival, double dval){
	int n = 0;
	b[n++] = 0; b[n++] = minor;
	b[n++] = id >> 24; b[n++] = id >> 16; b[n++] = id >> 8; b[n++] = id;
	uint32_t t = (uint32_t)(int)now;                                   // record.c:66
	b[n++] = t >> 24; b[n++] = t >> 16; b[n++] = t >> 8; b[n++] = t;
	uint16_t ut = (uint16_t)(int)((now-(double)((int)now))*1000);      // record.c:69
	b[n++] = ut >> 8; b[n++] = ut;
	b[n++] = (uint8_t)(int8_t)bytes;
	if (bytes == 1) b[n++] = (uint8_t)ival;
	else if (bytes == 4) { b[n++] = ival >> 24; b[n++] = ival >> 16; b[n++] = ival >> 8; b[n++] = ival; }
	else {                                                             // record.c:98-106
		double tmp = dval;
		uint32_t ip = (uint32_t)(int)tmp;
		b[n++] = ip >> 24; b[n++] = ip >> 16; b[n++] = ip >> 8; b[n++] = ip;
		tmp = tmp-(double)((int)tmp);
		uint16_t fp = (uint16_t)(int)(tmp*1000);
		b[n++] = fp >> 8; b[n++] = fp;
	}
	return n;
}
__device__ __forceinline__ void write_record(Ctx &c, uint8_t minor, uint32_t id, int bytes, uint32_t ival, double dval){
	const Sim &S = *c.S;
	if (!S.want_rec) return;
	int n = 13+(bytes == 1 ? 1 : bytes == 4 ? 4 : 6);
	if (c.par) {
		if (c.rec_n+n > NS_STAGE_REC) { set_status(c, NS_E_TABLE_FULL); return; }
		L0 rec_encode(c.ps->st_rec+(u64)c.ev*NS_STAGE_REC+c.rec_n, c.now, minor, id, bytes, ival, dval);
		c.rec_n += n;
		return;
	}
	if (c.rec_len+n > S.rec_cap) { set_status(c, NS_E_RECORDS_LOST); return; }   // a record was DROPPED: hard error, unlike the soft stop before a batch
	L0 {
		uint8_t b[19];
		rec_encode(b, c.now, minor, id, bytes, ival, dval);
		uint8_t *dst = S.rec+c.rec_len;
		for (int i = 0; i < n; i++) dst[i] = b[i];
	}
	c.rec_len += n;
}
__device__ __forceinline__ void write_usage(Ctx &c, int dir, int n, double usage){  // record_wrapper.h:6-9
	write_record(c, R_USAGE|dir, c.S->n_id[n], -1, 0, usage);
}
__device__ __forceinline__ void trace_row(Ctx &c, uint32_t type, uint32_t ord, double v0, double v1){
	const Sim &S = *c.S;
	if (!S.want_tr) return;
	if (c.par) {
		L0 { ns_trace_row32 r; r.type = type; r.ord = ord; r.time = c.now; r.v0 = v0; r.v1 = v1; c.ps->st_tr[c.ev] = r; }
		c.tr_n = 1;
		return;
	}
	if (c.tr_len >= S.tr_cap) return;
	L0 { ns_trace_row32 r; r.type = type; r.ord = ord; r.time = c.now; r.v0 = v0; r.v1 = v1; S.tr[c.tr_len] = r; }
	c.tr_len++;
}

// ---------------------------------------------------------------- event store
// event_add's time checks, event.h:58-70.  Returns false where the reference would abort().
__device__ __forceinline__ bool ev_time_ok(Ctx &c, double &t){
	if (t < c.now) {
		if (t < c.now-NS_EPS) { set_status(c, NS_E_BACK_IN_TIME); return false; }
		t = c.now;
	}
	if (isnan(t)) { set_status(c, NS_E_NAN_TIME); return false; }
	return true;
}
// take n consecutive sequence numbers (each successful event_add: one skip-list insert, one
// random() burned, skiplist.h:34-40,83)
__device__ __forceinline__ u64 stage_index(const Ctx &c, u64 j){
	return j < NS_STAGE_INLINE ? (u64)c.ev*NS_STAGE_INLINE+j : c.ps->inline_n+(u64)c.ovf_base+(j-NS_STAGE_INLINE);
}
// before a handler may stage more than NS_STAGE_INLINE inserts (a full spread) it takes a block
// of the overflow arena sized for its worst case
__device__ __forceinline__ void stage_ensure(Ctx &c, u64 need){
	if (!c.par || c.ovf_base >= 0) return;
	long long b = 0;
	L0 {
		b = (long long)atomicAdd((unsigned long long *)c.ps->ovf_top, (unsigned long long)need);
		if ((u64)b+need > c.ps->ovf_cap) { atomicCAS(&c.st->status, 0, NS_E_TABLE_FULL); b = 0; }
	}
	c.ovf_base = __shfl_sync(FULLMASK, b, 0);
}
__device__ __forceinline__ u64 seq_reserve(Ctx &c, u64 n){
	if (c.par) { u64 s0 = c.lseq; c.lseq += (unsigned)n; return s0; }
	u64 s = c.tail;
	if (c.tail+n-c.head > c.S->ring_cap) { set_status(c, NS_E_RING_FULL); return s; }
	c.tail += n; c.burns += n;
	return s;
}
__device__ __forceinline__ void ring_put(Ctx &c, u64 seq, double t, double speed, int f, int dir){
	const Sim &S = *c.S;
	if (c.par) {
		u64 i = stage_index(c, seq);
		c.ps->st_time[i] = t; c.ps->st_speed[i] = speed; c.ps->st_fd[i] = f*2+dir;
		return;
	}
	u64 slot = seq & (S.ring_cap-1);
	S.q_time[slot] = t; S.q_speed[slot] = speed; S.q_fd[slot] = f*2+dir; S.q_gen[slot] = S.f_gen[f];
	tmin_lower(&S.q_tmin[slot >> NS_TILE_SHIFT], t);
}
__device__ __forceinline__ void l_put(Ctx &c, int slot, u64 seq, double t){
	const Sim &S = *c.S;
	if (c.par) {
		// provisional sequence number (event, local index); the commit pass rewrites it
		S.l_time[slot] = t; S.l_seq[slot] = NS_SEQ_PROV | ((u64)c.ev << 24) | seq;
		tmin_lower(&S.l_tmin[slot >> NS_TILE_SHIFT], t);
		c.ps->st_fd[stage_index(c, seq)] = -(slot+1);
		return;
	}
	S.l_time[slot] = t; S.l_seq[slot] = seq;
	tmin_lower(&S.l_tmin[slot >> NS_TILE_SHIFT], t);
}
__device__ __forceinline__ void stage_pure_hole(Ctx &c, u64 seq){
	if (c.par) c.ps->st_fd[stage_index(c, seq)] = NS_HOLE_PURE;
}
// is_later_than, event.h:82-88 (inactive events are stored as time == +inf)
__device__ __forceinline__ bool is_later_than(double time, double etime){
	if (etime == NS_INF) return false;
	return time > etime+NS_EPS;
}

// queue_speed_event, flow.c:16-29
__device__ __forceinline__ void queue_speed_event(Ctx &c, int f, int dir, double speed){
	double t = c.now+c.S->f_delay[f];
	if (!ev_time_ok(c, t)) return;
	u64 seq = seq_reserve(c, 1);
	if (c.st->status) return;
	L0 ring_put(c, seq, t, speed, f, dir);
}

// get_share, flow.h:37-41
__device__ __forceinline__ double get_share(const Sim &S, int f, int dir){
	int n = S.f_peer[dir][f];
	double total = S.n_total[dir][n], mx = S.n_max[dir][n];
	double ub = S.f_ub[f];
	return total > mx ? ub*mx/total : ub;
}

// ---------------------------------------------------------------- ranges
// _range_update, range.h:91-118 (Kahan-compensated len += speed*dt); single-thread semantics,
// callers guard the stores.
__device__ __forceinline__ void range_update_1(const Sim &S, int r, double now){
	double last = S.r_last[r];
	if (now == last) return;
	int p = S.r_prod[r];
	if (p < 0) { S.r_last[r] = now; return; }
	double len = S.r_len[r], lenc = S.r_lenc[r];
	double delta = S.f_speed[1][p]*(now-last), t;
	delta = delta-lenc;
	t = delta+len;
	S.r_lenc[r] = (t-len)-delta;
	S.r_len[r] = t;
	S.r_last[r] = now;
}
// flow_range_update, range.h:120-125
__device__ __forceinline__ void flow_range_update(Ctx &c, int f){
	if (f < 0) return;
	const Sim &S = *c.S;
	int sr = S.f_srng[f], dr = S.f_drng[f];
	if (sr < 0 || dr < 0) { set_status(c, NS_E_NULL_RANGE); return; }
	L0 { range_update_1(S, sr, c.now); range_update_1(S, dr, c.now); }
	WSYNC();
}

// range_calc_and_requeue_events, range.c:16-79, split into a per-thread "plan" (all arithmetic and
// range updates, no sequence numbers) and a "commit" so that the victims of one full spread can be
// re-timed by different lanes while their event_add order stays the list order.
struct Retime { double drain_t, done_t; int add_drain, add_done; };
#define SP_U1 8     // chunks of 32 flows in flight in pass 1 of a full spread
#define SP_U2 4     // ... in pass 2

// the arithmetic of range.c:28-78 on values (shared by the table-driven and the register-resident paths)
__device__ __forceinline__ int retime_math(double now, double fbw, bool has_prod, double sgrow, double prod_done_t,
		double slen, double dlen, long long dstart, long long sstart, int npos, Retime &p, int &dtype){
	double drng_start = (double)dstart+dlen-(double)sstart;           // :36
	double drain_active = NS_INF;                                     // time of f->drain if active
	if (!has_prod) {                                                  // :40-44
		double t = now+(slen-drng_start)/fbw;
		dtype = EV_FLOW_DRAIN;
		bool bad = false;
		if (t < now) { if (t < now-NS_EPS) bad = true; t = now; }
		if (isnan(t)) return NS_E_NAN_TIME;
		if (bad) return NS_E_BACK_IN_TIME;
		p.add_drain = 1; p.drain_t = t; drain_active = t;
	} else {
		double drain_time = (slen-drng_start)/(fbw-sgrow);            // :38 (only read on this branch)
		if (!is_later_than(drain_time+now, prod_done_t) && sgrow < fbw) {   // :45-61
			dtype = EV_FLOW_SPEED_THROTTLE;
			double d0 = slen-drng_start;
			if (d0 > -NS_EPS && d0 < NS_EPS) {
				double d1 = sgrow-fbw;
				if (!(d1 > -NS_EPS && d1 < NS_EPS)) { p.add_drain = 1; p.drain_t = now; drain_active = now; }
			} else {
				double t = now+drain_time;
				bool bad = false;
				if (t < now) { if (t < now-NS_EPS) bad = true; t = now; }
				if (isnan(t)) return NS_E_NAN_TIME;
				if (bad) return NS_E_BACK_IN_TIME;
				p.add_drain = 1; p.drain_t = t; drain_active = t;
			}
		}
	}
	bool bad = false;
	double done_time = ((double)(u64)((u64)(long long)npos-(u64)dstart)-dlen)/fbw;  // :70
	if (!is_later_than(now+done_time, drain_active)) {                // :73-78
		double t = now+done_time;
		if (t < now) { if (t < now-NS_EPS) bad = true; t = now; }
		if (isnan(t)) bad = true;
		if (bad) return isnan(t) ? NS_E_NAN_TIME : NS_E_BACK_IN_TIME;
		p.add_done = 1; p.done_t = t;
	}
	return 0;
}

__device__ __forceinline__ void retime_plan_1(const Sim &S, SimState *st, int f, double now, Retime &p){
	p.add_drain = p.add_done = 0; p.drain_t = p.done_t = NS_INF;
	S.l_time[2*f] = NS_INF;                                           // range.c:19-20 event_remove
	S.l_time[2*f+1] = NS_INF;
	int sr = S.f_srng[f], dr = S.f_drng[f];
	if (sr < 0 || dr < 0) { if (st->status == 0) st->status = NS_E_NULL_RANGE; return; }
	range_update_1(S, sr, now);                                       // :21
	range_update_1(S, dr, now);
	double fbw = S.f_speed[1][f];
	if (fbw < NS_EPS) return;                                         // :26-27
	int prod = S.r_prod[sr];
	double sgrow = prod >= 0 ? S.f_speed[1][prod] : 0;                // :33
	int nx = S.r_next[dr];                                            // :64-69 (int npos)
	int npos = nx >= 0 ? (int)S.r_start[nx] : (int)S.r_total[dr];
	int dtype = -1;
	int err = retime_math(now, fbw, prod >= 0, sgrow, prod >= 0 ? S.l_time[2*prod] : NS_INF, S.r_len[sr], S.r_len[dr],
			      S.r_start[dr], S.r_start[sr], npos, p, dtype);
	if (dtype >= 0) S.f_drain_type[f] = (uint8_t)dtype;
	if (err) { if (st->status == 0) st->status = err; }
}
// sequence numbers: drain first, then done; a drain superseded by done keeps its burned number.
__device__ __forceinline__ void retime_commit_1(Ctx &c, int f, const Retime &p, u64 seq){
	if (p.add_done) { if (p.add_drain) stage_pure_hole(c, seq); l_put(c, 2*f, seq+p.add_drain, p.done_t); }
	else if (p.add_drain) l_put(c, 2*f+1, seq, p.drain_t);
}
__device__ __forceinline__ void range_calc_and_requeue_events(Ctx &c, int f){
	if (f < 0) return;
	cnt_add(c, NS_CNT_RETIMES, 1);
	Retime p;
	p.add_drain = p.add_done = 0;
	L0 retime_plan_1(*c.S, c.st, f, c.now, p);
	if (!c.scalar) {
		p.add_drain = __shfl_sync(FULLMASK, p.add_drain, 0);
		p.add_done = __shfl_sync(FULLMASK, p.add_done, 0);
	}
	int n = p.add_drain+p.add_done;
	if (n) {
		u64 seq = seq_reserve(c, n);
		L0 { if (c.st->status == 0 || c.st->status == NS_E_RECORDS_LOST) retime_commit_1(c, f, p, seq); }
	}
	WSYNC();
}
// range_update_consumer_events, range.h:65-69
__device__ __forceinline__ void range_update_consumer_events(Ctx &c, int r){
	for (int f = c.S->r_cons[r]; f >= 0; f = c.S->f_cnext[f])
		range_calc_and_requeue_events(c, f);
}

// ---------------------------------------------------------------- bwspread, flow.c:35-205
__device__ void bwspread(Ctx &c, int f, double amount, int dir, int close){
	const Sim &S = *c.S;
	const int lane = c.lane;
	cnt_add(c, NS_CNT_SOLVES, 1);
	double fs = S.f_speed[dir][f];
	if (close) amount = -fs;                                          // :51-54
	const int n = S.f_peer[dir][f];
	const double total = S.n_total[dir][n], mx = S.n_max[dir][n];     // :59
	const double used = S.n_usage[dir][n];
	const double ub = S.f_ub[f];
	const double share = total > mx ? ub*mx/total : ub;               // :61
	if (amount > NS_EPS && fs+amount > share) {                       // :64
		if (fs > share) { queue_speed_event(c, f, !dir, fs); return; } // :66-74
		queue_speed_event(c, f, !dir, share);                         // :75
		amount = share-fs;
	}
	if (amount > -NS_EPS && amount < NS_EPS) return;                  // :80-83
	fs += amount;                                                     // :85
	L0 S.f_speed[dir][f] = fs;
	if (amount < NS_EPS && amount > -64) {                            // :86-91
		double u = used+amount;
		L0 S.n_usage[dir][n] = u;
		write_usage(c, dir, n, u);
		WSYNC();
		return;
	}
	if (total < mx+NS_EPS) {                                          // :94-109
		if (!close || total+ub < mx+NS_EPS) {
			double u = used+amount;
			L0 S.n_usage[dir][n] = u;
			write_usage(c, dir, n, u);
			WSYNC();
			return;
		}
	}
	double spread_amount = amount;                                    // :112
	if (amount > -NS_EPS) {
		if (used+amount < mx+NS_EPS) {                                // :114-119
			double u = used+amount;
			L0 S.n_usage[dir][n] = u;
			write_usage(c, dir, n, u);
			WSYNC();
			return;
		}
		spread_amount = amount-mx+used;                               // :122
	}
	// ---- full spread over the node's flows, newest first (list order of flow.c:129,157)
	if (c.scalar) { set_status(c, NS_E_ARG); return; }                // classified away by sc_needs_full_spread
	WSYNC();
	const int g = n*2+dir;
	const int off = S.adj_off[g], len = S.adj_len[g];
	stage_ensure(c, (dir == RCV ? 3ull : 1ull)*(u64)S.adj_live[g]+8);   // RCV victims may also re-time (drain+done)
	cnt_add(c, NS_CNT_FULL_SPREADS, 1);
	cnt_add(c, NS_CNT_FLOW_VISITS, 2ull*(u64)S.adj_live[g]);
	const bool dec = amount < NS_EPS, inc = !dec && amount > NS_EPS;
	// Both passes walk the list in chunks of 32 (one flow per lane).  SP_U1/SP_U2 chunks are loaded
	// together so that their dependent gathers (index -> flow -> remote node) overlap; the sums are
	// then accumulated chunk by chunk, lane by lane, i.e. in exactly the list order.
	double e = 0;                                                     // :128-145
	for (int base = len-1; base >= 0; base -= 32*SP_U1) {
		double term[SP_U1];
		#pragma unroll
		for (int u = 0; u < SP_U1; u++) {
			int p = base-32*u-lane;
			int nf = p >= 0 ? S.adj_arena[off+p] : -1;
			double t = 0.0;
			if (nf >= 0 && nf != f) {
				double sp = S.f_speed[dir][nf];
				if (dec) { double ls = get_share(S, nf, !dir); if (sp < ls) t = ls-sp; }
				else if (inc) { double ls = S.f_ub[nf]*mx/total; if (sp > ls) t = sp-ls; }
			}
			term[u] = t;
		}
		#pragma unroll
		for (int u = 0; u < SP_U1; u++) {
			if (base-32*u < 0) break;
			#pragma unroll
			for (int l = 0; l < 32; l++) e += __shfl_sync(FULLMASK, term[u], l);
		}
	}
	if (dec && -amount > e) amount = -e;                              // :149-154
	const bool dec2 = amount < NS_EPS, inc2 = !dec2 && amount > NS_EPS;
	double new_use = 0;                                               // :156-200
	for (int base = len-1; base >= 0; base -= 32*SP_U2) {
		int nfv[SP_U2], adjv[SP_U2], okv[SP_U2];
		double valv[SP_U2], emitv[SP_U2], etv[SP_U2];
		Retime rpv[SP_U2];
		#pragma unroll
		for (int u = 0; u < SP_U2; u++) {
			int p = base-32*u-lane;
			int nf = p >= 0 ? S.adj_arena[off+p] : -1;
			double val = 0.0, emit_speed = 0;
			int adj = 0;
			if (nf == f) val = fs;
			else if (nf >= 0) {
				double sp = S.f_speed[dir][nf];
				val = sp;
				if (dec2) {
					double ls = get_share(S, nf, !dir);
					if (sp < ls) {
						double delta = ls-sp;
						double new_speed = sp-amount*delta/e;         // :171
						if (dir == SND) { S.f_speed[dir][nf] = new_speed; val = new_speed; }   // :172-176
						adj = 1; emit_speed = new_speed;              // :178
					}
				} else if (inc2) {
					double ls = S.f_ub[nf]*mx/total;
					if (sp > ls) {
						double delta = sp-ls;
						double ns = sp-spread_amount*delta/e;         // :187
						S.f_speed[dir][nf] = ns; val = ns;
						adj = 1; emit_speed = ns;                     // :189
					}
				}
			}
			nfv[u] = nf; adjv[u] = adj; valv[u] = val; emitv[u] = emit_speed;
		}
		// inserts in list order: [speed event][drain][done] per adjusted flow
		#pragma unroll
		for (int u = 0; u < SP_U2; u++) {
			double et = 0; int ok = 0;
			rpv[u].add_drain = rpv[u].add_done = 0;
			if (adjv[u]) {
				et = c.now+S.f_delay[nfv[u]];
				ok = 1;
				if (et < c.now) { if (et < c.now-NS_EPS) ok = 0; else et = c.now; }
				if (isnan(et)) ok = 0;
				if (!ok) atomicCAS(&c.st->status, 0, isnan(et) ? NS_E_NAN_TIME : NS_E_BACK_IN_TIME);
				if (dir == RCV && inc2) retime_plan_1(S, c.st, nfv[u], c.now, rpv[u]);   // :193-194
			}
			etv[u] = et; okv[u] = ok;
		}
		#pragma unroll
		for (int u = 0; u < SP_U2; u++) {
			if (base-32*u < 0) break;
			if (dir == RCV && inc2) cnt_add(c, NS_CNT_RETIMES, __popc(__ballot_sync(FULLMASK, adjv[u])));
			int k = okv[u]+rpv[u].add_drain+rpv[u].add_done;
			int incl = k;
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULLMASK, incl, d); if (lane >= d) incl += t; }
			int tot = __shfl_sync(FULLMASK, incl, 31);
			if (tot) {
				u64 seq0 = seq_reserve(c, tot);
				if (c.st->status == 0 || c.st->status == NS_E_RECORDS_LOST) {
					u64 seq = seq0+(u64)(incl-k);
					if (okv[u]) { ring_put(c, seq, etv[u], emitv[u], nfv[u], !dir); seq++; }
					if (adjv[u]) retime_commit_1(c, nfv[u], rpv[u], seq);
				}
			}
			#pragma unroll
			for (int l = 0; l < 32; l++) new_use += __shfl_sync(FULLMASK, valv[u], l);   // :199
		}
	}
	L0 S.n_usage[dir][n] = new_use;                                   // :201-202
	write_usage(c, dir, n, new_use);
	WSYNC();
}

// ---------------------------------------------------------------- adjacency arrays (node->conns[dir])
// list_add at head (flow.c:311-312) == append; list_del (flow.c:251-252) == tombstone.
__device__ __noinline__ void adj_append(Ctx &c, int g, int f, int dir){
	const Sim &S = *c.S;
	int off = S.adj_off[g], cap = S.adj_cap[g], len = S.adj_len[g], live = S.adj_live[g];
	if (len == cap) {
		if (live*2 <= cap && cap > 0) {
			// compact in place (drop tombstones, keep order)
			L0 {
				int w = 0;
				for (int p = 0; p < len; p++) {
					int x = S.adj_arena[off+p];
					if (x >= 0) { S.adj_arena[off+w] = x; S.f_pos[dir][x] = w; w++; }
				}
				S.adj_len[g] = w;
			}
			WSYNC();
			len = live;
		} else {
			int ncap = cap ? cap*2 : 4;
			int top = 0;
			if (c.pre_slot >= 0) { L0 top = atomicAdd(&c.st->adj_top, ncap); top = __shfl_sync(FULLMASK, top, 0); }
			else top = c.st->adj_top;
			if (top+ncap > S.adj_arena_cap) { set_status(c, NS_E_TABLE_FULL); return; }
			WSYNC();
			// move live entries to the new block, lanes strided; positions recomputed serially
			L0 {
				int w = 0;
				for (int p = 0; p < len; p++) {
					int x = S.adj_arena[off+p];
					if (x >= 0) { S.adj_arena[top+w] = x; S.f_pos[dir][x] = w; w++; }
				}
				S.adj_off[g] = top; S.adj_cap[g] = ncap; S.adj_len[g] = w;
				if (c.pre_slot < 0) c.st->adj_top = top+ncap;
			}
			WSYNC();
			off = top; len = live;
		}
	}
	L0 {
		S.adj_arena[off+len] = f; S.f_pos[dir][f] = len;
		S.adj_len[g] = len+1; S.adj_live[g] = live+1;
	}
	WSYNC();
}
__device__ __forceinline__ void adj_remove(Ctx &c, int g, int f, int dir){
	const Sim &S = *c.S;
	L0 {
		int off = S.adj_off[g], pos = S.f_pos[dir][f];
		S.adj_arena[off+pos] = -1;
		S.adj_live[g] -= 1;
		int len = S.adj_len[g];
		while (len > 0 && S.adj_arena[off+len-1] < 0) len--;      // trim trailing tombstones
		S.adj_len[g] = len;
	}
	WSYNC();
}

// ---------------------------------------------------------------- random draws (sim.c:117, flow.c:327)
__device__ __forceinline__ uint32_t draw_random(Ctx &c){
	if (c.pre_slot >= 0) { c.id_pos = c.lseq; return 0; }            // drawn in op order by the init commit
	uint32_t v = 0;
	L0 {
		SimState *st = c.st;
		int32_t f = st->rf, r = st->rr;
		if (c.burns) ns_rand_jump(st->rnd, f, r, c.burns);
		v = ns_rand_step(st->rnd, f, r);
		st->rf = f; st->rr = r;
		st->cnt[NS_CNT_DRAWS] += c.burns+1;
	}
	c.burns = 0;
	return __shfl_sync(FULLMASK, v, 0);
}

// ---------------------------------------------------------------- resources and ranges
__device__ __forceinline__ int store_get(const Sim &S, int node, uint32_t rid){      // store.h:5-10
	for (int r = S.n_res_head[node]; r >= 0; r = S.s_next[r])
		if (S.s_id[r] == rid) return r;
	return -1;
}
__device__ __noinline__ int resource_new(Ctx &c, int node, uint32_t rid, long long len){          // data.h:241-249, store.h:25-35
	const Sim &S = *c.S;
	int i = 0;
	if (c.pre_slot >= 0) { L0 i = atomicAdd(&c.st->hi_res, 1); i = __shfl_sync(FULLMASK, i, 0); }
	else i = c.st->hi_res;
	if (i >= S.cap_res) { set_status(c, NS_E_TABLE_FULL); return -1; }
	WSYNC();
	L0 {
		S.s_id[i] = rid; S.s_len[i] = len; S.s_node[i] = node; S.s_first[i] = -1;
		S.s_nprod[i] = 0; S.s_ncons[i] = 0;
		S.s_next[i] = S.n_res_head[node];
		__threadfence();                                              // concurrent store_get() of another op may walk this list
		S.n_res_head[node] = i;
		if (c.pre_slot < 0) c.st->hi_res = i+1;
	}
	WSYNC();
	return i;
}
// range_get, range.h:30-41 with range_include_cmp :10-20 (key is an int)
__device__ int range_get(const Sim &S, int res, int start){
	int r = S.s_first[res];
	for (; r >= 0; r = S.r_next[r]) {
		long long rs = S.r_start[r];
		int cmp;
		if (rs <= (long long)start) cmp = ((double)rs+S.r_len[r] >= (double)start) ? 0 : -1;
		else cmp = 1;
		if (cmp >= 0) break;
	}
	if (r < 0) return -1;
	if (S.r_start[r] > (long long)start) return -1;
	return r;
}
// resource_new_range, range.h:71-82: sorted insert by start; the skip-list insert burns one random()
__device__ __noinline__ int resource_new_range(Ctx &c, int res, long long start, double len, double last){
	const Sim &S = *c.S;
	int i = c.pre_slot >= 0 ? c.st->hi_range+c.pre_slot : c.st->free_range;      // parallel init: slot = base + ordinal
	if (c.pre_slot >= 0) { if (i >= S.cap_ranges) { set_status(c, NS_E_TABLE_FULL); return -1; } }
	else if (i >= 0) { WSYNC(); L0 c.st->free_range = S.r_next_free[i]; }
	else {
		i = c.st->hi_range;
		if (i >= S.cap_ranges) { set_status(c, NS_E_TABLE_FULL); return -1; }
		WSYNC();
		L0 c.st->hi_range = i+1;
	}
	c.burns += 1;
	L0 {
		S.r_start[i] = start; S.r_len[i] = len; S.r_lenc[i] = 0; S.r_last[i] = last;
		S.r_total[i] = S.s_len[res]; S.r_res[i] = res; S.r_prod[i] = -1; S.r_cons[i] = -1;
		int prev = -1, cur = S.s_first[res];
		while (cur >= 0 && (int)((u64)S.r_start[cur]-(u64)(long long)(int)start) < 0) { prev = cur; cur = S.r_next[cur]; }
		S.r_prev[i] = prev; S.r_next[i] = cur;
		if (prev >= 0) S.r_next[prev] = i; else S.s_first[res] = i;
		if (cur >= 0) S.r_prev[cur] = i;
	}
	WSYNC();
	return i;
}

// range_merge_with_next, range.c:83-124
__device__ __noinline__ void range_merge_with_next(Ctx &c, int r){
	const Sim &S = *c.S;
	int nr = S.r_next[r];
	if (nr < 0) return;
	WSYNC();
	L0 {
		S.r_len[r] = (double)(u64)((u64)S.r_start[nr]-(u64)S.r_start[r])+S.r_len[nr];   // :94
		int p = S.r_prod[nr];
		S.r_prod[r] = p;                                              // :96-98
		if (p >= 0) S.f_drng[p] = r;
		while (S.r_cons[nr] >= 0) {                                   // :100-107
			int f = S.r_cons[nr];
			int nx = S.f_cnext[f];
			S.r_cons[nr] = nx;
			if (nx >= 0) S.f_cprev[nx] = -1;
			S.f_cprev[f] = -1; S.f_cnext[f] = S.r_cons[r];
			if (S.r_cons[r] >= 0) S.f_cprev[S.r_cons[r]] = f;
			S.r_cons[r] = f;
			S.f_srng[f] = r;
		}
		int nn = S.r_next[nr];                                        // :109-110
		S.r_next[r] = nn;
		if (nn >= 0) S.r_prev[nn] = r;
		S.r_next_free[nr] = c.st->free_range; c.st->free_range = nr;
		S.r_res[nr] = -1;                                             // (a free slot: the host view skips it)
		range_update_1(S, r, c.now);                                  // :113
	}
	WSYNC();
	for (int f = S.r_cons[r]; f >= 0; f = S.f_cnext[f]) {             // :114-117
		L0 range_update_1(S, S.f_drng[f], c.now);
		WSYNC();
		range_calc_and_requeue_events(c, f);
	}
	int p = S.r_prod[r];
	if (p >= 0) {                                                     // :119-123
		L0 range_update_1(S, S.f_srng[p], c.now);
		WSYNC();
		range_calc_and_requeue_events(c, p);
	}
}

#ifdef NS_HOST_CORE_OPS
// node_del_resource, sim.h:47-53: range_clear + store_delete (only reached through the host core: p2p.c evicts resources)
__device__ __noinline__ void res_delete(Ctx &c, int res){
	const Sim &S = *c.S;
	WSYNC();
	L0 {
		const int node = S.s_node[res];
		int prev = -1, cur = S.n_res_head[node];
		while (cur >= 0 && cur != res) { prev = cur; cur = S.s_next[cur]; }
		if (cur == res) { if (prev >= 0) S.s_next[prev] = S.s_next[res]; else S.n_res_head[node] = S.s_next[res]; }
		for (int g = S.s_first[res]; g >= 0; ) {
			const int nx = S.r_next[g];
			S.r_res[g] = -1; S.r_next_free[g] = c.st->free_range; c.st->free_range = g;
			g = nx;
		}
		S.s_first[res] = -1;
		if (S.s_autodel) S.s_autodel[res] = 0;
	}
	WSYNC();
}
#endif

// ---------------------------------------------------------------- flows
// flow_create, flow.c:296-337
__device__ __noinline__ int flow_create(Ctx &c, int src, int dst, double ub, double delay){
	const Sim &S = *c.S;
	int f = c.pre_slot >= 0 ? c.pre_slot : c.st->free_flow;
	if (c.pre_slot >= 0) { if (f >= S.cap_flows) { set_status(c, NS_E_TABLE_FULL); return -1; } }
	else if (f >= 0) { WSYNC(); L0 c.st->free_flow = S.f_next_free[f]; }
	else {
		f = c.st->hi_flow;
		if (f >= S.cap_flows) { set_status(c, NS_E_TABLE_FULL); return -1; }
		WSYNC();
		L0 c.st->hi_flow = f+1;
	}
	L0 {
		S.f_ub[f] = ub; S.f_delay[f] = delay; S.f_speed[0][f] = 0; S.f_speed[1][f] = 0;
		S.f_peer[0][f] = src; S.f_peer[1][f] = dst;
		S.f_srng[f] = -1; S.f_drng[f] = -1; S.f_cnext[f] = -1; S.f_cprev[f] = -1;
		S.f_alive[f] = 1; S.f_drain_type[f] = 0; S.f_ord[f] = -1;
		S.l_time[2*f] = NS_INF; S.l_time[2*f+1] = NS_INF;
		S.n_total[0][src] += ub;                                      // :314-315
		S.n_total[1][dst] += ub;
	}
	WSYNC();
	adj_append(c, src*2+0, f, 0);                                     // :311-312
	adj_append(c, dst*2+1, f, 1);
	if (c.st->status) return f;
	double spd0 = get_share(S, f, 0);                                 // :317-318
	bwspread(c, f, spd0, 0, 0);
	queue_speed_event(c, f, RCV, S.f_speed[0][f]);                    // :322
	uint32_t id = draw_random(c);                                     // :326-330
	L0 S.f_id[f] = id;
	write_record(c, R_CONN_CREATE, S.n_id[src], 4, id, 0);            // :333-334
	write_record(c, R_CONN_DST, id, 4, S.n_id[dst], 0);
	WSYNC();
	return f;
}

// flow_close, flow.c:241-292
__device__ __noinline__ void flow_close(Ctx &c, int f){
	const Sim &S = *c.S;
	int src = S.f_peer[0][f], dst = S.f_peer[1][f];
	double ub = S.f_ub[f];
	WSYNC();
	L0 { S.n_total[0][src] -= ub; S.n_total[1][dst] -= ub; }          // :246-247
	WSYNC();
	flow_range_update(c, f);                                          // :248
	if (c.st->status) return;
	bwspread(c, f, S.f_speed[0][f], 0, 1);                            // :249-250
	bwspread(c, f, S.f_speed[1][f], 1, 1);
	adj_remove(c, src*2+0, f, 0);                                     // :251-252
	adj_remove(c, dst*2+1, f, 1);
	int sr = S.f_srng[f], dr = S.f_drng[f];
	L0 {
		S.f_gen[f] += 1;                                              // :260-270 cancels its pending speed events
		if (dr >= 0 && S.r_prod[dr] == f) S.r_prod[dr] = -1;          // :273-274
		S.s_ncons[S.r_res[sr]] -= 1;                                  // :275-276
		S.s_nprod[S.r_res[dr]] -= 1;
		int cp = S.f_cprev[f], cn = S.f_cnext[f];                     // :282
		if (cp >= 0) S.f_cnext[cp] = cn; else S.r_cons[sr] = cn;
		if (cn >= 0) S.f_cprev[cn] = cp;
		S.l_time[2*f] = NS_INF; S.l_time[2*f+1] = NS_INF;             // :283-284
		S.f_alive[f] = 0;
		int ord = S.f_ord[f];
		if (ord >= 0) S.ord2flow[ord] = -1;
		S.f_next_free[f] = c.st->free_flow; c.st->free_flow = f;
		c.st->cnt[NS_CNT_CLOSES] += 1; c.st->cnt[NS_CNT_LIVE_FLOWS] -= 1;
	}
#ifdef NS_HOST_CORE_OPS
	if (S.s_autodel) {                                                // :278-281
		const int sres = S.r_res[sr];
		if (S.s_autodel[sres] && S.s_ncons[sres] == 0 && S.s_nprod[sres] == 0) res_delete(c, sres);
	}
#endif
	write_record(c, R_CONN_CLOSE, S.f_id[f], 1, 0, 0);                // :289-290
	WSYNC();
}

// sim_establish_flow, sim.c:42-94
__device__ __noinline__ int sim_establish_flow(Ctx &c, uint32_t rid, long long start, int src, int dst, double ub, double delay){
	const Sim &S = *c.S;
	int sr = store_get(S, src, rid);
	if (sr < 0) return -1;                                            // :49-53
	int dr = store_get(S, dst, rid);
	if (dr < 0) {                                                     // :54-61
		dr = resource_new(c, dst, rid, S.s_len[sr]);
		if (dr >= 0) { L0 S.s_brate[dr] = S.s_brate[sr]; WSYNC(); }
	}
	if (dr < 0) return -1;
	if (range_get(S, dr, (int)start) >= 0) return -1;                 // :62-67
	int rng = range_get(S, sr, (int)start);                           // :69
	int f = flow_create(c, src, dst, ub, delay);                      // :70
	if (f < 0 || (c.st->status && c.st->status != NS_E_RECORDS_LOST)) return f;
	int dg = resource_new_range(c, dr, start, 0, c.now);              // :74-76
	if (dg < 0) return f;
	L0 { S.f_drng[f] = dg; S.r_prod[dg] = f; S.f_srng[f] = rng; }     // :75-77
	WSYNC();
	if (rng < 0) { set_status(c, NS_E_NULL_RANGE); return f; }
	if (S.r_prod[rng] >= 0) { L0 c.st->relay_seen = 1; }
	range_calc_and_requeue_events(c, f);                              // :79
	L0 {
		S.f_cprev[f] = -1; S.f_cnext[f] = S.r_cons[rng];              // :80
		if (S.r_cons[rng] >= 0) S.f_cprev[S.r_cons[rng]] = f;
		S.r_cons[rng] = f;
		S.s_ncons[sr] += 1; S.s_nprod[dr] += 1;                       // :82-83
	}
	WSYNC();
	int ph = S.r_prev[dg];                                            // :86-91
	if (ph >= 0) range_calc_and_requeue_events(c, S.r_prod[ph]);
	return f;
}


#ifndef NS_NO_CLIENTS   // (the grid engine of plain simulations is compiled without the behaviour hooks: ns_par.cu / ns_par_cl.cu)
// ---------------------------------------------------------------- behaviour library, client side
// modules/client_behaviour.c on the device (scripts with NS_SC_F_CLIENTS): the water-mark state machine a p2p
// client runs on every RCV speed change, FLOW_DONE and throttle of its flows, and on its own USER event d->e.
// Everything a handler touches belongs to the client node or to the flow that triggered it, so the hooks are
// legal wherever the DEFAULT handler of the same event is (one (node,RCV) group).  The C arithmetic of the
// reference is kept conversion for conversion (size_t/int/double mixes, int truncation of buffer_pos).
__device__ __forceinline__ int cl_slot(const Sim &S, int n){ return 2*S.cap_flows+S.cap_users+n; }
// sim_node_change_state, sim.h:25-28
__device__ __forceinline__ void node_change_state(Ctx &c, int n, int state){
	const Sim &S = *c.S;
	L0 S.n_state[n] = state;
	WSYNC();
	write_record(c, R_NODE_STATE, S.n_id[n], 1, (uint32_t)state & 0xffu, 0);
}
// event_remove(d->e), event.h:13-18
__device__ __forceinline__ void cl_event_remove(Ctx &c, int n){
	L0 c.S->l_time[cl_slot(*c.S, n)] = NS_INF;
	WSYNC();
}
// ue->type / ue->data / d->e->time set, then event_add(d->e), event.h:56-74 (an event at +inf burns its draw and never fires)
__device__ __forceinline__ void cl_event_add(Ctx &c, int n, int type, int rng, double t){
	const Sim &S = *c.S;
	if (!ev_time_ok(c, t)) return;
	u64 seq = seq_reserve(c, 1);
	if (c.st->status && c.st->status != NS_E_RECORDS_LOST) return;
	L0 { S.c_ue_type[n] = type; S.c_ue_rng[n] = rng; l_put(c, cl_slot(S, n), seq, t); }
	WSYNC();
}
// client_lowwm_event, client_behaviour.c:17-89
__device__ __noinline__ void client_lowwm_event(Ctx &c, int rng){
	const Sim &S = *c.S;
	const int n = S.s_node[S.r_res[rng]];
	L0 range_update_1(S, rng, c.now);                                 // :22
	WSYNC();
	if (S.n_state[n] != NST_PLAYING) return;                          // :23-24
	const int pos = S.c_buffer_pos[n];
	const long long rstart = S.r_start[rng], total = S.r_total[rng];
	if ((u64)rstart > (u64)(long long)pos) return;                    // :27-29 (size_t > int)
	cl_event_remove(c, n);                                            // :35
	const double br = S.c_bit_rate[n];
	const int prod = S.r_prod[rng];
	const double sgrow = prod >= 0 ? S.f_speed[1][prod] : 0;          // :40
	const double rlen = S.r_len[rng];
	if (S.r_next[rng] < 0) {                                          // :41-55
		double time2 = (double)((u64)total-(u64)(long long)pos)/br;
		double time3 = ((double)(u64)(total-rstart)-rlen)/sgrow;
		double d0 = (double)(u64)total-((double)(u64)rstart+rlen);
		if (d0 > -NS_EPS && d0 < NS_EPS) time3 = 0;
		if (time2 > time3) { cl_event_add(c, n, UE_DONE_PLAY, -1, c.now+time2); return; }
	}
	if (br < sgrow) return;                                           // :57-61
	const int limit = S.c_cfg[n].lowwm;
	const double time = ((((double)(u64)rstart+rlen)-(double)pos)-(double)limit)/(br-sgrow);   // :65
	// :78-88 (the reference compares the RELATIVE time with the absolute event times)
	if (prod < 0 || (!is_later_than(time, S.l_time[2*prod]) && !is_later_than(time, S.l_time[2*prod+1])))
		cl_event_add(c, n, UE_PAUSE_BUFFERING, rng, time+c.now);
}
// client_highwm_event, client_behaviour.c:91-139
__device__ __noinline__ void client_highwm_event(Ctx &c, int rng){
	const Sim &S = *c.S;
	L0 range_update_1(S, rng, c.now);                                 // :92
	WSYNC();
	const int n = S.s_node[S.r_res[rng]];
	if (S.n_state[n] != NST_STALE) return;                            // :98-100
	const int nh = S.r_next[rng];
	const int pos = S.c_buffer_pos[n];
	const long long rstart = S.r_start[rng], total = S.r_total[rng];
	if ((u64)(long long)pos < (u64)rstart) return;                    // :102-105 (int < size_t)
	cl_event_remove(c, n);                                            // :111
	const double rlen = S.r_len[rng];
	const int re = (int)(((double)(u64)rstart+rlen)-(double)pos);     // :113
	const int prod = S.r_prod[rng];
	const double sgrow = prod >= 0 ? S.f_speed[1][prod] : 0;
	const int highwm = S.c_cfg[n].highwm;
	double time = (double)(highwm-re)/sgrow;                          // :115
	if (re >= highwm) time = 0;
	if (nh < 0) {                                                     // :118-125: reaching the end counts as high water
		double time2 = ((double)(u64)(total-rstart)-rlen)/sgrow;
		double d0 = (double)(u64)total-((double)(u64)rstart+rlen);
		if (d0 > -NS_EPS && d0 < NS_EPS) time2 = 0;
		if (time2 < time) time = time2;
	}
	time += c.now;
	if (prod < 0 || (!is_later_than(time, S.l_time[2*prod]) && !is_later_than(time, S.l_time[2*prod+1])))
		cl_event_add(c, n, UE_DONE_BUFFERING, rng, time);
}
// the part client_speed_change / client_done / client_speed_throttle share (client_behaviour.c:270-317)
__device__ __forceinline__ void client_touch(Ctx &c, int f){
	const Sim &S = *c.S;
	const int n = S.f_peer[1][f];
	if (S.c_cfg[n].kind != NS_CK_CLIENT) return;
	L0 {
		if (S.n_state[n] == NST_PLAYING)                              // d->buffer_pos += d->bit_rate*(s->now-d->last_update)+eps (int += double)
			S.c_buffer_pos[n] = (int)((double)S.c_buffer_pos[n]+(S.c_bit_rate[n]*(c.now-S.c_last[n])+NS_EPS));
		S.c_last[n] = c.now;
	}
	WSYNC();
	const int rng = S.f_drng[f];
	client_lowwm_event(c, rng);
	client_highwm_event(c, rng);
}
// the client's own USER event: client_next_state_from_event + client_handle_next_state (client_behaviour.c:143-196),
// as modules/p2p.c:59-67 dispatches them (the next NEW_CONNECTION of a client that finished playing is the script's business)
__device__ __noinline__ void client_user_event(Ctx &c, int n){
	const Sim &S = *c.S;
	const int type = S.c_ue_type[n];
	int next = NST_DONE;
	if (type != UE_DONE_PLAY) {                                       // :145-151 returns before d->trigger is touched
		L0 S.c_trigger[n] = S.c_ue_rng[n];
		next = type == UE_PAUSE_BUFFERING ? NST_STALE : NST_PLAYING;
	}
	L0 S.c_next_state[n] = next;
	WSYNC();
	const int rng = S.c_trigger[n];
	const int o_state = S.n_state[n];
	if (o_state == next) return;                                      // :173-174
	node_change_state(c, n, next);
	if (o_state == NST_PLAYING) {                                     // :178-183
		L0 {
			S.c_buffer_pos[n] = (int)((double)S.c_buffer_pos[n]+(S.c_bit_rate[n]*(c.now-S.c_last[n])+NS_EPS));
			S.c_last[n] = c.now;
		}
		WSYNC();
		if (next == NST_STALE) client_highwm_event(c, rng);
	} else if (o_state == NST_STALE) {                                // :184-190
		L0 S.c_last[n] = c.now;
		WSYNC();
		if (next == NST_PLAYING) client_lowwm_event(c, rng);
	}
}

#endif

// ---------------------------------------------------------------- handlers
// Would this SPEED_CHANGE reach bwspread's full spread (flow.c:126)?  Pure function of the current
// state, same branch sequence as flow.c:51-124 with close == 0.  Lets the grid-parallel path run the
// O(1) solves one per THREAD and keep whole warps for the O(degree) ones.
__device__ __forceinline__ bool sc_needs_full_spread(const Sim &S, int f, int type, double speed){
	double fs = S.f_speed[type][f];
	double amount = speed-fs;                                         // flow.c:359
	int n = S.f_peer[type][f];
	double total = S.n_total[type][n], mx = S.n_max[type][n], used = S.n_usage[type][n], ub = S.f_ub[f];
	double share = total > mx ? ub*mx/total : ub;
	if (amount > NS_EPS && fs+amount > share) { if (fs > share) return false; amount = share-fs; }
	if (amount > -NS_EPS && amount < NS_EPS) return false;
	if (amount < NS_EPS && amount > -64) return false;
	if (total < mx+NS_EPS) return false;
	if (amount > -NS_EPS && used+amount < mx+NS_EPS) return false;
	return true;
}

#ifdef NS_PREFETCH
// One simulation stepped by one warp / one CTA pays the latency of every dependent table walk of an event in full
// (there is nothing else to switch to), and the structure-of-arrays layout puts every field of a flow, node or
// range into a different line.  So the lines an event will touch are requested up front, in the three waves
// their addresses become known: the flow's columns; its node's and its two ranges' columns; the source range's
// producer and the next range.  Prefetches only: no value is used from here.  (Engines that keep thousands of
// events in flight — the grid kernel — do not need this and are compiled without it.)
__device__ __forceinline__ void pf(const void *p){ asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__device__ __forceinline__ void prefetch_event(const Sim &S, int f, int dir){
	pf(&S.f_speed[0][f]); pf(&S.f_speed[1][f]); pf(&S.f_ub[f]); pf(&S.f_delay[f]); pf(&S.f_id[f]); pf(&S.f_ord[f]);
	pf(&S.l_time[2*f]);
	const int n = S.f_peer[dir][f], sr = S.f_srng[f], dr = S.f_drng[f];
	pf(&S.n_total[dir][n]); pf(&S.n_max[dir][n]); pf(&S.n_usage[dir][n]); pf(&S.n_id[n]);
	if (dir == RCV && sr >= 0 && dr >= 0) {
		pf(&S.r_last[sr]); pf(&S.r_len[sr]); pf(&S.r_lenc[sr]); pf(&S.r_start[sr]);
		pf(&S.r_last[dr]); pf(&S.r_len[dr]); pf(&S.r_lenc[dr]); pf(&S.r_start[dr]); pf(&S.r_total[dr]); pf(&S.r_cons[dr]);
		if (S.has_clients) { pf(&S.c_cfg[n]); pf(&S.n_state[n]); pf(&S.c_buffer_pos[n]); pf(&S.c_bit_rate[n]); pf(&S.c_last[n]); pf(&S.r_res[dr]); }
		const int prod = S.r_prod[sr], nx = S.r_next[dr];
		if (prod >= 0) { pf(&S.f_speed[1][prod]); pf(&S.l_time[2*prod]); }
		if (nx >= 0) pf(&S.r_start[nx]);
	}
}
#endif

// handle_speed_change (flow.c:341-384) + speed_change_free (flow.c:386-389)
__device__ void handle_speed_change(Ctx &c, int f, int type, double speed){
	const Sim &S = *c.S;
#ifdef NS_PREFETCH
	if (c.scalar || c.lane == 0) prefetch_event(S, f, type);
#endif
	if (type == RCV) flow_range_update(c, f);                         // :355-357
	double delta = speed-S.f_speed[type][f];                          // :359
	bwspread(c, f, delta, type, 0);                                   // :361
	if (type == RCV) {                                                // :367-374
		range_calc_and_requeue_events(c, f);
		range_update_consumer_events(c, S.f_drng[f]);
	} else if (delta > 0)
		queue_speed_event(c, f, RCV, S.f_speed[0][f]);
	write_record(c, R_SPD|type, S.f_id[f], -1, 0, speed);             // :383
	trace_row(c, EV_SPEED_CHANGE|(type << 8), S.f_ord[f], speed, S.f_speed[type][f]);
#ifndef NS_NO_CLIENTS
	if (S.has_clients && type == RCV) client_touch(c, f);             // HNDR_USER: client_speed_change, client_behaviour.c:285-299
#endif
}


__device__ int note_created_flow(Ctx &c, int f);
#ifndef NS_NO_CLIENTS
// ---------------------------------------------------------------- client_new_play1 (NS_OP_PLAY1)
#define NS_PICK_MAX 128
// is_connected, flow.h:29-35: a flow src -> dst exists (walks dst's receive list, the short one)
__device__ __forceinline__ bool is_connected(const Sim &S, int src, int dst){
	const int g = dst*2+RCV, off = S.adj_off[g], len = S.adj_len[g];
	for (int p = 0; p < len; p++) { int f = S.adj_arena[off+p]; if (f >= 0 && S.f_peer[0][f] == src) return true; }
	return false;
}
// distance_metric / share_metric, modules/p2p_common.h:171-198
__device__ __forceinline__ int fetch_metric(const Sim &S, int metric, int node, int client){
	int d = S.c_cfg[node].time_zone-S.c_cfg[client].time_zone;
	if (d < 0) d = -d;
	if (metric == 0) return d;
	const double bw = 4000-100*d;
	const double total = S.n_total[0][node]+bw, mx = S.n_max[0][node];
	return total > mx ? (int)(-bw*mx/total) : (int)(-bw);
}
// client_recalc_state, client_behaviour.c:228-247
__device__ __forceinline__ int client_recalc_state(const Sim &S, int n, int res){
	const int pos = S.c_buffer_pos[n];
	const int rng = range_get(S, res, pos);
	if (rng < 0) return NST_STALE;
	const long long total = S.r_total[rng];
	if ((u64)(long long)pos == (u64)total) return NST_DONE;
	const double end = (double)(u64)S.r_start[rng]+S.r_len[rng];
	if (end == (double)(u64)total) return NST_PLAYING;
	if (end > (double)(pos+S.c_cfg[n].lowwm)) return NST_PLAYING;
	return NST_STALE;
}
// client_start_play, client_behaviour.c:249-268
__device__ __noinline__ void client_start_play(Ctx &c, int n, uint32_t rid){
	const Sim &S = *c.S;
	const int res = store_get(S, n, rid);
	if (res < 0) { set_status(c, NS_E_NULL_RANGE); return; }         // the reference dereferences NULL here
	int rng = range_get(S, res, 0);
	L0 { S.c_buffer_pos[n] = 0; S.c_last[n] = c.now; S.c_bit_rate[n] = (double)S.s_brate[res]; }
	WSYNC();
	node_change_state(c, n, client_recalc_state(S, n, res));
	L0 S.c_next_state[n] = S.n_state[n];
	WSYNC();
	if (rng < 0) rng = S.s_first[res];                                // nothing downloaded yet: the first range
	if (rng < 0) { set_status(c, NS_E_NULL_RANGE); return; }
	client_lowwm_event(c, rng);
	client_highwm_event(c, rng);
}
// client_new_play1, client_behaviour.c:364-387, with server_picker_opt1 (client_behaviour.h:68-88) and the
// reference's quick-select (modules/qselect.h:5-27: Lomuto, last element as pivot, `<=` goes left)
__device__ __noinline__ void client_new_play1(Ctx &c, int n, uint32_t rid, int metric){
	const Sim &S = *c.S;
	if (store_get(S, n, rid) >= 0) { client_start_play(c, n, rid); return; }   // :366-372
	int cnt = S.n_srv/2;
	if (cnt <= 0) cnt = 1;
	int v[NS_PICK_MAX], nd[NS_PICK_MAX];
	int cc = 0;
	if (S.n_srv > NS_PICK_MAX) { set_status(c, NS_E_TABLE_FULL); return; }
	for (int i = 0; i < S.n_srv; i++) {                               // list order: newest server first
		const int sv = S.srv_list[i];
		if (is_connected(S, sv, n)) continue;
		v[cc] = fetch_metric(S, metric, sv, n); nd[cc] = sv; cc++;
	}
	if (cc <= cnt) cnt = cc;
	else {
		int base = 0, nmemb = cc, kq = cnt;                           // qselect_eval(eval_table, c, *count)
		while (true) {
			int st = 0;
			const int pv = v[base+nmemb-1];
			for (int i = 0; i < nmemb-1; i++) {
				const int vi = v[base+i];
				if ((int)((unsigned)vi-(unsigned)pv) > 0) continue;   // cmp(a,b) = a->val-b->val
				int tv = v[base+st], tn = nd[base+st];
				v[base+st] = vi; nd[base+st] = nd[base+i];
				v[base+i] = tv; nd[base+i] = tn;
				st++;
			}
			int tv = v[base+st], tn = nd[base+st];
			v[base+st] = v[base+nmemb-1]; nd[base+st] = nd[base+nmemb-1];
			v[base+nmemb-1] = tv; nd[base+nmemb-1] = tn;
			if (kq-1 == st) break;
			if (st > kq-1) nmemb = st;
			else { base += st; nmemb -= st; kq -= st; }
		}
	}
	if (cnt <= 0) { set_status(c, NS_E_ARG); return; }               // assert(cnt), :380
	const int r1 = store_get(S, nd[0], rid);
	if (r1 < 0) { set_status(c, NS_E_NULL_RANGE); return; }
	const int strip = (int)((u64)S.s_len[r1]/(u64)(long long)cnt);   // :382
	for (int i = 0; i < cnt; i++) {                                   // client_new_connection, client_behaviour.c:198-226
		if (c.st->status) return;
		int d = S.c_cfg[nd[i]].time_zone-S.c_cfg[n].time_zone;
		if (d < 0) d = -d;
		const double ub = 4000-100*d, delay = 0.02*d+0.1;             // distance_based_bw / distance_based_delay, p2p_common.h:181-189,214-221
		const int f = sim_establish_flow(c, rid, (long long)(u64)(long long)(i*strip), nd[i], n, ub, delay);
		if (note_created_flow(c, f) < 0) return;
	}
	client_start_play(c, n, rid);
}

#endif

__device__ void run_group(Ctx &c, int g);

// flow_done_handler (flow.c:391-401), the script's HNDR_USER slot, flow_done_cleaner (flow.c:403-406)
__device__ __noinline__ void handle_flow_done(Ctx &c, int f){
	const Sim &S = *c.S;
	flow_range_update(c, f);
	int drng = S.f_drng[f];
	int next = S.r_next[drng];
	if (next >= 0) {
		flow_range_update(c, S.r_prod[next]);
		range_merge_with_next(c, drng);
	}
	int ord = S.f_ord[f];
	trace_row(c, EV_FLOW_DONE, ord, S.r_len[S.f_drng[f]], S.f_speed[1][f]);
#ifndef NS_NO_CLIENTS
	if (S.has_clients) client_touch(c, f);                            // HNDR_USER: client_done, client_behaviour.c:301-317 (via p2p_done)
#endif
	if (S.n_ondone) {
		int lo = 0, hi = S.n_ondone-1, g = -1;
		while (lo <= hi) {
			int m = (lo+hi)/2;
			int fo = S.ondone[m].flow_ord;
			if (fo == ord) { g = S.ondone[m].group; break; }
			if (fo < ord) lo = m+1; else hi = m-1;
		}
		if (g >= 0) run_group(c, g);
	}
	if (c.st->status) return;
	if (c.st->host_mask & (1u << EV_FLOW_DONE)) {                     // module handler runs between the DEFAULT
		L0 c.st->pending_close = f;                                   // and CLEANER handlers (data.h:126-130)
		WSYNC();
		return;
	}
	flow_close(c, f);
}

// flow_throttle_handler, flow.c:408-423
__device__ __noinline__ void handle_throttle(Ctx &c, int f){
	const Sim &S = *c.S;
	int prod = S.r_prod[S.f_srng[f]];
	if (prod < 0) { set_status(c, NS_E_NO_PRODUCER); return; }
	double delta = S.f_speed[1][prod]-S.f_speed[0][f];                // :412
	if (!(delta >= NS_EPS)) {
		flow_range_update(c, f);                                      // :416
		bwspread(c, f, delta, SND, 0);                                // :418
		delta = S.f_speed[1][prod]-S.f_speed[1][f];                   // :419
		bwspread(c, f, delta, RCV, 0);
		range_calc_and_requeue_events(c, f);                          // :421-422
		range_update_consumer_events(c, S.f_drng[f]);
	}
	trace_row(c, EV_FLOW_SPEED_THROTTLE, S.f_ord[f], S.f_speed[0][f], S.f_speed[1][f]);
#ifndef NS_NO_CLIENTS
	if (S.has_clients) client_touch(c, f);                            // HNDR_USER: client_speed_throttle, client_behaviour.c:270-281
#endif
}

// every sim_establish_flow call takes the next creation ordinal, whether it succeeded or not
__device__ int note_created_flow(Ctx &c, int f){
	const Sim &S = *c.S;
	int ord = c.st->n_ord;
	if (ord >= S.cap_ord) { set_status(c, NS_E_TABLE_FULL); return -1; }
	WSYNC();
	L0 {
		S.ord2flow[ord] = f;
		c.st->n_ord = ord+1;
		if (f >= 0) {
			S.f_ord[f] = ord;
			c.st->cnt[NS_CNT_CREATES] += 1;
			u64 live = ++c.st->cnt[NS_CNT_LIVE_FLOWS];
			if (live > c.st->cnt[NS_CNT_PEAK_FLOWS]) c.st->cnt[NS_CNT_PEAK_FLOWS] = live;
		}
	}
	WSYNC();
	return ord;
}

// one group of script operations == what oracle/ref_scenario.c run_group() does through the module API
__device__ __noinline__ void run_group(Ctx &c, int g){
	const Sim &S = *c.S;
	int lo = S.group_off[g], hi = S.group_off[g+1];
	for (int i = lo; i < hi; i++) {
		if (c.st->status) return;
		ns_sc_op op = S.ops[i];
		switch (op.kind) {
		case NS_OP_CREATE:
		case NS_OP_CREATE_RID: {
			int f = sim_establish_flow(c, op.kind == NS_OP_CREATE ? S.rids[op.c] : (uint32_t)op.c, op.start, op.a, op.b, op.x, op.y);
			if (note_created_flow(c, f) < 0) return;
			break; }
#ifndef NS_NO_CLIENTS
		case NS_OP_PLAY1:
			if (!S.has_clients || S.c_cfg[op.a].kind != NS_CK_CLIENT) { set_status(c, NS_E_BAD_SCENARIO); return; }
			client_new_play1(c, op.a, S.rids[op.c], op.b);
			break;
#endif
		case NS_OP_SCHED: {
			int u = c.st->free_user;
			if (u >= 0) { WSYNC(); L0 c.st->free_user = S.u_next_free[u]; }
			else {
				u = c.st->hi_user;
				if (u >= S.cap_users) { set_status(c, NS_E_TABLE_FULL); return; }
				WSYNC();
				L0 c.st->hi_user = u+1;
			}
			double t = op.c ? c.now+op.x : op.x;
			if (ev_time_ok(c, t)) {
				u64 seq = seq_reserve(c, 1);
				L0 { S.u_group[u] = op.a; l_put(c, 2*S.cap_flows+u, seq, t); }
			}
			WSYNC();
			break; }
		case NS_OP_CLOSE: {
			int f = (op.a < c.st->n_ord) ? S.ord2flow[op.a] : -1;
			if (f >= 0) flow_close(c, f);
			break; }
		case NS_OP_END:
			L0 c.st->exit_flag = 1;                                   // sim.h:42-45
			WSYNC();
			break;
		}
	}
}

// operations injected by a host-side module handler (sim_establish_flow / event_add / flow_close /
// sim_end called from inside a handler), executed at the current `now`
__device__ __noinline__ void run_host_ops(Ctx &c, int n){
	const Sim &S = *c.S;
	for (int i = 0; i < n; i++) {
		if (c.st->status) return;
		ns_sc_op op = S.host_ops[i];
		switch (op.kind) {
		case NS_OP_CREATE:
		case NS_OP_CREATE_RID: {
			int f = sim_establish_flow(c, op.kind == NS_OP_CREATE ? S.rids[op.c] : (uint32_t)op.c, op.start, op.a, op.b, op.x, op.y);
			int ord = c.st->n_ord;
			if (ord >= S.cap_ord) { set_status(c, NS_E_TABLE_FULL); return; }
			WSYNC();
			L0 {
				S.ord2flow[ord] = f;
				c.st->n_ord = ord+1;
				S.host_ops[i].a = f >= 0 ? ord : -1;                  // result: creation ordinal or -1
				if (f >= 0 && op.kind == NS_OP_CREATE_RID) {          // ... and where the host core finds the flow's ranges
					const int dg = S.f_drng[f];
					S.host_ops[i].b = dg >= 0 ? S.r_res[dg] : -1; S.host_ops[i].c = dg; S.host_ops[i].start = S.f_srng[f];
				}
				if (f >= 0) {
					S.f_ord[f] = ord;
					c.st->cnt[NS_CNT_CREATES] += 1;
					u64 live = ++c.st->cnt[NS_CNT_LIVE_FLOWS];
					if (live > c.st->cnt[NS_CNT_PEAK_FLOWS]) c.st->cnt[NS_CNT_PEAK_FLOWS] = live;
				}
			}
			WSYNC();
			break; }
		case NS_OP_SCHED: {                                            // a = host token of the USER event
			int u = c.st->free_user;
			if (u >= 0) { WSYNC(); L0 c.st->free_user = S.u_next_free[u]; }
			else {
				u = c.st->hi_user;
				if (u >= S.cap_users) { set_status(c, NS_E_TABLE_FULL); return; }
				WSYNC();
				L0 c.st->hi_user = u+1;
			}
			double t = op.c ? c.now+op.x : op.x;
			if (ev_time_ok(c, t)) {
				u64 seq = seq_reserve(c, 1);
				L0 { S.u_group[u] = -(op.a+1); l_put(c, 2*S.cap_flows+u, seq, t); S.host_ops[i].b = u; }   // result: the slot, for a later NS_OP_CANCEL
			}
			WSYNC();
			break; }
		case NS_OP_CANCEL: {                                           // event_remove of a module-owned event: the slot is free again
			const int u = op.a >= 0 ? op.a : S.host_ops[-op.a-1].b;   // (a < 0: the slot an NS_OP_SCHED earlier in this batch got)
			if (S.l_time[2*S.cap_flows+u] != NS_INF && S.u_group[u] < 0) {
				WSYNC();
				L0 { S.l_time[2*S.cap_flows+u] = NS_INF; S.u_next_free[u] = c.st->free_user; c.st->free_user = u; }
			}
			WSYNC();
			break; }
		case NS_OP_CLOSE: {
			int f = (op.a >= 0 && op.a < c.st->n_ord) ? S.ord2flow[op.a] : -1;
			if (f >= 0) flow_close(c, f);
			break; }
		case NS_OP_END:
			L0 c.st->exit_flag = 1;
			WSYNC();
			break;
		case NS_OP_NEWNODE: {                                          // sim_create_node, sim.c:113-125 (id drawn and recorded by the host core)
			const int n = op.a;
			L0 {
				S.n_max[0][n] = op.x; S.n_max[1][n] = op.y; S.n_usage[0][n] = 0; S.n_usage[1][n] = 0; S.n_total[0][n] = 0; S.n_total[1][n] = 0;
				S.n_id[n] = (uint32_t)op.c; S.n_res_head[n] = -1;
				S.adj_off[2*n] = 0; S.adj_cap[2*n] = 0; S.adj_len[2*n] = 0; S.adj_live[2*n] = 0;
				S.adj_off[2*n+1] = 0; S.adj_cap[2*n+1] = 0; S.adj_len[2*n+1] = 0; S.adj_live[2*n+1] = 0;
			}
			WSYNC();
			break; }
		case NS_OP_NEWRES: {                                           // sim_node_new_resource / sim_node_add_resource, sim.c:129-154
			int r = store_get(S, op.a, (uint32_t)op.c);
			int g = -1;
			if (r < 0) {
				r = resource_new(c, op.a, (uint32_t)op.c, op.start);
				if (r >= 0) {
					L0 S.s_brate[r] = op.b;
					WSYNC();
					g = resource_new_range(c, r, 0, (double)op.start, 0.0);   // range_new: calloc'ed, last_update = 0 (data.h:202-209)
					c.burns -= 1;                                     // the host core took the skip-list draw itself, in call order
				}
			} else r = -1;                                             // sim.c:144-146: already there
			L0 { S.host_ops[i].b = r; S.host_ops[i].c = g; }
			WSYNC();
			break; }
#ifdef NS_HOST_CORE_OPS
		case NS_OP_DELRES:
			if (op.a >= 0 && op.a < c.st->hi_res) res_delete(c, op.a);
			break;
		case NS_OP_AUTODEL:
			if (S.s_autodel && op.a >= 0 && op.a < c.st->hi_res) { L0 S.s_autodel[op.a] = 1; WSYNC(); }
			break;
#endif
		case NS_OP_RANGE_SET: {
			L0 { S.r_len[op.a] = op.x; S.r_lenc[op.a] = op.y; S.r_last[op.a] = __longlong_as_double(op.start); }
			WSYNC();
			break; }
		}
	}
}
// the status travels back with the ops (slot n of the batch), so the host needs no second read
__device__ __forceinline__ void host_ops_status(Ctx &c, int n){
	L0 { if (n < c.S->host_ops_cap) c.S->host_ops[n].kind = c.st->status; }
	WSYNC();
}

// ---------------------------------------------------------------- register-resident event chains
// A batch can hold thousands of SPEED_CHANGE events for ONE (node, direction) (every full spread at a
// neighbour re-notifies all its flows): that chain is strictly sequential and sets the critical path
// of the batch.  For nodes with <= 32 flows each lane keeps one flow of the node in registers
// (write-through), so an O(1) solve of the chain costs a ballot, the owner lane's arithmetic and a few
// stores instead of a walk through the tables.  Events that need a full spread go through bwspread()
// above and the lanes reload.  Only used inside relay-free parallel segments (source ranges static).
struct GFlow {
	int nf, ord; unsigned id;
	double ub, sp, delay, share;      // share = get_share(f, dir): constant inside a segment (flow.c:61)
	double rshare;                    // get_share(f, !dir), the share at the flow's other end (flow.c:140,165)
	int dr, sr, npos;
	long long d_start, s_start;
	double d_len, d_lenc, d_last, s_len, s_last;
};
__device__ __forceinline__ void ge_load(const Sim &S, GFlow &g, int off, int len, int lane, int dir, double total, double mx){
	g.nf = lane < len ? S.adj_arena[off+lane] : -1;
	if (g.nf < 0) return;
	int f = g.nf;
	g.ord = S.f_ord[f]; g.id = S.f_id[f]; g.ub = S.f_ub[f]; g.sp = S.f_speed[dir][f]; g.delay = S.f_delay[f];
	g.share = total > mx ? g.ub*mx/total : g.ub;
	g.rshare = get_share(S, f, !dir);
	if (dir == RCV) {
		g.dr = S.f_drng[f]; g.sr = S.f_srng[f];
		g.d_start = S.r_start[g.dr]; g.s_start = S.r_start[g.sr];
		g.d_len = S.r_len[g.dr]; g.d_lenc = S.r_lenc[g.dr]; g.d_last = S.r_last[g.dr];
		g.s_len = S.r_len[g.sr]; g.s_last = S.r_last[g.sr];
		int nx = S.r_next[g.dr];
		g.npos = nx >= 0 ? (int)S.r_start[nx] : (int)S.r_total[g.dr];
	}
}
// range_calc_and_requeue_events (range.c:16-79) of the lane's own flow on its registers (source range
// static: relay-free segment), write-through; fills the plan, the caller assigns sequence numbers
__device__ __forceinline__ void ge_retime_regs(const Sim &S, SimState *st, GFlow &g, double now, Retime &p){
	const int f = g.nf;
	p.add_drain = p.add_done = 0; p.drain_t = p.done_t = NS_INF;
	S.l_time[2*f] = NS_INF; S.l_time[2*f+1] = NS_INF;                 // range.c:19-20
	if (now != g.s_last) { g.s_last = now; S.r_last[g.sr] = now; }    // range.c:21 -> range.h:91-118
	if (now != g.d_last) {
		double delta = g.sp*(now-g.d_last), t;
		delta = delta-g.d_lenc;
		t = delta+g.d_len;
		g.d_lenc = (t-g.d_len)-delta;
		g.d_len = t;
		g.d_last = now;
		S.r_lenc[g.dr] = g.d_lenc; S.r_len[g.dr] = g.d_len; S.r_last[g.dr] = now;
	}
	if (g.sp < NS_EPS) return;                                        // range.c:26-27
	int dtype = -1;
	int err = retime_math(now, g.sp, false, 0.0, NS_INF, g.s_len, g.d_len, g.d_start, g.s_start, g.npos, p, dtype);
	if (dtype >= 0) S.f_drain_type[f] = (uint8_t)dtype;
	if (err) atomicCAS(&st->status, 0, err);
}
// queue_speed_event (flow.c:16-29) with the delay taken from the lane's registers
// staging pointers and flags of a chain, read once and kept in registers
struct GeStage {
	double *st_time, *st_speed; int *st_fd; uint8_t *st_rec; ns_trace_row32 *st_tr;
	unsigned *ev_nins; long long *ev_ovf; uint8_t *ev_nrec, *ev_ntr;
	int want_rec, want_tr;
};
__device__ __forceinline__ void ge_stage_load(GeStage &t, const Sim &S, const ParStage &ps){
	t.st_time = ps.st_time; t.st_speed = ps.st_speed; t.st_fd = ps.st_fd; t.st_rec = ps.st_rec; t.st_tr = ps.st_tr;
	t.ev_nins = ps.ev_nins; t.ev_ovf = ps.ev_ovf; t.ev_nrec = ps.ev_nrec; t.ev_ntr = ps.ev_ntr;
	t.want_rec = S.want_rec; t.want_tr = S.want_tr;
}
// one O(1) SPEED_CHANGE (handle_speed_change with every bwspread exit of flow.c:64-124) executed by the
// owner lane alone, everything it needs in registers; e = the event's index in the segment (its
// staging slot: at most 3 event_adds, 2 records, 1 trace row)
__device__ __forceinline__ void ge_simple_event(const GeStage &t, const Sim &S, SimState *st, GFlow &g, double &usage,
		unsigned node_id, int n, int type, double speed, double now, unsigned e, unsigned &n_ret){
	const int f = g.nf;
	unsigned lseq = 0, rec_n = 0;
	const u64 sbase = (u64)e*NS_STAGE_INLINE;
	uint8_t *rbase = t.st_rec+(u64)e*NS_STAGE_REC;
	if (type == RCV) {                                                // flow.c:355-357 -> range.h:91-125
		if (now != g.s_last) { g.s_last = now; S.r_last[g.sr] = now; }
		if (now != g.d_last) {
			double delta = g.sp*(now-g.d_last), tt;
			delta = delta-g.d_lenc;
			tt = delta+g.d_len;
			g.d_lenc = (tt-g.d_len)-delta;
			g.d_len = tt;
			g.d_last = now;
			S.r_lenc[g.dr] = g.d_lenc; S.r_len[g.dr] = g.d_len; S.r_last[g.dr] = now;
		}
	}
	double fs = g.sp;
	const double delta0 = speed-fs;                                   // flow.c:359
	double amount = delta0;
	const double share = g.share;                                     // flow.c:61
	const double et = now+g.delay;                                    // queue_speed_event, flow.c:22
	if (!(et >= now)) atomicCAS(&st->status, 0, isnan(et) ? NS_E_NAN_TIME : NS_E_BACK_IN_TIME);
	bool done_bw = false;
	if (amount > NS_EPS && fs+amount > share) {                       // flow.c:64-78
		double es = share;
		if (fs > share) { es = fs; done_bw = true; } else amount = share-fs;
		t.st_time[sbase+lseq] = et; t.st_speed[sbase+lseq] = es; t.st_fd[sbase+lseq] = f*2+(!type); lseq++;
	}
	if (!done_bw && !(amount > -NS_EPS && amount < NS_EPS)) {         // flow.c:80-124, the O(1) exits
		fs += amount;
		g.sp = fs; S.f_speed[type][f] = fs;
		usage += amount; S.n_usage[type][n] = usage;
		if (t.want_rec) rec_n += rec_encode(rbase+rec_n, now, R_USAGE|type, node_id, -1, 0, usage);   // record_wrapper.h:6-9
	}
	if (type == RCV) {                                                // flow.c:367-370 -> range.c:16-79
		n_ret++;
		Retime p;
		ge_retime_regs(S, st, g, now, p);
		if (p.add_done) {
			if (p.add_drain) { t.st_fd[sbase+lseq] = NS_HOLE_PURE; lseq++; }
			S.l_time[2*f] = p.done_t; S.l_seq[2*f] = NS_SEQ_PROV | ((u64)e << 24) | lseq;
			tmin_lower(&S.l_tmin[(2*f) >> NS_TILE_SHIFT], p.done_t);
			t.st_fd[sbase+lseq] = -(2*f+1); lseq++;
		} else if (p.add_drain) {
			S.l_time[2*f+1] = p.drain_t; S.l_seq[2*f+1] = NS_SEQ_PROV | ((u64)e << 24) | lseq;
			tmin_lower(&S.l_tmin[(2*f+1) >> NS_TILE_SHIFT], p.drain_t);
			t.st_fd[sbase+lseq] = -(2*f+1+1); lseq++;
		}
	} else if (delta0 > 0) {                                          // flow.c:371-374
		t.st_time[sbase+lseq] = et; t.st_speed[sbase+lseq] = g.sp; t.st_fd[sbase+lseq] = f*2+RCV; lseq++;
	}
	if (t.want_rec) rec_n += rec_encode(rbase+rec_n, now, R_SPD|type, g.id, -1, 0, speed);   // flow.c:383
	unsigned ntr = 0;
	if (t.want_tr) {
		ns_trace_row32 r; r.type = EV_SPEED_CHANGE|(type << 8); r.ord = g.ord; r.time = now; r.v0 = speed; r.v1 = g.sp;
		t.st_tr[e] = r; ntr = 1;
	}
	t.ev_nins[e] = lseq; t.ev_ovf[e] = -1; t.ev_nrec[e] = (uint8_t)rec_n; t.ev_ntr[e] = (uint8_t)ntr;
}

// one SPEED_CHANGE that reaches bwspread's full spread (flow.c:126-203), on the register-resident
// group: lane p holds the flow at list position p (newest = highest p), so the reference's in-order
// sums run from lane 31 down to lane 0.  Warp-uniform; `owner` is the lane holding the event's flow.
__device__ void ge_full_event(Ctx &c, const Sim &S, GFlow &g, double &usage, double total, double mx, unsigned node_id,
		int n, int dir, double speed, int owner){
	const int lane = c.lane;
	const double now = c.now;
	const bool me = lane == owner;
	const int live = __popc(__ballot_sync(FULLMASK, g.nf >= 0));
	cnt_add(c, NS_CNT_EVENTS, 1); cnt_add(c, NS_CNT_SPEED_CHANGE, 1); cnt_add(c, NS_CNT_SOLVES, 1);
	cnt_add(c, NS_CNT_FULL_SPREADS, 1); cnt_add(c, NS_CNT_FLOW_VISITS, 2ull*(u64)live);
	if (dir == RCV && me) {                                           // flow.c:355-357, with the OLD speed
		if (now != g.s_last) { g.s_last = now; S.r_last[g.sr] = now; }
		if (now != g.d_last) {
			double delta = g.sp*(now-g.d_last), t;
			delta = delta-g.d_lenc;
			t = delta+g.d_len;
			g.d_lenc = (t-g.d_len)-delta;
			g.d_len = t;
			g.d_last = now;
			S.r_lenc[g.dr] = g.d_lenc; S.r_len[g.dr] = g.d_len; S.r_last[g.dr] = now;
		}
	}
	double fs = __shfl_sync(FULLMASK, g.sp, owner);
	const double share = __shfl_sync(FULLMASK, g.share, owner), odelay = __shfl_sync(FULLMASK, g.delay, owner);
	const int of = __shfl_sync(FULLMASK, g.nf, owner);
	const double delta0 = speed-fs;                                   // flow.c:359
	double amount = delta0;
	if (amount > NS_EPS && fs+amount > share) {                       // flow.c:64-78 (fs <= share here)
		double t = now+odelay;
		if (ev_time_ok(c, t)) { u64 seq = seq_reserve(c, 1); if (me) ring_put(c, seq, t, share, of, !dir); }
		amount = share-fs;
	}
	fs += amount;                                                     // flow.c:85
	if (me) { g.sp = fs; S.f_speed[dir][of] = fs; }
	double spread_amount = amount;                                    // flow.c:112-124
	if (amount > -NS_EPS) spread_amount = amount-mx+usage;
	stage_ensure(c, (dir == RCV ? 3ull : 1ull)*(u64)live+8);
	const bool dec = amount < NS_EPS, inc = !dec && amount > NS_EPS;
	double term = 0.0;                                                // flow.c:128-145
	if (g.nf >= 0 && !me) {
		if (dec) { if (g.sp < g.rshare) term = g.rshare-g.sp; }
		else if (inc) { if (g.sp > g.share) term = g.sp-g.share; }
	}
	double e = 0;
	#pragma unroll
	for (int l = 31; l >= 0; l--) e += __shfl_sync(FULLMASK, term, l);
	if (dec && -amount > e) amount = -e;                              // flow.c:149-154
	const bool dec2 = amount < NS_EPS, inc2 = !dec2 && amount > NS_EPS;
	double val = 0.0, emit_speed = 0;                                 // flow.c:156-200
	int adj = 0;
	Retime rp; rp.add_drain = rp.add_done = 0;
	if (me) val = fs;
	else if (g.nf >= 0) {
		const double sp = g.sp;
		val = sp;
		if (dec2) {
			if (sp < g.rshare) {
				double delta = g.rshare-sp;
				double new_speed = sp-amount*delta/e;                 // :171
				if (dir == SND) { g.sp = new_speed; S.f_speed[dir][g.nf] = new_speed; val = new_speed; }
				adj = 1; emit_speed = new_speed;
			}
		} else if (inc2) {
			if (sp > g.share) {
				double delta = sp-g.share;
				double ns = sp-spread_amount*delta/e;                 // :187
				g.sp = ns; S.f_speed[dir][g.nf] = ns; val = ns;
				adj = 1; emit_speed = ns;
			}
		}
	}
	double et = 0; int ok = 0;
	if (adj) {
		et = now+g.delay;
		ok = 1;
		if (et < now) { if (et < now-NS_EPS) ok = 0; else et = now; }
		if (isnan(et)) ok = 0;
		if (!ok) atomicCAS(&c.st->status, 0, isnan(et) ? NS_E_NAN_TIME : NS_E_BACK_IN_TIME);
		if (dir == RCV && inc2) ge_retime_regs(S, c.st, g, now, rp);  // :193-194, with the NEW speed
	}
	if (dir == RCV && inc2) cnt_add(c, NS_CNT_RETIMES, __popc(__ballot_sync(FULLMASK, adj)));
	{
		int k = ok+rp.add_drain+rp.add_done;
		int incl = k;                                                 // suffix sum: list order is descending lane
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { int t = __shfl_down_sync(FULLMASK, incl, d); if (lane+d < 32) incl += t; }
		int tot = __shfl_sync(FULLMASK, incl, 0);
		if (tot) {
			u64 seq0 = seq_reserve(c, tot);
			u64 seq = seq0+(u64)(incl-k);
			if (ok) { ring_put(c, seq, et, emit_speed, g.nf, !dir); seq++; }
			if (adj) retime_commit_1(c, g.nf, rp, seq);
		}
	}
	double new_use = 0;
	#pragma unroll
	for (int l = 31; l >= 0; l--) new_use += __shfl_sync(FULLMASK, val, l);   // :199
	usage = new_use;                                                  // :201-202
	if (lane == 0) S.n_usage[dir][n] = usage;
	write_record(c, R_USAGE|dir, node_id, -1, 0, usage);
	// the rest of handle_speed_change (flow.c:367-383)
	if (dir == RCV) {
		cnt_add(c, NS_CNT_RETIMES, 1);
		Retime p; p.add_drain = p.add_done = 0;
		if (me) ge_retime_regs(S, c.st, g, now, p);
		p.add_drain = __shfl_sync(FULLMASK, p.add_drain, owner);
		p.add_done = __shfl_sync(FULLMASK, p.add_done, owner);
		int k = p.add_drain+p.add_done;
		if (k) { u64 seq = seq_reserve(c, k); if (me) retime_commit_1(c, of, p, seq); }
	} else if (delta0 > 0) {
		double t = now+odelay;
		if (ev_time_ok(c, t)) { u64 seq = seq_reserve(c, 1); if (me) ring_put(c, seq, t, g.sp, of, RCV); }
	}
	const unsigned oid = __shfl_sync(FULLMASK, g.id, owner);
	const int oord = __shfl_sync(FULLMASK, g.ord, owner);
	const double osp = __shfl_sync(FULLMASK, g.sp, owner);
	write_record(c, R_SPD|dir, oid, -1, 0, speed);
	trace_row(c, EV_SPEED_CHANGE|(dir << 8), oord, speed, osp);
}

}   // namespace
