// ns_seq.cuh — one simulation stepped by ONE warp: next-completion argmin, equal-timestamp batch
// extraction and strictly sequential handling of the batch.  This is the always-legal path (relay
// topologies, FLOW_DONE / USER / THROTTLE events) and the per-replica engine of ensembles; the
// grid-parallel path (ns_par.cuh) reuses its pieces.
//
// Replaces main.c:50-67: `while(!skip_list_empty) { e = event_pop(); now = e->time; handlers }`.
// event_pop (event.h:20-25) becomes: t = min over the per-tile minima of the two event stores;
// the batch is every pending event with time == t, in insertion order (FIFO ties, event.h:27-31).
#pragma once
#include "ns_model.cuh"
#include "ns_ana_dev.cuh"

namespace {   // internal linkage: every engine translation unit carries its own copy of the device code

__device__ __forceinline__ double warp_min_d(double v){
	#pragma unroll
	for (int d = 16; d; d >>= 1) v = fmin(v, __shfl_xor_sync(FULLMASK, v, d));
	return v;
}
__device__ __forceinline__ u64 warp_min_u64(u64 v){
	#pragma unroll
	for (int d = 16; d; d >>= 1) { u64 o = __shfl_xor_sync(FULLMASK, v, d); v = o < v ? o : v; }
	return v;
}


// Step 1: global minimum event time + ring head advance.
__device__ double seq_find_tmin(Ctx &c){
	const Sim &S = *c.S;
	const int lane = c.lane;
	double m = NS_INF;
	const u64 tmask = (u64)S.n_ring_tiles-1;
	u64 first = c.head >> NS_TILE_SHIFT;
	u64 last = c.tail ? (c.tail-1) >> NS_TILE_SHIFT : 0;
	u64 first_live = ~0ull;
	if (c.tail > c.head || (c.tail & (NS_TILE-1))) {
		for (u64 a = first+lane; a <= last; a += 32) {
			double v = S.q_tmin[a & tmask];
			if (v != NS_INF && a < first_live) first_live = a;
			m = fmin(m, v);
		}
	}
	first_live = warp_min_u64(first_live);
	c.head = (first_live == ~0ull) ? ((c.tail >> NS_TILE_SHIFT) << NS_TILE_SHIFT) : (first_live << NS_TILE_SHIFT);
	for (int t = lane; t < S.n_l_tiles; t += 32) m = fmin(m, S.l_tmin[t]);
	return warp_min_d(m);
}

// Step 2a: ordered extraction of the ring entries with time == m.  Returns the count (or -1 when the
// batch scratch overflows).  Tile minima are recomputed without the extracted entries.
__device__ int seq_collect_ring(Ctx &c, double m){
	const Sim &S = *c.S;
	const int lane = c.lane;
	const u64 tmask = (u64)S.n_ring_tiles-1;
	const u64 smask = S.ring_cap-1;
	int nb = 0;
	u64 first = c.head >> NS_TILE_SHIFT;
	u64 last = c.tail ? (c.tail-1) >> NS_TILE_SHIFT : 0;
	if (c.tail <= c.head) return 0;
	for (u64 a0 = first; a0 <= last; a0 += 32) {
		u64 a = a0+lane;
		bool hit = a <= last && S.q_tmin[a & tmask] == m;
		unsigned hits = __ballot_sync(FULLMASK, hit);
		while (hits) {
			int b = __ffs(hits)-1; hits &= hits-1;
			u64 at = a0+b;
			u64 base = at << NS_TILE_SHIFT;
			double nm = NS_INF;
			for (int i0 = 0; i0 < NS_TILE; i0 += 32*8) {              // 8 independent loads in flight per lane
				double tv[8];
				#pragma unroll
				for (int u = 0; u < 8; u++) tv[u] = S.q_time[(base+i0+32*u+lane) & smask];
				#pragma unroll
				for (int u = 0; u < 8; u++) {
					u64 seq = base+i0+32*u+lane;
					double t = tv[u];
					bool match = t == m && seq >= c.head && seq < c.tail;
					unsigned mm = __ballot_sync(FULLMASK, match);
					if (mm) {
						if (nb+__popc(mm) > S.batch_cap) return -1;
						if (match) { int pos = nb+__popc(mm & ((1u << lane)-1)); S.b_seq[pos] = seq; }
						nb += __popc(mm);
					}
					if (!match) nm = fmin(nm, t);
				}
			}
			nm = warp_min_d(nm);
			if (lane == 0) S.q_tmin[at & tmask] = nm;
		}
	}
	__syncwarp();
	return nb;
}

// Step 2b: long-lived events with time == m, sorted by sequence number.
__device__ __noinline__ int seq_collect_l(Ctx &c, double m, const SeqScratch &sc){
	const Sim &S = *c.S;
	const int lane = c.lane;
	int nl = 0;
	for (int t0 = 0; t0 < S.n_l_tiles; t0 += 32) {
		int t = t0+lane;
		bool hit = t < S.n_l_tiles && S.l_tmin[t] == m;
		unsigned hits = __ballot_sync(FULLMASK, hit);
		while (hits) {
			int b = __ffs(hits)-1; hits &= hits-1;
			int tile = t0+b;
			int base = tile << NS_TILE_SHIFT;
			double nm = NS_INF;
			for (int i0 = 0; i0 < NS_TILE; i0 += 32*8) {              // 8 independent loads in flight per lane
				double tv[8];
				#pragma unroll
				for (int u = 0; u < 8; u++) { int slot = base+i0+32*u+lane; tv[u] = slot < S.n_l ? S.l_time[slot] : NS_INF; }
				#pragma unroll
				for (int u = 0; u < 8; u++) {
					int slot = base+i0+32*u+lane;
					double tt = tv[u];
					bool match = tt == m;
					unsigned mm = __ballot_sync(FULLMASK, match);
					if (match) { int pos = nl+__popc(mm & ((1u << lane)-1)); sc.lb_seq[pos] = S.l_seq[slot]; sc.lb_ref[pos] = slot; }
					nl += __popc(mm);
					if (!match) nm = fmin(nm, tt);
				}
			}
			nm = warp_min_d(nm);
			if (lane == 0) S.l_tmin[tile] = nm;
		}
	}
	__syncwarp();
	if (nl > 1) {
		if (nl <= 64) {
			if (lane == 0) {
				for (int i = 1; i < nl; i++) {
					u64 s = sc.lb_seq[i]; int r = sc.lb_ref[i]; int j = i-1;
					while (j >= 0 && sc.lb_seq[j] > s) { sc.lb_seq[j+1] = sc.lb_seq[j]; sc.lb_ref[j+1] = sc.lb_ref[j]; j--; }
					sc.lb_seq[j+1] = s; sc.lb_ref[j+1] = r;
				}
			}
		} else {
			// rank sort (sequence numbers are unique)
			for (int i = lane; i < nl; i += 32) {
				u64 s = sc.lb_seq[i]; int rank = 0;
				for (int j = 0; j < nl; j++) rank += sc.lb_seq[j] < s;
				sc.lb_tmp_seq[rank] = s; sc.lb_tmp_ref[rank] = sc.lb_ref[i];
			}
			__syncwarp();
			for (int i = lane; i < nl; i += 32) { sc.lb_seq[i] = sc.lb_tmp_seq[i]; sc.lb_ref[i] = sc.lb_tmp_ref[i]; }
		}
		__syncwarp();
	}
	return nl;
}

// host-callback mode: remember the event and stop after it when its type is in the mask
__device__ __forceinline__ void host_report(Ctx &c, int type, int f, int dir, double speed, int token){
	SimState *st = c.st;
	if (!(st->host_mask & (1u << type))) return;
	L0 {
		st->le_type = type; st->le_flow = f; st->le_ord = f >= 0 ? c.S->f_ord[f] : -1; st->le_dir = dir;
		st->le_token = token; st->le_time = c.now; st->le_speed = speed;
		st->host_stop = 1;
	}
	WSYNC();
}

// handle one long-lived event (FLOW_DONE / FLOW_DRAIN / FLOW_SPEED_THROTTLE / USER); returns false
// when it had been cancelled or re-timed after extraction.
__device__ __noinline__ bool seq_handle_l(Ctx &c, double m, u64 seq, int slot){
	const Sim &S = *c.S;
	if (!(S.l_time[slot] == m && S.l_seq[slot] == seq)) return false;
	__syncwarp();
	L0 S.l_time[slot] = NS_INF;                                       // event_pop: active = false
	__syncwarp();
	cnt_add(c, NS_CNT_EVENTS, 1);
	if (slot < 2*S.cap_flows) {
		int f = slot >> 1;
		if (!(slot & 1)) { cnt_add(c, NS_CNT_FLOW_DONE, 1); handle_flow_done(c, f); host_report(c, EV_FLOW_DONE, f, 0, 0, -1); }
		else if (S.f_drain_type[f] == EV_FLOW_SPEED_THROTTLE) { cnt_add(c, NS_CNT_THROTTLE, 1); handle_throttle(c, f); host_report(c, EV_FLOW_SPEED_THROTTLE, f, 0, 0, -1); }
		else { cnt_add(c, NS_CNT_DRAIN, 1); trace_row(c, EV_FLOW_DRAIN, S.f_ord[f], 0, 0); host_report(c, EV_FLOW_DRAIN, f, 0, 0, -1); }   // no default handler (main.c:44-48)
#ifndef NS_NO_CLIENTS
	} else if (slot >= 2*S.cap_flows+S.cap_users) {                      // a client's own USER event (d->e, client_behaviour.c:31-34)
		const int n = slot-2*S.cap_flows-S.cap_users;
		cnt_add(c, NS_CNT_USER, 1);
		trace_row(c, EV_USER, 0x40000000u|(unsigned)n, (double)S.c_ue_type[n], 0);
		client_user_event(c, n);
#endif
	} else {
		int u = slot-2*S.cap_flows;
		int g = S.u_group[u];
		L0 { S.u_next_free[u] = c.st->free_user; c.st->free_user = u; }
		__syncwarp();
		cnt_add(c, NS_CNT_USER, 1);
		if (g < 0) {                                                  // a host-owned USER event: its handler runs on the host
			L0 { c.st->le_type = EV_USER; c.st->le_flow = -1; c.st->le_ord = -1; c.st->le_dir = 0; c.st->le_token = -g-1;
			     c.st->le_time = c.now; c.st->le_speed = 0; c.st->host_stop = 1; }
			WSYNC();
			return true;
		}
		trace_row(c, EV_USER, g, 0, 0);
		run_group(c, g);
	}
	return true;
}

// handle one ring (SPEED_CHANGE) entry
__device__ bool seq_handle_ring(Ctx &c, double m, u64 seq){
	const Sim &S = *c.S;
	u64 slot = seq & (S.ring_cap-1);
	if (S.q_time[slot] != m) return false;
	int fd = S.q_fd[slot];
	int f = fd >> 1, dir = fd & 1;
	double speed = S.q_speed[slot];
	bool live = S.f_gen[f] == S.q_gen[slot] && S.f_alive[f];
	__syncwarp();
	L0 S.q_time[slot] = NS_INF;
	__syncwarp();
	if (!live) return false;                                          // cancelled by flow_close (flow.c:260-270)
	cnt_add(c, NS_CNT_EVENTS, 1);
	cnt_add(c, NS_CNT_SPEED_CHANGE, 1);
	handle_speed_change(c, f, dir, speed);
	host_report(c, EV_SPEED_CHANGE, f, dir, speed, -1);
	return true;
}

// records are appended by the handlers; a batch is only started when its worst case fits (two records
// per SPEED_CHANGE, plus a reserve for the creates/closes of USER groups and FLOW_DONE handlers)
#define NS_REC_PER_EVENT 64
#define NS_REC_PER_OP 192          // one scripted create / close: its own records plus the usage records of its solves
#define NS_REC_RESERVE (1u << 20)
// script operations an L-table event will run (a USER event's group, the on-done group of a FLOW_DONE)
__device__ __forceinline__ int l_entry_ops(const Sim &S, int slot){
	if (slot >= 2*S.cap_flows+S.cap_users) return 0;                   // a client's own event: one node-state record at most
	if (slot >= 2*S.cap_flows) { int g = S.u_group[slot-2*S.cap_flows]; return g >= 0 ? (S.group_off[g+1]-S.group_off[g])*(S.has_clients ? S.n_srv/2+2 : 1) : 0; }   // a PLAY1 op is up to nsvr/2 creates
	if ((slot & 1) || !S.n_ondone) return 0;
	const int ord = S.f_ord[slot >> 1];
	int lo = 0, hi = S.n_ondone-1;
	while (lo <= hi) {
		int mid = (lo+hi)/2, fo = S.ondone[mid].flow_ord;
		if (fo == ord) { int g = S.ondone[mid].group; return S.group_off[g+1]-S.group_off[g]; }
		if (fo < ord) lo = mid+1; else hi = mid-1;
	}
	return 0;
}
__device__ __forceinline__ bool rec_room_short(const Ctx &c, u64 n_events, u64 n_ops){
	const Sim &S = *c.S;
	if (!S.want_rec) return false;
	u64 reserve = S.rec_cap > 8ull*NS_REC_RESERVE ? NS_REC_RESERVE : S.rec_cap/8;
	return c.rec_len+n_events*NS_REC_PER_EVENT+n_ops*NS_REC_PER_OP+reserve > S.rec_cap;
}
__device__ __forceinline__ bool seq_should_stop(Ctx &c){
	SimState *st = c.st;
	if (st->status || st->exit_flag || st->host_stop) return true;
	if (st->max_events && st->cnt[NS_CNT_EVENTS] >= st->max_events) return true;
	return false;
}

// the staged event_adds / records of one event, written at their final positions (shared with the grid engine)
__device__ __forceinline__ void par_commit_event(const Sim &S, const ParStage &ps, int e, u64 seq0, u64 rec0, u64 tr0, int lane0, int lstep, bool with_records = true){
	unsigned nins = ps.ev_nins[e];
	long long ovf = ps.ev_ovf[e];
	for (unsigned j = lane0; j < nins; j += lstep) {
		u64 si = j < NS_STAGE_INLINE ? (u64)e*NS_STAGE_INLINE+j : ps.inline_n+(u64)ovf+(j-NS_STAGE_INLINE);
		int fd = ps.st_fd[si];
		u64 seq = seq0+j;
		if (fd >= 0) {
			u64 slot = seq & (S.ring_cap-1);
			double t = ps.st_time[si];
			S.q_time[slot] = t; S.q_speed[slot] = ps.st_speed[si]; S.q_fd[slot] = fd; S.q_gen[slot] = S.f_gen[fd >> 1];
			double *tm = &S.q_tmin[slot >> NS_TILE_SHIFT];
			if (*(volatile double *)tm > t) tmin_lower(tm, t);            // minima only fall during a commit
		} else if (fd != NS_HOLE_PURE) {
			int ls = -fd-1;
			u64 prov = NS_SEQ_PROV | ((u64)e << 24) | j;
			if (S.l_seq[ls] == prov) S.l_seq[ls] = seq;               // only the last add of a slot survives
		}
	}
	if (lane0 == 0 && with_records) {
		unsigned nrec = ps.ev_nrec[e];
		if (nrec) { const uint8_t *src = ps.st_rec+(u64)e*NS_STAGE_REC; uint8_t *dst = S.rec+rec0; for (unsigned k = 0; k < nrec; k++) dst[k] = src[k]; }
		if (ps.ev_ntr[e] && tr0 < S.tr_cap) S.tr[tr0] = ps.st_tr[e];
	}
}

// ---------------------------------------------------------------- wide runs: one event per LANE
// A run of SPEED_CHANGE events of one timestamp, handled 32 at a time by the lanes of the stepping warp.
// Same legality argument as the grid engine (SURVEY §8a conflict analysis): the unit of mutual exclusion is
// the (node, direction) group, so inside a chunk of 32 consecutive events a lane may run as soon as every
// EARLIER event of its group is done; events whose solve takes one of bwspread's O(1) exits run as scalar
// threads with their event_adds / records staged in shared memory (the latency of 32 dependent table walks
// overlaps instead of adding up).  An event that needs a full spread waits until everything before it has
// been committed and is then handled by the whole warp, unstaged, at its final sequence numbers.  The
// commit is an in-order prefix scan, so sequence numbers, random() burns and the record stream are those of
// the strictly sequential order.  Not used for relay topologies (a range with a producer that is also
// consumed couples groups) and in host-callback mode (every event stops).
#define NS_WIDE 32
struct WideSmem {
	double st_time[NS_STAGE_INLINE*NS_WIDE], st_speed[NS_STAGE_INLINE*NS_WIDE];
	int st_fd[NS_STAGE_INLINE*NS_WIDE];
	ns_trace_row32 st_tr[NS_WIDE];
	long long ev_ovf[NS_WIDE];
	u64 ovf_top;
	unsigned ev_nins[NS_WIDE], ev_idpos[NS_WIDE];
	uint8_t ev_nrec[NS_WIDE], ev_ntr[NS_WIDE];
	uint8_t st_rec[NS_STAGE_REC*NS_WIDE];
	ParStage ps;
};
__device__ __forceinline__ void wide_init(WideSmem &W, int lane){
	if (lane == 0) {
		ParStage &ps = W.ps;
		ps.st_time = W.st_time; ps.st_speed = W.st_speed; ps.st_fd = W.st_fd;
		ps.inline_n = NS_STAGE_INLINE*NS_WIDE; ps.ovf_cap = 0; ps.ovf_top = &W.ovf_top;
		ps.st_rec = W.st_rec; ps.st_tr = W.st_tr;
		ps.ev_nins = W.ev_nins; ps.ev_ovf = W.ev_ovf; ps.ev_nrec = W.ev_nrec; ps.ev_ntr = W.ev_ntr; ps.ev_idpos = W.ev_idpos;
		W.ovf_top = 0;
	}
	__syncwarp();
}

// in-order commit of the staged events [a, b) of the chunk (one lane per event)
__device__ __forceinline__ void wide_commit(Ctx &c, WideSmem &W, int a, int b){
	if (a >= b) return;
	const Sim &S = *c.S;
	const int lane = c.lane;
	const bool mine = lane >= a && lane < b;
	unsigned ni = mine ? W.ev_nins[lane] : 0u, nr = mine ? W.ev_nrec[lane] : 0u, nt = mine ? W.ev_ntr[lane] : 0u;
	unsigned xi = ni, xr = nr, xt = nt;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		unsigned yi = __shfl_up_sync(FULLMASK, xi, d), yr = __shfl_up_sync(FULLMASK, xr, d), yt = __shfl_up_sync(FULLMASK, xt, d);
		if (lane >= d) { xi += yi; xr += yr; xt += yt; }
	}
	const unsigned ti = __shfl_sync(FULLMASK, xi, 31), tr = __shfl_sync(FULLMASK, xr, 31), tt = __shfl_sync(FULLMASK, xt, 31);
	if (c.tail+ti-c.head > S.ring_cap) { set_status(c, NS_E_RING_FULL); return; }
	if (S.want_rec && c.rec_len+tr > S.rec_cap) { set_status(c, NS_E_RECORDS_LOST); return; }
	if (mine && (ni|nr|nt)) par_commit_event(S, W.ps, lane, c.tail+(xi-ni), c.rec_len+(xr-nr), c.tr_len+(xt-nt), 0, 1, true);
	c.tail += ti; c.burns += ti;
	if (S.want_rec) c.rec_len += tr;
	if (S.want_tr) { c.tr_len += tt; if (c.tr_len > S.tr_cap) c.tr_len = S.tr_cap; }
	__syncwarp();
}

#ifndef NS_NO_WIDE
// ring entries b_seq[i, end) of timestamp m, all SPEED_CHANGE events; advances i past what it handled
__device__ __noinline__ void seq_wide_run(Ctx &c, double m, int &i, int end, WideSmem &W){
	const Sim &S = *c.S;
	const int lane = c.lane;
	const u64 smask = S.ring_cap-1;
	while (i < end) {
		if (seq_should_stop(c)) return;
		int n = end-i < NS_WIDE ? end-i : NS_WIDE;
		{	// a max_events bound counts handled (live) events: never take more entries than the budget allows
			const u64 ev_max = c.st->max_events;
			if (ev_max) { const u64 left = ev_max-c.st->cnt[NS_CNT_EVENTS]; if ((u64)n > left) n = (int)left; }
		}
		u64 seq = 0, slot = 0; int f = -1, dir = 0, key = -1-lane; double speed = 0; bool live = false;
		if (lane < n) {
			seq = S.b_seq[i+lane]; slot = seq & smask;
			if (S.q_time[slot] == m) {
				const int fd = S.q_fd[slot];
				f = fd >> 1; dir = fd & 1; speed = S.q_speed[slot];
				live = S.f_gen[f] == S.q_gen[slot] && S.f_alive[f];
				if (live) key = S.f_peer[dir][f]*2+dir;
				else S.q_time[slot] = NS_INF;                             // cancelled by flow_close (flow.c:260-270): popped, not handled
			}
			W.ev_nins[lane] = 0; W.ev_nrec[lane] = 0; W.ev_ntr[lane] = 0; W.ev_ovf[lane] = -1;
		}
		__syncwarp();
		const unsigned peers = __match_any_sync(FULLMASK, key);
		const unsigned pred = peers & ((1u << lane)-1u);                  // earlier events of my group
		unsigned done = __ballot_sync(FULLMASK, !live);
		int committed = 0;
		u64 lc[LC_N];
		#pragma unroll
		for (int q = 0; q < LC_N; q++) lc[q] = 0;
		while (done != FULLMASK) {
			const bool ready = !((done >> lane) & 1u) && (pred & ~done) == 0u;
			const bool full = ready && sc_needs_full_spread(S, f, dir, speed);
			if (ready && !full) {
				Ctx cs = c;
				cs.lane = 0; cs.par = 1; cs.scalar = 1; cs.ps = &W.ps;
				cs.ev = (unsigned)lane; cs.lseq = 0; cs.rec_n = 0; cs.tr_n = 0; cs.ovf_base = -1;
				#pragma unroll
				for (int q = 0; q < LC_N; q++) cs.lc[q] = 0;
				cs.lc[LC_EVENTS] = 1; cs.lc[LC_SPEED_CHANGE] = 1;
				handle_speed_change(cs, f, dir, speed);
				W.ev_nins[lane] = cs.lseq; W.ev_nrec[lane] = (uint8_t)cs.rec_n; W.ev_ntr[lane] = (uint8_t)cs.tr_n;
				S.q_time[slot] = NS_INF;                                  // event_pop
				#pragma unroll
				for (int q = 0; q < LC_N; q++) lc[q] += cs.lc[q];
			}
			__syncwarp();
			const unsigned newly = __ballot_sync(FULLMASK, ready && !full);
			done |= newly;
			if (newly == 0u) {
				// no lane could run: the earliest pending event needs a full spread (everything before it is done)
				const int F = __ffs(~done)-1;
				wide_commit(c, W, committed, F);
				const u64 sF = __shfl_sync(FULLMASK, seq, F);
				if (!c.st->status) seq_handle_ring(c, m, sF);
				committed = F+1;
				done |= 1u << F;
			}
		}
		wide_commit(c, W, committed, n);
		#pragma unroll
		for (int q = 0; q < LC_N; q++) for (int d = 16; d; d >>= 1) lc[q] += __shfl_xor_sync(FULLMASK, lc[q], d);
		L0 {
			u64 *cn = c.st->cnt;
			cn[NS_CNT_EVENTS] += lc[LC_EVENTS]; cn[NS_CNT_SPEED_CHANGE] += lc[LC_SPEED_CHANGE]; cn[NS_CNT_SOLVES] += lc[LC_SOLVES];
			cn[NS_CNT_FULL_SPREADS] += lc[LC_FULL_SPREADS]; cn[NS_CNT_FLOW_VISITS] += lc[LC_FLOW_VISITS]; cn[NS_CNT_RETIMES] += lc[LC_RETIMES];
			cn[NS_CNT_PAR_EVENTS] += lc[LC_EVENTS];
		}
		__syncwarp();
		i += n;
	}
}
#endif

// fold the record buffer of an ensemble replica in place (one warp; rec_len = the bytes in the buffer): returns the new length
struct FoldSmem { uint8_t tile[ANA_TILE+32]; int starts[34]; uint8_t tailb[8]; };
__device__ __noinline__ u64 ens_fold_records(const Sim &S, SimState *st, const EnsFold &F, u64 rec_len, FoldSmem &Fs, int lane){
	const u64 base = st->rec_base;
	if (lane == 0) st->rec_len = rec_len;                             // what the analysis reads
	__syncwarp();
	__threadfence_block();
	if (F.ana) ana_process(*F.ana, Fs.tile, Fs.starts, lane, 0);
	const u64 nw = rec_len >> 3;                                      // whole words (rec_base stays a multiple of 8)
	u64 a = 0;
	const u64 *w = (const u64 *)S.rec;
	for (u64 j = lane; j < nw; j += 32) a += ns_mix64(w[j]+0x9e3779b97f4a7c15ULL*((base >> 3)+j+1));
	for (int d = 16; d; d >>= 1) a += __shfl_xor_sync(FULLMASK, a, d);
	const u64 tail = rec_len & 7;
	if ((u64)lane < tail) Fs.tailb[lane] = S.rec[(nw << 3)+lane];
	__syncwarp();
	if ((u64)lane < tail) S.rec[lane] = Fs.tailb[lane];
	if (lane == 0) { F.acc->hash += a; st->rec_base = base+(nw << 3); st->rec_len = tail; }
	__syncwarp();
	return tail;
}

// put the unprocessed tail of a batch back under the tile minima (max_events / sim_end mid-batch)
__device__ __noinline__ void seq_restore(Ctx &c, double m, int i, int nr, int j, int nl, const SeqScratch &sc){
	const Sim &S = *c.S;
	for (int k = i+c.lane; k < nr; k += 32) {
		u64 seq = S.b_seq[k];
		u64 slot = seq & (S.ring_cap-1);
		if (S.q_time[slot] == m) tmin_lower(&S.q_tmin[slot >> NS_TILE_SHIFT], m);
	}
	for (int k = j+c.lane; k < nl; k += 32) {
		int slot = sc.lb_ref[k];
		if (S.l_time[slot] == m && S.l_seq[slot] == sc.lb_seq[k]) tmin_lower(&S.l_tmin[slot >> NS_TILE_SHIFT], m);
	}
	__syncwarp();
}

// process batch entries [i,nr) x [j,nl) in sequence order, strictly sequentially; stops early on
// status / sim_end / max_events / host stop and reports how far it got
#define NS_WIDE_MIN_RUN 1      // every run: one handler path for all speed events is what many one-warp simulations per GPU need (profiles/README.md r2-o, r2-q)
__device__ __noinline__ void seq_process(Ctx &c, double m, int &i, int nr, int &j, int nl, const SeqScratch &sc, WideSmem *W){
	const Sim &S = *c.S;
#ifdef NS_NO_WIDE      // the grid engine's translation unit: its sparse phases go to k_run_seq anyway, and k_run_par's code layout is touchy (profiles/README.md r1-f)
	const bool wide_on = false;
#else
	const bool wide_on = W && (S.features & NS_FEAT_WIDE) && !c.st->host_mask;
#endif
	while (i < nr || j < nl) {
		if (seq_should_stop(c)) return;
		u64 sr = i < nr ? S.b_seq[i] : ~0ull;
		u64 sl = j < nl ? sc.lb_seq[j] : ~0ull;
		if (sr < sl) {
			if (wide_on && nr-i >= NS_WIDE_MIN_RUN && !c.st->relay_seen) {
				// the run of SPEED_CHANGE events up to the next long-lived event of the batch
				int end = nr;
				if (j < nl) { int lo = i+1, hi = nr; while (lo < hi) { int mid = (lo+hi) >> 1; if (S.b_seq[mid] < sl) lo = mid+1; else hi = mid; } end = lo; }
#ifndef NS_NO_WIDE
				if (end-i >= NS_WIDE_MIN_RUN) { seq_wide_run(c, m, i, end, *W); continue; }
#endif
			}
			seq_handle_ring(c, m, sr); i++;
		}
		else { seq_handle_l(c, m, sl, sc.lb_ref[j]); j++; }
	}
}

#define NS_SPARSE_MAX_HIT_TILES (4*1024/NS_TILE)
#define NS_ENS_MAX_HIT_TILES (64*1024/NS_TILE)
__device__ void seq_run(const Sim *Sp, const SeqScratch &sc, int lane, WideSmem *W, int ens_phase = 0, const EnsFold *fold = nullptr, FoldSmem *Fs = nullptr){
	Ctx c;
	ctx_load(c, Sp, lane);
	if (ens_phase && !c.st->sparse_mode) return;                         // ensemble, sparse phase: this replica is with the CTA engine (or done)
	if (W) wide_init(*W, lane);
	if (c.st->pending_close >= 0) {                                   // flow_done_cleaner of the event the host just saw
		int f = c.st->pending_close;
		WSYNC();
		L0 c.st->pending_close = -1;
		WSYNC();
		flow_close(c, f);
	}
#ifdef NS_PROFILE
	u64 sp_t = 0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(sp_t));
#define SEQ_MARK(id) do { u64 t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); if (lane == 0) c.st->phase_ns[id] += t_-sp_t; sp_t = t_; } while (0)
#else
#define SEQ_MARK(id) do { } while (0)
#endif
	while (true) {
		if (seq_should_stop(c)) break;
		SEQ_MARK(6);
		double m = seq_find_tmin(c);
		SEQ_MARK(0);
		if (m == NS_INF) break;                                       // queue empty (main.c:50)
		if (c.st->sparse_mode && !c.st->relay_seen) {
			// sparse phase of a big simulation: a timestamp spread over many ring tiles is a batch for
			// the grid engine — leave it in place and hand back
			const Sim &S = *c.S;
			const u64 tmask = (u64)S.n_ring_tiles-1;
			int hits = 0;
			u64 first = c.head >> NS_TILE_SHIFT, last = c.tail ? (c.tail-1) >> NS_TILE_SHIFT : 0;
			if (c.tail > c.head) for (u64 a = first+lane; a <= last; a += 32) hits += S.q_tmin[a & tmask] == m;
			for (int d = 16; d; d >>= 1) hits += __shfl_xor_sync(FULLMASK, hits, d);
			// (an ensemble switches engines in rounds over all replicas: there the warp keeps medium batches — its lanes take 32 events at a time)
			if (hits > (ens_phase ? NS_ENS_MAX_HIT_TILES : NS_SPARSE_MAX_HIT_TILES)) { L0 { c.st->want_dense = 1; if (ens_phase) c.st->sparse_mode = 0; } __syncwarp(); break; }
		}
		SEQ_MARK(1);
		int nr = seq_collect_ring(c, m);
		if (nr < 0) { set_status(c, NS_E_BATCH_FULL); __syncwarp(); break; }
		SEQ_MARK(2);
		int nl = seq_collect_l(c, m, sc);
		SEQ_MARK(3);
		if (nr+nl == 0) continue;                                     // only stale lower bounds surfaced
		u64 n_ops = 0;
		for (int k = lane; k < nl; k += 32) n_ops += (u64)l_entry_ops(*c.S, sc.lb_ref[k]);
		for (int d = 16; d; d >>= 1) n_ops += __shfl_xor_sync(FULLMASK, n_ops, d);
		if (rec_room_short(c, nr+nl, n_ops)) {                        // soft stop BEFORE the batch: nothing is lost
			seq_restore(c, m, 0, nr, 0, nl, sc);
			if (fold && c.rec_len > 64) {                                 // an ensemble replica folds its records itself and goes on
				c.rec_len = ens_fold_records(*c.S, c.st, *fold, c.rec_len, *Fs, lane);
				continue;
			}
			L0 c.st->host_stop = 2;
			__syncwarp();
			break;
		}
		c.now = m;                                                    // main.c:60
		cnt_add(c, NS_CNT_BATCHES, 1);
		L0 { if ((u64)(nr+nl) > c.st->cnt[NS_CNT_MAX_BATCH]) c.st->cnt[NS_CNT_MAX_BATCH] = nr+nl; }
		int pi = 0, pj = 0;
		SEQ_MARK(4);
		seq_process(c, m, pi, nr, pj, nl, sc, W);
		SEQ_MARK(5);
		if (pi < nr || pj < nl) seq_restore(c, m, pi, nr, pj, nl, sc);   // max_events / sim_end mid-batch
		__syncwarp();
	}
	L0 c.st->cnt[NS_CNT_SEQ] = c.tail;
	ctx_flush(c);
}

}   // namespace
