// ns_par.cuh — the grid-parallel engine: one persistent cooperative kernel steps the simulation,
// one equal-timestamp batch per iteration, all SMs working on the same batch.
//
//   A  argmin            min over the per-tile minima of both event stores        (event_pop, event.h:20-25)
//   B  ordered extract   count / scan / write of the ring entries with time == t  (FIFO ties, event.h:27-31)
//   C  pieces            the batch is cut, in sequence-number order, into pieces (par_plan_piece): long
//                        runs of SPEED_CHANGE events of a relay-free simulation go through D-F, everything
//                        else (FLOW_DONE / USER / THROTTLE events, short runs, relay topologies, the last
//                        events before a max_events bound) through the single-warp code of ns_seq.cuh
//   D  grouping          events are bucketed by the (node, direction) whose rates they rewrite
//                        (bwspread, flow.c:56-60): that is the unit of mutual exclusion;
//                        different groups touch disjoint mutable state, same-group events keep FIFO order
//   E1 solve, threads    one THREAD per group for the events whose bwspread takes an O(1) exit
//                        (sc_needs_full_spread decides before anything is touched); a group that meets a
//                        full spread is handed over at that position
//   E2 solve, warps      one WARP per handed-over group: full-spread loops over the lanes, the reference's
//                        in-order fp64 sums reproduced from shuffled terms; nodes with <= 32 flows run their
//                        whole event chain register-resident (ge_chain); emissions staged under local indices
//   F  commit            prefix scan over the batch order of {event_adds, record bytes, trace rows}; staged
//                        speed events are copied into the ring at their final sequence numbers, records
//                        into the record stream — byte-identical to the sequential order
//
// Reads of state owned by another group (get_share(nf,!dir), flow.c:140,165: the remote node's
// total_bwupbound / maximum_bandwidth) are stable inside such a segment because only flow
// create/close write them (flow.c:246-247,314-315), and those are never part of it.
//
// Around it: k_init_par runs a module init made of flow creations in conflict-free waves; par_run (host)
// alternates this kernel with the single-warp k_run_seq when the simulation enters a sparse phase.
#pragma once
#include <cooperative_groups.h>
#include "ns_seq.cuh"
#include "ns_par_api.cuh"

namespace cgp = cooperative_groups;

namespace {   // internal linkage: every engine translation unit carries its own copy of the device code

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ double block_min_d(double v, double *sh){
	v = warp_min_d(v);
	int w = threadIdx.x >> 5, l = threadIdx.x & 31;
	__syncthreads();
	if (l == 0) sh[w] = v;
	__syncthreads();
	v = l < PAR_WARPS ? sh[l] : NS_INF;
	return warp_min_d(v);
}

// ascending in-place sort of a[0..n) by one warp (bitonic network with mirrored first stage, so a
// virtual +inf padding needs no storage)
__device__ __noinline__ void warp_sort_global(int *a, int n, int lane){
	int np = 1; while (np < n) np <<= 1;
	for (int k = 2; k <= np; k <<= 1) {
		for (int i = lane; i < np; i += 32) {
			int p = i ^ (k-1);
			if (p > i && p < n) { int x = a[i], y = a[p]; if (x > y) { a[i] = y; a[p] = x; } }
		}
		__syncwarp();
		for (int j = k >> 2; j > 0; j >>= 1) {
			for (int i = lane; i < np; i += 32) {
				int p = i ^ j;
				if (p > i && p < n) { int x = a[i], y = a[p]; if (x > y) { a[i] = y; a[p] = x; } }
			}
			__syncwarp();
		}
	}
}

// one staged SPEED_CHANGE event handled by a warp
__device__ void par_handle_event(Ctx &c, const ParScratch &P, int seg_lo, int e, double m){
	const Sim &S = *c.S;
	u64 seq = S.b_seq[seg_lo+e];
	u64 slot = seq & (S.ring_cap-1);
	int fd = S.q_fd[slot];
	double speed = S.q_speed[slot];
	c.ev = (unsigned)e; c.lseq = 0; c.rec_n = 0; c.tr_n = 0; c.ovf_base = -1;
	cnt_add(c, NS_CNT_EVENTS, 1);
	cnt_add(c, NS_CNT_SPEED_CHANGE, 1);
	handle_speed_change(c, fd >> 1, fd & 1, speed);
	L0 {
		const ParStage &ps = *c.ps;
		ps.ev_nins[c.ev] = c.lseq; ps.ev_ovf[c.ev] = c.ovf_base;
		ps.ev_nrec[c.ev] = (uint8_t)c.rec_n; ps.ev_ntr[c.ev] = (uint8_t)c.tr_n;
		S.q_time[slot] = NS_INF;                                      // event_pop
	}
	WSYNC();
}

// a chain of events of one small (node,dir) group with the node's flows held in registers (ns_model.cuh)
#define GE_MIN_CHAIN 6
__device__ void ge_chain(Ctx &c, const ParScratch &P, int g, int seg_lo, int off, int j, int cnt, double m){
	const Sim &S = *c.S;
	const int lane = c.lane, n = g >> 1, dir = g & 1;
	const u64 smask = S.ring_cap-1;
	const int a_off = S.adj_off[g], a_len = S.adj_len[g];
	const double total = S.n_total[dir][n], mx = S.n_max[dir][n];
	double usage = S.n_usage[dir][n];
	const unsigned node_id = S.n_id[n];
	GFlow gf;
	ge_load(S, gf, a_off, a_len, lane, dir, total, mx);
	GeStage stg;
	ge_stage_load(stg, S, *c.ps);
	unsigned n_simple = 0, n_ret = 0;                                 // this lane's O(1) events / their re-times
	for (int j0 = j; j0 < cnt; j0 += 32) {
		// fetch up to 32 events of the chain at once
		int my_e = -1, my_fd = 0; double my_speed = 0; u64 my_slot = 0;
		if (j0+lane < cnt) {
			my_e = P.grp_items[off+j0+lane];
			my_slot = S.b_seq[seg_lo+my_e] & smask;
			my_fd = S.q_fd[my_slot]; my_speed = S.q_speed[my_slot];
		}
		int nev = cnt-j0 < 32 ? cnt-j0 : 32;
		for (int k = 0; k < nev; k++) {
			int e = __shfl_sync(FULLMASK, my_e, k);
			int fd = __shfl_sync(FULLMASK, my_fd, k);
			double speed = __shfl_sync(FULLMASK, my_speed, k);
			u64 slot = __shfl_sync(FULLMASK, my_slot, k);
			int f = fd >> 1;
			unsigned own = __ballot_sync(FULLMASK, gf.nf == f);
			int owner = __ffs(own)-1;
			bool full = true;
			if (owner >= 0) {
				// same decision sequence as sc_needs_full_spread(), on the cached values
				double fs = __shfl_sync(FULLMASK, gf.sp, owner), share = __shfl_sync(FULLMASK, gf.share, owner);
				double amount = speed-fs;
				full = true;
				if (amount > NS_EPS && fs+amount > share) { if (fs > share) full = false; else amount = share-fs; }
				if (full) {
					if (amount > -NS_EPS && amount < NS_EPS) full = false;
					else if (amount < NS_EPS && amount > -64) full = false;
					else if (total < mx+NS_EPS) full = false;
					else if (amount > -NS_EPS && usage+amount < mx+NS_EPS) full = false;
				}
			}
			if (full && owner < 0) {
				// a flow this lane set does not hold (cannot happen for a consistent group): table-driven
				// handler; the cache is write-through, so the tables are current; reload afterwards
				__syncwarp();
				par_handle_event(c, P, seg_lo, e, m);
				ge_load(S, gf, a_off, a_len, lane, dir, total, mx);
				usage = S.n_usage[dir][n];
				continue;
			}
			if (full) {
				c.ev = (unsigned)e; c.lseq = 0; c.rec_n = 0; c.tr_n = 0; c.ovf_base = -1;
				ge_full_event(c, S, gf, usage, total, mx, node_id, n, dir, speed, owner);
				if (lane == 0) {
					const ParStage &ps = *c.ps;
					ps.ev_nins[e] = c.lseq; ps.ev_ovf[e] = c.ovf_base;
					ps.ev_nrec[e] = (uint8_t)c.rec_n; ps.ev_ntr[e] = (uint8_t)c.tr_n;
					S.q_time[slot] = NS_INF;                          // event_pop
				}
				continue;
			}
			if (lane == owner) {
				ge_simple_event(stg, S, c.st, gf, usage, node_id, n, dir, speed, m, (unsigned)e, n_ret);
				n_simple++;
				S.q_time[slot] = NS_INF;                              // event_pop
			}
			usage = __shfl_sync(FULLMASK, usage, owner);
		}
		__syncwarp();
	}
	// the lanes' private counters
	for (int d = 16; d; d >>= 1) { n_simple += __shfl_xor_sync(FULLMASK, n_simple, d); n_ret += __shfl_xor_sync(FULLMASK, n_ret, d); }
	c.lc[LC_EVENTS] += n_simple; c.lc[LC_SPEED_CHANGE] += n_simple; c.lc[LC_SOLVES] += n_simple; c.lc[LC_RETIMES] += n_ret;
	__syncwarp();
}

// ---------------------------------------------------------------- stable partition of a segment by group
// One LSD radix pass over ns (key, value) pairs by the whole grid: every CTA owns a contiguous chunk of the
// current order, counts its digits, all CTAs derive their per-digit bases from the (digit, CTA) table, and the
// chunk is scattered tile by tile in order (rank inside a warp by match_any, warps in order through shared
// memory).  Stable, deterministic, no sorting by anyone afterwards.  The first pass takes value = position.
#define RAD_BITS 9
#define RAD_BINS (1 << RAD_BITS)
struct RadSmem { int hist[RAD_BINS], base[RAD_BINS], wcnt[PAR_WARPS][RAD_BINS], scan_tmp[PAR_WARPS]; };
static_assert(RAD_BINS == 2*PAR_THREADS, "the bin scan gives every thread two bins");
// scope of the engine's barriers: the whole cooperative grid (one simulation per GPU) or one CTA (one replica per CTA)
#ifndef NS_SCOPE_CTA
#define NS_GRID_T cgp::grid_group
#define NS_SYNC(g) (g).sync()
#define NS_NB ((int)gridDim.x)
#define NS_BID ((int)blockIdx.x)
#else
struct CtaScope { };
#define NS_GRID_T CtaScope
#define NS_SYNC(g) __syncthreads()
#define NS_NB 1
#define NS_BID 0
#endif
__device__ __noinline__ void radix_pass(NS_GRID_T &grid, RadSmem &R, const int *src_key, const int *src_val, int *dst_key, int *dst_val,
		int ns, int shift, int *ghist){
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int G = NS_NB, b = NS_BID;
	const int per = (ns+G-1)/G;
	const int c_lo = b*per < ns ? b*per : ns, c_hi = c_lo+per < ns ? c_lo+per : ns;
	for (int i = tid; i < RAD_BINS; i += PAR_THREADS) R.hist[i] = 0;
	for (int i = tid; i < PAR_WARPS*RAD_BINS; i += PAR_THREADS) (&R.wcnt[0][0])[i] = 0;
	__syncthreads();
	for (int i = c_lo+tid; i < c_hi; i += PAR_THREADS) atomicAdd(&R.hist[(src_key[i] >> shift) & (RAD_BINS-1)], 1);
	__syncthreads();
	for (int i = tid; i < RAD_BINS; i += PAR_THREADS) ghist[i*G+b] = R.hist[i];
	NS_SYNC(grid);
	for (int bin = tid; bin < RAD_BINS; bin += PAR_THREADS) {
		int tot = 0, before = 0;
		const int *row = ghist+bin*G;
		for (int q = 0; q < G; q++) { int v = row[q]; if (q < b) before += v; tot += v; }
		R.hist[bin] = tot; R.base[bin] = before;
	}
	__syncthreads();
	{	// exclusive scan of the digit totals
		const int a0 = R.hist[2*tid], a1 = R.hist[2*tid+1];
		const int s = a0+a1;
		int incl = s;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULLMASK, incl, d); if (lane >= d) incl += t; }
		if (lane == 31) R.scan_tmp[warp] = incl;
		__syncthreads();
		int ex = incl-s;
		for (int w = 0; w < warp; w++) ex += R.scan_tmp[w];
		R.base[2*tid] += ex; R.base[2*tid+1] += ex+a0;
	}
	__syncthreads();
	for (int t0 = c_lo; t0 < c_hi; t0 += PAR_THREADS) {
		const int i = t0+tid;
		const bool valid = i < c_hi;
		const int k = valid ? src_key[i] : 0, v = valid ? (src_val ? src_val[i] : i) : 0;
		const int d = valid ? ((k >> shift) & (RAD_BINS-1)) : RAD_BINS+lane;
		const unsigned peers = __match_any_sync(FULLMASK, d);
		const int rank = __popc(peers & ((1u << lane)-1));
		const bool leader = valid && rank == 0;
		if (leader) R.wcnt[warp][d] = __popc(peers);
		__syncthreads();
		if (valid) {
			int o = R.base[d];
			for (int w = 0; w < warp; w++) o += R.wcnt[w][d];
			dst_key[o+rank] = k; dst_val[o+rank] = v;
		}
		__syncthreads();
		if (leader) { atomicAdd(&R.base[d], __popc(peers)); R.wcnt[warp][d] = 0; }
		__syncthreads();
	}
	NS_SYNC(grid);
}

// ---------------------------------------------------------------- chains, second form
// The same chain split the way its data dependences run.  Inside one (node,dir) group at one `now`
// the only loop-carried state of the O(1) solves (flow.c:64-124) is the speed of each flow and the
// node's usage; everything an event EMITS — its records (record.c:53-116), its event_adds, the
// completion re-time of range.c:16-79 — depends on that state only through the event's own
// (speed after, usage after, branch taken).  So, 32 events at a time:
//   pass A  walks the events in order carrying (gf.sp per lane, usage) and leaves each event's
//           {speed after, usage after, branch flags} in the lane of the same index;
//   advance the byte progress of every touched flow once, with the speed it had before its first event of
//           the run (flow.c:355-357 -> range.h:91-125 is a no-op for later events at the same `now`);
//   pass B  gives each lane ONE event: owner constants come by indexed shuffle, records are built as words,
//           event_adds staged, the L-table entry written by the last event of each flow.
// A full spread in the middle ends the run: pass B for what came before, ge_full_event, next run.
#define GE_EMIT 1u        // flow.c:64-78: the flow is at its share, tell the other end
#define GE_DONE_BW 2u     // ... and bwspread returned right there (flow.c:66-74)
#define GE_CHANGED 4u     // speed and usage changed: R_USAGE record (flow.c:85-124)
#define GE_DELTA_POS 8u   // flow.c:371: delta > 0

// one 19-byte record with a 6-byte fixed-point value (record.c:53-116) as little-endian words; h0/h1 carry
// the time fields, which are the same for every record of a chain
__device__ __forceinline__ uint32_t bswap32(uint32_t x){ return __byte_perm(x, 0, 0x0123); }
__device__ __forceinline__ void rec19_time(double now, u64 &h0, u64 &h1){
	uint32_t t = (uint32_t)(int)now;                                  // record.c:66
	uint16_t ut = (uint16_t)(int)((now-(double)((int)now))*1000);     // record.c:69
	uint32_t bt = bswap32(t);
	h0 = (u64)(bt & 0xffffu) << 48;                                   // bytes 6,7
	h1 = (u64)(bt >> 16) | ((u64)(((ut & 0xffu) << 8) | (ut >> 8)) << 16) | (0xffull << 32);   // bytes 8..12 (bytes = -1)
}
__device__ __forceinline__ void rec19_words(u64 h0, u64 h1, uint8_t minor, uint32_t id, double dval, u64 &w0, u64 &w1, u64 &w2){
	double tmp = dval;                                                // record.c:98-106
	uint32_t ip = (uint32_t)(int)tmp;
	tmp = tmp-(double)((int)tmp);
	uint16_t fp = (uint16_t)(int)(tmp*1000);
	uint32_t bi = bswap32(ip);
	w0 = ((u64)minor << 8) | ((u64)bswap32(id) << 16) | h0;
	w1 = h1 | ((u64)(bi & 0xffffffu) << 40);
	w2 = (u64)(bi >> 24) | ((u64)(((fp & 0xffu) << 8) | (fp >> 8)) << 8);
}

__device__ void ge_chain2(Ctx &c, const ParScratch &P, int g, int seg_lo, int off, int j, int cnt, double m){
	const Sim &S = *c.S;
	SimState *st = c.st;
	const int lane = c.lane, n = g >> 1, dir = g & 1;
	const int a_off = S.adj_off[g], a_len = S.adj_len[g];
	const double total = S.n_total[dir][n], mx = S.n_max[dir][n];
	double usage = S.n_usage[dir][n];
	const unsigned node_id = S.n_id[n];
	const double now = m;
	const bool sat = !(total < mx+NS_EPS);                            // flow.c:94: an unsaturated node never spreads
	GFlow gf;
	ge_load(S, gf, a_off, a_len, lane, dir, total, mx);
	GeStage stg;
	ge_stage_load(stg, S, *c.ps);
	u64 h0, h1;
	rec19_time(now, h0, h1);
	unsigned n_simple = 0;
	// the events of the chain come packed in partition order (k_run_par phase D3): one coalesced load per field,
	// and the next block is fetched while this one runs
	int nx_e = -1, nx_fd = 0, nx_own = lane; double nx_speed = 0; u64 nx_slot = 0;
	if (j+lane < cnt) { const int q = off+j+lane; nx_e = P.grp_items[q]; nx_fd = P.pk_fd[q]; nx_speed = P.pk_speed[q]; nx_slot = P.pk_slot[q]; nx_own = P.pk_own[q]; }
	for (int j0 = j; j0 < cnt; j0 += 32) {
		const int nev = cnt-j0 < 32 ? cnt-j0 : 32;
		// ---- fetch: lane k holds event k of the block and the lane that owns its flow (list position == lane, ge_load)
		int my_e = nx_e, my_fd = nx_fd, own = nx_own; double my_speed = nx_speed; u64 my_slot = nx_slot;
		nx_e = -1; nx_own = lane;
		if (j0+32+lane < cnt) { const int q = off+j0+32+lane; nx_e = P.grp_items[q]; nx_fd = P.pk_fd[q]; nx_speed = P.pk_speed[q]; nx_slot = P.pk_slot[q]; nx_own = P.pk_own[q]; }
		bool bad = false;
		if (lane < nev && (unsigned)own >= 32u) { bad = true; own = lane; }
		{
			const int own_nf = __shfl_sync(FULLMASK, gf.nf, own);
			if (lane < nev && own_nf != (my_fd >> 1)) bad = true;
		}
		if (__any_sync(FULLMASK, bad)) {
			// a flow the lanes do not hold (cannot happen for a consistent group): the table-driven handler for
			// this block; the tables are current (write-back below), reload afterwards
			for (int k = 0; k < nev; k++) par_handle_event(c, P, seg_lo, __shfl_sync(FULLMASK, my_e, k), m);
			ge_load(S, gf, a_off, a_len, lane, dir, total, mx);
			usage = S.n_usage[dir][n];
			continue;
		}
		unsigned dirty = 0;                                           // lanes whose gf.sp moved in this block
		int k0 = 0;
		while (k0 < nev) {
			// ---- pass A: the recurrence, warp-uniform
			const double sp0 = gf.sp;
			double ev_fs = 0, ev_use = 0; unsigned ev_fl = 0;
			bool hit_full = false;
			int k = k0;
			int owner_n = __shfl_sync(FULLMASK, own, k0 & 31);
			double speed_n = __shfl_sync(FULLMASK, my_speed, k0 & 31), share_n = __shfl_sync(FULLMASK, gf.share, owner_n);
			for (; k < nev; k++) {
				// what does not depend on the carried state (the event, its owner, the owner's share) is fetched one
				// iteration ahead, so only the speed shuffle and the arithmetic sit on the chain
				const int owner = owner_n;
				const double speed = speed_n, share = share_n;
				owner_n = __shfl_sync(FULLMASK, own, (k+1) & 31);
				speed_n = __shfl_sync(FULLMASK, my_speed, (k+1) & 31);
				share_n = __shfl_sync(FULLMASK, gf.share, owner_n);
				const double fs = __shfl_sync(FULLMASK, gf.sp, owner);
				double amount = speed-fs;                                 // flow.c:359
				unsigned fl = amount > 0 ? GE_DELTA_POS : 0u;
				if (amount > NS_EPS && fs+amount > share) {               // flow.c:64-78
					fl |= GE_EMIT;
					if (fs > share) fl |= GE_DONE_BW; else amount = share-fs;
				}
				double nfs = fs;
				if (!(fl & GE_DONE_BW) && !(amount > -NS_EPS && amount < NS_EPS)) {   // flow.c:80-83
					if (sat && !(amount < NS_EPS && amount > -64) && !(amount > -NS_EPS && usage+amount < mx+NS_EPS)) { hit_full = true; break; }
					nfs = fs+amount;                                      // flow.c:85
					usage += amount;                                      // flow.c:86-124, the O(1) exits
					fl |= GE_CHANGED;
					if (lane == owner) gf.sp = nfs;
					dirty |= 1u << owner;
				}
				if (lane == k) { ev_fs = nfs; ev_use = usage; ev_fl = fl; }
			}
			const int k1 = k;
			const bool act = lane >= k0 && lane < k1;
			if (k1 > k0) {
				n_simple += (unsigned)(k1-k0);
				// ---- byte progress of the touched flows, once, with the speed before the run (range.h:91-125)
				if (dir == RCV) {
					const unsigned touched = __reduce_or_sync(FULLMASK, act ? (1u << own) : 0u);
					if ((touched >> lane) & 1u) {
						if (now != gf.s_last) { gf.s_last = now; S.r_last[gf.sr] = now; }
						if (now != gf.d_last) {
							double delta = sp0*(now-gf.d_last), tt;
							delta = delta-gf.d_lenc;
							tt = delta+gf.d_len;
							gf.d_lenc = (tt-gf.d_len)-delta;
							gf.d_len = tt;
							gf.d_last = now;
							S.r_lenc[gf.dr] = gf.d_lenc; S.r_len[gf.dr] = gf.d_len; S.r_last[gf.dr] = now;
						}
					}
				}
				// ---- pass B: one event per lane; the owner's constants by indexed shuffle
				const unsigned o_id = __shfl_sync(FULLMASK, gf.id, own);
				const int o_ord = __shfl_sync(FULLMASK, gf.ord, own);
				const double o_delay = __shfl_sync(FULLMASK, gf.delay, own), o_share = __shfl_sync(FULLMASK, gf.share, own);
				double o_dlen = 0, o_slen = 0; long long o_dstart = 0, o_sstart = 0; int o_npos = 0;
				if (dir == RCV) {
					o_dlen = __shfl_sync(FULLMASK, gf.d_len, own); o_slen = __shfl_sync(FULLMASK, gf.s_len, own);
					o_dstart = __shfl_sync(FULLMASK, gf.d_start, own); o_sstart = __shfl_sync(FULLMASK, gf.s_start, own);
					o_npos = __shfl_sync(FULLMASK, gf.npos, own);
				}
				// the last event of each flow in the run owns the flow's L-table entries
				const unsigned peers = __match_any_sync(FULLMASK, act ? own : 32+lane);
				const bool last = act && lane == 31-__clz(peers);
				if (act) {
					const int f = my_fd >> 1;
					const unsigned e = (unsigned)my_e;
					const u64 sbase = (u64)e*NS_STAGE_INLINE;
					unsigned lseq = 0;
					const double et = now+o_delay;                        // queue_speed_event, flow.c:22
					if (!(et >= now)) atomicCAS(&st->status, 0, isnan(et) ? NS_E_NAN_TIME : NS_E_BACK_IN_TIME);
					if (ev_fl & GE_EMIT) {
						stg.st_time[sbase+lseq] = et; stg.st_speed[sbase+lseq] = (ev_fl & GE_DONE_BW) ? ev_fs : o_share;
						stg.st_fd[sbase+lseq] = f*2+(!dir); lseq++;
					}
					if (dir == RCV) {                                     // flow.c:367-370 -> range.c:16-79
						Retime p;
						p.add_drain = p.add_done = 0; p.drain_t = p.done_t = NS_INF;
						if (!(ev_fs < NS_EPS)) {                          // range.c:26-27
							int dtype = -1;
							int err = retime_math(now, ev_fs, false, 0.0, NS_INF, o_slen, o_dlen, o_dstart, o_sstart, o_npos, p, dtype);
							if (dtype >= 0) S.f_drain_type[f] = (uint8_t)dtype;
							if (err) atomicCAS(&st->status, 0, err);
						}
						if (p.add_done) {
							if (p.add_drain) { stg.st_fd[sbase+lseq] = NS_HOLE_PURE; lseq++; }
							if (last) {
								S.l_time[2*f] = p.done_t; S.l_seq[2*f] = NS_SEQ_PROV | ((u64)e << 24) | lseq;
								S.l_time[2*f+1] = NS_INF;
								tmin_lower(&S.l_tmin[(2*f) >> NS_TILE_SHIFT], p.done_t);
							}
							stg.st_fd[sbase+lseq] = -(2*f+1); lseq++;
						} else if (p.add_drain) {
							if (last) {
								S.l_time[2*f+1] = p.drain_t; S.l_seq[2*f+1] = NS_SEQ_PROV | ((u64)e << 24) | lseq;
								S.l_time[2*f] = NS_INF;
								tmin_lower(&S.l_tmin[(2*f+1) >> NS_TILE_SHIFT], p.drain_t);
							}
							stg.st_fd[sbase+lseq] = -(2*f+1+1); lseq++;
						} else if (last) { S.l_time[2*f] = NS_INF; S.l_time[2*f+1] = NS_INF; }
					} else if (ev_fl & GE_DELTA_POS) {                    // flow.c:371-374
						stg.st_time[sbase+lseq] = et; stg.st_speed[sbase+lseq] = ev_fs; stg.st_fd[sbase+lseq] = f*2+RCV; lseq++;
					}
					unsigned rec_n = 0;
					if (stg.want_rec) {
						u64 *rw = (u64 *)(stg.st_rec+(u64)e*NS_STAGE_REC);
						u64 a0, a1, a2;
						rec19_words(h0, h1, R_SPD|dir, o_id, my_speed, a0, a1, a2);      // flow.c:383
						if (ev_fl & GE_CHANGED) {
							u64 u0, u1, u2;
							rec19_words(h0, h1, R_USAGE|dir, node_id, ev_use, u0, u1, u2);   // record_wrapper.h:6-9
							rw[0] = u0; rw[1] = u1; rw[2] = u2 | (a0 << 24);
							rw[3] = (a0 >> 40) | (a1 << 24); rw[4] = (a1 >> 40) | (a2 << 24);
							rec_n = 38;
						} else { rw[0] = a0; rw[1] = a1; rw[2] = a2; rec_n = 19; }
					}
					unsigned ntr = 0;
					if (stg.want_tr) {
						ns_trace_row32 r; r.type = EV_SPEED_CHANGE|(dir << 8); r.ord = o_ord; r.time = now; r.v0 = my_speed; r.v1 = ev_fs;
						stg.st_tr[e] = r; ntr = 1;
					}
					stg.ev_nins[e] = lseq; stg.ev_ovf[e] = -1; stg.ev_nrec[e] = (uint8_t)rec_n; stg.ev_ntr[e] = (uint8_t)ntr;
					S.q_time[my_slot] = NS_INF;                           // event_pop
				}
			}
			if (!hit_full) break;
			// ---- event k1 reaches the full spread (flow.c:126-203): warp-cooperative, on the registers
			{
				const int e = __shfl_sync(FULLMASK, my_e, k1), owner = __shfl_sync(FULLMASK, own, k1);
				const double speed = __shfl_sync(FULLMASK, my_speed, k1);
				const u64 slot = __shfl_sync(FULLMASK, my_slot, k1);
				c.ev = (unsigned)e; c.lseq = 0; c.rec_n = 0; c.tr_n = 0; c.ovf_base = -1;
				ge_full_event(c, S, gf, usage, total, mx, node_id, n, dir, speed, owner);
				if (lane == 0) {
					const ParStage &ps = *c.ps;
					ps.ev_nins[e] = c.lseq; ps.ev_ovf[e] = c.ovf_base;
					ps.ev_nrec[e] = (uint8_t)c.rec_n; ps.ev_ntr[e] = (uint8_t)c.tr_n;
					S.q_time[slot] = NS_INF;                              // event_pop
				}
			}
			k0 = k1+1;
		}
		// ---- write-back: the speeds that moved, the usage
		if ((dirty >> lane) & 1u) S.f_speed[dir][gf.nf] = gf.sp;
		if (lane == 0) S.n_usage[dir][n] = usage;
		__syncwarp();
	}
	c.lc[LC_EVENTS] += n_simple; c.lc[LC_SPEED_CHANGE] += n_simple; c.lc[LC_SOLVES] += n_simple;
	if (dir == RCV) c.lc[LC_RETIMES] += n_simple;
	__syncwarp();
}

// ---------------------------------------------------------------- hub chains: one CTA, shared memory
// A saturated sender with hundreds or thousands of flows (a p2p server / seeder) answers nearly every event
// with a full spread (flow.c:126-203), and a batch holds a chain of them for the same node.  One CTA takes
// the whole chain: the node's flows {speed, share, remote share, delay} are loaded into shared memory once
// (list position p, newest = highest, flow.c:311-312), every spread runs its two passes over them with all
// threads, and the reference's two in-order fp64 sums (e, new_use; flow.c:128-145,199) are taken by one thread
// over the order-preserving compaction of the NON-ZERO terms (x + 0.0 == x, and the sums never reach -0.0, so
// dropping exact zeros leaves every partial sum unchanged).  SND groups only: RCV victims re-time (flow.c:193).
#define HUB_CAP 3584
#define HUB_MIN_DEG 160
#define HUB_ROWS (HUB_CAP/PAR_THREADS)
#define HUB_SORT_MAX 4096
struct HubSmem {
	double sp[HUB_CAP], shr[HUB_CAP], rsh[HUB_CAP], dly[HUB_CAP], term[HUB_CAP];
	int nf[HUB_CAP];
	int cnt_a[HUB_ROWS*PAR_WARPS], cnt_b[HUB_ROWS*PAR_WARPS];     // per (row, warp): selected entries -> exclusive offsets
	int ev_e[PAR_THREADS], ev_pf[PAR_THREADS]; double ev_speed[PAR_THREADS]; u64 ev_slot[PAR_THREADS];
	double bcast_d; long long bcast_ll; int tot_a, tot_b, work;
};
// exclusive scan of the (row, warp) counts in row-major order by warp 0; returns nothing, leaves the total in *tot
__device__ __forceinline__ void hub_scan(int *cnt, int n, int *tot, int lane){
	const int per = (HUB_ROWS*PAR_WARPS+31)/32;
	int v[per], s = 0;
	#pragma unroll
	for (int q = 0; q < per; q++) { int i = lane*per+q; v[q] = i < n ? cnt[i] : 0; s += v[q]; }
	int incl = s;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULLMASK, incl, d); if (lane >= d) incl += t; }
	int run = incl-s;
	#pragma unroll
	for (int q = 0; q < per; q++) { int i = lane*per+q; if (i < n) cnt[i] = run; run += v[q]; }
	if (lane == 31) *tot = incl;
}
__device__ __forceinline__ double hub_serial_sum(const double *t, int n){
	double a = 0;
	int i = 0;
	for (; i+8 <= n; i += 8) {
		double x0 = t[i], x1 = t[i+1], x2 = t[i+2], x3 = t[i+3], x4 = t[i+4], x5 = t[i+5], x6 = t[i+6], x7 = t[i+7];
		a += x0; a += x1; a += x2; a += x3; a += x4; a += x5; a += x6; a += x7;
	}
	for (; i < n; i++) a += t[i];
	return a;
}
__device__ void hub_chain_cta(const Sim &S, SimState *st, const ParScratch &P, HubSmem &H, int g, int seg_lo, double m, u64 *lc){
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int n = g >> 1;                                             // dir == SND
	const ParStage &ps = P.stage;
	const int off = P.grp_off[g], cnt = P.grp_cnt[g];
	int j = P.defer_pos[g];
	const int a_off = S.adj_off[g], len = S.adj_len[g], live = S.adj_live[g];
	const double total = S.n_total[SND][n], mx = S.n_max[SND][n];
	double usage = S.n_usage[SND][n];
	const unsigned node_id = S.n_id[n];
	const bool sat = !(total < mx+NS_EPS);
	const int rows = (len+PAR_THREADS-1)/PAR_THREADS;
	u64 h0, h1;
	rec19_time(m, h0, h1);
	__syncthreads();
	// ---- the chain in batch order
	if (j < 0) {
		int *items = P.grp_items+off;
		if (cnt <= HUB_SORT_MAX) {
			int *a = (int *)H.term;
			int np = 1; while (np < cnt) np <<= 1;
			for (int i = tid; i < np; i += PAR_THREADS) a[i] = i < cnt ? items[i] : 0x7fffffff;
			__syncthreads();
			for (int k = 2; k <= np; k <<= 1)
				for (int jj = k >> 1; jj > 0; jj >>= 1) {
					for (int i = tid; i < np; i += PAR_THREADS) {
						int ixj = i ^ jj;
						if (ixj > i) { int x = a[i], y = a[ixj]; bool asc = (i & k) == 0; if ((x > y) == asc) { a[i] = y; a[ixj] = x; } }
					}
					__syncthreads();
				}
			for (int i = tid; i < cnt; i += PAR_THREADS) items[i] = a[i];
		} else if (warp == 0) warp_sort_global(items, cnt, lane);
		j = 0;
		__syncthreads();
	}
	// ---- the node's flows, once
	for (int p = tid; p < len; p += PAR_THREADS) {
		int nf = S.adj_arena[a_off+p];
		H.nf[p] = nf;
		if (nf >= 0) {
			double ub = S.f_ub[nf];
			H.sp[p] = S.f_speed[SND][nf];
			H.shr[p] = total > mx ? ub*mx/total : ub;                     // flow.c:61; == ub*mx/total inside a spread (total > max there)
			H.rsh[p] = get_share(S, nf, RCV);                             // flow.c:140,165
			H.dly[p] = S.f_delay[nf];
		}
	}
	for (int j0 = j; j0 < cnt; j0 += PAR_THREADS) {
		const int nev = cnt-j0 < PAR_THREADS ? cnt-j0 : PAR_THREADS;
		__syncthreads();
		if (tid < nev) {
			const int q = off+j0+tid;
			const int f = P.pk_fd[q] >> 1;
			int pf = P.pk_own[q];
			if ((unsigned)pf >= (unsigned)len || H.nf[pf] != f) pf = -1;
			H.ev_e[tid] = P.grp_items[q]; H.ev_pf[tid] = pf; H.ev_speed[tid] = P.pk_speed[q]; H.ev_slot[tid] = P.pk_slot[q];
		}
		__syncthreads();
		for (int i = 0; i < nev; i++) {
			const int e = H.ev_e[i], pf = H.ev_pf[i];
			const double speed = H.ev_speed[i];
			if (pf < 0) { if (tid == 0) atomicCAS(&st->status, 0, NS_E_INTERNAL); continue; }
			const int f = H.nf[pf];
			double fs = H.sp[pf];
			const double share = H.shr[pf];
			const double et = m+H.dly[pf];                                // queue_speed_event, flow.c:22
			__syncthreads();                                              // everyone has read sp[pf] before thread 0 rewrites it
			const double delta0 = speed-fs;                               // flow.c:359
			double amount = delta0;
			const u64 sbase = (u64)e*NS_STAGE_INLINE;
			unsigned lseq = 0;
			bool done_bw = false, changed = false, full = false;
			if (amount > NS_EPS && fs+amount > share) {                   // flow.c:64-78
				double es = share;
				if (fs > share) { es = fs; done_bw = true; } else amount = share-fs;
				if (tid == 0) { ps.st_time[sbase] = et; ps.st_speed[sbase] = es; ps.st_fd[sbase] = f*2+RCV; }
				lseq = 1;
			}
			if (!done_bw && !(amount > -NS_EPS && amount < NS_EPS)) {     // flow.c:80-85
				fs += amount;
				changed = true;
				full = sat && !(amount < NS_EPS && amount > -64) && !(amount > -NS_EPS && usage+amount < mx+NS_EPS);
				if (!full) usage += amount;                               // flow.c:86-124
			}
			long long ovf_base = -1;
			if (full) {                                                   // flow.c:126-203
				if (tid == 0) {
					H.sp[pf] = fs; S.f_speed[SND][f] = fs;
					u64 need = (u64)live+8;
					long long b = (long long)atomicAdd((unsigned long long *)ps.ovf_top, (unsigned long long)need);
					if ((u64)b+need > ps.ovf_cap) { atomicCAS(&st->status, 0, NS_E_TABLE_FULL); b = 0; }
					H.bcast_ll = b;
				}
				double spread_amount = amount;                            // flow.c:112-122
				if (amount > -NS_EPS) spread_amount = amount-mx+usage;
				const bool dec = amount < NS_EPS, inc = !dec && amount > NS_EPS;
				__syncthreads();
				ovf_base = H.bcast_ll;
				// pass 1 (flow.c:128-145): terms in list order, non-zero ones compacted in order
				double term[HUB_ROWS]; unsigned bal[HUB_ROWS];
				#pragma unroll
				for (int r = 0; r < HUB_ROWS; r++) {
					double t = 0.0;
					if (r < rows) {
						int p = len-1-r*PAR_THREADS-tid;
						if (p >= 0 && p != pf && H.nf[p] >= 0) {
							double sp = H.sp[p];
							if (dec) { double ls = H.rsh[p]; if (sp < ls) t = ls-sp; }
							else if (inc) { double ls = H.shr[p]; if (sp > ls) t = sp-ls; }
						}
						unsigned b = __ballot_sync(FULLMASK, t != 0.0);
						bal[r] = b;
						if (lane == 0) H.cnt_a[r*PAR_WARPS+warp] = __popc(b);
					}
					term[r] = t;
				}
				__syncthreads();
				if (warp == 0) hub_scan(H.cnt_a, rows*PAR_WARPS, &H.tot_a, lane);
				__syncthreads();
				#pragma unroll
				for (int r = 0; r < HUB_ROWS; r++)
					if (r < rows && term[r] != 0.0) H.term[H.cnt_a[r*PAR_WARPS+warp]+__popc(bal[r] & ((1u << lane)-1))] = term[r];
				__syncthreads();
				if (tid == 0) H.bcast_d = hub_serial_sum(H.term, H.tot_a);
				__syncthreads();
				const double esum = H.bcast_d;
				if (dec && -amount > esum) amount = -esum;                // flow.c:149-154
				const bool dec2 = amount < NS_EPS, inc2 = !dec2 && amount > NS_EPS;
				// pass 2 (flow.c:156-200): new speeds, one event per adjusted flow in list order, values for new_use
				double val[HUB_ROWS];
				#pragma unroll
				for (int r = 0; r < HUB_ROWS; r++) {
					double v = 0.0; int ok = 0;
					if (r < rows) {
						int p = len-1-r*PAR_THREADS-tid;
						if (p == pf) v = fs;
						else if (p >= 0 && H.nf[p] >= 0) {
							const int nf = H.nf[p];
							double sp = H.sp[p];
							v = sp;
							int adj = 0;
							if (dec2) {
								double ls = H.rsh[p];
								if (sp < ls) { double delta = ls-sp; double nsp = sp-amount*delta/esum; H.sp[p] = nsp; S.f_speed[SND][nf] = nsp; v = nsp; adj = 1; }   // :171-178
							} else if (inc2) {
								double ls = H.shr[p];
								if (sp > ls) { double delta = sp-ls; double nsp = sp-spread_amount*delta/esum; H.sp[p] = nsp; S.f_speed[SND][nf] = nsp; v = nsp; adj = 1; }   // :187-189
							}
							if (adj) {
								double t = m+H.dly[p];
								ok = 1;
								if (t < m) { if (t < m-NS_EPS) ok = 0; }
								if (isnan(t)) ok = 0;
								if (!ok) atomicCAS(&st->status, 0, isnan(t) ? NS_E_NAN_TIME : NS_E_BACK_IN_TIME);
							}
						}
						unsigned b = __ballot_sync(FULLMASK, ok), bv = __ballot_sync(FULLMASK, v != 0.0);
						bal[r] = b;
						if (lane == 0) { H.cnt_a[r*PAR_WARPS+warp] = __popc(b); H.cnt_b[r*PAR_WARPS+warp] = __popc(bv); }
					}
					val[r] = v;                                               // an adjusted flow emits its new speed, which is its value
				}
				__syncthreads();
				if (warp == 0) hub_scan(H.cnt_a, rows*PAR_WARPS, &H.tot_a, lane);
				if (warp == 1) hub_scan(H.cnt_b, rows*PAR_WARPS, &H.tot_b, lane);
				__syncthreads();
				#pragma unroll
				for (int r = 0; r < HUB_ROWS; r++) {
					if (r >= rows) continue;
					const unsigned lt = (1u << lane)-1;
					const unsigned bv = __ballot_sync(FULLMASK, val[r] != 0.0);
					if ((bal[r] >> lane) & 1u) {
						int p = len-1-r*PAR_THREADS-tid;
						u64 q = (u64)lseq+(u64)(H.cnt_a[r*PAR_WARPS+warp]+__popc(bal[r] & lt));
						u64 si = q < NS_STAGE_INLINE ? sbase+q : ps.inline_n+(u64)ovf_base+(q-NS_STAGE_INLINE);
						double t = m+H.dly[p];
						if (t < m) t = m;
						ps.st_time[si] = t; ps.st_speed[si] = val[r]; ps.st_fd[si] = H.nf[p]*2+RCV;
					}
					if (val[r] != 0.0) H.term[H.cnt_b[r*PAR_WARPS+warp]+__popc(bv & lt)] = val[r];
				}
				__syncthreads();
				lseq += (unsigned)H.tot_a;
				if (tid == 0) H.bcast_d = hub_serial_sum(H.term, H.tot_b);   // flow.c:199
				__syncthreads();
				usage = H.bcast_d;                                        // flow.c:201
			}
			if (tid == 0) {
				if (!(et >= m)) atomicCAS(&st->status, 0, isnan(et) ? NS_E_NAN_TIME : NS_E_BACK_IN_TIME);
				if (changed && !full) { H.sp[pf] = fs; S.f_speed[SND][f] = fs; }
				if (changed) S.n_usage[SND][n] = usage;
				if (delta0 > 0) {                                         // flow.c:371-374
					u64 q = lseq;
					u64 si = q < NS_STAGE_INLINE ? sbase+q : ps.inline_n+(u64)ovf_base+(q-NS_STAGE_INLINE);
					ps.st_time[si] = et; ps.st_speed[si] = fs; ps.st_fd[si] = f*2+RCV;
				}
				unsigned rec_n = 0;
				if (S.want_rec) {
					u64 *rw = (u64 *)(ps.st_rec+(u64)e*NS_STAGE_REC);
					u64 a0, a1, a2;
					rec19_words(h0, h1, R_SPD|SND, S.f_id[f], speed, a0, a1, a2);        // flow.c:383
					if (changed) {
						u64 u0, u1, u2;
						rec19_words(h0, h1, R_USAGE|SND, node_id, usage, u0, u1, u2);     // record_wrapper.h:6-9
						rw[0] = u0; rw[1] = u1; rw[2] = u2 | (a0 << 24);
						rw[3] = (a0 >> 40) | (a1 << 24); rw[4] = (a1 >> 40) | (a2 << 24);
						rec_n = 38;
					} else { rw[0] = a0; rw[1] = a1; rw[2] = a2; rec_n = 19; }
				}
				unsigned ntr = 0;
				if (S.want_tr) {
					ns_trace_row32 r; r.type = EV_SPEED_CHANGE|(SND << 8); r.ord = S.f_ord[f]; r.time = m; r.v0 = speed; r.v1 = fs;
					ps.st_tr[e] = r; ntr = 1;
				}
				ps.ev_nins[e] = lseq+(delta0 > 0 ? 1u : 0u); ps.ev_ovf[e] = ovf_base; ps.ev_nrec[e] = (uint8_t)rec_n; ps.ev_ntr[e] = (uint8_t)ntr;
				S.q_time[H.ev_slot[i]] = NS_INF;                          // event_pop
				lc[LC_EVENTS] += 1; lc[LC_SPEED_CHANGE] += 1; lc[LC_SOLVES] += 1;
				if (full) { lc[LC_FULL_SPREADS] += 1; lc[LC_FLOW_VISITS] += 2ull*(u64)live; }
			}
			__syncthreads();
		}
	}
	if (tid == 0) { S.n_usage[SND][n] = usage; P.grp_cnt[g] = 0; }
	__syncthreads();
}

// copy the staged outputs of segment-local event e to their final places
// the commit pass of k_run_par: per-iteration scratch in dynamic shared memory
#define COMMIT_TILES 8
struct CommitSmem { u64 tmin[COMMIT_TILES]; u64 pad_[2]; unsigned char rec[PAR_THREADS*NS_STAGE_REC+32]; };
// the event_adds of one small event (<= 32), by its own thread: inline stage slots loaded together, ring tile
// minima collected in shared memory (one global update per tile and iteration instead of one per event_add)
__device__ __forceinline__ void commit_small_inserts(const Sim &S, const ParStage &ps, CommitSmem &Cs, int e, unsigned nins, u64 seq0, u64 run_ins){
	const long long ovf = nins > NS_STAGE_INLINE ? ps.ev_ovf[e] : -1;
	int fdv[NS_STAGE_INLINE]; double tv[NS_STAGE_INLINE], sv[NS_STAGE_INLINE]; uint32_t gv[NS_STAGE_INLINE];
	#pragma unroll
	for (int j = 0; j < NS_STAGE_INLINE; j++) {
		fdv[j] = NS_HOLE_PURE; tv[j] = 0; sv[j] = 0;
		if ((unsigned)j < nins) { const u64 si = (u64)e*NS_STAGE_INLINE+j; fdv[j] = ps.st_fd[si]; tv[j] = ps.st_time[si]; sv[j] = ps.st_speed[si]; }
	}
	#pragma unroll
	for (int j = 0; j < NS_STAGE_INLINE; j++) gv[j] = fdv[j] >= 0 ? S.f_gen[fdv[j] >> 1] : 0u;
	for (unsigned j = 0; j < nins; j++) {
		int fd; double t, sp; uint32_t gen;
		if (j < NS_STAGE_INLINE) { fd = fdv[j]; t = tv[j]; sp = sv[j]; gen = gv[j]; }
		else { const u64 si = ps.inline_n+(u64)ovf+(j-NS_STAGE_INLINE); fd = ps.st_fd[si]; t = ps.st_time[si]; sp = ps.st_speed[si]; gen = fd >= 0 ? S.f_gen[fd >> 1] : 0u; }
		const u64 seq = seq0+j;
		if (fd >= 0) {
			const u64 slot = seq & (S.ring_cap-1);
			S.q_time[slot] = t; S.q_speed[slot] = sp; S.q_fd[slot] = fd; S.q_gen[slot] = gen;
			const u64 rel = (seq >> NS_TILE_SHIFT)-(run_ins >> NS_TILE_SHIFT);
			if (rel < COMMIT_TILES) atomicMin((unsigned long long *)&Cs.tmin[rel], (unsigned long long)__double_as_longlong(t));
			else { double *tm = &S.q_tmin[slot >> NS_TILE_SHIFT]; if (*(volatile double *)tm > t) tmin_lower(tm, t); }
		} else if (fd != NS_HOLE_PURE) {
			const int ls = -fd-1;
			const u64 prov = NS_SEQ_PROV | ((u64)e << 24) | j;
			if (S.l_seq[ls] == prov) S.l_seq[ls] = seq;                   // only the last add of a slot survives
		}
	}
}

__device__ __forceinline__ u64 gtimer(){ u64 t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// phase ids for SimState::phase_ns
enum { PH_ARGMIN = 0, PH_EXTRACT, PH_SEQ, PH_GROUP, PH_SOLVE, PH_COMMIT, PH_TAIL, PH_KERNEL, PH_SOLVE1, PH_N };
// phase timing and the slowest-group probe exist only in profiling builds (make PROF=1 -> libnetsim_b200_prof.so)
#ifdef NS_PROFILE
#define PHASE_MARK(id) do { if (leader) { u64 t_ = gtimer(); st->phase_ns[id] += t_-t_ph; t_ph = t_; } } while (0)
#else
#define PHASE_MARK(id) do { } while (0)
#endif

}   // namespace
