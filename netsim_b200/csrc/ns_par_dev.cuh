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

// copy the staged outputs of segment-local event e to their final places
__device__ __forceinline__ void par_commit_event(const Sim &S, const ParStage &ps, int e, u64 seq0, u64 rec0, u64 tr0, int lane0, int lstep){
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
	if (lane0 == 0) {
		unsigned nrec = ps.ev_nrec[e];
		if (nrec) { const uint8_t *src = ps.st_rec+(u64)e*NS_STAGE_REC; uint8_t *dst = S.rec+rec0; for (unsigned k = 0; k < nrec; k++) dst[k] = src[k]; }
		if (ps.ev_ntr[e] && tr0 < S.tr_cap) S.tr[tr0] = ps.st_tr[e];
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
