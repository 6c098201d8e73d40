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

namespace cgp = cooperative_groups;

#define PAR_THREADS 256
#ifndef PAR_MIN_BLOCKS
#define PAR_MIN_BLOCKS 1       // CTAs per SM the persistent kernel is compiled for (register budget)
#endif
#define PAR_WARPS (PAR_THREADS/32)
#define PAR_MIN_BATCH 48          // smaller batches are cheaper on the single-warp path
#define SPARSE_STREAK 6            // that many small batches in a row: the grid barriers cost more than the events
#define SPARSE_MAX_TILES 4096     // ... provided one warp can scan the tile minima of both event stores

struct ParScratch {
	int enabled;
	int seg_cap;                  // events per parallel segment
	ParStage stage;               // staging arrays (device pointers)
	ParStage *d_stage;
	// per-timestamp
	int *tile_cnt, *tile_off;     // per ring-window tile: matches, exclusive offset
	// grouping
	int *grp_cnt, *grp_off, *grp_fill, *active, *grp_items, *b_grp;
	int *deferred, *defer_pos;     // groups handed from the per-thread pass to the per-warp pass
	// parallel module init: CREATE ops of group 0 scheduled into conflict-free waves (host, par_plan_init)
	int n_init, n_waves; int *wave_off, *wave_ops; u64 *cta_draws;
	// scan
	u64 *cta_ins; unsigned *cta_rec, *cta_tr;
	// control block (device)
	struct Ctl {
		u64 tmin_bits, first_live;
		int nb, nl, n_active, cursor, work, seg_lo, seg_hi, n_deferred, dbg_g, dbg_deg;
		int piece, last, pos, li, l_lo, l_hi;      // the current piece of the batch (par_plan_piece)
		int small_streak;                          // consecutive small batches (sparse phase detection)
		u64 dbg_key;
		u64 tot_ins; unsigned tot_rec, tot_tr;
		u64 lc[LC_N];
	} *ctl;
};

__global__ void k_init_groups(const Sim *sims, int n);
__global__ void k_run_seq(const Sim *sims, const SeqScratch *scr, int n);

template <typename A> static int par_alloc(ParScratch &p, const Sim &S, bool enable, int n_init, A alloc){
	memset(&p, 0, sizeof p);
	p.enabled = enable;
	if (!enable) return 0;
	int rc;
	p.n_init = n_init;
	p.seg_cap = std::max((int)std::min<u64>((u64)S.batch_cap, 1u << 20), n_init);
	ParStage &st = p.stage;
	st.inline_n = (u64)NS_STAGE_INLINE*p.seg_cap;
	st.ovf_cap = std::max<u64>(1u << 22, 64ull*S.cap_flows);
	u64 n = st.inline_n+st.ovf_cap;
#define PA(ptr, cnt) if ((rc = alloc((void **)&(ptr), (size_t)(cnt)*sizeof(*(ptr))))) return rc
	PA(st.st_time, n); PA(st.st_speed, n); PA(st.st_fd, n); PA(st.ovf_top, 1);
	PA(st.st_rec, (u64)NS_STAGE_REC*p.seg_cap); PA(st.st_tr, p.seg_cap);
	PA(st.ev_nins, p.seg_cap); PA(st.ev_ovf, p.seg_cap); PA(st.ev_nrec, p.seg_cap); PA(st.ev_ntr, p.seg_cap); PA(st.ev_idpos, p.seg_cap);
	PA(p.d_stage, 1);
	PA(p.tile_cnt, S.n_ring_tiles+1); PA(p.tile_off, S.n_ring_tiles+1);
	PA(p.grp_cnt, 2*S.n_nodes+1); PA(p.grp_off, 2*S.n_nodes+1); PA(p.grp_fill, 2*S.n_nodes+1);
	PA(p.active, p.seg_cap); PA(p.grp_items, p.seg_cap); PA(p.b_grp, p.seg_cap);
	PA(p.deferred, p.seg_cap); PA(p.defer_pos, 2*S.n_nodes+1);
	PA(p.cta_ins, 4096); PA(p.cta_rec, 4096); PA(p.cta_tr, 4096); PA(p.cta_draws, 4096);
	if (n_init) { PA(p.wave_off, n_init+2); PA(p.wave_ops, n_init); }
	PA(p.ctl, 1);
#undef PA
	cudaError_t e;
	if ((e = cudaMemcpy(p.d_stage, &p.stage, sizeof p.stage, cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
	if ((e = cudaMemset(p.grp_cnt, 0, (size_t)(2*S.n_nodes+1)*4)) != cudaSuccess) return (int)e;
	if ((e = cudaMemset(p.ctl, 0, sizeof *p.ctl)) != cudaSuccess) return (int)e;
	if ((e = cudaMemset(st.ovf_top, 0, 8)) != cudaSuccess) return (int)e;
	return 0;
}

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
#define PHASE_MARK(id) do { if (leader) { u64 t_ = gtimer(); st->phase_ns[id] += t_-t_ph; t_ph = t_; } } while (0)

// ---------------------------------------------------------------- pieces of one batch
// A batch is handled as a sequence of pieces in sequence-number order: long runs of SPEED_CHANGE events
// go through the parallel phases D-F, everything else (FLOW_DONE / USER / THROTTLE / DRAIN events and
// the short runs between them, relay simulations, the last events before a max_events bound)
// through the single-warp path.  The leader thread plans one piece at a time because a sequential
// piece may change what follows (flow_close cancels events, a USER group may create a relay).
enum { PIECE_END = 0, PIECE_SEQ = 1, PIECE_PAR = 2 };
__device__ __forceinline__ int lower_bound_seq(const u64 *a, int lo, int hi, u64 key){
	while (lo < hi) { int mid = (lo+hi) >> 1; if (a[mid] < key) lo = mid+1; else hi = mid; }
	return lo;
}
__device__ __noinline__ void par_finish_piece(const Sim &S, SimState *st, ParScratch::Ctl *ctl, const ParScratch &P){
	st->tail += ctl->tot_ins; st->burns += ctl->tot_ins;
	st->rec_len += S.want_rec ? ctl->tot_rec : 0;
	u64 ntr = st->tr_len+ctl->tot_tr;
	st->tr_len = S.want_tr ? (ntr < S.tr_cap ? ntr : S.tr_cap) : 0;
	st->cnt[NS_CNT_EVENTS] += ctl->lc[LC_EVENTS]; st->cnt[NS_CNT_SPEED_CHANGE] += ctl->lc[LC_SPEED_CHANGE];
	st->cnt[NS_CNT_SOLVES] += ctl->lc[LC_SOLVES]; st->cnt[NS_CNT_FULL_SPREADS] += ctl->lc[LC_FULL_SPREADS];
	st->cnt[NS_CNT_FLOW_VISITS] += ctl->lc[LC_FLOW_VISITS]; st->cnt[NS_CNT_PAR_EVENTS] += ctl->lc[LC_EVENTS];
	st->cnt[NS_CNT_RETIMES] += ctl->lc[LC_RETIMES];
	st->cnt[NS_CNT_WAVES] += 1;
	{ u64 k_ = ctl->dbg_key; st->phase_ns[13] += k_ >> 20; if ((k_ >> 20) > st->phase_ns[10]) { st->phase_ns[10] = k_ >> 20; st->phase_ns[11] = ctl->dbg_deg; st->phase_ns[12] = k_ & 0xfffff; } ctl->dbg_key = 0; st->phase_ns[14] += ctl->n_deferred; }
	st->cnt[NS_CNT_SEQ] = st->tail;
	for (int q = 0; q < LC_N; q++) ctl->lc[q] = 0;
	ctl->pos = ctl->seg_hi;
	ctl->n_active = 0; ctl->cursor = 0; ctl->work = 0; ctl->n_deferred = 0;
	ctl->tot_ins = 0; ctl->tot_rec = 0; ctl->tot_tr = 0;
	*P.stage.ovf_top = 0;
}
__device__ __noinline__ void par_plan_piece(const Sim &S, SimState *st, ParScratch::Ctl *ctl, const ParScratch &P, const SeqScratch &sc, int nb, int nl){
	if (ctl->piece == PIECE_PAR) par_finish_piece(S, st, ctl, P);
	ctl->piece = PIECE_END; ctl->last = 0;
	const u64 ev_max = st->max_events, evn = st->cnt[NS_CNT_EVENTS];
	if (st->status || st->exit_flag || st->host_stop || (ev_max && evn >= ev_max)) return;
	const int pos = ctl->pos, li = ctl->li;
	if (pos >= nb && li >= nl) return;
	const u64 budget = ev_max ? ev_max-evn : ~0ull;
	if (S.want_rec) {                                                 // soft stop before a piece whose records might not fit
		u64 reserve = S.rec_cap > 8ull*NS_REC_RESERVE ? NS_REC_RESERVE : S.rec_cap/8;
		u64 left = (u64)(nb-pos)+(u64)(nl-li);
		u64 next = left < (u64)P.seg_cap ? left : (u64)P.seg_cap;
		if (st->rec_len+next*NS_REC_PER_EVENT+reserve > S.rec_cap) { st->host_stop = 2; return; }
	}
	const bool no_par = st->relay_seen || budget < PAR_MIN_BATCH;
	const u64 ls = li < nl ? sc.lb_seq[li] : ~0ull;
	int end = pos;
	if (pos < nb && S.b_seq[pos] < ls) end = lower_bound_seq(S.b_seq, pos, nb, ls);
	if (!no_par && end-pos >= PAR_MIN_BATCH) {
		u64 lim = budget < (u64)P.seg_cap ? budget : (u64)P.seg_cap;
		int hi = (u64)(end-pos) <= lim ? end : pos+(int)lim;
		ctl->piece = PIECE_PAR; ctl->seg_lo = pos; ctl->seg_hi = hi;
		ctl->last = hi == nb && li >= nl;
		return;
	}
	int e2 = end, l2 = li;
	if (no_par) { e2 = nb; l2 = nl; }
	else while (true) {
		if (l2 >= nl) { e2 = nb; break; }
		l2++;                                                         // the next long-lived event
		u64 ls2 = l2 < nl ? sc.lb_seq[l2] : ~0ull;
		int nx = lower_bound_seq(S.b_seq, e2, nb, ls2);
		if (nx-e2 >= PAR_MIN_BATCH) break;                            // a long run follows: it gets its own parallel piece
		e2 = nx;
		if (l2 >= nl && e2 >= nb) break;
	}
	ctl->piece = PIECE_SEQ; ctl->seg_lo = pos; ctl->seg_hi = e2; ctl->l_lo = li; ctl->l_hi = l2;
	ctl->last = e2 == nb && l2 == nl;
}

// ---------------------------------------------------------------- the persistent kernel
__global__ void __launch_bounds__(PAR_THREADS, PAR_MIN_BLOCKS) k_run_par(const Sim *sp, const SeqScratch *scrp, ParScratch P){
	cgp::grid_group grid = cgp::this_grid();
	const Sim &S = *sp;
	SimState *st = S.st;
	ParScratch::Ctl *ctl = P.ctl;
	const SeqScratch sc = *scrp;
	__shared__ double sh_d[PAR_WARPS];
	__shared__ u64 sh_u[PAR_WARPS];
	__shared__ unsigned sh_r[PAR_WARPS], sh_t[PAR_WARPS];
	__shared__ int sh_big[PAR_THREADS]; __shared__ int sh_nbig, sh_status;
	__shared__ u64 sh_bseq[PAR_THREADS], sh_brec[PAR_THREADS], sh_btr[PAR_THREADS];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int gw = blockIdx.x*PAR_WARPS+warp, nw = gridDim.x*PAR_WARPS;
	const u64 gtid = (u64)blockIdx.x*PAR_THREADS+tid, gth = (u64)gridDim.x*PAR_THREADS;
	const u64 tmask = (u64)S.n_ring_tiles-1, smask = S.ring_cap-1;
	const bool leader = blockIdx.x == 0 && tid == 0;
	u64 t_ph = gtimer();
	const u64 t_kernel0 = t_ph;

	while (true) {
		// ---- A: argmin over tile minima ------------------------------------------------------
		if (leader) { ctl->tmin_bits = ~0ull; ctl->first_live = ~0ull; ctl->nl = 0; ctl->nb = 0; }
		grid.sync();
		if (st->status || st->exit_flag || st->host_stop || st->want_sparse || (st->max_events && st->cnt[NS_CNT_EVENTS] >= st->max_events)) break;
		const u64 head = st->head, tail = st->tail;
		const u64 t_first = head >> NS_TILE_SHIFT, t_last = tail ? (tail-1) >> NS_TILE_SHIFT : 0;
		const u64 n_wt = tail > head ? t_last-t_first+1 : 0;
		{
			double m = NS_INF; u64 fl = ~0ull;
			for (u64 k = gtid; k < n_wt; k += gth) {
				double v = S.q_tmin[(t_first+k) & tmask];
				if (v != NS_INF && t_first+k < fl) fl = t_first+k;
				m = fmin(m, v);
			}
			for (u64 k = gtid; k < (u64)S.n_l_tiles; k += gth) m = fmin(m, S.l_tmin[k]);
			m = block_min_d(m, sh_d);
			fl = warp_min_u64(fl);
			if (lane == 0 && fl != ~0ull) atomicMin((unsigned long long *)&ctl->first_live, fl);
			if (tid == 0 && m != NS_INF) atomicMin((unsigned long long *)&ctl->tmin_bits, (u64)__double_as_longlong(m));
		}
		grid.sync();
		PHASE_MARK(PH_ARGMIN);
		const u64 mb = ctl->tmin_bits;
		if (mb == ~0ull) break;                                       // queue empty
		const double m = __longlong_as_double((long long)mb);
		// ---- B1: per-tile match counts; block 0 warp 0 also extracts the long-lived events -----
		for (u64 k = gw; k < n_wt; k += nw) {
			u64 at = t_first+k;
			int cnt = 0;
			if (S.q_tmin[at & tmask] == m) {
				u64 base = at << NS_TILE_SHIFT;
				#pragma unroll
				for (int i = 0; i < NS_TILE; i += 32) { u64 s = base+i+lane; cnt += (S.q_time[s & smask] == m && s >= head && s < tail); }
				for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(FULLMASK, cnt, d);
			}
			if (lane == 0) P.tile_cnt[k] = cnt;
		}
		if (gw == nw-1) {
			Ctx c; ctx_load(c, sp, lane);
			int nl = seq_collect_l(c, m, sc);
			if (lane == 0) ctl->nl = nl;
		}
		grid.sync();
		// ---- B2: exclusive scan of tile counts (block 0) ---------------------------------------
		if (blockIdx.x == 0) {
			__shared__ int sh_scan[PAR_THREADS];
			u64 per = (n_wt+PAR_THREADS-1)/PAR_THREADS;
			u64 lo = tid*per, hi = lo+per < n_wt ? lo+per : n_wt;
			int s = 0;
			for (u64 k = lo; k < hi; k++) s += P.tile_cnt[k];
			sh_scan[tid] = s;
			__syncthreads();
			if (tid == 0) { int acc = 0; for (int i = 0; i < PAR_THREADS; i++) { int v = sh_scan[i]; sh_scan[i] = acc; acc += v; } ctl->nb = acc; }
			__syncthreads();
			int acc = sh_scan[tid];
			for (u64 k = lo; k < hi; k++) { int v = P.tile_cnt[k]; P.tile_off[k] = acc; acc += v; }
		}
		grid.sync();
		const int nb = ctl->nb, nl = ctl->nl;
		if (nb > S.batch_cap) { if (leader) atomicCAS(&st->status, 0, NS_E_BATCH_FULL); continue; }
		// ---- B3: write the batch in order, refresh the tile minima ------------------------------
		for (u64 k = gw; k < n_wt; k += nw) {
			u64 at = t_first+k;
			if (S.q_tmin[at & tmask] != m) continue;
			u64 base = at << NS_TILE_SHIFT;
			int off = P.tile_off[k];
			double nm = NS_INF;
			for (int i0 = 0; i0 < NS_TILE; i0 += 32*8) {                 // 8 independent loads in flight per lane
				double tv[8];
				#pragma unroll
				for (int u = 0; u < 8; u++) tv[u] = S.q_time[(base+i0+32*u+lane) & smask];
				#pragma unroll
				for (int u = 0; u < 8; u++) {
					u64 s = base+i0+32*u+lane;
					double t = tv[u];
					bool match = t == m && s >= head && s < tail;
					unsigned mm = __ballot_sync(FULLMASK, match);
					if (match) S.b_seq[off+__popc(mm & ((1u << lane)-1))] = s;
					off += __popc(mm);
					if (!match) nm = fmin(nm, t);
				}
			}
			nm = warp_min_d(nm);
			if (lane == 0) S.q_tmin[at & tmask] = nm;
		}
		if (leader) {
			u64 fl = ctl->first_live;
			st->head = fl == ~0ull ? ((tail >> NS_TILE_SHIFT) << NS_TILE_SHIFT) : (fl << NS_TILE_SHIFT);
		}
		grid.sync();
		PHASE_MARK(PH_EXTRACT);
		if (nb+nl == 0) continue;                                     // only stale lower bounds surfaced
		// ---- C: the batch as a sequence of pieces (par_plan_piece) -----------------------------------
		if (leader) {
			st->now = m;
			st->cnt[NS_CNT_BATCHES] += 1;
			if ((u64)(nb+nl) > st->cnt[NS_CNT_MAX_BATCH]) st->cnt[NS_CNT_MAX_BATCH] = nb+nl;
			ctl->pos = 0; ctl->li = 0; ctl->piece = PIECE_END;
			// sparse phase: many small batches in a row and event stores one warp can scan -> hand over to k_run_seq
			ctl->small_streak = nb+nl < PAR_MIN_BATCH ? ctl->small_streak+1 : 0;
			if (ctl->small_streak >= SPARSE_STREAK && n_wt+(u64)S.n_l_tiles <= SPARSE_MAX_TILES) { st->want_sparse = 1; ctl->small_streak = 0; }
			ctl->n_active = 0; ctl->cursor = 0; ctl->work = 0; ctl->n_deferred = 0;
			ctl->tot_ins = 0; ctl->tot_rec = 0; ctl->tot_tr = 0;
			*P.stage.ovf_top = 0;
		}
		bool done_all = false;                                        // the last piece covered the rest of the batch
		while (true) {
			if (leader) par_plan_piece(S, st, ctl, P, sc, nb, nl);
			grid.sync();
			const int piece = ctl->piece, last = ctl->last;
			if (piece == PIECE_END) break;
			if (piece == PIECE_SEQ) {
				if (blockIdx.x == 0 && warp == 0) {
					Ctx c; ctx_load(c, sp, lane);
					c.now = m;
					int pi = ctl->seg_lo, pj = ctl->l_lo;
					const int pe = ctl->seg_hi, le = ctl->l_hi;
					seq_process(c, m, pi, pe, pj, le, sc);
					if (last && (pi < pe || pj < le)) seq_restore(c, m, pi, pe, pj, le, sc);   // stopped inside the last piece
					L0 { st->cnt[NS_CNT_SEQ] = c.tail; ctl->pos = last ? nb : pi; ctl->li = last ? nl : pj; }
					ctx_flush(c);
				}
				PHASE_MARK(PH_SEQ);
				if (last) { done_all = true; break; }
				continue;
			}
			const int seg_lo = ctl->seg_lo, seg_hi = ctl->seg_hi;
			const int ns = seg_hi-seg_lo;
			// ---- D1: group of every event; cancelled events drop out ---------------------------
			for (u64 e = gtid; e < (u64)ns; e += gth) {
				u64 s = S.b_seq[seg_lo+e];
				u64 slot = s & smask;
				int fd = S.q_fd[slot];
				int f = fd >> 1, dir = fd & 1;
				int g = -1;
				if (S.f_gen[f] == S.q_gen[slot] && S.f_alive[f]) {
					g = S.f_peer[dir][f]*2+dir;
					if (atomicAdd(&P.grp_cnt[g], 1) == 0) P.active[atomicAdd(&ctl->n_active, 1)] = g;
				} else {
					S.q_time[slot] = NS_INF;                          // flow_close cancelled it (flow.c:260-270)
					P.stage.ev_nins[e] = 0; P.stage.ev_nrec[e] = 0; P.stage.ev_ntr[e] = 0; P.stage.ev_ovf[e] = -1;
				}
				P.b_grp[e] = g;
			}
			grid.sync();
			const int n_active = ctl->n_active;
			// ---- D2: slice of grp_items per active group ---------------------------------------
			for (u64 k = gtid; k < (u64)n_active; k += gth) {
				int g = P.active[k];
				P.grp_off[g] = atomicAdd(&ctl->cursor, P.grp_cnt[g]);
				P.grp_fill[g] = 0;
			}
			grid.sync();
			// ---- D3: fill -------------------------------------------------------------------------
			for (u64 e = gtid; e < (u64)ns; e += gth) {
				int g = P.b_grp[e];
				if (g >= 0) P.grp_items[P.grp_off[g]+atomicAdd(&P.grp_fill[g], 1)] = (int)e;
			}
			grid.sync();
			PHASE_MARK(PH_GROUP);
			// ---- E1: one THREAD per group: O(1) solves (no full spread), events in sequence order ----
			{
				Ctx c;
				c.S = sp; c.st = st; c.lane = 0; c.par = 1; c.scalar = 1; c.ps = P.d_stage; c.now = m;
				c.head = head; c.tail = 0; c.burns = 0; c.rec_len = 0; c.tr_len = 0;
				c.ev = 0; c.lseq = 0; c.rec_n = 0; c.tr_n = 0; c.ovf_base = -1;
				for (int q = 0; q < LC_N; q++) c.lc[q] = 0;
				for (u64 k = gtid; k < (u64)n_active; k += gth) {
					int g = P.active[k];
					int cnt = P.grp_cnt[g], off = P.grp_off[g];
					int j = 0;
					if (cnt > 32) j = -1;                             // long chains: sorted and run by a warp
					else {
						int *it = P.grp_items+off;
						for (int a = 1; a < cnt; a++) {               // insertion sort by batch position
							int x = it[a], b = a-1;
							while (b >= 0 && it[b] > x) { it[b+1] = it[b]; b--; }
							it[b+1] = x;
						}
						for (; j < cnt; j++) {
							int e = it[j];
							u64 slot = S.b_seq[seg_lo+e] & smask;
							int fd = S.q_fd[slot];
							if (sc_needs_full_spread(S, fd >> 1, fd & 1, S.q_speed[slot])) break;
							par_handle_event(c, P, seg_lo, e, m);
						}
					}
					if (j < cnt) { P.defer_pos[g] = j; P.deferred[atomicAdd(&ctl->n_deferred, 1)] = g; }
					else P.grp_cnt[g] = 0;                            // ready for the next segment
				}
				for (int q = 0; q < LC_N; q++) if (c.lc[q]) atomicAdd((unsigned long long *)&ctl->lc[q], c.lc[q]);
			}
			grid.sync();
			PHASE_MARK(PH_SOLVE1);
			// ---- E2: one WARP per remaining group: full spreads strided over the lanes ----------------
			{
				const int n_def = ctl->n_deferred;
				Ctx c;
				c.S = sp; c.st = st; c.lane = lane; c.par = 1; c.scalar = 0; c.ps = P.d_stage; c.now = m;
				c.head = head; c.tail = 0; c.burns = 0; c.rec_len = 0; c.tr_len = 0;
				c.ev = 0; c.lseq = 0; c.rec_n = 0; c.tr_n = 0; c.ovf_base = -1;
				for (int q = 0; q < LC_N; q++) c.lc[q] = 0;
				while (true) {
					int k = 0;
					if (lane == 0) k = atomicAdd(&ctl->work, 1);
					k = __shfl_sync(FULLMASK, k, 0);
					if (k >= n_def) break;
					int g = P.deferred[k];
					int cnt = P.grp_cnt[g], off = P.grp_off[g];
					int j = P.defer_pos[g];
					u64 t_g0 = gtimer();
					if (j < 0) { warp_sort_global(P.grp_items+off, cnt, lane); j = 0; }
					int j0 = j;
					if (cnt-j >= GE_MIN_CHAIN && S.adj_len[g] <= 32) ge_chain(c, P, g, seg_lo, off, j, cnt, m);
					else for (; j < cnt; j++) par_handle_event(c, P, seg_lo, P.grp_items[off+j], m);
					if (lane == 0) {
						P.grp_cnt[g] = 0;                             // ready for the next segment
						u64 dt = gtimer()-t_g0;                       // profiling aid: the slowest group of the segment
						u64 key = (dt << 20) | (u64)(cnt-j0 < 0xfffff ? cnt-j0 : 0xfffff);
						if (atomicMax((unsigned long long *)&ctl->dbg_key, key) < key) { ctl->dbg_g = g; ctl->dbg_deg = S.adj_live[g]; }
					}
				}
				if (lane == 0) for (int q = 0; q < LC_N; q++) if (c.lc[q]) atomicAdd((unsigned long long *)&ctl->lc[q], c.lc[q]);
			}
			grid.sync();
			PHASE_MARK(PH_SOLVE);
			// ---- F1: per-CTA sums over a contiguous chunk of the segment ----------------------------
			const int per = (ns+gridDim.x-1)/gridDim.x;
			const int c_lo = (int)blockIdx.x*per < ns ? (int)blockIdx.x*per : ns;
			const int c_hi = c_lo+per < ns ? c_lo+per : ns;
			{
				u64 a = 0; unsigned r = 0, t = 0;
				for (int e = c_lo+tid; e < c_hi; e += PAR_THREADS) { a += P.stage.ev_nins[e]; r += P.stage.ev_nrec[e]; t += P.stage.ev_ntr[e]; }
				for (int d = 16; d; d >>= 1) { a += __shfl_xor_sync(FULLMASK, a, d); r += __shfl_xor_sync(FULLMASK, r, d); t += __shfl_xor_sync(FULLMASK, t, d); }
				__syncthreads();
				if (lane == 0) { sh_u[warp] = a; sh_r[warp] = r; sh_t[warp] = t; }
				__syncthreads();
				if (tid == 0) {
					for (int w = 1; w < PAR_WARPS; w++) { a += sh_u[w]; r += sh_r[w]; t += sh_t[w]; }
					P.cta_ins[blockIdx.x] = a; P.cta_rec[blockIdx.x] = r; P.cta_tr[blockIdx.x] = t;
				}
			}
			grid.sync();
			// ---- F2: chunk bases, capacity checks, ordered commit --------------------------------
			{
				u64 b_ins = 0, t_ins = 0; unsigned b_rec = 0, t_rec = 0, b_tr = 0, t_tr = 0;
				for (int b = 0; b < (int)gridDim.x; b++) {
					u64 a = P.cta_ins[b]; unsigned r = P.cta_rec[b], t = P.cta_tr[b];
					if (b < (int)blockIdx.x) { b_ins += a; b_rec += r; b_tr += t; }
					t_ins += a; t_rec += r; t_tr += t;
				}
				const u64 tail0 = st->tail, rec0 = st->rec_len, tr0 = st->tr_len;
				if (tid == 0) sh_status = st->status;
				__syncthreads();
				bool okc = sh_status == 0;            // block-uniform: the CAS below only fires when okc is false anyway
				if (tail0+t_ins-st->head > S.ring_cap) { okc = false; if (leader) atomicCAS(&st->status, 0, NS_E_RING_FULL); }
				if (S.want_rec && rec0+t_rec > S.rec_cap) { okc = false; if (leader) atomicCAS(&st->status, 0, NS_E_RECORDS_FULL); }
				if (leader) { ctl->tot_ins = t_ins; ctl->tot_rec = t_rec; ctl->tot_tr = t_tr; }
				if (okc) {
					// block-wide running scan over the chunk, PAR_THREADS events at a time
					u64 run_ins = tail0+b_ins, run_rec = rec0+b_rec, run_tr = tr0+b_tr;
					for (int e0 = c_lo; e0 < c_hi; e0 += PAR_THREADS) {
						int e = e0+tid;
						unsigned ni = 0, nr = 0, nt = 0;
						if (e < c_hi) { ni = P.stage.ev_nins[e]; nr = P.stage.ev_nrec[e]; nt = P.stage.ev_ntr[e]; }
						u64 xi = ni; unsigned xr = nr, xt = nt;
						for (int d = 1; d < 32; d <<= 1) {
							u64 yi = __shfl_up_sync(FULLMASK, xi, d); unsigned yr = __shfl_up_sync(FULLMASK, xr, d), yt = __shfl_up_sync(FULLMASK, xt, d);
							if (lane >= d) { xi += yi; xr += yr; xt += yt; }
						}
						__syncthreads();
						if (lane == 31) { sh_u[warp] = xi; sh_r[warp] = xr; sh_t[warp] = xt; }
						if (tid == 0) sh_nbig = 0;
						__syncthreads();
						u64 wi = 0; unsigned wr = 0, wt = 0;
						u64 tot_i = 0; unsigned tot_r = 0, tot_t = 0;
						for (int w = 0; w < PAR_WARPS; w++) {
							if (w < warp) { wi += sh_u[w]; wr += sh_r[w]; wt += sh_t[w]; }
							tot_i += sh_u[w]; tot_r += sh_r[w]; tot_t += sh_t[w];
						}
						u64 my_seq = run_ins+wi+xi-ni, my_rec = run_rec+wr+xr-nr, my_tr = run_tr+wt+xt-nt;
						if (e < c_hi) {
							if (ni > 32) {                            // a full spread: copied by a whole warp below
								int q = atomicAdd(&sh_nbig, 1);
								sh_big[q] = e; sh_bseq[q] = my_seq; sh_brec[q] = my_rec; sh_btr[q] = my_tr;
							} else par_commit_event(S, P.stage, e, my_seq, my_rec, my_tr, 0, 1);
						}
						__syncthreads();
						int nbig = sh_nbig;
						for (int q = warp; q < nbig; q += PAR_WARPS)
							par_commit_event(S, P.stage, sh_big[q], sh_bseq[q], sh_brec[q], sh_btr[q], lane, 32);
						run_ins += tot_i; run_rec += tot_r; run_tr += tot_t;
					}
				}
			}
			grid.sync();
			PHASE_MARK(PH_COMMIT);
			if (last) { if (leader) { par_finish_piece(S, st, ctl, P); ctl->piece = PIECE_END; } done_all = true; break; }
		}
		// a stop condition hit inside the batch: what was extracted but not handled goes back under the
		// tile minima (the whole grid lowers one minimum per touched ring tile; the few long-lived
		// events are re-armed by the leader)
		if (!done_all) {
			const int lo = ctl->pos, l0 = ctl->li;
			if (lo < nb) {
				for (u64 k = lo+gtid; k < (u64)nb; k += gth) {
					u64 tl = S.b_seq[k] >> NS_TILE_SHIFT;
					if (k == (u64)lo || (S.b_seq[k-1] >> NS_TILE_SHIFT) != tl) tmin_lower(&S.q_tmin[tl & tmask], m);
				}
			}
			if (leader) for (int k = l0; k < nl; k++) {
				int slot = sc.lb_ref[k];
				if (S.l_time[slot] == m && S.l_seq[slot] == sc.lb_seq[k]) tmin_lower(&S.l_tmin[slot >> NS_TILE_SHIFT], m);
			}
			PHASE_MARK(PH_TAIL);
		}
	}
	if (leader) { u64 t_ = gtimer(); st->phase_ns[PH_TAIL] += t_-t_ph; st->phase_ns[PH_KERNEL] += t_-t_kernel0; }
}

// ---------------------------------------------------------------- parallel module init
// A module's init (modules/p2p.c:122-203 style) establishes its first flows one after the other.
// sim_establish_flow (sim.c:42-94) only touches its two end nodes: the sender's SND group, resource
// and consumer list, the receiver's RCV adjacency, resources and ranges; its bwspread is an increase
// (flow.c:317-318), which never reads another node (flow.c:138).  Two CREATE ops that share no node
// therefore commute, and the host schedules group 0 into waves such that every op comes after all
// earlier ops touching either of its nodes (par_plan_init).  Inside a wave one warp runs one op with
// its event_adds / records staged; sequence numbers, random() draws (flow ids, flow.c:327) and
// record offsets are assigned afterwards in OP order by a prefix scan, so the result is the
// sequential one bit for bit.
__global__ void __launch_bounds__(PAR_THREADS, PAR_MIN_BLOCKS) k_init_par(const Sim *sp, ParScratch P){
	cgp::grid_group grid = cgp::this_grid();
	const Sim &S = *sp;
	SimState *st = S.st;
	ParScratch::Ctl *ctl = P.ctl;
	const ParStage &ps = P.stage;
	__shared__ u64 sh_u[PAR_WARPS], sh_d[PAR_WARPS];
	__shared__ unsigned sh_r[PAR_WARPS];
	__shared__ int sh_big[PAR_THREADS]; __shared__ int sh_nbig;
	__shared__ u64 sh_bseq[PAR_THREADS], sh_brec[PAR_THREADS];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int gw = blockIdx.x*PAR_WARPS+warp, nw = gridDim.x*PAR_WARPS;
	const bool leader = blockIdx.x == 0 && tid == 0;
	const int n0 = P.n_init;
	if (leader) { ctl->n_active = 0; for (int q = 0; q < LC_N; q++) ctl->lc[q] = 0; *P.stage.ovf_top = 0; }
	grid.sync();
	// ---- waves of commuting CREATE ops, one warp per op ---------------------------------------
	{
		Ctx c;
		c.S = sp; c.st = st; c.lane = lane; c.par = 1; c.scalar = 0; c.ps = P.d_stage; c.now = st->now;
		c.head = st->head; c.tail = 0; c.burns = 0; c.rec_len = 0; c.tr_len = 0;
		for (int q = 0; q < LC_N; q++) c.lc[q] = 0;
		int created = 0;
		for (int w = 0; w < P.n_waves; w++) {
			const int lo = P.wave_off[w], hi = P.wave_off[w+1];
			for (int i = lo+gw; i < hi; i += nw) {
				const int k = P.wave_ops[i];
				const ns_sc_op op = S.ops[k];
				c.pre_slot = k; c.ev = (unsigned)k; c.lseq = 0; c.rec_n = 0; c.tr_n = 0; c.ovf_base = -1; c.id_pos = 0;
				int f = sim_establish_flow(c, S.rids[op.c], op.start, op.a, op.b, op.x, op.y);
				if (lane == 0) {
					S.ord2flow[k] = f;
					if (f >= 0) S.f_ord[f] = k;
					ps.ev_nins[k] = c.lseq; ps.ev_ovf[k] = c.ovf_base; ps.ev_nrec[k] = (uint8_t)c.rec_n; ps.ev_ntr[k] = 0;
					ps.ev_idpos[k] = f >= 0 ? c.id_pos : 0xffffffffu;
				}
				created += f >= 0;
				__syncwarp();
			}
			if (w+1 < P.n_waves) grid.sync();
		}
		if (lane == 0) {
			for (int q = 0; q < LC_N; q++) if (c.lc[q]) atomicAdd((unsigned long long *)&ctl->lc[q], c.lc[q]);
			if (created) atomicAdd(&ctl->n_active, created);
		}
	}
	grid.sync();
	// ---- per-CTA sums over a contiguous chunk of the op order -----------------------------------
	const int per = (n0+gridDim.x-1)/gridDim.x;
	const int c_lo = (int)blockIdx.x*per < n0 ? (int)blockIdx.x*per : n0;
	const int c_hi = c_lo+per < n0 ? c_lo+per : n0;
	{
		u64 a = 0, d = 0; unsigned r = 0;
		for (int e = c_lo+tid; e < c_hi; e += PAR_THREADS) {
			unsigned ni = ps.ev_nins[e];
			a += ni; r += ps.ev_nrec[e];
			if (ps.ev_idpos[e] != 0xffffffffu) d += (u64)ni+2;            // + flow id + the range's skip-list insert
		}
		for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(FULLMASK, a, o); d += __shfl_xor_sync(FULLMASK, d, o); r += __shfl_xor_sync(FULLMASK, r, o); }
		if (lane == 0) { sh_u[warp] = a; sh_d[warp] = d; sh_r[warp] = r; }
		__syncthreads();
		if (tid == 0) {
			for (int w = 1; w < PAR_WARPS; w++) { a += sh_u[w]; d += sh_d[w]; r += sh_r[w]; }
			P.cta_ins[blockIdx.x] = a; P.cta_draws[blockIdx.x] = d; P.cta_rec[blockIdx.x] = r;
		}
	}
	grid.sync();
	u64 b_ins = 0, t_ins = 0, b_dr = 0, t_dr = 0; unsigned b_rec = 0, t_rec = 0;
	for (int b = 0; b < (int)gridDim.x; b++) {
		u64 a = P.cta_ins[b], d = P.cta_draws[b]; unsigned r = P.cta_rec[b];
		if (b < (int)blockIdx.x) { b_ins += a; b_dr += d; b_rec += r; }
		t_ins += a; t_dr += d; t_rec += r;
	}
	const u64 tail0 = st->tail, rec0 = st->rec_len;
	bool okc = true;
	if (tail0+t_ins-st->head > S.ring_cap) { okc = false; if (leader) atomicCAS(&st->status, 0, NS_E_RING_FULL); }
	if (S.want_rec && rec0+t_rec > S.rec_cap) { okc = false; if (leader) atomicCAS(&st->status, 0, NS_E_RECORDS_FULL); }
	if (okc) {
		// ---- flow ids: this chunk's slice of the random() stream, walked in op order -----------
		if (tid == 0) {
			int32_t rs[31]; int32_t rf = st->rf, rr = st->rr;
			for (int i = 0; i < 31; i++) rs[i] = st->rnd[i];
			ns_rand_jump(rs, rf, rr, st->burns+b_dr);
			for (int e = c_lo; e < c_hi; e++) {
				unsigned idp = ps.ev_idpos[e];
				if (idp == 0xffffffffu) continue;
				unsigned ni = ps.ev_nins[e];
				for (unsigned q = 0; q < idp; q++) ns_rand_step(rs, rf, rr);
				uint32_t id = ns_rand_step(rs, rf, rr);
				for (unsigned q = idp; q < ni+1; q++) ns_rand_step(rs, rf, rr);
				int f = S.ord2flow[e];
				S.f_id[f] = id;
				if (S.want_rec) {                                         // patch R_CONN_CREATE value / R_CONN_DST id (flow.c:333-334)
					uint8_t *rb = ps.st_rec+(u64)e*NS_STAGE_REC+ps.ev_nrec[e];
					uint8_t *cr = rb-34, *ds = rb-17;
					cr[13] = id >> 24; cr[14] = id >> 16; cr[15] = id >> 8; cr[16] = id;
					ds[2] = id >> 24; ds[3] = id >> 16; ds[4] = id >> 8; ds[5] = id;
				}
			}
		}
		__syncthreads();
		// ---- ordered commit of the staged event_adds and records -------------------------------
		u64 run_ins = tail0+b_ins, run_rec = rec0+b_rec;
		for (int e0 = c_lo; e0 < c_hi; e0 += PAR_THREADS) {
			int e = e0+tid;
			unsigned ni = 0, nr = 0;
			if (e < c_hi) { ni = ps.ev_nins[e]; nr = ps.ev_nrec[e]; }
			u64 xi = ni; unsigned xr = nr;
			for (int d = 1; d < 32; d <<= 1) {
				u64 yi = __shfl_up_sync(FULLMASK, xi, d); unsigned yr = __shfl_up_sync(FULLMASK, xr, d);
				if (lane >= d) { xi += yi; xr += yr; }
			}
			__syncthreads();
			if (lane == 31) { sh_u[warp] = xi; sh_r[warp] = xr; }
			if (tid == 0) sh_nbig = 0;
			__syncthreads();
			u64 wi = 0, tot_i = 0; unsigned wr = 0, tot_r = 0;
			for (int w = 0; w < PAR_WARPS; w++) {
				if (w < warp) { wi += sh_u[w]; wr += sh_r[w]; }
				tot_i += sh_u[w]; tot_r += sh_r[w];
			}
			u64 my_seq = run_ins+wi+xi-ni, my_rec = run_rec+wr+xr-nr;
			if (e < c_hi) {
				if (ni > 32) { int q = atomicAdd(&sh_nbig, 1); sh_big[q] = e; sh_bseq[q] = my_seq; sh_brec[q] = my_rec; }
				else par_commit_event(S, ps, e, my_seq, my_rec, 0, 0, 1);
			}
			__syncthreads();
			int nbig = sh_nbig;
			for (int q = warp; q < nbig; q += PAR_WARPS) par_commit_event(S, ps, sh_big[q], sh_bseq[q], sh_brec[q], 0, lane, 32);
			run_ins += tot_i; run_rec += tot_r;
		}
	}
	grid.sync();
	if (leader && st->status == 0) {
		int32_t rf = st->rf, rr = st->rr;
		ns_rand_jump(st->rnd, rf, rr, st->burns+t_dr);
		st->rf = rf; st->rr = rr;
		st->cnt[NS_CNT_DRAWS] += st->burns+t_dr; st->burns = 0;
		st->tail += t_ins; st->rec_len += S.want_rec ? t_rec : 0;
		st->n_ord = n0; st->hi_flow = n0; st->hi_range += n0;
		u64 cr = (u64)ctl->n_active;
		st->cnt[NS_CNT_CREATES] += cr; st->cnt[NS_CNT_LIVE_FLOWS] += cr;
		if (st->cnt[NS_CNT_LIVE_FLOWS] > st->cnt[NS_CNT_PEAK_FLOWS]) st->cnt[NS_CNT_PEAK_FLOWS] = st->cnt[NS_CNT_LIVE_FLOWS];
		st->cnt[NS_CNT_SOLVES] += ctl->lc[LC_SOLVES]; st->cnt[NS_CNT_FULL_SPREADS] += ctl->lc[LC_FULL_SPREADS];
		st->cnt[NS_CNT_FLOW_VISITS] += ctl->lc[LC_FLOW_VISITS]; st->cnt[NS_CNT_RETIMES] += ctl->lc[LC_RETIMES];
		st->cnt[NS_CNT_SEQ] = st->tail;
		for (int q = 0; q < LC_N; q++) ctl->lc[q] = 0;
		ctl->n_active = 0;
		*P.stage.ovf_top = 0;
	}
}

// host: wave schedule of group 0 (all CREATE): wave(op) = 1 + max wave of earlier ops sharing a node
static int par_plan_init(ParScratch &p, const ns_sc_op *ops, int n0, int n_nodes){
	std::vector<int> last(n_nodes, 0), wave(n0);
	int nwv = 0;
	for (int k = 0; k < n0; k++) {
		int w = 1+std::max(last[ops[k].a], last[ops[k].b]);
		wave[k] = w; last[ops[k].a] = w; last[ops[k].b] = w;
		if (w > nwv) nwv = w;
	}
	std::vector<int> off(nwv+2, 0), order(n0);
	for (int k = 0; k < n0; k++) off[wave[k]]++;                          // wave numbers start at 1
	for (int w = 1; w <= nwv; w++) off[w] += off[w-1];                    // off[w] = end of wave w
	std::vector<int> start(nwv+1, 0);
	for (int w = 1; w <= nwv; w++) start[w] = off[w-1];
	std::vector<int> cur(start);
	for (int k = 0; k < n0; k++) order[cur[wave[k]]++] = k;               // stable: op order inside a wave
	std::vector<int> woff(nwv+1);
	for (int w = 0; w <= nwv; w++) woff[w] = w == 0 ? 0 : off[w];
	p.n_waves = nwv;
	cudaError_t e;
	if ((e = cudaMemcpy(p.wave_off, woff.data(), (size_t)(nwv+1)*4, cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
	if ((e = cudaMemcpy(p.wave_ops, order.data(), (size_t)n0*4, cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
	return 0;
}

// ---------------------------------------------------------------- host launchers
static int par_grid(const void *kernel, int sm_count, int *grid){
	int per_sm = 0;
	cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, PAR_THREADS, 0);
	if (e != cudaSuccess) return (int)e;
	if (per_sm < 1) return NS_E_ARG;
	*grid = std::min(sm_count*per_sm, 4096);
	return 0;
}
static int par_init_group0(const Sim &S, const Sim *d_sim, ParScratch &p, int sm_count, cudaStream_t stream, uint64_t *launches){
	if (p.n_init > 0 && p.n_waves > 0) {
		int grid = 0, rc = par_grid((const void *)k_init_par, sm_count, &grid);
		if (rc) return rc;
		void *args[] = { (void *)&d_sim, (void *)&p };
		cudaError_t e = cudaLaunchCooperativeKernel((void *)k_init_par, dim3(grid), dim3(PAR_THREADS), args, 0, stream);
		if (e != cudaSuccess) return (int)e;
		(*launches)++;
		return 0;
	}
	k_init_groups<<<1, 32, 0, stream>>>(d_sim, 1);
	(*launches)++;
	return 0;
}
static int par_run(const Sim &S, const Sim *d_sim, const SeqScratch *d_scr, ParScratch &p, int sm_count, cudaStream_t stream, uint64_t *launches){
	int grid = 0, rc = par_grid((const void *)k_run_par, sm_count, &grid);
	if (rc) return rc;
	// The grid engine runs until it is done or asks for the single-warp engine (a sparse phase: many
	// timestamps with a handful of events each, where six grid barriers per timestamp cost more than
	// the events); that one hands back at the first timestamp that spans many ring tiles.
	for (int round = 0; round < 1 << 20; round++) {
		void *args[] = { (void *)&d_sim, (void *)&d_scr, (void *)&p };
		cudaError_t e = cudaLaunchCooperativeKernel((void *)k_run_par, dim3(grid), dim3(PAR_THREADS), args, 0, stream);
		if (e != cudaSuccess) return (int)e;
		(*launches)++;
		int flags[3];                                                     // sparse_mode, want_sparse, want_dense
		if ((e = cudaMemcpyAsync(flags, &S.st->sparse_mode, sizeof flags, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return (int)e;
		if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return (int)e;
		if (!flags[1]) return 0;
		int on[3] = { 1, 0, 0 };
		if ((e = cudaMemcpyAsync(&S.st->sparse_mode, on, sizeof on, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return (int)e;
		k_run_seq<<<1, 32, 0, stream>>>(d_sim, d_scr, 1);
		(*launches)++;
		if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
		if ((e = cudaMemcpyAsync(flags, &S.st->sparse_mode, sizeof flags, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return (int)e;
		if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return (int)e;
		int off[3] = { 0, 0, 0 };
		if ((e = cudaMemcpyAsync(&S.st->sparse_mode, off, sizeof off, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return (int)e;
		if (!flags[2]) return 0;                                          // it finished (or stopped) by itself
	}
	return 0;
}
