// ns_par.cu — the grid-parallel engine's persistent kernel (k_run_par) and its host loop.
#include <vector>
#include "ns_par_dev.cuh"

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
#ifdef NS_PROFILE
	{ u64 k_ = ctl->dbg_key; st->phase_ns[13] += k_ >> 20; if ((k_ >> 20) > st->phase_ns[10]) { st->phase_ns[10] = k_ >> 20; st->phase_ns[11] = ctl->dbg_deg; st->phase_ns[12] = k_ & 0xfffff; } ctl->dbg_key = 0; st->phase_ns[14] += ctl->n_deferred; }
#endif
	st->cnt[NS_CNT_SEQ] = st->tail;
	for (int q = 0; q < LC_N; q++) ctl->lc[q] = 0;
	ctl->pos = ctl->seg_hi;
	ctl->n_active = 0; ctl->cursor = 0; ctl->work = 0; ctl->n_deferred = 0; ctl->n_hub = 0; ctl->hub_work = 0; ctl->n_long = 0; ctl->n_big = 0;
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
		u64 ops = 0;                                                  // scripted creates / closes the long-lived events of the batch will run
		for (int k = li; k < nl; k++) ops += (u64)l_entry_ops(S, sc.lb_ref[k]);
		if (st->rec_len+next*NS_REC_PER_EVENT+ops*NS_REC_PER_OP+reserve > S.rec_cap) { st->host_stop = 2; return; }
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
	__shared__ int sh_status, sh_nbtot;
	__shared__ int sh_ctab[1024];                                 // per-CTA tile-count prefixes (grid size <= 1024 CTAs)
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int gw = blockIdx.x*PAR_WARPS+warp, nw = gridDim.x*PAR_WARPS;
	const u64 gtid = (u64)blockIdx.x*PAR_THREADS+tid, gth = (u64)gridDim.x*PAR_THREADS;
	const u64 tmask = (u64)S.n_ring_tiles-1, smask = S.ring_cap-1;
	const bool leader = blockIdx.x == 0 && tid == 0;
	extern __shared__ __align__(16) unsigned char dyn_smem[];
	HubSmem &Hs = *reinterpret_cast<HubSmem *>(dyn_smem);
	RadSmem &Rs = *reinterpret_cast<RadSmem *>(dyn_smem);       // the partition phase and the hub phase never overlap
	const bool hub_on = (S.features & NS_FEAT_HUBCTA) != 0;
#ifdef NS_PROFILE
	u64 t_ph = gtimer();
	const u64 t_kernel0 = t_ph;
#endif

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
		// ---- B1: ONE sweep over the tiles that hit: match count, match bit mask (32 words per tile) and the tile's
		//      new minimum; the last warp also extracts the long-lived events ------------------------------------
		for (u64 k = gw; k < n_wt; k += nw) {
			const u64 at = t_first+k;
			int cnt = 0;
			if (S.q_tmin[at & tmask] == m) {
				const u64 base = at << NS_TILE_SHIFT;
				double nm = NS_INF;
				unsigned my_word = 0;                                     // lane i keeps the ballot of row i
				for (int i0 = 0; i0 < NS_TILE; i0 += 32*8) {                 // 8 independent loads in flight per lane
					double tv[8];
					#pragma unroll
					for (int u = 0; u < 8; u++) tv[u] = S.q_time[(base+i0+32*u+lane) & smask];
					#pragma unroll
					for (int u = 0; u < 8; u++) {
						const u64 s = base+i0+32*u+lane;
						const double t = tv[u];
						const bool match = t == m && s >= head && s < tail;
						const unsigned mm = __ballot_sync(FULLMASK, match);
						if (lane == (i0 >> 5)+u) my_word = mm;
						cnt += __popc(mm);
						if (!match) nm = fmin(nm, t);
					}
				}
				nm = warp_min_d(nm);
				P.tile_mask[k*32+lane] = my_word;
				if (lane == 0) S.q_tmin[at & tmask] = nm;                 // nobody else looks at this tile's minimum before the next argmin
			}
			if (lane == 0) P.tile_cnt[k] = cnt;
		}
		if (gw == nw-1) {
			Ctx c; ctx_load(c, sp, lane);
			int nl = seq_collect_l(c, m, sc);
			if (lane == 0) ctl->nl = nl;
		}
		grid.sync();
		// ---- B2: exclusive scan of the tile counts: every CTA scans a contiguous share, the CTA totals are
		//      combined by everyone after the barrier -----------------------------------------------------------
		const u64 t_per = (n_wt+gridDim.x-1)/gridDim.x;
		{
			const u64 lo = (u64)blockIdx.x*t_per < n_wt ? (u64)blockIdx.x*t_per : n_wt, hi = lo+t_per < n_wt ? lo+t_per : n_wt;
			int carry = 0;
			for (u64 k0 = lo; k0 < hi; k0 += PAR_THREADS) {
				const u64 k = k0+tid;
				const int v = k < hi ? P.tile_cnt[k] : 0;
				int incl = v;
				#pragma unroll
				for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(FULLMASK, incl, d); if (lane >= d) incl += t; }
				__syncthreads();
				if (lane == 31) sh_r[warp] = (unsigned)incl;
				__syncthreads();
				int wb = 0, tot = 0;
				for (int w = 0; w < PAR_WARPS; w++) { if (w < warp) wb += (int)sh_r[w]; tot += (int)sh_r[w]; }
				if (k < hi) P.tile_off[k] = carry+wb+incl-v;
				carry += tot;
			}
			if (tid == 0) P.cta_tile[blockIdx.x] = carry;
		}
		grid.sync();
		{
			int acc = 0;                                                  // exclusive prefix of the CTA totals, in shared memory
			for (int q = tid; q < (int)gridDim.x; q += PAR_THREADS) sh_ctab[q] = P.cta_tile[q];
			__syncthreads();
			if (tid == 0) { for (int q = 0; q < (int)gridDim.x; q++) { int v = sh_ctab[q]; sh_ctab[q] = acc; acc += v; } sh_nbtot = acc; }
			__syncthreads();
		}
		const int nb = sh_nbtot, nl = ctl->nl;
		if (nb > S.batch_cap) { if (leader) atomicCAS(&st->status, 0, NS_E_BATCH_FULL); grid.sync(); continue; }
		// ---- B3: write the batch in order from the bit masks (no second sweep over the ring) -----------------------
		for (u64 k = gw; k < n_wt; k += nw) {
			if (P.tile_cnt[k] == 0) continue;
			const u64 base = (t_first+k) << NS_TILE_SHIFT;
			int off = P.tile_off[k]+sh_ctab[k/t_per];
			const unsigned my_word = P.tile_mask[k*32+lane];
			#pragma unroll 8
			for (int i = 0; i < 32; i++) {
				const unsigned mm = __shfl_sync(FULLMASK, my_word, i);
				if ((mm >> lane) & 1u) S.b_seq[off+__popc(mm & ((1u << lane)-1))] = base+(u64)(i*32+lane);
				off += __popc(mm);
			}
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
			ctl->n_active = 0; ctl->cursor = 0; ctl->work = 0; ctl->n_deferred = 0; ctl->n_hub = 0; ctl->hub_work = 0; ctl->n_long = 0; ctl->n_big = 0;
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
			// ---- D1: key of every event = the (node,dir) group it rewrites; cancelled events get the sentinel ----
			const int NG = 2*S.n_nodes;
			for (u64 e = gtid; e < (u64)ns; e += gth) {
				u64 s = S.b_seq[seg_lo+e];
				u64 slot = s & smask;
				int fd = S.q_fd[slot];
				int f = fd >> 1, dir = fd & 1;
				int g = NG;
				if (S.f_gen[f] == S.q_gen[slot] && S.f_alive[f]) g = S.f_peer[dir][f]*2+dir;
				else {
					S.q_time[slot] = NS_INF;                          // flow_close cancelled it (flow.c:260-270)
					P.stage.ev_nins[e] = 0; P.stage.ev_nrec[e] = 0; P.stage.ev_ntr[e] = 0; P.stage.ev_ovf[e] = -1;
				}
				P.b_grp[e] = g;
			}
			grid.sync();
			// ---- D2: STABLE partition of the segment by group (LSD radix, RAD_BITS per pass): afterwards the events
			//      of a group are contiguous in grp_items AND in batch order, so no chain ever needs sorting
			{
				int kb = 1; while ((NG >> kb) != 0) kb++;
				const int np = (kb+RAD_BITS-1)/RAD_BITS;
				const int *sk = P.b_grp, *sv = nullptr;
				for (int i = 0; i < np; i++) {
					const bool fin = ((np-1-i) & 1) == 0;
					int *dk = fin ? P.srt_key : P.srt_key2, *dv = fin ? P.grp_items : P.srt_val2;
					radix_pass(grid, Rs, sk, sv, dk, dv, ns, i*RAD_BITS, P.rad_hist);
					sk = dk; sv = dv;
				}
			}
			// ---- D3: group boundaries ------------------------------------------------------------------
			for (u64 i = gtid; i < (u64)ns; i += gth) {
				const int k = P.srt_key[i];
				if (k >= NG) continue;
				if (i == 0 || P.srt_key[i-1] != k) { P.grp_off[k] = (int)i; P.active[atomicAdd(&ctl->n_active, 1)] = k; }
				if (i+1 == (u64)ns || P.srt_key[i+1] != k) P.grp_fill[k] = (int)i+1;
				// the event itself, packed in partition order: the chains read these with coalesced loads
				const u64 slot = S.b_seq[seg_lo+P.grp_items[i]] & smask;
				const int fd = S.q_fd[slot];
				P.pk_fd[i] = fd; P.pk_speed[i] = S.q_speed[slot]; P.pk_slot[i] = slot; P.pk_own[i] = S.f_pos[fd & 1][fd >> 1];
			}
			grid.sync();
			const int n_active = ctl->n_active;
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
					const int off = P.grp_off[g], cnt = P.grp_fill[g]-off;
					P.grp_cnt[g] = cnt;
					int j = 0;
					if (cnt <= 32) {                                  // long chains go straight to a warp / a CTA
						const int *it = P.grp_items+off;
						for (; j < cnt; j++) {
							int e = it[j];
							u64 slot = S.b_seq[seg_lo+e] & smask;
							int fd = S.q_fd[slot];
							if (sc_needs_full_spread(S, fd >> 1, fd & 1, S.q_speed[slot])) break;
							par_handle_event(c, P, seg_lo, e, m);
						}
					}
					if (j < cnt) {
						P.defer_pos[g] = j;
						// a large sender group goes to a whole CTA (hub_chain_cta), everything else to one warp — the long
						// chains first, since the segment ends with its slowest group
						if (hub_on && (g & 1) == SND && S.adj_live[g] >= HUB_MIN_DEG && S.adj_len[g] <= HUB_CAP) P.deferred[P.seg_cap-1-atomicAdd(&ctl->n_hub, 1)] = g;
						else if (cnt-j > 128) P.defer_long[atomicAdd(&ctl->n_long, 1)] = g;
						else P.deferred[atomicAdd(&ctl->n_deferred, 1)] = g;
					}
					else P.grp_cnt[g] = 0;                            // ready for the next segment
				}
				for (int q = 0; q < LC_N; q++) if (c.lc[q]) atomicAdd((unsigned long long *)&ctl->lc[q], c.lc[q]);
			}
			grid.sync();
			PHASE_MARK(PH_SOLVE1);
			// ---- E3: one CTA per large sender group, its flows resident in shared memory ------------------
			// (before E2 in program order, concurrent with it in time: CTAs without a hub go straight on)
			if (ctl->n_hub) {
				const int n_hub = ctl->n_hub;
				u64 hlc[LC_N];
				for (int q = 0; q < LC_N; q++) hlc[q] = 0;
				while (true) {
					__syncthreads();
					if (tid == 0) Hs.work = atomicAdd(&ctl->hub_work, 1);
					__syncthreads();
					const int k = Hs.work;
					if (k >= n_hub) break;
					hub_chain_cta(S, st, P, Hs, P.deferred[P.seg_cap-1-k], seg_lo, m, hlc);
				}
				if (tid == 0) for (int q = 0; q < LC_N; q++) if (hlc[q]) atomicAdd((unsigned long long *)&ctl->lc[q], hlc[q]);
			}
			// ---- E2: one WARP per remaining group: full spreads strided over the lanes ----------------
			{
				const int n_lg = ctl->n_long, n_def = n_lg+ctl->n_deferred;
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
					int g = k < n_lg ? P.defer_long[k] : P.deferred[k-n_lg];
					int cnt = P.grp_cnt[g], off = P.grp_off[g];
					int j = P.defer_pos[g];
#ifdef NS_PROFILE
					u64 t_g0 = gtimer();
#endif
					if (j < 0) { warp_sort_global(P.grp_items+off, cnt, lane); j = 0; }
#ifdef NS_PROFILE
					int j0 = j;
#endif
					if (cnt-j >= GE_MIN_CHAIN && S.adj_len[g] <= 32) { if (S.features & NS_FEAT_CHAIN2) ge_chain2(c, P, g, seg_lo, off, j, cnt, m); else ge_chain(c, P, g, seg_lo, off, j, cnt, m); }
					else for (; j < cnt; j++) par_handle_event(c, P, seg_lo, P.grp_items[off+j], m);
					if (lane == 0) {
						P.grp_cnt[g] = 0;                             // ready for the next segment
#ifdef NS_PROFILE
						u64 dt = gtimer()-t_g0;                       // the slowest group of the segment
						u64 key = (dt << 20) | (u64)(cnt-j0 < 0xfffff ? cnt-j0 : 0xfffff);
						if (atomicMax((unsigned long long *)&ctl->dbg_key, key) < key) { ctl->dbg_g = g; ctl->dbg_deg = S.adj_live[g]; }
#endif
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
				if (S.want_rec && rec0+t_rec > S.rec_cap) { okc = false; if (leader) atomicCAS(&st->status, 0, NS_E_RECORDS_LOST); }
				if (leader) { ctl->tot_ins = t_ins; ctl->tot_rec = t_rec; ctl->tot_tr = t_tr; }
				if (okc) {
					// block-wide running scan over the chunk, PAR_THREADS events at a time.  Small events (<= 32
					// event_adds) are committed by their own thread with the loads batched; the records of the 256
					// events are laid out contiguously in shared memory and leave as one coalesced copy; the
					// event_adds of big events (full spreads) wait for F3, where the whole grid shares them.
					CommitSmem &Cs = *reinterpret_cast<CommitSmem *>(dyn_smem);
					u64 run_ins = tail0+b_ins, run_rec = rec0+b_rec, run_tr = tr0+b_tr;
					if (tid < COMMIT_TILES) Cs.tmin[tid] = ~0ull;
					for (int e0 = c_lo; e0 < c_hi; e0 += PAR_THREADS) {
						const int e = e0+tid;
						unsigned ni = 0, nr = 0, nt = 0;
						if (e < c_hi) { ni = P.stage.ev_nins[e]; nr = P.stage.ev_nrec[e]; nt = P.stage.ev_ntr[e]; }
						u64 xi = ni; unsigned xr = nr, xt = nt;
						for (int d = 1; d < 32; d <<= 1) {
							u64 yi = __shfl_up_sync(FULLMASK, xi, d); unsigned yr = __shfl_up_sync(FULLMASK, xr, d), yt = __shfl_up_sync(FULLMASK, xt, d);
							if (lane >= d) { xi += yi; xr += yr; xt += yt; }
						}
						__syncthreads();
						if (lane == 31) { sh_u[warp] = xi; sh_r[warp] = xr; sh_t[warp] = xt; }
						__syncthreads();
						u64 wi = 0; unsigned wr = 0, wt = 0;
						u64 tot_i = 0; unsigned tot_r = 0, tot_t = 0;
						for (int w = 0; w < PAR_WARPS; w++) {
							if (w < warp) { wi += sh_u[w]; wr += sh_r[w]; wt += sh_t[w]; }
							tot_i += sh_u[w]; tot_r += sh_r[w]; tot_t += sh_t[w];
						}
						const u64 my_seq = run_ins+wi+xi-ni, my_rec = run_rec+wr+xr-nr, my_tr = run_tr+wt+xt-nt;
						const unsigned pad = (unsigned)((uintptr_t)(S.rec+run_rec) & 15u);     // shared copy keeps the destination's alignment
						if (e < c_hi) {
							if (ni > 32) {                            // a full spread: its event_adds go to F3
								int q = atomicAdd(&ctl->n_big, 1);
								if (q < P.big_cap) { P.big_e[q] = e; P.big_seq[q] = my_seq; }
								else atomicCAS(&st->status, 0, NS_E_INTERNAL);
							} else if (ni) commit_small_inserts(S, P.stage, Cs, e, ni, my_seq, run_ins);
							if (nr) {                                 // record bytes -> shared memory at their final relative offset
								const u64 *src = (const u64 *)(P.stage.st_rec+(u64)e*NS_STAGE_REC);
								unsigned char *dstb = Cs.rec+pad+(unsigned)(my_rec-run_rec);
								#pragma unroll
								for (int w = 0; w < NS_STAGE_REC/8; w++) {
									if (8u*w < nr) {
										const u64 x = src[w];
										#pragma unroll
										for (int bb = 0; bb < 8; bb++) if (8u*w+bb < nr) dstb[8*w+bb] = (unsigned char)(x >> (8*bb));
									}
								}
							}
							if (nt && my_tr < S.tr_cap) S.tr[my_tr] = P.stage.st_tr[e];
						}
						__syncthreads();
						if (tot_r) {                                  // one coalesced copy of the iteration's records
							unsigned char *dst = S.rec+run_rec;
							const unsigned n = tot_r, head = n < ((16u-pad) & 15u) ? n : ((16u-pad) & 15u);
							for (unsigned i = tid; i < head; i += PAR_THREADS) dst[i] = Cs.rec[pad+i];
							const unsigned nvec = (n-head) >> 4;
							for (unsigned v = tid; v < nvec; v += PAR_THREADS) *(uint4 *)(dst+head+16u*v) = *(const uint4 *)(Cs.rec+pad+head+16u*v);
							for (unsigned i = head+16u*nvec+tid; i < n; i += PAR_THREADS) dst[i] = Cs.rec[pad+i];
						}
						if (tid < COMMIT_TILES) {                     // one tile-minimum update per touched ring tile
							const u64 tb = Cs.tmin[tid];
							if (tb != ~0ull) {
								const double t = __longlong_as_double((long long)tb);
								double *tm = &S.q_tmin[((run_ins >> NS_TILE_SHIFT)+tid) & tmask];
								if (*(volatile double *)tm > t) tmin_lower(tm, t);
								Cs.tmin[tid] = ~0ull;
							}
						}
						run_ins += tot_i; run_rec += tot_r; run_tr += tot_t;
					}
				}
			}
			grid.sync();
			// ---- F3: the event_adds of the big events, one warp per event, over the whole grid -------------------
			{
				const int n_big = ctl->n_big < P.big_cap ? ctl->n_big : P.big_cap;
				for (int q = gw; q < n_big; q += nw) par_commit_event(S, P.stage, P.big_e[q], P.big_seq[q], 0, 0, lane, 32, false);
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
#ifdef NS_PROFILE
	if (leader) { u64 t_ = gtimer(); st->phase_ns[PH_TAIL] += t_-t_ph; st->phase_ns[PH_KERNEL] += t_-t_kernel0; }
#endif
}

// ---------------------------------------------------------------- host launchers
int par_grid(const void *kernel, int sm_count, int *grid, size_t dyn_smem){
	int per_sm = 0;
	cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, PAR_THREADS, dyn_smem);
	if (e != cudaSuccess) return (int)e;
	if (per_sm < 1) return NS_E_ARG;
	*grid = std::min(sm_count*per_sm, 1024);                                 // k_run_par keeps per-CTA tables of 1024 entries
	return 0;
}
int par_run(const Sim &S, const Sim *d_sim, const SeqScratch *d_scr, ParScratch &p, int sm_count, cudaStream_t stream, uint64_t *launches){
	{	// per device: the hub path keeps a node's flows in dynamic shared memory
		cudaError_t e0 = cudaFuncSetAttribute(k_run_par, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HubSmem));
		if (e0 != cudaSuccess) return (int)e0;
	}
	int grid = 0, rc = par_grid((const void *)k_run_par, sm_count, &grid, sizeof(HubSmem));
	if (rc) return rc;
	// The grid engine runs until it is done or asks for the single-warp engine (a sparse phase: many
	// timestamps with a handful of events each, where six grid barriers per timestamp cost more than
	// the events); that one hands back at the first timestamp that spans many ring tiles.
	for (int round = 0; round < 1 << 20; round++) {
		void *args[] = { (void *)&d_sim, (void *)&d_scr, (void *)&p };
		cudaError_t e = cudaLaunchCooperativeKernel((void *)k_run_par, dim3(grid), dim3(PAR_THREADS), args, sizeof(HubSmem), stream);
		if (e != cudaSuccess) return (int)e;
		(*launches)++;
		int flags[3];                                                     // sparse_mode, want_sparse, want_dense
		if ((e = cudaMemcpyAsync(flags, &S.st->sparse_mode, sizeof flags, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return (int)e;
		if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return (int)e;
		if (!flags[1]) return 0;
		int on[3] = { 1, 0, 0 };
		if ((e = cudaMemcpyAsync(&S.st->sparse_mode, on, sizeof on, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return (int)e;
		{ int rc2 = seq_launch_run(d_sim, d_scr, 1, stream); if (rc2) return rc2; }
		(*launches)++;
		if ((e = cudaMemcpyAsync(flags, &S.st->sparse_mode, sizeof flags, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return (int)e;
		if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return (int)e;
		int off[3] = { 0, 0, 0 };
		if ((e = cudaMemcpyAsync(&S.st->sparse_mode, off, sizeof off, cudaMemcpyHostToDevice, stream)) != cudaSuccess) return (int)e;
		if (!flags[2]) return 0;                                          // it finished (or stopped) by itself
	}
	return 0;
}

int par_preload(void){
	cudaFuncAttributes fa;
	cudaError_t e = cudaFuncGetAttributes(&fa, k_run_par);
	return e == cudaSuccess ? 0 : (int)e;
}
