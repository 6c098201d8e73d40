// ns_init.cu — the parallel module init (k_init_par): a module's first sim_establish_flow calls in
// conflict-free waves.
#include <vector>
#include "ns_par_dev.cuh"

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
	if (S.want_rec && rec0+t_rec > S.rec_cap) { okc = false; if (leader) atomicCAS(&st->status, 0, NS_E_RECORDS_LOST); }
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
int par_plan_init(ParScratch &p, const ns_sc_op *ops, int n0, int n_nodes){
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

int par_init_group0(const Sim &S, const Sim *d_sim, ParScratch &p, int sm_count, cudaStream_t stream, uint64_t *launches){
	if (p.n_init > 0 && p.n_waves > 0) {
		int grid = 0, rc = par_grid((const void *)k_init_par, sm_count, &grid, 0);
		if (rc) return rc;
		void *args[] = { (void *)&d_sim, (void *)&p };
		cudaError_t e = cudaLaunchCooperativeKernel((void *)k_init_par, dim3(grid), dim3(PAR_THREADS), args, 0, stream);
		if (e != cudaSuccess) return (int)e;
		(*launches)++;
		return 0;
	}
	int rc = seq_launch_init_groups(d_sim, 1, stream);
	(*launches)++;
	return rc;
}
