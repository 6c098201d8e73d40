// ns_par_api.cuh — what the host layer (ns_engine.cu) sees of the two device engines: the scratch
// structs it allocates and the launchers.  The kernels live in their own translation units
// (ns_par.cu: k_run_par; ns_init.cu: k_init_par; ns_seq.cu: k_run_seq, k_init_groups, k_exec_ops) so
// that they compile in parallel and a change to one engine does not rebuild the others.
#pragma once
#include <cuda_runtime.h>
#include <string.h>
#include <algorithm>
#include "ns_sim.cuh"

#define PAR_THREADS 256
#ifndef PAR_MIN_BLOCKS
#define PAR_MIN_BLOCKS 1       // CTAs per SM the persistent kernel is compiled for (register budget)
#endif
#define PAR_WARPS (PAR_THREADS/32)
#ifndef PAR_MIN_BATCH
#define PAR_MIN_BATCH 48          // smaller batches are cheaper on the single-warp path
#endif
#define SPARSE_STREAK 6            // that many small batches in a row: the grid barriers cost more than the events
#define SPARSE_MAX_TILES 4096     // ... provided one warp can scan the tile minima of both event stores

struct ParScratch {
	int enabled;
	int seg_cap;                  // events per parallel segment
	ParStage stage;               // staging arrays (device pointers)
	ParStage *d_stage;
	// per-timestamp
	int *tile_cnt, *tile_off;     // per ring-window tile: matches, exclusive offset
	unsigned *tile_mask; int *cta_tile;   // ... the match bits of a tile (NS_TILE_ROWS words), per-CTA totals of the count scan
	// grouping
	int *grp_cnt, *grp_off, *grp_fill, *active, *grp_items, *b_grp;
	int *deferred, *defer_pos;     // groups handed from the per-thread pass to the per-warp pass
	int *defer_long;               // ... the ones with long chains (taken first)
	int *srt_key, *srt_key2, *srt_val2, *rad_hist;   // stable partition of a segment by group (radix_pass)
	int *big_e; u64 *big_seq; int big_cap;           // ... their segment index and first sequence number
	int *pk_fd, *pk_own; double *pk_speed; u64 *pk_slot;   // the events in partition order: flow*2+dir, list position of the flow, speed, ring slot
	// parallel module init: CREATE ops of group 0 scheduled into conflict-free waves (host, par_plan_init)
	int n_init, n_waves; int *wave_off, *wave_ops; u64 *cta_draws;
	// scan
	u64 *cta_ins; unsigned *cta_rec, *cta_tr;
	// control block (device)
	struct Ctl {
		u64 tmin_bits, first_live;
		int nb, nl, n_active, cursor, work, seg_lo, seg_hi, n_deferred, dbg_g, dbg_deg;
		int n_long;                                // chains of more than 128 events: taken first by the warp pass (defer_long)
		int n_big;                                 // events with more than 32 event_adds: committed by the whole grid (F3)
		int n_hub, hub_work;                       // groups handed to whole CTAs (hub_chain_cta); stored from the END of `deferred`
		int piece, last, pos, li, l_lo, l_hi;      // the current piece of the batch (par_plan_piece)
		int small_streak;                          // consecutive small batches (sparse phase detection)
		u64 dbg_key, dbg_key2, dbg_hub, dbg_hub_end, dbg_hub_sum;            // profiling builds: slowest chain / table-driven group / hub CTA of the segment
		u64 tot_ins; unsigned tot_rec, tot_tr;
		u64 lc[LC_N];
	} *ctl;
};

// sizes of the scratch: the grid engine of one big simulation (defaults) or one CTA per replica of an ensemble (small)
struct ParSizes { u64 max_seg = 1u << 20, min_ovf = 1u << 22, ovf_per_flow = 64; int grid_max = 4096; bool dry = false; };
template <typename A> static int par_alloc(ParScratch &p, const Sim &S, bool enable, int n_init, A alloc, const ParSizes &z = ParSizes()){
	memset(&p, 0, sizeof p);
	p.enabled = enable;
	if (!enable) return 0;
	int rc;
	p.n_init = n_init;
	p.seg_cap = std::max((int)std::min<u64>((u64)S.batch_cap, z.max_seg), n_init);
	ParStage &st = p.stage;
	st.inline_n = (u64)NS_STAGE_INLINE*p.seg_cap;
	st.ovf_cap = std::max<u64>(z.min_ovf, z.ovf_per_flow*(u64)S.cap_flows);
	u64 n = st.inline_n+st.ovf_cap;
#define PA(ptr, cnt) if ((rc = alloc((void **)&(ptr), (size_t)(cnt)*sizeof(*(ptr))))) return rc
	PA(st.st_time, n); PA(st.st_speed, n); PA(st.st_fd, n); PA(st.ovf_top, 1);
	PA(st.st_rec, (u64)NS_STAGE_REC*p.seg_cap); PA(st.st_tr, p.seg_cap);
	PA(st.ev_nins, p.seg_cap); PA(st.ev_ovf, p.seg_cap); PA(st.ev_nrec, p.seg_cap); PA(st.ev_ntr, p.seg_cap); PA(st.ev_idpos, p.seg_cap);
	PA(p.d_stage, 1);
	PA(p.tile_cnt, S.n_ring_tiles+1); PA(p.tile_off, S.n_ring_tiles+1);
	PA(p.tile_mask, (size_t)(S.n_ring_tiles+1)*NS_TILE_ROWS); PA(p.cta_tile, z.grid_max);
	PA(p.grp_cnt, 2*S.n_nodes+1); PA(p.grp_off, 2*S.n_nodes+1); PA(p.grp_fill, 2*S.n_nodes+1);
	PA(p.active, p.seg_cap); PA(p.grp_items, p.seg_cap); PA(p.b_grp, p.seg_cap);
	PA(p.deferred, p.seg_cap); PA(p.defer_pos, 2*S.n_nodes+1); PA(p.defer_long, p.seg_cap/128+2);
	PA(p.srt_key, p.seg_cap); PA(p.srt_key2, p.seg_cap); PA(p.srt_val2, p.seg_cap); PA(p.rad_hist, 512*z.grid_max);
	p.big_cap = (int)std::min<u64>(1u << 18, (u64)p.seg_cap); PA(p.big_e, p.big_cap); PA(p.big_seq, p.big_cap);
	PA(p.pk_fd, p.seg_cap); PA(p.pk_own, p.seg_cap); PA(p.pk_speed, p.seg_cap); PA(p.pk_slot, p.seg_cap);
	PA(p.cta_ins, z.grid_max); PA(p.cta_rec, z.grid_max); PA(p.cta_tr, z.grid_max); PA(p.cta_draws, z.grid_max);
	if (n_init) { PA(p.wave_off, n_init+2); PA(p.wave_ops, n_init); }
	PA(p.ctl, 1);
#undef PA
	if (z.dry) return 0;                                              // measuring pass of an arena (ns_engine.cu)
	cudaError_t e;
	if ((e = cudaMemcpy(p.d_stage, &p.stage, sizeof p.stage, cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
	if ((e = cudaMemset(p.grp_cnt, 0, (size_t)(2*S.n_nodes+1)*4)) != cudaSuccess) return (int)e;
	if ((e = cudaMemset(p.ctl, 0, sizeof *p.ctl)) != cudaSuccess) return (int)e;
	if ((e = cudaMemset(st.ovf_top, 0, 8)) != cudaSuccess) return (int)e;
	return 0;
}


// ---- launchers (each defined next to its kernel)
// ns_seq.cu
int seq_launch_run(const Sim *d_sims, const SeqScratch *d_scr, int n, cudaStream_t stream, int ens_phase = 0, EnsAcc *acc = nullptr, AnaStream *ana = nullptr);   // ens_phase: only the replicas in their sparse phase; acc: ensemble replicas fold their records in place
int seq_launch_init_groups(const Sim *d_sims, int n, cudaStream_t stream);
int seq_launch_exec_ops(const Sim *d_sim, int n_ops, cudaStream_t stream);
int seq_preload(void);
// ns_init.cu
int par_plan_init(ParScratch &p, const ns_sc_op *ops, int n0, int n_nodes);
int par_init_group0(const Sim &S, const Sim *d_sim, ParScratch &p, int sm_count, cudaStream_t stream, uint64_t *launches);
// ns_ens.cu: one CTA per replica (the batch engine at CTA scope)
int ens_launch_run(const Sim *d_sims, const SeqScratch *d_scr, const ParScratch *d_pars, int n, cudaStream_t stream, EnsAcc *acc, AnaStream *ana);
int ens_preload(void);
// ns_par_cl.cu: k_run_par with the behaviour-library hooks
const void *par_kernel_cl(void);
// ns_par.cu
int par_run(const Sim &S, const Sim *d_sim, const SeqScratch *d_scr, ParScratch &p, int sm_count, cudaStream_t stream, uint64_t *launches);
int par_preload(void);
int par_grid(const void *kernel, int sm_count, int *grid, size_t dyn_smem);
