// ns_par.cuh — grid-parallel handling of equal-timestamp SPEED_CHANGE batches (placeholder: falls
// through to the single-warp engine until the parallel kernels land).
#pragma once
#include "ns_seq.cuh"

struct ParScratch { int enabled; };

__global__ void k_init_groups(const Sim *sims, int n);
__global__ void k_run_seq(const Sim *sims, const SeqScratch *scr, int n);

template <typename A> static int par_alloc(ParScratch &p, const Sim &S, bool enable, A alloc){ p.enabled = enable; return 0; }
static int par_init_group0(const Sim &S, const Sim *d_sim, ParScratch &p, int sm_count, cudaStream_t stream, uint64_t *launches){
	k_init_groups<<<1, 32, 0, stream>>>(d_sim, 1);
	(*launches)++;
	return 0;
}
static int par_run(const Sim &S, const Sim *d_sim, const SeqScratch *d_scr, ParScratch &p, int sm_count, cudaStream_t stream, uint64_t *launches){
	k_run_seq<<<1, 32, 0, stream>>>(d_sim, d_scr, 1);
	(*launches)++;
	return 0;
}
