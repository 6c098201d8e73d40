// ns_ens.cu — replica ensembles: the batch engine of ns_par_kernel.inc at CTA scope.  One CTA steps one
// replica (argmin, ordered extraction, (node,dir) grouping, per-thread / per-warp / whole-CTA solves, ordered
// commit), its barriers are __syncthreads, and a launch covers every replica of the ensemble: the CTAs that
// do not fit wait on the hardware block scheduler, which is the work queue of the replicas (run
// times vary several-fold by seed).  Short runs of events go through the one-event-per-lane path of ns_seq.cuh.
// A replica that enters a sparse phase (a streak of tiny batches: a swarm draining to completion is ten thousands
// of single-event timestamps) leaves the CTA engine: one SM holds one such CTA but dozens of single warps, so the
// host loop (ns_ens_run) alternates this kernel with k_run_seq over the replicas that asked for it.
#define NS_SCOPE_CTA
#define NS_PREFETCH      // one CTA per simulation: request an event's lines up front (ns_model.cuh)
#define PAR_MIN_BATCH 33           // up to one warp's worth of events: the one-event-per-lane path; above: the whole CTA
#include "ns_par_dev.cuh"

#define NS_KERNEL_DECL __global__ void __launch_bounds__(PAR_THREADS, 1) k_run_ens(const Sim *sims, const SeqScratch *scrs, const ParScratch *pars, int n_sims, EnsAcc *ens_acc, AnaStream *ens_ana)
#define NS_KERNEL_PROLOGUE \
	CtaScope grid; \
	if ((int)blockIdx.x >= n_sims) return; \
	const Sim *sp = sims+blockIdx.x; const SeqScratch *scrp = scrs+blockIdx.x; const ParScratch &P = pars[blockIdx.x]; \
	if (sp->st->sparse_mode) return;          /* in its sparse phase: stepped by its warp of k_run_seq */
#include "ns_par_kernel.inc"

int ens_preload(void){
	cudaError_t e = cudaFuncSetAttribute(k_run_ens, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HubSmem));
	return e == cudaSuccess ? 0 : (int)e;
}
int ens_launch_run(const Sim *d_sims, const SeqScratch *d_scr, const ParScratch *d_pars, int n, cudaStream_t stream, EnsAcc *acc, AnaStream *ana){
	k_run_ens<<<n, PAR_THREADS, sizeof(HubSmem), stream>>>(d_sims, d_scr, d_pars, n, acc, ana);
	cudaError_t e = cudaGetLastError();
	return e == cudaSuccess ? 0 : (int)e;
}
