// ns_ens.cu — replica ensembles: the batch engine of ns_par_kernel.inc at CTA scope.  One CTA steps one
// replica (argmin, ordered extraction, (node,dir) grouping, per-thread / per-warp / whole-CTA solves, ordered
// commit), its barriers are __syncthreads, and a launch covers every replica of the ensemble: the CTAs that
// do not fit wait on the hardware block scheduler, which is the work queue of the replicas (run
// times vary several-fold by seed).  Short runs of events go through the one-event-per-lane path of ns_seq.cuh.
#define NS_SCOPE_CTA
#define PAR_MIN_BATCH 33           // up to one warp's worth of events: the one-event-per-lane path; above: the whole CTA
#include "ns_par_dev.cuh"

#define NS_KERNEL_DECL __global__ void __launch_bounds__(PAR_THREADS, 1) k_run_ens(const Sim *sims, const SeqScratch *scrs, const ParScratch *pars, int n_sims)
#define NS_KERNEL_PROLOGUE \
	CtaScope grid; \
	if ((int)blockIdx.x >= n_sims) return; \
	const Sim *sp = sims+blockIdx.x; const SeqScratch *scrp = scrs+blockIdx.x; const ParScratch &P = pars[blockIdx.x];
#include "ns_par_kernel.inc"

int ens_preload(void){
	cudaError_t e = cudaFuncSetAttribute(k_run_ens, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HubSmem));
	return e == cudaSuccess ? 0 : (int)e;
}
int ens_launch_run(const Sim *d_sims, const SeqScratch *d_scr, const ParScratch *d_pars, int n, cudaStream_t stream){
	k_run_ens<<<n, PAR_THREADS, sizeof(HubSmem), stream>>>(d_sims, d_scr, d_pars, n);
	cudaError_t e = cudaGetLastError();
	return e == cudaSuccess ? 0 : (int)e;
}
