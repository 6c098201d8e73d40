// ns_par.cu — the grid-parallel engine's persistent kernel (k_run_par) and its host loop.
#include <vector>
#define NS_NO_WIDE      // one-event-per-lane runs live in k_run_seq / k_run_ens only: k_run_par's code layout is touchy (profiles/README.md r1-f)
#define NS_NO_CLIENTS   // ... and so do the behaviour-library hooks: simulations with client behaviour run k_run_par_cl (ns_par_cl.cu)
#include "ns_par_dev.cuh"

#define NS_KERNEL_DECL __global__ void __launch_bounds__(PAR_THREADS, PAR_MIN_BLOCKS) k_run_par(const Sim *sp, const SeqScratch *scrp, ParScratch P)
#define NS_KERNEL_PROLOGUE cgp::grid_group grid = cgp::this_grid();
#include "ns_par_kernel.inc"


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
	const void *kernel = S.has_clients ? par_kernel_cl() : (const void *)k_run_par;   // same engine, with / without the behaviour hooks
	{	// per device: the hub path keeps a node's flows in dynamic shared memory
		cudaError_t e0 = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HubSmem));
		if (e0 != cudaSuccess) return (int)e0;
	}
	int grid = 0, rc = par_grid(kernel, sm_count, &grid, sizeof(HubSmem));
	if (rc) return rc;
	// The grid engine runs until it is done or asks for the single-warp engine (a sparse phase: many
	// timestamps with a handful of events each, where six grid barriers per timestamp cost more than
	// the events); that one hands back at the first timestamp that spans many ring tiles.
	for (int round = 0; round < 1 << 20; round++) {
		void *args[] = { (void *)&d_sim, (void *)&d_scr, (void *)&p };
		cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(PAR_THREADS), args, sizeof(HubSmem), stream);
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
