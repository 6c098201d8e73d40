// ns_seq.cu — kernels of the single-warp engine (ns_seq.cuh) and their launchers: one warp per
// simulation, strictly sequential event handling.  Used for replica ensembles, module inits that are
// not pure flow creation, the host-callback mode and the sparse phases of the grid engine.
#define NS_PREFETCH      // one warp per simulation: request an event's lines up front (ns_model.cuh)
#define NS_HOST_CORE_OPS // k_exec_ops / k_run_seq serve the host core: resource deletion (node_del_resource, auto_delete in flow_close)
#include "ns_seq.cuh"
#include "ns_par_api.cuh"

// run the script's init group (module *_init: first flows and scheduled USER events)
__global__ void __launch_bounds__(32) k_init_groups(const Sim *sims, int n){
	int s = blockIdx.x;
	if (s >= n) return;
	Ctx c;
	ctx_load(c, &sims[s], threadIdx.x);
	run_group(c, 0);
	L0 c.st->cnt[NS_CNT_SEQ] = c.tail;
	ctx_flush(c);
}

// operations injected by a host-side module handler, at the current simulation time
__global__ void __launch_bounds__(32) k_exec_ops(const Sim *sp, int n_ops){
	Ctx c;
	ctx_load(c, sp, threadIdx.x);
	run_host_ops(c, n_ops);
	L0 c.st->cnt[NS_CNT_SEQ] = c.tail;
	ctx_flush(c);
	host_ops_status(c, n_ops);
}

// one warp per simulation, strictly sequential event handling
__global__ void __launch_bounds__(32) k_run_seq(const Sim *sims, const SeqScratch *scr, int n, int ens_phase, EnsAcc *acc, AnaStream *ana){
	int s = blockIdx.x;
	if (s >= n) return;
	if (ens_phase < 0) {                                                 // host-callback mode (ns_dev_run_host): a fresh run to the next stop
		if (threadIdx.x == 0) { sims[s].st->host_stop = 0; sims[s].st->max_events = 0; }
		__syncwarp();
		ens_phase = 0;
	}
	__shared__ WideSmem W;
	__shared__ FoldSmem Fs;
	EnsFold F;
	F.acc = acc ? acc+s : nullptr; F.ana = ana ? ana+s : nullptr;
	seq_run(&sims[s], scr[s], threadIdx.x, &W, ens_phase, acc ? &F : nullptr, &Fs);
}

static int launched(void){ cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? 0 : (int)e; }
int seq_launch_run(const Sim *d_sims, const SeqScratch *d_scr, int n, cudaStream_t stream, int ens_phase, EnsAcc *acc, AnaStream *ana){
	k_run_seq<<<n, 32, 0, stream>>>(d_sims, d_scr, n, ens_phase, acc, ana);
	return launched();
}
int seq_launch_init_groups(const Sim *d_sims, int n, cudaStream_t stream){
	k_init_groups<<<n, 32, 0, stream>>>(d_sims, n);
	return launched();
}
int seq_launch_exec_ops(const Sim *d_sim, int n_ops, cudaStream_t stream){
	k_exec_ops<<<1, 32, 0, stream>>>(d_sim, n_ops);
	return launched();
}
int seq_preload(void){
	cudaFuncAttributes fa;
	cudaError_t e = cudaFuncGetAttributes(&fa, k_run_seq);
	return e == cudaSuccess ? 0 : (int)e;
}
