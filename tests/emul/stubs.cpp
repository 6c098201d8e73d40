// tests/emul/stubs.cpp — TEST INFRASTRUCTURE: the grid-scope batch engine (ns_par.cu, ns_par_cl.cu, ns_init.cu: cooperative
// kernels of 148 x 256 threads with grid barriers) is outside the lock-step model; asking for it is an error.  Single simulations
// run NS_MODE_SEQ (one warp); ensembles also run the batch engine at CTA scope (ns_ens.cu: the same ns_par_kernel.inc, one CTA
// per replica, __syncthreads barriers).
#include "ns_par_api.cuh"
#define NOT_EMULATED NS_E_NO_DEVICE
int par_plan_init(ParScratch &, const ns_sc_op *, int, int){ return NOT_EMULATED; }
int par_init_group0(const Sim &, const Sim *, ParScratch &, int, cudaStream_t, uint64_t *){ return NOT_EMULATED; }
const void *par_kernel_cl(void){ return nullptr; }
int par_run(const Sim &, const Sim *, const SeqScratch *, ParScratch &, int, cudaStream_t, uint64_t *){ return NOT_EMULATED; }
int par_preload(void){ return NOT_EMULATED; }
int par_grid(const void *, int, int *, size_t){ return NOT_EMULATED; }
