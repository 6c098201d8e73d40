// tests/emul/stubs.cpp — TEST INFRASTRUCTURE: the grid-scope and CTA-scope batch engines (ns_par.cu, ns_par_cl.cu, ns_init.cu,
// ns_ens.cu: cooperative kernels of 148 x 256 threads) are outside the lock-step model; asking for them is an error, so the
// model only ever runs NS_MODE_SEQ (one warp per simulation).
#include "ns_par_api.cuh"
#define NOT_EMULATED NS_E_NO_DEVICE
int par_plan_init(ParScratch &, const ns_sc_op *, int, int){ return NOT_EMULATED; }
int par_init_group0(const Sim &, const Sim *, ParScratch &, int, cudaStream_t, uint64_t *){ return NOT_EMULATED; }
int ens_launch_run(const Sim *, const SeqScratch *, const ParScratch *, int, cudaStream_t, EnsAcc *, AnaStream *){ return NOT_EMULATED; }
int ens_preload(void){ return NOT_EMULATED; }
const void *par_kernel_cl(void){ return nullptr; }
int par_run(const Sim &, const Sim *, const SeqScratch *, ParScratch &, int, cudaStream_t, uint64_t *){ return NOT_EMULATED; }
int par_preload(void){ return NOT_EMULATED; }
int par_grid(const void *, int, int *, size_t){ return NOT_EMULATED; }
