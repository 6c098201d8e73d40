// ns_par_cl.cu — the grid engine of ns_par_kernel.inc once more, this time WITH the behaviour-library hooks of
// ns_model.cuh (client water-mark machine, NS_OP_PLAY1): k_run_par_cl steps simulations whose script carries
// NS_SC_F_CLIENTS.  Plain simulations keep k_run_par (ns_par.cu), whose code layout the extra code would disturb.
#define NS_NO_WIDE
#include "ns_par_dev.cuh"

#define NS_KERNEL_DECL __global__ void __launch_bounds__(PAR_THREADS, PAR_MIN_BLOCKS) k_run_par_cl(const Sim *sp, const SeqScratch *scrp, ParScratch P)
#define NS_KERNEL_PROLOGUE cgp::grid_group grid = cgp::this_grid();
#include "ns_par_kernel.inc"

const void *par_kernel_cl(void){ return (const void *)k_run_par_cl; }
