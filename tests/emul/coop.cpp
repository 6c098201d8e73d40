// tests/emul/coop.cpp — TEST INFRASTRUCTURE: the kernels of netsim_b200/csrc that are launched with cudaLaunchCooperativeKernel
// (a bare function pointer and an array of argument addresses), registered with their signatures.
#include "ns_par_api.cuh"
void k_run_par(const Sim *sp, const SeqScratch *scrp, ParScratch P);
void k_run_par_cl(const Sim *sp, const SeqScratch *scrp, ParScratch P);
void k_init_par(const Sim *sp, ParScratch P);
namespace {
struct Reg {
	Reg(){
		emu::register_coop((const void *)k_run_par, "k_run_par", [](void **a){ k_run_par(*(const Sim **)a[0], *(const SeqScratch **)a[1], *(ParScratch *)a[2]); });
		emu::register_coop((const void *)k_run_par_cl, "k_run_par_cl", [](void **a){ k_run_par_cl(*(const Sim **)a[0], *(const SeqScratch **)a[1], *(ParScratch *)a[2]); });
		emu::register_coop((const void *)k_init_par, "k_init_par", [](void **a){ k_init_par(*(const Sim **)a[0], *(ParScratch *)a[1]); });
	}
} reg;
}
