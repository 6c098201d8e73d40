// ns_ana.cuh — per-stream state of the on-device analyzer reductions (ns_ana.cu)
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/netsim_b200.h"

struct AnaStream {
	const uint8_t *rec; const unsigned long long *rec_base, *rec_len;   // the replica's record buffer and its window of the stream
	struct ns_ana_summary *st;                                        // running summary (device)
	int n_nodes; uint32_t hmask; const uint32_t *hkey; const int *hval;   // node id -> index
	uint8_t *ntype;                                                   // 0 server, 1 cloud, 2 client (analyzers.c:40-46)
	double *u_time[2], *u_val[2];                                     // per node and direction: last usage record
	int8_t *n_state; double *n_since;                                 // per node: last state record
	double period;
};
int ana_launch(AnaStream *d_streams, int n, int final_flush, cudaStream_t stream);
