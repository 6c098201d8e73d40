// ns_ana.cu — the analyzer's reductions on the device, straight from the record stream (record.c:53-116 format), so
// that a replica ensemble reports summaries instead of gigabytes of records:
//   node classes        analyzer/analyzers.c:80-101 (node_tracker): a node that ever entered N_SERVER is a server,
//                       N_CLOUD a cloud node, everything else a client
//   bandwidth by class  analyzers.c:226-297 (node_type_speed) + :153-180 (print_speed_per): the time average, per
//                       period, of the summed R_USAGE of all nodes of a class and direction
//   stale clients       analyzers.c:384-431 (stale_client): how many clients are in N_STALE over time; here its time
//                       integral per period (client-seconds) and its maximum
// One warp per stream.  The stream is variable-length (13-byte header + 1/4/6-byte value), so a tile is staged in
// shared memory, lane 0 walks the record boundaries there and the lanes decode one record each.  The reductions are
// written per NODE (integral of a node's usage / stale time between its own consecutive records), which makes them
// independent of the order across nodes: same-node records inside a tile are chained with match_any, everything else
// goes through the per-node tables.  Floating-point sums are order-dependent: results agree with the sequential
// analyzer to ~1e-12 relative, not bit for bit (the analyzer itself prints 6 decimals).
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/netsim_b200.h"
#include "ns_ana_dev.cuh"


__global__ void __launch_bounds__(32) k_ana(AnaStream *streams, int n, int final_flush){
	const int s = blockIdx.x;
	if (s >= n) return;
	__shared__ uint8_t tile[ANA_TILE+32];
	__shared__ int starts[34];
	ana_process(streams[s], tile, starts, threadIdx.x, final_flush);
}

int ana_launch(AnaStream *d_streams, int n, int final_flush, cudaStream_t stream){
	if (n <= 0) return 0;
	k_ana<<<n, 32, 0, stream>>>(d_streams, n, final_flush);
	cudaError_t e = cudaGetLastError();
	return e == cudaSuccess ? 0 : (int)e;
}
