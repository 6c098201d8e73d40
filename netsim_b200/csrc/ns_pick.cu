// ns_pick.cu — batched source selection for simultaneous NEW_CONNECTION events.
//   mode 1: server_picker_opt1 (modules/client_behaviour.h:68-88): keep the k smallest metric values
//           with the reference's own quick-select (modules/qselect.h:5-27, Lomuto partition, last
//           element as pivot, `<=` goes left).  The ORDER the partition leaves matters: slot i
//           decides which source serves strip i*strip (client_behaviour.c:382-385), so the kernel
//           replays the exact swap sequence, one request per warp with the table staged in
//           shared memory.
//   mode 2: server_picker_opt2 (client_behaviour.h:90-124): first strict minimum over the holders.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/netsim_b200.h"

#define PICK_SMEM_ENTRIES 1024

// one warp per request; tables up to PICK_SMEM_ENTRIES are staged in shared memory (coalesced
// loads), larger ones are partitioned in place in global memory.
__global__ void __launch_bounds__(128) k_pick(int mode, int n_req, const int32_t *off, const int32_t *val, const int32_t *node,
		       const int32_t *kk, int32_t *out_val, int32_t *out_node, int32_t *out_cnt){
	__shared__ int32_t s_val[4][PICK_SMEM_ENTRIES];
	__shared__ int32_t s_node[4][PICK_SMEM_ENTRIES];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int q = blockIdx.x*4+w;
	if (q >= n_req) return;
	const int lo = off[q], c = off[q+1]-lo;
	if (mode == 2) {
		// first strict minimum == minimum of (val, index) pairs
		long long best = 0x7fffffffffffffffLL;
		for (int i = lane; i < c; i += 32) {
			long long key = ((long long)val[lo+i] << 32) | (unsigned)i;
			if (val[lo+i] != INT32_MAX && key < best) best = key;
		}
		for (int d = 16; d; d >>= 1) { long long o = __shfl_xor_sync(0xffffffffu, best, d); best = o < best ? o : best; }
		if (lane == 0) {
			if (best == 0x7fffffffffffffffLL) { out_cnt[q] = 0; if (c > 0) out_node[lo] = -1; }
			else { int i = (int)(best & 0xffffffff); out_cnt[q] = 1; out_node[lo] = node[lo+i]; out_val[lo] = val[lo+i]; }
		}
		return;
	}
	const int k = kk[q];
	int32_t *v, *nd;
	const bool in_smem = c <= PICK_SMEM_ENTRIES;
	if (in_smem) { v = s_val[w]; nd = s_node[w]; }
	else { v = out_val+lo; nd = out_node+lo; }
	for (int i = lane; i < c; i += 32) { v[i] = val[lo+i]; nd[i] = node[lo+i]; }
	__syncwarp();
	if (c <= k || k <= 0) {                                              // client_behaviour.h:77-80
		if (lane == 0) out_cnt[q] = c < 0 ? 0 : (k <= 0 ? 0 : c);
	} else {
		if (lane == 0) {
			// qselect_eval(v, c, k): iterative form of the tail recursion
			int base = 0, nmemb = c, kq = k;
			while (true) {
				int st = 0;
				const int32_t pv = v[base+nmemb-1];
				for (int i = 0; i < nmemb-1; i++) {
					int32_t vi = v[base+i];
					if ((int32_t)((uint32_t)vi-(uint32_t)pv) > 0) continue;
					int32_t tv = v[base+st], tn = nd[base+st];
					v[base+st] = vi; nd[base+st] = nd[base+i];
					v[base+i] = tv; nd[base+i] = tn;
					st++;
				}
				int32_t tv = v[base+st], tn = nd[base+st];
				v[base+st] = v[base+nmemb-1]; nd[base+st] = nd[base+nmemb-1];
				v[base+nmemb-1] = tv; nd[base+nmemb-1] = tn;
				if (kq-1 == st) break;
				if (st > kq-1) nmemb = st;
				else { base += st; nmemb -= st; kq -= st; }
			}
			out_cnt[q] = k;
		}
	}
	__syncwarp();
	if (in_smem) for (int i = lane; i < c; i += 32) { out_val[lo+i] = v[i]; out_node[lo+i] = nd[i]; }
}

// launch on device pointers (used by ns_dev_pick_live, which evaluates the metric on the device first)
int ns_pick_launch(int mode, int n_req, const int32_t *d_off, const int32_t *d_val, const int32_t *d_node, const int32_t *d_k,
		   int32_t *d_ov, int32_t *d_on, int32_t *d_oc, cudaStream_t stream){
	if (n_req <= 0) return 0;
	k_pick<<<(n_req+3)/4, 128, 0, stream>>>(mode, n_req, d_off, d_val, d_node, d_k, d_ov, d_on, d_oc);
	cudaError_t e = cudaGetLastError();
	return e == cudaSuccess ? 0 : (int)e;
}

#define CKP(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "netsim_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return (int)e_; } } while (0)

extern "C" int ns_dev_pick_batch(int device, int mode, int n_req, const int32_t *off, const int32_t *val,
		       const int32_t *node, const int32_t *k, int32_t *out_val, int32_t *out_node,
		       int32_t *out_cnt, float *device_ms){
	if (n_req < 0 || !off || (mode != 1 && mode != 2)) return NS_E_ARG;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) return NS_E_NO_DEVICE;
	CKP(cudaSetDevice(device));
	if (n_req == 0) return 0;
	const int total = off[n_req];
	int32_t *d_off, *d_val, *d_node, *d_k = nullptr, *d_ov, *d_on, *d_oc;
	size_t tb = (size_t)(total > 0 ? total : 1)*4;
	struct Tmp {                                                         // released on every way out
		void *p[8]; int n = 0; cudaEvent_t e0 = nullptr, e1 = nullptr;
		cudaError_t get(int32_t **out, size_t bytes){ cudaError_t e = cudaMalloc((void **)out, bytes); if (e == cudaSuccess) p[n++] = *out; return e; }
		~Tmp(){ for (int i = 0; i < n; i++) cudaFree(p[i]); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
	} tmp;
	CKP(tmp.get(&d_off, (size_t)(n_req+1)*4)); CKP(tmp.get(&d_val, tb)); CKP(tmp.get(&d_node, tb));
	CKP(tmp.get(&d_ov, tb)); CKP(tmp.get(&d_on, tb)); CKP(tmp.get(&d_oc, (size_t)n_req*4));
	CKP(tmp.get(&d_k, (size_t)n_req*4));
	CKP(cudaMemcpy(d_off, off, (size_t)(n_req+1)*4, cudaMemcpyHostToDevice));
	if (total > 0) { CKP(cudaMemcpy(d_val, val, tb, cudaMemcpyHostToDevice)); CKP(cudaMemcpy(d_node, node, tb, cudaMemcpyHostToDevice)); }
	if (k) CKP(cudaMemcpy(d_k, k, (size_t)n_req*4, cudaMemcpyHostToDevice)); else CKP(cudaMemset(d_k, 0, (size_t)n_req*4));
	cudaEvent_t &e0 = tmp.e0, &e1 = tmp.e1;
	CKP(cudaEventCreate(&e0)); CKP(cudaEventCreate(&e1));
	CKP(cudaEventRecord(e0));
	k_pick<<<(n_req+3)/4, 128>>>(mode, n_req, d_off, d_val, d_node, d_k, d_ov, d_on, d_oc);
	CKP(cudaGetLastError());
	CKP(cudaEventRecord(e1));
	CKP(cudaEventSynchronize(e1));
	if (device_ms) CKP(cudaEventElapsedTime(device_ms, e0, e1));
	if (total > 0) { CKP(cudaMemcpy(out_val, d_ov, tb, cudaMemcpyDeviceToHost)); CKP(cudaMemcpy(out_node, d_on, tb, cudaMemcpyDeviceToHost)); }
	CKP(cudaMemcpy(out_cnt, d_oc, (size_t)n_req*4, cudaMemcpyDeviceToHost));
	return 0;
}
