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
			long long key = (long long)val[lo+i]*4294967296LL | (unsigned)i;   // (value, index) as one key; a multiplication, not << 32: the value may be negative
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

// Small tables (the reference's: c <= nsvr or ncld, tens of entries): one THREAD per request, 128 requests per CTA, every
// lane busy with its own quick-select.  The tables sit in shared memory with the request index as the fastest dimension
// ([entry][thread]), so the lanes of a warp never collide on a bank while each replays its own swap sequence.
#define PICK_SMALL_MAX 64
#define PICK_SMALL_THREADS 128
__global__ void __launch_bounds__(PICK_SMALL_THREADS) k_pick_small(int mode, int n_req, const int32_t *off, const int32_t *val, const int32_t *node,
		       const int32_t *kk, int32_t *out_val, int32_t *out_node, int32_t *out_cnt){
	extern __shared__ int32_t sm[];
	int32_t *v = sm+threadIdx.x, *nd = sm+PICK_SMALL_MAX*PICK_SMALL_THREADS+threadIdx.x;
	#define V(i) v[(i)*PICK_SMALL_THREADS]
	#define N(i) nd[(i)*PICK_SMALL_THREADS]
	const int q = blockIdx.x*PICK_SMALL_THREADS+threadIdx.x;
	if (q >= n_req) return;
	const int lo = off[q], c = off[q+1]-lo;
	if (mode == 2) {                                                      // first strict minimum (client_behaviour.h:94-117)
		int best = -1; int32_t bv = INT32_MAX;
		for (int i = 0; i < c; i++) { const int32_t x = val[lo+i]; if (x < bv) { bv = x; best = i; } }
		if (best < 0) { out_cnt[q] = 0; if (c > 0) out_node[lo] = -1; }
		else { out_cnt[q] = 1; out_node[lo] = node[lo+best]; out_val[lo] = bv; }
		return;
	}
	const int k = kk[q];
	for (int i = 0; i < c; i++) { V(i) = val[lo+i]; N(i) = node[lo+i]; }
	if (c <= k || k <= 0) out_cnt[q] = c < 0 ? 0 : (k <= 0 ? 0 : c);      // client_behaviour.h:77-80
	else {
		int base = 0, nmemb = c, kq = k;                                  // qselect_eval(v, c, k), qselect.h:5-27
		while (true) {
			int st = 0;
			const int32_t pv = V(base+nmemb-1);
			for (int i = 0; i < nmemb-1; i++) {
				const int32_t vi = V(base+i);
				if ((int32_t)((uint32_t)vi-(uint32_t)pv) > 0) continue;
				const int32_t tv = V(base+st), tn = N(base+st);
				V(base+st) = vi; N(base+st) = N(base+i);
				V(base+i) = tv; N(base+i) = tn;
				st++;
			}
			const int32_t tv = V(base+st), tn = N(base+st);
			V(base+st) = V(base+nmemb-1); N(base+st) = N(base+nmemb-1);
			V(base+nmemb-1) = tv; N(base+nmemb-1) = tn;
			if (kq-1 == st) break;
			if (st > kq-1) nmemb = st;
			else { base += st; nmemb -= st; kq -= st; }
		}
		out_cnt[q] = k;
	}
	for (int i = 0; i < c; i++) { out_val[lo+i] = V(i); out_node[lo+i] = N(i); }
	#undef V
	#undef N
}

// launch on device pointers (used by ns_dev_pick_live, which evaluates the metric on the device first); max_c = the largest table
int ns_pick_launch(int mode, int n_req, const int32_t *d_off, const int32_t *d_val, const int32_t *d_node, const int32_t *d_k,
		   int32_t *d_ov, int32_t *d_on, int32_t *d_oc, cudaStream_t stream, int max_c){
	if (n_req <= 0) return 0;
	if (max_c <= PICK_SMALL_MAX && n_req >= 32) {
		const size_t smem = (size_t)2*PICK_SMALL_MAX*PICK_SMALL_THREADS*4;
		static bool attr_set[64];
		int dev = 0; cudaGetDevice(&dev);
		if (!attr_set[dev & 63]) { cudaFuncSetAttribute(k_pick_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set[dev & 63] = true; }
		k_pick_small<<<(n_req+PICK_SMALL_THREADS-1)/PICK_SMALL_THREADS, PICK_SMALL_THREADS, smem, stream>>>(mode, n_req, d_off, d_val, d_node, d_k, d_ov, d_on, d_oc);
	} else
		k_pick<<<(n_req+3)/4, 128, 0, stream>>>(mode, n_req, d_off, d_val, d_node, d_k, d_ov, d_on, d_oc);
	cudaError_t e = cudaGetLastError();
	return e == cudaSuccess ? 0 : (int)e;
}

#define CKP(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "netsim_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return (int)e_; } } while (0)

// device buffers of the batch entry point: grown on demand, kept per device for the life of the process (a batch of
// simultaneous NEW_CONNECTION events is a per-timestamp call: no cudaMalloc / cudaFree / event creation on that path)
#include <mutex>
struct PickCache { void *p[7] = {nullptr}; size_t cap[7] = {0}; cudaEvent_t e0 = nullptr, e1 = nullptr; cudaStream_t st = nullptr; };
static PickCache g_pick[64];
static std::mutex g_pick_mu;
static cudaError_t pick_buf(PickCache &c, int i, size_t bytes, int32_t **out){
	if (bytes > c.cap[i]) {
		if (c.p[i]) cudaFree(c.p[i]);
		c.p[i] = nullptr; c.cap[i] = 0;
		size_t want = bytes+(bytes >> 1)+256;
		cudaError_t e = cudaMalloc(&c.p[i], want);
		if (e != cudaSuccess) return e;
		c.cap[i] = want;
	}
	*out = (int32_t *)c.p[i];
	return cudaSuccess;
}
extern "C" int ns_dev_pick_batch(int device, int mode, int n_req, const int32_t *off, const int32_t *val,
		       const int32_t *node, const int32_t *k, int32_t *out_val, int32_t *out_node,
		       int32_t *out_cnt, float *device_ms){
	if (n_req < 0 || !off || (mode != 1 && mode != 2)) return NS_E_ARG;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device || device < 0 || device >= 64) return NS_E_NO_DEVICE;
	CKP(cudaSetDevice(device));
	if (n_req == 0) return 0;
	std::lock_guard<std::mutex> lock(g_pick_mu);
	PickCache &c = g_pick[device];
	if (!c.st) { CKP(cudaStreamCreate(&c.st)); CKP(cudaEventCreate(&c.e0)); CKP(cudaEventCreate(&c.e1)); }
	const int total = off[n_req];
	int max_c = 0;
	for (int q = 0; q < n_req; q++) { if (off[q+1] < off[q]) return NS_E_ARG; max_c = off[q+1]-off[q] > max_c ? off[q+1]-off[q] : max_c; }
	int32_t *d_off, *d_val, *d_node, *d_k, *d_ov, *d_on, *d_oc;
	const size_t tb = (size_t)(total > 0 ? total : 1)*4;
	CKP(pick_buf(c, 0, (size_t)(n_req+1)*4, &d_off)); CKP(pick_buf(c, 1, tb, &d_val)); CKP(pick_buf(c, 2, tb, &d_node));
	CKP(pick_buf(c, 3, tb, &d_ov)); CKP(pick_buf(c, 4, tb, &d_on)); CKP(pick_buf(c, 5, (size_t)n_req*4, &d_oc)); CKP(pick_buf(c, 6, (size_t)n_req*4, &d_k));
	CKP(cudaMemcpyAsync(d_off, off, (size_t)(n_req+1)*4, cudaMemcpyHostToDevice, c.st));
	if (total > 0) { CKP(cudaMemcpyAsync(d_val, val, tb, cudaMemcpyHostToDevice, c.st)); CKP(cudaMemcpyAsync(d_node, node, tb, cudaMemcpyHostToDevice, c.st)); }
	if (k) CKP(cudaMemcpyAsync(d_k, k, (size_t)n_req*4, cudaMemcpyHostToDevice, c.st)); else CKP(cudaMemsetAsync(d_k, 0, (size_t)n_req*4, c.st));
	CKP(cudaEventRecord(c.e0, c.st));
	{ int rc = ns_pick_launch(mode, n_req, d_off, d_val, d_node, d_k, d_ov, d_on, d_oc, c.st, max_c); if (rc) return rc; }
	CKP(cudaEventRecord(c.e1, c.st));
	if (total > 0) { CKP(cudaMemcpyAsync(out_val, d_ov, tb, cudaMemcpyDeviceToHost, c.st)); CKP(cudaMemcpyAsync(out_node, d_on, tb, cudaMemcpyDeviceToHost, c.st)); }
	CKP(cudaMemcpyAsync(out_cnt, d_oc, (size_t)n_req*4, cudaMemcpyDeviceToHost, c.st));
	CKP(cudaStreamSynchronize(c.st));
	if (device_ms) CKP(cudaEventElapsedTime(device_ms, c.e0, c.e1));
	return 0;
}
