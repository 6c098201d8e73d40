// tests/emul/emu_threads.cpp — TEST INFRASTRUCTURE: a second runtime behind cuda_runtime.h, for ThreadSanitizer only
// (`make tsan`): one OS thread per CUDA thread, blocks one after the other, every rendezvous a mutex + condition variable, so that
// the sanitizer sees exactly the synchronisation the device code asks for (collectives, __syncwarp, __syncthreads, grid.sync() with
// one CTA, the two ends of an `L0 { }` block) and reports every pair of accesses by two CUDA threads to the same address, one of
// them a write, with nothing of that in between: a race between lanes that run independently (the scalar threads of the
// one-event-per-lane path, the one-thread-per-group pass of the batch engine), or a dependence on warp-synchronous execution.
// The deterministic fiber runtime (emu.cpp) is the one the test suite uses; this one is ~50x slower and its schedule is the OS's.
#include "cuda_runtime.h"
#include <stdio.h>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

namespace emu {
struct Barrier {
	std::mutex m; std::condition_variable cv;
	int expected = 0, arrived = 0; unsigned gen = 0;
	const char *what = "";
	uint64_t tag = 0; bool tag_set = false; int t_line = 0; const char *t_file = "";
	void wait(uint64_t my_tag, const char *kernel, int line, const char *file){
		std::unique_lock<std::mutex> l(m);
		const unsigned g = gen;
		if (!tag_set) { tag = my_tag; t_line = line; t_file = file; tag_set = true; }
		else if (tag != my_tag || line != t_line || strcmp(file, t_file)) {
			fprintf(stderr, "emu: %s: divergent %s collective: one thread is at %s:%d after %llu rendezvous, another at %s:%d after %llu\n",
				kernel, what, t_file, t_line, (unsigned long long)tag, file, line, (unsigned long long)my_tag);
			abort();
		}
		if (++arrived >= expected) { arrived = 0; gen++; tag_set = false; cv.notify_all(); return; }
		if (!cv.wait_for(l, std::chrono::seconds(600), [&]{ return gen != g; })) {
			fprintf(stderr, "emu: %s: %s rendezvous timed out: %d of %d threads arrived\n", kernel, what, arrived, expected);
			abort();
		}
	}
	void leave(){
		std::lock_guard<std::mutex> l(m);
		expected--;
		if (expected > 0 && arrived >= expected) { arrived = 0; gen++; tag_set = false; cv.notify_all(); }
	}
};
struct Warp { Barrier bar; uint64_t slots[32]; unsigned live = 0; };
struct Block { Barrier bar; std::vector<Warp> warps; std::vector<uint64_t> dyn; bool coop = false; };
struct Me { Block *blk = nullptr; Warp *w = nullptr; int lane = 0; const char *kernel = ""; uint64_t n_warp = 0, n_block = 0; };
static thread_local Me me;
static thread_local const char *seq_kernel = "";

static void need_threads(const char *what){
	if (me.w) return;
	fprintf(stderr, "emu: kernel %s calls %s but is launched in sequential mode\n", seq_kernel, what);
	abort();
}
void warp_sync(int line, const char *file){ need_threads("a warp collective"); me.w->bar.wait(++me.n_warp, me.kernel, line, file); }
void block_sync(void){ need_threads("__syncthreads"); me.blk->bar.wait(++me.n_block, me.kernel, 0, "__syncthreads"); }
void grid_sync(void){ need_threads("grid.sync()"); if (!me.blk->coop) abort(); me.blk->bar.wait(++me.n_block, me.kernel, 0, "grid.sync()"); }   // one CTA: the grid is the block
uint64_t *warp_slots(void){ need_threads("a warp collective"); return me.w->slots; }
unsigned warp_live(void){ return __atomic_load_n(&me.w->live, __ATOMIC_SEQ_CST); }
int lane_id(void){ return me.w ? me.lane : (int)(threadIdx.x & 31); }
void *dyn_smem(void){ need_threads("dynamic shared memory"); return me.blk->dyn.data(); }

static const std::set<std::string> threaded = { "k_run_seq", "k_init_groups", "k_exec_ops", "k_stream_hash", "k_ens_drain", "k_ens_digest",
	"k_pick", "k_pick_small", "k_ana", "k_run_ens", "k_run_par", "k_run_par_cl", "k_init_par" };

static void run_block(const char *kernel, dim3 grid, dim3 block, unsigned b, size_t dyn_smem, const std::function<void()> &body, bool coop){
	const unsigned nt = block.x*block.y*block.z;
	Block blk;
	const unsigned nw = (nt+31)/32;
	blk.warps = std::vector<Warp>(nw);
	blk.dyn.assign((dyn_smem+7)/8+1, 0xCDCDCDCDCDCDCDCDull);
	blk.bar.expected = (int)nt; blk.bar.what = "block"; blk.coop = coop;
	for (unsigned w = 0; w < nw; w++) {
		const unsigned n = std::min(32u, nt-32*w);
		blk.warps[w].bar.expected = (int)n; blk.warps[w].bar.what = "warp";
		blk.warps[w].live = n == 32 ? 0xffffffffu : ((1u << n)-1u);
		memset(blk.warps[w].slots, 0, sizeof blk.warps[w].slots);
	}
	std::vector<std::thread> th;
	th.reserve(nt);
	for (unsigned t = 0; t < nt; t++)
		th.emplace_back([&, t]{
			threadIdx = { t, 0, 0 }; blockIdx = { b, 0, 0 }; blockDim = block; gridDim = grid;
			me = Me(); me.blk = &blk; me.w = &blk.warps[t >> 5]; me.lane = (int)(t & 31); me.kernel = kernel;
			body();
			__atomic_fetch_and(&me.w->live, ~(1u << me.lane), __ATOMIC_SEQ_CST);
			me.w->bar.leave();
			blk.bar.leave();
		});
	for (auto &x : th) x.join();
}
void launch(const char *kernel, dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body){
	const std::string name(kernel);
	if (name == "k_stream_hash") grid.x = 1;
	const unsigned nt = block.x*block.y*block.z;
	if (!threaded.count(name)) {
		seq_kernel = kernel;
		blockDim = block; gridDim = grid;
		for (unsigned b = 0; b < grid.x; b++) for (unsigned t = 0; t < nt; t++) { blockIdx = { b, 0, 0 }; threadIdx = { t, 0, 0 }; body(); }
		return;
	}
	for (unsigned b = 0; b < grid.x; b++) run_block(kernel, grid, block, b, dyn_smem, body, false);
}
void launch_coop(const char *kernel, dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body){
	if (grid.x != 1) { fprintf(stderr, "emu (threads): cooperative launches run with one CTA here (EMU_SMS=1)\n"); abort(); }
	run_block(kernel, grid, block, 0, dyn_smem, body, true);
}
struct CoopKernel { std::string name; std::function<void(void **)> invoke; };
static std::map<const void *, CoopKernel> &coop_table(void){ static std::map<const void *, CoopKernel> t; return t; }
void register_coop(const void *kernel, const char *name, std::function<void(void **)> invoke){ coop_table()[kernel] = CoopKernel{ name, std::move(invoke) }; }
}   // namespace emu

// ---------------------------------------------------------------- runtime API
struct emu_stream { int id; };
struct emu_event { std::chrono::steady_clock::time_point t; };
const char *cudaGetErrorString(cudaError_t e){ return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "error"; }
static bool zero_fill(void){ static int z = -1; if (z < 0) { const char *e = getenv("EMU_ZERO"); z = e && *e && *e != '0'; } return z != 0; }
cudaError_t cudaMalloc(void **p, size_t n){
	*p = malloc(n ? n : 1);
	if (!*p) return cudaErrorMemoryAllocation;
	memset(*p, zero_fill() ? 0 : 0xCD, n);                  // cudaMalloc does not clear: make a read of never-written memory visible
	return cudaSuccess;
}
cudaError_t cudaFree(void *p){ free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned){ *p = calloc(1, n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p){ free(p); return cudaSuccess; }
cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned){ *d = h; return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind){ if (n) memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t){ if (n) memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t n){ if (n) memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t){ if (n) memset(p, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreate(cudaStream_t *s){ *s = new emu_stream{ 0 }; return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned){ *s = new emu_stream{ 1 }; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t){ return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s){ delete s; return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e){ *e = new emu_event{ std::chrono::steady_clock::now() }; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned){ return cudaEventCreate(e); }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t){ e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b){ *ms = std::chrono::duration<float, std::milli>(b->t-a->t).count(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e){ delete e; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void){ return cudaSuccess; }
cudaError_t cudaGetLastError(void){ return cudaSuccess; }
cudaError_t cudaSetDevice(int){ return cudaSuccess; }
cudaError_t cudaGetDevice(int *d){ *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n){ *n = 1; return cudaSuccess; }
cudaError_t cudaLaunchCooperativeKernel(const void *kernel, dim3 grid, dim3 block, void **args, size_t dyn_smem, cudaStream_t){
	auto &t = emu::coop_table();
	auto it = t.find(kernel);
	if (it == t.end()) { fprintf(stderr, "emu: cooperative launch of a kernel that is not registered (coop.cpp)\n"); abort(); }
	emu::launch_coop(it->second.name.c_str(), grid, block, dyn_smem, [&]{ it->second.invoke(args); });
	return cudaSuccess;
}
// EMU_SMS: the number of multiprocessors the model reports = the CTAs of a cooperative launch (k_run_par, k_init_par)
static int emu_sms(void){ const char *e = getenv("EMU_SMS"); int n = e ? atoi(e) : 1; return n < 1 ? 1 : n > 16 ? 16 : n; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int){ memset(p, 0, sizeof *p); p->multiProcessorCount = emu_sms(); strcpy(p->name, "lock-step warp model (tests/emul)"); return cudaSuccess; }
