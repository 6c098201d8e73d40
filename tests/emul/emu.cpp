// tests/emul/emu.cpp — TEST INFRASTRUCTURE: the runtime behind tests/emul/cuda_runtime.h (see there).
#include "cuda_runtime.h"
#include <stdio.h>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

// The context switch between fibers: the callee-saved registers and the stack pointer (x86-64 System V).  swapcontext would do,
// but it saves and restores the signal mask with two system calls per switch, and a collective of one warp is 128 switches.
#if defined(__x86_64__)
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
.size emu_switch,.-emu_switch
)");
#else
#error "tests/emul needs an x86-64 host (emu_switch)"
#endif

namespace emu {
// One CUDA thread = one fiber of the calling OS thread; the threads of a block are scheduled one at a time, each
// running until it finishes or reaches a rendezvous (a warp collective, __syncthreads, the two ends of an `L0 { }` block).  The
// schedule is deterministic: between two rendezvous the threads take their turn in index order (EMU_ORDER=rev: highest first).
// Code whose lanes only communicate through rendezvous gives the same answer in both orders; the goldens are run in both.
struct Rendezvous {
	int expected = 0, arrived = 0; unsigned gen = 0;
	const char *what = "";
	// the threads that meet must come from the same call site after the same number of rendezvous
	uint64_t tag = 0; bool tag_set = false; int t_line = 0; const char *t_file = "";
};
struct Fiber {
	void *sp = nullptr; void *stack = nullptr; size_t stack_size = 0;
	bool done = false;
	Rendezvous *waiting = nullptr; unsigned wait_gen = 0; int w_line = 0; const char *w_file = "";
	unsigned tid = 0; uint64_t n_warp = 0, n_block = 0, n_grid = 0;
	void *fake = nullptr;                                   // AddressSanitizer's fake-stack handle across a switch
};
struct Warp { Rendezvous bar; uint64_t slots[32]; unsigned live = 0; };
struct Grid;
struct Block {
	Rendezvous bar; std::vector<Warp> warps; std::vector<uint64_t> dyn; std::vector<Fiber> fibers;
	Rendezvous grid_bar; bool grid_ready = false; Grid *grid_ctx = nullptr;   // cooperative launches: this block's side of grid.sync()
	void *sched_sp = nullptr; void *sched_fake = nullptr;
	const void *sched_bottom = nullptr; size_t sched_size = 0;
	const std::function<void()> *body = nullptr; dim3 block, grid; unsigned bidx = 0; const char *kernel = "";
	Fiber *cur = nullptr;
};
static thread_local Block *B = nullptr;                     // the block the calling OS thread is running (nullptr: sequential mode)
static thread_local const char *seq_kernel = "";

static void need_threads(const char *what){
	if (B && B->cur) return;
	fprintf(stderr, "emu: kernel %s calls %s but is launched in sequential mode: add it to the threaded set in emu.cpp\n", seq_kernel, what);
	abort();
}
#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/common_interface_defs.h>
#define SWITCH_BEGIN(save, bottom, size) __sanitizer_start_switch_fiber(save, bottom, size)
#define SWITCH_END(fake, ob, os) __sanitizer_finish_switch_fiber(fake, ob, os)
#else
#define SWITCH_BEGIN(save, bottom, size) do { } while (0)
#define SWITCH_END(fake, ob, os) do { } while (0)
#endif
static void to_scheduler(bool dying){
	Fiber *f = B->cur;
	SWITCH_BEGIN(dying ? nullptr : &f->fake, B->sched_bottom, B->sched_size);
	emu_switch(&f->sp, B->sched_sp);
	SWITCH_END(f->fake, nullptr, nullptr);
}
static void set_ids(const Fiber *f){
	threadIdx = { f->tid, 0, 0 }; blockIdx = { B->bidx, 0, 0 }; blockDim = B->block; gridDim = B->grid;
}
static void fiber_main(void){
	SWITCH_END(nullptr, &B->sched_bottom, &B->sched_size);
	Fiber *f = B->cur;
	set_ids(f);
	(*B->body)();
	f->done = true;
	Warp &w = B->warps[f->tid >> 5];
	w.live &= ~(1u << (f->tid & 31));
	for (Rendezvous *r : { &w.bar, &B->bar }) { r->expected--; if (r->expected > 0 && r->arrived >= r->expected) { r->arrived = 0; r->gen++; r->tag_set = false; } }
	B->grid_bar.expected--;
	if (B->grid_bar.expected > 0 && B->grid_bar.arrived >= B->grid_bar.expected) B->grid_ready = true;   // the others of this block already wait for the grid
	to_scheduler(true);
}
static void meet(Rendezvous &r, uint64_t my_tag, int line, const char *file){
	Fiber *f = B->cur;
	if (!r.tag_set) { r.tag = my_tag; r.t_line = line; r.t_file = file; r.tag_set = true; }
	else if (r.tag != my_tag || line != r.t_line || strcmp(file, r.t_file)) {
		fprintf(stderr, "emu: %s: divergent %s collective: one thread is at %s:%d after %llu rendezvous, thread %u at %s:%d after %llu\n",
			B->kernel, r.what, r.t_file, r.t_line, (unsigned long long)r.tag, f->tid, file, line, (unsigned long long)my_tag);
		abort();
	}
	const unsigned g = r.gen;
	if (++r.arrived >= r.expected) {
		if (&r == &B->grid_bar) B->grid_ready = true;        // released by the block's scheduler once every block is here
		else { r.arrived = 0; r.gen++; r.tag_set = false; }
	}
	f->waiting = &r; f->wait_gen = g; f->w_line = line; f->w_file = file;
	to_scheduler(false);                                     // (also the last one to arrive: the next stretch starts with the highest lane again)
	set_ids(f);
}
void warp_sync(int line, const char *file){ need_threads("a warp collective"); Fiber *f = B->cur; meet(B->warps[f->tid >> 5].bar, ++f->n_warp, line, file); }
void block_sync(void){ need_threads("__syncthreads"); Fiber *f = B->cur; meet(B->bar, ++f->n_block, 0, "__syncthreads"); }
void grid_sync(void){
	need_threads("grid.sync()");
	if (!B->grid_ctx) { fprintf(stderr, "emu: %s calls grid.sync() but was not launched cooperatively\n", B->kernel); abort(); }
	Fiber *f = B->cur; meet(B->grid_bar, ++f->n_grid, 0, "grid.sync()");
}
uint64_t *warp_slots(void){ need_threads("a warp collective"); return B->warps[B->cur->tid >> 5].slots; }
unsigned warp_live(void){ return B->warps[B->cur->tid >> 5].live; }
int lane_id(void){ return B && B->cur ? (int)(B->cur->tid & 31) : (int)(threadIdx.x & 31); }
void *dyn_smem(void){ need_threads("dynamic shared memory"); return B->dyn.data(); }

// kernels whose threads meet (warp collectives, __syncthreads, shared memory): one fiber per CUDA thread, blocks one after the
// other.  Everything else is plain data-parallel code and runs its threads one after the other as plain calls.
static const std::set<std::string> threaded = { "k_run_seq", "k_init_groups", "k_exec_ops", "k_stream_hash", "k_ens_drain", "k_ens_digest",
	"k_pick", "k_pick_small", "k_ana", "k_run_ens" };
// grid-stride kernels whose result does not depend on the grid: fewer blocks, same answer
static unsigned clamp_grid(const std::string &k, unsigned g){ if (k == "k_stream_hash") return 1; return g; }
// The order in which the threads of a block take their turn between two rendezvous: lowest first, or EMU_ORDER=rev.  (It found one
// place that depended on it — ana_process read a per-node table and wrote it back with no __syncwarp in between — since fixed;
// tests/emul/README.md.)
static bool lanes_forward(const std::string &){
	const char *e = getenv("EMU_ORDER");
	return !(e && !strcmp(e, "rev"));
}
#define EMU_STACK (4u << 20)
// fiber stacks are reused from launch to launch (the host core launches a kernel per event)
static thread_local std::vector<void *> stack_pool;
static void *stack_get(void){
	if (!stack_pool.empty()) { void *p = stack_pool.back(); stack_pool.pop_back(); return p; }
	void *p = malloc(EMU_STACK);
	if (!p) { fprintf(stderr, "emu: out of memory for a fiber stack\n"); abort(); }
	return p;
}
static void stack_put(void *p){ if (stack_pool.size() < 512) stack_pool.push_back(p); else free(p); }

// the blocks of a cooperative launch meet here: one OS thread per block, every block alive at once
struct Grid {
	std::mutex m; std::condition_variable cv;
	int expected = 0, arrived = 0; unsigned gen = 0;
	void wait(const char *kernel){
		std::unique_lock<std::mutex> l(m);
		const unsigned g = gen;
		if (++arrived >= expected) { arrived = 0; gen++; cv.notify_all(); return; }
		if (!cv.wait_for(l, std::chrono::seconds(300), [&]{ return gen != g; })) {
			fprintf(stderr, "emu: %s: grid.sync() timed out: %d of %d blocks arrived\n", kernel, arrived, expected);
			abort();
		}
	}
	void leave(){ std::lock_guard<std::mutex> l(m); expected--; if (expected > 0 && arrived >= expected) { arrived = 0; gen++; cv.notify_all(); } }
};

// one block, to completion, on the calling OS thread
static void run_block(const char *kernel, dim3 grid, dim3 block, unsigned b, size_t dyn_smem, const std::function<void()> &body, bool forward, Grid *gctx){
	const unsigned nt = block.x*block.y*block.z;
	Block blk;
	const unsigned nw = (nt+31)/32;
	blk.warps = std::vector<Warp>(nw);
	blk.dyn.assign((dyn_smem+7)/8+1, 0xCDCDCDCDCDCDCDCDull);
	blk.bar.expected = (int)nt; blk.bar.what = "__syncthreads";
	blk.grid_bar.expected = (int)nt; blk.grid_bar.what = "grid.sync()"; blk.grid_ctx = gctx;
	blk.body = &body; blk.block = block; blk.grid = grid; blk.bidx = b; blk.kernel = kernel;
	for (unsigned w = 0; w < nw; w++) {
		const unsigned n = std::min(32u, nt-32*w);
		blk.warps[w].bar.expected = (int)n; blk.warps[w].bar.what = "warp";
		blk.warps[w].live = n == 32 ? 0xffffffffu : ((1u << n)-1u);
		memset(blk.warps[w].slots, 0, sizeof blk.warps[w].slots);
	}
	blk.fibers = std::vector<Fiber>(nt);
	B = &blk;
	for (unsigned t = 0; t < nt; t++) {
		Fiber &f = blk.fibers[t];
		f.tid = t; f.stack_size = EMU_STACK; f.stack = stack_get();
		// a fresh fiber "returns" from emu_switch into fiber_main: six zeroed registers, the entry point, a null return address
		uintptr_t top = ((uintptr_t)f.stack+f.stack_size) & ~(uintptr_t)15;
		void **sp = (void **)(top-64);
		for (int k = 0; k < 6; k++) sp[k] = nullptr;
		sp[6] = (void *)fiber_main; sp[7] = nullptr;
		f.sp = sp;
	}
	unsigned left = nt;
	while (left) {
		bool progress = false;
		for (unsigned k = 0; k < nt; k++) {
			Fiber &f = blk.fibers[forward ? k : nt-1-k];
			if (f.done) continue;
			if (f.waiting) { if (f.waiting->gen == f.wait_gen) continue; f.waiting = nullptr; }
			blk.cur = &f;
			SWITCH_BEGIN(&blk.sched_fake, f.stack, f.stack_size);
			emu_switch(&blk.sched_sp, f.sp);
			SWITCH_END(blk.sched_fake, nullptr, nullptr);
			blk.cur = nullptr;
			progress = true;
			if (f.done) { left--; stack_put(f.stack); f.stack = nullptr; }
		}
		if (!progress && blk.grid_ready && left) {          // every live thread of this block is at grid.sync(): meet the other blocks
			gctx->wait(kernel);
			Rendezvous &r = blk.grid_bar;
			r.arrived = 0; r.gen++; r.tag_set = false; blk.grid_ready = false;
			progress = true;
		}
		if (!progress) {
			fprintf(stderr, "emu: %s block %u: dead-lock, %u threads wait for a rendezvous the others never reach:\n", kernel, b, left);
			for (auto &f : blk.fibers) if (!f.done) fprintf(stderr, "  thread %u waits at %s:%d (%s)\n", f.tid, f.w_file, f.w_line, f.waiting ? f.waiting->what : "?");
			abort();
		}
	}
	if (gctx) gctx->leave();
	B = nullptr;
}

void launch(const char *kernel, dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body){
	const std::string name(kernel);
	grid.x = clamp_grid(name, grid.x);
	const unsigned nt = block.x*block.y*block.z;
	if (B) { fprintf(stderr, "emu: launch of %s from inside a kernel\n", kernel); abort(); }
	if (!threaded.count(name)) {
		seq_kernel = kernel;
		blockDim = block; gridDim = grid;
		for (unsigned b = 0; b < grid.x; b++) for (unsigned t = 0; t < nt; t++) { blockIdx = { b, 0, 0 }; threadIdx = { t, 0, 0 }; body(); }
		return;
	}
	const bool forward = lanes_forward(name);
	for (unsigned b = 0; b < grid.x; b++) run_block(kernel, grid, block, b, dyn_smem, body, forward, nullptr);
}

void launch_coop(const char *kernel, dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body){
	if (B) { fprintf(stderr, "emu: launch of %s from inside a kernel\n", kernel); abort(); }
	const bool forward = lanes_forward(kernel);
	Grid g;
	g.expected = (int)grid.x;
	std::vector<std::thread> th;
	for (unsigned b = 0; b < grid.x; b++) th.emplace_back([&, b]{ run_block(kernel, grid, block, b, dyn_smem, body, forward, &g); });
	for (auto &t : th) t.join();
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
