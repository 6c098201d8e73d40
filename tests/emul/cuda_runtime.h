/* tests/emul/cuda_runtime.h — TEST INFRASTRUCTURE, never part of the product.
 *
 * A lock-step warp model on the CPU: enough of the CUDA runtime and of the device intrinsics that netsim_b200/csrc's single-warp
 * engine (ns_seq.cu + ns_model.cuh + ns_sim.cuh), its pickers (ns_pick.cu), its analyzer kernel (ns_ana.cu) and the host layer
 * (ns_engine.cu) compile as plain C++ (tests/emul/prep.py rewrites only the <<<>>> launch syntax).  "Device memory" is malloc'd
 * host memory; a launched block is one fiber per CUDA thread, scheduled one at a time and deterministically (emu.cpp); warp
 * collectives (__shfl*_sync, __ballot_sync, __match_any_sync, __syncwarp) and __syncthreads are rendezvous points of the
 * participating threads: a collective that not every live lane reaches, or that lanes reach from different call sites, is
 * reported here where it is undefined on the GPU.
 *
 * What it is for: the device code under AddressSanitizer / UBSan (compute-sanitizer is closed on the GPU pool), and a
 * differential check of the device code against the CPU oracle without a GPU.  It is not a fallback: the product library is
 * netsim_b200/libnetsim_b200.so, which fails with NS_E_NO_DEVICE without a GPU; this model is ~10^4 times slower than one. */
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
/* (before __noinline__ becomes a macro: libstdc++ spells the attribute __attribute__((__noinline__))) */
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#ifdef EMU_THREADS
#define __shared__ static                   /* emu_threads.cpp: one OS thread per CUDA thread, one block at a time */
#else
#define __shared__ static thread_local      /* one OS thread runs one block at a time (emu.cpp): per-block storage */
#endif
#define __restrict__

struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

/* ---------------------------------------------------------------- runtime API (host memory, synchronous) */
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaHostAllocMapped = 2, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct emu_stream; struct emu_event;
typedef emu_stream *cudaStream_t;
typedef emu_event *cudaEvent_t;
struct cudaDeviceProp { int multiProcessorCount; char name[64]; };
struct cudaFuncAttributes { int numRegs; };
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaMalloc(void **p, size_t n);
template <typename T> cudaError_t cudaMalloc(T **p, size_t n){ return cudaMalloc((void **)p, n); }
cudaError_t cudaFree(void *p);
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned flags);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned flags);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t s = nullptr);
cudaError_t cudaMemset(void *p, int v, size_t n);
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t s = nullptr);
cudaError_t cudaStreamCreate(cudaStream_t *s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaDeviceSynchronize(void);
cudaError_t cudaGetLastError(void);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
template <typename F> cudaError_t cudaFuncGetAttributes(cudaFuncAttributes *a, F){ a->numRegs = 0; return cudaSuccess; }
template <typename F> cudaError_t cudaFuncSetAttribute(F, int, int){ return cudaSuccess; }
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, const void *, int, size_t){ *n = 1; return cudaSuccess; }
/* cooperative launches: the kernel arrives as a bare pointer, so the kernels that are launched this way register an invoker (coop.cpp) */
cudaError_t cudaLaunchCooperativeKernel(const void *kernel, dim3 grid, dim3 block, void **args, size_t dyn_smem, cudaStream_t s);
cudaError_t cudaGetDevice(int *d);

/* ---------------------------------------------------------------- launches and rendezvous */
namespace emu {
void launch(const char *kernel, dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body);
void launch_coop(const char *kernel, dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body);   /* every block alive at once: one OS thread per block */
void register_coop(const void *kernel, const char *name, std::function<void(void **)> invoke);
void grid_sync(void);                       /* every live thread of every block of a cooperative launch */
void *dyn_smem(void);                       /* `extern __shared__ T x[];` of the calling thread's block */
void warp_sync(int line = 0, const char *file = "");   /* every live lane of the calling thread's warp; (line, file) = the call site, which must be the same for all of them */
void block_sync(void);                      /* every live thread of the block */
uint64_t *warp_slots(void);                 /* 32 exchange words of the calling thread's warp */
unsigned warp_live(void);                   /* lanes of the warp that have not returned from the kernel */
int lane_id(void);
template <typename T> inline uint64_t to_bits(T v){ uint64_t b = 0; static_assert(sizeof(T) <= 8, "shuffle operand"); memcpy(&b, &v, sizeof v); return b; }
template <typename T> inline T from_bits(uint64_t b){ T v; memcpy(&v, &b, sizeof v); return v; }
template <typename T> inline T exchange(T v, int src, int line, const char *file){
	uint64_t *s = warp_slots();
	s[lane_id()] = to_bits(v);
	warp_sync(line, file);
	T r = from_bits<T>(s[src & 31]);
	warp_sync(line, file);
	return r;
}
/* `L0 { ... }` of ns_model.cuh (lane 0 acts for the warp) under the lock-step reading of a converged warp: every lane has executed
 * what precedes the block before lane 0 runs it, and sees its effects afterwards.  prep.py swaps the macro's definition for this
 * guard (two rendezvous around the body; none in scalar mode, where the lanes are independent threads). */
struct L0Guard {
	const bool scalar, mine; bool entered = false; const int line; const char *const file;
	L0Guard(bool scalar_, int lane, int line_, const char *file_) : scalar(scalar_), mine(scalar_ || lane == 0), line(line_), file(file_) { if (!scalar) warp_sync(line, file); }
	bool body(){ if (!entered) { entered = true; if (mine) return true; } if (!scalar) warp_sync(line, file); return false; }
};
}
#define EMU_L0 for (emu::L0Guard g_(c.scalar, c.lane, __LINE__, __FILE__); g_.body(); )
#define EMU_FULL 0xffffffffu
#define EMU_SITE int line = __builtin_LINE(), const char *file = __builtin_FILE()
inline void __syncwarp(unsigned mask = EMU_FULL, EMU_SITE){ (void)mask; emu::warp_sync(line, file); }
inline void __syncthreads(void){ emu::block_sync(); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src, int width = 32, EMU_SITE){ (void)width; return emu::exchange(v, src, line, file); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int d, int width = 32, EMU_SITE){ (void)width; return emu::exchange(v, emu::lane_id()^d, line, file); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32, EMU_SITE){ (void)width; int l = emu::lane_id(); return emu::exchange(v, l >= (int)d ? l-(int)d : l, line, file); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32, EMU_SITE){ (void)width; int l = emu::lane_id(); return emu::exchange(v, l+(int)d < 32 ? l+(int)d : l, line, file); }
inline unsigned __ballot_sync(unsigned, int pred, EMU_SITE){
	uint64_t *s = emu::warp_slots();
	s[emu::lane_id()] = pred ? 1 : 0;
	emu::warp_sync(line, file);
	unsigned r = 0, live = emu::warp_live();
	for (int i = 0; i < 32; i++) if (((live >> i) & 1u) && s[i]) r |= 1u << i;
	emu::warp_sync(line, file);
	return r;
}
inline int __any_sync(unsigned m, int pred, EMU_SITE){ return __ballot_sync(m, pred, line, file) != 0; }
inline int __all_sync(unsigned m, int pred, EMU_SITE){ return __ballot_sync(m, !pred, line, file) == 0; }
template <typename T> inline unsigned __match_any_sync(unsigned, T v, EMU_SITE){
	uint64_t *s = emu::warp_slots();
	s[emu::lane_id()] = emu::to_bits(v);
	emu::warp_sync(line, file);
	unsigned r = 0, live = emu::warp_live();
	const uint64_t mine = emu::to_bits(v);
	for (int i = 0; i < 32; i++) if (((live >> i) & 1u) && s[i] == mine) r |= 1u << i;
	emu::warp_sync(line, file);
	return r;
}
template <typename T, typename F> inline T emu_reduce(T v, F f, int line, const char *file){
	uint64_t *s = emu::warp_slots();
	s[emu::lane_id()] = emu::to_bits(v);
	emu::warp_sync(line, file);
	const unsigned live = emu::warp_live();
	bool first = true; T r = v;
	for (int i = 0; i < 32; i++) if ((live >> i) & 1u) { T x = emu::from_bits<T>(s[i]); r = first ? x : f(r, x); first = false; }
	emu::warp_sync(line, file);
	return r;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v, EMU_SITE){ return emu_reduce(v, [](unsigned a, unsigned b){ return a | b; }, line, file); }
inline unsigned __reduce_and_sync(unsigned, unsigned v, EMU_SITE){ return emu_reduce(v, [](unsigned a, unsigned b){ return a & b; }, line, file); }
inline unsigned __reduce_add_sync(unsigned, unsigned v, EMU_SITE){ return emu_reduce(v, [](unsigned a, unsigned b){ return a+b; }, line, file); }
inline int __reduce_add_sync(unsigned, int v, EMU_SITE){ return emu_reduce(v, [](int a, int b){ return a+b; }, line, file); }
inline unsigned __reduce_min_sync(unsigned, unsigned v, EMU_SITE){ return emu_reduce(v, [](unsigned a, unsigned b){ return a < b ? a : b; }, line, file); }
inline int __reduce_min_sync(unsigned, int v, EMU_SITE){ return emu_reduce(v, [](int a, int b){ return a < b ? a : b; }, line, file); }
inline unsigned __reduce_max_sync(unsigned, unsigned v, EMU_SITE){ return emu_reduce(v, [](unsigned a, unsigned b){ return a > b ? a : b; }, line, file); }
inline int __reduce_max_sync(unsigned, int v, EMU_SITE){ return emu_reduce(v, [](int a, int b){ return a > b ? a : b; }, line, file); }
inline void __threadfence(void){ __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block(void){ __atomic_thread_fence(__ATOMIC_SEQ_CST); }

/* ---------------------------------------------------------------- scalar intrinsics */
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned sel){
	const uint64_t src = ((uint64_t)y << 32) | x;
	unsigned r = 0;
	for (int i = 0; i < 4; i++) { const unsigned n = (sel >> (4*i)) & 0xF; unsigned b = (unsigned)(src >> (8*(n & 7))) & 0xFF; if (n & 8) b = (b & 0x80) ? 0xFF : 0; r |= b << (8*i); }
	return r;
}
inline int __popc(unsigned v){ return __builtin_popcount(v); }
inline int __popcll(unsigned long long v){ return __builtin_popcountll(v); }
inline int __ffs(int v){ return __builtin_ffs(v); }
inline int __ffsll(long long v){ return __builtin_ffsll(v); }
inline int __clz(int v){ return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v){ return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline double __longlong_as_double(long long v){ double d; memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d){ long long v; memcpy(&v, &d, 8); return v; }
inline int min(int a, int b){ return a < b ? a : b; }
inline int max(int a, int b){ return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b){ return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b){ return a > b ? a : b; }
inline long long min(long long a, long long b){ return a < b ? a : b; }
inline long long max(long long a, long long b){ return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b){ return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b){ return a > b ? a : b; }
inline unsigned long min(unsigned long a, unsigned long b){ return a < b ? a : b; }
inline unsigned long max(unsigned long a, unsigned long b){ return a > b ? a : b; }
inline double min(double a, double b){ return fmin(a, b); }
inline double max(double a, double b){ return fmax(a, b); }

/* ---------------------------------------------------------------- atomics (device scope == process scope here) */
#define EMU_ATOMICS(T) \
	inline T atomicAdd(T *p, T v){ return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); } \
	inline T atomicCAS(T *p, T cmp, T v){ __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; } \
	inline T atomicMin(T *p, T v){ T old = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { } return old; } \
	inline T atomicMax(T *p, T v){ T old = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { } return old; } \
	inline T atomicExch(T *p, T v){ return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); } \
	inline T atomicOr(T *p, T v){ return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); } \
	inline T atomicAnd(T *p, T v){ return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
EMU_ATOMICS(int)
EMU_ATOMICS(unsigned)
EMU_ATOMICS(unsigned long long)
inline double atomicAdd(double *p, double v){
	uint64_t *q = (uint64_t *)p, old = __atomic_load_n(q, __ATOMIC_SEQ_CST);
	while (true) { double nv = emu::from_bits<double>(old)+v; uint64_t nb = emu::to_bits(nv); if (__atomic_compare_exchange_n(q, &old, nb, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return emu::from_bits<double>(old); }
}
