// ns_engine.cu — kernels and the extern "C" layer (include/netsim_b200.h).
//
// Host side here is what netsim keeps on the host: parsing the module's script, sim_create_node /
// sim_node_new_resource / sim_node_add_resource (sim.c:113-154: ids from random(), the R_NODE_CREATE
// records) and laying out the device tables.  Everything from the first sim_establish_flow on —
// rate solves, byte progress, completion times, event selection — runs in the kernels.
// There is no CPU fallback: every entry point fails with NS_E_NO_DEVICE when CUDA is unusable.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <unordered_set>

#include "ns_sim.cuh"
#include "ns_par_api.cuh"
#include "ns_ana.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "netsim_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return (int)e_; } } while (0)

// ---------------------------------------------------------------- small kernels
__global__ void k_fill_f64(double *p, u64 n, double v){
	for (u64 i = blockIdx.x*(u64)blockDim.x+threadIdx.x; i < n; i += (u64)gridDim.x*blockDim.x) p[i] = v;
}
__global__ void k_fill_i32(int *p, u64 n, int v){
	for (u64 i = blockIdx.x*(u64)blockDim.x+threadIdx.x; i < n; i += (u64)gridDim.x*blockDim.x) p[i] = v;
}

struct FlowInfoDev { int alive, src, dst, pad; unsigned flow_id, resource_id; double ub, delay, speed0, speed1; long long drng_start; double drng_len; };
__global__ void k_flow_info(const Sim *sp, int ord, FlowInfoDev *out){
	const Sim &S = *sp;
	FlowInfoDev r;
	memset(&r, 0, sizeof r);
	int f = (ord >= 0 && ord < S.st->n_ord) ? S.ord2flow[ord] : -1;
	if (f >= 0) {
		r.alive = 1; r.src = S.f_peer[0][f]; r.dst = S.f_peer[1][f]; r.flow_id = S.f_id[f];
		r.ub = S.f_ub[f]; r.delay = S.f_delay[f]; r.speed0 = S.f_speed[0][f]; r.speed1 = S.f_speed[1][f];
		int dr = S.f_drng[f];
		if (dr >= 0) { r.drng_start = S.r_start[dr]; r.drng_len = S.r_len[dr]; r.resource_id = S.s_id[S.r_res[dr]]; }
	}
	*out = r;
}
struct RangeInfoDev { long long start, total_len; double len; int producer_ord, pad; };
__global__ void k_node_ranges(const Sim *sp, int node, unsigned rid, RangeInfoDev *out, int max, int *n_out){
	const Sim &S = *sp;
	int n = 0;
	for (int r = S.n_res_head[node]; r >= 0; r = S.s_next[r]) {
		if (S.s_id[r] != rid) continue;
		for (int g = S.s_first[r]; g >= 0 && n < max; g = S.r_next[g]) {
			out[n].start = S.r_start[g]; out[n].total_len = S.r_total[g]; out[n].len = S.r_len[g];
			out[n].producer_ord = S.r_prod[g] >= 0 ? S.f_ord[S.r_prod[g]] : -1; out[n].pad = 0;
			n++;
		}
		break;
	}
	*n_out = n;
}

// order-sensitive, parallel digest of a byte stream: sum over little-endian 64-bit words w_j of
// mix64(w_j + K*(j+1)), finalised with the length (host mirror: stream_hash in netsim_b200/engine.py).
__global__ void k_stream_hash(const uint8_t *p, u64 len, u64 *out){
	u64 nw = (len+7)/8;
	u64 acc = 0;
	for (u64 j = blockIdx.x*(u64)blockDim.x+threadIdx.x; j < nw; j += (u64)gridDim.x*blockDim.x) {
		u64 w = 0;
		u64 b0 = j*8;
		for (int k = 0; k < 8; k++) if (b0+k < len) w |= (u64)p[b0+k] << (8*k);
		acc += ns_mix64(w+0x9e3779b97f4a7c15ULL*(j+1));
	}
	for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(FULLMASK, acc, d);
	if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

__global__ void k_flow_state(const Sim *sp, double *out4, int n){
	const Sim &S = *sp;
	for (int i = blockIdx.x*blockDim.x+threadIdx.x; i < n; i += gridDim.x*blockDim.x) {
		int f = S.ord2flow[i];
		double a = 0, s0 = 0, s1 = 0, l = 0;
		if (f >= 0) { a = 1; s0 = S.f_speed[0][f]; s1 = S.f_speed[1][f]; int dr = S.f_drng[f]; l = dr >= 0 ? S.r_len[dr] : 0; }
		out4[4*i] = a; out4[4*i+1] = s0; out4[4*i+2] = s1; out4[4*i+3] = l;
	}
}
__global__ void k_node_state(const Sim *sp, double *out4, int n){
	const Sim &S = *sp;
	for (int i = blockIdx.x*blockDim.x+threadIdx.x; i < n; i += gridDim.x*blockDim.x) {
		out4[4*i] = S.n_usage[0][i]; out4[4*i+1] = S.n_usage[1][i];
		out4[4*i+2] = S.n_total[0][i]; out4[4*i+3] = S.n_total[1][i];
	}
}

// fetch metric of every candidate, evaluated against the LIVE node table of the simulation:
// metric 0 = distance_metric (modules/p2p_common.h:171-179), metric 1 = share_metric (:190-198, with
// distance_based_bw :181-189); unusable candidates (server_picker_opt2's filters) get INT32_MAX
__global__ void k_pick_metric(const Sim *sp, int metric, int n_req, const int32_t *off, const int32_t *cand_node, const int32_t *cand_tz,
			      const int32_t *client_tz, const uint8_t *usable, int32_t *val){
	const Sim &S = *sp;
	for (int q = blockIdx.x; q < n_req; q += gridDim.x) {
		const int ctz = client_tz[q];
		for (int i = off[q]+threadIdx.x; i < off[q+1]; i += blockDim.x) {
			int d = cand_tz[i]-ctz;
			if (d < 0) d = -d;
			int v = d;
			if (metric == 1) {
				const int n = cand_node[i];
				double bw = 4000-100*d;
				double total = S.n_total[0][n]+bw;
				double mx = S.n_max[0][n];
				v = total > mx ? (int)(-bw*mx/total) : (int)(-bw);
			}
			if (usable && !usable[i]) v = INT32_MAX;
			val[i] = v;
		}
	}
}
int ns_pick_launch(int mode, int n_req, const int32_t *d_off, const int32_t *d_val, const int32_t *d_node, const int32_t *d_k,
		   int32_t *d_ov, int32_t *d_on, int32_t *d_oc, cudaStream_t stream, int max_c);

// ---------------------------------------------------------------- host-side simulation object
struct HostSim {
	Sim S;                         // device pointers
	SeqScratch scr;
	ParScratch par;
	std::vector<void *> allocs;
	size_t bytes = 0;
	uint32_t seed = 1;
	int n_creates = 0;
	// ensembles put all tables of all replicas into ONE allocation: a measuring pass (arena_mode 1) sums the sizes,
	// the real pass (arena_mode 2) carves them out of [arena, arena+arena_size)
	int arena_mode = 0;
	char *arena = nullptr; size_t arena_size = 0, arena_used = 0;
	std::vector<std::pair<char *, size_t>> zones;   // NS_REDZONE=1: (allocation base, payload bytes) of every table
	int cap_nodes = 0;                              // rows of the node tables (>= S.n_nodes: NS_OP_NEWNODE appends)
};

// Red zones (compute-sanitizer is closed on the GPU pool): with NS_REDZONE=1 every table gets NS_RZ bytes of a fixed pattern
// in front of it and behind it (behind: right after its last byte, then padding); ns_dev_check / ns_ens_check compare them
// after a run, so a store that runs off either end of a table shows up as a violation instead of as silent corruption.
#define NS_RZ 256
#define NS_RZ_BYTE 0xA5
static bool redzones_on(void){ static int on = -1; if (on < 0) { const char *e = getenv("NS_REDZONE"); on = e && *e && *e != '0'; } return on != 0; }
static int dev_alloc(HostSim &h, void **p, size_t bytes){
	if (bytes == 0) bytes = 8;
	const bool rz = redzones_on();
	const size_t total = rz ? bytes+2*NS_RZ : bytes;
	char *base = nullptr;
	if (h.arena_mode) {
		const size_t b = (total+255) & ~(size_t)255;
		if (h.arena_mode == 1) { *p = (void *)(uintptr_t)256; h.arena_used += b; return 0; }
		if (h.arena_used+b > h.arena_size) return NS_E_INTERNAL;
		base = h.arena+h.arena_used; h.arena_used += b; h.bytes += b;
	} else {
		CK(cudaMalloc((void **)&base, total));
		h.allocs.push_back(base);
		h.bytes += total;
	}
	if (rz) {
		CK(cudaMemset(base, NS_RZ_BYTE, NS_RZ)); CK(cudaMemset(base+NS_RZ+bytes, NS_RZ_BYTE, NS_RZ));
		h.zones.push_back({ base, bytes });
		*p = base+NS_RZ;
	} else *p = base;
	return 0;
}
__global__ void k_check_zones(const char *const *base, const size_t *bytes, int n, int *bad){
	for (int z = blockIdx.x; z < n; z += gridDim.x) {
		const unsigned char *a = (const unsigned char *)base[z], *b = a+NS_RZ+bytes[z];
		int v = 0;
		for (int i = threadIdx.x; i < NS_RZ; i += blockDim.x) v += (a[i] != NS_RZ_BYTE)+(b[i] != NS_RZ_BYTE);
		if (v) atomicAdd(bad, v);
	}
}
// number of red-zone bytes that no longer hold the pattern (0 = clean; -1: red zones were not enabled)
static int check_zones(const std::vector<std::pair<char *, size_t>> &zones, cudaStream_t stream, int *out){
	if (!redzones_on()) { *out = -1; return 0; }
	const int n = (int)zones.size();
	*out = 0;
	if (!n) return 0;
	std::vector<const char *> hb(n); std::vector<size_t> hs(n);
	for (int i = 0; i < n; i++) { hb[i] = zones[i].first; hs[i] = zones[i].second; }
	const char **db; size_t *ds; int *dbad;
	CK(cudaMalloc(&db, sizeof(char *)*n)); CK(cudaMalloc(&ds, sizeof(size_t)*n)); CK(cudaMalloc(&dbad, 4));
	CK(cudaMemcpy(db, hb.data(), sizeof(char *)*n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(ds, hs.data(), sizeof(size_t)*n, cudaMemcpyHostToDevice));
	CK(cudaMemset(dbad, 0, 4));
	k_check_zones<<<std::min(n, 1024), 64, 0, stream>>>(db, ds, n, dbad);
	CK(cudaMemcpyAsync(out, dbad, 4, cudaMemcpyDeviceToHost, stream));
	CK(cudaStreamSynchronize(stream));
	cudaFree(db); cudaFree(ds); cudaFree(dbad);
	return 0;
}
// short-lived device buffers of one call
struct DevTmp {
	std::vector<void *> p;
	template <typename T> cudaError_t get(T **out, size_t bytes){
		cudaError_t e = cudaMalloc((void **)out, bytes ? bytes : 8);
		if (e == cudaSuccess) p.push_back(*out);
		return e;
	}
	~DevTmp(){ for (void *q : p) cudaFree(q); }
};
template <typename T> static int dalloc(HostSim &h, T **p, size_t n){ return dev_alloc(h, (void **)p, n*sizeof(T)); }
static int fill_f64(double *p, u64 n, double v){ if (!n) return 0; k_fill_f64<<<(unsigned)std::min<u64>((n+255)/256, 4096), 256>>>(p, n, v); CK(cudaGetLastError()); return 0; }
static int fill_i32(int *p, u64 n, int v){ if (!n) return 0; k_fill_i32<<<(unsigned)std::min<u64>((n+255)/256, 4096), 256>>>(p, n, v); CK(cudaGetLastError()); return 0; }
static void free_sim(HostSim &h){ for (void *p : h.allocs) cudaFree(p); h.allocs.clear(); }

static u64 next_pow2(u64 x){ u64 p = 1; while (p < x) p <<= 1; return p; }

static int host_rec(std::vector<uint8_t> &rec, double now, uint8_t minor, uint32_t id, int bytes, uint32_t ival){
	// record.c:53-116 for the init-time records written on the host (R_NODE_CREATE, sim.c:123)
	uint8_t b[19]; int n = 0;
	b[n++] = 0; b[n++] = minor;
	b[n++] = id >> 24; b[n++] = id >> 16; b[n++] = id >> 8; b[n++] = id;
	uint32_t t = (uint32_t)(int)now;
	b[n++] = t >> 24; b[n++] = t >> 16; b[n++] = t >> 8; b[n++] = t;
	uint16_t ut = (uint16_t)(int)((now-((int)now))*1000);
	b[n++] = ut >> 8; b[n++] = ut;
	b[n++] = (uint8_t)(int8_t)bytes;
	if (bytes == 1) b[n++] = (uint8_t)ival;
	else { b[n++] = ival >> 24; b[n++] = ival >> 16; b[n++] = ival >> 8; b[n++] = ival; }
	rec.insert(rec.end(), b, b+n);
	return n;
}

// Build one simulation's device tables from a scenario script.
static int build_sim(HostSim &h, const void *scenario, size_t len, const ns_dev_opts &o, bool parallel_scratch, bool ens_cta = false){
	if (len < sizeof(ns_sc_header)) return NS_E_BAD_SCENARIO;
	ns_sc_header H;
	memcpy(&H, scenario, sizeof H);
	if (H.magic != NS_SC_MAGIC || H.version != NS_SC_VERSION || H.n_groups < 1 || H.n_nodes < 0) return NS_E_BAD_SCENARIO;
	if (H.n_resources < 0 || H.n_copies < 0 || H.n_ops < 0 || H.n_ondone < 0) return NS_E_BAD_SCENARIO;
	size_t need = sizeof H+(size_t)H.n_nodes*16+(size_t)H.n_resources*sizeof(ns_sc_resource)+(size_t)H.n_copies*sizeof(ns_sc_copy)+
		((size_t)H.n_groups+1)*4+(size_t)H.n_ops*sizeof(ns_sc_op)+(size_t)H.n_ondone*sizeof(ns_sc_ondone);
	if (len < need) return NS_E_BAD_SCENARIO;
	const char *p = (const char *)scenario+sizeof H;
	const double *max_up = (const double *)p; p += (size_t)H.n_nodes*8;
	const double *max_down = (const double *)p; p += (size_t)H.n_nodes*8;
	const ns_sc_resource *rs = (const ns_sc_resource *)p; p += (size_t)H.n_resources*sizeof *rs;
	const ns_sc_copy *cp = (const ns_sc_copy *)p; p += (size_t)H.n_copies*sizeof *cp;
	const int32_t *goff = (const int32_t *)p; p += ((size_t)H.n_groups+1)*4;
	// (the op table follows 4-byte offsets, so where it lies it may not be aligned for its 8-byte members: read it from a copy)
	std::vector<ns_sc_op> ops_copy((size_t)H.n_ops);
	if (H.n_ops) memcpy(ops_copy.data(), p, (size_t)H.n_ops*sizeof(ns_sc_op));
	const ns_sc_op *ops = ops_copy.data(); p += (size_t)H.n_ops*sizeof *ops;
	const ns_sc_ondone *od = (const ns_sc_ondone *)p; p += (size_t)H.n_ondone*sizeof *od;
	const bool has_clients = (H.flags & NS_SC_F_CLIENTS) != 0;
	const ns_sc_client *cl = (const ns_sc_client *)p;
	if (has_clients && len < need+(size_t)H.n_nodes*sizeof(ns_sc_client)) return NS_E_BAD_SCENARIO;
	if (H.flags & ~NS_SC_F_CLIENTS) return NS_E_BAD_SCENARIO;
	// group g = ops[goff[g] .. goff[g+1]): offsets must be monotone and inside the op table
	if (goff[0] < 0 || goff[H.n_groups] > H.n_ops) return NS_E_BAD_SCENARIO;
	for (int g = 0; g < H.n_groups; g++) if (goff[g] > goff[g+1]) return NS_E_BAD_SCENARIO;
	// the on-done table is binary-searched on the device: sorted, unique
	for (int i = 1; i < H.n_ondone; i++) if (od[i-1].flow_ord >= od[i].flow_ord) return NS_E_BAD_SCENARIO;
	int n_create = 0, n_sched = 0, n_play = 0;
	for (int i = 0; i < H.n_ops; i++) {
		if (ops[i].kind == NS_OP_CLOSE && ops[i].a < 0) return NS_E_BAD_SCENARIO;
		if (ops[i].kind == NS_OP_CREATE) {
			n_create++;
			if (ops[i].a < 0 || ops[i].a >= H.n_nodes || ops[i].b < 0 || ops[i].b >= H.n_nodes || ops[i].c < 0 || ops[i].c >= H.n_resources) return NS_E_BAD_SCENARIO;
		} else if (ops[i].kind == NS_OP_SCHED) {
			n_sched++;
			if (ops[i].a < 0 || ops[i].a >= H.n_groups) return NS_E_BAD_SCENARIO;
		} else if (ops[i].kind == NS_OP_PLAY1) {
			if (!has_clients || ops[i].a < 0 || ops[i].a >= H.n_nodes || ops[i].c < 0 || ops[i].c >= H.n_resources || (ops[i].b != 0 && ops[i].b != 1)) return NS_E_BAD_SCENARIO;
			if (cl[ops[i].a].kind != NS_CK_CLIENT) return NS_E_BAD_SCENARIO;
			n_play++;
		}
	}
	std::vector<int> srv_list;                                        // list_add at head: newest server first (p2p_common.h:112)
	if (has_clients) for (int i = H.n_nodes-1; i >= 0; i--) if (cl[i].kind == NS_CK_SERVER) srv_list.push_back(i);
	n_create += n_play*std::max(1, (int)srv_list.size()/2);           // a PLAY1 establishes up to nsvr/2 flows
	for (int i = 0; i < H.n_ondone; i++) if (od[i].group < 0 || od[i].group >= H.n_groups) return NS_E_BAD_SCENARIO;
	h.seed = H.seed; h.n_creates = n_create;

	Sim &S = h.S;
	memset(&S, 0, sizeof S);
	const int N = (H.flags & NS_SC_F_CLIENTS) ? H.n_nodes : std::max(H.n_nodes, o.cap_nodes);   // node tables: room for NS_OP_NEWNODE (host core)
	h.cap_nodes = N;
	S.n_nodes = H.n_nodes;
	S.cap_flows = o.cap_flows > 0 ? o.cap_flows : std::max(16, n_create);
	const int F = S.cap_flows;
	// a caller that creates flows beyond the script (the host core: the module IS the script) sizes the tables
	// through cap_flows / cap_ord / cap_ranges / cap_users; the static op counts are only the default
	const int base = std::max(n_create, F);
	S.cap_ord = o.cap_ord > 0 ? o.cap_ord : 2*base+16;
	S.cap_ranges = o.cap_ranges > 0 ? o.cap_ranges : 2*base+H.n_resources+H.n_copies+16;
	S.cap_res = H.n_resources+H.n_copies+std::max(base, S.cap_ord)+16;
	S.cap_users = o.cap_users > 0 ? o.cap_users : 2*n_sched+16;
	S.n_groups = H.n_groups; S.n_ops = H.n_ops; S.n_ondone = H.n_ondone; S.n_rids = H.n_resources;
	S.ring_cap = o.ring_cap ? next_pow2(std::max<u64>(o.ring_cap, NS_TILE)) : next_pow2(std::min<u64>(std::max<u64>(1u << 16, 128ull*F), 1ull << 28));
	S.n_ring_tiles = (int)(S.ring_cap >> NS_TILE_SHIFT);
	S.has_clients = has_clients; S.n_srv = (int)srv_list.size();
	S.n_l = 2*F+S.cap_users+(has_clients ? N : 0);                    // + one slot per node for a client's own USER event
	S.n_l_tiles = (S.n_l+NS_TILE-1)/NS_TILE;
	S.adj_arena_cap = 16*F+8*N+64;
	S.batch_cap = (int)std::min<u64>(S.ring_cap/2, 1ull << 30);
	S.want_rec = o.want_records; S.want_tr = o.want_trace;
	{	// engine paths that can be switched off for A/B runs: ns_dev_opts.reserved[0] or NS_FEATURES_OFF (bit mask)
		unsigned off = (unsigned)o.features_off;
		const char *env = getenv("NS_FEATURES_OFF");
		if (env) off |= (unsigned)strtoul(env, nullptr, 0);
		S.features = (int)(NS_FEAT_ALL & ~off);
		const char *e1 = getenv("NS_E1_MAX");
		S.e1_max = e1 ? atoi(e1) : 8;
		const char *hm = getenv("NS_HUB_MIN_DEG");
		S.hub_min_deg = hm ? atoi(hm) : 160;
	}
	S.rec_cap = o.want_records ? (o.record_cap ? o.record_cap : std::max<u64>(64u << 20, 64ull*N+4096ull*F)) : 0;
	S.tr_cap = o.want_trace ? (o.trace_cap ? o.trace_cap : (1u << 20)) : 0;

	int rc;
#define DA(ptr, n) if ((rc = dalloc(h, &(ptr), (size_t)(n)))) return rc
	DA(S.st, 1);
	for (int d = 0; d < 2; d++) { DA(S.n_max[d], N); DA(S.n_usage[d], N); DA(S.n_total[d], N); }
	DA(S.n_id, N); DA(S.n_res_head, N);
	DA(S.adj_off, 2*N); DA(S.adj_cap, 2*N); DA(S.adj_len, 2*N); DA(S.adj_live, 2*N); DA(S.adj_arena, S.adj_arena_cap);
	DA(S.f_ub, F); DA(S.f_speed[0], F); DA(S.f_speed[1], F); DA(S.f_delay, F);
	DA(S.f_peer[0], F); DA(S.f_peer[1], F); DA(S.f_id, F); DA(S.f_gen, F);
	DA(S.f_srng, F); DA(S.f_drng, F); DA(S.f_cnext, F); DA(S.f_cprev, F); DA(S.f_pos[0], F); DA(S.f_pos[1], F);
	DA(S.f_ord, F); DA(S.f_next_free, F); DA(S.f_alive, F); DA(S.f_drain_type, F);
	const int R = S.cap_ranges;
	DA(S.r_start, R); DA(S.r_total, R); DA(S.r_len, R); DA(S.r_lenc, R); DA(S.r_last, R);
	DA(S.r_prod, R); DA(S.r_res, R); DA(S.r_next, R); DA(S.r_prev, R); DA(S.r_cons, R); DA(S.r_next_free, R);
	const int Q = S.cap_res;
	if (o.cap_nodes > 0 && !(H.flags & NS_SC_F_CLIENTS)) DA(S.s_autodel, Q);   // a host core is driving (NS_OP_NEWNODE ...)
	DA(S.s_id, Q); DA(S.s_len, Q); DA(S.s_node, Q); DA(S.s_first, Q); DA(S.s_nprod, Q); DA(S.s_ncons, Q); DA(S.s_next, Q); DA(S.s_brate, Q);
	ns_sc_client *d_cl = nullptr; int *d_srv = nullptr;
	if (has_clients) {
		DA(d_cl, N); DA(d_srv, std::max(1, S.n_srv)); DA(S.n_state, N);
		DA(S.c_bit_rate, N); DA(S.c_last, N); DA(S.c_buffer_pos, N); DA(S.c_next_state, N); DA(S.c_trigger, N); DA(S.c_ue_type, N); DA(S.c_ue_rng, N);
	}
	S.c_cfg = d_cl; S.srv_list = d_srv;
	DA(S.q_time, S.ring_cap); DA(S.q_speed, S.ring_cap); DA(S.q_fd, S.ring_cap); DA(S.q_gen, S.ring_cap); DA(S.q_tmin, S.n_ring_tiles);
	DA(S.l_time, (size_t)S.n_l_tiles*NS_TILE); DA(S.l_seq, (size_t)S.n_l_tiles*NS_TILE); DA(S.l_tmin, S.n_l_tiles);
	DA(S.u_group, S.cap_users); DA(S.u_next_free, S.cap_users);
	ns_sc_op *d_ops; int *d_goff; ns_sc_ondone *d_od; uint32_t *d_rids;
	DA(d_ops, H.n_ops); DA(d_goff, H.n_groups+1); DA(d_od, H.n_ondone); DA(d_rids, H.n_resources);
	S.ops = d_ops; S.group_off = d_goff; S.ondone = d_od; S.rids = d_rids;
	DA(S.ord2flow, S.cap_ord);
	S.host_ops_cap = 256; DA(S.host_ops, S.host_ops_cap);
	DA(S.b_seq, S.batch_cap); DA(S.b_ref, 8);
	DA(h.scr.lb_seq, S.n_l); DA(h.scr.lb_ref, S.n_l); DA(h.scr.lb_tmp_seq, S.n_l); DA(h.scr.lb_tmp_ref, S.n_l);
	if (S.rec_cap) DA(S.rec, S.rec_cap);
	if (S.tr_cap) DA(S.tr, S.tr_cap);
	// parallel module init: group 0 made only of CREATE ops (flows established by the module's *_init)
	int n_init = 0;
	if (parallel_scratch && goff[1]-goff[0] >= 256) {
		n_init = goff[1]-goff[0];
		for (int i = goff[0]; i < goff[1]; i++) if (ops[i].kind != NS_OP_CREATE) { n_init = 0; break; }
		if (goff[0] != 0) n_init = 0;
		// the wave schedule assumes sim_establish_flow touches only its two end nodes; a RELAY create (its source is
		// still being filled by an earlier flow of the init) also re-times that producer (sim.c:86-91, range.c:16-79
		// reads the producer's speed and done time): such inits run on the sequential warp
		if (n_init) {
			std::unordered_set<u64> filling;
			for (int i = goff[0]; i < goff[1] && n_init; i++) {
				if (filling.count(((u64)(uint32_t)ops[i].a << 32) | (uint32_t)ops[i].c)) n_init = 0;
				filling.insert(((u64)(uint32_t)ops[i].b << 32) | (uint32_t)ops[i].c);
			}
		}
	}
	ParSizes psz;
	if (ens_cta) { psz.max_seg = 1u << 14; psz.min_ovf = 1u << 16; psz.ovf_per_flow = 8; psz.grid_max = 1; n_init = 0; }   // one CTA per replica (ns_ens.cu)
	psz.dry = h.arena_mode == 1;
	if ((rc = par_alloc(h.par, S, parallel_scratch, n_init, [&](void **pp, size_t b){ return dev_alloc(h, pp, b); }, psz))) return rc;
	if (n_init && (rc = par_plan_init(h.par, ops, n_init, N))) return rc;
#undef DA
	if (h.arena_mode == 1) return 0;                                  // measuring pass: sizes only

	// ---- host-side init: sim_create_node (sim.c:113-125), sim_node_new_resource (:129-140),
	//      sim_node_add_resource (:141-154)
	SimState st;
	memset(&st, 0, sizeof st);
	ns_rand_state g;
	ns_rand_seed_host(&g, H.seed);
	u64 draws = 0;
	std::vector<uint8_t> rec;
	std::vector<uint32_t> nid(N);
	std::unordered_set<uint32_t> used;
	used.reserve((size_t)N*2+16);
	for (int i = 0; i < H.n_nodes; i++) {
		uint32_t id;
		do { id = ns_rand_next_host(&g); draws++; } while (!used.insert(id).second);
		nid[i] = id;
		if (o.want_records && (S.features & NS_FEAT_INIT_REC)) host_rec(rec, 0.0, R_NODE_CREATE, id, 1, 0);
		if (has_clients && cl[i].kind == NS_CK_SERVER && o.want_records)
			host_rec(rec, 0.0, R_NODE_STATE, id, 1, NST_SERVER);     // p2p_new_server: sim_node_change_state(n, N_SERVER), p2p_common.h:106
	}
	std::vector<int> res_head(N, -1);
	std::vector<uint32_t> s_id; std::vector<long long> s_len; std::vector<int> s_node, s_first, s_next, s_brate;
	std::vector<long long> r_start, r_total; std::vector<double> r_len; std::vector<int> r_res;
	std::vector<uint32_t> rids(H.n_resources);
	auto store_find = [&](int node, uint32_t rid){ for (int r = res_head[node]; r >= 0; r = s_next[r]) if (s_id[r] == rid) return r; return -1; };
	auto add_full = [&](int node, uint32_t rid, long long L, int brate){
		int r = (int)s_id.size();
		s_id.push_back(rid); s_len.push_back(L); s_node.push_back(node); s_next.push_back(res_head[node]); res_head[node] = r; s_brate.push_back(brate);
		s_first.push_back((int)r_start.size());
		r_start.push_back(0); r_total.push_back(L); r_len.push_back((double)L); r_res.push_back(r);
		ns_rand_next_host(&g); draws++;                               // resource_new_range -> skip_list_insert
	};
	for (int i = 0; i < H.n_resources; i++) {
		if (rs[i].owner < 0 || rs[i].owner >= H.n_nodes) return NS_E_BAD_SCENARIO;
		uint32_t rid = ns_rand_next_host(&g); draws++;
		while (store_find(rs[i].owner, rid) >= 0) { rid = ns_rand_next_host(&g); draws++; }
		rids[i] = rid;
		add_full(rs[i].owner, rid, rs[i].len, rs[i].bit_rate);
	}
	for (int i = 0; i < H.n_copies; i++) {
		if (cp[i].node < 0 || cp[i].node >= H.n_nodes || cp[i].res < 0 || cp[i].res >= H.n_resources) return NS_E_BAD_SCENARIO;
		uint32_t rid = rids[cp[i].res];
		if (store_find(cp[i].node, rid) >= 0) continue;
		add_full(cp[i].node, rid, rs[cp[i].res].len, rs[cp[i].res].bit_rate);
	}
	const int nres0 = (int)s_id.size(), nrng0 = (int)r_start.size();
	if (rec.size() > S.rec_cap && o.want_records) return NS_E_RECORDS_LOST;

	// ---- fill and upload
	for (int d = 0; d < 2; d++) { if ((rc = fill_f64(S.n_usage[d], N, 0))) return rc; if ((rc = fill_f64(S.n_total[d], N, 0))) return rc; }
	if (H.n_nodes) {
		CK(cudaMemcpy(S.n_max[0], max_up, (size_t)H.n_nodes*8, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(S.n_max[1], max_down, (size_t)H.n_nodes*8, cudaMemcpyHostToDevice));
	}
	CK(cudaMemcpy(S.n_id, nid.data(), (size_t)N*4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(S.n_res_head, res_head.data(), (size_t)N*4, cudaMemcpyHostToDevice));
	if ((rc = fill_i32(S.adj_off, 2*N, 0)) || (rc = fill_i32(S.adj_cap, 2*N, 0)) || (rc = fill_i32(S.adj_len, 2*N, 0)) || (rc = fill_i32(S.adj_live, 2*N, 0))) return rc;
	if (S.s_autodel) CK(cudaMemset(S.s_autodel, 0, (size_t)S.cap_res));
	CK(cudaMemset(S.f_gen, 0, (size_t)F*4));
	CK(cudaMemset(S.f_alive, 0, F));
	if ((rc = fill_f64(S.q_time, S.ring_cap, INFINITY)) || (rc = fill_f64(S.q_tmin, S.n_ring_tiles, INFINITY))) return rc;
	if ((rc = fill_f64(S.l_time, (u64)S.n_l_tiles*NS_TILE, INFINITY)) || (rc = fill_f64(S.l_tmin, S.n_l_tiles, INFINITY))) return rc;
	if ((rc = fill_i32(S.ord2flow, S.cap_ord, -1))) return rc;
	if (nres0) {
		CK(cudaMemcpy(S.s_id, s_id.data(), (size_t)nres0*4, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(S.s_len, s_len.data(), (size_t)nres0*8, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(S.s_node, s_node.data(), (size_t)nres0*4, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(S.s_first, s_first.data(), (size_t)nres0*4, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(S.s_next, s_next.data(), (size_t)nres0*4, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(S.s_brate, s_brate.data(), (size_t)nres0*4, cudaMemcpyHostToDevice));
		CK(cudaMemset(S.s_nprod, 0, (size_t)nres0*4)); CK(cudaMemset(S.s_ncons, 0, (size_t)nres0*4));
		CK(cudaMemcpy(S.r_start, r_start.data(), (size_t)nrng0*8, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(S.r_total, r_total.data(), (size_t)nrng0*8, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(S.r_len, r_len.data(), (size_t)nrng0*8, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(S.r_res, r_res.data(), (size_t)nrng0*4, cudaMemcpyHostToDevice));
		if ((rc = fill_f64(S.r_lenc, nrng0, 0)) || (rc = fill_f64(S.r_last, nrng0, 0))) return rc;
		if ((rc = fill_i32(S.r_prod, nrng0, -1)) || (rc = fill_i32(S.r_next, nrng0, -1)) || (rc = fill_i32(S.r_prev, nrng0, -1)) || (rc = fill_i32(S.r_cons, nrng0, -1))) return rc;
	}
	if (has_clients) {
		CK(cudaMemcpy(d_cl, cl, (size_t)N*sizeof *cl, cudaMemcpyHostToDevice));
		if (S.n_srv) CK(cudaMemcpy(d_srv, srv_list.data(), (size_t)S.n_srv*4, cudaMemcpyHostToDevice));
		std::vector<int> st0(N);
		for (int i = 0; i < N; i++) st0[i] = cl[i].kind == NS_CK_SERVER ? NST_SERVER : cl[i].kind == NS_CK_CLIENT ? NST_IDLE : NST_OFFLINE;   // p2p.c:168, data.h:226-236
		CK(cudaMemcpy(S.n_state, st0.data(), (size_t)N*4, cudaMemcpyHostToDevice));
		CK(cudaMemset(S.c_bit_rate, 0, (size_t)N*8)); CK(cudaMemset(S.c_last, 0, (size_t)N*8)); CK(cudaMemset(S.c_buffer_pos, 0, (size_t)N*4));
		CK(cudaMemcpy(S.c_next_state, st0.data(), (size_t)N*4, cudaMemcpyHostToDevice));
		if ((rc = fill_i32(S.c_trigger, N, -1)) || (rc = fill_i32(S.c_ue_type, N, 0)) || (rc = fill_i32(S.c_ue_rng, N, -1))) return rc;
	}
	CK(cudaMemcpy(d_ops, ops, (size_t)H.n_ops*sizeof *ops, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(d_goff, goff, (size_t)(H.n_groups+1)*4, cudaMemcpyHostToDevice));
	if (H.n_ondone) CK(cudaMemcpy(d_od, od, (size_t)H.n_ondone*sizeof *od, cudaMemcpyHostToDevice));
	if (H.n_resources) CK(cudaMemcpy(d_rids, rids.data(), (size_t)H.n_resources*4, cudaMemcpyHostToDevice));
	if (!rec.empty()) CK(cudaMemcpy(S.rec, rec.data(), rec.size(), cudaMemcpyHostToDevice));

	memcpy(st.rnd, g.st, sizeof st.rnd); st.rf = g.f; st.rr = g.r;
	st.free_flow = st.free_range = st.free_user = -1;
	st.pending_close = -1;
	st.hi_range = nrng0; st.hi_res = nres0;
	st.rec_len = rec.size();
	st.cnt[NS_CNT_DRAWS] = draws;
	st.seed = H.seed;
	CK(cudaMemcpy(S.st, &st, sizeof st, cudaMemcpyHostToDevice));
	return 0;
}

// ---------------------------------------------------------------- C ABI: single simulation
struct ns_dev {
	int device = 0, mode = 0;
	HostSim h;
	Sim *d_sim = nullptr; SeqScratch *d_scr = nullptr;
	cudaStream_t stream = nullptr, copy_stream = nullptr;   // records leave on a side stream (ns_dev_records_async)
	cudaEvent_t e0 = nullptr, e1 = nullptr, e_run = nullptr;
	uint64_t launches = 0;
	int sm_count = 0;
	size_t rec_total_host = 0;      // record bytes in the buffer when the last run returned
	size_t rec_base_host = 0;       // ... and the stream offset of the buffer's first byte
	u64 *d_hash = nullptr;
	void *view_buf = nullptr; size_t view_cap = 0;   // ns_dev_host_view
};

extern "C" int ns_dev_count(void){
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

static int read_state(ns_dev *d, SimState *st){
	CK(cudaMemcpyAsync(st, d->h.S.st, sizeof *st, cudaMemcpyDeviceToHost, d->stream));
	CK(cudaStreamSynchronize(d->stream));
	return 0;
}

static int dev_init(ns_dev *d, const void *scenario, size_t len, const ns_dev_opts &o);
extern "C" void ns_dev_destroy(ns_dev *d);
extern "C" int ns_dev_create(ns_dev **out, const void *scenario, size_t len, const ns_dev_opts *opts){
	if (!out || !scenario) return NS_E_ARG;
	*out = nullptr;
	ns_dev_opts o;
	memset(&o, 0, sizeof o);
	if (opts) o = *opts;
	if (ns_dev_count() <= o.device) return NS_E_NO_DEVICE;
	CK(cudaSetDevice(o.device));
	ns_dev *d = new ns_dev();
	d->device = o.device; d->mode = o.mode;
	int rc = dev_init(d, scenario, len, o);
	if (rc) { ns_dev_destroy(d); return rc; }          // releases whatever dev_init had built
	*out = d;
	return 0;
}
static int dev_init(ns_dev *d, const void *scenario, size_t len, const ns_dev_opts &o){
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, o.device));
	d->sm_count = prop.multiProcessorCount;
	int rc = build_sim(d->h, scenario, len, o, o.mode != NS_MODE_SEQ);
	if (rc) return rc;
	CK(cudaStreamCreate(&d->stream)); CK(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
	CK(cudaEventCreate(&d->e0)); CK(cudaEventCreate(&d->e1)); CK(cudaEventCreateWithFlags(&d->e_run, cudaEventDisableTiming));
	CK(cudaMalloc(&d->d_sim, sizeof(Sim))); CK(cudaMalloc(&d->d_scr, sizeof(SeqScratch))); CK(cudaMalloc(&d->d_hash, 8));
	CK(cudaMemcpy(d->d_sim, &d->h.S, sizeof(Sim), cudaMemcpyHostToDevice));
	CK(cudaMemcpy(d->d_scr, &d->h.scr, sizeof(SeqScratch), cudaMemcpyHostToDevice));
	CK(cudaDeviceSynchronize());
	if ((o.mode != NS_MODE_SEQ && (rc = par_preload())) || (rc = seq_preload())) return rc;   // load the kernels now, not inside the first timed run
	// module init: group 0 of the script (first sim_establish_flow calls, scheduled USER events)
	if (d->mode == NS_MODE_SEQ) {
		if ((rc = seq_launch_init_groups(d->d_sim, 1, d->stream))) return rc;
		d->launches++;
	} else {
		rc = par_init_group0(d->h.S, d->d_sim, d->h.par, d->sm_count, d->stream, &d->launches);
		if (rc) return rc;
	}
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(d->stream));
	return 0;
}

extern "C" int ns_dev_run(ns_dev *d, uint64_t max_events, float *device_ms){
	if (!d) return NS_E_ARG;
	CK(cudaSetDevice(d->device));
	CK(cudaMemcpyAsync(&d->h.S.st->max_events, &max_events, 8, cudaMemcpyHostToDevice, d->stream));
	{ int zero = 0; CK(cudaMemcpyAsync(&d->h.S.st->host_stop, &zero, 4, cudaMemcpyHostToDevice, d->stream)); }
	CK(cudaEventRecord(d->e0, d->stream));
	if (d->mode == NS_MODE_SEQ) {
		int rc = seq_launch_run(d->d_sim, d->d_scr, 1, d->stream);
		if (rc) return rc;
		d->launches++;
	} else {
		int rc = par_run(d->h.S, d->d_sim, d->d_scr, d->h.par, d->sm_count, d->stream, &d->launches);
		if (rc) return rc;
	}
	CK(cudaEventRecord(d->e1, d->stream));
	CK(cudaStreamSynchronize(d->stream));
	if (device_ms) CK(cudaEventElapsedTime(device_ms, d->e0, d->e1));
	SimState st; int rc = read_state(d, &st); if (rc) return rc;
	d->rec_total_host = st.rec_len;
	d->rec_base_host = st.rec_base;
	if (st.status == 0 && st.host_stop == 2) return NS_E_RECORDS_FULL;   // soft: drain the records, run again
	return st.status;
}

extern "C" int ns_dev_status(ns_dev *d){
	if (!d) return NS_E_ARG;
	SimState st; int rc = read_state(d, &st); if (rc) return rc;
	return st.status;
}
extern "C" double ns_dev_now(ns_dev *d){
	if (!d) return 0;
	SimState st; if (read_state(d, &st)) return 0;
	return st.now;
}
extern "C" int ns_dev_counters(ns_dev *d, uint64_t out[NS_CNT_COUNT]){
	if (!d || !out) return NS_E_ARG;
	SimState st; int rc = read_state(d, &st); if (rc) return rc;
	for (int i = 0; i < NS_CNT_COUNT; i++) out[i] = st.cnt[i];
	out[NS_CNT_SEQ] = st.tail;
	out[NS_CNT_DRAWS] = st.cnt[NS_CNT_DRAWS]+st.burns;
	return 0;
}
extern "C" int ns_dev_records(ns_dev *d, void *dst, size_t cap, size_t offset, size_t *len, size_t *total){
	if (!d) return NS_E_ARG;
	SimState st; int rc = read_state(d, &st); if (rc) return rc;
	const size_t end = st.rec_base+st.rec_len;                   // offsets are positions in the whole stream
	if (total) *total = end;
	size_t n = 0;
	if (offset >= st.rec_base && offset < end) n = std::min<size_t>(cap, end-offset);
	if (n && dst) CK(cudaMemcpy(dst, d->h.S.rec+(offset-st.rec_base), n, cudaMemcpyDeviceToHost));
	if (len) *len = n;
	return 0;
}
// everything produced so far has been copied out: the buffer starts over (after NS_E_RECORDS_FULL)
extern "C" int ns_dev_records_drain(ns_dev *d){
	if (!d) return NS_E_ARG;
	CK(cudaStreamSynchronize(d->copy_stream));
	SimState st; int rc = read_state(d, &st); if (rc) return rc;
	u64 nb = st.rec_base+st.rec_len, zero = 0;
	CK(cudaMemcpy(&d->h.S.st->rec_base, &nb, 8, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(&d->h.S.st->rec_len, &zero, 8, cudaMemcpyHostToDevice));
	d->rec_base_host = nb; d->rec_total_host = 0;
	return 0;
}
// Records go back to the host on a side stream: the copy of the span produced by the last run overlaps
// the next run (the engine only appends beyond it).  dst should be pinned host memory.
extern "C" int ns_dev_records_async(ns_dev *d, void *dst, size_t cap, size_t offset, size_t *len, size_t *total){
	if (!d || !dst) return NS_E_ARG;
	const size_t base = d->rec_base_host, end = base+d->rec_total_host;
	size_t n = 0;
	if (offset >= base && offset < end) n = std::min<size_t>(cap, end-offset);
	if (n) CK(cudaMemcpyAsync(dst, d->h.S.rec+(offset-base), n, cudaMemcpyDeviceToHost, d->copy_stream));
	if (len) *len = n;
	if (total) *total = end;
	return 0;
}
extern "C" int ns_dev_records_wait(ns_dev *d){
	if (!d) return NS_E_ARG;
	CK(cudaStreamSynchronize(d->copy_stream));
	return 0;
}
static int stream_hash(ns_dev *d, u64 rec_len, u64 *out){
	CK(cudaMemsetAsync(d->d_hash, 0, 8, d->stream));
	if (rec_len) { k_stream_hash<<<296, 256, 0, d->stream>>>(d->h.S.rec, rec_len, d->d_hash); d->launches++; }
	CK(cudaMemcpyAsync(out, d->d_hash, 8, cudaMemcpyDeviceToHost, d->stream));
	CK(cudaStreamSynchronize(d->stream));
	*out = ns_mix64(*out ^ rec_len);
	return 0;
}
extern "C" int ns_dev_digest(ns_dev *d, ns_digest *out){
	if (!d || !out) return NS_E_ARG;
	SimState st; int rc = read_state(d, &st); if (rc) return rc;
	out->seed = st.seed; out->events = st.cnt[NS_CNT_EVENTS]; out->solves = st.cnt[NS_CNT_SOLVES];
	out->record_bytes = st.rec_base+st.rec_len; out->t_end = st.now; out->status = st.status;
	return stream_hash(d, st.rec_len, (u64 *)&out->record_fnv1a);   // digest of the bytes still in the buffer (the whole stream unless drained)
}
extern "C" int ns_dev_trace(ns_dev *d, ns_trace_row32 *dst, size_t cap_rows, size_t *rows){
	if (!d) return NS_E_ARG;
	SimState st; int rc = read_state(d, &st); if (rc) return rc;
	size_t n = std::min<size_t>(cap_rows, st.tr_len);
	if (n && dst) CK(cudaMemcpy(dst, d->h.S.tr, n*sizeof *dst, cudaMemcpyDeviceToHost));
	if (rows) *rows = st.tr_len;
	return 0;
}
extern "C" int ns_dev_flow_state(ns_dev *d, double *out4, size_t max_flows, size_t *n){
	if (!d || !out4) return NS_E_ARG;
	SimState st; int rc = read_state(d, &st); if (rc) return rc;
	size_t k = std::min<size_t>(max_flows, (size_t)st.n_ord);
	if (k) {
		double *tmp; DevTmp bufs; CK(bufs.get(&tmp, k*32));
		k_flow_state<<<(unsigned)((k+255)/256), 256, 0, d->stream>>>(d->d_sim, tmp, (int)k); d->launches++;
		CK(cudaMemcpyAsync(out4, tmp, k*32, cudaMemcpyDeviceToHost, d->stream));
		CK(cudaStreamSynchronize(d->stream));
	}
	if (n) *n = k;
	return 0;
}
extern "C" int ns_dev_node_state(ns_dev *d, double *out4, size_t max_nodes, size_t *n){
	if (!d || !out4) return NS_E_ARG;
	size_t k = std::min<size_t>(max_nodes, (size_t)d->h.S.n_nodes);
	if (k) {
		double *tmp; DevTmp bufs; CK(bufs.get(&tmp, k*32));
		k_node_state<<<(unsigned)((k+255)/256), 256, 0, d->stream>>>(d->d_sim, tmp, (int)k); d->launches++;
		CK(cudaMemcpyAsync(out4, tmp, k*32, cudaMemcpyDeviceToHost, d->stream));
		CK(cudaStreamSynchronize(d->stream));
	}
	if (n) *n = k;
	return 0;
}
// batched source selection against the live simulation state (the score map + the selection, on the device)
extern "C" int ns_dev_pick_live(ns_dev *d, int mode, int metric, int n_req, const int32_t *off, const int32_t *cand_node,
		const int32_t *cand_tz, const int32_t *client_tz, const uint8_t *usable, const int32_t *k,
		int32_t *out_val, int32_t *out_node, int32_t *out_cnt, float *device_ms){
	if (!d || n_req < 0 || !off || (mode != 1 && mode != 2) || (metric != 0 && metric != 1)) return NS_E_ARG;
	if (n_req == 0) return 0;
	if (!client_tz || !out_cnt || off[0] < 0) return NS_E_ARG;
	for (int q = 0; q < n_req; q++) if (off[q+1] < off[q]) return NS_E_ARG;
	const int total = off[n_req];
	if (total > 0 && (!cand_node || !cand_tz || !out_val || !out_node)) return NS_E_ARG;
	CK(cudaSetDevice(d->device));
	for (int i = 0; i < total; i++) if (cand_node[i] < 0 || cand_node[i] >= d->h.S.n_nodes) return NS_E_ARG;
	const size_t tb = (size_t)std::max(total, 1)*4;
	int32_t *d_off, *d_node, *d_tz, *d_ctz, *d_val, *d_k, *d_ov, *d_on, *d_oc; uint8_t *d_us = nullptr;
	DevTmp tmp;                                                          // freed on every way out
	CK(tmp.get(&d_off, (size_t)(n_req+1)*4)); CK(tmp.get(&d_node, tb)); CK(tmp.get(&d_tz, tb)); CK(tmp.get(&d_ctz, (size_t)n_req*4));
	CK(tmp.get(&d_val, tb)); CK(tmp.get(&d_k, (size_t)n_req*4)); CK(tmp.get(&d_ov, tb)); CK(tmp.get(&d_on, tb)); CK(tmp.get(&d_oc, (size_t)n_req*4));
	CK(cudaMemcpyAsync(d_off, off, (size_t)(n_req+1)*4, cudaMemcpyHostToDevice, d->stream));
	CK(cudaMemcpyAsync(d_ctz, client_tz, (size_t)n_req*4, cudaMemcpyHostToDevice, d->stream));
	if (total > 0) {
		CK(cudaMemcpyAsync(d_node, cand_node, tb, cudaMemcpyHostToDevice, d->stream));
		CK(cudaMemcpyAsync(d_tz, cand_tz, tb, cudaMemcpyHostToDevice, d->stream));
		if (usable) { CK(tmp.get(&d_us, (size_t)total)); CK(cudaMemcpyAsync(d_us, usable, (size_t)total, cudaMemcpyHostToDevice, d->stream)); }
	}
	if (k) CK(cudaMemcpyAsync(d_k, k, (size_t)n_req*4, cudaMemcpyHostToDevice, d->stream)); else CK(cudaMemsetAsync(d_k, 0, (size_t)n_req*4, d->stream));
	CK(cudaEventRecord(d->e0, d->stream));
	k_pick_metric<<<std::min(n_req, 1184), 128, 0, d->stream>>>(d->d_sim, metric, n_req, d_off, d_node, d_tz, d_ctz, d_us, d_val);
	CK(cudaGetLastError());
	int max_c = 0;
	for (int q = 0; q < n_req; q++) max_c = std::max(max_c, off[q+1]-off[q]);
	int rc = ns_pick_launch(mode, n_req, d_off, d_val, d_node, d_k, d_ov, d_on, d_oc, d->stream, max_c);
	d->launches += 2;
	CK(cudaEventRecord(d->e1, d->stream));
	if (rc) return rc;
	if (total > 0) { CK(cudaMemcpyAsync(out_val, d_ov, tb, cudaMemcpyDeviceToHost, d->stream)); CK(cudaMemcpyAsync(out_node, d_on, tb, cudaMemcpyDeviceToHost, d->stream)); }
	CK(cudaMemcpyAsync(out_cnt, d_oc, (size_t)n_req*4, cudaMemcpyDeviceToHost, d->stream));
	CK(cudaStreamSynchronize(d->stream));
	if (device_ms) CK(cudaEventElapsedTime(device_ms, d->e0, d->e1));
	return 0;
}

// ---- host-callback mode (modules whose own handlers must run on the host) ----
extern "C" int ns_dev_set_host_mask(ns_dev *d, uint32_t mask){
	if (!d) return NS_E_ARG;
	CK(cudaMemcpy(&d->h.S.st->host_mask, &mask, 4, cudaMemcpyHostToDevice));
	return 0;
}
extern "C" int ns_dev_run_host(ns_dev *d, ns_host_event *ev){
	if (!d || !ev) return NS_E_ARG;
	CK(cudaSetDevice(d->device));
	int rc = seq_launch_run(d->d_sim, d->d_scr, 1, d->stream, -1);   // callbacks serialise the run anyway; -1: the kernel clears host_stop / max_events itself
	if (rc) return rc;
	d->launches++;
	SimState st; rc = read_state(d, &st); if (rc) return rc;
	ev->type = st.le_type; ev->flow_ord = st.le_ord; ev->dir = st.le_dir; ev->token = st.le_token;
	ev->time = st.le_time; ev->speed = st.le_speed; ev->stopped = st.host_stop; ev->status = st.status;
	ev->now = st.now; ev->exited = st.exit_flag;
	return st.status;
}
extern "C" int ns_dev_exec(ns_dev *d, ns_sc_op *ops, int n){
	if (!d || !ops || n < 0 || n >= d->h.S.host_ops_cap) return NS_E_ARG;   // (one slot carries the status back)
	if (n == 0) return 0;
	int new_nodes = 0;
	for (int i = 0; i < n; i++) {                                        // the device indexes its tables with these
		const ns_sc_op &op = ops[i];
		if (op.kind == NS_OP_CREATE) { if (op.a < 0 || op.a >= d->h.S.n_nodes || op.b < 0 || op.b >= d->h.S.n_nodes || op.c < 0 || op.c >= d->h.S.n_rids) return NS_E_ARG; }
		else if (op.kind == NS_OP_CLOSE) { if (op.a < 0) return NS_E_ARG; }
		else if (op.kind == NS_OP_CANCEL) { if (op.a >= d->h.S.cap_users || (op.a < 0 && (-op.a-1 >= i || ops[-op.a-1].kind != NS_OP_SCHED))) return NS_E_ARG; }
		else if (op.kind == NS_OP_NEWNODE) { if (op.a != d->h.S.n_nodes+new_nodes || op.a >= d->h.cap_nodes) return op.a >= d->h.cap_nodes ? NS_E_TABLE_FULL : NS_E_ARG; new_nodes++; }
		else if (op.kind == NS_OP_NEWRES) { if (op.a < 0 || op.a >= d->h.S.n_nodes+new_nodes || op.start < 0) return NS_E_ARG; }
		else if (op.kind == NS_OP_CREATE_RID) { if (op.a < 0 || op.a >= d->h.S.n_nodes+new_nodes || op.b < 0 || op.b >= d->h.S.n_nodes+new_nodes) return NS_E_ARG; }
		else if (op.kind == NS_OP_DELRES || op.kind == NS_OP_AUTODEL) { if (op.a < 0 || op.a >= d->h.S.cap_res || !d->h.S.s_autodel) return NS_E_ARG; }
		else if (op.kind == NS_OP_RANGE_SET) { if (op.a < 0 || op.a >= d->h.S.cap_ranges) return NS_E_ARG; }
		else if (op.kind != NS_OP_SCHED && op.kind != NS_OP_END) return NS_E_ARG;
	}
	CK(cudaSetDevice(d->device));
	CK(cudaMemcpyAsync(d->h.S.host_ops, ops, (size_t)n*sizeof *ops, cudaMemcpyHostToDevice, d->stream));
	{ int rc = seq_launch_exec_ops(d->d_sim, n, d->stream); if (rc) return rc; }
	d->launches++;
	std::vector<ns_sc_op> back((size_t)n+1);
	CK(cudaMemcpyAsync(back.data(), d->h.S.host_ops, ((size_t)n+1)*sizeof *ops, cudaMemcpyDeviceToHost, d->stream));
	if (new_nodes) {                                                     // the node count is part of the device's Sim
		d->h.S.n_nodes += new_nodes;
		CK(cudaMemcpyAsync(&d->d_sim->n_nodes, &d->h.S.n_nodes, sizeof(int), cudaMemcpyHostToDevice, d->stream));
	}
	CK(cudaStreamSynchronize(d->stream));
	memcpy(ops, back.data(), (size_t)n*sizeof *ops);
	return back[n].kind;                                                 // the simulation's status after the ops (host_ops_status)
}

// ---- the host core's view of the mutable state (every live flow, every range) and the shared random() stream ----
__global__ void k_host_view(const Sim *sp, ns_view_hdr *hdr, double *nodes4, ns_view_flow *fl, int max_f, ns_view_range *rg, int max_r, int max_n){
	const Sim &S = *sp;
	const int hf = min(S.st->hi_flow, max_f), hr = min(S.st->hi_range, max_r), nn = min(S.n_nodes, max_n);
	const int gt = blockIdx.x*blockDim.x+threadIdx.x, gn = gridDim.x*blockDim.x;
	if (gt == 0) { hdr->hi_flow = S.st->hi_flow; hdr->hi_range = S.st->hi_range; hdr->n_nodes = S.n_nodes; hdr->pad = 0; }
	for (int n = gt; n < nn; n += gn) { nodes4[4*n] = S.n_usage[0][n]; nodes4[4*n+1] = S.n_usage[1][n]; nodes4[4*n+2] = S.n_total[0][n]; nodes4[4*n+3] = S.n_total[1][n]; }
	for (int f = gt; f < hf; f += gn) {
		ns_view_flow v;
		memset(&v, 0, sizeof v);
		v.slot = -1;
		if (S.f_alive[f]) {
			v.ord = S.f_ord[f]; v.slot = f; v.src = S.f_peer[0][f]; v.dst = S.f_peer[1][f]; v.srng = S.f_srng[f]; v.drng = S.f_drng[f];
			v.drain_type = S.f_drain_type[f]; v.flow_id = S.f_id[f];
			v.speed[0] = S.f_speed[0][f]; v.speed[1] = S.f_speed[1][f]; v.done_t = S.l_time[2*f]; v.drain_t = S.l_time[2*f+1];
		}
		fl[f] = v;
	}
	for (int r = gt; r < hr; r += gn) {
		ns_view_range v;
		memset(&v, 0, sizeof v);
		v.slot = -1;
		if (S.r_res[r] >= 0) {                                           // (-1: a free slot, range_merge_with_next / res_delete)
			v.slot = r; v.res = S.r_res[r]; v.producer_ord = S.r_prod[r] >= 0 ? S.f_ord[S.r_prod[r]] : -1; v.next = S.r_next[r]; v.prev = S.r_prev[r];
			v.start = S.r_start[r]; v.total_len = S.r_total[r]; v.len = S.r_len[r]; v.lenc = S.r_lenc[r]; v.last_update = S.r_last[r];
		}
		rg[r] = v;
	}
}
extern "C" int ns_dev_host_view(ns_dev *d, const ns_view_hdr **hdr, const double **nodes4, const ns_view_flow **flows, const ns_view_range **ranges){
	if (!d || !hdr || !nodes4 || !flows || !ranges) return NS_E_ARG;
	CK(cudaSetDevice(d->device));
	// called once per event in host-callback mode: one pinned, device-mapped buffer for the life of the handle, sized for the
	// tables' capacities; the kernel writes it over PCIe, the host reads it after one stream synchronisation
	const size_t o_nodes = 64, o_flows = o_nodes+sizeof(double)*4*(size_t)d->h.cap_nodes;
	const size_t o_ranges = o_flows+sizeof(ns_view_flow)*(size_t)d->h.S.cap_flows, total = o_ranges+sizeof(ns_view_range)*(size_t)d->h.S.cap_ranges;
	if (!d->view_buf) {
		CK(cudaHostAlloc(&d->view_buf, total, cudaHostAllocMapped));
		d->view_cap = total;
	}
	char *base = (char *)d->view_buf;
	void *dbase = nullptr;
	CK(cudaHostGetDevicePointer(&dbase, d->view_buf, 0));
	char *db = (char *)dbase;
	k_host_view<<<64, 256, 0, d->stream>>>(d->d_sim, (ns_view_hdr *)db, (double *)(db+o_nodes), (ns_view_flow *)(db+o_flows), d->h.S.cap_flows,
	                                        (ns_view_range *)(db+o_ranges), d->h.S.cap_ranges, d->h.cap_nodes);
	d->launches++;
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(d->stream));
	*hdr = (const ns_view_hdr *)base; *nodes4 = (const double *)(base+o_nodes); *flows = (const ns_view_flow *)(base+o_flows); *ranges = (const ns_view_range *)(base+o_ranges);
	return 0;
}
extern "C" int ns_dev_rng_get(ns_dev *d, ns_rand_state *g){
	if (!d || !g) return NS_E_ARG;
	SimState st; int rc = read_state(d, &st); if (rc) return rc;
	if (st.burns) {                                                      // burned draws are applied lazily (ns_rand_jump)
		ns_rand_jump(st.rnd, st.rf, st.rr, st.burns);
		st.cnt[NS_CNT_DRAWS] += st.burns; st.burns = 0;
		CK(cudaMemcpy(&d->h.S.st->burns, &st.burns, sizeof st.burns, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(&d->h.S.st->rnd, st.rnd, sizeof st.rnd+2*sizeof(int32_t), cudaMemcpyHostToDevice));
		CK(cudaMemcpy(&d->h.S.st->cnt[NS_CNT_DRAWS], &st.cnt[NS_CNT_DRAWS], 8, cudaMemcpyHostToDevice));
	}
	memcpy(g->st, st.rnd, sizeof st.rnd); g->f = st.rf; g->r = st.rr;
	return 0;
}
extern "C" int ns_dev_rng_set(ns_dev *d, const ns_rand_state *g){
	if (!d || !g) return NS_E_ARG;
	CK(cudaSetDevice(d->device));
	int32_t w[33];
	memcpy(w, g->st, sizeof g->st); w[31] = g->f; w[32] = g->r;
	CK(cudaMemcpy(&d->h.S.st->rnd, w, sizeof w, cudaMemcpyHostToDevice));   // rnd[31], rf, rr are laid out together (ns_sim.cuh)
	return 0;
}
extern "C" int ns_dev_flow_info(ns_dev *d, int ord, ns_flow_info *out){
	if (!d || !out) return NS_E_ARG;
	FlowInfoDev *tmp, h;
	DevTmp bufs;
	CK(bufs.get(&tmp, sizeof *tmp));
	k_flow_info<<<1, 1, 0, d->stream>>>(d->d_sim, ord, tmp); d->launches++;
	CK(cudaMemcpyAsync(&h, tmp, sizeof h, cudaMemcpyDeviceToHost, d->stream));
	CK(cudaStreamSynchronize(d->stream));
	out->alive = h.alive; out->src = h.src; out->dst = h.dst; out->flow_id = h.flow_id; out->resource_id = h.resource_id;
	out->ub = h.ub; out->delay = h.delay; out->speed[0] = h.speed0; out->speed[1] = h.speed1;
	out->drng_start = h.drng_start; out->drng_len = h.drng_len;
	return 0;
}
extern "C" int ns_dev_node_ranges(ns_dev *d, int node, uint32_t rid, ns_range_info *out, int max, int *n){
	if (!d || !out || !n || node < 0 || node >= d->h.S.n_nodes || max <= 0) return NS_E_ARG;
	RangeInfoDev *tmp; int *dn;
	DevTmp bufs;
	CK(bufs.get(&tmp, sizeof(RangeInfoDev)*max)); CK(bufs.get(&dn, 4));
	k_node_ranges<<<1, 1, 0, d->stream>>>(d->d_sim, node, rid, tmp, max, dn); d->launches++;
	std::vector<RangeInfoDev> h(max);
	CK(cudaMemcpyAsync(h.data(), tmp, sizeof(RangeInfoDev)*max, cudaMemcpyDeviceToHost, d->stream));
	CK(cudaMemcpyAsync(n, dn, 4, cudaMemcpyDeviceToHost, d->stream));
	CK(cudaStreamSynchronize(d->stream));
	for (int i = 0; i < *n; i++) { out[i].start = h[i].start; out[i].total_len = h[i].total_len; out[i].len = h[i].len; out[i].producer_ord = h[i].producer_ord; out[i].pad = 0; }
	return 0;
}
extern "C" int ns_dev_node_id(ns_dev *d, int node, uint32_t *id){
	if (!d || !id || node < 0 || node >= d->h.S.n_nodes) return NS_E_ARG;
	CK(cudaMemcpy(id, d->h.S.n_id+node, 4, cudaMemcpyDeviceToHost));
	return 0;
}
extern "C" int ns_dev_resource_id(ns_dev *d, int res, uint32_t *id){
	if (!d || !id || res < 0 || res >= d->h.S.n_rids) return NS_E_ARG;
	CK(cudaMemcpy(id, d->h.S.rids+res, 4, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int ns_dev_check(ns_dev *d, int *violations){
	if (!d || !violations) return NS_E_ARG;
	CK(cudaSetDevice(d->device));
	return check_zones(d->h.zones, d->stream, violations);
}
extern "C" int ns_dev_phase_ns(ns_dev *d, uint64_t out[16]){
	if (!d || !out) return NS_E_ARG;
	SimState st; int rc = read_state(d, &st); if (rc) return rc;
	for (int i = 0; i < 16; i++) out[i] = st.phase_ns[i];
	return 0;
}
extern "C" uint64_t ns_dev_launches(ns_dev *d){ return d ? d->launches : 0; }
extern "C" void ns_dev_destroy(ns_dev *d){
	if (!d) return;
	cudaSetDevice(d->device);
	free_sim(d->h);
	if (d->view_buf) cudaFreeHost(d->view_buf);
	if (d->d_sim) cudaFree(d->d_sim);
	if (d->d_scr) cudaFree(d->d_scr);
	if (d->d_hash) cudaFree(d->d_hash);
	if (d->e0) cudaEventDestroy(d->e0);
	if (d->e1) cudaEventDestroy(d->e1);
	if (d->e_run) cudaEventDestroy(d->e_run);
	if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
	if (d->stream) cudaStreamDestroy(d->stream);
	delete d;
}

// ---------------------------------------------------------------- C ABI: replica ensembles
// Per-replica record digest without ever copying the records out: the stream hash is a sum over the stream's
// 64-bit words (k_stream_hash), so a bounded record buffer can be folded into a running sum whenever it fills
// up (the engine's soft stop, host_stop == 2) — a word-aligned prefix is hashed at its stream position, the
// odd tail bytes move to the front of the buffer — and the simulation goes on.  One CTA per replica, one
// launch for the whole ensemble, one 4-byte read-back per round (how many replicas still want to run).
__global__ void __launch_bounds__(256) k_ens_drain(const Sim *sims, EnsAcc *acc, int n, int *n_more){
	const int s = blockIdx.x;
	if (s >= n) return;
	const Sim &S = sims[s];
	SimState *st = S.st;
	__shared__ u64 sh[8];
	__shared__ uint8_t tailb[8];
	const u64 len = st->rec_len, base = st->rec_base;
	const bool soft = st->status == 0 && st->host_stop == 2;
	if (!soft) return;                                                // finished, failed or bounded by max_events: nothing to fold
	const u64 nw = len >> 3;                                          // whole words (rec_base stays a multiple of 8)
	u64 a = 0;
	const u64 *w = (const u64 *)S.rec;
	for (u64 j = threadIdx.x; j < nw; j += blockDim.x) a += ns_mix64(w[j]+0x9e3779b97f4a7c15ULL*((base >> 3)+j+1));
	for (int d = 16; d; d >>= 1) a += __shfl_xor_sync(FULLMASK, a, d);
	if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
	const u64 tail = len & 7;
	if (threadIdx.x < tail) tailb[threadIdx.x] = S.rec[(nw << 3)+threadIdx.x];
	__syncthreads();
	if (threadIdx.x < tail) S.rec[threadIdx.x] = tailb[threadIdx.x];
	if (threadIdx.x == 0) {
		for (int k = 1; k < 8; k++) a += sh[k];
		acc[s].hash += a;
		st->rec_base = base+(nw << 3); st->rec_len = tail;
		st->host_stop = 0;
		atomicAdd(n_more, 1);
	}
}
// hand-overs between the two replica engines (CTA per replica <-> warp per replica) asked for during the round
__global__ void k_ens_handover(const Sim *sims, int n, int *n_more){
	for (int s = blockIdx.x*blockDim.x+threadIdx.x; s < n; s += gridDim.x*blockDim.x) {
		SimState *st = sims[s].st;
		if (st->status) continue;
		if (st->want_sparse || st->want_dense) { st->want_sparse = 0; st->want_dense = 0; atomicAdd(n_more, 1); }
	}
}
__global__ void __launch_bounds__(256) k_ens_digest(const Sim *sims, const EnsAcc *acc, int n, ns_digest *out){
	const int s = blockIdx.x;
	if (s >= n) return;
	const Sim &S = sims[s];
	const SimState *st = S.st;
	__shared__ u64 sh[8];
	const u64 len = st->rec_len, base = st->rec_base;
	const u64 nw = (len+7) >> 3;
	u64 a = 0;
	for (u64 j = threadIdx.x; j < nw; j += blockDim.x) {
		u64 x = 0;
		const u64 b0 = j*8;
		if (b0+8 <= len) x = ((const u64 *)S.rec)[j];
		else for (int k = 0; k < 8; k++) if (b0+k < len) x |= (u64)S.rec[b0+k] << (8*k);
		a += ns_mix64(x+0x9e3779b97f4a7c15ULL*((base >> 3)+j+1));
	}
	for (int d = 16; d; d >>= 1) a += __shfl_xor_sync(FULLMASK, a, d);
	if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int k = 1; k < 8; k++) a += sh[k];
		ns_digest d;
		d.seed = st->seed; d.events = st->cnt[NS_CNT_EVENTS]; d.solves = st->cnt[NS_CNT_SOLVES];
		d.record_bytes = base+len; d.record_fnv1a = ns_mix64((acc[s].hash+a) ^ (base+len)); d.t_end = st->now;
		d.status = st->status ? st->status : (st->host_stop == 2 ? NS_E_RECORDS_FULL : 0);
		out[s] = d;
	}
}

struct ns_ens;
static bool S_rec_ok(ns_ens *e);
struct ns_ens {
	int device = 0, n = 0;
	std::vector<HostSim> sims;
	Sim *d_sims = nullptr; SeqScratch *d_scr = nullptr; ParScratch *d_pars = nullptr;
	int cta_engine = 1;             // one CTA per replica (k_run_ens); 0 (ns_dev_opts.mode == NS_MODE_SEQ): one warp per replica (k_run_seq)
	cudaStream_t stream = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
	uint64_t launches = 0;
	char *arena = nullptr;          // every table of every replica
	// analyzer reductions (ns_ana.cu): per-replica stream state, one arena
	bool ana_on = false; AnaStream *d_ana = nullptr; ns_ana_summary *d_sum = nullptr; char *ana_arena = nullptr;
	EnsAcc *d_acc = nullptr; int *d_more = nullptr; ns_digest *d_dig = nullptr;
};

static bool S_rec_ok(ns_ens *e){ for (auto &h : e->sims) if (!h.S.want_rec) return false; return true; }   // the analysis reads records
static int ens_init(ns_ens *e, const void *const *scenarios, const size_t *lens, int n, const ns_dev_opts &o);
extern "C" void ns_ens_destroy(ns_ens *e);
extern "C" int ns_ens_create(ns_ens **out, const void *const *scenarios, const size_t *lens, int n, const ns_dev_opts *opts){
	if (!out || !scenarios || !lens || n <= 0) return NS_E_ARG;
	*out = nullptr;
	ns_dev_opts o;
	memset(&o, 0, sizeof o);
	if (opts) o = *opts;
	if (ns_dev_count() <= o.device) return NS_E_NO_DEVICE;
	CK(cudaSetDevice(o.device));
	ns_ens *e = new ns_ens();
	e->device = o.device; e->n = n;
	int rc = ens_init(e, scenarios, lens, n, o);
	if (rc) { ns_ens_destroy(e); return rc; }          // releases whatever ens_init had built
	*out = e;
	return 0;
}
static int ens_init(ns_ens *e, const void *const *scenarios, const size_t *lens, int n, const ns_dev_opts &o){
	e->sims.resize(n);
	e->cta_engine = o.mode != NS_MODE_SEQ;
	const bool cta = e->cta_engine != 0;
	std::vector<Sim> hs(n); std::vector<SeqScratch> hscr(n); std::vector<ParScratch> hpar(n);
	size_t total = 0;
	std::vector<size_t> need(n);
	for (int i = 0; i < n; i++) {                                     // measuring pass
		e->sims[i].arena_mode = 1; e->sims[i].arena_used = 0;
		int rc = build_sim(e->sims[i], scenarios[i], lens[i], o, cta, cta);
		if (rc) return rc;
		need[i] = e->sims[i].arena_used; total += need[i];
	}
	CK(cudaMalloc(&e->arena, total ? total : 8));
	CK(cudaStreamCreate(&e->stream));
	size_t off = 0;
	for (int i = 0; i < n; i++) {
		HostSim &h = e->sims[i];
		h.arena_mode = 2; h.arena = e->arena+off; h.arena_size = need[i]; h.arena_used = 0;
		off += need[i];
		int rc = build_sim(h, scenarios[i], lens[i], o, cta, cta);
		if (rc) return rc;
		hs[i] = h.S; hscr[i] = h.scr; hpar[i] = h.par;
	}
	CK(cudaEventCreate(&e->e0)); CK(cudaEventCreate(&e->e1));
	CK(cudaMalloc(&e->d_sims, sizeof(Sim)*n)); CK(cudaMalloc(&e->d_scr, sizeof(SeqScratch)*n));
	CK(cudaMalloc(&e->d_acc, sizeof(EnsAcc)*n)); CK(cudaMalloc(&e->d_more, 4)); CK(cudaMalloc(&e->d_dig, sizeof(ns_digest)*std::max(n, 4)));
	CK(cudaMemset(e->d_acc, 0, sizeof(EnsAcc)*n));
	CK(cudaMemcpy(e->d_sims, hs.data(), sizeof(Sim)*n, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(e->d_scr, hscr.data(), sizeof(SeqScratch)*n, cudaMemcpyHostToDevice));
	if (cta) {
		CK(cudaMalloc(&e->d_pars, sizeof(ParScratch)*n));
		CK(cudaMemcpy(e->d_pars, hpar.data(), sizeof(ParScratch)*n, cudaMemcpyHostToDevice));
		int rc = ens_preload(); if (rc) return rc;
	}
	CK(cudaDeviceSynchronize());
	{ int rc = seq_launch_init_groups(e->d_sims, n, e->stream); if (rc) return rc; }
	e->launches++;
	CK(cudaStreamSynchronize(e->stream));
	return 0;
}
// every replica runs until its queue is empty, its script ends it, or it has handled max_events_each events; a
// replica whose record buffer fills up is folded into its running digest and resumed (no host copy of records)
extern "C" int ns_ens_run(ns_ens *e, uint64_t max_events_each, float *device_ms){
	if (!e) return NS_E_ARG;
	CK(cudaSetDevice(e->device));
	for (int i = 0; i < e->n; i++)
		CK(cudaMemcpyAsync(&e->sims[i].S.st->max_events, &max_events_each, 8, cudaMemcpyHostToDevice, e->stream));
	CK(cudaEventRecord(e->e0, e->stream));
	for (int round = 0; round < 1 << 20; round++) {
		CK(cudaMemsetAsync(e->d_more, 0, 4, e->stream));
		if (e->cta_engine) {
			// dense phases: one CTA per replica; sparse phases (streaks of tiny batches): one warp per replica, so that
			// every replica of the ensemble is resident at once.  A replica switches engine at most once per round.
			AnaStream *ana = e->ana_on ? e->d_ana : nullptr;
			int rc = ens_launch_run(e->d_sims, e->d_scr, e->d_pars, e->n, e->stream, e->d_acc, ana); if (rc) return rc;
			rc = seq_launch_run(e->d_sims, e->d_scr, e->n, e->stream, 1, e->d_acc, ana); if (rc) return rc;
			k_ens_handover<<<(e->n+127)/128, 128, 0, e->stream>>>(e->d_sims, e->n, e->d_more);
			e->launches += 2;
		} else {
			int rc = seq_launch_run(e->d_sims, e->d_scr, e->n, e->stream, 0, e->d_acc, e->ana_on ? e->d_ana : nullptr); if (rc) return rc;
		}
		if (e->ana_on) { int rc = ana_launch(e->d_ana, e->n, 0, e->stream); if (rc) return rc; e->launches++; }   // before anything is folded away
		k_ens_drain<<<e->n, 256, 0, e->stream>>>(e->d_sims, e->d_acc, e->n, e->d_more);
		CK(cudaGetLastError());
		e->launches += 2;
		int more = 0;
		CK(cudaMemcpyAsync(&more, e->d_more, 4, cudaMemcpyDeviceToHost, e->stream));
		CK(cudaStreamSynchronize(e->stream));
		if (!more) break;
	}
	CK(cudaEventRecord(e->e1, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	if (device_ms) CK(cudaEventElapsedTime(device_ms, e->e0, e->e1));
	return 0;
}
// one kernel for all replicas (hash of what is still buffered + the running sums), one copy back
extern "C" int ns_ens_digests(ns_ens *e, ns_digest *out, int n){
	if (!e || !out) return NS_E_ARG;
	if (n > e->n) n = e->n;
	if (n <= 0) return 0;
	CK(cudaSetDevice(e->device));
	k_ens_digest<<<n, 256, 0, e->stream>>>(e->d_sims, e->d_acc, n, e->d_dig);
	CK(cudaGetLastError());
	e->launches++;
	CK(cudaMemcpyAsync(out, e->d_dig, sizeof(ns_digest)*n, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return 0;
}
// ---- analyzer reductions over every replica's record stream (ns_ana.cu)
extern "C" int ns_ens_analysis_begin(ns_ens *e, double period){
	if (!e || !(period > 0) || e->ana_on || !S_rec_ok(e)) return NS_E_ARG;   // (the analysis reads the record stream)
	CK(cudaSetDevice(e->device));
	const int n = e->n;
	auto al = [](size_t b){ return (b+255) & ~(size_t)255; };
	std::vector<size_t> off(n+1, 0);
	std::vector<uint32_t> hsz(n);
	for (int i = 0; i < n; i++) {
		const size_t N = (size_t)e->sims[i].S.n_nodes;
		uint32_t h = 16; while (h < 2*N+2) h <<= 1;
		hsz[i] = h;
		off[i+1] = off[i]+al(h*4)+al(h*4)+al(N)+4*al(N*8)+al(N)+al(N*8);
	}
	CK(cudaMalloc(&e->ana_arena, off[n] ? off[n] : 8));
	CK(cudaMalloc(&e->d_ana, sizeof(AnaStream)*n)); CK(cudaMalloc(&e->d_sum, sizeof(ns_ana_summary)*n));
	CK(cudaMemset(e->ana_arena, 0, off[n] ? off[n] : 8));
	CK(cudaMemset(e->d_sum, 0, sizeof(ns_ana_summary)*n));
	std::vector<AnaStream> hs(n);
	for (int i = 0; i < n; i++) {
		const Sim &S = e->sims[i].S;
		const size_t N = (size_t)S.n_nodes;
		char *p = e->ana_arena+off[i];
		AnaStream &A = hs[i];
		A.rec = S.rec; A.rec_base = &S.st->rec_base; A.rec_len = &S.st->rec_len;
		A.st = e->d_sum+i; A.n_nodes = (int)N; A.hmask = hsz[i]-1; A.period = period;
		uint32_t *hk = (uint32_t *)p; p += al(hsz[i]*4);
		int *hv = (int *)p; p += al(hsz[i]*4);
		A.hkey = hk; A.hval = hv;
		A.ntype = (uint8_t *)p; p += al(N);
		for (int d = 0; d < 2; d++) { A.u_time[d] = (double *)p; p += al(N*8); A.u_val[d] = (double *)p; p += al(N*8); }
		A.n_state = (int8_t *)p; p += al(N);
		A.n_since = (double *)p; p += al(N*8);
		// node id -> index (the ids were drawn on the host side of build_sim and live in S.n_id)
		std::vector<uint32_t> ids(N), keys(hsz[i], 0); std::vector<int> vals(hsz[i], -1);
		if (N) CK(cudaMemcpy(ids.data(), S.n_id, N*4, cudaMemcpyDeviceToHost));
		for (size_t k = 0; k < N; k++) {
			uint32_t h = (ids[k]*2654435761u) & A.hmask;
			while (vals[h] >= 0) h = (h+1) & A.hmask;
			keys[h] = ids[k]; vals[h] = (int)k;
		}
		CK(cudaMemcpy(hk, keys.data(), (size_t)hsz[i]*4, cudaMemcpyHostToDevice));
		CK(cudaMemcpy(hv, vals.data(), (size_t)hsz[i]*4, cudaMemcpyHostToDevice));
		if (N) { CK(cudaMemset(A.ntype, 2, N)); CK(cudaMemset(A.n_state, 0xff, N)); }     // everything is a client until a state record says otherwise; no state yet
	}
	CK(cudaMemcpy(e->d_ana, hs.data(), sizeof(AnaStream)*n, cudaMemcpyHostToDevice));
	e->ana_on = true;
	return 0;
}
extern "C" int ns_ens_analysis(ns_ens *e, ns_ana_summary *out, int n){
	if (!e || !out || !e->ana_on) return NS_E_ARG;
	if (n > e->n) n = e->n;
	CK(cudaSetDevice(e->device));
	{ int rc = ana_launch(e->d_ana, e->n, 1, e->stream); if (rc) return rc; }
	e->launches++;
	CK(cudaMemcpyAsync(out, e->d_sum, sizeof(ns_ana_summary)*n, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return 0;
}
// the engine counters of all replicas, summed on the device (one kernel, one copy)
__global__ void k_ens_counters(const Sim *sims, int n, u64 *out){
	for (int s = blockIdx.x*blockDim.x+threadIdx.x; s < n; s += gridDim.x*blockDim.x) {
		const SimState *st = sims[s].st;
		for (int q = 0; q < NS_CNT_COUNT; q++) {
			u64 v = q == NS_CNT_SEQ ? st->tail : q == NS_CNT_DRAWS ? st->cnt[q]+st->burns : st->cnt[q];
			if (q == NS_CNT_MAX_BATCH) atomicMax((unsigned long long *)&out[q], v); else atomicAdd((unsigned long long *)&out[q], v);
		}
	}
}
extern "C" int ns_ens_counters_sum(ns_ens *e, uint64_t out[NS_CNT_COUNT]){
	if (!e || !out) return NS_E_ARG;
	CK(cudaSetDevice(e->device));
	u64 *d_out = (u64 *)e->d_dig;                                     // the digest buffer is free between digests
	CK(cudaMemsetAsync(d_out, 0, sizeof(u64)*NS_CNT_COUNT, e->stream));
	k_ens_counters<<<(e->n+127)/128, 128, 0, e->stream>>>(e->d_sims, e->n, d_out);
	CK(cudaGetLastError());
	e->launches++;
	CK(cudaMemcpyAsync(out, d_out, sizeof(u64)*NS_CNT_COUNT, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return 0;
}
extern "C" int ns_ens_check(ns_ens *e, int *violations){
	if (!e || !violations) return NS_E_ARG;
	CK(cudaSetDevice(e->device));
	std::vector<std::pair<char *, size_t>> all;
	for (auto &h : e->sims) all.insert(all.end(), h.zones.begin(), h.zones.end());
	return check_zones(all, e->stream, violations);
}
extern "C" uint64_t ns_ens_launches(ns_ens *e){ return e ? e->launches : 0; }
extern "C" void ns_ens_destroy(ns_ens *e){
	if (!e) return;
	cudaSetDevice(e->device);
	for (auto &h : e->sims) free_sim(h);
	if (e->arena) cudaFree(e->arena);
	if (e->d_sims) cudaFree(e->d_sims);
	if (e->d_scr) cudaFree(e->d_scr);
	if (e->d_pars) cudaFree(e->d_pars);
	if (e->d_ana) cudaFree(e->d_ana);
	if (e->d_sum) cudaFree(e->d_sum);
	if (e->ana_arena) cudaFree(e->ana_arena);
	if (e->d_acc) cudaFree(e->d_acc);
	if (e->d_more) cudaFree(e->d_more);
	if (e->d_dig) cudaFree(e->d_dig);
	if (e->e0) cudaEventDestroy(e->e0);
	if (e->e1) cudaEventDestroy(e->e1);
	if (e->stream) cudaStreamDestroy(e->stream);
	delete e;
}

// ---------------------------------------------------------------- host helpers
extern "C" void ns_rand_seed_host(ns_rand_state *g, uint32_t seed){
	int32_t word = seed ? (int32_t)seed : 1;
	g->st[0] = word;
	for (int i = 1; i < 31; i++) {
		long hi = word/127773, lo = word%127773;
		long w = 16807*lo-2836*hi;
		if (w < 0) w += 2147483647;
		word = (int32_t)w;
		g->st[i] = word;
	}
	g->f = 3; g->r = 0;
	for (int i = 0; i < 310; i++) ns_rand_step(g->st, g->f, g->r);
}
extern "C" uint32_t ns_rand_next_host(ns_rand_state *g){ return ns_rand_step(g->st, g->f, g->r); }
extern "C" void ns_rand_jump_host(ns_rand_state *g, uint64_t n){ ns_rand_jump(g->st, g->f, g->r, n); }

extern "C" const char *ns_strerror(int err){
	switch (err) {
	case NS_OK: return "ok";
	case NS_E_BAD_SCENARIO: return "malformed scenario script";
	case NS_E_BACK_IN_TIME: return "event added back in time (reference aborts, event.h:60-64)";
	case NS_E_NAN_TIME: return "event added at NaN (reference aborts, event.h:67-70)";
	case NS_E_NULL_RANGE: return "flow source range missing (reference dereferences NULL)";
	case NS_E_NO_PRODUCER: return "throttle on a range without producer (flow.c:411)";
	case NS_E_RING_FULL: return "pending-event ring full";
	case NS_E_TABLE_FULL: return "device table capacity exhausted";
	case NS_E_RECORDS_FULL: return "record buffer full";
	case NS_E_NO_DEVICE: return "no usable CUDA device";
	case NS_E_ARG: return "bad argument";
	case NS_E_BATCH_FULL: return "equal-timestamp batch exceeds scratch capacity";
	case NS_E_INTERNAL: return "engine invariant failed";
	case NS_E_RECORDS_LOST: return "record buffer overflowed inside a batch: records were dropped";
	default: return err > 0 ? "CUDA error" : "unknown error";
	}
}
