// ns_sim.cuh — device tables (structure of arrays in HBM) and small helpers shared by the kernels.
//
// The reference keeps one calloc'd struct per node/flow/range/event, linked by list_head, uthash and
// skip lists (data.h:15-151).  Here every field is a column; links are 32-bit indices:
//   nodes    data.h:100-112   max/usage/total per direction, id, resource list head
//   adj      node->conns[dir] (list.h lists, newest first, flow.c:311-312) as growable index arrays:
//            iteration newest-first == descending position, removed entries are tombstones (-1)
//   flows    data.h:44-87     ub, speed[2], delay, peers, ranges, generation (cancels pending events)
//   ranges   data.h:15-24     start/total_len/len/lenc/last_update/producer, sorted list per resource
//   ring     every event_add (event.h:56-74) takes the next sequence number `tail`; SPEED_CHANGE
//            events (flow.h:5-11) live in ring slot seq & (cap-1), so ring order == FIFO order of
//            equal timestamps (event.h:27-31) and an equal-time batch is an ordered compaction.
//   L table  long-lived events: FLOW_DONE (slot 2f) / FLOW_DRAIN|THROTTLE (slot 2f+1) of each flow
//            (flow->done, flow->drain, data.h:84) and USER events, with explicit seq.
// Both event stores keep a per-tile (NS_TILE entries) lower bound of their minimum time (q_tmin/l_tmin): the
// next-completion argmin reads only those, the batch extraction only the tiles that hit.
#pragma once
#include <stdint.h>
#include <math.h>
#include "../../include/ns_scenario.h"
#include "../../include/netsim_b200.h"

#define NS_TILE 256              // entries per tile of the event stores (one minimum per tile): 8 rows of 32 lanes = one 8-deep batch of loads
#define NS_TILE_SHIFT 8
#define NS_TILE_ROWS (NS_TILE/32)
#define NS_EPS 1e-4                       /* sim.c:11 */
#define NS_INF (__longlong_as_double(0x7ff0000000000000LL))
#define FULLMASK 0xffffffffu
#define NS_FEAT_CHAIN2 1u   /* small-node event chains: scalar recurrence pass + lane-parallel emission pass */
#define NS_FEAT_HUBCTA 2u   /* full spreads of large SND groups by a whole CTA from shared memory */
#define NS_FEAT_WIDE 4u     // single-warp engine: runs of SPEED_CHANGE events handled one event per lane (ns_seq.cuh)
#define NS_FEAT_INIT_REC 8u // the R_NODE_CREATE records of the script's nodes are written by ns_dev_create (off: the caller writes them itself, in its own order)
#define NS_FEAT_ALL 0xffffu

enum { SND = 0, RCV = 1 };
enum { EV_FLOW_DRAIN = 0, EV_FLOW_DONE = 1, EV_FLOW_SPEED_THROTTLE = 2, EV_PACKET_DONE = 3, EV_SPEED_CHANGE = 4, EV_USER = 5 };
enum { NST_OFFLINE = 0, NST_STALE, NST_PLAYING, NST_DONE, NST_IDLE, NST_SERVER, NST_CLOUD, NST_DYING };   // enum node_state, data.h:89-98
enum { UE_DONE_PLAY = 0, UE_PAUSE_BUFFERING = 1, UE_DONE_BUFFERING = 2 };                                 // enum ue_type, p2p_common.h:41-49
enum { R_USAGE = 0, R_SPD = 2, R_CONN_CREATE = 4, R_CONN_DST = 5, R_CONN_CLOSE = 6, R_NODE_STATE = 7, R_NODE_CREATE = 8 };

typedef unsigned long long u64;

struct SimState {                 // mutable scalars of one simulation
	double now;
	u64 head, tail;               // ring window [head, tail) in absolute sequence numbers
	u64 burns;                    // random() draws burned since the last real draw (skiplist.h:34-40)
	int32_t rnd[31]; int32_t rf, rr;
	int status, exit_flag, relay_seen, pad0;
	u64 rec_len, tr_len;
	u64 rec_base;                 // record bytes already drained by the host (ns_dev_records_drain)
	u64 cnt[NS_CNT_COUNT];
	int free_flow, hi_flow, free_range, hi_range, hi_res, free_user, hi_user, n_ord;
	int adj_top, pad1;
	u64 max_events;
	u64 seed;
	u64 phase_ns[16];             // leader-thread wall time per engine phase (profiling aid)
	// host-callback mode (modules with their own handlers)
	unsigned host_mask;           // bit t set: stop after every handled event of type t
	int host_stop;                // 1: stopped for a host handler; 2: record buffer nearly full (drain, run again)
	int pending_close;            // FLOW_DONE: flow whose flow_close (HNDR_CLEANER) runs on resume
	int sparse_mode;              // k_run_seq launched by the grid engine's host loop for a sparse phase
	int want_sparse, want_dense;  // engine hand-over requests between k_run_par and k_run_seq (ns_par.cuh)
	int le_type, le_flow, le_ord, le_dir, le_token;   // the event the host must see
	double le_time, le_speed;
};

struct Sim {
	// capacities
	int n_nodes, cap_flows, cap_ranges, cap_res, cap_users, cap_ord;
	int n_groups, n_ops, n_ondone, n_rids;
	u64 ring_cap; int n_ring_tiles; int n_l, n_l_tiles;
	int adj_arena_cap, batch_cap;
	u64 rec_cap, tr_cap;
	int want_rec, want_tr;
	int features;                 // NS_FEAT_* bits (ns_dev_opts.reserved[0] / NS_FEATURES): engine paths that can be switched off for A/B runs
	SimState *st;
	// nodes
	double *n_max[2], *n_usage[2], *n_total[2]; uint32_t *n_id; int *n_res_head;
	// adjacency: group g = node*2+dir
	int *adj_off, *adj_cap, *adj_len, *adj_live, *adj_arena;
	// flows
	double *f_ub, *f_speed[2], *f_delay; int *f_peer[2]; uint32_t *f_id, *f_gen;
	int *f_srng, *f_drng, *f_cnext, *f_cprev, *f_pos[2], *f_ord, *f_next_free; uint8_t *f_alive, *f_drain_type;
	// ranges
	long long *r_start, *r_total; double *r_len, *r_lenc, *r_last;
	int *r_prod, *r_res, *r_next, *r_prev, *r_cons, *r_next_free;
	// resources
	uint32_t *s_id; long long *s_len; int *s_node, *s_first, *s_nprod, *s_ncons, *s_next;
	// speed-event ring
	double *q_time, *q_speed; int *q_fd; uint32_t *q_gen; double *q_tmin;
	// long-lived event table
	double *l_time; u64 *l_seq; double *l_tmin; int *u_group, *u_next_free;
	// behaviour library, client side (scripts with NS_SC_F_CLIENTS): struct def_user (modules/p2p_common.h:8-29) and
	// node->state as columns; the client's own USER event d->e lives in L-table slot 2*cap_flows+cap_users+node
	int has_clients, n_srv;
	const ns_sc_client *c_cfg; const int *srv_list;   // script table; server nodes, newest first (list_add at head, p2p_common.h:112)
	int *n_state;                                     // enum node_state, data.h:89-98
	double *c_bit_rate, *c_last; int *c_buffer_pos, *c_next_state, *c_trigger, *c_ue_type, *c_ue_rng;
	int *s_brate;                                     // resource->bit_rate (data.h:29)
	// script
	const ns_sc_op *ops; const int *group_off; const ns_sc_ondone *ondone; const uint32_t *rids;
	int *ord2flow;
	ns_sc_op *host_ops; int host_ops_cap;   // operations injected by host-side handlers (ns_dev_exec)
	// per-timestamp batch scratch
	u64 *b_seq; int *b_ref;       // ring entries: ref = -1; L entries: ref = slot
	// outputs
	uint8_t *rec; ns_trace_row32 *tr;
	// host core only (k_exec_ops / k_run_seq in host-callback mode): resource->auto_delete (data.h:34), null otherwise
	uint8_t *s_autodel;
	// batch engine: groups with more events than this skip the one-thread-per-group pass and get a warp (NS_E1_MAX, default 32)
	int e1_max;
	int hub_min_deg;              // sender groups with at least this many flows run their chain on a whole CTA (NS_HUB_MIN_DEG, default HUB_MIN_DEG)
};

// Replica ensembles: a replica whose bounded record buffer fills up folds it on the spot — the analyzer reductions take
// what they need (ns_ana_dev.cuh), a word-aligned prefix goes into the running stream hash at its stream position, the
// odd tail bytes move to the front — and goes on (ns_ens_*; no records ever leave the device).
struct EnsAcc { u64 hash; };
struct AnaStream;
struct EnsFold { EnsAcc *acc; AnaStream *ana; };   // ana may be null (no analysis requested)

// Staging area of the grid-parallel path: an event handled out of order cannot know its final
// sequence numbers / record offset, so its event_adds, records and trace row are parked here under
// LOCAL indices and committed in batch order by a prefix scan (ns_par.cuh).
#define NS_STAGE_INLINE 4
#define NS_STAGE_REC 56
#define NS_SEQ_PROV (1ull << 63)
#define NS_HOLE_PURE (-2147483647-1)
struct ParStage {
	double *st_time, *st_speed; int *st_fd;     // [NS_STAGE_INLINE*seg_cap] inline, then overflow
	u64 inline_n, ovf_cap; u64 *ovf_top;
	uint8_t *st_rec; ns_trace_row32 *st_tr;
	unsigned *ev_nins; long long *ev_ovf; uint8_t *ev_nrec, *ev_ntr;
	unsigned *ev_idpos;           // parallel creates: event_adds of the op before its flow-id draw (flow.c:327)
};
enum { LC_EVENTS = 0, LC_SPEED_CHANGE, LC_SOLVES, LC_FULL_SPREADS, LC_FLOW_VISITS, LC_RETIMES, LC_N };

// scratch layout inside b_seq/b_ref: [0, batch_cap) ring part, L part in lb arrays
struct SeqScratch { u64 *lb_seq; int *lb_ref; u64 *lb_tmp_seq; int *lb_tmp_ref; };

// ---------------------------------------------------------------- glibc random(), TYPE_3
// r[k] = r[k-31] + r[k-3] (mod 2^32), output r[k] >> 1.  Shared by host init (sim_create_node ids)
// and the device (flow ids, flow.c:327).
struct NsRand { int32_t *st; int32_t *f, *r; };

__host__ __device__ inline uint32_t ns_rand_step(int32_t *st, int32_t &f, int32_t &r){
	uint32_t v = (uint32_t)st[f]+(uint32_t)st[r];
	st[f] = (int32_t)v;
	if (++f >= 31) f = 0;
	if (++r >= 31) r = 0;
	return v >> 1;
}

// polynomial arithmetic modulo x^31 - x^28 - 1 over Z/2^32 (the recurrence's characteristic
// polynomial), used to skip n draws in O(31^2 log n).
__host__ __device__ inline void ns_poly_mulmod(const uint32_t *a, const uint32_t *b, uint32_t *out){
	uint32_t t[61];
	for (int i = 0; i < 61; i++) t[i] = 0;
	for (int i = 0; i < 31; i++) {
		uint32_t ai = a[i];
		if (!ai) continue;
		for (int j = 0; j < 31; j++) t[i+j] += ai*b[j];
	}
	for (int k = 60; k >= 31; k--) {      // x^k = x^(k-3) + x^(k-31)
		uint32_t c = t[k];
		t[k-3] += c; t[k-31] += c;
	}
	for (int i = 0; i < 31; i++) out[i] = t[i];
}

__host__ __device__ inline void ns_rand_jump(int32_t *st, int32_t &f, int32_t &r, u64 n){
	// stepping costs ~10 instructions per draw, the polynomial jump ~13 000 whatever n is (measured: 12 % of all instructions of a
	// p2p replica were jumps over a few hundred burned draws, profiles/README.md r2-p): jump only over long stretches
	if (n < 1024) { for (u64 i = 0; i < n; i++) ns_rand_step(st, f, r); return; }
	// chronological window s_0..s_30 (oldest first): the next draw overwrites the oldest word
	// st[f] with st[f]+st[r], r = f+28 (mod 31), i.e. s_31 = s_0 + s_28.
	uint32_t s[61];
	for (int i = 0; i < 31; i++) s[i] = (uint32_t)st[(f+i)%31];
	for (int i = 31; i < 61; i++) s[i] = s[i-31]+s[i-3];
	uint32_t res[31], base[31], tmp[31];
	for (int i = 0; i < 31; i++) { res[i] = 0; base[i] = 0; }
	res[0] = 1; base[1] = 1;
	u64 e = n;
	while (e) {
		if (e & 1) { ns_poly_mulmod(res, base, tmp); for (int i = 0; i < 31; i++) res[i] = tmp[i]; }
		e >>= 1;
		if (e) { ns_poly_mulmod(base, base, tmp); for (int i = 0; i < 31; i++) base[i] = tmp[i]; }
	}
	uint32_t ns[31];
	for (int i = 0; i < 31; i++) {
		uint32_t acc = 0;
		for (int j = 0; j < 31; j++) acc += res[j]*s[i+j];
		ns[i] = acc;
	}
	for (int i = 0; i < 31; i++) st[i] = (int32_t)ns[i];
	f = 0; r = 28;                        // oldest word now at index 0
}

__host__ __device__ inline u64 ns_mix64(u64 z){
	z = (z ^ (z >> 30))*0xbf58476d1ce4e5b9ULL;
	z = (z ^ (z >> 27))*0x94d049bb133111ebULL;
	return z ^ (z >> 31);
}
