/* netsim_b200.h — C ABI of the B200 device layer for netsim's flow/event hot path.
 *
 * Plain C, no torch/CUDA types in any signature. Every entry point returns 0 on success or a
 * negative ns_err / a positive CUDA error code; nothing throws across the boundary. The host core
 * (plain C, netsim's sim.c/main.c side) owns every handle.
 *
 * What each entry point replaces in the reference (paths relative to yshui/netsim):
 *
 *   ns_dev_create        sim_state_new (data.h:215-224) + the module's *_init building nodes,
 *                        resources and first flows (sim.c:113-154, sim.c:42-94); the device tables
 *                        are the SoA form of struct node/flow/range/event (data.h:15-151).
 *   ns_dev_run           the event loop main.c:50-67 with its default handlers main.c:44-48:
 *                        event_pop (event.h:20-25) -> device argmin + equal-timestamp batch;
 *                        handle_speed_change (flow.c:341-384), bwspread (flow.c:35-205),
 *                        queue_speed_event (flow.c:16-29), _range_update (range.h:91-118),
 *                        range_calc_and_requeue_events (range.c:16-79), flow_done_handler /
 *                        flow_close (flow.c:391-406, 241-292), flow_throttle_handler (flow.c:408-423),
 *                        range_merge_with_next (range.c:83-124), event_add/remove (event.h:13-74).
 *   ns_dev_records       write_record's byte stream (record.c:27-117), to be appended by the host
 *                        to the mmap'd record file (record.c:119-153).
 *   ns_dev_pick_batch    server_picker_opt1 + qselect_eval (modules/client_behaviour.h:63-88,
 *                        modules/qselect.h:5-27) and server_picker_opt2 (client_behaviour.h:90-124)
 *                        for a batch of simultaneous NEW_CONNECTION events.
 *   ns_ens_*             replica ensembles: independent simulations side by side on
 *                        one GPU, one digest per replica.
 *   ns_rand_*            glibc random()/srandom() (TYPE_3) as used by sim.c:117,132, flow.c:327 and
 *                        skiplist.h:34-40, with O(log n) jump-ahead.
 */
#ifndef NETSIM_B200_H
#define NETSIM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum ns_err {
	NS_OK = 0,
	NS_E_BAD_SCENARIO = -1,
	NS_E_BACK_IN_TIME = -2,  /* event.h:60-64: the reference would abort() here */
	NS_E_NAN_TIME = -3,      /* event.h:67-70 */
	NS_E_NULL_RANGE = -4,    /* the reference would dereference a NULL range (sim.c:69 -> range.c:21) */
	NS_E_NO_PRODUCER = -5,   /* flow.c:411 */
	NS_E_RING_FULL = -6,     /* pending-event ring too small: recreate with a larger ring_cap */
	NS_E_TABLE_FULL = -7,    /* flow/range/resource/adjacency capacity exhausted */
	NS_E_RECORDS_FULL = -8,  /* ns_dev_run stopped BEFORE a batch whose records might not fit: copy the records
	                            out, ns_dev_records_drain, run again (nothing is lost) */
	NS_E_NO_DEVICE = -9,
	NS_E_ARG = -10,
	NS_E_BATCH_FULL = -11,
	NS_E_INTERNAL = -12,     /* an engine invariant failed (a bug, not an input error) */
	NS_E_RECORDS_LOST = -13, /* the record buffer overflowed INSIDE a batch and records were dropped (sticky, unlike the soft
	                            NS_E_RECORDS_FULL): recreate with a larger record_cap */
};

enum ns_counter {
	NS_CNT_EVENTS = 0, NS_CNT_SOLVES, NS_CNT_FULL_SPREADS, NS_CNT_FLOW_VISITS, NS_CNT_SPEED_CHANGE,
	NS_CNT_FLOW_DONE, NS_CNT_DRAIN, NS_CNT_THROTTLE, NS_CNT_USER, NS_CNT_CREATES, NS_CNT_CLOSES,
	NS_CNT_DRAWS, NS_CNT_SEQ, NS_CNT_LIVE_FLOWS, NS_CNT_PEAK_FLOWS, NS_CNT_BATCHES,
	NS_CNT_PAR_EVENTS,   /* events handled by the grid-parallel path */
	NS_CNT_MAX_BATCH, NS_CNT_WAVES,
	NS_CNT_RETIMES,     /* range_calc_and_requeue_events invocations (range.c:16) */
	NS_CNT_COUNT
};

enum ns_mode {
	NS_MODE_AUTO = 0,   /* grid-parallel batches where legal, sequential otherwise */
	NS_MODE_SEQ = 1,    /* one warp, strictly sequential (always legal; the parity anchor) */
};

typedef struct ns_dev ns_dev;
typedef struct ns_ens ns_ens;

struct ns_dev_opts {
	int device;            /* CUDA device ordinal */
	int mode;              /* enum ns_mode */
	int want_records;      /* build the record byte stream */
	int want_trace;        /* fp64 event trace rows (tests) */
	uint64_t ring_cap;     /* pending speed-event ring entries, 0 = auto */
	uint64_t record_cap;   /* bytes, 0 = auto */
	uint64_t trace_cap;    /* rows, 0 = auto */
	int cap_flows;         /* concurrent flows, 0 = auto (number of CREATE ops) */
	int features_off;      /* bit mask of engine paths to switch OFF (A/B measurements), 0 = everything on */
	int cap_ord;           /* flows ever created (creation ordinals), 0 = auto (2x the script's CREATE ops, >= 2x cap_flows) */
	int cap_ranges;        /* ranges, 0 = auto */
	int cap_users;         /* pending USER events, 0 = auto (2x the script's SCHED ops) */
	int cap_nodes;         /* nodes, 0 = the script's (a host core that creates nodes later through NS_OP_NEWNODE reserves room here) */
};

struct ns_trace_row32 { uint32_t type, ord; double time, v0, v1; };

struct ns_digest {          /* one per simulation */
	uint64_t seed, events, solves, record_bytes, record_fnv1a;
	double t_end;
	int64_t status;
};

int  ns_dev_count(void);
int  ns_dev_create(ns_dev **out, const void *scenario, size_t len, const struct ns_dev_opts *opts);
/* handle events until the queue is empty, the script ended the run, or max_events (0 = no limit)
 * events have been handled in total; device_ms (may be NULL) receives the CUDA-event time. */
int  ns_dev_run(ns_dev *d, uint64_t max_events, float *device_ms);
int  ns_dev_status(ns_dev *d);
double ns_dev_now(ns_dev *d);
int  ns_dev_counters(ns_dev *d, uint64_t out[NS_CNT_COUNT]);
/* copy record bytes [offset, offset+cap) of the STREAM (offsets keep counting across drains); *len = bytes
 * copied, *total = bytes produced so far. */
int  ns_dev_records(ns_dev *d, void *dst, size_t cap, size_t offset, size_t *len, size_t *total);
/* the same copy issued on a side stream (dst: pinned host memory): it overlaps the next ns_dev_run, which
 * only appends beyond the span; ns_dev_records_wait blocks until every issued copy has landed.
 * This is the pinned record ring of the host's record writer (record.c:27-117 stays on the host). */
int  ns_dev_records_async(ns_dev *d, void *dst, size_t cap, size_t offset, size_t *len, size_t *total);
int  ns_dev_records_wait(ns_dev *d);
int  ns_dev_records_drain(ns_dev *d);   /* all records produced so far were copied: reuse the buffer */
int  ns_dev_digest(ns_dev *d, struct ns_digest *out);
int  ns_dev_trace(ns_dev *d, struct ns_trace_row32 *dst, size_t cap_rows, size_t *rows);
/* per creation ordinal {alive, speed[0], speed[1], drng->len}; per node {usage[0], usage[1],
 * total_bwupbound[0], total_bwupbound[1]} */
int  ns_dev_flow_state(ns_dev *d, double *out4, size_t max_flows, size_t *n);
int  ns_dev_node_state(ns_dev *d, double *out4, size_t max_nodes, size_t *n);
/* ---- host-callback mode: modules that register their own handlers (sim_register_handler,
 * sim.c:96-111) get them called on the host between the device's DEFAULT and CLEANER work.
 * ns_dev_set_host_mask: bit t = stop after every handled event of enum event_type t (data.h:114-122).
 * ns_dev_run_host: run until such an event, the end of the queue or sim_end; for FLOW_DONE the flow is
 *   still alive when the host sees it (flow_done_cleaner runs on the next call), like HNDR_USER
 *   between HNDR_DEFAULT and HNDR_CLEANER (data.h:126-130).
 * ns_dev_exec: what a handler may do to the core — NS_OP_CREATE / NS_OP_CREATE_RID (sim_establish_flow; on return .a is the
 *   flow's creation ordinal or -1), NS_OP_SCHED (event_add of a USER event; .a is a host token that
 *   comes back in ns_host_event.token), NS_OP_CANCEL (event_remove), NS_OP_CLOSE (flow_close), NS_OP_END (sim_end), and what
 *   the host core needs for modules that create things while the simulation runs (p2p.c): NS_OP_NEWNODE (sim_create_node),
 *   NS_OP_NEWRES (sim_node_new_resource / sim_node_add_resource), NS_OP_DELRES / NS_OP_AUTODEL (node_del_resource,
 *   resource->auto_delete), NS_OP_RANGE_SET (a module-side _range_update written back); include/ns_scenario.h has the operands. */
struct ns_host_event { int32_t type, flow_ord, dir, token; double time, speed, now; int32_t stopped, status, exited, pad; };
struct ns_flow_info { int32_t alive, src, dst; uint32_t flow_id, resource_id; int32_t pad; double ub, delay, speed[2]; int64_t drng_start; double drng_len; };
struct ns_range_info { int64_t start, total_len; double len; int32_t producer_ord, pad; };
struct ns_sc_op;
struct ns_rand_state;
/* the whole mutable state a module handler may read, for the host core's mirrors (netsim_b200/host/core.c): node usage, every
 * flow slot and every range slot below the high-water marks, written by one kernel straight into pinned host memory the library
 * owns (valid until the next call).  Dead slots have slot == -1.  done_t / drain_t = +inf when the event is not pending
 * (event.h:76-88 is_active).  nodes4[4*i..] = {bandwidth_usage[0], [1], total_bwupbound[0], [1]} of node i. */
struct ns_view_hdr   { int32_t hi_flow, hi_range, n_nodes, pad; };
struct ns_view_flow  { int32_t ord, slot, src, dst, srng, drng, drain_type, pad; uint32_t flow_id, pad2; double speed[2], done_t, drain_t; };
struct ns_view_range { int32_t slot, res, producer_ord, next, prev, pad; int64_t start, total_len; double len, lenc, last_update; };
int  ns_dev_host_view(ns_dev *d, const struct ns_view_hdr **hdr, const double **nodes4, const struct ns_view_flow **flows, const struct ns_view_range **ranges);
/* the simulation's random() stream (sim.c:117,132, flow.c:327, skiplist.h:34-40 share ONE generator with the module's own
 * random() calls): get applies the pending burns and returns the state, set installs a state (after the host drew from it) */
int  ns_dev_rng_get(ns_dev *d, struct ns_rand_state *g);
int  ns_dev_rng_set(ns_dev *d, const struct ns_rand_state *g);
int  ns_dev_set_host_mask(ns_dev *d, uint32_t mask);
int  ns_dev_run_host(ns_dev *d, struct ns_host_event *ev);
int  ns_dev_exec(ns_dev *d, struct ns_sc_op *ops, int n);
int  ns_dev_flow_info(ns_dev *d, int ord, struct ns_flow_info *out);
int  ns_dev_node_ranges(ns_dev *d, int node, uint32_t rid, struct ns_range_info *out, int max, int *n);
int  ns_dev_node_id(ns_dev *d, int node, uint32_t *id);
int  ns_dev_resource_id(ns_dev *d, int res, uint32_t *id);

/* leader-thread nanoseconds per engine phase of the grid-parallel path:
 * {argmin, extract, sequential, group, solve, commit} — a profiling aid */
int  ns_dev_phase_ns(ns_dev *d, uint64_t out[16]);
/* memory-safety aid (compute-sanitizer is not available on every pool): with the environment variable NS_REDZONE=1 every
 * device table is allocated between two 256-byte red zones; *violations = red-zone bytes that were overwritten (0 = clean,
 * -1 = red zones not enabled) */
int  ns_dev_check(ns_dev *d, int *violations);
int  ns_ens_check(ns_ens *e, int *violations);
uint64_t ns_dev_launches(ns_dev *d);
void ns_dev_destroy(ns_dev *d);

/* replica ensembles: n independent scenarios resident on one GPU.  opts->mode NS_MODE_AUTO: one CTA per replica, the
 * hardware block scheduler is the replicas' work queue; NS_MODE_SEQ: one warp per replica.  ns_ens_run returns when
 * every replica is done (or has handled max_events_each events); the records are not copied out: a replica whose
 * bounded record buffer fills up is folded into its running digest on the device and resumed. */
int  ns_ens_create(ns_ens **out, const void *const *scenarios, const size_t *lens, int n,
		   const struct ns_dev_opts *opts);
int  ns_ens_run(ns_ens *e, uint64_t max_events_each, float *device_ms);
int  ns_ens_digests(ns_ens *e, struct ns_digest *out, int n);
/* analyzer reductions on the device, from each replica's record stream while it is produced (analyzer/analyzers.c:80-101 node
 * classes, :226-297 + :153-180 per-class bandwidth per period, :384-431 stale clients): call ns_ens_analysis_begin before ns_ens_run,
 * ns_ens_analysis afterwards.  usage[class][dir][k] = integral over period k of the summed R_USAGE of the class's nodes (Kbit;
 * divide by the period for the analyzer's per-period average), class 0 server / 1 cloud / 2 client, dir 0 out / 1 in;
 * stale[k] = client-seconds spent in N_STALE in period k.  Sums are order-dependent: ~1e-12 relative to the sequential analyzer. */
#define NS_ANA_BUCKETS 24
struct ns_ana_summary {
	uint64_t off, records, usage_records, state_records;
	double t_last;
	double usage[3][2][NS_ANA_BUCKETS];
	double stale[NS_ANA_BUCKETS];
	int32_t stale_now, stale_max;
};
int  ns_ens_analysis_begin(ns_ens *e, double period);
int  ns_ens_analysis(ns_ens *e, struct ns_ana_summary *out, int n);
int  ns_ens_counters_sum(ns_ens *e, uint64_t out[NS_CNT_COUNT]);   /* enum ns_counter, summed over the replicas (max for NS_CNT_MAX_BATCH) */
uint64_t ns_ens_launches(ns_ens *e);
void ns_ens_destroy(ns_ens *e);

/* batched source selection. Candidate tables are CSR: request q owns entries
 * [off[q], off[q+1]) of val/node (already filtered by is_connected / usability).
 * mode 1: keep the k[q] smallest by Lomuto quick-select; out_node/out_val (same CSR layout) hold
 *         the permuted table, out_cnt[q] = min(c, k).  mode 2: first strict minimum -> out_node[q]
 *         at off[q] (or -1), out_cnt[q] = 0/1.  Host buffers in, host buffers out. */
int  ns_dev_pick_batch(int device, int mode, int n_req, const int32_t *off, const int32_t *val,
		       const int32_t *node, const int32_t *k, int32_t *out_val, int32_t *out_node,
		       int32_t *out_cnt, float *device_ms);

/* the same selection with the score map on the device too, against the LIVE node table of a simulation:
 * metric 0 = distance_metric (modules/p2p_common.h:171-179), 1 = share_metric (:190-198: what the candidate's
 * upload share would be with one more flow, from its current total_bwupbound[0] / maximum_bandwidth[0]).
 * cand_node = device node index, cand_tz / client_tz = time zones (def_user.time_zone), usable (may be NULL)
 * = server_picker_opt2's filters (is_resource_usable / is_connected, client_behaviour.h:98-105) per candidate. */
int  ns_dev_pick_live(ns_dev *d, int mode, int metric, int n_req, const int32_t *off, const int32_t *cand_node,
		      const int32_t *cand_tz, const int32_t *client_tz, const uint8_t *usable, const int32_t *k,
		      int32_t *out_val, int32_t *out_node, int32_t *out_cnt, float *device_ms);

/* glibc TYPE_3 random(): state is 31 words + two indices */
struct ns_rand_state { int32_t st[31]; int32_t f, r; };
void ns_rand_seed_host(struct ns_rand_state *g, uint32_t seed);
uint32_t ns_rand_next_host(struct ns_rand_state *g);
void ns_rand_jump_host(struct ns_rand_state *g, uint64_t n);   /* == n calls of ns_rand_next_host */

const char *ns_strerror(int err);

#ifdef __cplusplus
}
#endif
#endif
