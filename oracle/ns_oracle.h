/* ns_oracle.h — TEST INFRASTRUCTURE: CPU restatement of netsim's flow/event hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may use it. */
#ifndef NS_ORACLE_H
#define NS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

typedef struct ns_oracle ns_oracle;

enum ns_oracle_status {
	NS_O_OK = 0,
	NS_O_BAD_SCENARIO = 1,
	NS_O_BACK_IN_TIME = 2,   /* event.h:60-64 would abort() */
	NS_O_NAN_TIME = 3,       /* event.h:67-70 would abort() */
	NS_O_NULL_RANGE = 4,     /* reference dereferences a NULL range (sim.c:69 -> range.c:21) */
	NS_O_NO_PRODUCER = 5,    /* flow.c:411 assert / NULL dereference */
};

enum ns_oracle_counter {
	NS_C_EVENTS = 0, NS_C_SOLVES, NS_C_FULL_SPREADS, NS_C_FLOW_VISITS, NS_C_SPEED_CHANGE,
	NS_C_FLOW_DONE, NS_C_DRAIN, NS_C_THROTTLE, NS_C_USER, NS_C_CREATES, NS_C_CLOSES,
	NS_C_DRAWS, NS_C_SEQ, NS_C_LIVE_FLOWS, NS_C_PEAK_FLOWS, NS_C_PENDING, NS_C_COUNT
};

struct ns_trace_row { uint32_t type, ord; double time, v0, v1; };

ns_oracle *ns_oracle_create(const void *scenario, size_t len, int want_trace, int want_records);
int ns_oracle_status(const ns_oracle *o);
/* pop/handle events until the queue is empty, the script ends the run, or max_events (0 = no
 * limit) events have been handled in total. Returns the status. */
int ns_oracle_run(ns_oracle *o, uint64_t max_events);
double ns_oracle_now(const ns_oracle *o);
const uint8_t *ns_oracle_records(const ns_oracle *o, size_t *len);
const struct ns_trace_row *ns_oracle_trace(const ns_oracle *o, size_t *rows);
void ns_oracle_counters(const ns_oracle *o, uint64_t out[NS_C_COUNT]);
/* state read-back for kernel-level parity: per flow ordinal {alive, speed0, speed1, drng_len} */
size_t ns_oracle_flow_state(const ns_oracle *o, double *out4, size_t max_flows);
size_t ns_oracle_node_state(const ns_oracle *o, double *usage_total4, size_t max_nodes);
void ns_oracle_destroy(ns_oracle *o);

/* glibc random()/srandom() TYPE_3 restatement (the id stream of sim.c:117,132 flow.c:327 and the
 * skip-list height draw of skiplist.h:34-40). */
struct ns_rand { int32_t st[31]; int f, r; };
void ns_rand_seed(struct ns_rand *g, uint32_t seed);
uint32_t ns_rand_next(struct ns_rand *g);

/* modules/qselect.h:5-27 instantiated as client_behaviour.h:63-65: Lomuto quick-select on
 * (node, val) pairs; leaves the k smallest in v[0..k-1] in the order the partition leaves them. */
void ns_oracle_qselect(int32_t *val, int32_t *node, size_t nmemb, size_t k);
/* server_picker_opt1 (client_behaviour.h:68-88) over a candidate table already filtered by
 * is_connected: returns the count kept (min(c,k)) and permutes val/node in place. */
int ns_oracle_pick_opt1(int32_t *val, int32_t *node, int c, int k);
/* server_picker_opt2 (client_behaviour.h:90-124): first strict minimum, -1 if none usable */
int ns_oracle_pick_opt2(const int32_t *val, const uint8_t *usable, int c);

/* distance_metric (metric 0) / share_metric (metric 1), modules/p2p_common.h:171-198 */
int ns_oracle_metric(int metric, int tz_node, int tz_client, double total_up, double max_up);

#endif
