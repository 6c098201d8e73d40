/* ns_scenario.h — binary "scenario script" format shared by the three executors of the
 * netsim flow/event hot path:
 *   1. the UNMODIFIED reference core (/root/reference, built into oracle/_ref/) driven by
 *      oracle/ref_scenario.c, a behaviour module written against the reference's module API
 *      (sim.h:14-23, data.h:124-136);
 *   2. the CPU restatement oracle/ns_oracle.c;
 *   3. the sm_100a CUDA engine netsim_b200/csrc/ (through include/netsim_b200.h).
 * A scenario is everything a behaviour module would do to the core, written down as data:
 * nodes with their capacities (modules/p2p.c:144-170 writes node->maximum_bandwidth the same way),
 * resources (sim_node_new_resource / sim_node_add_resource, sim.c:129-154), and groups of
 * operations that run at init, at scheduled USER events (event.h:56-74) or when a given flow
 * completes (a HNDR_USER handler on FLOW_DONE, like modules/test04.c:23-35).
 * All integers little-endian, doubles IEEE-754 binary64.
 */
#ifndef NS_SCENARIO_H
#define NS_SCENARIO_H
#include <stdint.h>

#define NS_SC_MAGIC   0x4353534eu /* "NSSC" */
#define NS_SC_VERSION 1u

enum ns_op_kind {
	NS_OP_CREATE = 1, /* sim_establish_flow(rid[c], start, node[a], node[b]); bwcalc -> x, dlycalc -> y */
	NS_OP_SCHED  = 2, /* event_add(event_new(c ? now+x : x, USER, group a)) */
	NS_OP_CLOSE  = 3, /* flow_close(flow created by the a-th executed CREATE), if still alive */
	NS_OP_END    = 4, /* sim_end(): the event loop stops before the next pop (main.c:51-52) */
	NS_OP_CANCEL = 6, /* ns_dev_exec only: event_remove (event.h:13-18) of a module-owned USER event; a = the slot ns_dev_exec
	                     returned in .b for its NS_OP_SCHED */
	NS_OP_PLAY1  = 5, /* client_new_play1(rid[c], node[a]) (modules/client_behaviour.c:364-387) with fetch metric b
	                     (0 = distance_metric, 1 = share_metric, p2p_common.h:171-198): server_picker_opt1 over the
	                     server nodes + qselect, one sim_establish_flow per kept server at strip i*strip with
	                     distance_based_bw / distance_based_delay, then client_start_play.  Needs NS_SC_F_CLIENTS. */
	/* ---- ns_dev_exec only: what the host core (netsim_b200/host) needs to mirror a module that is code, not data ---- */
	NS_OP_NEWNODE    = 7,  /* sim_create_node (sim.c:113-125) with the id the host drew: a = node index (== the current node
	                          count, < ns_dev_opts.cap_nodes), c = node id, x / y = maximum_bandwidth[0 / 1] */
	NS_OP_NEWRES     = 8,  /* sim_node_new_resource / sim_node_add_resource (sim.c:129-154): a complete copy of resource id c,
	                          `start` Kbit long, bit rate b, on node a; on return .b = its slot in the resource table, .c = the
	                          slot of its range [0, len) */
	NS_OP_CREATE_RID = 9,  /* NS_OP_CREATE with the resource given by its id (c) instead of a script index; on return .a = creation
	                          ordinal or -1, .b = the destination's resource slot, .c = the new destination range's slot,
	                          .start = the source range's slot */
	NS_OP_DELRES     = 11, /* node_del_resource (sim.h:47-53): a = resource slot; its ranges are freed, the node forgets it */
	NS_OP_AUTODEL    = 12, /* the module set resource->auto_delete (resource.c:63): a = resource slot; flow_close deletes the resource
	                          when its last consumer / producer has gone (flow.c:279-281) */
	NS_OP_RANGE_SET  = 10, /* a module handler advanced a range itself (_range_update, range.h:91-118, from is_resource_usable /
	                          client_lowwm_event ...): a = range slot, x = len, y = lenc, start = the bits of last_update */
};

/* header.flags */
#define NS_SC_F_CLIENTS 1u  /* a struct ns_sc_client per node follows the on-done table: the nodes run the behaviour
                               library's client side (modules/client_behaviour.c: water-mark state machine on every
                               RCV speed change / FLOW_DONE / throttle, its USER events, R_NODE_STATE records) */
enum ns_client_kind {
	NS_CK_PLAIN  = 0,   /* no behaviour (user_data == NULL) */
	NS_CK_CLIENT = 1,   /* p2p client: def_user with lowwm/highwm, state N_IDLE (modules/p2p.c:166-174) */
	NS_CK_SERVER = 2,   /* p2p_new_server (p2p_common.h:101-115): state N_SERVER, member of the picker's server list */
};
struct ns_sc_client { int32_t kind, time_zone, lowwm, highwm; };

struct ns_sc_header {
	uint32_t magic, version;
	uint32_t seed;       /* srandom(seed) before anything else; 1 == glibc default state */
	uint32_t flags;      /* NS_SC_F_* */
	int32_t n_nodes, n_resources, n_copies, n_groups, n_ops, n_ondone;
	int32_t pad[2];
};                           /* 48 bytes, followed by the arrays below in this order */
/* double max_up[n_nodes];  double max_down[n_nodes];          node->maximum_bandwidth[0|1] */
struct ns_sc_resource { int32_t owner; int32_t bit_rate; int64_t len; };   /* sim_node_new_resource */
struct ns_sc_copy     { int32_t node;  int32_t res; };                    /* sim_node_add_resource */
/* int32_t group_off[n_groups+1];   group g = ops[group_off[g] .. group_off[g+1]); group 0 runs at init */
struct ns_sc_op { int32_t kind, a, b, c; int64_t start; double x, y; };  /* 40 bytes */
struct ns_sc_ondone { int32_t flow_ord; int32_t group; };                /* sorted by flow_ord, unique */
/* struct ns_sc_client client[n_nodes];   only with NS_SC_F_CLIENTS */

#endif
