/* data.h — the structs a behaviour module may touch, as HOST MIRRORS of the device tables.
 * Field names, types and containers follow the reference (data.h:15-179) because modules read them directly
 * (SURVEY §8b): list_head conns / consumers, uthash handles, a skip list of ranges per resource whose level-0 links
 * modules walk (->ranges.next[0] / .prev[0]).  The numeric fields (speeds, usage, range progress, completion times)
 * are refreshed from the device before every module handler runs; structure (lists, hashes, range order) is kept by
 * netsim_b200/host/core.c as flows are established, closed and merged.  The trailing "host core" fields of each
 * struct are ours. */
#pragma once
#include <stdbool.h>
#include <stddef.h>
#include <sys/types.h>
#include "list.h"
#include "skiplist.h"
#include "uthash.h"
#include "common.h"

struct sim_state;
struct flow;
struct node;
struct resource;

struct range {
	size_t start, total_len;      /* Kbit */
	double len, lenc;             /* lenc: Kahan compensation (range.h:91-106) */
	double last_update;
	struct skip_list_head ranges;
	struct list_head consumers;
	struct flow *producer;
	struct resource *owner;
	/* host core */
	int dev_slot;                 /* slot in the device's range table */
	double dev_len, dev_lenc, dev_last;   /* what the device holds: a module-side _range_update is written back */
	unsigned seen;                /* refresh epoch */
};

struct resource {
	id_t resource_id;
	size_t len;
	int bit_rate;
	struct skip_list_head ranges;
	UT_hash_handle hh;
	struct node *owner;
	int nproducer, nconsumer;
	bool auto_delete;
	/* host core */
	int dev_slot;                 /* slot in the device's resource table, -1 while it only exists on the host */
	int all_idx;                  /* position in the host core's list of resource copies */
	bool autodel_pushed;          /* the device knows about auto_delete */
};

struct store {
	size_t total_size;
	struct resource *rsrc_hash;
};

enum peer_type { P_SND = 0, P_RCV };

struct event;
struct flow {
	id_t flow_id, resource_id;
	UT_hash_handle hh;
	double bwupbound;
	double speed[2];
	double delay;
	struct node *peer[2];
	struct list_head conns[2];
	struct list_head spd_evs;
	UT_hash_handle hh2;
	id_t dst_id;
	size_t start;
	double begin_time;
	struct event *done, *drain;
	struct range *drng, *srng;
	struct list_head consumers;
	/* host core */
	int ord;                      /* creation ordinal == device handle */
	int live_idx;
	bool closed;
	unsigned seen;
};

enum node_state { N_OFFLINE = 0, N_STALE, N_PLAYING, N_DONE, N_IDLE, N_SERVER, N_CLOUD, N_DYING };

struct node {
	double maximum_bandwidth[2];
	double bandwidth_usage[2];
	double total_bwupbound[2];
	void *user_data;
	struct store *store;
	struct list_head conns[2];
	struct flow *outs;            /* hash (hh2) of the outgoing flows by destination id: is_connected, flow.h:29-35 */
	id_t node_id;
	UT_hash_handle hh;
	enum node_state state;
	/* host core */
	int index;                    /* device handle */
	bool on_device;
};

enum event_type { FLOW_DRAIN, FLOW_DONE, FLOW_SPEED_THROTTLE, PACKET_DONE, SPEED_CHANGE, USER, LAST_EVENT };

typedef void (*event_handler_func)(struct event *, struct sim_state *);

enum handler_priority { HNDR_DEFAULT = 0, HNDR_USER, HNDR_CLEANER };

struct event_handler { event_handler_func f; struct list_head handlers; enum handler_priority pri; };

struct event {
	enum event_type type;
	double time;
	bool active;
	void *data;
	struct skip_list_head events;
	bool auto_free;
	/* host core */
	int token;                    /* host-owned USER events: handle on the device */
	int dev_slot;                 /* ... and the device slot of its pending instance (event_remove) */
	int q_idx;                    /* ... or, while its event_add waits in the host core's queue, its position there */
};

#define SND 0
#define RCV 1
#define R_BANDWIDTH_USAGE 0
#define R_SPEED_CHANGE 2

struct sim_state {
	double now;
	struct skip_list_head events;          /* (unused: the event queue lives on the device) */
	struct list_head handlers[LAST_EVENT];
	struct node *nodes;                    /* hash by node_id */
	struct flow *flows;                    /* hash by flow_id */
	void *record_tail, *record_head;
	size_t record_file_size;
	int record_fd;
	double (*bwcalc)(void *src, void *dst);
	double (*dlycalc)(void *src, void *dst);
	void *user_data;
	bool exit;
	/* host core */
	void *core;                   /* struct ns_host_core, netsim_b200/host/core.c */
};

static inline struct store *store_new(void){ return talloc(1, struct store); }
static inline struct event *event_new(double time, enum event_type t, void *d){
	struct event *e = talloc(1, struct event);
	e->time = time; e->type = t; e->data = d;
	e->token = -1; e->dev_slot = -1; e->q_idx = -1;
	return e;
}
static inline void event_free(struct event *e){ free(e); }
static inline struct range *range_new(size_t start, size_t len){
	struct range *rng = talloc(1, struct range);
	rng->start = start; rng->len = len; rng->dev_slot = -1;
	INIT_LIST_HEAD(&rng->consumers);
	return rng;
}
static inline struct resource *resource_new(id_t id, size_t s){
	struct resource *r = talloc(1, struct resource);
	skip_list_init_head(&r->ranges);
	r->len = s; r->resource_id = id; r->dev_slot = -1;
	return r;
}

static inline const char *strstate(int state){
	switch (state) {
	case N_DONE: return "Done";
	case N_PLAYING: return "Playing";
	case N_IDLE: return "Idle";
	case N_STALE: return "Stale";
	case N_OFFLINE: return "Offline";
	}
	return NULL;
}
