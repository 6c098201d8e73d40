/* data.h — the structs a behaviour module may touch, as HOST MIRRORS of the device tables.
 * Field names follow the reference (data.h:15-179) because modules read them directly
 * (SURVEY §8b); the hot fields (speed, bandwidth_usage, range len ...) are refreshed from the device
 * before every module handler runs.  Containers (list_head, uthash, skip lists) are gone: the
 * device core is index based, the mirrors keep plain next pointers. */
#pragma once
#include <stdbool.h>
#include <stddef.h>
#include <sys/types.h>
#include "common.h"

struct sim_state;
struct flow;
struct node;
struct resource;

struct range {
	size_t start, total_len;      /* Kbit */
	double len, lenc;
	double last_update;
	struct flow *producer;
	struct resource *owner;
	struct range *next;           /* sorted by start */
};

struct resource {
	id_t resource_id;
	size_t len;
	int bit_rate;
	struct node *owner;
	int nproducer, nconsumer;
	bool auto_delete;
	struct range *first;          /* refreshed by print_range / range_get */
	struct resource *next_in_store;
	int script_index;             /* catalogue entry on the device */
};

struct store {
	size_t total_size;
	struct resource *head;
};

enum peer_type { P_SND = 0, P_RCV };

struct flow {
	id_t flow_id, resource_id;
	double bwupbound;
	double speed[2];
	double delay;
	struct node *peer[2];
	id_t dst_id;
	size_t start;
	double begin_time;
	struct event *done, *drain;
	struct range *drng, *srng;
	int ord;                      /* creation ordinal == device handle */
	bool closed;
	struct flow *out_next;        /* the sender's outgoing flows (node->outs there): is_connected */
};

enum node_state { N_OFFLINE = 0, N_STALE, N_PLAYING, N_DONE, N_IDLE, N_SERVER, N_CLOUD, N_DYING };

struct node {
	double maximum_bandwidth[2];
	double bandwidth_usage[2];
	double total_bwupbound[2];
	void *user_data;
	struct store *store;
	id_t node_id;
	enum node_state state;
	int index;                    /* device handle */
	struct flow *outs;            /* outgoing flows, newest first */
};

enum event_type { FLOW_DRAIN, FLOW_DONE, FLOW_SPEED_THROTTLE, PACKET_DONE, SPEED_CHANGE, USER, LAST_EVENT };

typedef void (*event_handler_func)(struct event *, struct sim_state *);

enum handler_priority { HNDR_DEFAULT = 0, HNDR_USER, HNDR_CLEANER };

struct event {
	enum event_type type;
	double time;
	bool active;
	void *data;
	bool auto_free;
	int token;                    /* host-owned USER events: handle on the device */
	int dev_slot;                 /* ... and the device slot of its pending instance (event_remove) */
};

#define SND 0
#define RCV 1

struct sim_state {
	double now;
	struct node *nodes;
	struct flow *flows;
	double (*bwcalc)(void *src, void *dst);
	double (*dlycalc)(void *src, void *dst);
	void *user_data;
	bool exit;
	void *core;                   /* struct ns_host_core, netsim_b200/host/core.c */
};

static inline struct event *event_new(double time, enum event_type t, void *d){
	struct event *e = talloc(1, struct event);
	e->time = time; e->type = t; e->data = d;
	e->token = -1; e->dev_slot = -1;
	return e;
}
static inline void event_free(struct event *e){ free(e); }

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
