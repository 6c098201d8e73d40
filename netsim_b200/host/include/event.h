/* event.h — event_add / event_remove for the events a MODULE owns (USER events).  Core events live on
 * the device; a module's event_add becomes an NS_OP_SCHED through ns_dev_exec (event.h:56-74 there).  A flow's
 * done / drain events are mirrors: active + time are refreshed from the device, which is all is_later_than reads. */
#pragma once
#ifndef LOG_DOMAIN
# define LOG_DOMAIN "event"     /* (the reference's event.h names the domain of whoever includes it first) */
#endif
#include "log.h"
#include "skiplist.h"
#include "data.h"
#include <math.h>

void event_add(struct event *e, struct sim_state *s);
void event_remove(struct event *e);
static inline bool is_active(struct event *e){ return e && e->active; }
static inline bool is_later_than(double time, struct event *b){   /* event.h:82-88 */
	if (!b->active) return false;
	return time > b->time+eps;
}
