/* event.h — event_add / event_remove for the events a MODULE owns (USER events).  Core events live on
 * the device; a module's event_add becomes an NS_OP_SCHED through ns_dev_exec (event.h:56-74 there). */
#pragma once
#include "log.h"
#include "data.h"

void event_add(struct event *e, struct sim_state *s);
void event_remove(struct event *e);
static inline bool is_active(struct event *e){ return e && e->active; }
