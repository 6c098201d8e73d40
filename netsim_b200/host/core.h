/* core.h — entry points of the host core used by main.c (the driver) */
#pragma once
#include "data.h"
struct sim_state *ns_host_state_new(int device);
int ns_host_run(struct sim_state *s);
void ns_host_finish(struct sim_state *s);
