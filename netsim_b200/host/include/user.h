#pragma once
#include "data.h"
void user_speed_change(struct event *e, struct sim_state *s);   /* empty stubs, user.c:7-13 there */
void user_done(struct event *e, struct sim_state *s);
