/* ref_pick.c — TEST INFRASTRUCTURE.  The reference's OWN candidate selection, compiled from the sources
 * where they lie: this file includes modules/client_behaviour.h of the reference, i.e. the qselect
 * instantiation of client_behaviour.h:63-65 over modules/qselect.h:5-27 and the fetch metrics of
 * modules/p2p_common.h:171-198, and exposes them on flat arrays so that tests/golden/make_golden.py can
 * record what the reference does (tests/golden/pick*.npz).  Built by oracle/Makefile `ref` into
 * oracle/_ref/libref_pick.so; nothing in the product links or loads it. */
#include <assert.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#define LOG_DOMAIN "ref_pick"
#include "client_behaviour.h"

int global_log_level = 0;
const double eps = 1e-4;

/* qselect_eval(eval_table, c, k) exactly as server_picker_opt1 calls it (client_behaviour.h:77-87):
 * returns the count kept, val/node permuted in place */
int ref_pick_opt1(int32_t *val, int32_t *node, int c, int k){
	if (c <= k) return c;
	struct nv_pair *t = calloc((size_t)c, sizeof *t);
	for (int i = 0; i < c; i++) { t[i].n = (struct node *)(intptr_t)(node[i]+1); t[i].val = val[i]; }
	qselect_eval(t, (size_t)c, (size_t)k);
	for (int i = 0; i < c; i++) { node[i] = (int32_t)((intptr_t)t[i].n-1); val[i] = t[i].val; }
	free(t);
	return k;
}

/* distance_metric (metric 0) / share_metric (metric 1) on two real struct node + def_user */
int ref_metric(int metric, int tz_node, int tz_client, double total_up, double max_up){
	struct node n, c;
	struct def_user dn, dc;
	memset(&n, 0, sizeof n); memset(&c, 0, sizeof c); memset(&dn, 0, sizeof dn); memset(&dc, 0, sizeof dc);
	dn.time_zone = tz_node; dc.time_zone = tz_client;
	n.user_data = &dn; c.user_data = &dc;
	n.total_bwupbound[0] = total_up; n.maximum_bandwidth[0] = max_up;
	return metric == 0 ? distance_metric(&n, &c) : share_metric(&n, &c);
}
