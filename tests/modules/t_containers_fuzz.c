/* t_containers_fuzz.c — randomized check of the host core's containers (netsim_b200/host/include) against plain arrays: a hash with
 * insertion-ordered iteration under random adds / finds / deletes (also inside HASH_ITER), and a skip list under random inserts /
 * deletes / find_ge / find_le with ->next[0] / ->prev[0] walks.  Run by tests/test_host.py on the CPU. */
#include <stdio.h>
#include <stdlib.h>
#include "uthash.h"
#include "list.h"
#include "skiplist.h"
const double eps = 1e-4;
static unsigned long long rs = 88172645463325252ULL;
static unsigned rnd(void){ rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (unsigned)(rs >> 11); }
struct e { int id; int alive; UT_hash_handle hh; };
struct sk { int k; int alive; struct skip_list_head l; };
static int cmp(struct skip_list_head *a, void *key){ struct sk *s = skip_list_entry(a, struct sk, l); return s->k-*(int *)key; }
#define N 4000
int main(void){
	static struct e el[N]; static int order[N*4]; int n_order = 0;      /* ids in insertion order (with holes = -1) */
	struct e *h = NULL, *o, *t;
	int bad = 0;
	for (int step = 0; step < 60000 && !bad; step++) {
		unsigned r = rnd();
		int id = (int)(rnd() % N);
		if (r % 3 == 0 && !el[id].alive && n_order < N*4) { el[id].id = id; el[id].alive = 1; HASH_ADD_INT(h, id, &el[id]); order[n_order++] = id; }
		else if (r % 3 == 1 && el[id].alive) { HASH_DEL(h, &el[id]); el[id].alive = 0; for (int i = 0; i < n_order; i++) if (order[i] == id) order[i] = -1; }
		else { HASH_FIND_INT(h, &id, o); if ((o != NULL) != (el[id].alive != 0) || (o && o != &el[id])) bad = 1; }
		if (step % 5000 == 4999) {                                       /* iteration == insertion order; delete every third inside the loop */
			int i = 0, k = 0;
			HASH_ITER(hh, h, o, t) {
				while (i < n_order && order[i] < 0) i++;
				if (i >= n_order || order[i] != o->id) { bad = 2; break; }
				if (k++ % 3 == 0) { HASH_DEL(h, o); o->alive = 0; order[i] = -1; }
				i++;
			}
			while (i < n_order && order[i] < 0) i++;
			if (!bad && i != n_order) bad = 3;
			unsigned cnt = 0; for (int q = 0; q < N; q++) cnt += el[q].alive;
			if (!bad && HASH_COUNT(h) != cnt) bad = 4;
		}
	}
	/* skip list: unique keys */
	static struct sk sl[N]; struct skip_list_head head; skip_list_init_head(&head);
	for (int step = 0; step < 60000 && !bad; step++) {
		unsigned r = rnd();
		int k = (int)(rnd() % N);
		if (r % 3 == 0 && !sl[k].alive) { sl[k].k = k; sl[k].alive = 1; skip_list_insert(&head, &sl[k].l, &k, cmp); }
		else if (r % 3 == 1 && sl[k].alive) { skip_list_delete(&sl[k].l); sl[k].alive = 0; }
		else {
			struct skip_list_head *ge = skip_list_find_ge(&head, &k, cmp), *le = skip_list_find_le(&head, &k, cmp);
			int wge = -1, wle = -1;
			for (int q = k; q < N; q++) if (sl[q].alive) { wge = q; break; }
			for (int q = k; q >= 0; q--) if (sl[q].alive) { wle = q; break; }
			if ((ge ? skip_list_entry(ge, struct sk, l)->k : -1) != wge || (le ? skip_list_entry(le, struct sk, l)->k : -1) != wle) bad = 5;
		}
		if (step % 5000 == 4999) {
			int prev = -1, cnt = 0, want = 0;
			struct skip_list_head *last = &head;
			for (struct skip_list_head *p = head.next[0]; p; p = p->next[0]) {
				struct sk *s = skip_list_entry(p, struct sk, l);
				if (s->k <= prev || !s->alive || p->prev[0] != last) { bad = 6; break; }
				prev = s->k; last = p; cnt++;
			}
			for (int q = 0; q < N; q++) want += sl[q].alive;
			if (!bad && cnt != want) bad = 7;
		}
	}
	printf("containers fuzz: %s (%d)\n", bad ? "FAILED" : "ok", bad);
	return bad;
}
