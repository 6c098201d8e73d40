/* t_containers.c — the host core's containers (netsim_b200/host/include: uthash.h, list.h, skiplist.h) behave like the reference's:
 * hashes iterate in insertion order and survive deletion inside HASH_ITER, skip lists stay sorted and answer find_ge / find_le,
 * lists are stacks (list_add = newest first).  Run by tests/test_host.py on the CPU. */
#include <stdio.h>
#include "uthash.h"
#include "list.h"
#include "skiplist.h"
const double eps = 1e-4;
struct e { int id; int v; UT_hash_handle hh; };
struct sk { double k; struct skip_list_head l; };
static int cmp(struct skip_list_head *a, void *key){ struct sk *s = skip_list_entry(a, struct sk, l); double k = *(double *)key; return s->k < k ? -1 : s->k > k ? 1 : 0; }
int main(void){
	struct e *h = NULL, *o, *t;
	for (int i = 0; i < 1000; i++) { struct e *x = calloc(1, sizeof *x); x->id = i*7919 % 10007; x->v = i; HASH_ADD_INT(h, id, x); }
	int k = 3*7919 % 10007; HASH_FIND_INT(h, &k, o); printf("find %d -> %d count %u\n", k, o ? o->v : -1, HASH_COUNT(h));
	int n = 0, ok = 1; HASH_ITER(hh, h, o, t) { if (o->v != n) ok = 0; if (n % 2) { HASH_DEL(h, o); free(o); } n++; }
	printf("iter in order %d, left %u\n", ok, HASH_COUNT(h));
	n = 0; HASH_ITER(hh, h, o, t) { if (o->v != 2*n) ok = 0; n++; HASH_DEL(h, o); free(o); }
	printf("order after del %d n %d head %p\n", ok, n, (void *)h);
	struct skip_list_head sh; skip_list_init_head(&sh);
	for (int i = 0; i < 200; i++) { struct sk *s = calloc(1, sizeof *s); s->k = (i*37) % 200; skip_list_insert(&sh, &s->l, &s->k, cmp); }
	double prev = -1; int cnt = 0; for (struct skip_list_head *p = sh.next[0]; p; p = p->next[0]) { struct sk *s = skip_list_entry(p, struct sk, l); if (s->k <= prev) ok = 0; prev = s->k; cnt++; }
	double q = 17.5; struct sk *ge = skip_list_entry(skip_list_find_ge(&sh, &q, cmp), struct sk, l), *le = skip_list_entry(skip_list_find_le(&sh, &q, cmp), struct sk, l);
	printf("skip sorted %d n %d ge %.0f le %.0f\n", ok, cnt, ge->k, le->k);
	struct li { int v; struct list_head l; } a = { 1 }, b = { 2 }, c3 = { 3 }, *it;
	LIST_HEAD(lh);
	list_add(&a.l, &lh); list_add(&b.l, &lh); list_add(&c3.l, &lh); list_del(&b.l);
	int seq = 0; list_for_each_entry(it, &lh, l) seq = seq*10+it->v;
	int rseq = 0; list_for_each_entry_reverse(it, &lh, l) rseq = rseq*10+it->v;
	printf("list %d %d empty %d\n", seq, rseq, list_empty(&lh));
	return !(ok && cnt == 200 && seq == 31 && rseq == 13);
}
