/* ns_oracle_main.c — TEST INFRASTRUCTURE: command-line front end of the CPU restatement.
 * usage: ns_oracle <scenario> [--record FILE] [--trace FILE] [--max-events N] [--stats FILE] */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "ns_oracle.h"

int main(int argc, char **argv){
	const char *sc = NULL, *rec = NULL, *tr = NULL, *st = NULL;
	unsigned long long maxev = 0;
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "--record") && i+1 < argc) rec = argv[++i];
		else if (!strcmp(argv[i], "--trace") && i+1 < argc) tr = argv[++i];
		else if (!strcmp(argv[i], "--stats") && i+1 < argc) st = argv[++i];
		else if (!strcmp(argv[i], "--max-events") && i+1 < argc) maxev = strtoull(argv[++i], NULL, 10);
		else sc = argv[i];
	}
	if (!sc) { fprintf(stderr, "usage: %s <scenario> [--record F] [--trace F] [--max-events N] [--stats F]\n", argv[0]); return 2; }
	FILE *f = fopen(sc, "rb");
	if (!f) { perror(sc); return 2; }
	fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
	void *buf = malloc(n);
	if (fread(buf, 1, n, f) != (size_t)n) return 2;
	fclose(f);
	struct timespec t0, t1, t2;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	ns_oracle *o = ns_oracle_create(buf, n, tr != NULL, rec != NULL);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	int status = ns_oracle_run(o, maxev);
	clock_gettime(CLOCK_MONOTONIC, &t2);
	uint64_t c[NS_C_COUNT];
	ns_oracle_counters(o, c);
	if (rec) {
		size_t len; const uint8_t *p = ns_oracle_records(o, &len);
		FILE *g = fopen(rec, "wb");
		uint8_t h[4] = { len>>24, len>>16, len>>8, len };
		fwrite(h, 1, 4, g); fwrite(p, 1, len, g); fclose(g);
	}
	if (tr) {
		size_t rows; const struct ns_trace_row *p = ns_oracle_trace(o, &rows);
		FILE *g = fopen(tr, "wb"); fwrite(p, sizeof *p, rows, g); fclose(g);
	}
	FILE *g = st ? fopen(st, "w") : stdout;
	fprintf(g, "{\"status\": %d, \"events\": %llu, \"solves\": %llu, \"full_spreads\": %llu, \"flow_visits\": %llu, "
		"\"speed_change\": %llu, \"flow_done\": %llu, \"drain\": %llu, \"throttle\": %llu, \"user\": %llu, "
		"\"creates\": %llu, \"closes\": %llu, \"draws\": %llu, \"peak_flows\": %llu, \"pending\": %llu, "
		"\"init_seconds\": %.9f, \"loop_seconds\": %.9f, \"now\": %.17g}\n", status,
		(unsigned long long)c[NS_C_EVENTS], (unsigned long long)c[NS_C_SOLVES], (unsigned long long)c[NS_C_FULL_SPREADS],
		(unsigned long long)c[NS_C_FLOW_VISITS], (unsigned long long)c[NS_C_SPEED_CHANGE], (unsigned long long)c[NS_C_FLOW_DONE],
		(unsigned long long)c[NS_C_DRAIN], (unsigned long long)c[NS_C_THROTTLE], (unsigned long long)c[NS_C_USER],
		(unsigned long long)c[NS_C_CREATES], (unsigned long long)c[NS_C_CLOSES], (unsigned long long)c[NS_C_DRAWS],
		(unsigned long long)c[NS_C_PEAK_FLOWS], (unsigned long long)c[NS_C_PENDING],
		(t1.tv_sec-t0.tv_sec)+1e-9*(t1.tv_nsec-t0.tv_nsec), (t2.tv_sec-t1.tv_sec)+1e-9*(t2.tv_nsec-t1.tv_nsec),
		ns_oracle_now(o));
	if (st) fclose(g);
	ns_oracle_destroy(o);
	return status;
}
