/* timeshim.c — TEST INFRASTRUCTURE. LD_PRELOAD shim that pins time() so modules/p2p.c:123
 * (srandom(time(0))) becomes reproducible: time() returns NS_FAKE_TIME (default 1). */
#include <stdlib.h>
#include <time.h>
time_t time(time_t *t){
	const char *e = getenv("NS_FAKE_TIME");
	time_t v = e ? (time_t)strtoll(e, NULL, 10) : 1;
	if (t) *t = v;
	return v;
}
