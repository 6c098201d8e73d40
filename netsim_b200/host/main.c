/* main.c — `netsim_dev <module-path-without-.so>`: the reference's driver (main.c:19-69 there) on
 * the device core: dlopen the behaviour module, call <basename>_init, run the event loop. */
#include <dlfcn.h>
#include <libgen.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LOG_DOMAIN "main"
#include "log.h"
#include "data.h"
#include "core.h"

int global_log_level = LOG_INFO;

int main(int argc, const char **argv){
	if (argc != 2) {
		fprintf(stderr, "Usage: %s module\n", argv[0]);
		return 1;
	}
	char *filename = strdup(argv[1]);
	char *tmp = malloc(strlen(filename)+6);
	sprintf(tmp, "%s.so", filename);
	void *mh = dlopen(tmp, RTLD_LAZY);
	if (!mh) {
		fprintf(stderr, "%s\n", dlerror());
		return 1;
	}
	sprintf(tmp, "%s_init", basename(filename));
	void (*minit)(struct sim_state *) = (void (*)(struct sim_state *))dlsym(mh, tmp);
	if (!minit) {
		fprintf(stderr, "%s\n", dlerror());
		return 1;
	}
	const char *dev = getenv("NS_DEVICE");
	struct sim_state *s = ns_host_state_new(dev ? atoi(dev) : 0);
	minit(s);
	free(filename);
	free(tmp);
	int rc = ns_host_run(s);
	ns_host_finish(s);
	return rc ? 1 : 0;
}
