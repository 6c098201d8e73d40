/* common.h — host shim for netsim modules on the B200 device layer.
 * Same names as the reference header of this name (common.h:1-20) so that module sources compile
 * unchanged; written for the index-based device core, not copied. */
#pragma once
#include <stddef.h>
#include <stdlib.h>

#define talloc(nmemb, type) ((type *)calloc((nmemb), sizeof(type)))
#define fequ(a, b) ((a)-(b) > -eps && (a)-(b) < eps)
/* the struct that embeds `member` at address ptr */
#define container_of(ptr, type, member) ((type *)((char *)(ptr)-offsetof(type, member)))

extern const double eps;   /* 1e-4, sim.c:11 */
