/* recfile.h — saving a record byte stream as file(s) in the reference's layout (recfile.c) */
#pragma once
#include <stddef.h>
#include <stdint.h>
#define NS_RECFILE_MAX 0xFFFFFFFFull      /* what the uint32 length field of record.c:111-115 can say */
/* split_bytes 0: one file like the reference (returns 1 and complains when len wraps the length field);
 * otherwise pieces <path>, <path>.1, ... of at most split_bytes payload bytes each, cut at record boundaries.
 * 0 ok, 1 wrapped length, -1 I/O error, -2 split_bytes < one record, -3 buf is not a record stream, -4 the preamble fills a piece */
#define NS_RECFILE_PREAMBLE 1             /* continuation pieces start with the node-identity records of the earlier pieces */
int ns_recfile_save(const char *path, const uint8_t *buf, size_t len, size_t split_bytes, int flags, int *n_pieces);
size_t ns_recfile_split_from_env(void);   /* NS_RECORD_SPLIT */
int ns_recfile_flags_from_env(void);      /* NS_RECORD_SPLIT_PREAMBLE */
