/* recfile.c — the record FILE of the host core: the device's record byte stream saved in the reference's on-disk
 * layout (record.c:111-116 and :119-153 — a 4-byte big-endian payload length, the payload, the file size 1 MiB
 * doubled as record.c:38-51 does; what analyzer/record_reader.c:11-27 maps).
 *
 * The reference's length field is a uint32: a run whose records pass 4 GiB wraps it ("overflows
 * its 4 GiB record header") and the file becomes unreadable.  Default here = the reference's behaviour plus a loud
 * message.  Behind a flag (NS_RECORD_SPLIT) the stream is cut at RECORD boundaries into
 * <path>, <path>.1, <path>.2, ... — every piece a complete file in the reference's layout, so the unmodified analyzer
 * reads each of them — and nothing is lost.  The analyzer's node tracker dereferences a node it has not seen created
 * (analyzers.c:106-110), so with NS_RECORD_SPLIT_PREAMBLE=1 every continuation piece starts with a copy of the
 * R_NODE_CREATE / server- and cloud-class R_NODE_STATE records of the pieces before it (original timestamps): the
 * per-class reports then run on each piece alone; without it the pieces are an exact partition of the stream. */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "recfile.h"

/* length of the record that starts at p (13-byte header, then 1 / 2 / 4 bytes or a 4+2-byte double; record.c:76-106);
 * 0 = not a record */
static size_t rec_size(const uint8_t *p, size_t left){
	if (left < 13) return 0;
	int8_t bytes = (int8_t)p[12];
	size_t v = bytes == 1 ? 1 : bytes == 2 ? 2 : bytes == 4 ? 4 : bytes == -1 ? 6 : 0;
	if (!v || 13+v > left) return 0;
	return 13+v;
}

static int save_piece(const char *path, const uint8_t *pre, size_t pre_len, const uint8_t *buf, size_t len){
	FILE *f = fopen(path, "wb");
	if (!f) return -1;
	size_t tot = pre_len+len;
	uint8_t h[4] = { (uint8_t)(tot >> 24), (uint8_t)(tot >> 16), (uint8_t)(tot >> 8), (uint8_t)tot };   /* record.c:111-115 */
	int bad = fwrite(h, 1, 4, f) != 4 || (pre_len && fwrite(pre, 1, pre_len, f) != pre_len) || (len && fwrite(buf, 1, len, f) != len);
	size_t size = 1024*1024;                                /* record.c:133 initial size, doubled at :38-51 */
	while (4+tot+19 >= size) size <<= 1;
	if (ftruncate(fileno(f), (off_t)size)) bad = 1;
	if (fclose(f)) bad = 1;
	return bad ? -1 : 0;
}

/* the records the analyzer's node tracker needs before it sees anything else of a node (analyzer/analyzers.c:81-116):
 * R_NODE_CREATE, and the R_NODE_STATE records that make a node a server or a cloud node */
static bool is_node_identity(const uint8_t *r){
	if (r[0] != 0) return false;
	if (r[1] == 8) return true;                             /* R_NODE_CREATE, record.h:15 */
	return r[1] == 7 && r[12] == 1 && (r[13] == 5 || r[13] == 6);   /* R_NODE_STATE to N_SERVER / N_CLOUD, data.h:88-97 */
}

int ns_recfile_save(const char *path, const uint8_t *buf, size_t len, size_t split_bytes, int flags, int *n_pieces){
	if (n_pieces) *n_pieces = 0;
	if (split_bytes > NS_RECFILE_MAX) split_bytes = NS_RECFILE_MAX;
	if (!split_bytes) {                                     /* the reference's behaviour: one file, the length mod 2^32 */
		if (len > NS_RECFILE_MAX)
			fprintf(stderr, "record file %s: %zu bytes of records do not fit the 32-bit length field (record.c:111-115); "
				"set NS_RECORD_SPLIT=1 to get a series of readable files\n", path, len);
		if (save_piece(path, NULL, 0, buf, len)) return -1;
		if (n_pieces) *n_pieces = 1;
		return len > NS_RECFILE_MAX ? 1 : 0;
	}
	if (split_bytes < 19) return -2;                        /* no room for one record */
	size_t at = 0; int k = 0, rc = 0;
	size_t plen = strlen(path);
	char *name = malloc(plen+16);
	uint8_t *pre = NULL; size_t pre_len = 0, pre_cap = 0;  /* NS_RECFILE_PREAMBLE: node identities seen in earlier pieces */
	if (!name) return -1;
	do {
		if (pre_len+19 > split_bytes) { rc = -4; break; }   /* the preamble leaves no room for a record */
		size_t end = at;                                    /* the longest run of whole records that fits */
		while (end < len) {
			size_t r = rec_size(buf+end, len-end);
			if (!r) { rc = -3; break; }                     /* not a record stream */
			if (pre_len+(end+r-at) > split_bytes) break;
			end += r;
		}
		if (rc) break;
		if (k) snprintf(name, plen+16, "%s.%d", path, k); else memcpy(name, path, plen+1);
		if (save_piece(name, pre, pre_len, buf+at, end-at)) { rc = -1; break; }
		if (flags & NS_RECFILE_PREAMBLE)
			for (size_t p = at; p < end; ) {
				size_t r = rec_size(buf+p, len-p);
				if (is_node_identity(buf+p)) {
					if (pre_len+r > pre_cap) {
						pre_cap = pre_cap ? pre_cap*2 : 1 << 16;
						uint8_t *np = realloc(pre, pre_cap);
						if (!np) { rc = -1; break; }
						pre = np;
					}
					memcpy(pre+pre_len, buf+p, r); pre_len += r;
				}
				p += r;
			}
		if (rc) break;
		at = end; k++;
	} while (at < len);
	free(name); free(pre);
	if (rc) return rc;
	if (n_pieces) *n_pieces = k;
	return 0;
}

/* NS_RECORD_SPLIT: unset / 0 = off; 1 = pieces of at most 2^32-1 payload bytes; any larger number = that many bytes */
size_t ns_recfile_split_from_env(void){
	const char *e = getenv("NS_RECORD_SPLIT");
	if (!e || !*e) return 0;
	unsigned long long v = strtoull(e, NULL, 0);
	if (v == 0) return 0;
	if (v == 1) return NS_RECFILE_MAX;
	return (size_t)v;
}
int ns_recfile_flags_from_env(void){
	const char *e = getenv("NS_RECORD_SPLIT_PREAMBLE");
	return e && *e && *e != '0' ? NS_RECFILE_PREAMBLE : 0;
}
