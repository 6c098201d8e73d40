// ns_ana_dev.cuh — the analyzer reductions as a warp-level device function (included by ns_ana.cu for the stand-alone kernel and
// by the replica engines, which fold their own record buffer when it fills up instead of stopping: ns_seq.cuh, ns_par_kernel.inc)
#pragma once
#include "ns_ana.cuh"
#define FULL 0xffffffffu
#define ANA_TILE 1024
namespace {
__device__ __forceinline__ int ana_lookup(const AnaStream &A, uint32_t id){
	uint32_t h = (id*2654435761u) & A.hmask;
	for (int probe = 0; probe <= (int)A.hmask; probe++) {
		const int v = A.hval[h];
		if (v < 0) return -1;
		if (A.hkey[h] == id) return v;
		h = (h+1) & A.hmask;
	}
	return -1;
}
// add v*(b-a) to the per-period buckets the interval [a, b) overlaps
__device__ __forceinline__ void ana_spread(double *bucket, int nb, double period, double a, double b, double v){
	if (!(b > a) || v == 0.0) return;
	int k = (int)(a/period);
	while (a < b && k < nb) {
		const double e = fmin(b, (double)(k+1)*period);
		atomicAdd(&bucket[k], v*(e-a));
		a = e; k++;
	}
}

// one warp: everything in the stream's buffer that has not been analysed yet (tile / starts: ANA_TILE+32 bytes and 34 ints of shared memory)
__device__ __noinline__ void ana_process(AnaStream &A, uint8_t *tile, int *starts, int lane, int final_flush){
	const uint64_t base = *(volatile const unsigned long long *)A.rec_base, len = *(volatile const unsigned long long *)A.rec_len;
	uint64_t p = A.st->off > base ? A.st->off-base : 0;          // first unanalysed byte, relative to the buffer
	const double period = A.period;
	const int nb = NS_ANA_BUCKETS;
	while (p < len) {
		const int avail = (int)((len-p) < ANA_TILE ? (len-p) : ANA_TILE);
		for (int i = lane; i < avail; i += 32) tile[i] = A.rec[p+i];
		__syncwarp();
		if (lane == 0) {                                              // record boundaries of up to 32 whole records
			int q = 0, k = 0;
			while (k < 32 && q+13 < avail) {
				const int8_t bytes = (int8_t)tile[q+12];
				const int sz = 13+(bytes == 1 ? 1 : bytes == 4 ? 4 : bytes == 2 ? 2 : 6);
				if (q+sz > avail) break;
				starts[k++] = q; q += sz;
			}
			starts[k] = q; starts[33] = k;
		}
		__syncwarp();
		const int nrec = starts[33];
		if (nrec == 0) break;                                         // cannot happen: records are written whole
		int minor = -1, node = -1, u8v = 0; double t = 0, val = 0;
		if (lane < nrec) {
			const uint8_t *r = tile+starts[lane];
			minor = r[0] == 0 ? r[1] : -1;                            // major 0 only (analyzers.c:83-84)
			const uint32_t id = ((uint32_t)r[2] << 24) | ((uint32_t)r[3] << 16) | ((uint32_t)r[4] << 8) | r[5];
			const uint32_t ts = ((uint32_t)r[6] << 24) | ((uint32_t)r[7] << 16) | ((uint32_t)r[8] << 8) | r[9];
			const uint32_t tms = ((uint32_t)r[10] << 8) | r[11];
			t = (double)ts+((double)tms)/1000.0;                      // record_reader.c:40-43
			const int8_t bytes = (int8_t)r[12];
			if (bytes == 1) u8v = r[13];
			else if (bytes == -1) {
				const uint32_t ip = ((uint32_t)r[13] << 24) | ((uint32_t)r[14] << 16) | ((uint32_t)r[15] << 8) | r[16];
				val = (double)ip+((double)(((uint32_t)r[17] << 8) | r[18]))/1000.0;   // record_reader.c:62-68
			}
			if (minor == 0 || minor == 1 || minor == 7) node = ana_lookup(A, id);
		}
		// ---- node classes first (a class change and a usage record of the same node in one tile keep their order:
		//      state records of servers precede any traffic, analyzers.c:86-92)
		if (minor == 7 && node >= 0) {
			if (u8v == 6) A.ntype[node] = 1;                          // N_CLOUD -> CLD
			else if (u8v == 5) A.ntype[node] = 0;                     // N_SERVER -> SVR
		}
		__syncwarp();
		// ---- usage records: integral of the node's previous usage since its previous record
		{
			const bool us = (minor == 0 || minor == 1) && node >= 0;
			const int dir = minor & 1;
			const unsigned peers = __match_any_sync(FULL, us ? node*2+dir : -1-lane);
			const unsigned before = peers & ((1u << lane)-1u);
			const int pl = before ? 31-__clz(before) : -1;            // previous record of the same (node, dir) in this tile
			const double pt_l = __shfl_sync(FULL, t, pl < 0 ? lane : pl), pv_l = __shfl_sync(FULL, val, pl < 0 ? lane : pl);
			double pt = pt_l, pv = pv_l;
			if (us && pl < 0) { pt = A.u_time[dir][node]; pv = A.u_val[dir][node]; }   // first of its (node, dir) in the tile: the table
			__syncwarp();                                             // every lane has read the table before any lane writes it back
			if (us) {
				const int cls = A.ntype[node];
				ana_spread(A.st->usage[cls][dir], nb, period, pt, t, pv);
				if (!(peers & ~((2u << lane)-1u))) { A.u_time[dir][node] = t; A.u_val[dir][node] = val; }   // last of its (node, dir) in the tile
			}
		}
		// ---- client state records: stale time per node, concurrent stale count in stream order
		{
			const bool sr = minor == 7 && node >= 0 && A.ntype[node] == 2;
			const unsigned peers = __match_any_sync(FULL, sr ? node : -1-lane);
			const unsigned before = peers & ((1u << lane)-1u);
			const int pl = before ? 31-__clz(before) : -1;
			const int ps_l = __shfl_sync(FULL, u8v, pl < 0 ? lane : pl);
			const double pt_l = __shfl_sync(FULL, t, pl < 0 ? lane : pl);
			int delta = 0;
			int ps = ps_l; double pt = pt_l;
			if (sr && pl < 0) { ps = A.n_state[node]; pt = A.n_since[node]; }
			__syncwarp();                                             // (as above: gather, then update)
			if (sr) {
				if (ps == 1) ana_spread(A.st->stale, nb, period, pt, t, 1.0);          // was N_STALE since pt
				if (ps != u8v) delta = (u8v == 1) - (ps == 1);                            // analyzers.c:411-424
				if (!(peers & ~((2u << lane)-1u))) { A.n_state[node] = (int8_t)u8v; A.n_since[node] = t; }
			}
			int run = delta;
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(FULL, run, d); if (lane >= d) run += y; }
			int cur = A.st->stale_now+run;
			int mx = cur;
			#pragma unroll
			for (int d = 16; d; d >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, d));
			const int tot = __shfl_sync(FULL, run, 31);
			const unsigned nsr = __popc(__ballot_sync(FULL, sr)), nus = __popc(__ballot_sync(FULL, (minor == 0 || minor == 1) && node >= 0));
			double tmax = lane < nrec ? t : 0.0;
			#pragma unroll
			for (int d = 16; d; d >>= 1) tmax = fmax(tmax, __shfl_xor_sync(FULL, tmax, d));
			if (lane == 0) {
				A.st->stale_now += tot;
				if (mx > A.st->stale_max) A.st->stale_max = mx;
				A.st->records += nrec; A.st->state_records += nsr; A.st->usage_records += nus;
				if (tmax > A.st->t_last) A.st->t_last = tmax;
			}
		}
		__syncwarp();
		p += starts[nrec];
	}
	if (lane == 0) A.st->off = base+len;
	__syncwarp();
	if (final_flush) {
		// what every node still carries runs until the last record of the stream
		const double t_end = A.st->t_last;
		for (int i = lane; i < A.n_nodes; i += 32) {
			const int cls = A.ntype[i];
			for (int d = 0; d < 2; d++) { ana_spread(A.st->usage[cls][d], nb, period, A.u_time[d][i], t_end, A.u_val[d][i]); A.u_time[d][i] = t_end; }
			if (cls == 2 && A.n_state[i] == 1) { ana_spread(A.st->stale, nb, period, A.n_since[i], t_end, 1.0); A.n_since[i] = t_end; }
		}
	}
}


}   // namespace
