/* k_gridsweep.cu -- the perspective-grid build of one view (FrustumGrid::buildGrid, frustum_grid.h:210-366; rows a2-a8)
 * as THREE launches, each a single sweep over its data:
 *
 *   k_sweep_setup   per triangle: Moeller-Trumbore record + clamped cell rectangle + pair count (a2/a3,
 *                   grid_kernel.cu:13-36,164-243) and, in the same pass, the exclusive scan of the counts (a4,
 *                   frustum_grid.h:249; decoupled look-back between tiles).  The rectangle also tells how many pairs
 *                   the triangle puts into every cell ROW and every cell COLUMN, so the digit histograms of both sort
 *                   passes are accumulated here as difference arrays (4 shared-memory atomics per triangle) and never
 *                   need a pass over the pairs.
 *   k_sweep_pass<0> per output pair, GENERATED not read: pair p of the triangle-major emission order of DSFillkernel
 *                   (a5, grid_kernel.cu:245-332) -> (cell, triangle), per-cell histogram `span` (a7), and the first
 *                   pass of a stable LSD sort keyed on the cell row.
 *   k_sweep_pass<1> second pass keyed on the cell column -> list ordered by cell = column*rows + row, ascending
 *                   triangle id inside a cell (what cudppSort's stability gives the reference, frustum_grid.h:298);
 *                   the first tiles of the same launch scan `span` over ALL cells into `offset` (a7, frustum_grid.h:361).
 *
 * The key is (column, row) with at most 1024 values each, so two passes always suffice (the 8-bit LSD sort needed 2-3 passes
 * of 3 launches each).  Every pass is "one sweep": a tile ranks its 4096 pairs locally (warp match-any, stable), publishes its
 * per-digit counts and obtains the counts of all earlier tiles by decoupled look-back per digit; the digit's global start
 * comes from the histogram k_sweep_setup already made.  A pair is read once and written once per pass: 8 B written by pass
 * 0, 8 B read + 8 B written by pass 1.
 *
 * All tiles are handed out by atomic tickets, so a tile only ever waits for tiles that are running or finished: no
 * cooperative launch, no co-residency requirement, any grid size.  Look-back descriptors carry the build's sequence number
 * (tag), so nothing has to be cleared between builds or passes; the few counters are reset by the last block of each launch.
 * Results are bit-identical to k_grid.cu/k_sort.cu and k_gridfused.cu (all three are tested against the oracle).
 */
#include "k_grid_dev.cuh"

#define SW_THREADS 512
#define SW_WARPS (SW_THREADS / 32)
#define SW_ITEMS 8
#define SW_TILE (SW_THREADS * SW_ITEMS) /* pairs per tile */
#define SW_CELL_TILE (SW_THREADS * 8)   /* cells per tile of the offset scan */
#define SW_HSTRIDE (UG_SWEEP_MAXB + 2)

/* descriptor word: tag (28 bits) | flag (2) | value (34).  One 64-bit store publishes all three, so no fence is needed. */
#define SW_AGG 1ull
#define SW_INC 2ull
#define SW_VMASK 0x3ffffffffull
static __device__ __forceinline__ u64 sw_word(u32 tag, u64 flag, u64 value)
{
	return ((u64)(tag & 0xfffffffu) << 36) | (flag << 34) | (value > SW_VMASK ? SW_VMASK : value);
}
static __device__ __forceinline__ bool sw_ready(u64 w, u32 tag) { return (u32)(w >> 36) == (tag & 0xfffffffu) && ((w >> 34) & 3ull) != 0ull; }
static __device__ __forceinline__ bool sw_inclusive(u64 w) { return ((w >> 34) & 3ull) == SW_INC; }

/* single-value look-back by warp 0 (a window of 32 predecessors per step); returns the sum of all earlier tiles */
static __device__ __forceinline__ u64 sw_lookback(volatile u64 *desc, u32 tile, u64 total, u32 tag)
{
	const u32 lane = ug_lane();
	if (tile == 0) {
		if (lane == 0) desc[0] = sw_word(tag, SW_INC, total);
		return 0;
	}
	if (lane == 0) desc[tile] = sw_word(tag, SW_AGG, total);
	u64 prefix = 0;
	int look = (int)tile - 1;
	while (true) {
		const int idx = look - (int)lane;
		u64 w = sw_word(tag, SW_INC, 0); /* before tile 0 */
		if (idx >= 0) {
			do { w = desc[idx]; } while (!sw_ready(w, tag));
		}
		const u32 incl = __ballot_sync(UG_FULL_MASK, sw_inclusive(w));
		u64 part = w & SW_VMASK;
		if (incl) part = ((int)lane <= __ffs(incl) - 1) ? part : 0ull;
		for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(UG_FULL_MASK, part, o);
		prefix += part;
		if (incl) break;
		look -= 32;
	}
	if (lane == 0) desc[tile] = sw_word(tag, SW_INC, prefix + total);
	return prefix;
}

/* look-back of one digit by one thread; SW_WIN predecessors in flight per step.  All tiles of the first wave publish their
 * counts at about the same time, so tile k sums about k of them: the window is what keeps that walk short.  (A variant with
 * one state flag per tile and plain count rows was measured too: equal on configs B and D, 30 % slower on E - two fences and
 * three block barriers per tile cost more than the per-digit descriptors.) */
#ifndef SW_WIN
#define SW_WIN 8
#endif
#ifndef SW_WING
#define SW_WING 4 /* window per lane when several lanes share a digit's look-back (k_sweep_pass) */
#endif
static __device__ __forceinline__ u32 sw_lookback_bin(const volatile u64 *desc, u32 NB, u32 tile, u32 d, u32 tag)
{
	u32 excl = 0;
	int j = (int)tile - 1;
	while (j >= 0) {
		u64 w[SW_WIN];
		const int n = j + 1 < SW_WIN ? j + 1 : SW_WIN;
		bool ok;
		do {
			ok = true;
#pragma unroll
			for (int k = 0; k < SW_WIN; k++)
				if (k < n) { w[k] = desc[(size_t)(j - k) * NB + d]; ok = ok && sw_ready(w[k], tag); }
		} while (!ok);
#pragma unroll
		for (int k = 0; k < SW_WIN; k++)
			if (k < n) {
				excl += (u32)(w[k] & SW_VMASK);
				if (sw_inclusive(w[k])) return excl;
			}
		j -= SW_WIN;
	}
	return excl;
}

struct SweepArgs {
	const ViewConst *vc;
	const int *faces;
	const float *verts;
	int T, C;
	u32 cap;
	float4 *rec;
	uint2 *rect;
	u32 *count, *scan;
	u32 *keys[2], *vals[2];
	u32 *span, *offset, *counters;
	u32 offset_bias; /* added to every offset written (a light-grid shard's slot base) */
	u32 *hist;    /* [2 parities][2 passes][SW_HSTRIDE] difference arrays of the digit histograms */
	u32 *base;    /* [2 passes][UG_SWEEP_MAXB] where the first key of every digit goes */
	u64 *desc_a;  /* look-back descriptors of the count scan  (tiles of SW_THREADS triangles) */
	u64 *desc_b;  /* look-back descriptors of the offset scan (tiles of SW_CELL_TILE cells) */
	u64 *desc;    /* [tile][bins] look-back descriptors of the sort passes */
	u32 *tile_first; /* [pair tiles + 1] the triangle that owns the first pair of every tile of SW_TILE pairs */
	u32 *ctl;     /* [0] build sequence number, [1],[2] ticket / finished blocks of setup, [3],[4] pass 0, [5],[6] pass 1 */
};

/* ------------------------------------------------------------------------------------------------ launch 1 */
__global__ void __launch_bounds__(SW_THREADS) k_sweep_setup(const SweepArgs a)
{
	__shared__ float s_cc[64];
	__shared__ u32 s_warp[SW_WARPS];
	__shared__ u64 s_pref;
	__shared__ u32 s_tile, s_last;
	__shared__ u32 s_d0[UG_SWEEP_MAXB + 1], s_d1[UG_SWEEP_MAXB + 1];
	ug_pdl_enter<UG_PDL_SWEEP>();
	const u32 tid = threadIdx.x;
	const int T = a.T;
	if (tid < 64) s_cc[tid] = a.vc->cc[tid];
	for (u32 i = tid; i <= UG_SWEEP_MAXB; i += SW_THREADS) { s_d0[i] = 0; s_d1[i] = 0; }
	const u32 epoch = a.ctl[0] + 1u; /* this build; the last block publishes it when the launch is over */
	u32 *hist = a.hist + (size_t)(epoch & 1u) * 2 * SW_HSTRIDE;
	u32 *hist_next = a.hist + (size_t)((epoch & 1u) ^ 1u) * 2 * SW_HSTRIDE;
	for (u32 i = blockIdx.x * SW_THREADS + tid; i < (u32)a.C; i += gridDim.x * SW_THREADS) a.span[i] = 0;
	for (u32 i = blockIdx.x * SW_THREADS + tid; i < 2 * SW_HSTRIDE; i += gridDim.x * SW_THREADS) hist_next[i] = 0;
	__syncthreads();
	const GridGeo geo = ug_grid_geo(a.vc, s_cc);
	const u32 tiles = ((u32)T + SW_THREADS - 1) / SW_THREADS;
	while (true) {
		__syncthreads();
		if (tid == 0) s_tile = atomicAdd(&a.ctl[1], 1u);
		__syncthreads();
		const u32 tile = s_tile;
		if (tile >= tiles) break;
		const int t = (int)(tile * SW_THREADS + tid);
		u32 c = 0;
		if (t < T) {
			c = ug_tri_setup(s_cc, geo, a.faces, a.verts, t, a.rec, a.rect);
			a.count[t] = c;
			if (c) { /* sx pairs in each of its cell rows, sy in each of its cell columns */
				const uint2 r = a.rect[t];
				const u32 gx0 = r.x & 0xffffu, gx1 = r.x >> 16, ry0 = (r.y & 0xffffu) - (u32)geo.gy_lo, ry1 = (r.y >> 16) - (u32)geo.gy_lo;
				const u32 sx = gx1 - gx0 + 1u, sy = ry1 - ry0 + 1u;
				atomicAdd(&s_d0[ry0], sx); atomicAdd(&s_d0[ry1 + 1u], 0u - sx);
				atomicAdd(&s_d1[gx0], sy); atomicAdd(&s_d1[gx1 + 1u], 0u - sy);
			}
		}
		u32 total;
		const u32 excl = ug_block_scan<SW_THREADS>(c, s_warp, total);
		if (tid < 32) {
			const u64 p = sw_lookback(a.desc_a, tile, total, epoch);
			if (tid == 0) s_pref = p;
		}
		__syncthreads();
		if (t < T) {
			const u64 pos = s_pref + excl;
			a.scan[t] = (u32)pos;
			if (c && pos + c <= (u64)a.cap) { /* pass 0 finds the triangles of a tile of pairs without searching the scan */
				const u32 m_hi = (u32)((pos + c - 1) / SW_TILE);
				for (u32 m = (u32)((pos + SW_TILE - 1) / SW_TILE); m <= m_hi; m++) a.tile_first[m] = (u32)t;
			}
		}
		if (tile == tiles - 1 && tid == 0) {
			const u64 Ptrue = s_pref + total;
			const bool over = Ptrue > (u64)a.cap;
			a.scan[T] = (u32)(Ptrue > 0xffffffffull ? 0xffffffffull : Ptrue);
			a.counters[0] = over ? 0u : (u32)Ptrue; /* an overflowed frame is rendered again with a larger arena: nothing is emitted */
			a.counters[1] = over ? 1u : 0u;
			a.counters[2] = (u32)(Ptrue > 0xffffffffull ? 0xffffffffull : Ptrue);
		}
	}
	__syncthreads();
	for (u32 i = tid; i <= UG_SWEEP_MAXB; i += SW_THREADS) {
		if (s_d0[i]) atomicAdd(&hist[i], s_d0[i]);
		if (s_d1[i]) atomicAdd(&hist[SW_HSTRIDE + i], s_d1[i]);
	}
	__threadfence();
	__syncthreads();
	if (tid == 0) s_last = (atomicAdd(&a.ctl[2], 1u) == gridDim.x - 1) ? 1u : 0u;
	__syncthreads();
	if (!s_last) return;
	/* last block: every block's histogram contribution is in.  count[d] = running sum of the differences, base[d] = number
	 * of pairs in smaller digits; two bins per thread */
	__threadfence();
	for (int pass = 0; pass < 2; pass++) {
		const u32 v0 = __ldcg(&hist[pass * SW_HSTRIDE + 2 * tid]), v1 = __ldcg(&hist[pass * SW_HSTRIDE + 2 * tid + 1]);
		u32 total;
		const u32 e = ug_block_scan<SW_THREADS>(v0 + v1, s_warp, total);
		const u32 c0 = e + v0, c1 = c0 + v1;
		const u32 e2 = ug_block_scan<SW_THREADS>(c0 + c1, s_warp, total);
		a.base[pass * UG_SWEEP_MAXB + 2 * tid] = e2;
		a.base[pass * UG_SWEEP_MAXB + 2 * tid + 1] = e2 + c0;
	}
	if (tid == 0) { a.ctl[0] = epoch; a.ctl[1] = 0; a.ctl[2] = 0; }
}

/* ------------------------------------------------------------------------------------------------ launches 2 and 3 */
template <int PASS>
__global__ void __launch_bounds__(SW_THREADS, 2) k_sweep_pass(const SweepArgs a)
{
	/* per warp and digit; [NB] = dummy bin of out-of-range lanes.  Pass 0 first uses the same bytes as s_tri[SW_TILE]: the
	 * triangle of every pair of the tile. */
	__shared__ __align__(16) unsigned short s_cnt[SW_WARPS][UG_SWEEP_MAXB + 2];
	__shared__ u32 s_gbase[UG_SWEEP_MAXB];                        /* where the tile's first key of each digit goes */
	__shared__ u32 s_warp[SW_WARPS];
	__shared__ u64 s_pref;
	__shared__ u32 s_tile;
	ug_pdl_enter<UG_PDL_SWEEP>();
	const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u, lt = (1u << lane) - 1u;
	const int T = a.T, C = a.C;
	const u32 epoch = a.ctl[0];
	const u32 tag = epoch * 2u + (u32)PASS;
	const u32 P = a.counters[0];
	const u32 ntiles = (P + SW_TILE - 1) / SW_TILE;
	const u32 noff = PASS == 1 ? ((u32)C + 1u + SW_CELL_TILE - 1) / SW_CELL_TILE : 0u;
	const int gy_lo = a.vc->gy_lo, gy_cnt = a.vc->gy_cnt;
	const u32 NB = PASS == 0 ? (u32)gy_cnt : (u32)a.vc->nbx;
	const u32 *base = a.base + PASS * UG_SWEEP_MAXB;
	u32 KG = 1; /* lanes that share one digit's look-back (a power of two, KG * NB <= threads) */
#ifndef SW_NO_WIDE_LOOKBACK
	while (KG < 32u && NB * KG * 2u <= SW_THREADS) KG <<= 1;
#endif
	u32 *ticket = &a.ctl[3 + 2 * PASS], *finished = &a.ctl[4 + 2 * PASS];
	const u32 *keys_in = a.keys[0], *vals_in = a.vals[0];
	u32 *keys_out = a.keys[PASS == 0 ? 0 : 1], *vals_out = a.vals[PASS == 0 ? 0 : 1];
	while (true) {
		__syncthreads();
		if (tid == 0) s_tile = atomicAdd(ticket, 1u);
		__syncthreads();
		u32 tile = s_tile;
		if (tile >= noff + ntiles) break;
		if (PASS == 1 && tile < noff) { /* offset = exclusive scan of span over ALL cells (frustum_grid.h:361) */
			const u32 b0 = tile * SW_CELL_TILE + tid * 8u;
			u32 v[8], sum = 0;
#pragma unroll
			for (int k = 0; k < 8; k++) { v[k] = (b0 + k < (u32)C) ? a.span[b0 + k] : 0u; sum += v[k]; }
			u32 total;
			u32 run = ug_block_scan<SW_THREADS>(sum, s_warp, total);
			if (tid < 32) {
				const u64 p = sw_lookback(a.desc_b, tile, total, epoch);
				if (tid == 0) s_pref = p;
			}
			__syncthreads();
			run += (u32)s_pref;
#pragma unroll
			for (int k = 0; k < 8; k++) { if (b0 + k <= (u32)C) a.offset[b0 + k] = run + a.offset_bias; run += v[k]; }
			continue;
		}
		tile -= noff;
		const u32 lo = tile * SW_TILE, hi = min(lo + SW_TILE, P);
		u32 key[SW_ITEMS], val[SW_ITEMS], dr[SW_ITEMS]; /* dr = digit << 16 | rank inside (warp, digit) */
		if (PASS == 0) {
			/* Pair i of the emission order belongs to the last triangle t with scan[t] <= i.  Every triangle of the tile that
			 * has pairs marks the slot of its first pair; a running maximum over the slots spreads the ids (they ascend). */
			static_assert(sizeof(s_cnt) >= SW_TILE * sizeof(u32), "s_tri aliases s_cnt");
			u32 *s_tri = (u32 *)&s_cnt[0][0];
#pragma unroll
			for (int k = 0; k < SW_ITEMS; k++) s_tri[k * SW_THREADS + tid] = 0;
			const u32 tlo = a.tile_first[tile], thi = (tile + 1 < ntiles) ? a.tile_first[tile + 1] : (u32)T - 1u;
			__syncthreads();
			for (u32 t = tlo + tid; t <= thi; t += SW_THREADS) {
				if (a.count[t]) {
					const u32 pos = a.scan[t];
					const u32 idx = pos > lo ? pos - lo : 0u;
					if (idx < SW_TILE && (pos >= lo || t == tlo)) s_tri[idx] = t + 1u;
				}
			}
			__syncthreads();
			{ /* running maximum: thread j owns slots [8j, 8j+8) */
				uint4 x0 = ((uint4 *)s_tri)[2 * tid], x1 = ((uint4 *)s_tri)[2 * tid + 1];
				x0.y = max(x0.y, x0.x); x0.z = max(x0.z, x0.y); x0.w = max(x0.w, x0.z);
				x1.x = max(x1.x, x0.w); x1.y = max(x1.y, x1.x); x1.z = max(x1.z, x1.y); x1.w = max(x1.w, x1.z);
				u32 m = x1.w;
				for (int o = 1; o < 32; o <<= 1) {
					const u32 y = __shfl_up_sync(UG_FULL_MASK, m, o);
					if ((int)lane >= o) m = max(m, y);
				}
				if (lane == 31) s_warp[warp] = m;
				__syncthreads();
				u32 before = __shfl_up_sync(UG_FULL_MASK, m, 1);
				if (lane == 0) before = 0;
				for (u32 w = 0; w < warp; w++) before = max(before, s_warp[w]);
				x0.x = max(x0.x, before); x0.y = max(x0.y, before); x0.z = max(x0.z, before); x0.w = max(x0.w, before);
				x1.x = max(x1.x, before); x1.y = max(x1.y, before); x1.z = max(x1.z, before); x1.w = max(x1.w, before);
				((uint4 *)s_tri)[2 * tid] = x0; ((uint4 *)s_tri)[2 * tid + 1] = x1;
			}
			__syncthreads();
#pragma unroll
			for (int r = 0; r < SW_ITEMS; r++) val[r] = s_tri[warp * (32 * SW_ITEMS) + r * 32 + lane] - 1u;
			__syncthreads(); /* s_tri is dead: the same bytes become the counters */
		}
		for (u32 i = tid; i < SW_WARPS * (UG_SWEEP_MAXB + 2) / 2; i += SW_THREADS) ((u32 *)&s_cnt[0][0])[i] = 0;
#pragma unroll
		for (int r = 0; r < SW_ITEMS; r++) { /* warp w owns pairs [lo + w*32*ITEMS, +32*ITEMS): round r covers 32 consecutive pairs */
			const u32 i = lo + warp * (32 * SW_ITEMS) + r * 32 + lane;
			u32 d = NB;
			key[r] = 0;
			if (i < hi) {
				if (PASS == 0) { /* cell of pair i from its triangle's rectangle; grid_kernel.cu:318-325: i over x, j over y */
					const u32 t = val[r];
					const uint2 rc = a.rect[t];
					const u32 gx0 = rc.x & 0xffffu, ry0 = rc.y & 0xffffu, sy = (rc.y >> 16) - ry0 + 1u;
					const u32 local = i - a.scan[t];
					const u32 ix = local / sy, iy = local - ix * sy;
					d = ry0 - (u32)gy_lo + iy;
					key[r] = (gx0 + ix) * (u32)gy_cnt + d;
					atomicAdd(&a.span[key[r]], 1u);
				} else {
					key[r] = keys_in[i];
					val[r] = vals_in[i];
					d = key[r] / (u32)gy_cnt;
				}
			} else if (PASS == 1) {
				val[r] = 0;
			}
			dr[r] = d;
		}
		__syncthreads();
#pragma unroll
		for (int r = 0; r < SW_ITEMS; r++) { /* stable rank of every pair among the tile's pairs of the same digit */
			const u32 d = dr[r];
			const u32 peers = __match_any_sync(UG_FULL_MASK, d);
			const u32 prev = s_cnt[warp][d];
			__syncwarp();
			if ((peers & lt) == 0) s_cnt[warp][d] = (unsigned short)(prev + __popc(peers));
			__syncwarp();
			dr[r] = (d << 16) | (prev + __popc(peers & lt));
		}
		__syncthreads();
		if (KG == 1) {
			for (u32 d = tid; d < NB; d += SW_THREADS) { /* per-warp counts -> per-warp starts; the tile's count; all earlier tiles' */
				u32 run = 0;
#pragma unroll
				for (int w = 0; w < SW_WARPS; w++) {
					const u32 c = s_cnt[w][d];
					s_cnt[w][d] = (unsigned short)run;
					run += c;
				}
				volatile u64 *mine = a.desc + (size_t)tile * NB + d;
				u32 excl = 0;
				if (tile == 0) {
					*mine = sw_word(tag, SW_INC, run);
				} else {
					*mine = sw_word(tag, SW_AGG, run);
					excl = sw_lookback_bin(a.desc, NB, tile, d, tag);
					*mine = sw_word(tag, SW_INC, (u64)excl + run);
				}
				s_gbase[d] = base[d] + excl;
			}
		} else {
			/* Fewer digits than threads: KG consecutive lanes share a digit, lane q of the group looks at predecessors
			 * [j - q*SW_WING - SW_WING + 1, j - q*SW_WING] (a narrower window per lane than the one-lane walk: registers), so one round
			 * covers KG * SW_WING of them.  The sum ends at the nearest
			 * descriptor that already carries an inclusive prefix (or at tile 0). */
			const u32 d = tid / KG, q = tid & (KG - 1u);
			const bool act = d < NB;
			const volatile u64 *vdesc = a.desc; /* polled: every read must reach memory */
			volatile u64 *mine = a.desc + (size_t)tile * NB + (act ? d : 0u);
			u32 run = 0, excl = 0;
			if (act && q == 0) {
#pragma unroll
				for (int w = 0; w < SW_WARPS; w++) {
					const u32 c = s_cnt[w][d];
					s_cnt[w][d] = (unsigned short)run;
					run += c;
				}
				*mine = sw_word(tag, tile == 0 ? SW_INC : SW_AGG, run);
			}
			if (tile > 0) {
				int j = (int)tile - 1 - (int)(q * SW_WING);
				bool done = !act;
				while (!__all_sync(UG_FULL_MASK, done)) {
					u32 part = 0;
					bool found = false;
					if (!done) {
						if (j < 0) found = true; /* nothing before tile 0 */
						else {
							u64 w[SW_WING];
							const int n = j + 1 < SW_WING ? j + 1 : SW_WING;
							bool ok;
							do {
								ok = true;
#pragma unroll
								for (int k = 0; k < SW_WING; k++)
									if (k < n) { w[k] = vdesc[(size_t)(j - k) * NB + d]; ok = ok && sw_ready(w[k], tag); }
							} while (!ok);
#pragma unroll
							for (int k = 0; k < SW_WING; k++)
								if (k < n && !found) {
									part += (u32)(w[k] & SW_VMASK);
									found = sw_inclusive(w[k]);
								}
							if (n < SW_WING) found = true; /* the window reached tile 0 */
						}
					}
					const u32 fm = __ballot_sync(UG_FULL_MASK, found && !done);
					const u32 gm = (fm >> (lane & ~(KG - 1u))) & (KG == 32u ? 0xffffffffu : ((1u << KG) - 1u)); /* my group's lanes */
					const u32 qf = gm ? (u32)__ffs(gm) - 1u : KG; /* the nearest window that holds the end of the sum */
					u32 contrib = (!done && q <= qf) ? part : 0u;
					for (u32 o = KG >> 1; o > 0; o >>= 1) contrib += __shfl_xor_sync(UG_FULL_MASK, contrib, o);
					if (!done) {
						excl += contrib;
						if (gm) done = true;
						else j -= (int)(KG * SW_WING);
					}
				}
			}
			if (act && q == 0) {
				if (tile > 0) *mine = sw_word(tag, SW_INC, (u64)excl + run);
				s_gbase[d] = base[d] + excl;
			}
		}
		__syncthreads();
#pragma unroll
		for (int r = 0; r < SW_ITEMS; r++) {
			const u32 i = lo + warp * (32 * SW_ITEMS) + r * 32 + lane;
			if (i < hi) {
				const u32 d = dr[r] >> 16;
				const u32 pos = s_gbase[d] + s_cnt[warp][d] + (dr[r] & 0xffffu);
				if (pos < a.cap) { keys_out[pos] = key[r]; vals_out[pos] = val[r]; }
			}
		}
	}
	__syncthreads();
	if (tid == 0 && atomicAdd(finished, 1u) == gridDim.x - 1) { *ticket = 0; *finished = 0; }
}

/* host side ------------------------------------------------------------------------------------------ */
size_t ugk_sweep_desc_a(int T) { return (size_t)(T + SW_THREADS - 1) / SW_THREADS + 2; }
size_t ugk_sweep_desc_b(int C) { return ((size_t)C + 1 + SW_CELL_TILE - 1) / SW_CELL_TILE + 2; }
size_t ugk_sweep_desc(size_t cap, int bins) { return ((cap + SW_TILE - 1) / SW_TILE + 1) * (size_t)bins; }
size_t ugk_sweep_hist_words() { return (size_t)4 * SW_HSTRIDE; }
size_t ugk_sweep_tiles(size_t cap) { return (cap + SW_TILE - 1) / SW_TILE + 2; }

/* pass 0 writes buffer pair 0 from nothing, pass 1 sorts 0 -> 1: the ordered list ends in pair 1.
 * part: 1 = the per-triangle launch only, 2 = the two sort passes only, 3 = the whole build.  (A strip's light grid is not
 * needed before its shadow rays: its sort passes can be held back until the camera grid is done.) */
int ugk_grid_sweep(cudaStream_t s, const ViewConst *vc, const int *faces, const float *verts, GridBuffers &g, int sms, int *launches, int part)
{
	SweepArgs a;
	a.vc = vc; a.faces = faces; a.verts = verts;
	a.T = g.T; a.C = g.C; a.cap = (u32)g.cap;
	a.rec = g.tri_rec; a.rect = g.rect; a.count = g.count; a.scan = g.scan;
	a.keys[0] = g.keys[0]; a.keys[1] = g.keys[1]; a.vals[0] = g.vals[0]; a.vals[1] = g.vals[1];
	a.span = g.span; a.offset = g.offset; a.counters = g.counters; a.offset_bias = g.offset_bias;
	a.hist = g.sweep_hist; a.base = g.sweep_base; a.desc_a = g.sweep_desc_a; a.desc_b = g.sweep_desc_b; a.desc = g.sweep_desc; a.ctl = g.sweep_ctl;
	a.tile_first = g.sweep_tile_first;
	const unsigned tiles_a = (unsigned)((g.T + SW_THREADS - 1) / SW_THREADS);
	const unsigned max_blocks = (unsigned)sms * 2u;
	if (part & 1) {
		ug_launch(UG_PDL_SWEEP, k_sweep_setup, dim3(tiles_a < max_blocks ? (tiles_a ? tiles_a : 1u) : max_blocks), dim3(SW_THREADS), s, a);
		if (launches) *launches += 1;
	}
	if (part & 2) {
		const size_t tiles_p = (g.cap + SW_TILE - 1) / SW_TILE;
		const size_t noff = ((size_t)g.C + 1 + SW_CELL_TILE - 1) / SW_CELL_TILE;
		ug_launch(UG_PDL_SWEEP, k_sweep_pass<0>, dim3((unsigned)(tiles_p < max_blocks ? (tiles_p ? tiles_p : 1) : max_blocks)), dim3(SW_THREADS), s, a);
		ug_launch(UG_PDL_SWEEP, k_sweep_pass<1>, dim3((unsigned)(tiles_p + noff < max_blocks ? tiles_p + noff : max_blocks)), dim3(SW_THREADS), s, a);
		if (launches) *launches += 2;
	}
	return 1;
}
