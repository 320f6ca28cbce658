/* k_ugridfused.cu -- the world-space uniform grid of row R built by ONE persistent cooperative kernel.
 *
 * As a chain (k_reflect.cu: box reduce, resolution, count, scan, emit, 3 radix passes, scan, records, cells = 17
 * launches) the build moves a few MB but takes 120 us alone and ~290 us inside the frame, where its short kernels
 * queue behind the persistent traversal kernels of the other branches (measured with UG_TIMELINE) - and the
 * reflection rays wait for it.  Here one block per SM stays resident and grid-wide barriers separate the phases:
 *
 *   phase 0  bounding box of the vertices (block reduce + ordered-uint atomics); zero span[] and the fill cursors
 *   phase 1  every block derives the same resolution from the box; per triangle (v0,e1,e2), grown box, cell box,
 *            pair count, and the exclusive scan of the counts in the same pass (decoupled look-back)
 *   phase 2  per pair: cell id, span histogram (the pair's cell and triangle are kept for phase 4)
 *   phase 3  offset = exclusive scan of span over the cells (decoupled look-back)
 *   phase 4  per pair: slot = offset[cell] + cursor[cell]++, write the 16-byte pair record there
 *   phase 5  per cell: (offset, span, union box)
 *
 * No sort: the order of a cell's list does not matter to the DDA (nearest hit, ties to the lower id, is order
 * independent), so a counting fill replaces the three radix passes of the chain.  The list order inside a cell is
 * therefore not deterministic; the hits are.  Every scratch word is reset inside the launch (graph-replayable).
 */
#include "k_ugrid_dev.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define UF_THREADS 512

struct UGridFusedArgs {
	const float *verts;
	const int *faces;
	int V, T, max_cells;
	float cells_per_tri;
	u32 cap;
	u32 *box;        /* [6] ordered-uint bounding box, reset by the launch itself */
	UGridConst *uc;  /* out: what the traversal reads */
	float4 *geo, *tbox;
	uint4 *rect;
	u32 *scan;       /* [T+1] */
	u32 *pair_cell, *pair_tri, *cursor; /* scratch: [cap], [cap], [max_cells] */
	u32 *span, *offset, *counters;
	uint4 *prec, *cellrec;
	u64 *state_a, *state_b;
};

__global__ void __launch_bounds__(UF_THREADS, 1) k_ugrid_fused(const UGridFusedArgs a)
{
	cg::grid_group grid = cg::this_grid();
	__shared__ u32 s_warp[UF_THREADS / 32];
	__shared__ u32 s_pref;
	__shared__ int s_range[2];
	__shared__ UGridConst s_u;
	const u32 G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
	const int T = a.T;

	/* ---------------- phase 0: bounding box; zero the per-cell counters of the whole arena */
	{
		float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
		for (int v = (int)(b * UF_THREADS + tid); v < a.V; v += (int)(G * UF_THREADS))
			for (int k = 0; k < 3; k++) { const float x = a.verts[3 * (size_t)v + k]; lo[k] = fminf(lo[k], x); hi[k] = fmaxf(hi[k], x); }
		for (int k = 0; k < 3; k++) {
			for (int o = 16; o > 0; o >>= 1) {
				lo[k] = fminf(lo[k], __shfl_xor_sync(UG_FULL_MASK, lo[k], o));
				hi[k] = fmaxf(hi[k], __shfl_xor_sync(UG_FULL_MASK, hi[k], o));
			}
			if (ug_lane() == 0) { atomicMin(&a.box[k], f2ord(lo[k])); atomicMax(&a.box[3 + k], f2ord(hi[k])); }
		}
		for (u32 i = b * UF_THREADS + tid; i < (u32)a.max_cells; i += G * UF_THREADS) { a.span[i] = 0; a.cursor[i] = 0; }
		const u32 tiles_b = ((u32)a.max_cells + 1 + UF_THREADS * 4 - 1) / (UF_THREADS * 4);
		for (u32 i = b * UF_THREADS + tid; i < tiles_b + 1; i += G * UF_THREADS) a.state_b[i] = 0; /* used in phase 3 */
	}
	grid.sync();

	/* ---------------- phase 1: resolution; triangles -> geometry, boxes, counts; scan of the counts */
	if (tid == 0) {
		ugrid_resolution(a.box, T, a.max_cells, a.cells_per_tri, &s_u); /* same inputs, same result in every block */
		if (b == 0) *a.uc = s_u;
	}
	__syncthreads();
	const UGridConst u = s_u;
	const u32 tiles_a = ((u32)T + UF_THREADS - 1) / UF_THREADS;
	for (u32 tile = b; tile < tiles_a; tile += G) {
		const int t = (int)(tile * UF_THREADS + tid);
		u32 c = 0;
		if (t < T) c = ugrid_tri_setup(u, a.faces, a.verts, t, a.geo, a.tbox, a.rect);
		u32 total;
		const u32 excl = ug_block_scan<UF_THREADS>(c, s_warp, total);
		if (tid < 32) {
			const u32 p = ug_lookback(a.state_a, tile, total);
			if (tid == 0) s_pref = p;
		}
		__syncthreads();
		if (t < T) a.scan[t] = s_pref + excl;
		if (tile == tiles_a - 1 && tid == 0) {
			const u32 Ptrue = s_pref + total, Pc = min(Ptrue, a.cap);
			a.scan[T] = Ptrue;
			a.counters[0] = Pc;
			a.counters[1] = (Ptrue > a.cap) ? 1u : 0u; /* arena overflow: the host grows it and re-runs the frame */
			a.counters[2] = Ptrue;
		}
		__syncthreads();
	}
	grid.sync();

	/* ---------------- phase 2: pairs -> cell ids + span histogram */
	if (b == 0 && tid < 6) a.box[tid] = (tid < 3) ? 0xffffffffu : 0u;                              /* every block has read the box */
	for (u32 i = b * UF_THREADS + tid; i < tiles_a + 1; i += G * UF_THREADS) a.state_a[i] = 0;    /* phase 1 is over */
	const u32 P = a.counters[0];
	u32 lo, hi;
	rs_chunk(P, G, b, lo, hi);
	if (tid < 2 && lo < hi) s_range[tid] = ug_upper_bound(a.scan, 0, T + 1, tid == 0 ? lo : hi - 1) - 1;
	__syncthreads();
	{
		const int tlo = s_range[0], thi = s_range[1];
		for (u32 p = lo + tid; p < hi; p += UF_THREADS) {
			const int t = ug_upper_bound(a.scan, tlo, thi + 1, p) - 1;
			const u32 key = ugrid_pair_key(a.rect[t], p - a.scan[t], u.n[1], u.n[2]);
			a.pair_cell[p] = key;
			a.pair_tri[p] = (u32)t;
			atomicAdd(&a.span[key], 1u);
		}
	}
	grid.sync();

	/* ---------------- phase 3: offset = exclusive scan of span over the cells */
	const u32 C = (u32)u.cells;
	{
		const u32 per = UF_THREADS * 4, tiles_b = (C + 1 + per - 1) / per;
		for (u32 tile = b; tile < tiles_b; tile += G) {
			const u32 base = tile * per + tid * 4;
			u32 v[4], sum = 0;
#pragma unroll
			for (int k = 0; k < 4; k++) { v[k] = (base + k < C) ? a.span[base + k] : 0u; sum += v[k]; }
			u32 total;
			u32 run = ug_block_scan<UF_THREADS>(sum, s_warp, total);
			if (tid < 32) {
				const u32 p = ug_lookback(a.state_b, tile, total);
				if (tid == 0) s_pref = p;
			}
			__syncthreads();
			run += s_pref;
#pragma unroll
			for (int k = 0; k < 4; k++) { if (base + k <= C) a.offset[base + k] = run; run += v[k]; }
			__syncthreads();
		}
	}
	grid.sync();

	/* ---------------- phase 4: counting fill of the pair records */
	for (u32 p = lo + tid; p < hi; p += UF_THREADS) {
		const u32 key = a.pair_cell[p], tri = a.pair_tri[p];
		const u32 slot = a.offset[key] + atomicAdd(&a.cursor[key], 1u);
		a.prec[slot] = ugrid_pair_record(u, key, tri, a.tbox);
	}
	grid.sync();

	/* ---------------- phase 5: cell records */
	for (u32 c = b * UF_THREADS + tid; c < C; c += G * UF_THREADS) a.cellrec[c] = ugrid_cell_record(a.offset[c], a.span[c], a.prec);
}

/* host side ------------------------------------------------------------------------------------------ */
static int g_ugrid_fused_ok = -1;

bool ugk_ugrid_fused_supported(int device)
{
	if (g_ugrid_fused_ok < 0) {
		int coop = 0, n = 0;
		cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
		if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_ugrid_fused, UF_THREADS, 0) != cudaSuccess) n = 0;
		g_ugrid_fused_ok = (coop && n >= 1) ? 1 : 0;
	}
	return g_ugrid_fused_ok == 1;
}

size_t ugk_ugrid_fused_state_a(int T) { return (size_t)(T + UF_THREADS - 1) / UF_THREADS + 2; }
size_t ugk_ugrid_fused_state_b(int C) { return ((size_t)C + 1 + UF_THREADS * 4 - 1) / (UF_THREADS * 4) + 2; }

/* false if the cooperative launch was refused (the caller falls back to the kernel chain) */
bool ugk_ugrid_fused(cudaStream_t s, ug_context *ctx)
{
	ReflectState &rf = ctx->refl;
	GridBuffers &g = rf.grid;
	UGridFusedArgs a;
	a.verts = ctx->d_vertexlist; a.faces = ctx->d_facelist; a.V = ctx->num_vertices; a.T = ctx->num_faces; a.max_cells = g.C;
	a.cells_per_tri = ctx->refl_cells_per_tri; a.cap = (u32)g.cap;
	a.box = rf.box; a.uc = rf.uc; a.geo = rf.geo; a.tbox = rf.tbox; a.rect = rf.rect; a.scan = g.scan;
	a.pair_cell = g.keys[0]; a.pair_tri = g.vals[0]; a.cursor = g.keys[1]; /* the sort buffers of the chain double as scratch (cap >= max_cells) */
	a.span = g.span; a.offset = g.offset; a.counters = g.counters; a.prec = rf.prec; a.cellrec = rf.cellrec;
	a.state_a = g.fused_state_a; a.state_b = g.fused_state_b;
	void *args[] = {(void *)&a};
	return cudaLaunchCooperativeKernel((const void *)k_ugrid_fused, dim3((unsigned)ctx->num_sms), dim3(UF_THREADS), args, 0, s) == cudaSuccess;
}
