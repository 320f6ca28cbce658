/* k_gridfused.cu -- the whole perspective-grid build of one view (FrustumGrid::buildGrid,
 * frustum_grid.h:210-366) as ONE persistent cooperative kernel.
 *
 * The build moves a few MB (config B: 48 B x 80 000 triangles in, 0.35 M pairs) - microseconds of HBM
 * time - so the multi-kernel chain of k_grid.cu/k_sort.cu (11 launches + 4 memsets per view) is bound by
 * launch and dependency latency, not by bandwidth.  Here one block per SM stays resident and the phases
 * are separated by grid-wide barriers (cooperative launch guarantees co-residency, so blocks may wait for
 * one another):
 *
 *   phase 0  per triangle: record + rectangle + pair count (a2/a3), the exclusive scan of the counts
 *            in the same pass (decoupled look-back between tiles, a4), zero span[] and the digit totals
 *   phase 1  per output pair: key/value emission in triangle-major order (a5) + span histogram (a7)
 *            + the first radix pass's digit histogram of the block's own chunk
 *   phase 2  exclusive scan of span over ALL cells -> offset (a7); stable scatter of radix pass 0
 *   phase 3+ (per further 8-bit pass) chunk histogram | barrier | scatter                      (a6)
 *
 * Every scratch word is reset by a later phase of the same launch, so the kernel is replayable inside
 * a CUDA graph without memset nodes.  Results are bit-identical to the multi-kernel path (both are
 * tested against the oracle).
 */
#include "k_grid_dev.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define GF_THREADS 512
#define GF_ROUNDS 4   /* radix sub-tile = 2048 keys */

struct GridFusedArgs {
	const ViewConst *vc;
	const int *faces;
	const float *verts;
	int T, C, passes;
	u32 cap;
	float4 *rec;
	uint2 *rect;
	u32 *count, *scan;
	u32 *keys[2], *vals[2];
	u32 *span, *offset, *counters;
	u32 *hist;      /* [G][256] block digit histograms of the current pass */
	u32 *tot;       /* [4][256] digit totals per pass */
	u64 *state_a;   /* look-back descriptors of the count scan  (tiles of GF_THREADS triangles) */
	u64 *state_b;   /* look-back descriptors of the offset scan (tiles of GF_THREADS*4 cells) */
};

__global__ void __launch_bounds__(GF_THREADS, 2) k_grid_fused(const GridFusedArgs a)
{
	cg::grid_group grid = cg::this_grid();
	__shared__ float s_cc[64];
	__shared__ u32 s_warp[GF_THREADS / 32];
	__shared__ u32 s_pref;
	__shared__ int s_range[2];
	__shared__ u32 s_hist[257];
	__shared__ u32 s_cnt[GF_THREADS / 32][257];
	__shared__ u32 s_run[256];
	const u32 G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
	const int T = a.T, C = a.C;
	if (tid < 64) s_cc[tid] = a.vc->cc[tid];
	const int gy_lo = a.vc->gy_lo, gy_cnt = a.vc->gy_cnt;
	__syncthreads();
	const GridGeo geo = ug_grid_geo(a.vc, s_cc);

	gf_stamp(a.counters, 0);
	/* ---------------- phase 0: triangles -> record, rectangle, count; scan of the counts */
	for (u32 i = b * GF_THREADS + tid; i < (u32)C; i += G * GF_THREADS) a.span[i] = 0;
	for (u32 i = b * GF_THREADS + tid; i < 4 * 256; i += G * GF_THREADS) a.tot[i] = 0;
	{
		const u32 tiles_b = ((u32)C + 1 + GF_THREADS * 4 - 1) / (GF_THREADS * 4);
		for (u32 i = b * GF_THREADS + tid; i < tiles_b + 1; i += G * GF_THREADS) a.state_b[i] = 0; /* used in phase 2 */
	}
	const u32 tiles_a = ((u32)T + GF_THREADS - 1) / GF_THREADS;
	for (u32 tile = b; tile < tiles_a; tile += G) {
		const int t = (int)(tile * GF_THREADS + tid);
		u32 c = 0;
		if (t < T) {
			c = ug_tri_setup(s_cc, geo, a.faces, a.verts, t, a.rec, a.rect);
			a.count[t] = c;
		}
		u32 total;
		const u32 excl = ug_block_scan<GF_THREADS>(c, s_warp, total);
		if (tid < 32) {
			const u32 p = ug_lookback(a.state_a, tile, total);
			if (tid == 0) s_pref = p;
		}
		__syncthreads();
		if (t < T) a.scan[t] = s_pref + excl;
		if (tile == tiles_a - 1 && tid == 0) {
			const u64 P64 = (u64)s_pref + total; /* saturates: see ug_lookback */
			const u32 Ptrue = P64 > 0xffffffffull ? 0xffffffffu : (u32)P64, Pc = min(Ptrue, a.cap);
			a.scan[T] = Ptrue;
			a.counters[0] = Pc;
			a.counters[1] = (Ptrue > a.cap) ? 1u : 0u; /* arena overflow: the host grows it and re-runs the frame */
			a.counters[2] = Ptrue;
		}
		__syncthreads();
	}
	gf_stamp(a.counters, 1);
	grid.sync();
	gf_stamp(a.counters, 2);

	/* ---------------- phase 1: emit the block's chunk of pairs (+ span histogram, digit histogram of pass 0) */
	for (u32 i = b * GF_THREADS + tid; i < tiles_a + 1; i += G * GF_THREADS) a.state_a[i] = 0; /* phase 0 is over: ready for the next launch */
	const u32 P = a.counters[0];
	u32 lo, hi;
	rs_chunk(P, G, b, lo, hi);
	if (tid < 257) s_hist[tid] = 0;
	if (tid < 2 && lo < hi) s_range[tid] = ug_upper_bound(a.scan, 0, T + 1, tid == 0 ? lo : hi - 1) - 1;
	__syncthreads();
	{
		const int tlo = s_range[0], thi = s_range[1];
		u32 *keys = a.keys[0], *vals = a.vals[0];
		for (u32 base = lo; base < hi; base += GF_THREADS) {
			const u32 p = base + tid;
			const bool valid = p < hi;
			u32 d = 256u;
			if (valid) {
				const int t = ug_upper_bound(a.scan, tlo, thi + 1, p) - 1;
				const u32 key = ug_pair_key(a.rect[t], p - a.scan[t], gy_lo, gy_cnt);
				keys[p] = key;
				vals[p] = (u32)t;
				atomicAdd(&a.span[key], 1u);
				d = key & 255u;
			}
			const u32 peers = __match_any_sync(UG_FULL_MASK, d);
			if (valid && (peers & ((1u << ug_lane()) - 1u)) == 0) atomicAdd(&s_hist[d], __popc(peers));
		}
	}
	__syncthreads();
	if (tid < 256) {
		a.hist[(size_t)b * 256 + tid] = s_hist[tid];
		if (s_hist[tid]) atomicAdd(&a.tot[tid], s_hist[tid]);
	}
	gf_stamp(a.counters, 3);
	grid.sync();
	gf_stamp(a.counters, 4);

	/* ---------------- phase 2: offset = exclusive scan of span over ALL cells (frustum_grid.h:361) */
	{
		const u32 per = GF_THREADS * 4, tiles_b = ((u32)C + 1 + per - 1) / per;
		for (u32 tile = b; tile < tiles_b; tile += G) {
			const u32 base = tile * per + tid * 4;
			u32 v[4], sum = 0;
#pragma unroll
			for (int k = 0; k < 4; k++) { v[k] = (base + k < (u32)C) ? a.span[base + k] : 0u; sum += v[k]; }
			u32 total;
			u32 run = ug_block_scan<GF_THREADS>(sum, s_warp, total);
			if (tid < 32) {
				const u32 p = ug_lookback(a.state_b, tile, total);
				if (tid == 0) s_pref = p;
			}
			__syncthreads();
			run += s_pref;
#pragma unroll
			for (int k = 0; k < 4; k++) { if (base + k <= (u32)C) a.offset[base + k] = run; run += v[k]; }
			__syncthreads();
		}
	}

	gf_stamp(a.counters, 5);
	/* ---------------- radix passes: scatter pass 0, then (histogram | barrier | scatter) per further pass */
	int cur = 0;
	for (int pass = 0; pass < a.passes; pass++) {
		const int shift = 8 * pass;
		if (pass > 0) {
			if (tid < 257) s_hist[tid] = 0;
			__syncthreads();
			rs_block_hist<GF_THREADS>(a.keys[cur], lo, hi, shift, s_hist);
			__syncthreads();
			if (tid < 256) {
				a.hist[(size_t)b * 256 + tid] = s_hist[tid];
				if (s_hist[tid]) atomicAdd(&a.tot[pass * 256 + tid], s_hist[tid]);
			}
			gf_stamp(a.counters, 8);
			grid.sync();
			gf_stamp(a.counters, 9);
		}
		{ /* where this block's keys of digit `tid` go: all smaller digits + the same digit in earlier blocks */
			u32 total;
			const u32 digit_start = ug_block_scan<GF_THREADS>(tid < 256 ? a.tot[pass * 256 + tid] : 0u, s_warp, total);
			if (tid < 256) {
				u32 before = 0;
#pragma unroll 8
				for (u32 bb = 0; bb < b; bb++) before += a.hist[(size_t)bb * 256 + tid];
				s_run[tid] = digit_start + before;
			}
		}
		__syncthreads();
		gf_stamp(a.counters, pass == 0 ? 6 : 10);
		rs_block_scatter<GF_THREADS, GF_ROUNDS>(a.keys[cur], a.vals[cur], a.keys[cur ^ 1], a.vals[cur ^ 1], lo, hi, shift, s_cnt, s_run);
		cur ^= 1;
		gf_stamp(a.counters, pass == 0 ? 7 : 11);
		if (pass + 1 < a.passes) grid.sync(); /* the next pass reads what every block scattered (and overwrites hist) */
	}
}

/* host side ------------------------------------------------------------------------------------------ */
static int g_fused_blocks_per_sm = -1;

bool ugk_grid_fused_supported(int device)
{
	int coop = 0;
	cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
	if (!coop) return false;
	if (g_fused_blocks_per_sm < 0) {
		int n = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_grid_fused, GF_THREADS, 0) != cudaSuccess) n = 0;
		g_fused_blocks_per_sm = n;
	}
	return g_fused_blocks_per_sm >= 1;
}

size_t ugk_grid_fused_state_a(int T) { return (size_t)(T + GF_THREADS - 1) / GF_THREADS + 2; }
size_t ugk_grid_fused_state_b(int C) { return ((size_t)C + 1 + GF_THREADS * 4 - 1) / (GF_THREADS * 4) + 2; }
size_t ugk_grid_fused_hist(int sms) { return (size_t)sms * 256; }

/* returns the index of the buffer pair holding the sorted list, or -1 if the launch failed */
int ugk_grid_fused(cudaStream_t s, const ViewConst *vc, const int *faces, const float *verts, GridBuffers &g, int sms)
{
	GridFusedArgs a;
	a.vc = vc; a.faces = faces; a.verts = verts;
	a.T = g.T; a.C = g.C; a.passes = (g.key_bits + 7) / 8; a.cap = (u32)g.cap;
	a.rec = g.tri_rec; a.rect = g.rect; a.count = g.count; a.scan = g.scan;
	a.keys[0] = g.keys[0]; a.keys[1] = g.keys[1]; a.vals[0] = g.vals[0]; a.vals[1] = g.vals[1];
	a.span = g.span; a.offset = g.offset; a.counters = g.counters;
	a.hist = g.fused_hist; a.tot = g.fused_tot; a.state_a = g.fused_state_a; a.state_b = g.fused_state_b;
	void *args[] = {(void *)&a};
	const cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_grid_fused, dim3((unsigned)sms), dim3(GF_THREADS), args, 0, s);
	if (e != cudaSuccess) return -1;
	return a.passes & 1;
}
