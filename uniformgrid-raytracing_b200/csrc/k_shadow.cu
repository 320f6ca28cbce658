/* k_shadow.cu -- shadow rays through a light-space perspective grid (a13, a16-a18).
 *
 *   k_regroup  one persistent kernel (cooperative launch, grid barriers between phases; the same code can
 *              also be launched phase by phase when cooperative launch is unavailable):
 *     phase A  misc_kernel.cu:63-129 mapSortkernel: per pixel, the light cell its shadow ray falls into
 *              (sentinel NBX*NBY = outside the light frustum) + histogram of rays per light cell
 *     phase B  ray ranges per light cell (exclusive scan) and the WORK-ITEM table sizes (second scan)
 *     phase C  DecisionData::sort_by_block_indices ... stream_compact (decision_data.h:171-271): regroup
 *              the rays by light cell.  The shadow bit is an OR over the cell's triangles, so the order
 *              of rays inside a cell cannot change any output and a counting sort (the histogram's own atomics
 *              hand out the ranks) replaces the reference's 15-bit radix sort + segmented scan + compaction.
 *              Also writes the item table.
 *   k_shadow   light_kernel.cu:52-270 mod_light_rckernel: any-hit traversal, one warp per work item.
 *
 * Work items.  The reference gives a light cell's rays to CTAs in chunks of 64 and every chunk walks the
 * whole list, so its time is set by the cell with the longest list (border cells collect every clamped
 * triangle).  Here an item is (cell, sub-range of <= rpi rays[, segment of the list]) where rpi shrinks as
 * the list grows (32 rays for <= 64 triangles, 16 for <= 128, ... 1 beyond 1024), and the warp is used as
 * an m x g grid (m rays x g = 32/m triangle groups): every lane tests every g-th staged triangle for its
 * ray and the hits are OR-ed with a ballot.  No lane ever owns more than ~64 sequential tests, short
 * cells with few rays still fill the warp, and a ray still stops at its first occluder.
 */
#include "k_shadow_dev.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define RG_THREADS 512
#ifndef RG_KA
#define RG_KA 2 /* pixels per thread and trip, ray -> cell map */
#endif
#ifndef RG_KC
#define RG_KC 4 /* pixels per thread and trip, scatter */
#endif
#define SH_WARPS 4
#define SH_SEGMAX 2048 /* lists longer than this are also cut along the triangle axis ... */
#define SH_SEGBUDGET 16 /* ... but only while the cell has few rays: at most 1 + SH_SEGBUDGET/rays segments, so a cell never needs more than
                           rays + SH_SEGBUDGET items and the table is bounded by N + (SH_SEGBUDGET + 1)*C (ugk_shadow_item_cap) */

/* rays per item for a list of `ntri` triangles */
static __host__ __device__ __forceinline__ u32 sh_rays_per_item(u32 ntri)
{
	if (ntri <= 64) return 32;
	if (ntri <= 128) return 16;
	if (ntri <= 256) return 8;
	if (ntri <= 512) return 4;
	if (ntri <= 1024) return 2;
	return 1;
}
/* segments of a list: many rays already give the cell enough items; few rays on a long list are what needs cutting */
static __host__ __device__ __forceinline__ u32 sh_segments(u32 rays, u32 ntri)
{
	if (ntri <= SH_SEGMAX) return 1;
	return min((ntri + SH_SEGMAX - 1) / SH_SEGMAX, 1u + SH_SEGBUDGET / rays);
}
/* The shape of a cell's items.  Few rays on a long list: few rays per item, the lanes split every staged batch (above).
 * A warp's worth of rays or more on a long list (crowded cells - the spherical light grid puts tens of thousands of rays
 * and thousands of triangles into one): 32 rays per item, one per lane, and the LIST is cut into segments of about
 * SH_SEG triangles, so every staged batch is tested by 32 rays instead of being fetched again for every ray or two.
 * Either way a cell needs at most rays + SH_SEGBUDGET items. */
#ifndef SH_SEG
#define SH_SEG 128
#endif
#ifndef SH_CROWD_NTRI
#define SH_CROWD_NTRI 256 /* lists longer than this count as long (measured: 64 costs config B 5 us, 512+ costs the crowded cases) */
#endif
struct ItemShape { u32 rpi, nseg, seglen; };
static __host__ __device__ __forceinline__ ItemShape sh_shape(u32 rays, u32 ntri)
{
	ItemShape s;
	if (rays >= 32 && ntri > SH_CROWD_NTRI) {
		const u32 chunks = (rays + 31) / 32;
		s.rpi = 32;
		s.nseg = min((ntri + SH_SEG - 1) / SH_SEG, max(1u, (rays + SH_SEGBUDGET) / chunks));
	} else {
		s.rpi = sh_rays_per_item(ntri);
		s.nseg = sh_segments(max(rays, 1u), ntri);
	}
	s.seglen = ((s.nseg == 1u ? ntri : (ntri + s.nseg - 1) / s.nseg) + 31u) & ~31u; /* whole staging batches (one segment, the usual case: no division) */
	return s;
}
static __host__ __device__ __forceinline__ u32 sh_items_of_cell(u32 rays, u32 ntri)
{
	if (!rays || !ntri) return 0;
	const ItemShape s = sh_shape(rays, ntri);
	return ((rays + s.rpi - 1) / s.rpi) * s.nseg;
}

struct RegroupArgs {
	const FrameConst *fc;
	int light, N, C;
	const float *t, *dir;
	const u32 *grid_span;       /* triangles per light cell (light grid) */
	u32 *light_cell;            /* [N]   d_map[N..2N) */
	u32 *ray_span;              /* [C+1] rays per cell, zeroed before the launch */
	u32 *ray_offset;            /* [C+2] */
	u32 *ray_rank;              /* [N]   position of every ray among the rays of its cell */
	u32 *ray_pix;               /* [N]   d_map[0..N) after the regroup */
	u32 *item_off;              /* [C+1] */
	uint2 *items;               /* (cell, local item index) */
	u32 *counters;              /* [0] number of items, [1] ticket (zeroed here) */
	u64 *state_a, *state_b;     /* look-back descriptors of the two scans; self-resetting */
	u32 item_cap;
};

/* PHASE < 0: all phases with grid barriers (cooperative launch); otherwise just that phase, compiled on its own so that
 * it gets the registers it needs and no more (the ray -> cell map is bound by the latency of its loads and atomics:
 * what hides it is resident warps) */
#ifndef RG_MINB_A
#define RG_MINB_A 2
#endif
template <int PHASE>
__global__ void __launch_bounds__(RG_THREADS, PHASE == 0 ? RG_MINB_A : 2) k_regroup(const RegroupArgs a)
{
	constexpr int phase = PHASE;
	__shared__ u32 s_warp[RG_THREADS / 32];
	__shared__ u32 s_pref[2];
	__shared__ float s_lat[100];
	ug_pdl_enter<UG_PDL_REGROUP>();
	const u32 G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
	const int N = a.N, C = a.C;
	const bool all = phase < 0;

	if (all || phase == 0) { /* ---- phase A: ray -> light cell, rays per cell */
		const ViewConst &cam = a.fc->cam;
		const ViewConst &lv = a.fc->light[a.light];
		const LightMap m = make_light_map(lv);
		const bool spherical = lv.grid_mode == UG_LIGHT_GRID_SPHERICAL;
		const float ex = cam.cc[0], ey = cam.cc[1], ez = cam.cc[2];
		if (b == 0 && tid == 0) a.counters[1] = 0;
		/* a.dir == nullptr: the direction is evaluated again from the lattice (the bits k_init_pixels stored) instead of read -
		 * 12 of the 16 bytes this phase reads per pixel */
		const bool recomp = a.dir == nullptr;
		const int W = cam.W, H = cam.H, row0 = cam.gy_lo * UG_TILE;
		if (recomp) {
			for (u32 i = tid; i < 100; i += RG_THREADS) s_lat[i] = cam.lattice[i];
			__syncthreads();
		}
		/* RG_KA pixels per thread and trip: all their loads are in flight before any arithmetic starts (the loop is a dependent
		 * chain load -> divide -> match -> atomic, and only 32 warps per SM are resident) */
		const int stride = (int)(G * RG_THREADS);
		for (int base = b * RG_THREADS; base < N; base += RG_KA * stride) {
			int pix[RG_KA];
			float tv[RG_KA], dx[RG_KA], dy[RG_KA], dz[RG_KA];
#pragma unroll
			for (int k = 0; k < RG_KA; k++) {
				pix[k] = base + k * stride + (int)tid;
				const bool valid = pix[k] < N;
				tv[k] = valid ? a.t[pix[k]] : 0.0f;
				dx[k] = 0.0f; dy[k] = 0.0f; dz[k] = 1.0f;
				if (valid && !recomp) { dx[k] = a.dir[3 * (size_t)pix[k]]; dy[k] = a.dir[3 * (size_t)pix[k] + 1]; dz[k] = a.dir[3 * (size_t)pix[k] + 2]; }
			}
			if (recomp) {
#pragma unroll
				for (int k = 0; k < RG_KA; k++)
					if (pix[k] < N) {
						const int pyl = pix[k] / W, px = pix[k] - pyl * W;
						ug_primary_dir(s_lat, ex, ey, ez, px, pyl + row0, W, H, dx[k], dy[k], dz[k]);
					}
			}
			u32 cell[RG_KA], old[RG_KA], lower[RG_KA];
#pragma unroll
			for (int k = 0; k < RG_KA; k++) {
				cell[k] = 0xffffffffu; old[k] = 0; lower[k] = 0;
				if (base + k * stride >= N) continue; /* warp-uniform: the rest of the trip lies past the end */
				const bool valid = pix[k] < N;
				if (valid) {
					cell[k] = spherical ? light_cell_sph(lv, ex, ey, ez, tv[k], dx[k], dy[k], dz[k]) : light_cell_of(m, ex, ey, ez, tv[k], dx[k], dy[k], dz[k]);
					a.light_cell[pix[k]] = cell[k];
				}
				/* rays of one warp mostly share a light cell: one atomic per distinct cell.  What it returns is where the group
				 * starts among the cell's rays, so the scatter phase needs no cursors of its own (any order inside a cell will do) */
				const u32 peers = __match_any_sync(UG_FULL_MASK, cell[k]);
				lower[k] = __popc(peers & ((1u << ug_lane()) - 1u));
				if (valid && lower[k] == 0) old[k] = atomicAdd(&a.ray_span[cell[k]], (u32)__popc(peers));
				old[k] |= (u32)(__ffs(peers) - 1) << 27; /* the leader's lane rides in the top bits until the atomics have returned */
			}
#pragma unroll
			for (int k = 0; k < RG_KA; k++) {
				if (base + k * stride >= N) continue;
				const u32 st = __shfl_sync(UG_FULL_MASK, old[k] & 0x7ffffffu, (int)(old[k] >> 27));
				if (pix[k] < N) a.ray_rank[pix[k]] = st + lower[k];
			}
		}
	}
	if (all) cg::this_grid().sync();

	if (all || phase == 1) { /* ---- phase B: ray_offset = exscan(ray_span[0..C]); item_off = exscan(items per cell [0..C-1]) */
		const u32 per = RG_THREADS * 4, n = (u32)C + 1, tiles = (n + 1 + per - 1) / per;
		for (u32 tile = b; tile < tiles; tile += G) {
			const u32 base = tile * per + tid * 4;
			u32 r[4], it[4], rs = 0, is = 0;
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const u32 c = base + k;
				r[k] = (c < n) ? a.ray_span[c] : 0u;
				it[k] = (c < (u32)C) ? sh_items_of_cell(r[k], a.grid_span[c]) : 0u;
				rs += r[k]; is += it[k];
			}
			u32 rt, itot;
			u32 rrun = ug_block_scan<RG_THREADS>(rs, s_warp, rt);
			u32 irun = ug_block_scan<RG_THREADS>(is, s_warp, itot);
			if (tid < 32) {
				const u32 p0 = ug_lookback(a.state_a, tile, rt);
				const u32 p1 = ug_lookback(a.state_b, tile, itot);
				if (tid == 0) { s_pref[0] = p0; s_pref[1] = p1; }
			}
			__syncthreads();
			rrun += s_pref[0]; irun += s_pref[1];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const u32 c = base + k;
				if (c <= n) a.ray_offset[c] = rrun;
				if (c <= (u32)C) a.item_off[c] = irun;
				rrun += r[k]; irun += it[k];
			}
			__syncthreads();
		}
	}
	if (all) cg::this_grid().sync();

	if (all || phase == 2) { /* ---- phase C: scatter the rays, write the item table, reset the scratch */
		const u32 per = RG_THREADS * 4, tiles = ((u32)C + 2 + per - 1) / per;
		for (u32 i = b * RG_THREADS + tid; i < tiles + 1; i += G * RG_THREADS) { a.state_a[i] = 0; a.state_b[i] = 0; }
		const int stride = (int)(G * RG_THREADS);
		for (int base = b * RG_THREADS; base < N; base += RG_KC * stride) { /* RG_KC pixels per trip: a ray goes to (start of its cell) + (its rank, from phase A) */
			int pix[RG_KC];
			u32 cell[RG_KC], rank[RG_KC], roff[RG_KC];
#pragma unroll
			for (int k = 0; k < RG_KC; k++) {
				pix[k] = base + k * stride + (int)tid;
				cell[k] = (pix[k] < N) ? a.light_cell[pix[k]] : 0u;
				rank[k] = (pix[k] < N) ? a.ray_rank[pix[k]] : 0u;
			}
#pragma unroll
			for (int k = 0; k < RG_KC; k++) roff[k] = (pix[k] < N) ? a.ray_offset[cell[k]] : 0u;
#pragma unroll
			for (int k = 0; k < RG_KC; k++)
				if (pix[k] < N) a.ray_pix[roff[k] + rank[k]] = (u32)pix[k];
		}
		const u32 total = a.item_off[C];
		if (b == 0 && tid == 0) a.counters[0] = min(total, a.item_cap);
		/* the item table: a thread per cell writes a short run itself (most cells hold a few items, or none - an image strip's
		 * rays reach only part of the light grid); a crowded cell (spherical light grids put tens of thousands of rays into
		 * one, and it has as many items as rays) is taken over by its whole warp */
		for (u32 c0 = (b * RG_THREADS + tid) & ~31u; c0 < (u32)C; c0 += G * RG_THREADS) {
			const u32 c = c0 + ug_lane();
			u32 o = 0, n = 0;
			if (c < (u32)C) { o = a.item_off[c]; n = a.item_off[c + 1] - o; }
			if (n <= 8)
				for (u32 q = 0; q < n; q++)
					if (o + q < a.item_cap) a.items[o + q] = make_uint2(c, q);
			u32 big = __ballot_sync(UG_FULL_MASK, n > 8);
			while (big) {
				const int L = __ffs(big) - 1;
				big &= big - 1;
				const u32 oc = __shfl_sync(UG_FULL_MASK, o, L), nc = __shfl_sync(UG_FULL_MASK, n, L);
				for (u32 q = ug_lane(); q < nc; q += 32)
					if (oc + q < a.item_cap) a.items[oc + q] = make_uint2(c0 + (u32)L, q);
			}
		}
	}
}

__global__ void __launch_bounds__(32 * SH_WARPS) k_shadow(const FrameConst *__restrict__ fc, int light, const u32 *__restrict__ span,
                                                        const u32 *__restrict__ offset, const u32 *__restrict__ value_list, const float4 *__restrict__ rec,
                                                        const u32 *__restrict__ ray_pix, const u32 *__restrict__ ray_span, const u32 *__restrict__ ray_offset,
                                                        const uint2 *__restrict__ items, u32 *__restrict__ counters, const float *__restrict__ t,
                                                        const float *__restrict__ dir, int *__restrict__ shadowed)
{
	__shared__ float4 s_rec[SH_WARPS][32 * 4];
	ug_pdl_enter<UG_PDL_SHADOW>();
	float4 *w_rec = s_rec[threadIdx.x >> 5];
	const u32 lane = ug_lane();
	const float ex = fc->cam.cc[0], ey = fc->cam.cc[1], ez = fc->cam.cc[2];
	const float lx = fc->light[light].cc[0], ly = fc->light[light].cc[1], lz = fc->light[light].cc[2];
	const u32 nitems = counters[0];
	while (true) {
		u32 i = 0;
		if (lane == 0) i = atomicAdd(&counters[1], 1u);
		i = __shfl_sync(UG_FULL_MASK, i, 0);
		if (i >= nitems) break;
		const uint2 it = items[i]; /* (cell, local item index) */
		const u32 cell = it.x, ntri = span[cell], toff = offset[cell], nrays = ray_span[cell], roff = ray_offset[cell];
		const ItemShape shp = sh_shape(nrays, ntri);
		const u32 rpi = shp.rpi, nseg = shp.nseg, seglen = shp.seglen;
		const u32 chunk = nseg == 1u ? it.y : it.y / nseg, seg = it.y - chunk * nseg;
		const u32 a = chunk * rpi, m = min(rpi, nrays - a);         /* this item's rays: [roff+a, roff+a+m) */
		const u32 t0 = toff + min(ntri, seg * seglen), t1 = toff + min(ntri, (seg + 1) * seglen);
		/* lane = (ray slot, triangle group) */
		u32 g, slot, grp; /* m is a power of two except in the last item of a cell: three integer divisions per item are worth a branch */
		if ((m & (m - 1u)) == 0u) { const u32 sh = (u32)__ffs((int)m) - 1u; g = 32u >> sh; slot = lane & (m - 1u); grp = lane >> sh; }
		else { g = 32u / m; slot = lane % m; grp = lane / m; }
		const bool usable = grp < g;
		const ShadowRay ray = make_shadow_ray(dir, t, (int)ray_pix[roff + a + slot], ex, ey, ez, lx, ly, lz);
		u32 slotmask = 0; /* the lanes that work on my ray */
		for (u32 q = 0; q < g; q++) slotmask |= 1u << (slot + q * m);
		bool done = (nseg > 1) && shadowed[ray.pix] == 1; /* already occluded by another segment / light: nothing to add to an OR */

		float4 p0 = make_float4(0, 0, 0, 0), p1 = p0, p2 = p0, p3 = p0;
		if (t0 + lane < t1) { /* first batch in flight */
			const float4 *src4 = rec + 4 * (size_t)value_list[t0 + lane];
			p0 = src4[0]; p1 = src4[1]; p2 = src4[2]; p3 = src4[3];
		}
		for (u32 base = t0; base < t1; base += 32) {
			const u32 cnt = min(32u, t1 - base);
			__syncwarp();
			if (lane < cnt) { w_rec[4 * lane + 0] = p0; w_rec[4 * lane + 1] = p1; w_rec[4 * lane + 2] = p2; w_rec[4 * lane + 3] = p3; }
			__syncwarp();
			if (base + 32 + lane < t1) { /* next batch's records in flight while this one is tested */
				const float4 *src4 = rec + 4 * (size_t)value_list[base + 32 + lane];
				p0 = src4[0]; p1 = src4[1]; p2 = src4[2]; p3 = src4[3];
			}
			if (usable && !done) {
				for (u32 k = grp; k < cnt; k += g) { /* (two tests in flight per ray, as in k_trace, was measured here: 0-9 % slower - an any-hit walk ends early) */
					float tt;
					if (!ug_intersect(w_rec[4 * k], w_rec[4 * k + 1], w_rec[4 * k + 2], w_rec[4 * k + 3], ray.Lx, ray.Ly, ray.Lz, tt)) continue;
					if (!(tt < 999999.9f) || tt == 0.0f) continue; /* intersectTri returns t only if t < oldt; `if (value)` */
					/* light_kernel.cu:193-202 + isSmaller :1-11 */
					const float qx = lx + tt * ray.Lx, qy = ly + tt * ray.Ly, qz = lz + tt * ray.Lz;
					const float ax = qx - lx, ay = qy - ly, az = qz - lz;
					const float dist_a = sqrtf(ax * ax + ay * ay + az * az);
					if (dist_a + 1e-03f < ray.dist_b) { done = true; shadowed[ray.pix] = 1; break; }
				}
			}
			const u32 hits = __ballot_sync(UG_FULL_MASK, done && usable);
			if (hits & slotmask) done = true;
			if (__all_sync(UG_FULL_MASK, done || !usable)) break; /* every ray of the item is occluded */
		}
	}
}

/* host side ------------------------------------------------------------------------------------------ */
bool ugk_regroup_fused_supported(int device)
{
	int coop = 0, n = 0;
	cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
	if (!coop) return false;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_regroup<-1>, RG_THREADS, 0) != cudaSuccess) return false;
	return n >= 2;
}
size_t ugk_shadow_item_cap(int N, int C) { return (size_t)N + (size_t)(SH_SEGBUDGET + 1) * (size_t)C; }
/* host view of the item policy, so a CPU test can check the bound the table's size rests on: items <= rays + SH_SEGBUDGET */
extern "C" void ug_shadow_item_shape(unsigned rays, unsigned ntri, unsigned out[4])
{
	const ItemShape sh = sh_shape(rays ? rays : 1u, ntri);
	out[0] = sh.rpi; out[1] = sh.nseg; out[2] = sh.seglen; out[3] = sh_items_of_cell(rays, ntri);
}
size_t ugk_regroup_state_entries(int C) { return ((size_t)C + 2 + RG_THREADS * 4 - 1) / (RG_THREADS * 4) + 2; }

/* getRayGridMapping + processData (per_frame_funcs.h:72-137) for one light; returns false if the launch failed.
 * phases: bit 0 = ray -> light cell + rays per cell (needs the primary hits only), bits 1, 2 = scans + scatter + item table
 * (need the light grid's cell spans as well); 7 = everything */
bool ugk_regroup(cudaStream_t s, const FrameConst *fc, int light, const GridBuffers &g, const LightRays &r, const float *t, const float *dir, int N, int C,
                 int sms, bool fused, int *launches, int phases, bool recompute_dir)
{
	RegroupArgs a;
	a.fc = fc; a.light = light; a.N = N; a.C = C; a.t = t; a.dir = recompute_dir ? nullptr : dir; a.grid_span = g.cell_span();
	a.light_cell = r.cell; a.ray_span = r.span; a.ray_offset = r.offset; a.ray_rank = r.rank; a.ray_pix = r.pix;
	a.item_off = r.item_off; a.items = r.items; a.counters = r.counters; a.state_a = r.state_a; a.state_b = r.state_b;
	a.item_cap = (u32)r.item_cap;
	if (fused && phases == 7) {
		void *args[] = {(void *)&a};
		if (cudaLaunchCooperativeKernel((const void *)k_regroup<-1>, dim3((unsigned)sms * 2), dim3(RG_THREADS), args, 0, s) == cudaSuccess) {
			if (launches) *launches += 1;
			return true;
		}
		cudaGetLastError();
		return false;
	}
	const unsigned blocks = (unsigned)sms * 2;
	if (phases & 1) { ug_launch(UG_PDL_REGROUP, k_regroup<0>, dim3((unsigned)sms * RG_MINB_A), dim3(RG_THREADS), s, a); if (launches) *launches += 1; }
	if (phases & 2) { ug_launch(UG_PDL_REGROUP, k_regroup<1>, dim3(blocks), dim3(RG_THREADS), s, a); if (launches) *launches += 1; }
	if (phases & 4) { ug_launch(UG_PDL_REGROUP, k_regroup<2>, dim3(blocks), dim3(RG_THREADS), s, a); if (launches) *launches += 1; }
	return true;
}

void ugk_shadow(cudaStream_t s, const FrameConst *fc, int light, const GridBuffers &g, const LightRays &r, const float *t, const float *dir, int *shadowed,
                int N, int C, int sms)
{
	if (C <= 0 || N <= 0) return;
	const size_t want = (r.item_cap + SH_WARPS - 1) / SH_WARPS;
	const unsigned blocks = (unsigned)(want < (size_t)sms * 16 ? want : (size_t)sms * 16);
	ug_launch(UG_PDL_SHADOW, k_shadow, dim3(blocks), dim3(32 * SH_WARPS), s, fc, light, g.cell_span(), g.cell_offset(), g.vals[g.final_buf], g.tri_rec, r.pix, r.span, r.offset, r.items,
	          r.counters, t, dir, shadowed);
}
