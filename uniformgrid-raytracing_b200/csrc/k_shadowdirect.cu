/* k_shadowdirect.cu -- shadow rays (rows a13, a16-a18) WITHOUT regrouping the rays.
 *
 * The reference sorts all N shadow rays by light cell (DecisionData, decision_data.h:171-271: key-value sort, segmented
 * scan, compaction) so that a CTA can walk one light cell's triangle list for 64 of them (mod_light_rckernel,
 * light_kernel.cu:52-270).  k_shadow.cu keeps that shape (counting sort + work-item table).  The shadow bit is an OR over the
 * triangles of the ray's own light cell, so nothing requires the rays of a cell to be together - only that every ray is
 * tested against its cell's list.  Here the rays stay where they are:
 *
 *   a warp takes an 8x4 block of pixels (half a tile: neighbouring shadow rays fall into the same few light cells),
 *   computes ray -> light cell for its 32 rays (mapSortkernel, misc_kernel.cu:63-129, or the spherical map :255-296),
 *   and for every DISTINCT cell among them walks that cell's list with the rays that share it: m rays x g = 32/m
 *   triangle groups, every lane tests every g-th staged triangle for its ray, hits are OR-ed by ballot, a ray stops at its
 *   first occluder and the walk stops when all m are occluded.
 *
 * No ray -> cell histogram, no scans, no scatter of pixel ids, no gather of (t, dir) by pixel id, no item table: three
 * launches and about 40 B of scattered traffic per pixel disappear.  What the regrouped path got from its item table is
 * kept differently: a (ray group, list) product that is too long for one warp (crowded border cells, spherical light grids)
 * is not walked inline but appended to a queue as (cell, pixel block, lane mask, list segment) and walked by k_shadow_heavy,
 * one segment per warp, so the long lists are spread over the machine.
 *
 * Triangle records are staged by TMA: every lane issues one 64-byte cp.async.bulk (global -> shared, completion counted on
 * an mbarrier) for the record its list entry names; two buffers per warp, so the next batch is in flight while the current
 * one is tested.  No registers, no STS and no __syncwarp between load and use.
 */
#include "k_shadow_dev.cuh"

#define SD_WARPS 4
#define SD_BATCH 32          /* records per staged batch: one per lane */
#define SD_INLINE_TESTS 2048 /* a (ray group) x (list) product with more ray-triangle tests than this is walked by k_shadow_heavy ... */
#define SD_UNIT_TESTS 4096   /* ... in list segments of about this many tests (whole batches, at most SD_SEGMAX triangles) */
#define SD_SEGMAX 2048

/* ---- TMA bulk copy + mbarrier (PTX ISA: cp.async.bulk, mbarrier) */
static __device__ __forceinline__ void sd_mbar_init(u32 bar, u32 count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
static __device__ __forceinline__ void sd_mbar_expect(u32 bar, u32 bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void sd_mbar_wait(u32 bar, u32 parity)
{
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "SD_WAIT:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra SD_DONE;\n"
	    "bra SD_WAIT;\n"
	    "SD_DONE:\n"
	    "}\n" ::"r"(bar),
	    "r"(parity)
	    : "memory");
}
static __device__ __forceinline__ void sd_bulk_g2s(u32 dst_smem, const void *src, u32 bytes, u32 bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
	             : "memory");
}

struct ShadowDirectArgs {
	const FrameConst *fc;
	int light, N, C;
	const u32 *span, *offset, *vals; /* the light grid: triangles per cell, list start, triangle ids */
	const float4 *rec;               /* [4T] ray-independent Moeller-Trumbore records for this light's origin */
	const float *t, *dir;
	int *shadowed;
	u32 *light_cell;                 /* [N] d_map[N..2N) */
	uint4 *heavy;                    /* queue of (cell, pixel block, lane mask, segment | segment length / 32 << 20) */
	u32 heavy_cap;
	u32 *counters;                   /* [0] queue length, [1] ticket of k_shadow_heavy, [2] ticket of k_shadow_direct */
};

/* One group: the lanes of `gmask` hold rays of light cell `cell`; walk list entries [t0, t1) for them.  Lane layout:
 * slot = lane % m is the ray, lane / m the triangle group.  `ray` is the calling lane's OWN ray (valid for lanes in gmask). */
static __device__ __forceinline__ void sd_walk(const ShadowDirectArgs &a, u32 gmask, u32 t0, u32 t1, const ShadowRay &own, float lx, float ly, float lz,
                                               float4 *w_rec, u32 bar0, u32 &phase0, u32 &phase1)
{
	const u32 lane = ug_lane();
	const u32 m = (u32)__popc(gmask), g = 32u / m, slot = lane % m, grp = lane / m;
	const bool usable = grp < g;
	const int src = (int)__fns(gmask, 0, (int)slot + 1); /* the lane that owns ray number `slot` of the group */
	ShadowRay ray;
	ray.Lx = __shfl_sync(UG_FULL_MASK, own.Lx, src);
	ray.Ly = __shfl_sync(UG_FULL_MASK, own.Ly, src);
	ray.Lz = __shfl_sync(UG_FULL_MASK, own.Lz, src);
	ray.dist_b = __shfl_sync(UG_FULL_MASK, own.dist_b, src);
	ray.pix = __shfl_sync(UG_FULL_MASK, own.pix, src);
	u32 slotmask = 0; /* the lanes that work on my ray */
	for (u32 q = 0; q < g; q++) slotmask |= 1u << (slot + q * m);
	bool done = false;
	const u32 rec_base = (u32)__cvta_generic_to_shared(w_rec);
#ifndef SD_NO_TMA
	/* batch b goes to buffer b & 1; its barrier completes when the 64 bytes of every issued record have landed */
	u32 nb = 0;
	{
		const u32 cnt = min((u32)SD_BATCH, t1 - t0);
		if (lane == 0) sd_mbar_expect(bar0, cnt * 64u);
		__syncwarp();
		if (lane < cnt) sd_bulk_g2s(rec_base + lane * 64u, a.rec + 4 * (size_t)a.vals[t0 + lane], 64u, bar0);
	}
	for (u32 base = t0; base < t1; base += SD_BATCH, nb++) {
		const u32 cnt = min((u32)SD_BATCH, t1 - base), buf = nb & 1u;
		if (base + SD_BATCH < t1) { /* next batch into the other buffer (its last readers passed the __syncwarp below) */
			const u32 cnt2 = min((u32)SD_BATCH, t1 - base - SD_BATCH), bar2 = bar0 + 8u * (buf ^ 1u);
			if (lane == 0) sd_mbar_expect(bar2, cnt2 * 64u);
			__syncwarp();
			if (lane < cnt2) sd_bulk_g2s(rec_base + (buf ^ 1u) * (SD_BATCH * 64u) + lane * 64u, a.rec + 4 * (size_t)a.vals[base + SD_BATCH + lane], 64u, bar2);
		}
		sd_mbar_wait(bar0 + 8u * buf, buf ? phase1 : phase0);
		if (buf) phase1 ^= 1u; else phase0 ^= 1u;
#else
	/* comparison build: the same walk with LDG.128 -> registers -> STS.128 staging (what k_shadow.cu does) */
	u32 nb = 0;
	float4 p0 = make_float4(0, 0, 0, 0), p1 = p0, p2 = p0, p3 = p0;
	if (t0 + lane < t1) {
		const float4 *src4 = a.rec + 4 * (size_t)a.vals[t0 + lane];
		p0 = src4[0]; p1 = src4[1]; p2 = src4[2]; p3 = src4[3];
	}
	(void)rec_base; (void)bar0;
	for (u32 base = t0; base < t1; base += SD_BATCH, nb++) {
		const u32 cnt = min((u32)SD_BATCH, t1 - base), buf = 0;
		__syncwarp();
		if (lane < cnt) { w_rec[4 * lane + 0] = p0; w_rec[4 * lane + 1] = p1; w_rec[4 * lane + 2] = p2; w_rec[4 * lane + 3] = p3; }
		__syncwarp();
		if (base + SD_BATCH + lane < t1) {
			const float4 *src4 = a.rec + 4 * (size_t)a.vals[base + SD_BATCH + lane];
			p0 = src4[0]; p1 = src4[1]; p2 = src4[2]; p3 = src4[3];
		}
#endif
		const float4 *cur = w_rec + buf * (SD_BATCH * 4);
		if (usable && !done) {
			for (u32 k = grp; k < cnt; k += g) { /* (two tests in flight per ray, as in k_trace, was measured here: 0-9 % slower - an any-hit walk ends early) */
				float tt;
				if (!ug_intersect(cur[4 * k], cur[4 * k + 1], cur[4 * k + 2], cur[4 * k + 3], ray.Lx, ray.Ly, ray.Lz, tt)) continue;
				if (!(tt < 999999.9f) || tt == 0.0f) continue; /* intersectTri returns t only if t < oldt; `if (value)` */
				/* light_kernel.cu:193-202 + isSmaller :1-11 */
				const float qx = lx + tt * ray.Lx, qy = ly + tt * ray.Ly, qz = lz + tt * ray.Lz;
				const float ax = qx - lx, ay = qy - ly, az = qz - lz;
				const float dist_a = sqrtf(ax * ax + ay * ay + az * az);
				if (dist_a + 1e-03f < ray.dist_b) { done = true; a.shadowed[ray.pix] = 1; break; }
			}
		}
		const u32 hits = __ballot_sync(UG_FULL_MASK, done && usable);
		if (hits & slotmask) done = true;
		const bool all_done = __all_sync(UG_FULL_MASK, done || !usable); /* also orders this batch's reads before the buffer is refilled */
		if (all_done) {
#ifndef SD_NO_TMA
			if (base + SD_BATCH < t1) { /* a batch is still in flight: take it in before the buffers are reused */
				const u32 b2 = buf ^ 1u;
				sd_mbar_wait(bar0 + 8u * b2, b2 ? phase1 : phase0);
				if (b2) phase1 ^= 1u; else phase0 ^= 1u;
			}
#endif
			break;
		}
	}
	__syncwarp();
}

/* the calling lane's shadow ray and light cell (misc_kernel.cu:74-116 / :255-296; light_kernel.cu:164-184: same vector) */
static __device__ __forceinline__ u32 sd_ray_and_cell(const ViewConst &lv, const LightMap &map, bool spherical, float ex, float ey, float ez, float tv, float dx,
                                                      float dy, float dz, int pix, ShadowRay &ray)
{
	const float Px = ex + tv * dx, Py = ey + tv * dy, Pz = ez + tv * dz;
	ray.Lx = Px - map.lx; ray.Ly = Py - map.ly; ray.Lz = Pz - map.lz;
	const float bx = ray.Lx, by = ray.Ly, bz = ray.Lz;
	ug_normalize(ray.Lx, ray.Ly, ray.Lz);
	ray.dist_b = sqrtf(bx * bx + by * by + bz * bz);
	ray.pix = pix;
	return spherical ? light_cell_sph(lv, ex, ey, ez, tv, dx, dy, dz) : light_cell_of(map, ex, ey, ez, tv, dx, dy, dz);
}

/* pixel of lane `lane` in pixel block `blk` of a strip W pixels wide: blocks are the two 8x4 halves of the 8x8 tiles */
static __device__ __forceinline__ int sd_pixel(u32 blk, u32 lane, int W)
{
	const u32 tiles_x = (u32)W / UG_TILE;
	const u32 tile = blk >> 1, half = blk & 1u;
	const u32 ty = tile / tiles_x, tx = tile - ty * tiles_x;
	return (int)((ty * UG_TILE + half * 4u + (lane >> 3)) * (u32)W + tx * UG_TILE + (lane & 7u));
}

__global__ void __launch_bounds__(32 * SD_WARPS, 8) k_shadow_direct(const ShadowDirectArgs a)
{
	__shared__ __align__(128) float4 s_rec[SD_WARPS][2 * SD_BATCH * 4];
	__shared__ __align__(8) unsigned long long s_bar[SD_WARPS][2];
	const u32 lane = ug_lane(), warp = threadIdx.x >> 5;
	float4 *w_rec = s_rec[warp];
	const u32 bar0 = (u32)__cvta_generic_to_shared(&s_bar[warp][0]);
	if (lane == 0) { sd_mbar_init(bar0, 1); sd_mbar_init(bar0 + 8u, 1); }
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncwarp();
	u32 phase0 = 0, phase1 = 0;
	const ViewConst &lv = a.fc->light[a.light];
	const LightMap map = make_light_map(lv);
	const bool spherical = lv.grid_mode == UG_LIGHT_GRID_SPHERICAL;
	const float ex = a.fc->cam.cc[0], ey = a.fc->cam.cc[1], ez = a.fc->cam.cc[2];
	const int W = a.fc->cam.W;
	const u32 nblk = (u32)a.N / 32u, C = (u32)a.C;
	while (true) { /* pixel blocks are handed out by ticket: their cost varies with the lists their rays fall into */
		u32 blk = 0;
		if (lane == 0) blk = atomicAdd(&a.counters[2], 1u);
		blk = __shfl_sync(UG_FULL_MASK, blk, 0);
		if (blk >= nblk) break;
		const int pix = sd_pixel(blk, lane, W);
		const float tv = a.t[pix], dx = a.dir[3 * (size_t)pix], dy = a.dir[3 * (size_t)pix + 1], dz = a.dir[3 * (size_t)pix + 2];
		ShadowRay own;
		const u32 cell = sd_ray_and_cell(lv, map, spherical, ex, ey, ez, tv, dx, dy, dz, pix, own);
		a.light_cell[pix] = cell;
		u32 ntri = 0, toff = 0;
		if (cell < C) { ntri = a.span[cell]; toff = a.offset[cell]; }
		const u32 peers = __match_any_sync(UG_FULL_MASK, cell);
		u32 leaders = __ballot_sync(UG_FULL_MASK, ntri != 0 && (peers & ((1u << lane) - 1u)) == 0); /* first lane of every cell that has triangles */
		while (leaders) {
			const int L = __ffs(leaders) - 1;
			leaders &= leaders - 1;
			const u32 gmask = __shfl_sync(UG_FULL_MASK, peers, L), n = __shfl_sync(UG_FULL_MASK, ntri, L), o = __shfl_sync(UG_FULL_MASK, toff, L);
			const u32 mm = (u32)__popc(gmask);
			if (mm * n > SD_INLINE_TESTS) { /* too much for one warp in one go: one queue entry per list segment, walked by whichever warp is free */
				const u32 seglen = min((u32)SD_SEGMAX, max(32u, ((SD_UNIT_TESTS / mm) + 31u) & ~31u));
				const u32 nseg = (n + seglen - 1) / seglen;
				u32 q = 0xffffffffu; /* reserve nseg entries; the counter only ever grows, a reservation that does not fit is not made */
				if (lane == 0) {
					u32 old = *(volatile u32 *)&a.counters[0];
					while (old + nseg <= a.heavy_cap) {
						const u32 seen = atomicCAS(&a.counters[0], old, old + nseg);
						if (seen == old) { q = old; break; }
						old = seen;
					}
				}
				q = __shfl_sync(UG_FULL_MASK, q, 0);
				if (q != 0xffffffffu) {
					const u32 c = __shfl_sync(UG_FULL_MASK, cell, L);
					for (u32 sgm = lane; sgm < nseg; sgm += 32) a.heavy[q + sgm] = make_uint4(c, blk, gmask, sgm | ((seglen / 32u) << 20)); /* segment, its length in batches */
					continue;
				}
				/* queue full: walk it here */
			}
			sd_walk(a, gmask, o, o + n, own, map.lx, map.ly, map.lz, w_rec, bar0, phase0, phase1);
		}
	}
}

/* the queued (ray group, list segment) products, one per warp at a time */
__global__ void __launch_bounds__(32 * SD_WARPS) k_shadow_heavy(const ShadowDirectArgs a)
{
	__shared__ __align__(128) float4 s_rec[SD_WARPS][2 * SD_BATCH * 4];
	__shared__ __align__(8) unsigned long long s_bar[SD_WARPS][2];
	const u32 lane = ug_lane(), warp = threadIdx.x >> 5;
	float4 *w_rec = s_rec[warp];
	const u32 bar0 = (u32)__cvta_generic_to_shared(&s_bar[warp][0]);
	if (lane == 0) { sd_mbar_init(bar0, 1); sd_mbar_init(bar0 + 8u, 1); }
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncwarp();
	u32 phase0 = 0, phase1 = 0;
	const ViewConst &lv = a.fc->light[a.light];
	const LightMap map = make_light_map(lv);
	const bool spherical = lv.grid_mode == UG_LIGHT_GRID_SPHERICAL;
	const float ex = a.fc->cam.cc[0], ey = a.fc->cam.cc[1], ez = a.fc->cam.cc[2];
	const int W = a.fc->cam.W;
	const u32 nq = min(a.counters[0], a.heavy_cap);
	while (true) {
		u32 i = 0;
		if (lane == 0) i = atomicAdd(&a.counters[1], 1u);
		i = __shfl_sync(UG_FULL_MASK, i, 0);
		if (i >= nq) break;
		const uint4 it = a.heavy[i]; /* (cell, pixel block, lane mask, segment) */
		const int pix = sd_pixel(it.y, lane, W);
		ShadowRay own;
		own.Lx = own.Ly = own.Lz = 0.0f; own.dist_b = 0.0f; own.pix = pix;
		bool occluded = true;
		if ((it.z >> lane) & 1u) {
			const float tv = a.t[pix], dx = a.dir[3 * (size_t)pix], dy = a.dir[3 * (size_t)pix + 1], dz = a.dir[3 * (size_t)pix + 2];
			sd_ray_and_cell(lv, map, spherical, ex, ey, ez, tv, dx, dy, dz, pix, own);
			occluded = a.shadowed[pix] == 1; /* by another segment or light: nothing to add to an OR */
		}
		const u32 gmask = it.z & ~__ballot_sync(UG_FULL_MASK, occluded);
		if (!gmask) continue;
		const u32 n = a.span[it.x], o = a.offset[it.x];
		const u32 sgm = it.w & 0xfffffu, seglen = (it.w >> 20) * 32u;
		const u32 t0 = o + min(n, sgm * seglen), t1 = o + min(n, (sgm + 1) * seglen);
		if (t0 < t1) sd_walk(a, gmask, t0, t1, own, map.lx, map.ly, map.lz, w_rec, bar0, phase0, phase1);
	}
}

__global__ void k_shadow_direct_reset(u32 *counters) { counters[0] = 0; counters[1] = 0; counters[2] = 0; }

size_t ugk_shadow_heavy_cap(int N) { return (size_t)N / 2 + 4096; }

/* mapSortkernel + processData + check_for_shadows (per_frame_funcs.h:72-159) for one light, rays in screen order */
void ugk_shadow_direct(cudaStream_t s, const FrameConst *fc, int light, const GridBuffers &g, const LightRays &r, const float *t, const float *dir, int *shadowed,
                       int N, int C, int sms, int *launches)
{
	if (C <= 0 || N <= 0) return;
	ShadowDirectArgs a;
	a.fc = fc; a.light = light; a.N = N; a.C = C;
	a.span = g.cell_span(); a.offset = g.cell_offset(); a.vals = g.vals[g.final_buf]; a.rec = g.tri_rec;
	a.t = t; a.dir = dir; a.shadowed = shadowed; a.light_cell = r.cell; a.heavy = r.heavy; a.heavy_cap = (u32)r.heavy_cap; a.counters = r.counters;
	k_shadow_direct_reset<<<1, 1, 0, s>>>(r.counters);
	const size_t want = ((size_t)N / 32 + SD_WARPS - 1) / SD_WARPS;
	const unsigned blocks = (unsigned)(want < (size_t)sms * 16 ? (want ? want : 1) : (size_t)sms * 16);
	k_shadow_direct<<<blocks, 32 * SD_WARPS, 0, s>>>(a);
	k_shadow_heavy<<<(unsigned)sms * 8, 32 * SD_WARPS, 0, s>>>(a);
	if (launches) *launches += 3;
}
