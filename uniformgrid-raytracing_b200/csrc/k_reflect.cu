/* k_reflect.cu -- row R: reflection rays through a world-space UNIFORM grid with a 3-D DDA and ray compaction
 * by scan between bounces (BASELINE north_star item 5).
 *
 * The reference contains no reflection / uniform-grid / DDA code (uniform_grid.h is a dead clone of the
 * perspective build), so the semantics are the authored spec in oracle/ug_oracle.c (orc_reflect):
 * nearest hit = minimum over ALL triangles of the signed t > 0 computed with intersectTri's arithmetic
 * (light_kernel.cu:13-50), ties to the lower triangle id.  That definition is grid independent; the uniform grid
 * only has to be conservative:
 *   - a triangle is listed in every cell its bounding box (grown by a small margin) overlaps,
 *   - the DDA tests whole cell lists, keeps (t, id) with an order-independent rule (smaller t, then smaller id;
 *     a triangle may be met in several cells) and only stops when the best t lies clearly before the exit of the
 *     current cell.
 * Grid build = the same count -> scan -> emit -> stable radix sort -> span/offset chain as the perspective grids
 * with a 3-D key (k_sort.cu primitives); the resolution is chosen on the device from the scene's bounding box.
 */
#include "k_shade_dev.cuh"
#include "k_ugrid_dev.cuh"

#ifndef UG_REFL_LOOK
#define UG_REFL_LOOK 2            /* DDA look-ahead (cells whose spans are fetched together) */
#endif

/* ---------------------------------------------------------------- bounding box + resolution (kernel-chain build) */
__global__ void k_aabb_reset(u32 *box) { if (threadIdx.x < 3) box[threadIdx.x] = 0xffffffffu; else if (threadIdx.x < 6) box[threadIdx.x] = 0u; }

__global__ void __launch_bounds__(256) k_aabb(const float *__restrict__ verts, int V, u32 *box)
{
	float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
	for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x)
		for (int k = 0; k < 3; k++) { const float x = verts[3 * (size_t)v + k]; lo[k] = fminf(lo[k], x); hi[k] = fmaxf(hi[k], x); }
	for (int k = 0; k < 3; k++) {
		for (int o = 16; o > 0; o >>= 1) {
			lo[k] = fminf(lo[k], __shfl_xor_sync(UG_FULL_MASK, lo[k], o));
			hi[k] = fmaxf(hi[k], __shfl_xor_sync(UG_FULL_MASK, hi[k], o));
		}
		if (ug_lane() == 0) { atomicMin(&box[k], f2ord(lo[k])); atomicMax(&box[3 + k], f2ord(hi[k])); }
	}
}

__global__ void k_ugrid_setup(const u32 *box, int T, int max_cells, float cells_per_tri, UGridConst *uc)
{
	if (threadIdx.x || blockIdx.x) return;
	ugrid_resolution(box, T, max_cells, cells_per_tri, uc);
}

/* ---------------------------------------------------------------- grid build (kernel chain; k_ugridfused.cu is the one-launch version) */
__global__ void __launch_bounds__(256) k_ugrid_count(const UGridConst *__restrict__ ucp, const int *__restrict__ faces, const float *__restrict__ verts, int T,
                                                    float4 *__restrict__ geo, float4 *__restrict__ tbox, uint4 *__restrict__ rect, u32 *__restrict__ count)
{
	const UGridConst u = *ucp;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) count[t] = ugrid_tri_setup(u, faces, verts, t, geo, tbox, rect);
}

__global__ void __launch_bounds__(256) k_ugrid_emit(const UGridConst *__restrict__ ucp, const u32 *__restrict__ scan, const uint4 *__restrict__ rect, int T,
                                                   u32 cap, u32 *__restrict__ keys, u32 *__restrict__ vals, u32 *__restrict__ span, u32 *__restrict__ counters)
{
	__shared__ int s_range[2];
	const int ny = ucp->n[1], nz = ucp->n[2];
	const u32 Ptrue = scan[T], P = min(Ptrue, cap);
	if (blockIdx.x == 0 && threadIdx.x == 0) { counters[0] = P; counters[1] = (Ptrue > cap) ? 1u : 0u; counters[2] = Ptrue; }
	const u32 per = 512;
	for (u32 blk = blockIdx.x * per; blk < P; blk += gridDim.x * per) {
		const u32 blk_end = min(blk + per, P);
		__syncthreads();
		if (threadIdx.x < 2) s_range[threadIdx.x] = ug_upper_bound(scan, 0, T + 1, threadIdx.x == 0 ? blk : blk_end - 1) - 1;
		__syncthreads();
		const int tlo = s_range[0], thi = s_range[1];
		for (u32 p = blk + threadIdx.x; p < blk_end; p += blockDim.x) {
			const int t = ug_upper_bound(scan, tlo, thi + 1, p) - 1;
			const u32 key = ugrid_pair_key(rect[t], p - scan[t], ny, nz);
			keys[p] = key;
			vals[p] = (u32)t;
			atomicAdd(&span[key], 1u);
		}
	}
}

/* counting fill instead of a sort: the order of a cell's list does not matter to the DDA (nearest hit with ties to the
 * lower id is order independent), so every pair takes the next free slot of its cell */
__global__ void __launch_bounds__(256) k_ugrid_fill(const UGridConst *__restrict__ ucp, const u32 *__restrict__ counters, const u32 *__restrict__ keys,
                                                   const u32 *__restrict__ vals, const float4 *__restrict__ tbox, const u32 *__restrict__ offset,
                                                   u32 *__restrict__ cursor, uint4 *__restrict__ prec)
{
	const UGridConst u = *ucp;
	const u32 P = counters[0];
	for (u32 p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
		const u32 key = keys[p];
		prec[offset[key] + atomicAdd(&cursor[key], 1u)] = ugrid_pair_record(u, key, vals[p], tbox);
	}
}

__global__ void __launch_bounds__(256) k_ugrid_cells(const UGridConst *__restrict__ ucp, const u32 *__restrict__ span, const u32 *__restrict__ offset,
                                                    const uint4 *__restrict__ prec, uint4 *__restrict__ cellrec)
{
	const u32 C = (u32)ucp->cells;
	for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) cellrec[c] = ugrid_cell_record(offset[c], span[c], prec);
}

/* ---------------------------------------------------------------- rays */
/* intersectTri (light_kernel.cu:13-50) for an arbitrary origin: tvec = O - v0 */
static __device__ __forceinline__ bool intersect_full(float4 v0, float4 e1, float4 e2, float Ox, float Oy, float Oz, float dx, float dy, float dz, float &t)
{
	const float tx = Ox - v0.x, ty = Oy - v0.y, tz = Oz - v0.z;
	float px, py, pz, qx, qy, qz;
	ug_cross(px, py, pz, dx, dy, dz, e2.x, e2.y, e2.z);
	const float det = ug_dot(e1.x, e1.y, e1.z, px, py, pz);
	if (fabsf(det) <= 1e-21f) return false;
	const float inv = 1.0f / det;
	const float u = ug_dot(tx, ty, tz, px, py, pz) * inv;
	if (u < 0.0f || u > 1.0f) return false;
	ug_cross(qx, qy, qz, tx, ty, tz, e1.x, e1.y, e1.z);
	const float v = ug_dot(dx, dy, dz, qx, qy, qz) * inv;
	if (v < 0.0f || u + v > 1.0f) return false;
	t = ug_dot(e2.x, e2.y, e2.z, qx, qy, qz) * inv;
	return true;
}

/* spec reflect(P, d, tri): n = normalize(cross(e1,e2)) facing the incoming ray, O = P + 1e-3 n, r = d - (2 d.n) n */
static __device__ __forceinline__ void reflect_ray(float4 e1, float4 e2, float Px, float Py, float Pz, float dx, float dy, float dz, float *O, float *r)
{
	float nx, ny, nz;
	ug_cross(nx, ny, nz, e1.x, e1.y, e1.z, e2.x, e2.y, e2.z);
	ug_normalize(nx, ny, nz);
	if (ug_dot(nx, ny, nz, dx, dy, dz) > 0) { nx = -nx; ny = -ny; nz = -nz; }
	const float s = 2.0f * ug_dot(dx, dy, dz, nx, ny, nz);
	O[0] = Px + 1e-3f * nx; O[1] = Py + 1e-3f * ny; O[2] = Pz + 1e-3f * nz;
	r[0] = dx - s * nx; r[1] = dy - s * ny; r[2] = dz - s * nz;
}

struct ReflectArgs {
	const FrameConst *fc;
	const UGridConst *uc;
	const float4 *geo;           /* [3T] v0, e1, e2 */
	const uint4 *prec;           /* [P] per-pair record: triangle id + its box clipped to the cell in 16-bit cell units */
	const uint4 *cellrec;        /* [C] (offset, span, union box lo, hi) of every cell */
	const u32 *span, *offset, *vals;
	const int *mat_idx;
	const float *mat_list, *mat_reflect;
	int mat_count, N, bounces;
	const float *t, *dir;
	const int *id;
	unsigned char *rgb;
	unsigned char *rgb_remote; /* fused multi-GPU gather: rank 0's copy of these pixels (P2P stores), or null */
	float *O, *r, *w, *k0;       /* per-pixel ray state [3N],[3N], running weight [N], first-hit reflectivity [N] */
	float *term;                 /* [bounces][3N]: what bounce b adds to the pixel (w * local colour) */
	u32 *depth;                  /* [N] number of valid terms */
	const u32 *list;             /* compacted pixel ids of this bounce's rays */
	u32 *next;                   /* ... and of the rays it spawns (compaction, see ug_append) */
	u32 *counts;                 /* [UG_MAX_BOUNCES + 1] rays per bounce: counts[b-1] = rays of bounce b */
};

/* Ray compaction: the survivors of a warp take consecutive slots - an exclusive scan of the warp's live flags (ballot +
 * popc) places each one, a single atomicAdd per warp reserves the run.  List order is not deterministic and does not
 * matter: every ray writes only its own pixel.  All 32 lanes must call. */
static __device__ __forceinline__ void ug_append(bool live, u32 value, u32 *__restrict__ list, u32 *counter)
{
	const u32 m = __ballot_sync(UG_FULL_MASK, live);
	if (m == 0) return;
	const int leader = __ffs(m) - 1;
	u32 base = 0;
	if ((int)ug_lane() == leader) base = atomicAdd(counter, (u32)__popc(m));
	base = __shfl_sync(UG_FULL_MASK, base, leader);
	if (live) list[base + (u32)__popc(m & ((1u << ug_lane()) - 1u))] = value;
}

/* bounce 0: which pixels spawn a reflection ray */
__global__ void __launch_bounds__(256) k_refl_init(const ReflectArgs a)
{
	ug_pdl_enter<UG_PDL_REFLECT>();
	const float ex = a.fc->cam.cc[0], ey = a.fc->cam.cc[1], ez = a.fc->cam.cc[2];
	/* a warp takes an 8x4 pixel block, not 32 pixels of one row: neighbouring lanes then spawn rays with neighbouring
	 * origins in both directions, which walk the same cells for longer (W is a multiple of 8, the strip height of 8 rows) */
	const int W = a.fc->cam.W, bw = W >> 3;
	for (int w0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); w0 < a.N; w0 += gridDim.x * blockDim.x) { /* warp-uniform trip count */
		const int blk = w0 >> 5, lane = (int)ug_lane();
		const int by = blk / bw, bx = blk - by * bw;
		const int pix = (by * 4 + (lane >> 3)) * W + bx * 8 + (lane & 7);
		bool live = false;
		if (pix < a.N) {
			const int tri = a.id[pix];
			const float tv = a.t[pix];
			if (tri >= 0 && tv > 0.0f) {
				const int m = a.mat_idx[tri];
				const float k = (m >= 0 && m < a.mat_count) ? a.mat_reflect[m] : 0.0f;
				if (k > 0) {
					const float dx = a.dir[3 * (size_t)pix], dy = a.dir[3 * (size_t)pix + 1], dz = a.dir[3 * (size_t)pix + 2];
					const float Px = ex + tv * dx, Py = ey + tv * dy, Pz = ez + tv * dz;
					float O[3], r[3];
					reflect_ray(a.geo[3 * (size_t)tri + 1], a.geo[3 * (size_t)tri + 2], Px, Py, Pz, dx, dy, dz, O, r);
					for (int c = 0; c < 3; c++) {
						a.O[3 * (size_t)pix + c] = O[c];
						a.r[3 * (size_t)pix + c] = r[c];
					}
					a.w[pix] = k;
					a.k0[pix] = k;
					a.depth[pix] = 0;
					live = true;
				}
			}
		}
		ug_append(live, (u32)pix, a.next, a.counts);
	}
}

/* one bounce for every live ray: 3-D DDA through the uniform grid, shade, spawn the next ray or finish the pixel */
#ifndef UG_REFL_MINB
#define UG_REFL_MINB 7 /* 72 registers, no spills (8 blocks/SM = 64 registers spilled 40 bytes; 6, 7 and 8 time the same: 137-138 us on config B) */
#endif
__global__ void __launch_bounds__(128, UG_REFL_MINB) k_refl_trace(const ReflectArgs a, int bounce)
{
	ug_pdl_enter<UG_PDL_REFLECT>();
	const UGridConst u = *a.uc;
	const float *R = a.fc->shade_mv;
	const float l0 = a.fc->light_pos[0], l1 = a.fc->light_pos[1], l2 = a.fc->light_pos[2];
	const float Lvx = R[0] * l0 + R[4] * l1 + R[8] * l2, Lvy = R[1] * l0 + R[5] * l1 + R[9] * l2, Lvz = R[2] * l0 + R[6] * l1 + R[10] * l2;
	const u32 nlive = a.counts[bounce - 1];
	const float hmin = fminf(u.h[0], fminf(u.h[1], u.h[2])), hmax = fmaxf(u.h[0], fmaxf(u.h[1], u.h[2]));
	const int nz = u.n[2], nyz = u.n[1] * u.n[2];
	for (u32 i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0 < nlive; i0 += gridDim.x * blockDim.x) { /* warp-uniform trip count */
		const u32 i = i0 + ug_lane();
		const bool active = i < nlive;
		const int pix = (int)a.list[active ? i : nlive - 1];
		const float Ox = a.O[3 * (size_t)pix], Oy = a.O[3 * (size_t)pix + 1], Oz = a.O[3 * (size_t)pix + 2];
		const float rx = a.r[3 * (size_t)pix], ry = a.r[3 * (size_t)pix + 1], rz = a.r[3 * (size_t)pix + 2];
		/* entry into the grid's box (the origin sits on a surface, normally inside) */
		float t_in = 0.0f, t_out = 3.0e38f;
		bool miss_box = false;
		{
			const float O[3] = {Ox, Oy, Oz}, r[3] = {rx, ry, rz};
#pragma unroll
			for (int k = 0; k < 3; k++) {
				const float hi = u.lo[k] + u.h[k] * (float)u.n[k];
				if (r[k] != 0.0f) {
					float t0 = (u.lo[k] - O[k]) / r[k], t1 = (hi - O[k]) / r[k];
					if (t0 > t1) { const float s = t0; t0 = t1; t1 = s; }
					t_in = fmaxf(t_in, t0); t_out = fminf(t_out, t1);
				} else if (O[k] < u.lo[k] || O[k] > hi) miss_box = true;
			}
		}
		float best = 99999999.9f;
		int hit = -1;
		if (active && !miss_box && t_in <= t_out) {
			/* DDA state in scalars (arrays indexed by the stepping axis would live in local memory), stepped without branches */
			int cx = ucell(u, 0, Ox + t_in * rx), cy = ucell(u, 1, Oy + t_in * ry), cz = ucell(u, 2, Oz + t_in * rz);
			const int sx = (rx > 0.0f) ? 1 : ((rx < 0.0f) ? -1 : 0), sy = (ry > 0.0f) ? 1 : ((ry < 0.0f) ? -1 : 0), sz = (rz > 0.0f) ? 1 : ((rz < 0.0f) ? -1 : 0);
			float tmx = 3.0e38f, tmy = 3.0e38f, tmz = 3.0e38f, tdx = 0.0f, tdy = 0.0f, tdz = 0.0f;
			if (sx) { tmx = ((u.lo[0] + u.h[0] * (float)(cx + (sx > 0 ? 1 : 0))) - Ox) / rx; tdx = u.h[0] / fabsf(rx); }
			if (sy) { tmy = ((u.lo[1] + u.h[1] * (float)(cy + (sy > 0 ? 1 : 0))) - Oy) / ry; tdy = u.h[1] / fabsf(ry); }
			if (sz) { tmz = ((u.lo[2] + u.h[2] * (float)(cz + (sz > 0 ? 1 : 0))) - Oz) / rz; tdz = u.h[2] / fabsf(rz); }
			int cell = cx * nyz + cy * nz + cz;
			const int dcx = sx * nyz, dcy = sy * nz;
			/* A triangle can only be hit inside this cell where its box meets the piece of the ray inside the cell (a hit
			 * elsewhere is found in that other cell - the pieces tile the ray), so most of a crowded cell's list is
			 * rejected on six compares.  pad: rounding of O + t r, same form as the margin the boxes were grown by. */
			const float pad = 1e-4f * hmax + 1e-6f * fmaxf(fabsf(Ox) + t_out * fabsf(rx), fmaxf(fabsf(Oy) + t_out * fabsf(ry), fabsf(Oz) + t_out * fabsf(rz)));
			const float qsx = 65535.0f * u.inv_h[0], qsy = 65535.0f * u.inv_h[1], qsz = 65535.0f * u.inv_h[2];
			float t_prev = t_in;
			/* Most cells on a ray's way are empty, so a step would be one dependent load after another: walk LOOK cells
			 * ahead with the (cheap) DDA arithmetic, fetch their records together, then visit them in order. */
			constexpr int LOOK = UG_REFL_LOOK;
			bool finished = false;
			for (int guard = 0; guard < 3200 && !finished; guard++) {
				u32 cpk[LOOK];
				int cells[LOOK];
				uint4 cr[LOOK];
				float texit[LOOK];
				int nstep = 0;
				bool left = false;
#pragma unroll
				for (int j = 0; j < LOOK; j++) {
					cpk[j] = 0; texit[j] = 0.0f; cells[j] = 0;
					if (!left) {
						cpk[j] = (u32)cx | ((u32)cy << 10) | ((u32)cz << 20);
						cells[j] = cell;
						nstep = j + 1;
						const bool ax0 = (tmx <= tmy) && (tmx <= tmz), ax1 = !ax0 && (tmy <= tmz), ax2 = !ax0 && !ax1;
						texit[j] = ax0 ? tmx : (ax1 ? tmy : tmz);
						tmx += ax0 ? tdx : 0.0f; tmy += ax1 ? tdy : 0.0f; tmz += ax2 ? tdz : 0.0f;
						cx += ax0 ? sx : 0; cy += ax1 ? sy : 0; cz += ax2 ? sz : 0;
						cell += ax0 ? dcx : (ax1 ? dcy : sz);
						const int moved = ax0 ? sx : (ax1 ? sy : sz);
						left = (moved == 0) | ((u32)cx >= (u32)u.n[0]) | ((u32)cy >= (u32)u.n[1]) | ((u32)cz >= (u32)nz);
					}
				}
#pragma unroll
				for (int j = 0; j < LOOK; j++) cr[j] = (j < nstep) ? a.cellrec[cells[j]] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
				for (int j = 0; j < LOOK; j++) {
					if (j >= nstep || finished) continue;
					if (cr[j].y) {
						const int jx = (int)(cpk[j] & 1023u), jy = (int)((cpk[j] >> 10) & 1023u), jz = (int)(cpk[j] >> 20);
						const float clx = u.lo[0] + u.h[0] * (float)jx, cly = u.lo[1] + u.h[1] * (float)jy, clz = u.lo[2] + u.h[2] * (float)jz;
						const float ax_ = Ox + t_prev * rx, bx_ = Ox + texit[j] * rx, ay_ = Oy + t_prev * ry, by_ = Oy + texit[j] * ry, az_ = Oz + t_prev * rz, bz_ = Oz + texit[j] * rz;
						/* the ray piece in cell units, rounded outwards */
						const u32 sxl = (u32)floorf(q16(fminf(ax_, bx_) - pad, clx, qsx)), sxh = (u32)ceilf(q16(fmaxf(ax_, bx_) + pad, clx, qsx));
						const u32 syl = (u32)floorf(q16(fminf(ay_, by_) - pad, cly, qsy)), syh = (u32)ceilf(q16(fmaxf(ay_, by_) + pad, cly, qsy));
						const u32 szl = (u32)floorf(q16(fminf(az_, bz_) - pad, clz, qsz)), szh = (u32)ceilf(q16(fmaxf(az_, bz_) + pad, clz, qsz));
						/* the whole cell first (10-bit union box) */
						const u32 ulo = cr[j].z, uhi = cr[j].w;
						const bool cell_out = ((ulo & 1023u) > ((sxh + 63u) >> 6)) | ((uhi & 1023u) < (sxl >> 6)) | (((ulo >> 10) & 1023u) > ((syh + 63u) >> 6)) |
						                      (((uhi >> 10) & 1023u) < (syl >> 6)) | ((ulo >> 20) > ((szh + 63u) >> 6)) | ((uhi >> 20) < (szl >> 6));
						if (!cell_out) {
							const uint4 *rec = a.prec + cr[j].x;
							const u32 n = cr[j].y;
							uint4 cur = rec[0];
							for (u32 q = 0; q < n; q++) {
								const uint4 c = cur;
								if (q + 1 < n) cur = rec[q + 1]; /* next record in flight during this one's test */
								const bool out = ((c.y & 0xffffu) > sxh) | ((c.y >> 16) < sxl) | ((c.z & 0xffffu) > syh) | ((c.z >> 16) < syl) | ((c.w & 0xffffu) > szh) | ((c.w >> 16) < szl);
								if (out) continue;
								const int tri = (int)c.x;
								float tt;
								if (!intersect_full(a.geo[3 * (size_t)tri], a.geo[3 * (size_t)tri + 1], a.geo[3 * (size_t)tri + 2], Ox, Oy, Oz, rx, ry, rz, tt)) continue;
								if (tt > 0.0f && (tt < best || (tt == best && tri < hit))) { best = tt; hit = tri; }
							}
						}
					}
					t_prev = texit[j];
					/* clearly before this cell's exit: nothing in a later cell can win */
					if (hit >= 0 && best < texit[j] * 0.999f - 1e-4f * hmin) finished = true;
				}
				if (left) finished = true;
			}
		}
		const float O[3] = {Ox, Oy, Oz}, r[3] = {rx, ry, rz};
		/* spec: this bounce's term of the pixel sum (added in bounce order by k_refl_combine), then spawn the next ray */
		bool live = false;
		if (hit >= 0) {
			const float Px = O[0] + best * r[0], Py = O[1] + best * r[1], Pz = O[2] + best * r[2];
			const float4 e1 = a.geo[3 * (size_t)hit + 1], e2 = a.geo[3 * (size_t)hit + 2];
			const int m = a.mat_idx[hit];
			const bool mat_ok = m >= 0 && m < a.mat_count;
			u32 c8 = 0;
			if (mat_ok) c8 = ug_lambert_rgb8(R, Lvx, Lvy, Lvz, Px, Py, Pz, e1, e2, a.mat_list[m * UG_MATERIAL_SIZE + 3], a.mat_list[m * UG_MATERIAL_SIZE + 4],
			                                 a.mat_list[m * UG_MATERIAL_SIZE + 5]);
			const float local[3] = {(float)(c8 & 255u), (float)((c8 >> 8) & 255u), (float)((c8 >> 16) & 255u)};
			const float k = (bounce < a.bounces && mat_ok) ? a.mat_reflect[m] : 0.0f;
			const float w = a.w[pix];
			float *term = a.term + 3 * ((size_t)(bounce - 1) * (size_t)a.N + (size_t)pix);
			a.depth[pix] = (u32)bounce;
			if (!(k > 0)) {
				for (int c = 0; c < 3; c++) term[c] = w * local[c];
			} else {
				const float wk = w * (1.0f - k);
				for (int c = 0; c < 3; c++) term[c] = wk * local[c];
				a.w[pix] = w * k;
				float O2[3], r2[3];
				reflect_ray(e1, e2, Px, Py, Pz, r[0], r[1], r[2], O2, r2);
				for (int c = 0; c < 3; c++) { a.O[3 * (size_t)pix + c] = O2[c]; a.r[3 * (size_t)pix + c] = r2[c]; }
				live = true;
			}
		}
		ug_append(live, (u32)pix, a.next, a.counts + bounce);
	}
}

/* spec: pixel = (uchar)(((1-k0)*local0 + term_1) + term_2 ...) for every pixel that spawned a reflection ray; local0 is
 * the shaded (shadowed) colour, which is why this pass comes after k_shade while the bounces themselves do not */
__global__ void __launch_bounds__(256) k_refl_combine(const ReflectArgs a, const u32 *__restrict__ list0)
{
	ug_pdl_enter<UG_PDL_REFLECT>();
	const u32 n = a.counts[0];
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const size_t pix = list0[i];
		const float k = a.k0[pix];
		const u32 depth = a.depth[pix];
		float acc[3];
		for (int c = 0; c < 3; c++) acc[c] = (1.0f - k) * (float)a.rgb[3 * pix + c];
		for (u32 b = 0; b < depth; b++)
			for (int c = 0; c < 3; c++) acc[c] = acc[c] + a.term[3 * ((size_t)b * (size_t)a.N + pix) + c];
		for (int c = 0; c < 3; c++) {
			const unsigned char v = (unsigned char)acc[c];
			a.rgb[3 * pix + c] = v;
			if (a.rgb_remote) a.rgb_remote[3 * pix + c] = v;
		}
	}
}

/* host side ------------------------------------------------------------------------------------------ */
size_t ugk_ugrid_const_bytes(void) { return sizeof(UGridConst); }

/* the world-space uniform grid of the current vertices: independent of camera and lights, so the frame builds it on
 * its own branch of the graph */
void ugk_reflect_grid(cudaStream_t s, ug_context *ctx, int *launches)
{
	ReflectState &rf = ctx->refl;
	const int sms = ctx->num_sms, T = ctx->num_faces, V = ctx->num_vertices;
	GridBuffers &g = rf.grid;
	if (ctx->use_fused_ugrid) { /* one cooperative launch (k_ugridfused.cu); the chain below is the fallback */
		if (ugk_ugrid_fused(s, ctx)) {
			if (launches) *launches += 1;
			return;
		}
		cudaGetLastError();
		ctx->use_fused_ugrid = false;
	}
	k_aabb_reset<<<1, 32, 0, s>>>(rf.box);
	k_aabb<<<min((V + 255) / 256, sms * 4), 256, 0, s>>>(ctx->d_vertexlist, V, rf.box);
	k_ugrid_setup<<<1, 32, 0, s>>>(rf.box, T, g.C, ctx->refl_cells_per_tri, rf.uc);
	k_ugrid_count<<<min((T + 255) / 256, sms * 8), 256, 0, s>>>(rf.uc, ctx->d_facelist, ctx->d_vertexlist, T, rf.geo, rf.tbox, rf.rect, g.count);
	ugk_exclusive_scan(s, g.count, g.scan, (size_t)T, g.sort.scan_state, g.sort.scan_tiles, sms);
	cudaMemsetAsync(g.span, 0, sizeof(u32) * (size_t)g.C, s);
	cudaMemsetAsync(g.keys[1], 0, sizeof(u32) * (size_t)g.C, s); /* fill cursors (the second key buffer is free: nothing is sorted) */
	{
		const size_t want = (g.cap + 511) / 512;
		k_ugrid_emit<<<(unsigned)(want < (size_t)sms * 16 ? (want ? want : 1) : (size_t)sms * 16), 256, 0, s>>>(rf.uc, g.scan, rf.rect, T, (u32)g.cap, g.keys[0],
		                                                                                                    g.vals[0], g.span, g.counters);
	}
	int n = 6;
	g.final_buf = 0;
	ugk_exclusive_scan(s, g.span, g.offset, (size_t)g.C, g.sort.scan_state, g.sort.scan_tiles, sms);
	k_ugrid_fill<<<sms * 8, 256, 0, s>>>(rf.uc, g.counters, g.keys[0], g.vals[0], rf.tbox, g.offset, g.keys[1], rf.prec);
	k_ugrid_cells<<<sms * 8, 256, 0, s>>>(rf.uc, g.span, g.offset, rf.prec, rf.cellrec);
	n += 3;
	if (launches) *launches += n;
}

static ReflectArgs reflect_args(ug_context *ctx, int bounces)
{
	ReflectState &rf = ctx->refl;
	const GridBuffers &g = rf.grid;
	ReflectArgs a;
	a.fc = ctx->d_fc; a.uc = rf.uc; a.geo = rf.geo; a.prec = rf.prec; a.cellrec = rf.cellrec; a.span = g.span; a.offset = g.offset; a.vals = g.vals[g.final_buf];
	a.mat_idx = ctx->d_materiallist_index; a.mat_list = ctx->d_materiallist; a.mat_reflect = ctx->d_reflect; a.mat_count = ctx->num_materials;
	a.N = ctx->N; a.bounces = bounces; a.t = ctx->d_t; a.dir = ctx->d_dir; a.id = ctx->d_id; a.rgb = ctx->d_rgb; a.rgb_remote = ctx->rgb_remote;
	a.O = rf.O; a.r = rf.r; a.w = rf.w; a.k0 = rf.k0; a.term = rf.term; a.depth = rf.depth; a.list = rf.list0; a.next = rf.list0;
	a.counts = rf.ray_counts;
	return a;
}

/* the bounces themselves: they need the primary hits only (not the shadow bits or the shaded colour) */
void ugk_reflect_rays(cudaStream_t s, ug_context *ctx, int bounces, int *launches)
{
	ReflectState &rf = ctx->refl;
	const int sms = ctx->num_sms, N = ctx->N;
	int n = 0;
	ReflectArgs a = reflect_args(ctx, bounces);
	const int pb = min((N + 255) / 256, sms * 8);
	cudaMemsetAsync(rf.ray_counts, 0, sizeof(u32) * (UG_MAX_BOUNCES + 1), s);
	a.next = rf.list0; /* the rays of bounce 1 are the pixels the combine pass rewrites: their list is kept */
	k_refl_init<<<pb, 256, 0, s>>>(a);
	n += 1;
	for (int b = 1; b <= bounces; b++) {
		a.list = (b == 1) ? rf.list0 : rf.list[b & 1];
		a.next = rf.list[(b + 1) & 1];
		ug_launch(UG_PDL_REFLECT, k_refl_trace, dim3(min((N + 127) / 128, sms * 16)), dim3(128), s, a, b);
		n += 1;
	}
	if (launches) *launches += n;
}

void ugk_reflect_combine(cudaStream_t s, ug_context *ctx, int *launches)
{
	const ReflectArgs a = reflect_args(ctx, 0);
	ug_launch(UG_PDL_REFLECT, k_refl_combine, dim3(min((ctx->N + 255) / 256, ctx->num_sms * 8)), dim3(256), s, a, ctx->refl.list0);
	if (launches) *launches += 1;
}
