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
#include "k_grid_dev.cuh"

#ifndef UG_REFL_LOOK
#define UG_REFL_LOOK 2            /* DDA look-ahead (cells whose spans are fetched together) */
#endif
#ifndef UG_REFL_CELLS_PER_TRI
#define UG_REFL_CELLS_PER_TRI 4.0f /* uniform-grid resolution: cells per triangle */
#endif

struct UGridConst {
	float lo[3], h[3], inv_h[3];
	int n[3];
	int cells;
};

/* ---------------------------------------------------------------- bounding box + resolution */
static __device__ __forceinline__ u32 f2ord(float f) { const u32 u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
static __device__ __forceinline__ float ord2f(u32 u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

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

__global__ void k_ugrid_setup(const u32 *box, int T, int max_cells, UGridConst *uc)
{
	if (threadIdx.x || blockIdx.x) return;
	float lo[3], ext[3];
	for (int k = 0; k < 3; k++) {
		lo[k] = ord2f(box[k]);
		const float hi = ord2f(box[3 + k]);
		const float pad = 1e-3f * fmaxf(hi - lo[k], 1e-3f);
		lo[k] -= pad;
		ext[k] = (hi + pad) - lo[k];
	}
	/* roughly cubic cells, about one cell per triangle (coarser cells = fewer empty DDA steps), capped by the arena */
	float target = fminf(UG_REFL_CELLS_PER_TRI * (float)T, (float)max_cells);
	target = fmaxf(target, 8.0f);
	float hcell = cbrtf(ext[0] * ext[1] * ext[2] / target);
	int n[3];
	for (int iter = 0; iter < 64; iter++) {
		long long prod = 1;
		for (int k = 0; k < 3; k++) { n[k] = (int)ceilf(ext[k] / hcell); n[k] = max(1, min(n[k], 1024)); prod *= n[k]; }
		if (prod <= max_cells) break;
		hcell *= 1.1f;
	}
	for (int k = 0; k < 3; k++) {
		uc->lo[k] = lo[k];
		uc->n[k] = n[k];
		uc->h[k] = ext[k] / (float)n[k];
		uc->inv_h[k] = (float)n[k] / ext[k];
	}
	uc->cells = n[0] * n[1] * n[2];
}

/* ---------------------------------------------------------------- grid build */
static __device__ __forceinline__ int ucell(const UGridConst &u, int k, float x) { return ug_clampi((int)floorf((x - u.lo[k]) * u.inv_h[k]), 0, u.n[k] - 1); }

__global__ void __launch_bounds__(256) k_ugrid_count(const UGridConst *__restrict__ ucp, const int *__restrict__ faces, const float *__restrict__ verts, int T,
                                                    float4 *__restrict__ geo, uint4 *__restrict__ rect, u32 *__restrict__ count)
{
	const UGridConst u = *ucp;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
		const float *p0 = verts + 3 * (size_t)faces[3 * t], *p1 = verts + 3 * (size_t)faces[3 * t + 1], *p2 = verts + 3 * (size_t)faces[3 * t + 2];
		const float ax = p0[0], ay = p0[1], az = p0[2];
		/* (v0, e1, e2) with the staging arithmetic of light_kernel.cu:133-144 */
		geo[3 * (size_t)t + 0] = make_float4(ax, ay, az, 0.0f);
		geo[3 * (size_t)t + 1] = make_float4(p1[0] - ax, p1[1] - ay, p1[2] - az, 0.0f);
		geo[3 * (size_t)t + 2] = make_float4(p2[0] - ax, p2[1] - ay, p2[2] - az, 0.0f);
		u32 r[3], c = 1;
		for (int k = 0; k < 3; k++) {
			const float mn = fminf(p0[k], fminf(p1[k], p2[k])), mx = fmaxf(p0[k], fmaxf(p1[k], p2[k]));
			const float eps = 1e-4f * u.h[k] + 1e-6f * fmaxf(fabsf(mn), fabsf(mx)); /* conservative margin */
			const int i0 = ucell(u, k, mn - eps), i1 = ucell(u, k, mx + eps);
			r[k] = (u32)i0 | ((u32)i1 << 16);
			c *= (u32)(i1 - i0 + 1);
		}
		rect[t] = make_uint4(r[0], r[1], r[2], c);
		count[t] = c;
	}
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
			const uint4 r = rect[t];
			u32 local = p - scan[t];
			const u32 sy = (r.y >> 16) - (r.y & 0xffffu) + 1, sz = (r.z >> 16) - (r.z & 0xffffu) + 1;
			const u32 iz = local % sz; local /= sz;
			const u32 iy = local % sy, ix = local / sy;
			const u32 key = (((r.x & 0xffffu) + ix) * (u32)ny + ((r.y & 0xffffu) + iy)) * (u32)nz + ((r.z & 0xffffu) + iz);
			keys[p] = key;
			vals[p] = (u32)t;
			atomicAdd(&span[key], 1u);
		}
	}
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
	const u32 *span, *offset, *vals;
	const int *mat_idx;
	const float *mat_list, *mat_reflect;
	int mat_count, N, bounces;
	const float *t, *dir;
	const int *id;
	unsigned char *rgb;
	float *O, *r, *acc, *w;      /* per-pixel ray state [3N],[3N],[3N],[N] */
	u32 *live;                   /* [N] flags */
	const u32 *live_scan;        /* [N+1] */
	u32 *list;                   /* compacted pixel ids */
};

/* bounce 0: which pixels spawn a reflection ray */
__global__ void __launch_bounds__(256) k_refl_init(const ReflectArgs a)
{
	const float ex = a.fc->cam.cc[0], ey = a.fc->cam.cc[1], ez = a.fc->cam.cc[2];
	for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < a.N; pix += gridDim.x * blockDim.x) {
		u32 live = 0;
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
					a.acc[3 * (size_t)pix + c] = (1.0f - k) * (float)a.rgb[3 * (size_t)pix + c];
				}
				a.w[pix] = k;
				live = 1;
			}
		}
		a.live[pix] = live;
	}
}

__global__ void __launch_bounds__(256) k_refl_compact(const u32 *__restrict__ live, const u32 *__restrict__ scan, u32 *__restrict__ list, int N)
{
	for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < N; pix += gridDim.x * blockDim.x)
		if (live[pix]) list[scan[pix]] = (u32)pix;
}

/* one bounce for every live ray: 3-D DDA through the uniform grid, shade, spawn the next ray or finish the pixel */
__global__ void __launch_bounds__(128) k_refl_trace(const ReflectArgs a, int bounce)
{
	const UGridConst u = *a.uc;
	const float *R = a.fc->shade_mv;
	const float l0 = a.fc->light_pos[0], l1 = a.fc->light_pos[1], l2 = a.fc->light_pos[2];
	const float Lvx = R[0] * l0 + R[4] * l1 + R[8] * l2, Lvy = R[1] * l0 + R[5] * l1 + R[9] * l2, Lvz = R[2] * l0 + R[6] * l1 + R[10] * l2;
	const u32 nlive = a.live_scan[a.N];
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nlive; i += gridDim.x * blockDim.x) {
		const int pix = (int)a.list[i];
		float O[3], r[3];
		for (int c = 0; c < 3; c++) { O[c] = a.O[3 * (size_t)pix + c]; r[c] = a.r[3 * (size_t)pix + c]; }
		/* entry into the grid's box (the origin sits on a surface, normally inside) */
		float t_in = 0.0f, t_out = 3.0e38f;
		bool miss_box = false;
		for (int k = 0; k < 3; k++) {
			const float hi = u.lo[k] + u.h[k] * (float)u.n[k];
			if (r[k] != 0.0f) {
				float t0 = (u.lo[k] - O[k]) / r[k], t1 = (hi - O[k]) / r[k];
				if (t0 > t1) { const float s = t0; t0 = t1; t1 = s; }
				t_in = fmaxf(t_in, t0); t_out = fminf(t_out, t1);
			} else if (O[k] < u.lo[k] || O[k] > hi) miss_box = true;
		}
		float best = 99999999.9f;
		int hit = -1;
		if (!miss_box && t_in <= t_out) {
			int ci[3], step[3];
			float tmax[3], tdelta[3];
			for (int k = 0; k < 3; k++) {
				const float s = O[k] + t_in * r[k];
				ci[k] = ucell(u, k, s);
				step[k] = (r[k] > 0.0f) ? 1 : ((r[k] < 0.0f) ? -1 : 0);
				if (step[k] != 0) {
					const float bound = u.lo[k] + u.h[k] * (float)(ci[k] + (step[k] > 0 ? 1 : 0));
					tmax[k] = (bound - O[k]) / r[k];
					tdelta[k] = u.h[k] / fabsf(r[k]);
				} else { tmax[k] = 3.0e38f; tdelta[k] = 3.0e38f; }
			}
			const float hmin = fminf(u.h[0], fminf(u.h[1], u.h[2]));
			/* Most cells on a ray's way are empty, so a step would be one dependent load of span[cell] after another:
			 * walk LOOK cells ahead with the (cheap) DDA arithmetic, fetch their spans together, then visit them in order. */
			constexpr int LOOK = UG_REFL_LOOK;
			bool finished = false;
			for (int guard = 0; guard < 1100 && !finished; guard++) {
				u32 cells[LOOK], sp[LOOK];
				float texit[LOOK];
				int nstep = 0;
				bool left = false;
#pragma unroll
				for (int j = 0; j < LOOK; j++) {
					cells[j] = 0; texit[j] = 0.0f;
					if (!left) {
						cells[j] = ((u32)ci[0] * (u32)u.n[1] + (u32)ci[1]) * (u32)u.n[2] + (u32)ci[2];
						const int ax = (tmax[0] <= tmax[1]) ? ((tmax[0] <= tmax[2]) ? 0 : 2) : ((tmax[1] <= tmax[2]) ? 1 : 2);
						texit[j] = tmax[ax];
						nstep = j + 1;
						ci[ax] += step[ax];
						if (step[ax] == 0 || ci[ax] < 0 || ci[ax] >= u.n[ax]) left = true;
						tmax[ax] += tdelta[ax];
					}
				}
#pragma unroll
				for (int j = 0; j < LOOK; j++) sp[j] = (j < nstep) ? a.span[cells[j]] : 0u;
#pragma unroll
				for (int j = 0; j < LOOK; j++) {
					if (j >= nstep || finished) continue;
					if (sp[j]) {
						const u32 off = a.offset[cells[j]];
						for (u32 q = 0; q < sp[j]; q++) {
							const int tri = (int)a.vals[off + q];
							float tt;
							if (!intersect_full(a.geo[3 * (size_t)tri], a.geo[3 * (size_t)tri + 1], a.geo[3 * (size_t)tri + 2], O[0], O[1], O[2], r[0], r[1], r[2], tt)) continue;
							if (tt > 0.0f && (tt < best || (tt == best && tri < hit))) { best = tt; hit = tri; }
						}
					}
					/* clearly before this cell's exit: nothing in a later cell can win */
					if (hit >= 0 && best < texit[j] * 0.999f - 1e-4f * hmin) finished = true;
				}
				if (left) finished = true;
			}
		}
		/* spec: accumulate, then finish the pixel or spawn the next ray */
		float acc[3] = {a.acc[3 * (size_t)pix], a.acc[3 * (size_t)pix + 1], a.acc[3 * (size_t)pix + 2]};
		const float w = a.w[pix];
		u32 live = 0;
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
			if (!(k > 0)) {
				for (int c = 0; c < 3; c++) acc[c] = acc[c] + w * local[c];
			} else {
				const float wk = w * (1.0f - k);
				for (int c = 0; c < 3; c++) acc[c] = acc[c] + wk * local[c];
				a.w[pix] = w * k;
				float O2[3], r2[3];
				reflect_ray(e1, e2, Px, Py, Pz, r[0], r[1], r[2], O2, r2);
				for (int c = 0; c < 3; c++) { a.O[3 * (size_t)pix + c] = O2[c]; a.r[3 * (size_t)pix + c] = r2[c]; a.acc[3 * (size_t)pix + c] = acc[c]; }
				live = 1;
			}
		}
		a.live[pix] = live;
		if (!live)
			for (int c = 0; c < 3; c++) a.rgb[3 * (size_t)pix + c] = (unsigned char)acc[c];
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
	k_aabb_reset<<<1, 32, 0, s>>>(rf.box);
	k_aabb<<<min((V + 255) / 256, sms * 4), 256, 0, s>>>(ctx->d_vertexlist, V, rf.box);
	k_ugrid_setup<<<1, 32, 0, s>>>(rf.box, T, g.C, rf.uc);
	k_ugrid_count<<<min((T + 255) / 256, sms * 8), 256, 0, s>>>(rf.uc, ctx->d_facelist, ctx->d_vertexlist, T, rf.geo, rf.rect, g.count);
	ugk_exclusive_scan(s, g.count, g.scan, (size_t)T, g.sort.scan_state, g.sort.scan_tiles, sms);
	cudaMemsetAsync(g.span, 0, sizeof(u32) * (size_t)g.C, s);
	{
		const size_t want = (g.cap + 511) / 512;
		k_ugrid_emit<<<(unsigned)(want < (size_t)sms * 16 ? (want ? want : 1) : (size_t)sms * 16), 256, 0, s>>>(rf.uc, g.scan, rf.rect, T, (u32)g.cap, g.keys[0],
		                                                                                                    g.vals[0], g.span, g.counters);
	}
	int n = 6;
	g.final_buf = ugk_radix_sort_pairs(s, g.keys, g.vals, g.counters, g.cap, g.key_bits, g.sort, sms, &n);
	ugk_exclusive_scan(s, g.span, g.offset, (size_t)g.C, g.sort.scan_state, g.sort.scan_tiles, sms);
	n += 1;
	if (launches) *launches += n;
}

/* the bounces themselves, on top of the shaded frame */
void ugk_reflect_rays(cudaStream_t s, ug_context *ctx, int bounces, int *launches)
{
	ReflectState &rf = ctx->refl;
	const int sms = ctx->num_sms, N = ctx->N;
	const GridBuffers &g = rf.grid;
	int n = 0;
	ReflectArgs a;
	a.fc = ctx->d_fc; a.uc = rf.uc; a.geo = rf.geo; a.span = g.span; a.offset = g.offset; a.vals = g.vals[g.final_buf];
	a.mat_idx = ctx->d_materiallist_index; a.mat_list = ctx->d_materiallist; a.mat_reflect = ctx->d_reflect; a.mat_count = ctx->num_materials;
	a.N = N; a.bounces = bounces; a.t = ctx->d_t; a.dir = ctx->d_dir; a.id = ctx->d_id; a.rgb = ctx->d_rgb;
	a.O = rf.O; a.r = rf.r; a.acc = rf.acc; a.w = rf.w; a.live = rf.live; a.live_scan = rf.live_scan; a.list = rf.list;
	const int pb = min((N + 255) / 256, sms * 8);
	k_refl_init<<<pb, 256, 0, s>>>(a);
	n += 1;
	for (int b = 1; b <= bounces; b++) {
		ugk_exclusive_scan(s, rf.live, rf.live_scan, (size_t)N, rf.scan_state, rf.scan_tiles, sms);  /* compaction by scan */
		cudaMemcpyAsync(rf.ray_counts + (b - 1 < 7 ? b - 1 : 7), rf.live_scan + N, sizeof(u32), cudaMemcpyDeviceToDevice, s); /* live rays of bounce b */
		k_refl_compact<<<pb, 256, 0, s>>>(rf.live, rf.live_scan, rf.list, N);
		k_refl_trace<<<min((N + 127) / 128, sms * 16), 128, 0, s>>>(a, b);
		n += 3;
	}
	if (launches) *launches += n;
}
