/* k_ugrid_dev.cuh -- device helpers of the world-space uniform grid of row R, shared by the kernel chain in k_reflect.cu
 * and the fused cooperative build in k_ugridfused.cu (both must produce the same grid). */
#pragma once
#include "k_grid_dev.cuh"

struct UGridConst {
	float lo[3], h[3], inv_h[3];
	int n[3];
	int cells;
};

static __device__ __forceinline__ u32 f2ord(float f) { const u32 u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
static __device__ __forceinline__ float ord2f(u32 u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

/* resolution from the scene's bounding box: roughly cubic cells, cells_per_tri cells per triangle, capped by the arena */
static __device__ void ugrid_resolution(const u32 *box, int T, int max_cells, float cells_per_tri, UGridConst *uc)
{
	float lo[3], ext[3];
	for (int k = 0; k < 3; k++) {
		lo[k] = ord2f(box[k]);
		const float hi = ord2f(box[3 + k]);
		const float pad = 1e-3f * fmaxf(hi - lo[k], 1e-3f);
		lo[k] -= pad;
		ext[k] = (hi + pad) - lo[k];
	}
	/* a flat scene far from the origin loses the pad to rounding (extent 0), NaN / inf vertices give a useless box: every axis
	 * gets a positive, finite extent - the grid only has to be conservative, any resolution gives the same hits */
	{
		float m = 0.0f;
		for (int k = 0; k < 3; k++) if (ext[k] == ext[k] && ext[k] < 3.0e38f && ext[k] > m) m = ext[k];
		if (!(m > 0.0f)) m = 1.0f;
		for (int k = 0; k < 3; k++) {
			if (!(lo[k] == lo[k]) || fabsf(lo[k]) > 3.0e38f) lo[k] = 0.0f;
			if (!(ext[k] >= 1e-3f * m) || !(ext[k] < 3.0e38f)) ext[k] = fmaxf(1e-3f * m, 1e-30f);
		}
	}
	float target = fminf(cells_per_tri * (float)T, (float)max_cells);
	target = fmaxf(target, 8.0f);
	float hcell = cbrtf(ext[0] * ext[1] * ext[2] / target);
	int n[3];
	for (int iter = 0; iter < 64; iter++) {
		long long prod = 1;
		for (int k = 0; k < 3; k++) { n[k] = (int)ceilf(ext[k] / hcell); n[k] = max(1, min(n[k], 1024)); prod *= n[k]; }
		if (prod <= max_cells) break;
		hcell *= 1.1f;
	}
	while ((long long)n[0] * n[1] * n[2] > (long long)max_cells) { /* whatever the loop left: the cell arena is a hard bound */
		const int k = (n[0] >= n[1] && n[0] >= n[2]) ? 0 : (n[1] >= n[2] ? 1 : 2);
		if (n[k] <= 1) break;
		n[k] = (n[k] + 1) / 2;
	}
	for (int k = 0; k < 3; k++) {
		uc->lo[k] = lo[k];
		uc->n[k] = n[k];
		uc->h[k] = ext[k] / (float)n[k];
		uc->inv_h[k] = (float)n[k] / ext[k];
	}
	uc->cells = n[0] * n[1] * n[2];
}

static __device__ __forceinline__ int ucell(const UGridConst &u, int k, float x) { return ug_clampi((int)floorf((x - u.lo[k]) * u.inv_h[k]), 0, u.n[k] - 1); }

/* per triangle: (v0, e1, e2) with the staging arithmetic of light_kernel.cu:133-144, the grown box, its cell box; returns the pair count */
static __device__ __forceinline__ u32 ugrid_tri_setup(const UGridConst &u, const int *__restrict__ faces, const float *__restrict__ verts, int t,
                                                      float4 *__restrict__ geo, float4 *__restrict__ tbox, uint4 *__restrict__ rect)
{
	const float *p0 = verts + 3 * (size_t)faces[3 * t], *p1 = verts + 3 * (size_t)faces[3 * t + 1], *p2 = verts + 3 * (size_t)faces[3 * t + 2];
	const float ax = p0[0], ay = p0[1], az = p0[2];
	geo[3 * (size_t)t + 0] = make_float4(ax, ay, az, 0.0f);
	geo[3 * (size_t)t + 1] = make_float4(p1[0] - ax, p1[1] - ay, p1[2] - az, 0.0f);
	geo[3 * (size_t)t + 2] = make_float4(p2[0] - ax, p2[1] - ay, p2[2] - az, 0.0f);
	u32 r[3], c = 1;
	float blo[3], bhi[3];
	for (int k = 0; k < 3; k++) {
		const float mn = fminf(p0[k], fminf(p1[k], p2[k])), mx = fmaxf(p0[k], fmaxf(p1[k], p2[k]));
		const float eps = 1e-4f * u.h[k] + 1e-6f * fmaxf(fabsf(mn), fabsf(mx)); /* conservative margin */
		blo[k] = mn - eps; bhi[k] = mx + eps;
		const int i0 = ucell(u, k, blo[k]), i1 = ucell(u, k, bhi[k]);
		r[k] = (u32)i0 | ((u32)i1 << 16);
		c *= (u32)(i1 - i0 + 1);
	}
	tbox[2 * (size_t)t] = make_float4(blo[0], blo[1], blo[2], 0.0f); /* the box that was binned: the DDA culls against it */
	tbox[2 * (size_t)t + 1] = make_float4(bhi[0], bhi[1], bhi[2], 0.0f);
	rect[t] = make_uint4(r[0], r[1], r[2], c);
	return c;
}

/* cell id of the local-th cell of a triangle's cell box */
static __device__ __forceinline__ u32 ugrid_pair_key(const uint4 r, u32 local, int ny, int nz)
{
	const u32 sy = (r.y >> 16) - (r.y & 0xffffu) + 1, sz = (r.z >> 16) - (r.z & 0xffffu) + 1;
	const u32 iz = local % sz; local /= sz;
	const u32 iy = local % sy, ix = local / sy;
	return (((r.x & 0xffffu) + ix) * (u32)ny + ((r.y & 0xffffu) + iy)) * (u32)nz + ((r.z & 0xffffu) + iz);
}

/* ---- per-pair records.  The DDA's inner loop used to chase vals[p] -> box[tri] -> geo[tri]; the box of the triangle
 * CLIPPED TO THE CELL, in 16-bit cell-relative units rounded outwards, fits beside the triangle id in one 16-byte record
 * at the pair's own position, so a candidate costs one independent LDG.128.  q16 is the same monotone float expression
 * on both sides (build: box corner, traversal: ray piece), so "boxes overlap" in floats implies "overlap" in units. */
static __device__ __forceinline__ float q16(float x, float cell_lo, float scale) { return fminf(fmaxf((x - cell_lo) * scale, 0.0f), 65535.0f); }

static __device__ __forceinline__ uint4 ugrid_pair_record(const UGridConst &u, u32 cell, u32 tri, const float4 *__restrict__ tbox)
{
	const u32 nyz = (u32)u.n[1] * (u32)u.n[2];
	const u32 cx = cell / nyz, rem = cell - cx * nyz, cy = rem / (u32)u.n[2], cz = rem - cy * (u32)u.n[2];
	const float4 lo = tbox[2 * (size_t)tri], hi = tbox[2 * (size_t)tri + 1];
	const float clx = u.lo[0] + u.h[0] * (float)cx, cly = u.lo[1] + u.h[1] * (float)cy, clz = u.lo[2] + u.h[2] * (float)cz;
	const float sx = 65535.0f * u.inv_h[0], sy = 65535.0f * u.inv_h[1], sz = 65535.0f * u.inv_h[2];
	uint4 rec;
	rec.x = tri;
	rec.y = (u32)floorf(q16(lo.x, clx, sx)) | ((u32)ceilf(q16(hi.x, clx, sx)) << 16);
	rec.z = (u32)floorf(q16(lo.y, cly, sy)) | ((u32)ceilf(q16(hi.y, cly, sy)) << 16);
	rec.w = (u32)floorf(q16(lo.z, clz, sz)) | ((u32)ceilf(q16(hi.z, clz, sz)) << 16);
	return rec;
}

/* ---- per-cell records: (offset, span) plus the union of the cell's pair boxes in 10-bit cell units (outwards again), so
 * that a ray piece that misses everything in a crowded cell - e.g. the layer of cells holding a finely tessellated floor,
 * crossed by every ray that leaves that floor - rejects the whole list at once.  x = offset, y = span,
 * z = lo.x | lo.y << 10 | lo.z << 20, w = the same for hi. */
static __device__ __forceinline__ uint4 ugrid_cell_record(u32 off, u32 n, const uint4 *__restrict__ prec)
{
	u32 lo[3] = {1023u, 1023u, 1023u}, hi[3] = {0u, 0u, 0u};
	for (u32 q = 0; q < n; q++) {
		const uint4 r = prec[off + q];
		const u32 w[3] = {r.y, r.z, r.w};
#pragma unroll
		for (int k = 0; k < 3; k++) {
			lo[k] = min(lo[k], (w[k] & 0xffffu) >> 6);
			hi[k] = max(hi[k], min(((w[k] >> 16) + 63u) >> 6, 1023u));
		}
	}
	return make_uint4(off, n, lo[0] | (lo[1] << 10) | (lo[2] << 20), hi[0] | (hi[1] << 10) | (hi[2] << 20));
}
