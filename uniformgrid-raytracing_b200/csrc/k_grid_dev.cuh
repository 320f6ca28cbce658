/* k_grid_dev.cuh -- device functions of the perspective-grid build shared by the multi-kernel path
 * (k_grid.cu, k_sort.cu) and the fused persistent kernel (k_gridfused.cu). */
#ifndef UG_K_GRID_DEV_CUH
#define UG_K_GRID_DEV_CUH

#include "k_common.cuh"

/* grid_kernel.cu:132-146 - the exact ternaries (NaN behaviour differs from fminf/fmaxf) */
static __device__ __forceinline__ float ug_min3(float a, float b, float c) { return (a < b) ? ((a < c) ? a : c) : ((b < c) ? b : c); }
static __device__ __forceinline__ float ug_max3(float a, float b, float c) { return (a > b) ? ((a > c) ? a : c) : ((b > c) ? b : c); }
static __device__ __forceinline__ int ug_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* grid_kernel.cu:13-36 getTransformedVertex: modelview, divide by w, projection, divide by w */
static __device__ __forceinline__ void ug_project(const float *cc, float x, float y, float z, float &nx, float &ny, float &nz)
{
	const float *MV = cc + 16, *P = cc + 32;
	float a = ug_row(MV, 0, x, y, z, 1.0f), b = ug_row(MV, 1, x, y, z, 1.0f), c = ug_row(MV, 2, x, y, z, 1.0f), w = ug_row(MV, 3, x, y, z, 1.0f);
	a = a / w; b = b / w; c = c / w;
	const float qx = ug_row(P, 0, a, b, c, 1.0f), qy = ug_row(P, 1, a, b, c, 1.0f), qz = ug_row(P, 2, a, b, c, 1.0f), qw = ug_row(P, 3, a, b, c, 1.0f);
	nx = qx / qw; ny = qy / qw; nz = qz / qw;
}

struct GridGeo { int nbx, nby, gy_lo, gy_hi; float ox, oy, oz; int mode; float xM, yM; };
static __device__ __forceinline__ GridGeo ug_grid_geo(const ViewConst *vc, const float *s_cc)
{
	const int gy_lo = vc->gy_lo;
	const GridGeo g = {vc->nbx, vc->nby, gy_lo, gy_lo + vc->gy_cnt - 1, s_cc[0], s_cc[1], s_cc[2], vc->grid_mode, vc->xM, vc->yM};
	return g;
}

/* ---- a14/a15: the spherical ("effective") light grid, grid_kernel.cu:395-421 getEffective_x and :452-479 getEffective_y.
 * `v` = normalised direction from the light to the point; cc[16..] = the light camera's modelview, whose rows are
 * right (16,20,24), up (17,21,25) and -forward (18,22,26). */
static __device__ __forceinline__ int ug_effective_x(const float *cc, float vx, float vy, float vz, float maxa, int nbx)
{
	const float upd = vx * cc[17] + vy * cc[21] + vz * cc[25];
	float tx = vx - upd * cc[17], ty = vy - upd * cc[21], tz = vz - upd * cc[25];
	const float val = sqrtf(tx * tx + ty * ty + tz * tz); /* getMagnitude: ((0+x*x)+y*y)+z*z, sqrt -> sqrtf */
	tx /= val; ty /= val; tz /= val;
	const float fwd = tx * cc[18] + ty * cc[22] + tz * cc[26];
	const float angle = ug_acosf(fwd);
	const float rightd = tx * cc[16] + ty * cc[20] + tz * cc[24];
	const u32 half = (u32)(nbx / 2), k = (u32)ug_f2i((angle / maxa) * (float)(nbx / 2)); /* int + (int)float, stored unsigned, read back as int */
	return (int)(rightd > 0.0f ? half + k : half - k);
}
static __device__ __forceinline__ int ug_effective_y(const float *cc, float vx, float vy, float vz, float maxa, int nby)
{
	const float rd = vx * cc[16] + vy * cc[20] + vz * cc[24];
	float tx = vx - rd * cc[16], ty = vy - rd * cc[20], tz = vz - rd * cc[24];
	const float val = sqrtf(tx * tx + ty * ty + tz * tz);
	tx /= val; ty /= val; tz /= val;
	const float upd = tx * cc[17] + ty * cc[21] + tz * cc[25];
	const float fwd = tx * cc[18] + ty * cc[22] * tz * cc[26]; /* grid_kernel.cu:468: `a + b*c*d` as written */
	const float angle = ug_acosf(fwd);
	const float half = (float)(nby / 2), part = (angle / maxa) * half; /* :475-478: int + float -> float -> unsigned int */
	return (int)ug_f2u(upd > 0.0f ? half + part : half - part);
}
/* cell of one vertex, grid_kernel.cu:500-527: direction from the light, divided (not multiplied) by its length */
static __device__ __forceinline__ float ug_sph_vertex(const float *cc, const GridGeo &g, float x, float y, float z, int &bx, int &by)
{
	float px = x - cc[0], py = y - cc[1], pz = z - cc[2];
	const float radius = sqrtf(px * px + py * py + pz * pz);
	px /= radius; py /= radius; pz /= radius;
	bx = ug_effective_x(cc, px, py, pz, g.xM, g.nbx);
	by = ug_effective_y(cc, px, py, pz, g.yM, g.nby);
	return radius;
}
static __device__ __forceinline__ int ug_imin3(int a, int b, int c) { return (a < b) ? ((a < c) ? a : c) : ((b < c) ? b : c); }
static __device__ __forceinline__ int ug_imax3(int a, int b, int c) { return (a > b) ? ((a > c) ? a : c) : ((b > c) ? b : c); }
/* grid_kernel.cu:481-659 DS_spherical_Kernel: integer min/max of the three vertex cells, clamped (:615-625); the
 * wrap branch of size_x/size_y (:627-639) is algebraically gmax-gmin+1.  Kept out of line: the perspective path's
 * register budget (the fused cooperative build) must not pay for it. */
static __device__ __noinline__ void ug_sph_rect(const float *cc, const GridGeo &g, const float *p0, const float *p1, const float *p2, int &gx0, int &gx1, int &gy0,
                                               int &gy1, float &zmin)
{
	int x0, y0, x1, y1, x2, y2;
	const float r0 = ug_sph_vertex(cc, g, p0[0], p0[1], p0[2], x0, y0);
	const float r1 = ug_sph_vertex(cc, g, p1[0], p1[1], p1[2], x1, y1);
	const float r2 = ug_sph_vertex(cc, g, p2[0], p2[1], p2[2], x2, y2);
	zmin = ug_min3(r0, r1, r2); /* grid_kernel.cu:613: projCoordZ = the nearest vertex's distance */
	gx0 = ug_clampi(ug_imin3(x0, x1, x2), 0, g.nbx - 1);
	gy0 = ug_clampi(ug_imin3(y0, y1, y2), 0, g.nby - 1);
	gx1 = ug_clampi(ug_imax3(x0, x1, x2), 0, g.nbx - 1);
	gy1 = ug_clampi(ug_imax3(y0, y1, y2), 0, g.nby - 1);
}

/* One triangle of one view: the ray-independent Moeller-Trumbore record (trace_kernel.cu:158-174 /
 * light_kernel.cu:133-148 staging arithmetic + the hoisted qvec and dot(e2,qvec)), its clamped cell
 * rectangle (grid_kernel.cu:180-229) clipped to the owned cell rows, and its pair count (:231). */
static __device__ __forceinline__ u32 ug_tri_setup(const float *s_cc, const GridGeo &g, const int *__restrict__ faces, const float *__restrict__ verts, int t,
                                                   float4 *__restrict__ rec, uint2 *__restrict__ rect, float *zmin_out = nullptr)
{
	const int i0 = faces[3 * t], i1 = faces[3 * t + 1], i2 = faces[3 * t + 2];
	const float *p0 = verts + 3 * (size_t)i0, *p1 = verts + 3 * (size_t)i1, *p2 = verts + 3 * (size_t)i2;
	const float ax = p0[0], ay = p0[1], az = p0[2];
	const float bx = p1[0], by = p1[1], bz = p1[2];
	const float cx = p2[0], cy = p2[1], cz = p2[2];
	const float e1x = bx - ax, e1y = by - ay, e1z = bz - az;
	const float e2x = cx - ax, e2y = cy - ay, e2z = cz - az;
	const float tx = g.ox - ax, ty = g.oy - ay, tz = g.oz - az;
	float qx, qy, qz;
	ug_cross(qx, qy, qz, tx, ty, tz, e1x, e1y, e1z);       /* qvec = tvec x e1 */
	const float e2q = ug_dot(e2x, e2y, e2z, qx, qy, qz);   /* numerator of t */
	rec[4 * (size_t)t + 0] = make_float4(tx, ty, tz, e2q);
	rec[4 * (size_t)t + 1] = make_float4(e1x, e1y, e1z, 0.0f);
	rec[4 * (size_t)t + 2] = make_float4(e2x, e2y, e2z, 0.0f);
	rec[4 * (size_t)t + 3] = make_float4(qx, qy, qz, 0.0f);
	int gx0, gy0, gx1, gy1;
	float zmin;
	if (g.mode == UG_LIGHT_GRID_SPHERICAL) {
		ug_sph_rect(s_cc, g, p0, p1, p2, gx0, gx1, gy0, gy1, zmin);
	} else {
		float n0x, n0y, n0z, n1x, n1y, n1z, n2x, n2y, n2z;
		ug_project(s_cc, ax, ay, az, n0x, n0y, n0z);
		ug_project(s_cc, bx, by, bz, n1x, n1y, n1z);
		ug_project(s_cc, cx, cy, cz, n2x, n2y, n2z);
		const float xmin = ug_min3(n0x, n1x, n2x), ymin = ug_min3(n0y, n1y, n2y);
		const float xmax = ug_max3(n0x, n1x, n2x), ymax = ug_max3(n0y, n1y, n2y);
		gx0 = ug_clampi(ug_f2i(floorf(((xmin + 1.0f) / 2.0f) * (float)g.nbx)), 0, g.nbx - 1);
		gy0 = ug_clampi(ug_f2i(floorf(((ymin + 1.0f) / 2.0f) * (float)g.nby)), 0, g.nby - 1);
		gx1 = ug_clampi(ug_f2i(floorf(((xmax + 1.0f) / 2.0f) * (float)g.nbx)), 0, g.nbx - 1);
		gy1 = ug_clampi(ug_f2i(floorf(((ymax + 1.0f) / 2.0f) * (float)g.nby)), 0, g.nby - 1);
		zmin = ug_min3(n0z, n1z, n2z); /* grid_kernel.cu:203,211 projCoordZ */
	}
	if (zmin_out) *zmin_out = zmin;
	/* image-strip sharding: keep only the owned cell rows; the rectangle itself is the global one */
	const int ry0 = max(gy0, g.gy_lo), ry1 = min(gy1, g.gy_hi);
	const int sx = gx1 - gx0 + 1, sy = ry1 - ry0 + 1;
	const u32 c = (sx > 0 && sy > 0) ? (u32)(sx * sy) : 0u; /* NaN vertices can invert the rectangle: no cells */
	rect[t] = make_uint2((u32)gx0 | ((u32)gx1 << 16), (u32)ry0 | ((u32)(c ? ry1 : ry0) << 16));
	return c;
}

/* ---- depth slabs (NUM_SLABS > 1).  frustum_grid.h:227-235 reduces projCoordZ on the host:
 *   zMin = min(init, min{z : z >= 0}),  zMax = max(init, max z)   (NaN never wins a comparison)
 * here as atomics on zrange[0] (bits of a non-negative float order like the float) and zrange[1] (order-preserving code). */
static __device__ __forceinline__ u32 ug_fenc(float f) { const u32 b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
static __device__ __forceinline__ float ug_fdec(u32 e) { return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e); }
static __device__ __forceinline__ void ug_zrange_add(u32 *zrange, float z)
{
	u32 lo = 0xffffffffu, hi = 0u;
	if (z >= 0.0f) lo = __float_as_uint(z + 0.0f); /* -0 -> +0 */
	if (z == z) hi = ug_fenc(z + 0.0f);
	lo = __reduce_min_sync(UG_FULL_MASK, lo);
	hi = __reduce_max_sync(UG_FULL_MASK, hi);
	if (ug_lane() == 0) {
		if (lo != 0xffffffffu) atomicMin(&zrange[0], lo);
		if (hi != 0u) atomicMax(&zrange[1], hi);
	}
}
/* grid_kernel.cu:334-352 SlabKernel + the clamp of DSFillkernel (:300-311); a triangle SlabKernel skips (z < 0 or NaN) keeps
 * whatever d_zSlabs held - never initialised in the reference, 0 here */
static __device__ __forceinline__ u32 ug_slab_of(float p, const u32 *__restrict__ zrange, int S)
{
	if (!(p >= 0.0f)) return 0u;
	const float zMin = __uint_as_float(zrange[0]), zMax = ug_fdec(zrange[1]);
	u32 bin = ug_f2u((float)S * ((p - zMin) / (zMax - zMin)));
	if (bin >= (u32)S) bin = (u32)S - 1u;
	return (u32)ug_clampi((int)bin, 0, S - 1);
}

/* first index in [lo,hi) with a[idx] > x */
static __device__ __forceinline__ int ug_upper_bound(const u32 *__restrict__ a, int lo, int hi, u32 x)
{
	while (lo < hi) {
		const int mid = (lo + hi) >> 1;
		if (a[mid] > x) hi = mid; else lo = mid + 1;
	}
	return lo;
}

/* pair p of triangle t (grid_kernel.cu:318-325: i over x, j over y inside the rectangle) -> local cell key */
static __device__ __forceinline__ u32 ug_pair_key(uint2 r, u32 local, int gy_lo, int gy_cnt)
{
	const u32 gx0 = r.x & 0xffffu, ry0 = r.y & 0xffffu, ry1 = r.y >> 16;
	const u32 sy = ry1 - ry0 + 1;
	const u32 i = local / sy, j = local - i * sy;
	return (gx0 + i) * (u32)gy_cnt + (ry0 - (u32)gy_lo + j);
}

/* ---- scan building blocks of the persistent (cooperative) kernels */
/* Decoupled look-back: tile `tile` publishes `total`, returns the sum of all earlier tiles.  Called by
 * warp 0 of the block.  All tiles < `tile` are either finished or held by a resident block. */
/* phase time stamps for profiling (ns, low 32 bits of %globaltimer), written by block 0 into counters[8..] */
static __device__ __forceinline__ void gf_stamp(u32 *counters, int k)
{
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		unsigned long long t;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
		counters[8 + k] = (u32)t;
	}
}

static __device__ __forceinline__ u32 ug_lookback(volatile u64 *desc, u32 tile, u32 total)
{
	const u32 lane = ug_lane();
	u64 prefix = 0; /* sums saturate at 2^32 - 1 instead of wrapping: a saturated pair total exceeds every arena and raises the overflow flag */
	if (tile == 0) {
		if (lane == 0) desc[0] = (2ull << 32) | total;
		return 0;
	}
	if (lane == 0) desc[tile] = (1ull << 32) | total;
	int look = (int)tile - 1;
	while (true) {
		const int idx = look - (int)lane;
		u64 d = 2ull << 32; /* before tile 0: inclusive prefix 0 */
		if (idx >= 0) {
			do { d = desc[idx]; } while ((d >> 32) == 0);
		}
		const u32 flag = (u32)(d >> 32), val = (u32)d;
		const u32 incl = __ballot_sync(UG_FULL_MASK, flag == 2);
		u64 part = val;
		if (incl) part = ((int)lane <= __ffs(incl) - 1) ? val : 0u;
		for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(UG_FULL_MASK, part, o);
		prefix += part;
		if (incl) break;
		look -= 32;
	}
	if (prefix > 0xffffffffull) prefix = 0xffffffffull;
	if (lane == 0) desc[tile] = (2ull << 32) | (prefix + total > 0xffffffffull ? 0xffffffffull : prefix + total);
	return (u32)prefix;
}

/* block-wide exclusive scan of one value per thread; returns the thread's exclusive prefix, total in `total` */
template <int NT>
static __device__ __forceinline__ u32 ug_block_scan(u32 v, u32 *s_warp, u32 &total)
{
	u32 incl = v;
	for (int o = 1; o < 32; o <<= 1) {
		const u32 x = __shfl_up_sync(UG_FULL_MASK, incl, o);
		if ((int)ug_lane() >= o) incl += x;
	}
	__syncthreads();
	if (ug_lane() == 31) s_warp[threadIdx.x >> 5] = incl;
	__syncthreads();
	u32 off = 0;
	total = 0;
#pragma unroll
	for (int w = 0; w < NT / 32; w++) {
		const u32 c = s_warp[w];
		if (w < (int)(threadIdx.x >> 5)) off += c;
		total += c;
	}
	return off + incl - v;
}

/* ---- radix sort building blocks (k_sort.cu, k_gridfused.cu); NT = threads per block, ROUNDS = keys per thread and sub-tile */
#define RS_THREADS 256
#define RS_ROUNDS 8

static __device__ __forceinline__ void rs_chunk(u32 n, u32 G, u32 b, u32 &lo, u32 &hi)
{
	u32 chunk = (n + G - 1) / G;
	chunk = (chunk + 31u) & ~31u; /* warp rounds stay 128-byte aligned */
	const u64 l = (u64)chunk * b, h = l + chunk;
	lo = (u32)(l < n ? l : n);
	hi = (u32)(h < n ? h : n);
}

/* digit histogram of keys[lo,hi) into s_hist[256] (zeroed by the caller); warp-aggregated because
 * neighbouring pairs mostly carry the same cell id */
template <int NT>
static __device__ __forceinline__ void rs_block_hist(const u32 *__restrict__ keys, u32 lo, u32 hi, int shift, u32 *s_hist)
{
	for (u32 base = lo; base < hi; base += NT) {
		const u32 i = base + threadIdx.x;
		const bool valid = i < hi;
		const u32 d = valid ? ((keys[i] >> shift) & 255u) : 256u;
		const u32 peers = __match_any_sync(UG_FULL_MASK, d);
		if (valid && (peers & ((1u << ug_lane()) - 1u)) == 0) atomicAdd(&s_hist[d], __popc(peers));
	}
}

/* stable scatter of the block's chunk [lo,hi): s_run[d] holds where the block's next key of digit d
 * goes; s_cnt is (NT/32) x 257 scratch.  Order inside the block = (sub-tile, warp, round, lane) = input order. */
template <int NT, int ROUNDS>
static __device__ __forceinline__ void rs_block_scatter(const u32 *__restrict__ keys_in, const u32 *__restrict__ vals_in, u32 *__restrict__ keys_out,
                                                        u32 *__restrict__ vals_out, u32 lo, u32 hi, int shift, u32 (*s_cnt)[257], u32 *s_run)
{
	constexpr int WARPS = NT / 32;
	const u32 warp = threadIdx.x >> 5, lane = ug_lane(), lt = (1u << lane) - 1u;
	for (u32 tile = lo; tile < hi; tile += NT * ROUNDS) {
		for (u32 i = threadIdx.x; i < WARPS * 257; i += NT) (&s_cnt[0][0])[i] = 0;
		__syncthreads();
		u32 key[ROUNDS], rank[ROUNDS];
#pragma unroll
		for (int r = 0; r < ROUNDS; r++) { /* warp w owns keys [tile + w*32*ROUNDS, +32*ROUNDS): round r covers 32 consecutive keys */
			const u32 i = tile + warp * (32 * ROUNDS) + r * 32 + lane;
			const bool valid = i < hi;
			key[r] = valid ? keys_in[i] : 0xffffffffu;
			const u32 d = valid ? ((key[r] >> shift) & 255u) : 256u;
			const u32 peers = __match_any_sync(UG_FULL_MASK, d);
			const u32 prev = s_cnt[warp][d];
			__syncwarp();
			if ((peers & lt) == 0) s_cnt[warp][d] = prev + __popc(peers);
			__syncwarp();
			rank[r] = prev + __popc(peers & lt);
		}
		__syncthreads();
		for (u32 d = threadIdx.x; d < 256; d += NT) { /* per-warp counts -> per-warp start positions; advance the block's run */
			u32 run = s_run[d];
#pragma unroll
			for (int w = 0; w < WARPS; w++) {
				const u32 c = s_cnt[w][d];
				s_cnt[w][d] = run;
				run += c;
			}
			s_run[d] = run;
		}
		__syncthreads();
#pragma unroll
		for (int r = 0; r < ROUNDS; r++) {
			const u32 i = tile + warp * (32 * ROUNDS) + r * 32 + lane;
			if (i < hi) {
				const u32 d = (key[r] >> shift) & 255u;
				const u32 pos = s_cnt[warp][d] + rank[r];
				keys_out[pos] = key[r];
				vals_out[pos] = vals_in[i];
			}
		}
		__syncthreads();
	}
}

#endif
