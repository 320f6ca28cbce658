/* k_grid.cu -- perspective ("frustum") grid build, one call chain per view (the eye, every light):
 *
 *   k_setup_bin   a2/a3  grid_kernel.cu:13-36,164-243  per triangle: ray-independent Moeller-Trumbore
 *                        record for this view's origin + NDC bounding box -> clamped cell rectangle
 *                        -> pair count (d_sizeList)
 *   exclusive scan a4    frustum_grid.h:249            (k_sort.cu, replaces cudppScan)
 *   k_emit_pairs  a5     grid_kernel.cu:245-332        one thread per OUTPUT pair (load-balanced; the
 *                        reference lets one thread write up to NBX*NBY pairs), key = gx*NBY+gy,
 *                        emission order = triangle id, then x-major inside the rectangle; also the
 *                        per-cell histogram `span` (a7, misc_kernel.cu:36-60)
 *   radix sort    a6     frustum_grid.h:298            (k_sort.cu, replaces cudppSort; stable)
 *   exclusive scan a7    frustum_grid.h:361            offset over ALL cells, empty ones included
 *
 * HBM-bound streaming kernels: coalesced where the reference layout allows (faces/vertices are packed
 * 12-byte AoS, scene.h:19-23), 16-byte stores for everything this library lays out itself.
 */
#include "k_common.cuh"

/* grid_kernel.cu:132-146 - the exact ternaries (NaN behaviour differs from fminf/fmaxf) */
static __device__ __forceinline__ float min3(float a, float b, float c) { return (a < b) ? ((a < c) ? a : c) : ((b < c) ? b : c); }
static __device__ __forceinline__ float max3(float a, float b, float c) { return (a > b) ? ((a > c) ? a : c) : ((b > c) ? b : c); }

/* grid_kernel.cu:13-36 getTransformedVertex */
static __device__ __forceinline__ void project(const float *cc, float x, float y, float z, float &nx, float &ny, float &nz)
{
	const float *MV = cc + 16, *P = cc + 32;
	float a = ug_row(MV, 0, x, y, z, 1.0f), b = ug_row(MV, 1, x, y, z, 1.0f), c = ug_row(MV, 2, x, y, z, 1.0f), w = ug_row(MV, 3, x, y, z, 1.0f);
	a = a / w; b = b / w; c = c / w;
	const float qx = ug_row(P, 0, a, b, c, 1.0f), qy = ug_row(P, 1, a, b, c, 1.0f), qz = ug_row(P, 2, a, b, c, 1.0f), qw = ug_row(P, 3, a, b, c, 1.0f);
	nx = qx / qw; ny = qy / qw; nz = qz / qw;
}

static __device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(256) k_setup_bin(const ViewConst *__restrict__ vcp, const int *__restrict__ faces, const float *__restrict__ verts,
                                                  int T, float4 *__restrict__ rec, uint2 *__restrict__ rect, u32 *__restrict__ count)
{
	__shared__ float s_cc[64];
	__shared__ int s_geo[4];
	if (threadIdx.x < 64) s_cc[threadIdx.x] = vcp->cc[threadIdx.x];
	if (threadIdx.x == 64) { s_geo[0] = vcp->nbx; s_geo[1] = vcp->nby; s_geo[2] = vcp->gy_lo; s_geo[3] = vcp->gy_cnt; }
	__syncthreads();
	const int nbx = s_geo[0], nby = s_geo[1], gy_lo = s_geo[2], gy_hi = s_geo[2] + s_geo[3] - 1;
	const float ox = s_cc[0], oy = s_cc[1], oz = s_cc[2];
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
		const int i0 = faces[3 * t], i1 = faces[3 * t + 1], i2 = faces[3 * t + 2];
		const float *p0 = verts + 3 * (size_t)i0, *p1 = verts + 3 * (size_t)i1, *p2 = verts + 3 * (size_t)i2;
		const float ax = p0[0], ay = p0[1], az = p0[2];
		const float bx = p1[0], by = p1[1], bz = p1[2];
		const float cx = p2[0], cy = p2[1], cz = p2[2];
		/* trace_kernel.cu:158-174 / light_kernel.cu:133-148 staging arithmetic */
		const float e1x = bx - ax, e1y = by - ay, e1z = bz - az;
		const float e2x = cx - ax, e2y = cy - ay, e2z = cz - az;
		const float tx = ox - ax, ty = oy - ay, tz = oz - az;
		float qx, qy, qz;
		ug_cross(qx, qy, qz, tx, ty, tz, e1x, e1y, e1z);       /* qvec = tvec x e1 */
		const float e2q = ug_dot(e2x, e2y, e2z, qx, qy, qz);   /* numerator of t */
		rec[4 * (size_t)t + 0] = make_float4(tx, ty, tz, e2q);
		rec[4 * (size_t)t + 1] = make_float4(e1x, e1y, e1z, 0.0f);
		rec[4 * (size_t)t + 2] = make_float4(e2x, e2y, e2z, 0.0f);
		rec[4 * (size_t)t + 3] = make_float4(qx, qy, qz, 0.0f);
		/* grid_kernel.cu:180-231 */
		float n0x, n0y, n0z, n1x, n1y, n1z, n2x, n2y, n2z;
		project(s_cc, ax, ay, az, n0x, n0y, n0z);
		project(s_cc, bx, by, bz, n1x, n1y, n1z);
		project(s_cc, cx, cy, cz, n2x, n2y, n2z);
		const float xmin = min3(n0x, n1x, n2x), ymin = min3(n0y, n1y, n2y);
		const float xmax = max3(n0x, n1x, n2x), ymax = max3(n0y, n1y, n2y);
		const int gx0 = clampi(ug_f2i(floorf(((xmin + 1.0f) / 2.0f) * (float)nbx)), 0, nbx - 1);
		const int gy0 = clampi(ug_f2i(floorf(((ymin + 1.0f) / 2.0f) * (float)nby)), 0, nby - 1);
		const int gx1 = clampi(ug_f2i(floorf(((xmax + 1.0f) / 2.0f) * (float)nbx)), 0, nbx - 1);
		const int gy1 = clampi(ug_f2i(floorf(((ymax + 1.0f) / 2.0f) * (float)nby)), 0, nby - 1);
		/* image-strip sharding: keep only the owned cell rows; the rectangle itself is the global one */
		const int ry0 = max(gy0, gy_lo), ry1 = min(gy1, gy_hi);
		const int sx = gx1 - gx0 + 1, sy = ry1 - ry0 + 1;
		const u32 c = (sx > 0 && sy > 0) ? (u32)(sx * sy) : 0u; /* NaN vertices can invert the rectangle: no cells */
		rect[t] = make_uint2((u32)gx0 | ((u32)gx1 << 16), (u32)ry0 | ((u32)(c ? ry1 : ry0) << 16));
		count[t] = c;
	}
}

void ugk_setup_bin(cudaStream_t s, const ViewConst *vc, const int *faces, const float *verts, int T, GridBuffers &g, int sms)
{
	if (T <= 0) return;
	const int blocks = min((T + 255) / 256, sms * 8);
	k_setup_bin<<<blocks, 256, 0, s>>>(vc, faces, verts, T, g.tri_rec, g.rect, g.count);
}

void ugk_scan_counts(cudaStream_t s, GridBuffers &g, int sms)
{
	ugk_exclusive_scan(s, g.count, g.scan, (size_t)g.T, g.sort.scan_state, g.sort.scan_tiles, sms);
}

/* one thread per output pair; the owning triangle is found by binary search in the exclusive scan,
 * narrowed first to the block's own triangle range */
#define EMIT_THREADS 256
#define EMIT_PER_BLOCK (EMIT_THREADS * 8)

static __device__ __forceinline__ int upper_bound(const u32 *__restrict__ a, int lo, int hi, u32 x)
{ /* first index in [lo,hi) with a[idx] > x */
	while (lo < hi) {
		const int mid = (lo + hi) >> 1;
		if (a[mid] > x) hi = mid; else lo = mid + 1;
	}
	return lo;
}

__global__ void __launch_bounds__(EMIT_THREADS) k_emit_pairs(const ViewConst *__restrict__ vcp, const u32 *__restrict__ scan, const uint2 *__restrict__ rect,
                                                            int T, u32 cap, u32 *__restrict__ keys, u32 *__restrict__ vals, u32 *__restrict__ span,
                                                            u32 *__restrict__ counters)
{
	__shared__ int s_range[2];
	const u32 Ptrue = scan[T];
	const u32 P = min(Ptrue, cap);
	const int gy_lo = vcp->gy_lo, gy_cnt = vcp->gy_cnt;
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		counters[0] = P;
		counters[1] = (Ptrue > cap) ? 1u : 0u; /* arena overflow: the host grows it and re-runs the frame */
		counters[2] = Ptrue;
	}
	for (u32 blk = blockIdx.x * EMIT_PER_BLOCK; blk < P; blk += gridDim.x * EMIT_PER_BLOCK) {
		const u32 blk_end = min(blk + EMIT_PER_BLOCK, P);
		__syncthreads();
		if (threadIdx.x < 2) {
			const u32 x = threadIdx.x == 0 ? blk : blk_end - 1;
			s_range[threadIdx.x] = upper_bound(scan, 0, T + 1, x) - 1;
		}
		__syncthreads();
		const int tlo = s_range[0], thi = s_range[1];
		for (u32 p = blk + threadIdx.x; p < blk_end; p += EMIT_THREADS) {
			const int t = upper_bound(scan, tlo, thi + 1, p) - 1;
			const uint2 r = rect[t];
			const u32 local = p - scan[t];
			const u32 gx0 = r.x & 0xffffu, ry0 = r.y & 0xffffu, ry1 = r.y >> 16;
			const u32 sy = ry1 - ry0 + 1;
			const u32 i = local / sy, j = local - i * sy;      /* grid_kernel.cu:318-325: i over x, j over y */
			const u32 key = (gx0 + i) * (u32)gy_cnt + (ry0 - (u32)gy_lo + j);
			keys[p] = key;
			vals[p] = (u32)t;
			atomicAdd(&span[key], 1u);
		}
	}
}

void ugk_emit_pairs(cudaStream_t s, const ViewConst *vc, GridBuffers &g, int sms)
{
	cudaMemsetAsync(g.span, 0, sizeof(u32) * (size_t)g.C, s);
	const size_t want = (g.cap + EMIT_PER_BLOCK - 1) / EMIT_PER_BLOCK;
	const unsigned blocks = (unsigned)(want < (size_t)sms * 8 ? (want ? want : 1) : (size_t)sms * 8);
	k_emit_pairs<<<blocks, EMIT_THREADS, 0, s>>>(vc, g.scan, g.rect, g.T, (u32)g.cap, g.keys[0], g.vals[0], g.span, g.counters);
}

void ugk_cell_offsets(cudaStream_t s, GridBuffers &g, int sms)
{
	ugk_exclusive_scan(s, g.span, g.offset, (size_t)g.C, g.sort.scan_state, g.sort.scan_tiles, sms);
}
