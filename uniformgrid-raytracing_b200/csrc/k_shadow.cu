/* k_shadow.cu -- shadow rays through a light-space perspective grid (a13, a16-a18).
 *
 *   k_map_rays      misc_kernel.cu:63-129 mapSortkernel: per pixel, the light cell its shadow ray
 *                   falls into (sentinel NBX*NBY = outside the light frustum), plus the histogram of
 *                   rays per light cell
 *   (scan)          ray ranges per light cell
 *   k_scatter_rays  DecisionData::sort_by_block_indices ... stream_compact (decision_data.h:171-271):
 *                   regroup the rays by light cell.  The shadow bit is an OR over the cell's triangles,
 *                   so the order of rays inside a cell cannot change any output and a counting sort
 *                   with warp-aggregated cursors replaces the reference's 15-bit radix sort +
 *                   segmented scan + compaction
 *   k_shadow        light_kernel.cu:52-270 mod_light_rckernel: packets of <= 64 rays of ONE light cell
 *                   against that cell's triangle list (ray-independent records staged in shared
 *                   memory), any-hit with warp-ballot / block early termination
 */
#include "k_common.cuh"

#define SH_BATCH 64

/* shadow ray of one pixel: P = eye + t*dir (misses have t = -1 and still shoot, misc_kernel.cu:74-81),
 * L = normalize(P - light) */
static __device__ __forceinline__ void shadow_ray(const float *s_lat, const ViewConst &cam, float lx, float ly, float lz, int px, int py, float tv,
                                                  float &Px, float &Py, float &Pz, float &Lx, float &Ly, float &Lz)
{
	float dx, dy, dz;
	const float ex = cam.cc[0], ey = cam.cc[1], ez = cam.cc[2];
	ug_primary_dir(s_lat, ex, ey, ez, px, py, cam.W, cam.H, dx, dy, dz);
	Px = ex + tv * dx; Py = ey + tv * dy; Pz = ez + tv * dz;
	Lx = Px - lx; Ly = Py - ly; Lz = Pz - lz;
	ug_normalize(Lx, Ly, Lz);
}

__global__ void __launch_bounds__(256) k_map_rays(const FrameConst *__restrict__ fc, int light, const float *__restrict__ t, u32 *__restrict__ light_cell,
                                                 u32 *__restrict__ ray_span, int N)
{
	__shared__ float s_lat[100];
	__shared__ float s_cc[64];
	const ViewConst &cam = fc->cam;
	const ViewConst &lv = fc->light[light];
	for (int i = threadIdx.x; i < 100; i += blockDim.x) s_lat[i] = cam.lattice[i];
	if (threadIdx.x < 64) s_cc[threadIdx.x] = lv.cc[threadIdx.x];
	__syncthreads();
	const int W = cam.W, nbx = lv.nbx, nby = lv.nby;
	const float nearP = lv.near_plane, tanh_ = lv.tan_half;
	const float lx = s_cc[0], ly = s_cc[1], lz = s_cc[2];
	/* misc_kernel.cu:97-101: un-normalised near plane = MVP row 3 + row 2 */
	const float p0 = s_cc[48 + 3] + s_cc[48 + 2], p1 = s_cc[48 + 7] + s_cc[48 + 6], p2 = s_cc[48 + 11] + s_cc[48 + 10], p3 = s_cc[48 + 15] + s_cc[48 + 14];
	const float num = ug_dot(p0, p1, p2, lx, ly, lz) + p3;
	const float half = nearP * tanh_;                                                  /* float */
	const double den = __dmul_rn(__dmul_rn(2.0, (double)nearP), (double)tanh_);        /* 2.0 * near * tan: double */
	const int row0 = cam.gy_lo * UG_TILE;
	for (int base = blockIdx.x * blockDim.x; base < N; base += gridDim.x * blockDim.x) {
		const int pix = base + threadIdx.x;
		const bool valid = pix < N;
		u32 cell = 0xffffffffu;
		if (valid) {
			const int pyl = pix / W, px = pix - pyl * W;
			float Px, Py, Pz, Lx, Ly, Lz;
			shadow_ray(s_lat, cam, lx, ly, lz, px, pyl + row0, t[pix], Px, Py, Pz, Lx, Ly, Lz);
			const float s = -1.0f * num / ug_dot(p0, p1, p2, Lx, Ly, Lz);
			const float X = s * Lx, Y = s * Ly, Z = s * Lz;
			const float right = X * s_cc[16 + 0] + Y * s_cc[16 + 4] + Z * s_cc[16 + 8];
			const float up = X * s_cc[16 + 1] + Y * s_cc[16 + 5] + Z * s_cc[16 + 9];
			/* misc_kernel.cu:115-116: (float)/(double) -> double, * int, floor, -> int */
			const int blx = ug_d2i(floor(__dmul_rn(__ddiv_rn((double)(right + half), den), (double)nbx)));
			const int bly = ug_d2i(floor(__dmul_rn(__ddiv_rn((double)(up + half), den), (double)nby)));
			cell = (u32)(nbx * nby);
			if (blx >= 0 && blx <= nbx - 1 && bly >= 0 && bly <= nby - 1) cell = (u32)(blx * nby + bly);
			light_cell[pix] = cell;
		}
		/* rays of one warp mostly share a light cell: one atomic per distinct cell */
		const u32 peers = __match_any_sync(UG_FULL_MASK, cell);
		if (valid && (peers & ((1u << ug_lane()) - 1u)) == 0) atomicAdd(&ray_span[cell], (u32)__popc(peers));
	}
}

__global__ void __launch_bounds__(256) k_scatter_rays(const u32 *__restrict__ light_cell, const u32 *__restrict__ ray_offset, u32 *__restrict__ cursor,
                                                     u32 *__restrict__ ray_pix, int N)
{
	for (int base = blockIdx.x * blockDim.x; base < N; base += gridDim.x * blockDim.x) {
		const int pix = base + threadIdx.x;
		const bool valid = pix < N;
		const u32 cell = valid ? light_cell[pix] : 0xffffffffu;
		const u32 peers = __match_any_sync(UG_FULL_MASK, cell);
		const u32 lower = __popc(peers & ((1u << ug_lane()) - 1u));
		u32 start = 0;
		const int leader = __ffs(peers) - 1;
		if (valid && lower == 0) start = atomicAdd(&cursor[cell], (u32)__popc(peers));
		start = __shfl_sync(UG_FULL_MASK, start, leader);
		if (valid) ray_pix[ray_offset[cell] + start + lower] = (u32)pix;
	}
}

/* Work decomposition (same idea as k_trace): the unit of work is 64 CONSECUTIVE rays of the regrouped
 * ray array, not a light cell.  Rays are grouped by cell, so a packet holds a few runs of equal cell;
 * each run walks its cell's triangle list with any-hit early termination.  Lists longer than SH_SEG
 * (border cells of the light grid collect every clamped triangle) are cut into segments: segment 0 is
 * done inline, the others are queued as (ray range, cell, segment) entries and traced by a second launch
 * of the same inner loop, so no packet ever owns more than SH_SEG tests per ray.  The shadow bit is an
 * OR, so segments need no merge step. */
#define SH_SEG 256

struct ShadowRay { float Lx, Ly, Lz, dist_b; int pix; };

/* rays [r0,r1) of the packet against triangles [t0,t1) of one cell's list */
static __device__ __forceinline__ void shadow_run(float4 *s_rec, const float4 *__restrict__ rec, const u32 *__restrict__ value_list, u32 t0, u32 t1,
                                                  bool mine, const ShadowRay &ray, float lx, float ly, float lz, bool &done, int *__restrict__ shadowed)
{
	const int tid = threadIdx.x;
	for (u32 base = t0; base < t1; base += SH_BATCH) {
		const u32 cnt = min((u32)SH_BATCH, t1 - base);
		__syncthreads();
		if ((u32)tid < cnt) {
			const float4 *src = rec + 4 * (size_t)value_list[base + tid];
			s_rec[4 * tid + 0] = src[0];
			s_rec[4 * tid + 1] = src[1];
			s_rec[4 * tid + 2] = src[2];
			s_rec[4 * tid + 3] = src[3];
		}
		__syncthreads();
		if (mine && !done) {
			for (u32 k = 0; k < cnt; k++) {
				float tt;
				if (!ug_intersect(s_rec[4 * k], s_rec[4 * k + 1], s_rec[4 * k + 2], s_rec[4 * k + 3], ray.Lx, ray.Ly, ray.Lz, tt)) continue;
				if (!(tt < 999999.9f) || tt == 0.0f) continue; /* intersectTri returns t only if t < oldt; `if (value)` */
				/* light_kernel.cu:193-202 + isSmaller :1-11 */
				const float qx = lx + tt * ray.Lx, qy = ly + tt * ray.Ly, qz = lz + tt * ray.Lz;
				const float ax = qx - lx, ay = qy - ly, az = qz - lz;
				const float dist_a = sqrtf(ax * ax + ay * ay + az * az);
				if (dist_a + 1e-03f < ray.dist_b) { done = true; shadowed[ray.pix] = 1; break; }
			}
		}
		if (__syncthreads_and(done || !mine)) break; /* every ray of the run is occluded */
	}
}

static __device__ __forceinline__ ShadowRay make_shadow_ray(const float *s_lat, const ViewConst &cam, float lx, float ly, float lz, int pix, const float *__restrict__ t)
{
	ShadowRay r;
	const int W = cam.W, pyl = pix / W, px = pix - pyl * W;
	float Px, Py, Pz; /* light_kernel.cu:164-184 */
	shadow_ray(s_lat, cam, lx, ly, lz, px, pyl + cam.gy_lo * UG_TILE, t[pix], Px, Py, Pz, r.Lx, r.Ly, r.Lz);
	const float bx = Px - lx, by = Py - ly, bz = Pz - lz;
	r.dist_b = sqrtf(bx * bx + by * by + bz * bz); /* isSmaller's distance_b, light_kernel.cu:5 */
	r.pix = pix;
	return r;
}

__global__ void __launch_bounds__(64) k_shadow(const FrameConst *__restrict__ fc, int light, int C, const u32 *__restrict__ span, const u32 *__restrict__ offset,
                                               const u32 *__restrict__ value_list, const float4 *__restrict__ rec, const u32 *__restrict__ ray_pix,
                                               const u32 *__restrict__ light_cell, const u32 *__restrict__ ray_offset, const float *__restrict__ t,
                                               int *__restrict__ shadowed, uint4 *__restrict__ queue, u32 *__restrict__ queue_count, u32 queue_cap)
{
	__shared__ float4 s_rec[SH_BATCH * 4];
	__shared__ float s_lat[100];
	__shared__ u32 s_cell[64];
	__shared__ u32 s_slot;
	const ViewConst &cam = fc->cam;
	const ViewConst &lv = fc->light[light];
	const int tid = threadIdx.x;
	for (int i = tid; i < 100; i += 64) s_lat[i] = cam.lattice[i];
	const float lx = lv.cc[0], ly = lv.cc[1], lz = lv.cc[2];
	const u32 R = ray_offset[C]; /* rays that fall inside the light frustum; the rest stay lit */
	const u32 nitems = (R + 63u) / 64u;
	for (u32 item = blockIdx.x; item < nitems; item += gridDim.x) {
		const u32 first = item * 64u, nact = min(64u, R - first);
		__syncthreads();
		ShadowRay ray = {0.0f, 0.0f, 1.0f, 0.0f, 0};
		if ((u32)tid < nact) {
			const int pix = (int)ray_pix[first + tid];
			s_cell[tid] = light_cell[pix];
			ray = make_shadow_ray(s_lat, cam, lx, ly, lz, pix, t);
		}
		__syncthreads();
		bool done = (u32)tid >= nact;
		u32 r0 = 0;
		while (r0 < nact) {
			const u32 cell = s_cell[r0];
			u32 r1 = r0 + 1;
			while (r1 < nact && s_cell[r1] == cell) r1++;
			const u32 ntri = span[cell], toff = offset[cell];
			const bool mine = (u32)tid >= r0 && (u32)tid < r1;
			if (ntri) {
				const u32 segs = (ntri + SH_SEG - 1) / SH_SEG;
				u32 inline_segs = 1;
				if (segs > 1) { /* queue segments 1.. for the second launch; if the queue is full do them here */
					__syncthreads();
					if (tid == 0) {
						const u32 slot = atomicAdd(queue_count, segs - 1);
						if (slot + (segs - 1) <= queue_cap) {
							for (u32 k = 1; k < segs; k++) queue[slot + k - 1] = make_uint4(first + r0, r1 - r0, cell, k);
							s_slot = 1;
						} else {
							atomicSub(queue_count, segs - 1);
							s_slot = 0;
						}
					}
					__syncthreads();
					if (!s_slot) inline_segs = segs;
				}
				shadow_run(s_rec, rec, value_list, toff, toff + min(ntri, inline_segs * SH_SEG), mine, ray, lx, ly, lz, done, shadowed);
			}
			r0 = r1;
		}
	}
}

__global__ void __launch_bounds__(64) k_shadow_queue(const FrameConst *__restrict__ fc, int light, const u32 *__restrict__ span, const u32 *__restrict__ offset,
                                                     const u32 *__restrict__ value_list, const float4 *__restrict__ rec, const u32 *__restrict__ ray_pix,
                                                     const float *__restrict__ t, int *__restrict__ shadowed, const uint4 *__restrict__ queue,
                                                     const u32 *__restrict__ queue_count, u32 queue_cap)
{
	__shared__ float4 s_rec[SH_BATCH * 4];
	__shared__ float s_lat[100];
	const ViewConst &cam = fc->cam;
	const ViewConst &lv = fc->light[light];
	const int tid = threadIdx.x;
	for (int i = tid; i < 100; i += 64) s_lat[i] = cam.lattice[i];
	const float lx = lv.cc[0], ly = lv.cc[1], lz = lv.cc[2];
	const u32 n = min(*queue_count, queue_cap);
	for (u32 e = blockIdx.x; e < n; e += gridDim.x) {
		const uint4 q = queue[e]; /* (first ray, ray count, cell, segment) */
		__syncthreads();
		ShadowRay ray = {0.0f, 0.0f, 1.0f, 0.0f, 0};
		const bool mine = (u32)tid < q.y;
		bool done = true;
		if (mine) {
			const int pix = (int)ray_pix[q.x + tid];
			done = shadowed[pix] == 1; /* already occluded (by another segment or light): nothing to add to an OR */
			if (!done) ray = make_shadow_ray(s_lat, cam, lx, ly, lz, pix, t);
		}
		const u32 ntri = span[q.z], toff = offset[q.z];
		const u32 t0 = q.w * SH_SEG, t1 = min(ntri, t0 + SH_SEG);
		shadow_run(s_rec, rec, value_list, toff + t0, toff + t1, mine, ray, lx, ly, lz, done, shadowed);
	}
}

void ugk_map_rays(cudaStream_t s, const FrameConst *fc, int light, const float *t, u32 *light_cell, u32 *ray_span, int N, int sms)
{
	if (N <= 0) return;
	const int blocks = min((N + 255) / 256, sms * 8);
	k_map_rays<<<blocks, 256, 0, s>>>(fc, light, t, light_cell, ray_span, N);
}

void ugk_scatter_rays(cudaStream_t s, const u32 *light_cell, const u32 *ray_offset, u32 *cursor, u32 *ray_pix, int N, int sms)
{
	if (N <= 0) return;
	const int blocks = min((N + 255) / 256, sms * 8);
	k_scatter_rays<<<blocks, 256, 0, s>>>(light_cell, ray_offset, cursor, ray_pix, N);
}

void ugk_shadow(cudaStream_t s, const FrameConst *fc, int light, const GridBuffers &g, const LightRays &r, const float *t, int *shadowed, int N, int C, int sms)
{
	if (C <= 0 || N <= 0) return;
	const size_t items = ((size_t)N + 63) / 64;
	const unsigned blocks = (unsigned)(items < (size_t)sms * 32 ? items : (size_t)sms * 32);
	const u32 *vals = g.vals[g.final_buf];
	cudaMemsetAsync(r.queue_count, 0, sizeof(u32), s);
	k_shadow<<<blocks, 64, 0, s>>>(fc, light, C, g.span, g.offset, vals, g.tri_rec, r.pix, r.cell, r.offset, t, shadowed, r.queue, r.queue_count, (u32)r.queue_cap);
	const unsigned qblocks = (unsigned)(r.queue_cap < (size_t)sms * 32 ? r.queue_cap : (size_t)sms * 32);
	k_shadow_queue<<<qblocks, 64, 0, s>>>(fc, light, g.span, g.offset, vals, g.tri_rec, r.pix, t, shadowed, r.queue, r.queue_count, (u32)r.queue_cap);
}
