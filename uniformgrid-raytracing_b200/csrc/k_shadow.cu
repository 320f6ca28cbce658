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

__global__ void __launch_bounds__(256) k_map_rays(const FrameConst *__restrict__ fc, int light, const float *__restrict__ t, const float *__restrict__ dir,
                                                 u32 *__restrict__ light_cell, u32 *__restrict__ ray_span, int N)
{
	__shared__ float s_cc[64];
	const ViewConst &cam = fc->cam;
	const ViewConst &lv = fc->light[light];
	if (threadIdx.x < 64) s_cc[threadIdx.x] = lv.cc[threadIdx.x];
	__syncthreads();
	const int nbx = lv.nbx, nby = lv.nby;
	const float ex = cam.cc[0], ey = cam.cc[1], ez = cam.cc[2];
	const float nearP = lv.near_plane, tanh_ = lv.tan_half;
	const float lx = s_cc[0], ly = s_cc[1], lz = s_cc[2];
	/* misc_kernel.cu:97-101: un-normalised near plane = MVP row 3 + row 2 */
	const float p0 = s_cc[48 + 3] + s_cc[48 + 2], p1 = s_cc[48 + 7] + s_cc[48 + 6], p2 = s_cc[48 + 11] + s_cc[48 + 10], p3 = s_cc[48 + 15] + s_cc[48 + 14];
	const float num = ug_dot(p0, p1, p2, lx, ly, lz) + p3;
	const float half = nearP * tanh_;                                                  /* float */
	const double den = __dmul_rn(__dmul_rn(2.0, (double)nearP), (double)tanh_);        /* 2.0 * near * tan: double */
	for (int base = blockIdx.x * blockDim.x; base < N; base += gridDim.x * blockDim.x) {
		const int pix = base + threadIdx.x;
		const bool valid = pix < N;
		u32 cell = 0xffffffffu;
		if (valid) {
			/* misc_kernel.cu:74-88: P = eye + t*dir (misses have t = -1 and still shoot), L = normalize(P - light) */
			const float tv = t[pix], dx = dir[3 * (size_t)pix], dy = dir[3 * (size_t)pix + 1], dz = dir[3 * (size_t)pix + 2];
			const float Px = ex + tv * dx, Py = ey + tv * dy, Pz = ez + tv * dz;
			float Lx = Px - lx, Ly = Py - ly, Lz = Pz - lz;
			ug_normalize(Lx, Ly, Lz);
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

/* Work decomposition (same idea as k_trace): the unit of work is 32 CONSECUTIVE rays of the regrouped
 * ray array - one warp, no block barriers - not a light cell.  Rays are grouped by cell, so a packet
 * holds a few runs of equal cell; each run walks its cell's triangle list with any-hit early
 * termination.  A run of m < 32 rays would leave lanes idle, so the warp is used as an m x g grid
 * (g = 32/m): lane = (ray slot, triangle group), every group tests every g-th staged triangle for its
 * ray and the hits are OR-ed with a ballot - legal because the shadow bit is an OR over the list.
 * Lists longer than SH_SEG (border cells of the light grid collect every clamped triangle) are cut
 * into segments: segment 0 is done inline, the others are queued as (ray range, cell, segment) entries
 * and traced by a second launch of the same inner loop, so no packet owns more than SH_SEG tests per
 * ray.  Packets are handed out dynamically (atomic ticket) because their cost varies with the lists. */
#define SH_SEG 256
#define SH_WARPS 4

struct ShadowRay { float Lx, Ly, Lz, dist_b; int pix; };

/* light_kernel.cu:164-184: rebuild the shadow ray of one pixel from (t, dir) */
static __device__ __forceinline__ ShadowRay make_shadow_ray(const float *__restrict__ dir, const float *__restrict__ t, int pix, float ex, float ey, float ez,
                                                            float lx, float ly, float lz)
{
	ShadowRay r;
	const float tv = t[pix], dx = dir[3 * (size_t)pix], dy = dir[3 * (size_t)pix + 1], dz = dir[3 * (size_t)pix + 2];
	const float Px = ex + tv * dx, Py = ey + tv * dy, Pz = ez + tv * dz;
	r.Lx = Px - lx; r.Ly = Py - ly; r.Lz = Pz - lz;
	ug_normalize(r.Lx, r.Ly, r.Lz);
	const float bx = Px - lx, by = Py - ly, bz = Pz - lz;
	r.dist_b = sqrtf(bx * bx + by * by + bz * bz); /* isSmaller's distance_b, light_kernel.cu:5 */
	r.pix = pix;
	return r;
}

/* One run: the rays held by lanes [r0, r0+m) against triangles [t0,t1) of one cell's list.
 * `own` is the lane's own ray (valid for lanes of the run), `w_rec` the warp's staging area. */
static __device__ __forceinline__ void shadow_run_warp(float4 *w_rec, const float4 *__restrict__ rec, const u32 *__restrict__ value_list, u32 t0, u32 t1,
                                                       u32 r0, u32 m, const ShadowRay &own, bool own_done, float lx, float ly, float lz,
                                                       int *__restrict__ shadowed)
{
	const u32 lane = ug_lane();
	const u32 g = 32u / m, slot = lane % m, grp = lane / m;
	const bool usable = grp < g;
	const int src = (int)(r0 + slot);
	ShadowRay ray;
	ray.Lx = __shfl_sync(UG_FULL_MASK, own.Lx, src);
	ray.Ly = __shfl_sync(UG_FULL_MASK, own.Ly, src);
	ray.Lz = __shfl_sync(UG_FULL_MASK, own.Lz, src);
	ray.dist_b = __shfl_sync(UG_FULL_MASK, own.dist_b, src);
	ray.pix = __shfl_sync(UG_FULL_MASK, own.pix, src);
	bool done = __shfl_sync(UG_FULL_MASK, (int)own_done, src) != 0;
	u32 slotmask = 0; /* the lanes that work on my ray */
	for (u32 q = 0; q < g; q++) slotmask |= 1u << (slot + q * m);
	for (u32 base = t0; base < t1; base += 32) {
		const u32 cnt = min(32u, t1 - base);
		__syncwarp();
		if (lane < cnt) {
			const float4 *src4 = rec + 4 * (size_t)value_list[base + lane];
			w_rec[4 * lane + 0] = src4[0];
			w_rec[4 * lane + 1] = src4[1];
			w_rec[4 * lane + 2] = src4[2];
			w_rec[4 * lane + 3] = src4[3];
		}
		__syncwarp();
		if (usable && !done) {
			for (u32 k = grp; k < cnt; k += g) {
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
		if (__all_sync(UG_FULL_MASK, done || !usable)) break; /* every ray of the run is occluded */
	}
}

__global__ void __launch_bounds__(32 * SH_WARPS) k_shadow(const FrameConst *__restrict__ fc, int light, int C, const u32 *__restrict__ span,
                                                        const u32 *__restrict__ offset, const u32 *__restrict__ value_list, const float4 *__restrict__ rec,
                                                        const u32 *__restrict__ ray_pix, const u32 *__restrict__ light_cell, const u32 *__restrict__ ray_offset,
                                                        const float *__restrict__ t, const float *__restrict__ dir, int *__restrict__ shadowed,
                                                        uint4 *__restrict__ queue, u32 *__restrict__ queue_count, u32 queue_cap)
{
	__shared__ float4 s_rec[SH_WARPS][32 * 4];
	float4 *w_rec = s_rec[threadIdx.x >> 5];
	const u32 lane = ug_lane();
	const float ex = fc->cam.cc[0], ey = fc->cam.cc[1], ez = fc->cam.cc[2];
	const float lx = fc->light[light].cc[0], ly = fc->light[light].cc[1], lz = fc->light[light].cc[2];
	const u32 R = ray_offset[C]; /* rays that fall inside the light frustum; the rest stay lit */
	const u32 nitems = (R + 31u) / 32u;
	while (true) {
		u32 item = 0;
		if (lane == 0) item = atomicAdd(&queue_count[1], 1u);
		item = __shfl_sync(UG_FULL_MASK, item, 0);
		if (item >= nitems) break;
		const u32 first = item * 32u;
		const bool active = first + lane < R;
		ShadowRay own = {0.0f, 0.0f, 1.0f, 0.0f, 0};
		u32 cell = 0xffffffffu;
		if (active) {
			const int pix = (int)ray_pix[first + lane];
			cell = light_cell[pix];
			own = make_shadow_ray(dir, t, pix, ex, ey, ez, lx, ly, lz);
		}
		u32 remaining = __ballot_sync(UG_FULL_MASK, active);
		while (remaining) {
			const u32 r0 = __ffs(remaining) - 1;
			const u32 c = __shfl_sync(UG_FULL_MASK, cell, (int)r0);
			const u32 members = __ballot_sync(UG_FULL_MASK, active && cell == c);
			const u32 m = __popc(members);
			remaining &= ~members;
			const u32 ntri = span[c], toff = offset[c];
			if (!ntri) continue;
			const u32 segs = (ntri + SH_SEG - 1) / SH_SEG;
			u32 inline_segs = 1;
			if (segs > 1) { /* queue segments 1.. for the second launch; if the queue is full do them here */
				u32 ok = 0;
				if (lane == 0) {
					const u32 slot = atomicAdd(&queue_count[0], segs - 1);
					if (slot + (segs - 1) <= queue_cap) {
						for (u32 k = 1; k < segs; k++) queue[slot + k - 1] = make_uint4(first + r0, m, c, k);
						ok = 1;
					} else {
						atomicSub(&queue_count[0], segs - 1);
					}
				}
				ok = __shfl_sync(UG_FULL_MASK, ok, 0);
				if (!ok) inline_segs = segs;
			}
			shadow_run_warp(w_rec, rec, value_list, toff, toff + min(ntri, inline_segs * SH_SEG), r0, m, own, false, lx, ly, lz, shadowed);
		}
	}
}

__global__ void __launch_bounds__(32 * SH_WARPS) k_shadow_queue(const FrameConst *__restrict__ fc, int light, const u32 *__restrict__ span,
                                                              const u32 *__restrict__ offset, const u32 *__restrict__ value_list,
                                                              const float4 *__restrict__ rec, const u32 *__restrict__ ray_pix, const float *__restrict__ t,
                                                              const float *__restrict__ dir, int *__restrict__ shadowed, const uint4 *__restrict__ queue,
                                                              const u32 *__restrict__ queue_count, u32 queue_cap)
{
	__shared__ float4 s_rec[SH_WARPS][32 * 4];
	float4 *w_rec = s_rec[threadIdx.x >> 5];
	const u32 lane = ug_lane();
	const float ex = fc->cam.cc[0], ey = fc->cam.cc[1], ez = fc->cam.cc[2];
	const float lx = fc->light[light].cc[0], ly = fc->light[light].cc[1], lz = fc->light[light].cc[2];
	const u32 n = min(queue_count[0], queue_cap);
	const u32 warps = gridDim.x * SH_WARPS;
	for (u32 e = blockIdx.x * SH_WARPS + (threadIdx.x >> 5); e < n; e += warps) {
		const uint4 q = queue[e]; /* (first ray, ray count, cell, segment) */
		ShadowRay own = {0.0f, 0.0f, 1.0f, 0.0f, 0};
		bool own_done = true;
		if (lane < q.y) {
			const int pix = (int)ray_pix[q.x + lane];
			own_done = shadowed[pix] == 1; /* already occluded (another segment or light): nothing to add to an OR */
			own = make_shadow_ray(dir, t, pix, ex, ey, ez, lx, ly, lz);
		}
		const u32 ntri = span[q.z], toff = offset[q.z];
		const u32 t0 = q.w * SH_SEG, t1 = min(ntri, t0 + SH_SEG);
		/* the run sits in lanes [0, q.y) here */
		shadow_run_warp(w_rec, rec, value_list, toff + t0, toff + t1, 0, q.y, own, own_done, lx, ly, lz, shadowed);
	}
}

void ugk_map_rays(cudaStream_t s, const FrameConst *fc, int light, const float *t, const float *dir, u32 *light_cell, u32 *ray_span, int N, int sms)
{
	if (N <= 0) return;
	const int blocks = min((N + 255) / 256, sms * 8);
	k_map_rays<<<blocks, 256, 0, s>>>(fc, light, t, dir, light_cell, ray_span, N);
}

void ugk_scatter_rays(cudaStream_t s, const u32 *light_cell, const u32 *ray_offset, u32 *cursor, u32 *ray_pix, int N, int sms)
{
	if (N <= 0) return;
	const int blocks = min((N + 255) / 256, sms * 8);
	k_scatter_rays<<<blocks, 256, 0, s>>>(light_cell, ray_offset, cursor, ray_pix, N);
}

void ugk_shadow(cudaStream_t s, const FrameConst *fc, int light, const GridBuffers &g, const LightRays &r, const float *t, const float *dir, int *shadowed,
                int N, int C, int sms)
{
	if (C <= 0 || N <= 0) return;
	const size_t items = ((size_t)N + 31) / 32, want = (items + SH_WARPS - 1) / SH_WARPS;
	const unsigned blocks = (unsigned)(want < (size_t)sms * 16 ? want : (size_t)sms * 16);
	const u32 *vals = g.vals[g.final_buf];
	cudaMemsetAsync(r.queue_count, 0, 2 * sizeof(u32), s); /* [0] queue length, [1] packet ticket */
	k_shadow<<<blocks, 32 * SH_WARPS, 0, s>>>(fc, light, C, g.span, g.offset, vals, g.tri_rec, r.pix, r.cell, r.offset, t, dir, shadowed, r.queue,
	                                          r.queue_count, (u32)r.queue_cap);
	const size_t qwant = (r.queue_cap + SH_WARPS - 1) / SH_WARPS;
	const unsigned qblocks = (unsigned)(qwant < (size_t)sms * 16 ? qwant : (size_t)sms * 16);
	k_shadow_queue<<<qblocks, 32 * SH_WARPS, 0, s>>>(fc, light, g.span, g.offset, vals, g.tri_rec, r.pix, t, dir, shadowed, r.queue, r.queue_count,
	                                                 (u32)r.queue_cap);
}
