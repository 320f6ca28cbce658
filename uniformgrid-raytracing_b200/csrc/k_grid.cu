/* k_grid.cu -- perspective ("frustum") grid build, multi-kernel path; one call chain per view (the eye,
 * every light).  The fused single-kernel version of the same chain is k_gridfused.cu.
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
#include "k_grid_dev.cuh"

__global__ void __launch_bounds__(256) k_setup_bin(const ViewConst *__restrict__ vcp, const int *__restrict__ faces, const float *__restrict__ verts,
                                                  int T, float4 *__restrict__ rec, uint2 *__restrict__ rect, u32 *__restrict__ count,
                                                  float *__restrict__ projz, u32 *__restrict__ zrange)
{
	__shared__ float s_cc[64];
	if (threadIdx.x < 64) s_cc[threadIdx.x] = vcp->cc[threadIdx.x];
	__syncthreads();
	const GridGeo g = ug_grid_geo(vcp, s_cc);
	if (!projz) { /* NUM_SLABS 1: the committed configuration */
		for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x)
			count[t] = ug_tri_setup(s_cc, g, faces, verts, t, rec, rect);
		return;
	}
	for (int base = blockIdx.x * blockDim.x; base < T; base += gridDim.x * blockDim.x) { /* whole warps: the z range is reduced per warp */
		const int t = base + threadIdx.x;
		float z = __uint_as_float(0x7fc00000u); /* NaN: takes no part */
		if (t < T) {
			count[t] = ug_tri_setup(s_cc, g, faces, verts, t, rec, rect, &z);
			projz[t] = z;
		}
		ug_zrange_add(zrange, z);
	}
}

/* frustum_grid.h:227 / :385: zMin = +2, zMax = -2 (buildGrid) or 9999.9, -9999.9 (buildSphericalGrid) */
__global__ void k_slab_reset(const ViewConst *__restrict__ vcp, u32 *__restrict__ zrange)
{
	const bool sph = vcp->grid_mode == UG_LIGHT_GRID_SPHERICAL;
	zrange[0] = __float_as_uint(sph ? 9999.9f : 2.0f);
	zrange[1] = ug_fenc(sph ? -9999.9f : -2.0f);
}
void ugk_slab_reset(cudaStream_t s, const ViewConst *vc, GridBuffers &g) { k_slab_reset<<<1, 1, 0, s>>>(vc, g.zrange); }

/* the slabs of a cell are neighbours in the sorted list (key = cell*S + slab): one range per cell for the any-hit pass */
__global__ void __launch_bounds__(256) k_slab_union(const u32 *__restrict__ span, const u32 *__restrict__ offset, int cells, int S, u32 *__restrict__ cspan,
                                                   u32 *__restrict__ coffset)
{
	for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += gridDim.x * blockDim.x) {
		const u32 lo = offset[(size_t)c * S], hi = offset[(size_t)c * S + S - 1] + span[(size_t)c * S + S - 1];
		coffset[c] = lo;
		cspan[c] = hi - lo;
	}
}
void ugk_slab_union(cudaStream_t s, GridBuffers &g, int sms)
{
	if (g.S <= 1 || g.cells <= 0) return;
	k_slab_union<<<min((g.cells + 255) / 256, sms * 8), 256, 0, s>>>(g.span, g.offset, g.cells, g.S, g.cspan, g.coffset);
}

void ugk_setup_bin(cudaStream_t s, const ViewConst *vc, const int *faces, const float *verts, int T, GridBuffers &g, int sms)
{
	if (T <= 0) return;
	const int blocks = min((T + 255) / 256, sms * 8);
	k_setup_bin<<<blocks, 256, 0, s>>>(vc, faces, verts, T, g.tri_rec, g.rect, g.count, g.S > 1 ? g.projz : nullptr, g.zrange);
}

void ugk_scan_counts(cudaStream_t s, GridBuffers &g, int sms)
{
	ugk_exclusive_scan(s, g.count, g.scan, (size_t)g.T, g.sort.scan_state, g.sort.scan_tiles, sms);
}

/* one thread per output pair; the owning triangle is found by binary search in the exclusive scan,
 * narrowed first to the block's own triangle range */
#define EMIT_THREADS 256
#define EMIT_PER_BLOCK (EMIT_THREADS * 2)

__global__ void __launch_bounds__(EMIT_THREADS) k_emit_pairs(const ViewConst *__restrict__ vcp, const u32 *__restrict__ scan, const uint2 *__restrict__ rect,
                                                            int T, u32 cap, u32 *__restrict__ keys, u32 *__restrict__ vals, u32 *__restrict__ span,
                                                            u32 *__restrict__ counters, const float *__restrict__ projz, const u32 *__restrict__ zrange, int S)
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
			s_range[threadIdx.x] = ug_upper_bound(scan, 0, T + 1, x) - 1;
		}
		__syncthreads();
		const int tlo = s_range[0], thi = s_range[1];
		for (u32 p = blk + threadIdx.x; p < blk_end; p += EMIT_THREADS) {
			const int t = ug_upper_bound(scan, tlo, thi + 1, p) - 1;
			u32 key = ug_pair_key(rect[t], p - scan[t], gy_lo, gy_cnt);
			if (S > 1) key = key * (u32)S + ug_slab_of(projz[t], zrange, S); /* grid_kernel.cu:322 */
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
	const unsigned blocks = (unsigned)(want < (size_t)sms * 16 ? (want ? want : 1) : (size_t)sms * 16);
	k_emit_pairs<<<blocks, EMIT_THREADS, 0, s>>>(vc, g.scan, g.rect, g.T, (u32)g.cap, g.keys[0], g.vals[0], g.span, g.counters, g.projz, g.zrange, g.S);
}

void ugk_cell_offsets(cudaStream_t s, GridBuffers &g, int sms)
{
	ugk_exclusive_scan(s, g.span, g.offset, (size_t)g.C, g.sort.scan_state, g.sort.scan_tiles, sms);
}
