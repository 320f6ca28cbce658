/* k_trace.cu -- primary rays (a9-a12): rckernel_alpha, trace_kernel.cu:84-270, launched by
 * FrustumTracer::trace (frustum_tracer.h:35-58).
 *
 * One 8x8 pixel tile = one cell of the perspective grid = one ray packet of 64 threads.  The cell's
 * triangle list (a contiguous run of d_triangle_value_list) is walked in batches: each thread of the
 * packet fetches one 64-byte ray-independent triangle record (4 x LDG.128, written by k_setup_bin)
 * into shared memory, then every ray tests every staged triangle (broadcast LDS.128).
 * Nearest hit = smallest |t| with 0 < |t| < best, strict '<' so the FIRST triangle in list order wins
 * ties, and list order inside a cell is ascending triangle id (trace_kernel.cu:35-38).
 *
 * Compute bound (FP32 issue): ~40 FP32 instructions per ray-triangle test after hoisting the
 * ray-independent cross/dot products; see DESIGN.md for the roofline.
 */
#include "k_common.cuh"

#define TR_BATCH 64

template <bool FULL>
__global__ void __launch_bounds__(64) k_trace(const FrameConst *__restrict__ fc, const u32 *__restrict__ span, const u32 *__restrict__ offset,
                                              const u32 *__restrict__ value_list, const float4 *__restrict__ rec, float *__restrict__ out_t,
                                              int *__restrict__ out_id, int *__restrict__ out_shadowed, float *__restrict__ out_normal,
                                              float *__restrict__ out_dir)
{
	__shared__ float4 s_rec[TR_BATCH * 4];
	__shared__ u32 s_tri[TR_BATCH];
	__shared__ float s_lat[100];
	const ViewConst &vc = fc->cam;
	const int tid = threadIdx.x;
	for (int i = tid; i < 100; i += 64) s_lat[i] = vc.lattice[i];
	const int gy_cnt = vc.gy_cnt, gy_lo = vc.gy_lo, W = vc.W, H = vc.H;
	const float ex = vc.cc[0], ey = vc.cc[1], ez = vc.cc[2];
	const int cell = blockIdx.x;                 /* local cell id = gx*gy_cnt + (gy-gy_lo); trace_kernel.cu:138 */
	const int gx = cell / gy_cnt, gyl = cell - gx * gy_cnt;
	const int px = gx * UG_TILE + (tid & 7), pyl = gyl * UG_TILE + (tid >> 3); /* strip-local row */
	const int py = pyl + gy_lo * UG_TILE;
	const int pix = pyl * W + px;                /* trace_kernel.cu:91 */
	__syncthreads();
	float dx, dy, dz;
	ug_primary_dir(s_lat, ex, ey, ez, px, py, W, H, dx, dy, dz);

	const u32 n = span[cell], off = offset[cell];
	float best = 99999999.9f; /* trace_kernel.cu:126 */
	int hit = -1;             /* position in the batch-relative list: off + k */
	for (u32 base = 0; base < n; base += TR_BATCH) {
		const u32 cnt = min((u32)TR_BATCH, n - base);
		if ((u32)tid < cnt) {
			const u32 tri = value_list[off + base + tid];
			const float4 *src = rec + 4 * (size_t)tri;
			s_tri[tid] = tri;
			s_rec[4 * tid + 0] = src[0];
			s_rec[4 * tid + 1] = src[1];
			s_rec[4 * tid + 2] = src[2];
			s_rec[4 * tid + 3] = src[3];
		}
		__syncthreads();
		for (u32 k = 0; k < cnt; k++) {
			float t;
			if (ug_intersect(s_rec[4 * k], s_rec[4 * k + 1], s_rec[4 * k + 2], s_rec[4 * k + 3], dx, dy, dz, t)) {
				t = fabsf(t);                    /* trace_kernel.cu:35 */
				if (t < best && t > 0.0f) {      /* :38 */
					best = t;
					hit = (int)s_tri[k];
				}
			}
		}
		__syncthreads();
	}
	/* `hit` is remembered as the triangle id of the staged slot; since s_tri is overwritten per batch
	 * the id itself (not the slot) is kept. */
	bool keep = false;
	if (hit >= 0) { /* isWithin, trace_kernel.cu:56-82 with NUM_SLABS 1 */
		const float X = ex + best * dx, Y = ey + best * dy, Z = ez + best * dz;
		const float *M = vc.cc + 48;
		const float z = ug_row(M, 2, X, Y, Z, 1.0f) / ug_row(M, 3, X, Y, Z, 1.0f);
		keep = ug_f2i(floorf(z * 1.0f)) == 0;
	}
	if (keep) {
		out_t[pix] = best;
		out_id[pix] = hit;
	} else {
		out_t[pix] = -1.0f;
		out_id[pix] = -2;
	}
	out_shadowed[pix] = 0;
	if (FULL) { /* reference buffers, decision_data.h:66-70 */
		float nx = -1.0f, ny = -1.0f, nz = -1.0f;
		if (keep) {
			const float4 r1 = rec[4 * (size_t)hit + 1], r2 = rec[4 * (size_t)hit + 2];
			float ax = r1.x, ay = r1.y, az = r1.z, bx = r2.x, by = r2.y, bz = r2.z;
			ug_normalize(ax, ay, az);
			ug_normalize(bx, by, bz);
			ug_cross(nx, ny, nz, ax, ay, az, bx, by, bz);
			ug_normalize(nx, ny, nz);
			nx = (nx < 0) ? nx * -1.0f : nx; ny = (ny < 0) ? ny * -1.0f : ny; nz = (nz < 0) ? nz * -1.0f : nz;
		}
		out_normal[3 * (size_t)pix + 0] = nx; out_normal[3 * (size_t)pix + 1] = ny; out_normal[3 * (size_t)pix + 2] = nz;
		out_dir[3 * (size_t)pix + 0] = dx; out_dir[3 * (size_t)pix + 1] = dy; out_dir[3 * (size_t)pix + 2] = dz;
	}
}

void ugk_trace(cudaStream_t s, const FrameConst *fc, const GridBuffers &g, float *t, int *id, int *shadowed, float *normal, float *dir, int C, int)
{
	if (C <= 0) return;
	if (normal && dir)
		k_trace<true><<<C, 64, 0, s>>>(fc, g.span, g.offset, g.vals[g.final_buf], g.tri_rec, t, id, shadowed, normal, dir);
	else
		k_trace<false><<<C, 64, 0, s>>>(fc, g.span, g.offset, g.vals[g.final_buf], g.tri_rec, t, id, shadowed, nullptr, nullptr);
}
