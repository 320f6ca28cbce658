/* k_trace.cu -- primary rays (a9-a12): rckernel_alpha, trace_kernel.cu:84-270, launched by
 * FrustumTracer::trace (frustum_tracer.h:35-58).
 *
 * One 8x8 pixel tile = one cell of the perspective grid = one ray packet of 64 threads, as in the
 * reference.  What differs is the work decomposition: the reference gives every cell one CTA, so the
 * frame time is the time of the longest cell list (border cells collect every triangle that was
 * clamped into the frustum, grid_kernel.cu:217-229 - thousands of entries).  Here the unit of work is
 * a fixed-size RANGE of the sorted (cell,triangle) pair array:
 *
 *   item i  = pairs [i*S, (i+1)*S)                      (S = TR_ITEM, all items cost the same)
 *   a range covers a few whole cells and at most two partial ones; the cell of every pair is the
 *   sorted key, so runs are found from shared memory without touching span/offset.
 *   whole cell   -> nearest hit is final: isWithin + write-out straight from registers
 *   partial cell -> per-pixel 64-bit atomicMin of (t bits << 32 | pair position); the last of the
 *                   cell's items to arrive (atomic counter) finalises the 64 pixels.
 *
 * Nearest hit = smallest |t| with 0 < |t| < best, strict '<' so the FIRST triangle in list order wins
 * ties (trace_kernel.cu:35-38); positive floats order like their bit patterns and pair positions
 * ascend in list order, so min over (t bits, position) is exactly that rule.
 *
 * Triangle data: each thread of the packet fetches one 64-byte ray-independent record (4 x LDG.128,
 * written by k_setup_bin) into shared memory, every ray then tests every staged triangle with
 * broadcast LDS.128.  Compute bound (FP32 issue); see DESIGN.md for the roofline.
 */
#include "k_common.cuh"

/* ---- TMA bulk copy + mbarrier (cp.async.bulk / mbarrier PTX; SASS: UBLKCP, SYNCS) */
static __device__ __forceinline__ void tr_mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
static __device__ __forceinline__ void tr_mbar_expect(u32 bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
static __device__ __forceinline__ void tr_mbar_wait(u32 bar, u32 parity)
{
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "TR_WAIT:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra TR_DONE;\n"
	    "bra TR_WAIT;\n"
	    "TR_DONE:\n"
	    "}\n" ::"r"(bar),
	    "r"(parity)
	    : "memory");
}
static __device__ __forceinline__ void tr_bulk_g2s(u32 dst_smem, const void *src, u32 bytes, u32 bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
	             : "memory");
}

#ifndef TR_ITEM
#define TR_ITEM 128   /* pairs per work item */
#endif
#define TR_BATCH 64   /* triangles staged per round = threads per packet */

__global__ void __launch_bounds__(256) k_init_pixels(const FrameConst *__restrict__ fc, float *__restrict__ t, int *__restrict__ id,
                                                    int *__restrict__ shadowed, u64 *__restrict__ best, float *__restrict__ normal,
                                                    float *__restrict__ dir, int N)
{
	__shared__ float s_lat[100];
	const ViewConst &vc = fc->cam;
	for (int i = threadIdx.x; i < 100; i += blockDim.x) s_lat[i] = vc.lattice[i];
	__syncthreads();
	const int W = vc.W, H = vc.H, row0 = vc.gy_lo * UG_TILE;
	const float ex = vc.cc[0], ey = vc.cc[1], ez = vc.cc[2];
	for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < N; pix += gridDim.x * blockDim.x) {
		t[pix] = -1.0f;      /* trace_kernel.cu:257-263: the miss encoding; hits overwrite it */
		id[pix] = -2;
		shadowed[pix] = 0;
		best[pix] = ~0ull;
		/* d_primary_ray_direction (decision_data.h:14) is written for every pixel, hit or miss (trace_kernel.cu:264-266);
		 * the primary, shadow-map, shadow and shade kernels all read it instead of recomputing it */
		const int pyl = pix / W, px = pix - pyl * W;
		float dx, dy, dz;
		ug_primary_dir(s_lat, ex, ey, ez, px, pyl + row0, W, H, dx, dy, dz);
		dir[3 * (size_t)pix + 0] = dx; dir[3 * (size_t)pix + 1] = dy; dir[3 * (size_t)pix + 2] = dz;
		if (normal) { normal[3 * (size_t)pix + 0] = -1.0f; normal[3 * (size_t)pix + 1] = -1.0f; normal[3 * (size_t)pix + 2] = -1.0f; }
	}
}

/* the depth slab isWithin (trace_kernel.cu:56-82) assigns to the point at distance `best` along the ray: floor(ndc_z * NUM_SLABS) */
static __device__ __forceinline__ int hit_slab(const float *cc, float dx, float dy, float dz, float best, int S)
{
	const float X = cc[0] + best * dx, Y = cc[1] + best * dy, Z = cc[2] + best * dz;
	const float *M = cc + 48;
	const float z = ug_row(M, 2, X, Y, Z, 1.0f) / ug_row(M, 3, X, Y, Z, 1.0f);
	return ug_f2i(floorf(z * (float)S));
}

/* write-out (trace_kernel.cu:232-253) of one pixel's final hit */
static __device__ __forceinline__ void write_hit(float best, u32 tri, int pix, const float4 *__restrict__ rec, float *__restrict__ out_t, int *__restrict__ out_id,
                                                 float *__restrict__ out_normal);

/* isWithin (NUM_SLABS 1) + write-out of one pixel's final hit */
static __device__ __forceinline__ void finalize_pixel(const float *cc, float dx, float dy, float dz, float best, u32 tri, int pix,
                                                      const float4 *__restrict__ rec, float *__restrict__ out_t, int *__restrict__ out_id,
                                                      float *__restrict__ out_normal)
{
	if (hit_slab(cc, dx, dy, dz, best, 1) != 0) return; /* outside NDC z [0,1): stays a miss, even if a farther hit existed */
	write_hit(best, tri, pix, rec, out_t, out_id, out_normal);
}

static __device__ __forceinline__ void write_hit(float best, u32 tri, int pix, const float4 *__restrict__ rec, float *__restrict__ out_t, int *__restrict__ out_id,
                                                 float *__restrict__ out_normal)
{
	out_t[pix] = best;
	out_id[pix] = (int)tri;
	if (out_normal) {
		const float4 r1 = rec[4 * (size_t)tri + 1], r2 = rec[4 * (size_t)tri + 2];
		float ax = r1.x, ay = r1.y, az = r1.z, bx = r2.x, by = r2.y, bz = r2.z, nx, ny, nz;
		ug_normalize(ax, ay, az);
		ug_normalize(bx, by, bz);
		ug_cross(nx, ny, nz, ax, ay, az, bx, by, bz);
		ug_normalize(nx, ny, nz);
		nx = (nx < 0) ? nx * -1.0f : nx; ny = (ny < 0) ? ny * -1.0f : ny; nz = (nz < 0) ? nz * -1.0f : nz;
		out_normal[3 * (size_t)pix + 0] = nx; out_normal[3 * (size_t)pix + 1] = ny; out_normal[3 * (size_t)pix + 2] = nz;
	}
}

#ifndef TR_MINB
#define TR_MINB 16 /* 64 registers: no spills (the compiler's own choice was 48 with spills); measured -5 us on config B */
#endif
template <int ITEM>
__global__ void __launch_bounds__(64, TR_MINB) k_trace(const FrameConst *__restrict__ fc, const u32 *__restrict__ counters, const u32 *__restrict__ keys,
                                              const u32 *__restrict__ vals, const u32 *__restrict__ span, const u32 *__restrict__ offset,
                                              const float4 *__restrict__ rec, u64 *__restrict__ best_buf, u32 *__restrict__ cell_done,
                                              const float *__restrict__ dir, float *__restrict__ out_t, int *__restrict__ out_id,
                                              float *__restrict__ out_normal)
{
#ifndef TR_NO_TMA
	__shared__ __align__(128) float4 s_rec[2 * TR_BATCH * 4]; /* two staging windows, filled by TMA */
	__shared__ __align__(8) unsigned long long s_bar[2];     /* one mbarrier per window: completes when its records have landed */
#else
	__shared__ float4 s_rec[TR_BATCH * 4];
#endif
	__shared__ u32 s_key[ITEM + 2]; /* keys[lo-1 .. hi], so both ends can tell whole cells from partial ones */
	__shared__ u32 s_val[ITEM];
	__shared__ u32 s_head[ITEM / 32]; /* bit p = pair lo+p starts a run (its key differs from the previous pair's) */
	__shared__ float s_cc[64];
	__shared__ u32 s_ticket;
	ug_pdl_enter<UG_PDL_TRACE>();
	const ViewConst &vc = fc->cam;
	const int tid = threadIdx.x;
	s_cc[tid] = vc.cc[tid];
	const int gy_cnt = vc.gy_cnt, W = vc.W;
	const u32 rec_base0 = (u32)__cvta_generic_to_shared(s_rec);
	const u32 P = counters[0];
	const u32 nitems = (P + ITEM - 1) / ITEM;
#ifndef TR_NO_TMA
	const u32 bar0 = (u32)__cvta_generic_to_shared(&s_bar[0]);
	if (tid == 0) { tr_mbar_init(bar0, 1); tr_mbar_init(bar0 + 8u, 1); }
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	u32 wc = 0, phase0 = 0, phase1 = 0; /* windows staged so far (window w uses buffer w & 1), parity of each barrier's next completion */
#endif
	__syncthreads();
	for (u32 item = blockIdx.x; item < nitems; item += gridDim.x) {
		const u32 lo = item * ITEM, hi = min(lo + ITEM, P);
		__syncthreads(); /* previous item's shared data is dead */
		for (u32 i = tid; i < hi - lo + 2; i += 64) {
			const u32 g = lo + i; /* s_key[i] = keys[lo - 1 + i] */
			s_key[i] = (g == 0 || g > P) ? 0xffffffffu : keys[g - 1];
		}
		for (u32 i = tid; i < hi - lo; i += 64) s_val[i] = vals[lo + i];
		__syncthreads();
#pragma unroll
		for (u32 w = 0; w < ITEM / 64; w++) { /* run heads, 32 pairs per ballot */
			const u32 p = w * 64 + tid;
			const bool head = (p < hi - lo) && (p == 0 || s_key[p + 1] != s_key[p]);
			const u32 bits = __ballot_sync(UG_FULL_MASK, head);
			if ((tid & 31) == 0) s_head[p >> 5] = bits;
		}
		/* Staging is decoupled from the runs: windows of 64 consecutive pairs.  Every thread issues ONE 64-byte TMA bulk copy
		 * (cp.async.bulk global -> shared) for the record its pair names; an mbarrier counts the bytes.  The next window is
		 * issued into the other buffer before the current one is tested, so its latency hides behind 64 x 64 tests. */
#ifndef TR_NO_TMA
		{
			const u32 cnt = min((u32)TR_BATCH, hi - lo), b = wc & 1u;
			if (tid == 0) tr_mbar_expect(bar0 + 8u * b, cnt * 64u);
			if ((u32)tid < cnt) tr_bulk_g2s(rec_base0 + b * (TR_BATCH * 64u) + tid * 64u, rec + 4 * (size_t)s_val[tid], 64u, bar0 + 8u * b);
		}
#else
		float4 p0 = make_float4(0, 0, 0, 0), p1 = p0, p2 = p0, p3 = p0;
		if (lo + tid < hi) {
			const float4 *src = rec + 4 * (size_t)s_val[tid];
			p0 = src[0]; p1 = src[1]; p2 = src[2]; p3 = src[3];
		}
#endif

		/* state of the run being walked (a run = the part of one cell's list inside this item) */
		bool have = false;
		u32 cell = 0, run_start = lo, hit = 0xffffffffu;
		int pix = 0;
		float dx = 0.0f, dy = 0.0f, dz = 1.0f, best = 99999999.9f;

		for (u32 wlo = lo; wlo < hi; wlo += TR_BATCH) { /* the last window (whi == hi) also closes the last run */
			const u32 whi = min(wlo + TR_BATCH, hi);
#ifndef TR_NO_TMA
			const u32 b = wc & 1u;
			const u32 rec_base = rec_base0 + b * (TR_BATCH * 64u);
			__syncthreads(); /* everybody is done with the other buffer (the window before this one) */
			if (whi < hi) {
				const u32 cnt2 = min((u32)TR_BATCH, hi - whi), bar2 = bar0 + 8u * (b ^ 1u);
				if (tid == 0) tr_mbar_expect(bar2, cnt2 * 64u);
				if ((u32)tid < cnt2) tr_bulk_g2s(rec_base0 + (b ^ 1u) * (TR_BATCH * 64u) + tid * 64u, rec + 4 * (size_t)s_val[whi - lo + tid], 64u, bar2);
			}
			tr_mbar_wait(bar0 + 8u * b, b ? phase1 : phase0);
			if (b) phase1 ^= 1u; else phase0 ^= 1u;
			wc++;
#else
			const u32 rec_base = rec_base0;
			__syncthreads();
			if (wlo + tid < whi) { s_rec[4 * tid + 0] = p0; s_rec[4 * tid + 1] = p1; s_rec[4 * tid + 2] = p2; s_rec[4 * tid + 3] = p3; }
			__syncthreads();
			if (whi + tid < hi) {
				const float4 *src = rec + 4 * (size_t)s_val[whi - lo + tid];
				p0 = src[0]; p1 = src[1]; p2 = src[2]; p3 = src[3];
			}
#endif
			u32 p = wlo;
			while (true) {
				const bool at_end = p >= hi;
				const bool head = at_end || ((s_head[(p - lo) >> 5] >> ((p - lo) & 31)) & 1u);
				if (head && have) { /* the run [run_start, p) is complete */
					const bool whole = (s_key[run_start - lo] != cell) && (s_key[p - lo + 1] != cell);
					if (whole) {
						if (hit != 0xffffffffu) finalize_pixel(s_cc, dx, dy, dz, best, s_val[hit - lo], pix, rec, out_t, out_id, out_normal);
					} else {
						if (hit != 0xffffffffu) atomicMin((unsigned long long *)&best_buf[pix], ((u64)__float_as_uint(best) << 32) | (u64)hit);
						const u32 c_lo = offset[cell], c_hi = c_lo + span[cell];
						const u32 parts = (c_hi - 1) / ITEM - c_lo / ITEM + 1;
						__threadfence();
						__syncthreads();
						if (tid == 0) s_ticket = atomicAdd(&cell_done[cell], 1u);
						__syncthreads();
						if (s_ticket == parts - 1) { /* last item of this cell: every partial result is in best_buf */
							__threadfence();
							const u64 m = __ldcg((const unsigned long long *)&best_buf[pix]);
							if (m != ~0ull)
								finalize_pixel(s_cc, dx, dy, dz, __uint_as_float((u32)(m >> 32)), __ldg(&vals[(u32)m]), pix, rec, out_t, out_id, out_normal);
							if (tid == 0) cell_done[cell] = 0; /* ready for the next frame */
						}
					}
					have = false;
				}
				if (at_end || p >= whi) break;
				if (head) { /* a new run starts at p */
					cell = s_key[p - lo + 1];
					const int gx = (int)cell / gy_cnt, gyl = (int)cell - gx * gy_cnt;   /* trace_kernel.cu:138 (strip-local) */
					const int px = gx * UG_TILE + (tid & 7), pyl = gyl * UG_TILE + (tid >> 3);
					pix = pyl * W + px;                                                  /* trace_kernel.cu:91 */
					dx = dir[3 * (size_t)pix]; dy = dir[3 * (size_t)pix + 1]; dz = dir[3 * (size_t)pix + 2];
					best = 99999999.9f; /* trace_kernel.cu:126 */
					hit = 0xffffffffu;  /* absolute pair position of the best hit */
					run_start = p;
					have = true;
				}
				u32 e = whi; /* next run head after p, inside this window */
				{
					const u32 q = p - lo + 1;
					u32 w = q >> 5;
					if (w < ITEM / 32) {
						u32 bits = s_head[w] & (0xffffffffu << (q & 31));
						while (true) {
							if (bits) { e = min(whi, lo + (w << 5) + (u32)__ffs(bits) - 1); break; }
							if (++w >= ITEM / 32) break;
							bits = s_head[w];
						}
					}
				}
				u32 k = p;
#ifndef TR_NO_ILP2
				for (; k + 1 < e; k += 2) { /* two triangles in flight per ray (ug_isect_u); hits are taken in list order */
					const u32 ad = rec_base + (k - wlo) * 64u;
					const float4 a2 = ug_lds128(ad + 32), a1 = ug_lds128(ad + 16), b2 = ug_lds128(ad + 96), b1 = ug_lds128(ad + 80);
					const float4 a0 = ug_lds128(ad), b0 = ug_lds128(ad + 64);
					TriU A, B;
					const bool ha = ug_isect_u(a0, a1, a2, dx, dy, dz, A), hb = ug_isect_u(b0, b1, b2, dx, dy, dz, B);
					float t;
					if (ha && ug_isect_vt(a0, ug_lds128(ad + 48), dx, dy, dz, A, t)) {
						t = fabsf(t);                    /* trace_kernel.cu:35 */
						if (t < best && t > 0.0f) { best = t; hit = k; } /* :38 */
					}
					if (hb && ug_isect_vt(b0, ug_lds128(ad + 112), dx, dy, dz, B, t)) {
						t = fabsf(t);
						if (t < best && t > 0.0f) { best = t; hit = k + 1; }
					}
				}
#endif
				for (; k < e; k++) {
					float t;
					if (ug_intersect_lds(rec_base + (k - wlo) * 64u, dx, dy, dz, t)) {
						t = fabsf(t);                    /* trace_kernel.cu:35 */
						if (t < best && t > 0.0f) {      /* :38 */
							best = t;
							hit = k;
						}
					}
				}
				p = e;
			}
		}
	}
}

/* NUM_SLABS > 1: rckernel_alpha's slab loop, trace_kernel.cu:132-229, as the reference runs it - one CTA per
 * 8x8 tile walks the tile's cell slab by slab.  A ray is 0 (nothing yet), 1 (has a hit whose depth slab is not the one
 * just walked) or 2 (done); isWithin turns a done ray back into 0 in the next slab the tile walks, and the tile stops as
 * soon as all 64 rays are done after the same slab (metadata[3]).  That coupling is why this cannot use k_trace's
 * range decomposition; it is a fidelity mode, not a fast path. */
__global__ void __launch_bounds__(64) k_trace_slabs(const FrameConst *__restrict__ fc, const u32 *__restrict__ vals, const u32 *__restrict__ span,
                                                    const u32 *__restrict__ offset, const float4 *__restrict__ rec, const float *__restrict__ dir,
                                                    float *__restrict__ out_t, int *__restrict__ out_id, float *__restrict__ out_normal, int cells, int S)
{
	__shared__ float4 s_rec[TR_BATCH * 4];
	__shared__ float s_cc[64];
	const ViewConst &vc = fc->cam;
	const int tid = threadIdx.x;
	s_cc[tid] = vc.cc[tid];
	const int gy_cnt = vc.gy_cnt, W = vc.W;
	const u32 rec_base = (u32)__cvta_generic_to_shared(s_rec);
	__syncthreads();
	for (int cell = blockIdx.x; cell < cells; cell += gridDim.x) {
		const int gx = cell / gy_cnt, gyl = cell - gx * gy_cnt;
		const int pix = (gyl * UG_TILE + (tid >> 3)) * W + gx * UG_TILE + (tid & 7);
		const float dx = dir[3 * (size_t)pix], dy = dir[3 * (size_t)pix + 1], dz = dir[3 * (size_t)pix + 2];
		int rayDone = 0;
		float best = 99999999.9f;
		u32 hit = 0xffffffffu;
		for (int slab = 0; slab < S; slab++) {
			const u32 n = span[(size_t)cell * S + slab], off = offset[(size_t)cell * S + slab];
			for (u32 base = 0; base < n; base += TR_BATCH) {
				const u32 cnt = min((u32)TR_BATCH, n - base);
				__syncthreads();
				if ((u32)tid < cnt) {
					const float4 *src = rec + 4 * (size_t)vals[off + base + tid];
					s_rec[4 * tid + 0] = src[0]; s_rec[4 * tid + 1] = src[1]; s_rec[4 * tid + 2] = src[2]; s_rec[4 * tid + 3] = src[3];
				}
				__syncthreads();
				if (rayDone != 2)
					for (u32 k = 0; k < cnt; k++) {
						float t;
						if (ug_intersect_lds(rec_base + k * 64u, dx, dy, dz, t)) {
							t = fabsf(t);
							if (t < best && t > 0.0f) { best = t; hit = off + base + k; rayDone = 1; }
						}
					}
			}
			if (rayDone == 1) rayDone = (hit_slab(s_cc, dx, dy, dz, best, S) == slab) ? 2 : 1;
			else rayDone = 0;
			if (__syncthreads_and(rayDone == 2)) break;
		}
		if (rayDone == 2) write_hit(best, vals[hit], pix, rec, out_t, out_id, out_normal);
	}
}

void ugk_trace_slabs(cudaStream_t s, const FrameConst *fc, const GridBuffers &g, const float *dir, float *t, int *id, float *normal, int cells, int sms)
{
	if (cells <= 0) return;
	k_trace_slabs<<<min(cells, sms * 32), 64, 0, s>>>(fc, g.vals[g.final_buf], g.span, g.offset, g.tri_rec, dir, t, id, normal, cells, g.S);
}

void ugk_init_pixels(cudaStream_t s, const FrameConst *fc, float *t, int *id, int *shadowed, u64 *best, float *normal, float *dir, int N, int sms)
{
	if (N <= 0) return;
	k_init_pixels<<<min((N + 255) / 256, sms * 8), 256, 0, s>>>(fc, t, id, shadowed, best, normal, dir, N);
}

void ugk_trace(cudaStream_t s, const FrameConst *fc, const GridBuffers &g, u64 *best, u32 *cell_done, const float *dir, float *t, int *id, float *normal,
               int sms, bool small_items)
{
	/* Items of 128 pairs; 64 when the grid holds so few pairs that the items would not fill the GPU twice (16 blocks are
	 * resident per SM): the idle SMs of the last wave cost more than the per-item overhead (config B, 352 K pairs: 74.6 ->
	 * 70.3 us; config D, 3.4 M pairs: 547 -> 574 us, so large grids keep 128).  The caller decides from the pair count of
	 * the last finished frame. */
	const size_t item = small_items ? TR_ITEM / 2 : TR_ITEM;
	const size_t want = (g.cap + item - 1) / item;
	const unsigned blocks = (unsigned)(want < (size_t)sms * 32 ? (want ? want : 1) : (size_t)sms * 32);
	if (small_items)
		ug_launch(UG_PDL_TRACE, k_trace<TR_ITEM / 2>, dim3(blocks), dim3(64), s, fc, g.counters, g.keys[g.final_buf], g.vals[g.final_buf], g.span, g.offset, g.tri_rec, best, cell_done, dir, t, id, normal);
	else
		ug_launch(UG_PDL_TRACE, k_trace<TR_ITEM>, dim3(blocks), dim3(64), s, fc, g.counters, g.keys[g.final_buf], g.vals[g.final_buf], g.span, g.offset, g.tri_rec, best, cell_done, dir, t, id, normal);
}
