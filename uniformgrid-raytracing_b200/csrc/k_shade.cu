/* k_shade.cu -- K1 vertex transform (a1) and K10 shading (a19, a20).
 *
 *   k_transform   transformation_kernel.cu:4-18 copy_data_transform (Model::rotate_bunny, scene.h:122-139):
 *                 pure HBM streaming, 12 B in + 12 B out per vertex; threads move float4s over the
 *                 packed float3 array (4 vertices = 3 float4) when the sub-range is 16-byte aligned.
 *   k_shade       shader_kernel.cu:165-221 lambertian_shade + :46-86 lambert_color_pixel, with
 *                 shader_kernel.cu:347-359 shadow_kernel fused in (apply_shadow) so the image is
 *                 written once.  The normal is recomputed from the triangle id with the reference's
 *                 arithmetic instead of being stored per pixel; the direction comes from d_primary_ray_direction.
 */
#include "k_shade_dev.cuh"

/* one vertex, transformation_kernel.cu:10-16; cos/sin come from the host (same glibc values the
 * host-compiled reference gets; CUDA's cosf/sinf differ in the last ulp) */
static __device__ __forceinline__ void xform(float ix, float iy, float iz, float c, float s, float &ox, float &oy, float &oz)
{
	const float x = (ix - 12.0f) / 12.0f;
	const float y = (iy - 11.0f) / 12.0f;
	const float z = (iz - 4.5f) / 12.0f;
	ox = (x * c - y * s) * 9.0f + 14.5f;
	oy = (x * s + y * c) * 9.0f + 13.0f;
	oz = z * 9.0f + 4.0f;
}

__global__ void __launch_bounds__(256) k_transform_vec4(float4 *__restrict__ dst, const float4 *__restrict__ src, int groups, const FrameConst *__restrict__ fc)
{
	const float c = fc->rot_cos, s = fc->rot_sin;
	for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
		const float4 a = src[3 * (size_t)g], b = src[3 * (size_t)g + 1], d = src[3 * (size_t)g + 2];
		float4 oa, ob, od;
		xform(a.x, a.y, a.z, c, s, oa.x, oa.y, oa.z);
		xform(a.w, b.x, b.y, c, s, oa.w, ob.x, ob.y);
		xform(b.z, b.w, d.x, c, s, ob.z, ob.w, od.x);
		xform(d.y, d.z, d.w, c, s, od.y, od.z, od.w);
		dst[3 * (size_t)g] = oa; dst[3 * (size_t)g + 1] = ob; dst[3 * (size_t)g + 2] = od;
	}
}
__global__ void __launch_bounds__(256) k_transform_scalar(float *__restrict__ dst, const float *__restrict__ src, int first, int count, const FrameConst *__restrict__ fc)
{
	const float c = fc->rot_cos, s = fc->rot_sin;
	for (int v = first + blockIdx.x * blockDim.x + threadIdx.x; v < first + count; v += gridDim.x * blockDim.x)
		xform(src[3 * (size_t)v], src[3 * (size_t)v + 1], src[3 * (size_t)v + 2], c, s, dst[3 * (size_t)v], dst[3 * (size_t)v + 1], dst[3 * (size_t)v + 2]);
}

void ugk_transform(cudaStream_t s, float *vertexlist, const float *orig, int size, int offset, const FrameConst *fc, int sms)
{
	if (size <= 0) return;
	float *dst = vertexlist + 3 * (size_t)offset;
	const bool aligned = (((uintptr_t)dst | (uintptr_t)orig) & 15u) == 0;
	const int groups = aligned ? size / 4 : 0;
	if (groups) k_transform_vec4<<<min((groups + 255) / 256, sms * 8), 256, 0, s>>>((float4 *)dst, (const float4 *)orig, groups, fc);
	const int rest = size - 4 * groups;
	if (rest) k_transform_scalar<<<min((rest + 255) / 256, sms * 8), 256, 0, s>>>(dst, orig, 4 * groups, rest, fc);
}

/* 4 consecutive pixels per thread: 16-byte loads of t/id/shadowed, 3 x 16-byte loads of dir, the 12 RGB bytes leave
 * as three 32-bit stores (N is a multiple of 64, every buffer is 16-byte aligned).
 * RECOMP: the ray direction is evaluated again from the 5x5 lattice (ug_primary_dir, the expression k_init_pixels stored
 * d_primary_ray_direction with, so the bits are the same) instead of being read: 12 of the kernel's 24 input bytes per
 * pixel go away, and a pixel without a hit never needs its direction.  The four pixels share a screen row (W is a
 * multiple of the tile), so the row part of the lattice weights is common. */
template <bool SPOT, bool RECOMP, bool HOIST>
__global__ void __launch_bounds__(256) k_shade(const FrameConst *__restrict__ fc, const float *__restrict__ t, const float *__restrict__ dir,
                                              const int *__restrict__ id, const int *__restrict__ shadowed, const float4 *__restrict__ rec,
                                              const int *__restrict__ mat_idx, const float *__restrict__ mat_list, int mat_count,
                                              unsigned char *__restrict__ rgb, int *__restrict__ mat_out, int N, int apply_shadow,
                                              unsigned char *__restrict__ rgb_remote, const float4 *__restrict__ nv)
{
	ug_pdl_enter<UG_PDL_SHADE>();
	const ViewConst &cam = fc->cam;
	const float *R = fc->shade_mv, *leye = fc->shade_eye;
	const float ex = cam.cc[0], ey = cam.cc[1], ez = cam.cc[2];
	/* shader_kernel.cu:54-57: light 0 through the rotation part of the resident modelview */
	const float l0 = fc->light_pos[0], l1 = fc->light_pos[1], l2 = fc->light_pos[2];
	const float Lvx = R[0] * l0 + R[4] * l1 + R[8] * l2;
	const float Lvy = R[1] * l0 + R[5] * l1 + R[9] * l2;
	const float Lvz = R[2] * l0 + R[6] * l1 + R[10] * l2;
	const int groups = N / 4;
	__shared__ float s_lat[100];
	const int W = cam.W, H = cam.H, row0 = cam.gy_lo * UG_TILE;
	if (RECOMP) {
		for (int i = threadIdx.x; i < 100; i += blockDim.x) s_lat[i] = cam.lattice[i];
		__syncthreads();
	}
	for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
		const float4 tv = ((const float4 *)t)[g];
		const int4 tri = ((const int4 *)id)[g];
		const int4 sh = ((const int4 *)shadowed)[g];
		float4 d0, d1, d2;
		if (RECOMP) {
			const int pyl = (4 * g) / W, px = 4 * g - pyl * W;
			d0 = d1 = d2 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			if (tri.x >= 0 && tv.x > 0.0f) ug_primary_dir(s_lat, ex, ey, ez, px, pyl + row0, W, H, d0.x, d0.y, d0.z);
			if (tri.y >= 0 && tv.y > 0.0f) ug_primary_dir(s_lat, ex, ey, ez, px + 1, pyl + row0, W, H, d0.w, d1.x, d1.y);
			if (tri.z >= 0 && tv.z > 0.0f) ug_primary_dir(s_lat, ex, ey, ez, px + 2, pyl + row0, W, H, d1.z, d1.w, d2.x);
			if (tri.w >= 0 && tv.w > 0.0f) ug_primary_dir(s_lat, ex, ey, ez, px + 3, pyl + row0, W, H, d2.y, d2.z, d2.w);
		} else {
			d0 = ((const float4 *)dir)[3 * (size_t)g]; d1 = ((const float4 *)dir)[3 * (size_t)g + 1]; d2 = ((const float4 *)dir)[3 * (size_t)g + 2];
		}
		int m0, m1, m2, m3;
		const u32 c0 = shade_pixel<SPOT, HOIST>(R, leye, ex, ey, ez, Lvx, Lvy, Lvz, tv.x, d0.x, d0.y, d0.z, tri.x, sh.x, rec, mat_idx, mat_list, mat_count, apply_shadow, m0, nv);
		const u32 c1 = shade_pixel<SPOT, HOIST>(R, leye, ex, ey, ez, Lvx, Lvy, Lvz, tv.y, d0.w, d1.x, d1.y, tri.y, sh.y, rec, mat_idx, mat_list, mat_count, apply_shadow, m1, nv);
		const u32 c2 = shade_pixel<SPOT, HOIST>(R, leye, ex, ey, ez, Lvx, Lvy, Lvz, tv.z, d1.z, d1.w, d2.x, tri.z, sh.z, rec, mat_idx, mat_list, mat_count, apply_shadow, m2, nv);
		const u32 c3 = shade_pixel<SPOT, HOIST>(R, leye, ex, ey, ez, Lvx, Lvy, Lvz, tv.w, d2.y, d2.z, d2.w, tri.w, sh.w, rec, mat_idx, mat_list, mat_count, apply_shadow, m3, nv);
		if (mat_out) ((int4 *)mat_out)[g] = make_int4(m0, m1, m2, m3); /* shader_kernel.cu:176 */
		u32 *out = (u32 *)rgb + 3 * (size_t)g;
		const u32 w0 = c0 | (c1 << 24), w1 = (c1 >> 8) | (c2 << 16), w2 = (c2 >> 16) | (c3 << 8);
		out[0] = w0; out[1] = w1; out[2] = w2;
		if (rgb_remote) { /* multi-GPU gather fused into the shader: the same words go straight into rank 0's frame over NVLink (P2P stores) */
			u32 *far = (u32 *)rgb_remote + 3 * (size_t)g;
			far[0] = w0; far[1] = w1; far[2] = w2;
		}
	}
}

__global__ void __launch_bounds__(256) k_add_shadows(const int *__restrict__ shadowed, unsigned char *__restrict__ rgb, int N)
{
	for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < N; pix += gridDim.x * blockDim.x)
		if (shadowed[pix] == 1) {
			rgb[3 * (size_t)pix + 0] /= 3; rgb[3 * (size_t)pix + 1] /= 3; rgb[3 * (size_t)pix + 2] /= 3;
		}
}

void ugk_shade(cudaStream_t s, const FrameConst *fc, const float *t, const float *dir, const int *id, const int *shadowed, const float4 *rec, const int *mat_idx,
               const float *mat_list, int mat_count, unsigned char *rgb, int *mat_out, int N, int apply_shadow, int sms, bool spot, unsigned char *rgb_remote,
               bool recompute_dir, const float4 *nv)
{
	if (N <= 0) return;
	const int blocks = min((N / 4 + 255) / 256, sms * 8);
#define UG_SHADE_LAUNCH(SPOT, RECOMP, HOIST) \
	ug_launch(UG_PDL_SHADE, k_shade<SPOT, RECOMP, HOIST>, dim3(blocks), dim3(256), s, fc, t, dir, id, shadowed, rec, mat_idx, mat_list, mat_count, rgb, mat_out, N, apply_shadow, rgb_remote, nv)
#define UG_SHADE_PICK(SPOT, RECOMP) do { if (nv) UG_SHADE_LAUNCH(SPOT, RECOMP, true); else UG_SHADE_LAUNCH(SPOT, RECOMP, false); } while (0)
	if (spot) { if (recompute_dir) UG_SHADE_PICK(true, true); else UG_SHADE_PICK(true, false); }
	else { if (recompute_dir) UG_SHADE_PICK(false, true); else UG_SHADE_PICK(false, false); }
#undef UG_SHADE_PICK
#undef UG_SHADE_LAUNCH
}
/* Per triangle and frame: the view-space shading normal (ug_shade_normal) from the edges e1 = v1 - v0, e2 = v2 - v0 - the
 * subtractions ug_tri_setup stores in the records k_shade would otherwise read - and the modelview that is resident when
 * the frame shades.  Four of the five normalisations of lambert_color_pixel (shader_kernel.cu:46-86) depend on nothing
 * else, and with IEEE division and square root each of them is some forty instructions: at 4096x4096 k_shade was bound by
 * instruction issue, not by HBM.  Runs beside the grid builds (its own branch of the frame graph). */
__global__ void __launch_bounds__(256) k_shade_normals(const FrameConst *__restrict__ fc, const int *__restrict__ faces, const float *__restrict__ verts, int T,
                                                      float4 *__restrict__ nv)
{
	const float *R = fc->shade_mv;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
		const int i0 = faces[3 * t], i1 = faces[3 * t + 1], i2 = faces[3 * t + 2];
		const float *p0 = verts + 3 * (size_t)i0, *p1 = verts + 3 * (size_t)i1, *p2 = verts + 3 * (size_t)i2;
		const float ax = p0[0], ay = p0[1], az = p0[2];
		const float4 r1 = make_float4(p1[0] - ax, p1[1] - ay, p1[2] - az, 0.0f), r2 = make_float4(p2[0] - ax, p2[1] - ay, p2[2] - az, 0.0f);
		float x, y, z;
		ug_shade_normal(R, r1, r2, x, y, z);
		nv[t] = make_float4(x, y, z, 0.0f);
	}
}
void ugk_shade_normals(cudaStream_t s, const FrameConst *fc, const int *faces, const float *verts, int T, float4 *nv, int sms)
{
	if (T <= 0) return;
	k_shade_normals<<<min((T + 255) / 256, sms * 8), 256, 0, s>>>(fc, faces, verts, T, nv);
}

void ugk_add_shadows(cudaStream_t s, const int *shadowed, unsigned char *rgb, int N, int sms)
{
	if (N <= 0) return;
	k_add_shadows<<<min((N + 255) / 256, sms * 8), 256, 0, s>>>(shadowed, rgb, N);
}

/* development aid (env UG_TIMELINE=1): one-thread marker kernels between the stages of every branch of the frame graph,
 * each writing %globaltimer - the only way to see how the branches really interleave without a tracing tool */
__global__ void k_stamp(unsigned long long *slot)
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	*slot = t;
}
void ugk_stamp(cudaStream_t s, unsigned long long *slots, int slot) { if (slots) k_stamp<<<1, 1, 0, s>>>(slots + slot); }

/* The pair counts / overflow flags of every grid and the reflection ray counts ride back with the frame: one tiny kernel
 * writes them straight into the context's pinned (device-visible) block, instead of one copy-engine node per grid at the
 * tail of the frame graph (~5 us each in series). */
struct PublishArgs {
	const u32 *src[UG_MAX_LIGHTS + 5];
	int words[UG_MAX_LIGHTS + 5], dst_off[UG_MAX_LIGHTS + 5];
	int n;
};
__global__ void k_publish(const PublishArgs a, u32 *__restrict__ host_block)
{
	ug_pdl_enter<UG_PDL_PUBLISH>();
	for (int i = 0; i < a.n; i++)
		for (int w = threadIdx.x; w < a.words[i]; w += blockDim.x) host_block[a.dst_off[i] + w] = a.src[i][w];
}
/* word [1] of each listed counter block is that grid's overflow flag: their OR becomes this rank's contribution to the
 * frame's "done" handshake, so that all ranks learn in the same collective whether anyone has to run the frame again */
__global__ void k_overflow_any(const PublishArgs a, u32 *__restrict__ out)
{
	u32 any = 0;
	for (int i = 0; i < a.n; i++) any |= a.src[i][1];
	*out = any ? 1u : 0u;
}
void ugk_overflow_any(cudaStream_t s, const u32 *const *src, int n, u32 *out)
{
	PublishArgs a;
	a.n = n;
	for (int i = 0; i < n; i++) { a.src[i] = src[i]; a.words[i] = 0; a.dst_off[i] = 0; }
	k_overflow_any<<<1, 1, 0, s>>>(a, out);
}
void ugk_publish(cudaStream_t s, const u32 *const *src, const int *words, const int *dst_off, int n, u32 *host_block)
{
	PublishArgs a;
	a.n = n;
	for (int i = 0; i < n; i++) { a.src[i] = src[i]; a.words[i] = words[i]; a.dst_off[i] = dst_off[i]; }
	ug_launch(UG_PDL_PUBLISH, k_publish, dim3(1), dim3(64), s, a, host_block);
}

/* the frame's constant block comes from a ring of pinned (device-visible) host slots: a one-block kernel pulls the ~6 KB
 * of slot (seq % slots) over PCIe, stores them in device memory and advances the device-side sequence counter.  A kernel
 * node starts sooner than a copy-engine node at the head of the frame graph, and because the slot is chosen on the device
 * the captured graph is launched unchanged while the host fills the next slot (no host synchronisation per frame). */
static __device__ __forceinline__ unsigned long long ug_globaltimer()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}
__global__ void k_fetch_consts(const uint4 *__restrict__ host_ring, uint4 *__restrict__ dev_block, int n16, u32 *seq, int slots, unsigned long long *t_start)
{
	__shared__ u32 s_seq;
	if (threadIdx.x == 0) { s_seq = *seq; *t_start = ug_globaltimer(); /* the frame's first kernel: where ug_stats.ms_own starts */ }
	__syncthreads();
	const uint4 *src = host_ring + (size_t)(s_seq % (u32)slots) * (size_t)n16;
	for (int i = threadIdx.x; i < n16; i += blockDim.x) dev_block[i] = src[i];
	if (threadIdx.x == 0) *seq = s_seq + 1u;
}
void ugk_fetch_consts(cudaStream_t s, const void *host_ring, void *dev_block, size_t bytes, u32 *seq, int slots, unsigned long long *t_start)
{
	k_fetch_consts<<<1, 512, 0, s>>>((const uint4 *)host_ring, (uint4 *)dev_block, (int)(bytes / 16), seq, slots, t_start);
}

/* ---- the gather's handshakes as flags in rank 0's IPC-mapped buffer (the page behind the image slots) --------------------
 * Every rank counts its gathered frames in a device word of its own (seq); all ranks render the same sequence of gathered
 * frames, so the counts agree without being exchanged.
 *   ready   rank 0: "my stream is past the read-out of the frame these stores will overwrite" = flags->ready := seq (release).
 *           others: spin on rank 0's word over NVLink until it says seq (acquire) - issued on a side branch after the camera
 *           grid, needed only in front of the first remote store.
 *   done    others, after their last remote store: system fence, flags->done[rank] := seq | overflow << 31, then spin on
 *           flags->result.  rank 0: one lane per rank spins on done[] (local memory, written by the peers), ORs the overflow
 *           bits and publishes flags->result := seq | any << 31 - when that kernel retires the image is whole.
 * Two stores and one round trip over NVLink per frame replace two 8-rank ncclAllReduce launches (~25 us each on the frame's
 * tail, plus an SM per NCCL kernel while it waits for the slowest rank). */
struct CommFlags { u32 ready, result, pad[14]; u32 done[UG_MAX_RANKS]; };
static_assert(sizeof(CommFlags) <= 4096, "the flags live in one page behind rank 0's gather buffer");
static __device__ __forceinline__ u32 ld_acquire_sys(const u32 *p)
{
	u32 v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
static __device__ __forceinline__ void st_release_sys(u32 *p, u32 v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
static __device__ __forceinline__ bool seq_reached(u32 word, u32 seq) { return (int)((word & 0x7fffffffu) - seq) >= 0; }

__global__ void k_comm_ready(CommFlags *flags, u32 *seq, int rank)
{
	const u32 s = (*seq + 1u) & 0x7fffffffu;
	*seq = s;
	if (rank == 0) { __threadfence_system(); st_release_sys(&flags->ready, s); }
	else while (!seq_reached(ld_acquire_sys(&flags->ready), s)) __nanosleep(100);
}
__global__ void k_comm_done(CommFlags *flags, const u32 *seq, int rank, int world, u32 *over_word, const unsigned long long *t_start, u32 *own_ns)
{
	const u32 s = *seq, lane = threadIdx.x;
	if (lane == 0 && own_ns) *own_ns = (u32)(ug_globaltimer() - *t_start); /* this rank's own work ends here: what follows is waiting for the others */
	u32 over = over_word ? *over_word : 0u; /* in: this rank's pair arenas overflowed; out: anybody's did */
	if (rank != 0) {
		if (lane == 0) {
			__threadfence_system(); /* this rank's remote pixel stores (earlier kernels of the stream) before its flag */
			st_release_sys(&flags->done[rank], s | (over ? 0x80000000u : 0u));
			u32 w;
			while (!seq_reached(w = ld_acquire_sys(&flags->result), s)) __nanosleep(100);
			if (over_word) *over_word = w >> 31;
		}
		return;
	}
	if ((int)lane > 0 && (int)lane < world) {
		u32 w;
		while (!seq_reached(w = ld_acquire_sys(&flags->done[lane]), s)) __nanosleep(20);
		over |= w >> 31;
	}
	for (int o = 16; o > 0; o >>= 1) over |= __shfl_xor_sync(0xffffffffu, over, o);
	if (world > 32) { /* lanes beyond the warp's width: the first lane walks the rest */
		if (lane == 0)
			for (int r = 32; r < world; r++) {
				u32 w;
				while (!seq_reached(w = ld_acquire_sys(&flags->done[r]), s)) __nanosleep(20);
				over |= w >> 31;
			}
		over = __shfl_sync(0xffffffffu, over, 0) | over;
	}
	if (lane == 0) {
		if (over_word) *over_word = over ? 1u : 0u;
		__threadfence_system();
		st_release_sys(&flags->result, s | (over ? 0x80000000u : 0u));
	}
}
void ugk_comm_ready(cudaStream_t s, void *flags, u32 *seq, int rank) { k_comm_ready<<<1, 1, 0, s>>>((CommFlags *)flags, seq, rank); }
void ugk_comm_done(cudaStream_t s, void *flags, const u32 *seq, int rank, int world, u32 *over_word, const unsigned long long *t_start, u32 *own_ns)
{
	k_comm_done<<<1, 32, 0, s>>>((CommFlags *)flags, seq, rank, world, over_word, t_start, own_ns);
}
