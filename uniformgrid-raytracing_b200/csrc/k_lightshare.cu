/* k_lightshare.cu -- strip mode: a light grid is built ONCE per communicator, not once per rank.
 *
 * In the image-strip sharding of SURVEY.md 8(e) the camera grid shrinks with the strip, but shadow rays leave the strip, so every
 * rank needs every light grid in full (frustum_grid.h:210-366 for the light's view) - at 8 GPUs that replicated build is the
 * largest piece of work that does not shrink (config D: 170 us of the whole GPU per frame, beside a 105 us camera grid).
 * Here rank q builds the cell rows [nby*q/world, nby*(q+1)/world) of the light grid with the ordinary strip build
 * (k_gridsweep.cu with the light's gy_lo / gy_cnt) and this kernel stores the result into the same place of every rank's copy:
 *
 *   span / offset   C_q words each at the shard's cell base (cells are numbered shard-major: nbx*b_q + column*rows_q + row - b_q,
 *                   light_cell_of uses the same numbering); offsets already carry the shard's slot base (SweepArgs::offset_bias)
 *   vals            P_q triangle ids into slot q of the value list (the build itself wrote this rank's own copy)
 *
 * as 16-byte P2P stores over NVLink, then a system fence and one release store per peer of this build's sequence number into
 * the peer's arrival flag [light][q].  k_light_wait (one lane per rank, acquire loads of local memory) lets the consumers -
 * the item table of k_regroup and k_shadow - start.  Nothing reads a remote buffer.
 * A frame cannot overwrite lists a slower rank is still reading: every rank ends its frame in the gather's "done" handshake,
 * which is behind its own shadow rays.  */
#include "k_common.cuh"

#define LS_THREADS 256
struct LightPushArgs {
	u32 *peer[UG_MAX_RANKS];  /* every rank's buffer */
	const u32 *span, *offset, *vals, *counters;
	u32 *seq;                 /* [0] builds pushed so far, [1] blocks of this launch that are done */
	size_t span_w, offset_w, vals_w; /* word offsets of this light's arrays inside a buffer */
	u32 slot;                 /* pair capacity of one rank's slot */
	int light, rank, world, cells, cell_base;
};

static __device__ __forceinline__ void ls_st_release_sys(u32 *p, u32 v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
static __device__ __forceinline__ u32 ls_ld_acquire_sys(const u32 *p)
{
	u32 v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__global__ void __launch_bounds__(LS_THREADS) k_light_push(const LightPushArgs a)
{
	__shared__ u32 s_last;
	const u32 gtid = blockIdx.x * LS_THREADS + threadIdx.x, gsz = gridDim.x * LS_THREADS;
	const u32 P = min(a.counters[0], a.slot);
	const u32 P4 = P / 4;
	const uint4 *src4 = (const uint4 *)a.vals; /* slot bases are multiples of 4 words */
	for (int k = 0; k < a.world; k++) {
		const int p = (a.rank + 1 + k) % a.world; /* every rank starts with a different peer */
		u32 *dst = a.peer[p];
		for (u32 i = gtid; i < (u32)a.cells; i += gsz) {
			dst[a.span_w + (size_t)a.cell_base + i] = a.span[i];
			dst[a.offset_w + (size_t)a.cell_base + i] = a.offset[i];
		}
		if (p == a.rank) continue; /* the build wrote this rank's own slot */
		u32 *dv = dst + a.vals_w + (size_t)a.rank * a.slot;
		for (u32 i = gtid; i < P4; i += gsz) ((uint4 *)dv)[i] = src4[i];
		for (u32 i = 4 * P4 + gtid; i < P; i += gsz) dv[i] = a.vals[i];
	}
	__threadfence_system(); /* every thread's remote stores before the block reports */
	__syncthreads();
	if (threadIdx.x == 0) s_last = (atomicAdd(&a.seq[1], 1u) == gridDim.x - 1) ? 1u : 0u;
	__syncthreads();
	if (!s_last) return;
	__threadfence_system();
	if (threadIdx.x == 0) { s_last = a.seq[0] + 1u; a.seq[0] = s_last; a.seq[1] = 0; } /* one reader of the old count: the other warps take the new one from shared memory */
	__syncthreads();
	const u32 s = s_last;
	for (int p = threadIdx.x; p < a.world; p += LS_THREADS) ls_st_release_sys(a.peer[p] + (size_t)a.light * UG_MAX_RANKS + a.rank, s);
}

/* flags: this rank's own page; seq[0] was advanced by this rank's push of the same build (earlier in the stream) */
__global__ void k_light_wait(const u32 *flags, const u32 *seq, int light, int world)
{
	const u32 s = seq[0];
	for (int r = threadIdx.x; r < world; r += blockDim.x)
		while ((int)(ls_ld_acquire_sys(flags + (size_t)light * UG_MAX_RANKS + r) - s) < 0) __nanosleep(40);
}

void ugk_light_push(cudaStream_t s, ug_context *ctx, int l, const GridBuffers &built, int cell_base, int cells, int sms)
{
	LightPushArgs a;
	for (int r = 0; r < UG_MAX_RANKS; r++) a.peer[r] = r < ctx->comm_world ? (u32 *)ctx->lg_peer[r] : nullptr;
	a.span = built.span; a.offset = built.offset; a.vals = built.vals[built.final_buf]; a.counters = built.counters;
	a.seq = ctx->d_lg_seq + 2 * l;
	ug_light_layout(ctx, l, &a.span_w, &a.offset_w, &a.vals_w);
	a.slot = (u32)ctx->lg_slot;
	a.light = l; a.rank = ctx->comm_rank; a.world = ctx->comm_world; a.cells = cells; a.cell_base = cell_base;
	k_light_push<<<(unsigned)sms, LS_THREADS, 0, s>>>(a);
}
void ugk_light_wait(cudaStream_t s, ug_context *ctx, int l)
{
	k_light_wait<<<1, 64, 0, s>>>((const u32 *)ctx->lg_local, ctx->d_lg_seq + 2 * l, l, ctx->comm_world);
}
