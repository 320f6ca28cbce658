/* context.cu -- host side of libugrt.so: lifetime, the Model, persistent device arenas, the per-frame
 * sequencing that display() (main.cu:59-302) does in the reference, CUDA-graph capture of the frame,
 * read-backs.  Mirrors per_frame_funcs.h / frustum_grid.h / frustum_tracer.h / decision_data.h /
 * shader.h call for call, minus every cudaThreadSynchronize(), D2H scalar read-back, per-frame
 * cudaMalloc/cudaFree and cudppPlan of the reference (frustum_grid.h:221-272,285-290,317-337).
 */
#include "k_common.cuh"
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

int ug_fail(ug_context *ctx, int code, const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (ctx) ctx->err = buf;
	else fprintf(stderr, "ugrt: %s\n", buf);
	return code;
}

template <class T> static cudaError_t dmalloc(T **p, size_t count)
{
	if (*p) { cudaFree(*p); *p = nullptr; }
	return cudaMalloc((void **)p, sizeof(T) * (count ? count : 1));
}
template <class T> static void dfree(T *&p) { if (p) cudaFree(p); p = nullptr; }

static int bits_for(unsigned max_key_exclusive) /* number of key bits needed for keys in [0, max) */
{
	int b = 1;
	while (b < 32 && (1ull << b) < max_key_exclusive) b++;
	return b;
}

/* ------------------------------------------------------------------ lifetime */
#if UG_PDL
extern "C" const char *ug_version(void) { return "ugrt 0.2 (sm_100a, pdl)"; } /* built with -DUG_PDL=1: programmatic dependent launches along the frame's chain */
#else
extern "C" const char *ug_version(void) { return "ugrt 0.2 (sm_100a)"; }
#endif
extern "C" void ug_acosf_host(const float *x, float *y, size_t n)
{
	for (size_t i = 0; i < n; i++) y[i] = ug_acosf(x[i]);
}
extern "C" size_t ug_frame_const_bytes(void) { return sizeof(FrameConst); }

extern "C" int ug_create(ug_context **out, int device)
{
	if (!out) return UG_ERR_ARG;
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count <= 0)
		return ug_fail(nullptr, UG_ERR_CUDA, "no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(e));
	if (device < 0 || device >= count) return ug_fail(nullptr, UG_ERR_ARG, "device %d out of range (%d devices)", device, count);
	ug_context *ctx = new ug_context();
	ctx->device = device;
	UG_CUDA(ctx, cudaSetDevice(device));
	cudaDeviceProp prop;
	UG_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
	ctx->num_sms = prop.multiProcessorCount;
	/* One-launch cooperative builds win where the build is alone on the critical path (small scenes, one light); with
	 * more branches in flight (two lights, reflection) they cannot all be resident and the kernel chains, whose short
	 * kernels interleave with the other branches, come out ahead - measured per config in DESIGN.md section 7.  So the
	 * grid build is chosen per frame (fused_grid_now), the other two are opt-in.  UG_NO_FUSED / UG_FUSED_* override. */
	ctx->use_fused = ugk_grid_fused_supported(device) && !getenv("UG_NO_FUSED") && !getenv("UG_NO_FUSED_GRID");
	/* the default build of every perspective / spherical grid: three one-sweep launches (k_gridsweep.cu).  UG_NO_SWEEP, or any
	 * switch that forces one of the other two variants, turns it off. */
	ctx->use_sweep = !getenv("UG_NO_SWEEP") && !getenv("UG_NO_FUSED") && !getenv("UG_FUSED_GRID");
	ctx->direct_shadow = getenv("UG_SHADOW_DIRECT") != nullptr; /* measured slower than the regrouped path on every config: opt-in */
	ctx->fused_gather = !getenv("UG_NO_FUSED_GATHER");
	ctx->use_fused_regroup = ugk_regroup_fused_supported(device) && !getenv("UG_NO_FUSED") && getenv("UG_FUSED_REGROUP");
	ctx->use_fused_ugrid = ugk_ugrid_fused_supported(device) && !getenv("UG_NO_FUSED") && getenv("UG_FUSED_UGRID");
	ctx->force_fused_grid = getenv("UG_FUSED_GRID") != nullptr;
	if (getenv("UG_TIMELINE")) {
		UG_CUDA(ctx, cudaMalloc((void **)&ctx->d_stamps, 64 * sizeof(unsigned long long)));
		UG_CUDA(ctx, cudaMemset(ctx->d_stamps, 0, 64 * sizeof(unsigned long long)));
	}
	if (const char *e = getenv("UG_REFL_CELLS_PER_TRI")) {
		const float v = (float)atof(e);
		if (v >= 0.25f && v <= 256.0f) ctx->refl_cells_per_tri = v;
	}
	/* Priorities follow the slack of each branch of the frame graph (the graph is instantiated with per-node priorities): the
	 * camera grid, the primary rays and the reflection rays are the critical path (greatest); a light grid or the uniform grid is
	 * not needed before the primary rays are done (least).  Measured on the B200 this changes nothing (config D as 8 strips:
	 * camera grid done at 178 us beside the light-grid build with or without priorities, ~105 us alone) - the build kernels
	 * are persistent, all their blocks are resident at once and share the SMs' issue slots whatever the priority; what
	 * does help is giving the non-critical build fewer blocks (UG_LIGHT_BLOCKS).  UG_NO_PRIO: all equal. */
	int prio_least = 0, prio_greatest = 0;
	cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
	const bool prio = !getenv("UG_NO_PRIO");
	ctx->light_build_sms = -1; /* strip mode: two thirds of the SMs (set once the device is known); UG_LIGHT_BLOCKS=0: all of them */
	if (const char *e = getenv("UG_LIGHT_BLOCKS")) ctx->light_build_sms = atoi(e);
	UG_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->own_stream, cudaStreamNonBlocking, prio ? prio_greatest : prio_least));
	ctx->stream = ctx->own_stream;
	for (int i = 0; i < UG_MAX_LIGHTS; i++) {
		UG_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side[i], cudaStreamNonBlocking, prio_least));
		UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming));
		UG_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side_map[i], cudaStreamNonBlocking, prio ? prio_greatest : prio_least));
		UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_map[i], cudaEventDisableTiming));
	}
	UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
	UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork2, cudaEventDisableTiming));
	UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_init, cudaEventDisableTiming));
	UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_preshade, cudaEventDisableTiming));
	UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_pub, cudaEventDisableTiming));
	UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_ugrid, cudaEventDisableTiming));
	UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_camgrid, cudaEventDisableTiming));
	UG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
	UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_frame_done, cudaEventDisableTiming));
	UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copy_done, cudaEventDisableTiming));
	for (int i = 0; i < 4; i++) UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copy_ring[i], cudaEventDisableTiming));
	ctx->light_after_cam = getenv("UG_LIGHT_AFTER_CAM") != nullptr;
	if (const char *e = getenv("UG_DIR_RECOMPUTE")) ctx->dir_recompute = atoi(e) & 3;
	ctx->hoist_normals = !getenv("UG_NO_HOIST_NORMALS");
	if (const char *e = getenv("UG_TRACE_ITEM")) ctx->trace_item_forced = atoi(e) == 64 ? 64 : 128; /* tests force both item sizes of k_trace */
	UG_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side_init, cudaStreamNonBlocking, prio ? prio_greatest : prio_least));
	UG_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side_refl, cudaStreamNonBlocking, prio ? prio_greatest : prio_least));
	UG_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side_ugrid, cudaStreamNonBlocking, prio_least));
	UG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->side_comm, cudaStreamNonBlocking));
	UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
	UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_cam, cudaEventDisableTiming));
	UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_refl, cudaEventDisableTiming));
	UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_trace, cudaEventDisableTiming));
	UG_CUDA(ctx, cudaEventCreate(&ctx->ev_t0));
	UG_CUDA(ctx, cudaEventCreate(&ctx->ev_t1));
	for (auto &ev : ctx->ev_stage) UG_CUDA(ctx, cudaEventCreate(&ev));
	ctx->h_fc = (FrameConst *)calloc(1, sizeof(FrameConst));
	if (!ctx->h_fc) return ug_fail(ctx, UG_ERR_CUDA, "out of host memory");
	UG_CUDA(ctx, cudaMallocHost((void **)&ctx->h_ring, sizeof(FrameConst) * UG_FC_SLOTS));
	memset(ctx->h_ring, 0, sizeof(FrameConst) * UG_FC_SLOTS);
	for (int i = 0; i < UG_FC_SLOTS; i++) UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_ring[i], cudaEventDisableTiming));
	UG_CUDA(ctx, cudaMalloc((void **)&ctx->d_fc_seq, sizeof(u32)));
	UG_CUDA(ctx, cudaMemset(ctx->d_fc_seq, 0, sizeof(u32)));
	UG_CUDA(ctx, cudaMalloc((void **)&ctx->d_tstart, 16));
	UG_CUDA(ctx, cudaMemset(ctx->d_tstart, 0, 16));
	UG_CUDA(ctx, cudaMallocHost((void **)&ctx->h_counters, sizeof(u32) * UG_HC_WORDS));
	memset(ctx->h_counters, 0, sizeof(u32) * UG_HC_WORDS);
	UG_CUDA(ctx, cudaMalloc((void **)&ctx->d_fc, sizeof(FrameConst)));
	*out = ctx;
	return UG_OK;
}

static void free_grid(GridBuffers &g)
{
	dfree(g.tri_rec); dfree(g.rect); dfree(g.count); dfree(g.scan);
	dfree(g.keys[0]); dfree(g.keys[1]); dfree(g.vals[0]); dfree(g.vals[1]);
	dfree(g.span); dfree(g.offset); dfree(g.counters);
	dfree(g.sort.tile_hist); dfree(g.sort.tile_base); dfree(g.sort.scan_state);
	dfree(g.fused_hist); dfree(g.fused_tot); dfree(g.fused_state_a); dfree(g.fused_state_b);
	dfree(g.sweep_hist); dfree(g.sweep_base); dfree(g.sweep_ctl); dfree(g.sweep_tile_first); dfree(g.sweep_desc_a); dfree(g.sweep_desc_b); dfree(g.sweep_desc);
	dfree(g.projz); dfree(g.zrange); dfree(g.cspan); dfree(g.coffset);
	g = GridBuffers();
}
static void free_rays(LightRays &r)
{
	dfree(r.cell); dfree(r.pix); dfree(r.span); dfree(r.offset); dfree(r.rank);
	dfree(r.item_off); dfree(r.items); dfree(r.counters); dfree(r.state_a); dfree(r.state_b); dfree(r.heavy);
	r = LightRays();
}

static void free_reflect(ReflectState &r)
{
	free_grid(r.grid);
	dfree(r.box); dfree(r.geo); dfree(r.tbox); dfree(r.prec); dfree(r.cellrec);
	r.prec_cap = 0; dfree(r.rect); dfree(r.O); dfree(r.r); dfree(r.term); dfree(r.w); dfree(r.k0); dfree(r.list[0]); dfree(r.list[1]); dfree(r.list0); dfree(r.depth);
	dfree(r.ray_counts);
	r.term_bounces = 0;
	if (r.uc) cudaFree(r.uc);
	r.uc = nullptr;
	r.N = 0;
}

extern "C" void ug_destroy(ug_context *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	ug_drop_graphs(ctx);
	ug_gl_unregister_pbo(ctx);
	ugc_destroy(ctx);
	dfree(ctx->d_rgb_full);
	free_reflect(ctx->refl);
	dfree(ctx->d_reflect);
	free_grid(ctx->cam_grid);
	for (int i = 0; i < UG_MAX_LIGHTS; i++) { free_grid(ctx->light_grid[i]); free_rays(ctx->rays[i]); }
	dfree(ctx->d_vertexlist); dfree(ctx->d_materiallist); dfree(ctx->d_vertex_orig_list); dfree(ctx->d_facelist); dfree(ctx->d_materiallist_index); dfree(ctx->d_tri_nv);
	dfree(ctx->d_t); dfree(ctx->d_id); dfree(ctx->d_shadowed); dfree(ctx->d_matid); dfree(ctx->d_best); dfree(ctx->d_cell_done); dfree(ctx->d_normal); dfree(ctx->d_dir); dfree(ctx->d_rgb_own);
	if (ctx->h_rgb_pinned) cudaFreeHost(ctx->h_rgb_pinned);
	if (ctx->h_full_pinned) cudaFreeHost(ctx->h_full_pinned);
	free(ctx->h_fc);
	if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
	for (int i = 0; i < UG_FC_SLOTS; i++) if (ctx->ev_ring[i]) cudaEventDestroy(ctx->ev_ring[i]);
	dfree(ctx->d_fc_seq);
	dfree(ctx->d_tstart);
	dfree(ctx->d_lg_seq);
	if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
	for (int i = 0; i < 2; i++) {
		if (ctx->h_vstage[i]) cudaFreeHost(ctx->h_vstage[i]);
		if (ctx->ev_vstage[i]) cudaEventDestroy(ctx->ev_vstage[i]);
	}
	dfree(ctx->d_fc);
	for (int i = 0; i < UG_MAX_LIGHTS; i++) { cudaStreamDestroy(ctx->side[i]); cudaEventDestroy(ctx->ev_join[i]); cudaStreamDestroy(ctx->side_map[i]); cudaEventDestroy(ctx->ev_map[i]); }
	cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_fork2); cudaEventDestroy(ctx->ev_init); cudaEventDestroy(ctx->ev_preshade); cudaEventDestroy(ctx->ev_pub); cudaStreamDestroy(ctx->side_init); cudaStreamDestroy(ctx->side_refl); cudaStreamDestroy(ctx->side_ugrid); cudaEventDestroy(ctx->ev_ugrid); cudaEventDestroy(ctx->ev_camgrid); cudaStreamDestroy(ctx->copy_stream); cudaEventDestroy(ctx->ev_frame_done); cudaEventDestroy(ctx->ev_copy_done); for (int i = 0; i < 4; i++) cudaEventDestroy(ctx->ev_copy_ring[i]); cudaStreamDestroy(ctx->side_comm); cudaEventDestroy(ctx->ev_ready); cudaEventDestroy(ctx->ev_cam); cudaEventDestroy(ctx->ev_refl); cudaEventDestroy(ctx->ev_trace); cudaEventDestroy(ctx->ev_t0); cudaEventDestroy(ctx->ev_t1);
	for (auto &ev : ctx->ev_stage) cudaEventDestroy(ev);
	cudaStreamDestroy(ctx->own_stream);
	delete ctx;
}

void ug_drop_graphs(ug_context *ctx)
{
	for (int k = 0; k < 2; k++)
		if (ctx->graph_exec[k]) { cudaGraphExecDestroy(ctx->graph_exec[k]); ctx->graph_exec[k] = nullptr; }
}

extern "C" const char *ug_last_error(const ug_context *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int ug_set_stream(ug_context *ctx, void *cuda_stream)
{
	if (!ctx) return UG_ERR_ARG;
	ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
	ug_drop_graphs(ctx);
	return UG_OK;
}

extern "C" int ug_set_rgb_buffer(ug_context *ctx, void *device_ptr, size_t bytes)
{
	if (!ctx) return UG_ERR_ARG;
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->d_rgb_ext = (unsigned char *)device_ptr;
	ctx->rgb_ext_bytes = device_ptr ? bytes : 0;
	if (!device_ptr) ctx->d_rgb = ctx->d_rgb_own;
	ug_drop_graphs(ctx);
	return UG_OK;
}

extern "C" int ug_synchronize(ug_context *ctx)
{
	if (!ctx) return UG_ERR_ARG;
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return UG_OK;
}

/* ------------------------------------------------------------------ arenas */
static int alloc_grid(ug_context *ctx, GridBuffers &g, int T, int C, size_t cap, int S = 0, bool view_grid = true)
{
	if (S <= 0) S = g.S > 0 ? g.S : 1; /* regrowing keeps the slab count; C counts (cell, slab) entries */
	if (g.T == T && g.C == C && g.S == S && g.cap >= cap && g.tri_rec) return UG_OK;
	const size_t keep_cap = g.cap > cap ? g.cap : cap;
	free_grid(g);
	g.T = T; g.C = C; g.cap = keep_cap; g.S = S; g.cells = C / S;
	if (S > 1) {
		UG_CUDA(ctx, dmalloc(&g.projz, (size_t)T));
		UG_CUDA(ctx, dmalloc(&g.zrange, 4));
		UG_CUDA(ctx, dmalloc(&g.cspan, (size_t)g.cells + 1));
		UG_CUDA(ctx, dmalloc(&g.coffset, (size_t)g.cells + 1));
	}
	g.key_bits = bits_for((unsigned)C);
	UG_CUDA(ctx, dmalloc(&g.tri_rec, 4 * (size_t)T));
	UG_CUDA(ctx, dmalloc(&g.rect, (size_t)T));
	UG_CUDA(ctx, dmalloc(&g.count, (size_t)T));
	UG_CUDA(ctx, dmalloc(&g.scan, (size_t)T + 1));
	for (int k = 0; k < 2; k++) {
		UG_CUDA(ctx, dmalloc(&g.keys[k], g.cap));
		UG_CUDA(ctx, dmalloc(&g.vals[k], g.cap));
	}
	UG_CUDA(ctx, dmalloc(&g.span, (size_t)C));
	UG_CUDA(ctx, dmalloc(&g.offset, (size_t)C + 1));
	UG_CUDA(ctx, dmalloc(&g.counters, 32));
	UG_CUDA(ctx, cudaMemset(g.counters, 0, 32 * sizeof(u32)));
	const size_t he = ugk_sort_hist_entries(ctx->num_sms);
	UG_CUDA(ctx, dmalloc(&g.sort.tile_hist, he));
	UG_CUDA(ctx, dmalloc(&g.sort.tile_base, he + 1));
	size_t biggest = (size_t)(T > C ? T : C) + 2;
	if (biggest < he + 2) biggest = he + 2;
	g.sort.scan_tiles = ugk_scan_tiles(biggest) + 1;
	UG_CUDA(ctx, dmalloc(&g.sort.scan_state, g.sort.scan_tiles + 1));
	{ /* fused build scratch; the look-back descriptors start zeroed and are re-zeroed by the kernel itself */
		const size_t na = ugk_grid_fused_state_a(T), nb = ugk_grid_fused_state_b(C);
		UG_CUDA(ctx, dmalloc(&g.fused_hist, ugk_grid_fused_hist(ctx->num_sms)));
		UG_CUDA(ctx, dmalloc(&g.fused_tot, 4 * 256));
		UG_CUDA(ctx, dmalloc(&g.fused_state_a, na));
		UG_CUDA(ctx, dmalloc(&g.fused_state_b, nb));
		UG_CUDA(ctx, cudaMemset(g.fused_state_a, 0, sizeof(u64) * na));
		UG_CUDA(ctx, cudaMemset(g.fused_state_b, 0, sizeof(u64) * nb));
	}
	/* three-launch build: sorts on (cell column, cell row), one digit each.  Descriptors are tagged with the build's sequence
	 * number, so they are cleared once, here. */
	if (view_grid && S == 1 && ctx->nbx > 0 && ctx->nbx <= UG_SWEEP_MAXB && ctx->nby <= UG_SWEEP_MAXB && g.cap < (1ull << 30) && C <= ctx->nbx * ctx->nby) {
		const size_t na = ugk_sweep_desc_a(T), nb = ugk_sweep_desc_b(C), nd = ugk_sweep_desc(g.cap, ctx->nbx > ctx->nby ? ctx->nbx : ctx->nby);
		UG_CUDA(ctx, dmalloc(&g.sweep_hist, ugk_sweep_hist_words()));
		UG_CUDA(ctx, dmalloc(&g.sweep_base, 2 * UG_SWEEP_MAXB));
		UG_CUDA(ctx, dmalloc(&g.sweep_ctl, 8));
		UG_CUDA(ctx, dmalloc(&g.sweep_tile_first, ugk_sweep_tiles(g.cap)));
		UG_CUDA(ctx, cudaMemset(g.sweep_tile_first, 0, sizeof(u32) * ugk_sweep_tiles(g.cap)));
		UG_CUDA(ctx, dmalloc(&g.sweep_desc_a, na));
		UG_CUDA(ctx, dmalloc(&g.sweep_desc_b, nb));
		UG_CUDA(ctx, dmalloc(&g.sweep_desc, nd));
		UG_CUDA(ctx, cudaMemset(g.sweep_hist, 0, sizeof(u32) * ugk_sweep_hist_words()));
		UG_CUDA(ctx, cudaMemset(g.sweep_ctl, 0, sizeof(u32) * 8));
		UG_CUDA(ctx, cudaMemset(g.sweep_desc_a, 0, sizeof(u64) * na));
		UG_CUDA(ctx, cudaMemset(g.sweep_desc_b, 0, sizeof(u64) * nb));
		UG_CUDA(ctx, cudaMemset(g.sweep_desc, 0, sizeof(u64) * nd));
	}
	return UG_OK;
}

static int alloc_rays(ug_context *ctx, LightRays &r, int N, int C)
{
	free_rays(r);
	UG_CUDA(ctx, dmalloc(&r.cell, (size_t)N));
	UG_CUDA(ctx, dmalloc(&r.counters, 4));
	UG_CUDA(ctx, cudaMemset(r.counters, 0, 4 * sizeof(u32)));
	if (ctx->direct_shadow) { /* rays stay in screen order: only the queue of long (ray group, list segment) products */
		r.heavy_cap = ugk_shadow_heavy_cap(N);
		UG_CUDA(ctx, dmalloc(&r.heavy, r.heavy_cap));
		return UG_OK;
	}
	UG_CUDA(ctx, dmalloc(&r.pix, (size_t)N));
	UG_CUDA(ctx, dmalloc(&r.span, (size_t)C + 1));
	UG_CUDA(ctx, dmalloc(&r.offset, (size_t)C + 2));
	UG_CUDA(ctx, dmalloc(&r.rank, (size_t)N));
	UG_CUDA(ctx, dmalloc(&r.item_off, (size_t)C + 2));
	r.item_cap = ugk_shadow_item_cap(N, C); /* worst case: one item per ray (every list longer than 1024) + the segment budget of every cell */
	UG_CUDA(ctx, dmalloc(&r.items, r.item_cap));
	const size_t ns = ugk_regroup_state_entries(C);
	UG_CUDA(ctx, dmalloc(&r.state_a, ns));
	UG_CUDA(ctx, dmalloc(&r.state_b, ns));
	UG_CUDA(ctx, cudaMemset(r.state_a, 0, sizeof(u64) * ns));
	UG_CUDA(ctx, cudaMemset(r.state_b, 0, sizeof(u64) * ns));
	return UG_OK;
}

/* the per-pair records follow the pair arena of the uniform grid (also when it regrows after an overflow) */
static int alloc_reflect_records(ug_context *ctx)
{
	ReflectState &r = ctx->refl;
	if (r.prec && r.prec_cap >= r.grid.cap) return UG_OK;
	dfree(r.prec);
	r.prec_cap = 0;
	UG_CUDA(ctx, dmalloc(&r.prec, r.grid.cap));
	r.prec_cap = r.grid.cap;
	return UG_OK;
}

static int alloc_reflect(ug_context *ctx, int T, int N, int bounces)
{
	ReflectState &r = ctx->refl;
	if (r.grid.T == T && r.N == N && r.geo) {
		if (bounces > r.term_bounces) { /* deeper recursion than before: one more [3N] term plane per bounce */
			dfree(r.term);
			r.term_bounces = 0;
			UG_CUDA(ctx, dmalloc(&r.term, 3 * (size_t)N * (size_t)bounces));
			r.term_bounces = bounces;
		}
		return UG_OK;
	}
	const size_t keep = r.grid.cap;
	free_reflect(r);
	long long cells = (long long)(ctx->refl_cells_per_tri * (float)T) + 1;
	if (cells < 4096) cells = 4096;
	if (cells > (1ll << 23)) cells = 1ll << 23;
	const size_t cap = (size_t)24 * (size_t)T + (size_t)cells + (1u << 20);
	int rc = alloc_grid(ctx, r.grid, T, (int)cells, keep > cap ? keep : cap, 0, false);
	if (rc) return rc;
	UG_CUDA(ctx, dmalloc(&r.box, 8));
	{ /* empty box in the order-preserving encoding; the fused build leaves it like this after every launch */
		const u32 empty[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u};
		UG_CUDA(ctx, cudaMemcpy(r.box, empty, sizeof(empty), cudaMemcpyHostToDevice));
	}
	UG_CUDA(ctx, cudaMalloc((void **)&r.uc, ugk_ugrid_const_bytes()));
	UG_CUDA(ctx, dmalloc(&r.geo, 3 * (size_t)T));
	UG_CUDA(ctx, dmalloc(&r.tbox, 2 * (size_t)T));
	UG_CUDA(ctx, dmalloc(&r.cellrec, (size_t)cells));
	rc = alloc_reflect_records(ctx);
	if (rc) return rc;
	UG_CUDA(ctx, dmalloc(&r.rect, (size_t)T));
	UG_CUDA(ctx, dmalloc(&r.O, 3 * (size_t)N));
	UG_CUDA(ctx, dmalloc(&r.r, 3 * (size_t)N));
	UG_CUDA(ctx, dmalloc(&r.term, 3 * (size_t)N * (size_t)bounces));
	r.term_bounces = bounces;
	UG_CUDA(ctx, dmalloc(&r.w, (size_t)N));
	UG_CUDA(ctx, dmalloc(&r.k0, (size_t)N));
	UG_CUDA(ctx, dmalloc(&r.list0, (size_t)N));
	UG_CUDA(ctx, dmalloc(&r.depth, (size_t)N));
	UG_CUDA(ctx, dmalloc(&r.list[0], (size_t)N));
	UG_CUDA(ctx, dmalloc(&r.list[1], (size_t)N));
	UG_CUDA(ctx, dmalloc(&r.ray_counts, UG_MAX_BOUNCES + 1));
	UG_CUDA(ctx, cudaMemset(r.ray_counts, 0, (UG_MAX_BOUNCES + 1) * sizeof(u32)));
	r.N = N;
	return UG_OK;
}

static size_t default_cap(int T, int C) { return (size_t)16 * (size_t)T + (size_t)C + (1u << 20); }

/* ---- light grids sharded over the ranks of a strip gather (k_lightshare.cu): layout of every rank's buffer, in 32-bit words:
 * flags page | per light: span[C] | offset[C] | vals[world * slot] */
static size_t lg_align(size_t words) { return (words + 63) & ~(size_t)63; }
static size_t lg_words_per_light(const ug_context *ctx) { return 2 * lg_align((size_t)ctx->lg_cells + 1) + (size_t)ctx->comm_world * ctx->lg_slot; }
void ug_light_layout(const ug_context *ctx, int l, size_t *span_w, size_t *offset_w, size_t *vals_w)
{
	const size_t Cw = lg_align((size_t)ctx->lg_cells + 1), base = UG_COMM_FLAGS_BYTES / 4 + (size_t)l * lg_words_per_light(ctx);
	*span_w = base; *offset_w = base + Cw; *vals_w = base + 2 * Cw;
}
/* rows [lo, hi) of a light grid that rank q of the communicator builds */
static void lg_rows(const ug_context *ctx, int q, int *lo, int *hi) { *lo = ctx->nby * q / ctx->comm_world; *hi = ctx->nby * (q + 1) / ctx->comm_world; }
/* the same answer on every rank: it only depends on what all ranks share */
static bool lg_eligible(const ug_context *ctx, int num_lights)
{
	if (!ctx->comm || ctx->comm_world <= 1 || ctx->lg_disabled || num_lights <= 0) return false;
	if (ctx->strip_count != ctx->comm_world || ctx->strip_index != ctx->comm_rank) return false; /* frames that end in the strip gather */
	if (!ctx->use_push || !ctx->flag_handshake) return false; /* needs the mapped buffers and the frame-end handshake every rank waits in */
	if (!ctx->use_sweep || ctx->num_slabs != 1 || ctx->direct_shadow || ctx->use_fused_regroup || ctx->nby < ctx->comm_world) return false;
	for (int l = 0; l < num_lights; l++)
		if (!ctx->light_grid[l].sweep_desc) return false;
	return true;
}

/* make every buffer the current (model, screen, strip, lights, flags) needs exist */
static int ensure_buffers(ug_context *ctx, int num_lights, bool full, int bounces = 0)
{
	if (ctx->num_faces <= 0) return ug_fail(ctx, UG_ERR_ARG, "no model loaded");
	if (ctx->W <= 0) return ug_fail(ctx, UG_ERR_ARG, "screen size not set");
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	bool changed = false;
	const int T = ctx->num_faces;
	if (ctx->pix_alloc != (size_t)ctx->N || !ctx->d_t) {
		const size_t N = (size_t)ctx->N;
		UG_CUDA(ctx, dmalloc(&ctx->d_t, N));
		UG_CUDA(ctx, dmalloc(&ctx->d_id, N));
		UG_CUDA(ctx, dmalloc(&ctx->d_shadowed, N));
		UG_CUDA(ctx, dmalloc(&ctx->d_matid, N));
		UG_CUDA(ctx, dmalloc(&ctx->d_best, N));
		UG_CUDA(ctx, dmalloc(&ctx->d_dir, 3 * N));
		UG_CUDA(ctx, dmalloc(&ctx->d_rgb_own, 3 * N));
		if (ctx->h_rgb_pinned) cudaFreeHost(ctx->h_rgb_pinned);
		ctx->h_rgb_pinned = nullptr;
		UG_CUDA(ctx, cudaMallocHost((void **)&ctx->h_rgb_pinned, 3 * N));
		dfree(ctx->d_normal);
		ctx->full_alloc = false;
		ctx->pix_alloc = N;
		for (int l = 0; l < UG_MAX_LIGHTS; l++) free_rays(ctx->rays[l]); /* sized by N */
		changed = true;
	}
	const int S = ctx->num_slabs;
	if (ctx->cam_grid.T != T || ctx->cam_grid.C != ctx->C_cam * S || ctx->cam_grid.S != S || !ctx->cam_grid.tri_rec) {
		int rc = alloc_grid(ctx, ctx->cam_grid, T, ctx->C_cam * S, default_cap(T, ctx->C_cam * S), S);
		if (rc) return rc;
		UG_CUDA(ctx, dmalloc(&ctx->d_cell_done, (size_t)ctx->C_cam));
		UG_CUDA(ctx, cudaMemset(ctx->d_cell_done, 0, sizeof(u32) * (size_t)ctx->C_cam));
		changed = true;
	}
	for (int l = 0; l < num_lights; l++) {
		if (ctx->light_grid[l].T != T || ctx->light_grid[l].C != ctx->C_light * S || ctx->light_grid[l].S != S || !ctx->light_grid[l].tri_rec) {
			int rc = alloc_grid(ctx, ctx->light_grid[l], T, ctx->C_light * S, default_cap(T, ctx->C_light * S), S);
			if (rc) return rc;
			free_rays(ctx->rays[l]);
			changed = true;
		}
		if (!ctx->rays[l].cell) {
			int rc = alloc_rays(ctx, ctx->rays[l], ctx->N, ctx->C_light);
			if (rc) return rc;
			changed = true;
		}
	}
	{
		if (ctx->gl_resource && !ctx->gl_mapped) return ug_fail(ctx, UG_ERR_ARG, "the registered PBO is not mapped (ug_gl_map_pbo before the frame, main.cu:139)");
		unsigned char *want = (ctx->d_rgb_ext && ctx->rgb_ext_bytes >= 3 * (size_t)ctx->N) ? ctx->d_rgb_ext : ctx->d_rgb_own;
		if (ctx->d_rgb_ext && want != ctx->d_rgb_ext) return ug_fail(ctx, UG_ERR_ARG, "external RGB buffer holds %zu bytes, need %zu", ctx->rgb_ext_bytes, 3 * (size_t)ctx->N);
		if (want != ctx->d_rgb) { ctx->d_rgb = want; changed = true; }
	}
	if (bounces > 0 && (ctx->refl.grid.T != T || ctx->refl.N != ctx->N || !ctx->refl.geo || ctx->refl.term_bounces < bounces)) {
		int rc = alloc_reflect(ctx, T, ctx->N, bounces);
		if (rc) return rc;
		changed = true;
	}
	if (ctx->comm && ctx->comm_world > 1 && ctx->rgb_full_bytes != 2 * 3 * (size_t)ctx->W * ctx->H * (size_t)ctx->comm_world) {
		/* one gathered frame (strips) or one frame per rank (UG_FLAG_GATHER_FRAMES), times the two slots of a deferred gather;
		 * collective: every rank gets here in the same call */
		int rc = ugc_share_full(ctx, 2 * 3 * (size_t)ctx->W * ctx->H * (size_t)ctx->comm_world);
		if (rc) return rc;
		changed = true;
	}
	if (lg_eligible(ctx, num_lights)) {
		/* collective as well: one rank's slot of a light's value list has the capacity every rank's arena starts with (a shard
		 * holds about 1/world of the grid's pairs; a shard that does not fit ends the sharding, check_overflow) */
		const size_t slot = lg_align(default_cap(T, ctx->C_light));
		if (ctx->lg_slot != slot || ctx->lg_lights != num_lights || ctx->lg_cells != ctx->C_light || !ctx->lg_bytes) {
			ctx->lg_slot = slot; ctx->lg_lights = num_lights; ctx->lg_cells = ctx->C_light;
			const size_t bytes = UG_COMM_FLAGS_BYTES + 4 * (size_t)num_lights * lg_words_per_light(ctx);
			if (!ctx->d_lg_seq) UG_CUDA(ctx, cudaMalloc((void **)&ctx->d_lg_seq, sizeof(u32) * 2 * UG_MAX_LIGHTS));
			UG_CUDA(ctx, cudaMemset(ctx->d_lg_seq, 0, sizeof(u32) * 2 * UG_MAX_LIGHTS));
			int rc = ugc_share_light(ctx, bytes);
			if (rc) return rc;
			changed = true;
		}
	}
	if (full && !ctx->full_alloc) {
		UG_CUDA(ctx, dmalloc(&ctx->d_normal, 3 * (size_t)ctx->N));
		ctx->full_alloc = true;
		changed = true;
	}
	if (changed) ug_drop_graphs(ctx);
	return UG_OK;
}

/* ------------------------------------------------------------------ Model (scene.h) */
extern "C" int ug_model_set(ug_context *ctx, const float *vertexlist, int num_vertices, const int *facelist, int num_faces,
                            const int *materiallist_index, const float *materiallist, int num_materials)
{
	if (!ctx || !vertexlist || !facelist || num_vertices <= 0 || num_faces <= 0) return ug_fail(ctx, UG_ERR_ARG, "ug_model_set: bad arguments");
	for (size_t i = 0; i < 3 * (size_t)num_faces; i++)
		if (facelist[i] < 0 || facelist[i] >= num_vertices) return ug_fail(ctx, UG_ERR_ARG, "face %zu references vertex %d (have %d)", i / 3, facelist[i], num_vertices);
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->h_vertexlist.assign(vertexlist, vertexlist + 3 * (size_t)num_vertices);
	ctx->h_facelist.assign(facelist, facelist + 3 * (size_t)num_faces);
	if (materiallist_index) ctx->h_materiallist_index.assign(materiallist_index, materiallist_index + num_faces);
	else ctx->h_materiallist_index.assign(num_faces, -1);
	if (materiallist && num_materials > 0) ctx->h_materiallist.assign(materiallist, materiallist + (size_t)UG_MATERIAL_SIZE * num_materials);
	else { ctx->h_materiallist.clear(); num_materials = 0; }
	ctx->num_vertices = num_vertices; ctx->num_faces = num_faces; ctx->num_materials = num_materials;
	for (int k = 0; k < 3; k++) { ctx->bounds[k] = 9999.9f; ctx->bounds[3 + k] = -9999.9f; } /* scene.h:224-229 */
	for (size_t i = 0; i < 3 * (size_t)num_vertices; i++) {
		const int k = (int)(i % 3);
		if (vertexlist[i] < ctx->bounds[k]) ctx->bounds[k] = vertexlist[i];
		if (vertexlist[i] > ctx->bounds[3 + k]) ctx->bounds[3 + k] = vertexlist[i];
	}
	UG_CUDA(ctx, dmalloc(&ctx->d_vertexlist, 3 * (size_t)num_vertices + 4));
	UG_CUDA(ctx, dmalloc(&ctx->d_facelist, 3 * (size_t)num_faces));
	UG_CUDA(ctx, dmalloc(&ctx->d_materiallist_index, (size_t)num_faces));
	UG_CUDA(ctx, dmalloc(&ctx->d_tri_nv, (size_t)num_faces));
	UG_CUDA(ctx, dmalloc(&ctx->d_materiallist, (size_t)UG_MATERIAL_SIZE * (num_materials ? num_materials : 1)));
	UG_CUDA(ctx, cudaMemcpy(ctx->d_vertexlist, ctx->h_vertexlist.data(), sizeof(float) * 3 * (size_t)num_vertices, cudaMemcpyHostToDevice));
	UG_CUDA(ctx, cudaMemcpy(ctx->d_facelist, ctx->h_facelist.data(), sizeof(int) * 3 * (size_t)num_faces, cudaMemcpyHostToDevice));
	UG_CUDA(ctx, cudaMemcpy(ctx->d_materiallist_index, ctx->h_materiallist_index.data(), sizeof(int) * (size_t)num_faces, cudaMemcpyHostToDevice));
	if (num_materials)
		UG_CUDA(ctx, cudaMemcpy(ctx->d_materiallist, ctx->h_materiallist.data(), sizeof(float) * UG_MATERIAL_SIZE * (size_t)num_materials, cudaMemcpyHostToDevice));
	dfree(ctx->d_vertex_orig_list);
	dfree(ctx->d_reflect);
	ctx->h_reflect.clear();
	ctx->size_orig_list = ctx->offset_orig_list = 0;
	ug_drop_graphs(ctx);
	ctx->frame_valid = false;
	return UG_OK;
}

extern "C" int ug_model_load(ug_context *ctx, const char *obj_path, const char *material_file)
{
	if (!ctx || !obj_path) return ug_fail(ctx, UG_ERR_ARG, "ug_model_load: bad arguments");
	int nv = 0, nf = 0, nm = 0;
	std::vector<float> mats;
	/* per_app_funcs.h:33-34: materials first, then the model */
	if (material_file) {
		if (ug_material_parse(material_file, &nm, nullptr) != UG_OK) return ug_fail(ctx, UG_ERR_IO, "cannot read material file %s", material_file);
		mats.resize((size_t)UG_MATERIAL_SIZE * (nm ? nm : 1));
		ug_material_parse(material_file, &nm, mats.data());
	}
	if (ug_obj_parse(obj_path, &nv, &nf, nullptr, nullptr, nullptr) != UG_OK) return ug_fail(ctx, UG_ERR_IO, "cannot read %s", obj_path);
	if (nv <= 0 || nf <= 0) return ug_fail(ctx, UG_ERR_IO, "%s: %d vertices, %d faces", obj_path, nv, nf);
	std::vector<float> V(3 * (size_t)nv);
	std::vector<int> F(3 * (size_t)nf), M(nf);
	ug_obj_parse(obj_path, &nv, &nf, V.data(), F.data(), M.data());
	return ug_model_set(ctx, V.data(), nv, F.data(), nf, M.data(), nm ? mats.data() : nullptr, nm);
}

/* Keyframed scenes: Model::tmp_model (scene.h:70-120) re-reads f_<i>.obj every frame, overwrites h_vertexlist and does a
 * blocking cudaMemcpy of faces + vertices.  Here only the vertices move (the topology is fixed by contract - the
 * reference copies an unchanged h_facelist), through one of two pinned staging blocks, asynchronously on the frame's
 * stream: the caller may parse keyframe i+1 while frame i renders.  Pointers do not change, so the captured graph stays. */
extern "C" int ug_model_update_vertices(ug_context *ctx, const float *vertexlist, int num_vertices)
{
	if (!ctx || !vertexlist || ctx->num_vertices <= 0) return ug_fail(ctx, UG_ERR_ARG, "update_vertices: no model");
	if (num_vertices <= 0 || num_vertices > ctx->num_vertices)
		return ug_fail(ctx, UG_ERR_ARG, "update_vertices: %d vertices, the model holds %d", num_vertices, ctx->num_vertices);
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	const size_t n = 3 * (size_t)num_vertices;
	if (ctx->vstage_floats < 3 * (size_t)ctx->num_vertices) {
		for (int i = 0; i < 2; i++) {
			if (ctx->ev_vstage[i]) UG_CUDA(ctx, cudaEventSynchronize(ctx->ev_vstage[i]));
			if (ctx->h_vstage[i]) cudaFreeHost(ctx->h_vstage[i]);
			ctx->h_vstage[i] = nullptr;
			UG_CUDA(ctx, cudaMallocHost((void **)&ctx->h_vstage[i], sizeof(float) * 3 * (size_t)ctx->num_vertices));
			if (!ctx->ev_vstage[i]) UG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_vstage[i], cudaEventDisableTiming));
		}
		ctx->vstage_floats = 3 * (size_t)ctx->num_vertices;
	}
	const int b = ctx->vstage_next;
	ctx->vstage_next ^= 1;
	UG_CUDA(ctx, cudaEventSynchronize(ctx->ev_vstage[b])); /* the copy that last used this block (two updates ago) */
	float lo[3] = {9999.9f, 9999.9f, 9999.9f}, hi[3] = {-9999.9f, -9999.9f, -9999.9f}; /* scene.h:92-97 */
	float *st = ctx->h_vstage[b];
	for (size_t i = 0; i < n; i++) {
		const float v = vertexlist[i];
		st[i] = v;
		const int k = (int)(i % 3);
		if (v < lo[k]) lo[k] = v;
		if (v > hi[k]) hi[k] = v;
	}
	for (int k = 0; k < 3; k++) { ctx->bounds[k] = lo[k]; ctx->bounds[3 + k] = hi[k]; }
	std::copy(vertexlist, vertexlist + n, ctx->h_vertexlist.begin());
	UG_CUDA(ctx, cudaMemcpyAsync(ctx->d_vertexlist, st, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
	UG_CUDA(ctx, cudaEventRecord(ctx->ev_vstage[b], ctx->stream));
	ctx->frame_valid = false;
	return UG_OK;
}

extern "C" int ug_model_tmp_model(ug_context *ctx, const char *path, int i)
{
	if (!ctx || !path) return ug_fail(ctx, UG_ERR_ARG, "tmp_model: bad arguments");
	if (ctx->num_vertices <= 0) return ug_fail(ctx, UG_ERR_ARG, "tmp_model: no model (load the first keyframe with ug_model_load)");
	char file[4096];
	if (snprintf(file, sizeof(file), "%s/f_%d.obj", path, i) >= (int)sizeof(file)) return ug_fail(ctx, UG_ERR_ARG, "tmp_model: path too long"); /* scene.h:78-83 */
	int nv = 0;
	if (ug_obj_parse_vertices(file, &nv, nullptr, 0) != UG_OK) return ug_fail(ctx, UG_ERR_IO, "cannot read %s", file);
	if (nv <= 0 || nv > ctx->num_vertices) return ug_fail(ctx, UG_ERR_IO, "%s: %d vertices, the model holds %d", file, nv, ctx->num_vertices);
	std::vector<float> V(3 * (size_t)nv);
	ug_obj_parse_vertices(file, &nv, V.data(), nv);
	return ug_model_update_vertices(ctx, V.data(), nv);
}

extern "C" int ug_model_bounds(ug_context *ctx, float *min_max6)
{
	if (!ctx || !min_max6) return ug_fail(ctx, UG_ERR_ARG, "ug_model_bounds: bad arguments");
	memcpy(min_max6, ctx->bounds, sizeof(ctx->bounds));
	return UG_OK;
}

extern "C" int ug_model_set_reflect(ug_context *ctx, const float *reflect, int num_materials)
{
	if (!ctx || !reflect || num_materials != ctx->num_materials || num_materials <= 0) return ug_fail(ctx, UG_ERR_ARG, "ug_model_set_reflect: need %d values", ctx ? ctx->num_materials : 0);
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->h_reflect.assign(reflect, reflect + num_materials);
	UG_CUDA(ctx, dmalloc(&ctx->d_reflect, (size_t)num_materials));
	UG_CUDA(ctx, cudaMemcpy(ctx->d_reflect, reflect, sizeof(float) * (size_t)num_materials, cudaMemcpyHostToDevice));
	ug_drop_graphs(ctx);
	return UG_OK;
}

extern "C" int ug_model_init_orig_list(ug_context *ctx, int vertices_size, int vertices_offset)
{
	if (!ctx || ctx->num_vertices <= 0) return ug_fail(ctx, UG_ERR_ARG, "init_orig_list: no model");
	if (vertices_size < 0 || vertices_offset < 0 || vertices_offset + vertices_size > ctx->num_vertices)
		return ug_fail(ctx, UG_ERR_ARG, "init_orig_list: range [%d,+%d) outside %d vertices", vertices_offset, vertices_size, ctx->num_vertices);
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	UG_CUDA(ctx, dmalloc(&ctx->d_vertex_orig_list, 3 * (size_t)vertices_size + 4));
	UG_CUDA(ctx, cudaMemcpy(ctx->d_vertex_orig_list, ctx->h_vertexlist.data() + 3 * (size_t)vertices_offset, sizeof(float) * 3 * (size_t)vertices_size,
	                        cudaMemcpyHostToDevice));
	ctx->size_orig_list = vertices_size;
	ctx->offset_orig_list = vertices_offset;
	ug_drop_graphs(ctx);
	return UG_OK;
}

static void set_rotation(ug_context *ctx, float rot)
{
	ctx->h_fc->rot_cos = cosf(rot); /* glibc, like the host-compiled reference (transformation_kernel.cu:14-15) */
	ctx->h_fc->rot_sin = sinf(rot);
}

/* The constant block travels through a ring of pinned slots (ugrt_internal.h): copy the master block into the slot the
 * next k_fetch_consts launch will read.  The slot was last read UG_FC_SLOTS uploads ago; its event has long fired unless the
 * caller runs that many frames ahead of the GPU, so this never blocks in a normal frame loop. */
static int stage_consts(ug_context *ctx, int *slot_out)
{
	const int slot = (int)(ctx->fc_seq % UG_FC_SLOTS);
	if (ctx->ring_used[slot]) UG_CUDA(ctx, cudaEventSynchronize(ctx->ev_ring[slot]));
	memcpy(&ctx->h_ring[slot], ctx->h_fc, sizeof(FrameConst));
	*slot_out = slot;
	return UG_OK;
}
/* after the launch (graph or direct) that contains exactly one k_fetch_consts */
static int consts_launched(ug_context *ctx, int slot)
{
	ctx->fc_seq++;
	ctx->ring_used[slot] = true;
	UG_CUDA(ctx, cudaEventRecord(ctx->ev_ring[slot], ctx->stream));
	return UG_OK;
}
static int upload_consts(ug_context *ctx)
{
	int slot = 0;
	int rc = stage_consts(ctx, &slot);
	if (rc) return rc;
	ugk_fetch_consts(ctx->stream, ctx->h_ring, ctx->d_fc, sizeof(FrameConst), ctx->d_fc_seq, UG_FC_SLOTS, ctx->d_tstart);
	UG_CUDA(ctx, cudaGetLastError());
	return consts_launched(ctx, slot);
}

extern "C" int ug_model_rotate_bunny(ug_context *ctx, float rot_factor)
{
	if (!ctx || !ctx->d_vertex_orig_list) return ug_fail(ctx, UG_ERR_ARG, "rotate_bunny: init_orig_list not called");
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	set_rotation(ctx, rot_factor);
	int rc = upload_consts(ctx);
	if (rc) return rc;
	ugk_transform(ctx->stream, ctx->d_vertexlist, ctx->d_vertex_orig_list, ctx->size_orig_list, ctx->offset_orig_list, ctx->d_fc, ctx->num_sms);
	UG_CUDA(ctx, cudaGetLastError());
	return UG_OK;
}

/* ------------------------------------------------------------------ screen / strip */
/* cell rows [lo, hi) of strip r: an equal split of the nby rows, or the caller's boundaries (ug_set_strip_bounds) */
void ug_strip_rows(const ug_context *ctx, int r, int *lo, int *hi)
{
	const int sc = ctx->strip_count > 0 ? ctx->strip_count : 1;
	if ((int)ctx->strip_bounds.size() == sc + 1 && ctx->strip_bounds.back() == ctx->nby) {
		*lo = ctx->strip_bounds[r];
		*hi = ctx->strip_bounds[r + 1];
		return;
	}
	*lo = (int)((long long)ctx->nby * r / sc);
	*hi = (int)((long long)ctx->nby * (r + 1) / sc);
}

static void update_geometry(ug_context *ctx)
{
	ctx->nbx = ctx->W / UG_TILE;
	ctx->nby = ctx->H / UG_TILE;
	int hi = 0;
	ug_strip_rows(ctx, ctx->strip_index, &ctx->gy_lo, &hi);
	ctx->gy_cnt = hi - ctx->gy_lo;
	ctx->C_cam = ctx->nbx * ctx->gy_cnt;
	ctx->C_light = ctx->nbx * ctx->nby;
	ctx->N = ctx->W * ctx->gy_cnt * UG_TILE;
}

extern "C" int ug_set_screen(ug_context *ctx, int width, int height, float fovy)
{
	if (!ctx || width <= 0 || height <= 0 || width % UG_TILE || height % UG_TILE || width / UG_TILE > 65535 || height / UG_TILE > 65535)
		return ug_fail(ctx, UG_ERR_ARG, "screen %dx%d must be positive multiples of %d", width, height, UG_TILE);
	if (ctx->W == width && ctx->H == height && ctx->fovy == fovy) return UG_OK;
	ctx->W = width; ctx->H = height; ctx->fovy = fovy;
	update_geometry(ctx);
	ctx->frame_valid = false;
	return UG_OK;
}

/* NUM_SLABS (main.cu.h:18) made runtime: every grid is built with that many depth slabs from now on */
extern "C" int ug_set_slabs(ug_context *ctx, int num_slabs)
{
	if (!ctx || num_slabs < 1 || num_slabs > UG_MAX_SLABS) return ug_fail(ctx, UG_ERR_ARG, "num_slabs %d not in [1,%d]", num_slabs, UG_MAX_SLABS);
	if (ctx->num_slabs != num_slabs) { ctx->num_slabs = num_slabs; ctx->frame_valid = false; }
	return UG_OK;
}

extern "C" int ug_set_strip(ug_context *ctx, int strip_index, int strip_count)
{
	if (!ctx || strip_count <= 0 || strip_index < 0 || strip_index >= strip_count) return ug_fail(ctx, UG_ERR_ARG, "bad strip %d/%d", strip_index, strip_count);
	if (ctx->strip_index == strip_index && ctx->strip_count == strip_count) return UG_OK;
	ctx->strip_index = strip_index; ctx->strip_count = strip_count;
	ctx->strip_bounds.clear(); /* a different split: equal strips again */
	if (ctx->W > 0) update_geometry(ctx);
	ctx->frame_valid = false;
	return UG_OK;
}

/* Strips of unequal height: row_bounds[0] = 0 < ... < row_bounds[strip_count] = cell rows of the screen.  Image-space cost is
 * not uniform (config D: the slowest of 8 equal strips takes 1.34x the mean), so a host that measures its ranks can move
 * the boundaries; every rank of a communicator must be given the same array.  ug_set_strip with another split forgets them. */
extern "C" int ug_set_strip_bounds(ug_context *ctx, const int *row_bounds, int strip_count, int strip_index)
{
	if (!ctx || !row_bounds || strip_count <= 0 || strip_index < 0 || strip_index >= strip_count) return ug_fail(ctx, UG_ERR_ARG, "bad strip %d/%d", strip_index, strip_count);
	if (ctx->W <= 0) return ug_fail(ctx, UG_ERR_ARG, "ug_set_strip_bounds: ug_set_screen first");
	if (row_bounds[0] != 0 || row_bounds[strip_count] != ctx->H / UG_TILE) return ug_fail(ctx, UG_ERR_ARG, "strip bounds must run from 0 to %d cell rows", ctx->H / UG_TILE);
	for (int r = 0; r < strip_count; r++)
		if (row_bounds[r + 1] <= row_bounds[r]) return ug_fail(ctx, UG_ERR_ARG, "strip %d is empty", r);
	ctx->strip_index = strip_index; ctx->strip_count = strip_count;
	ctx->strip_bounds.assign(row_bounds, row_bounds + strip_count + 1);
	update_geometry(ctx);
	ctx->frame_valid = false;
	return UG_OK;
}

/* ------------------------------------------------------------------ stage sequencing */
struct Launches { int n = 0; };

/* programmatic dependent launch along the frame's kernel chain (k_common.cuh, UG_PDL builds only): which kernel groups are
 * launched with the attribute while this thread enqueues a frame.  Off for a frame that ends in a gather (its spinning
 * handshake kernels were not measured with it), for per-stage event timing (an event between two kernels is not a
 * programmatic edge), with UG_NO_PDL=1; UG_PDL_MASK=<bits> picks groups (UG_PDL_* in k_common.cuh). */
thread_local unsigned ugk_pdl_mask = 0;
static unsigned pdl_mask_for(const ug_context *ctx, bool stage_events)
{
#if UG_PDL
	if (stage_events || getenv("UG_NO_PDL") || (ctx->comm && ctx->comm_world > 1)) return 0u;
	if (const char *e = getenv("UG_PDL_MASK")) return (unsigned)atoi(e) & (unsigned)UG_PDL_ALL;
	return (unsigned)UG_PDL_ALL;
#else
	(void)ctx; (void)stage_events;
	return 0u;
#endif
}
struct PdlScope { /* the mask applies to enqueue_frame only: the stage entry points launch the plain way */
	explicit PdlScope(unsigned m) { ugk_pdl_mask = m; }
	~PdlScope() { ugk_pdl_mask = 0; }
};
static bool gather_deferred(const ug_context *ctx, const ug_frame_params *p);
static int gather_slot(const ug_context *ctx, const ug_frame_params *p);

/* FrustumGrid::buildGrid, frustum_grid.h:210-366, on stream s */
static void enqueue_grid(ug_context *ctx, cudaStream_t s, const ViewConst *vc, GridBuffers &g, Launches &L, int part = 3)
{
	/* A strip's light grid is as large as the whole frame's while its camera grid is small and alone on the critical path: the
	 * light build gets the blocks of two thirds of the SMs, so that the camera grid's blocks meet less competition for issue
	 * slots (config D, strip 5 of 8: camera grid done at 147 us instead of 178, frame 364 us instead of 402; with a third
	 * of the SMs the light grid itself becomes the critical path).  UG_LIGHT_BLOCKS=<SMs>, 0 = all. */
	const int lsms = ctx->light_build_sms < 0 ? (2 * ctx->num_sms + 2) / 3 : ctx->light_build_sms;
	const int sms = (&g != &ctx->cam_grid && ctx->strip_count > 1 && lsms > 0) ? lsms : ctx->num_sms;
	const bool want_fused = (&g == &ctx->cam_grid) ? ctx->fused_cam_now : ctx->fused_grid_now;
	if (g.S > 1) { /* depth slabs: kernel chain only - the slab of a triangle needs the z range of ALL triangles before any pair is emitted */
		ugk_slab_reset(s, vc, g);                                                                       L.n += 1;
	} else if (ctx->use_sweep && g.sweep_desc) {
		if (&g != &ctx->cam_grid && ctx->lg_now) { /* this rank's cell rows only, then into every rank's copy (k_lightshare.cu) */
			const int l = (int)(&g - ctx->light_grid);
			int lo, hi;
			lg_rows(ctx, ctx->comm_rank, &lo, &hi);
			size_t span_w, offset_w, vals_w;
			ug_light_layout(ctx, l, &span_w, &offset_w, &vals_w);
			GridBuffers gl = g;
			gl.C = ctx->nbx * (hi - lo);
			gl.vals[1] = (u32 *)ctx->lg_local + vals_w + (size_t)ctx->comm_rank * ctx->lg_slot; /* pass 1 sorts straight into this rank's slot */
			gl.cap = g.cap < ctx->lg_slot ? g.cap : ctx->lg_slot;
			gl.offset_bias = (u32)((size_t)ctx->comm_rank * ctx->lg_slot);
			gl.final_buf = g.final_buf = ugk_grid_sweep(s, vc, ctx->d_facelist, ctx->d_vertexlist, gl, ctx->num_sms, &L.n, part);
			if (part & 2) {
				ugk_light_push(s, ctx, l, gl, ctx->nbx * lo, gl.C, ctx->num_sms);
				ugk_light_wait(s, ctx, l);
				L.n += 2;
			}
			return;
		}
		g.final_buf = ugk_grid_sweep(s, vc, ctx->d_facelist, ctx->d_vertexlist, g, sms, &L.n, part);
		return;
	} else if (ctx->use_fused && want_fused && g.key_bits <= 32) {
		const int fb = ugk_grid_fused(s, vc, ctx->d_facelist, ctx->d_vertexlist, g, sms);
		if (fb >= 0) { g.final_buf = fb; L.n += 1; return; }
		cudaGetLastError();   /* cooperative launch refused: fall back to the multi-kernel chain for good */
		ctx->use_fused = false;
	}
	ugk_setup_bin(s, vc, ctx->d_facelist, ctx->d_vertexlist, g.T, g, sms);                              L.n += 1;
	ugk_scan_counts(s, g, sms);                                                                         L.n += 1;
	ugk_emit_pairs(s, vc, g, sms);                                                                      L.n += 1;
	g.final_buf = ugk_radix_sort_pairs(s, g.keys, g.vals, g.counters, g.cap, g.key_bits, g.sort, sms, &L.n);
	ugk_cell_offsets(s, g, sms);                                                                        L.n += 1;
	if (g.S > 1 && &g != &ctx->cam_grid) { ugk_slab_union(s, g, sms);                                   L.n += 1; }
}

/* main.cu:148-203 for one light, after its ViewConst is resident: ray->cell map, light grid, regroup, shadow rays */
static void enqueue_light_grid(ug_context *ctx, cudaStream_t s, int l, Launches &L, int part = 3)
{
	enqueue_grid(ctx, s, &ctx->d_fc->light[l], ctx->light_grid[l], L, part);
}
/* a build that can be cut after its per-triangle launch */
static bool light_grid_splits(const ug_context *ctx, int l)
{
	const GridBuffers &g = ctx->light_grid[l];
	return ctx->light_after_cam && ctx->strip_count > 1 && g.S <= 1 && ctx->use_sweep && g.sweep_desc;
}
/* the light grid as its consumers see it: with sharded light grids, this rank's copy of the lists all ranks built together */
static GridBuffers light_view(const ug_context *ctx, int l)
{
	GridBuffers g = ctx->light_grid[l];
	if (ctx->lg_now) {
		size_t span_w, offset_w, vals_w;
		ug_light_layout(ctx, l, &span_w, &offset_w, &vals_w);
		u32 *base = (u32 *)ctx->lg_local;
		g.span = base + span_w; g.offset = base + offset_w; g.vals[g.final_buf] = base + vals_w;
	}
	return g;
}
static void enqueue_light_zero(ug_context *ctx, cudaStream_t s, int l)
{ /* independent of the primary rays: issued early on the light's branch */
	if (ctx->direct_shadow) return;
	LightRays &r = ctx->rays[l];
	cudaMemsetAsync(r.span, 0, sizeof(u32) * ((size_t)ctx->C_light + 1), s);
}
/* in_frame: the caller is enqueue_frame, so the eye's ViewConst in d_fc is the one the stored directions were made from
 * and phase A may evaluate them again instead of reading them (the stage entry points always read the buffer) */
static void enqueue_light_regroup(ug_context *ctx, cudaStream_t s, int l, Launches &L, int phases = 7, bool in_frame = false)
{
	if (ctx->direct_shadow) return; /* the rays are not regrouped: ray -> light cell is part of the shadow pass */
	LightRays &r = ctx->rays[l];
	const GridBuffers lg = light_view(ctx, l);
	const bool recomp = in_frame && (ctx->dir_recompute & 2);
	if (!ugk_regroup(s, ctx->d_fc, l, lg, r, ctx->d_t, ctx->d_dir, ctx->N, ctx->C_light, ctx->num_sms, ctx->use_fused_regroup, &L.n, phases, recomp)) {
		ctx->use_fused_regroup = false; /* cooperative launch refused: phase-by-phase launches from now on */
		ugk_regroup(s, ctx->d_fc, l, lg, r, ctx->d_t, ctx->d_dir, ctx->N, ctx->C_light, ctx->num_sms, false, &L.n, phases, recomp);
	}
}
static void enqueue_light_shadow(ug_context *ctx, cudaStream_t s, int l, Launches &L)
{
	if (ctx->direct_shadow) {
		ugk_shadow_direct(s, ctx->d_fc, l, ctx->light_grid[l], ctx->rays[l], ctx->d_t, ctx->d_dir, ctx->d_shadowed, ctx->N, ctx->C_light, ctx->num_sms, &L.n);
		return;
	}
	ugk_shadow(s, ctx->d_fc, l, light_view(ctx, l), ctx->rays[l], ctx->d_t, ctx->d_dir, ctx->d_shadowed, ctx->N, ctx->C_light, ctx->num_sms); L.n += 1;
}
static void enqueue_light_rays(ug_context *ctx, cudaStream_t s, int l, Launches &L)
{
	enqueue_light_regroup(ctx, s, l, L);
	enqueue_light_shadow(ctx, s, l, L);
}

static void stage_mark(ug_context *ctx, int &k, bool on)
{
	if (on) cudaEventRecord(ctx->ev_stage[k++], ctx->stream);
}

/* pair counts + overflow flags of every grid ride back with the frame (no extra synchronisation later) */
static void enqueue_publish(ug_context *ctx, cudaStream_t s, const ug_frame_params *p, bool gathered, bool deferred, Launches &L)
{
	const int nl = p->num_lights;
	const bool reflect = p->bounces > 0 && ctx->d_reflect;
	const u32 *src[UG_MAX_LIGHTS + 5];
	int words[UG_MAX_LIGHTS + 5], off[UG_MAX_LIGHTS + 5], n = 0;
	src[n] = ctx->cam_grid.counters; words[n] = 4; off[n++] = 0;
	for (int l = 0; l < nl; l++) { src[n] = ctx->light_grid[l].counters; words[n] = 4; off[n++] = 4 * (1 + l); }
	if (reflect) {
		src[n] = ctx->refl.grid.counters; words[n] = 4; off[n++] = 4 * (1 + UG_MAX_LIGHTS);
		src[n] = ctx->refl.ray_counts; words[n] = UG_MAX_BOUNCES + 1; off[n++] = 4 * (2 + UG_MAX_LIGHTS);
	}
	if (ctx->overflow_in_frame) { src[n] = ugc_overflow_word(ctx); words[n] = 1; off[n++] = UG_HC_WORDS - 1; }
	ctx->own_in_frame = gathered && ugc_own_word(ctx) && !deferred;
	if (ctx->own_in_frame) { src[n] = ugc_own_word(ctx); words[n] = 1; off[n++] = UG_HC_OWN; }
	ugk_publish(s, src, words, off, n, ctx->h_counters);
	L.n += 1;
}

/* the whole frame; `parallel` forks the light-grid builds onto side streams */
static int enqueue_frame(ug_context *ctx, const ug_frame_params *p, bool parallel, bool stage_events, Launches &L)
{
	cudaStream_t s = ctx->stream;
	const PdlScope pdl(pdl_mask_for(ctx, stage_events));
	const int sms = ctx->num_sms, nl = p->num_lights;
	const bool full = (p->flags & UG_FLAG_FULL_BUFFERS) != 0;
	const bool reflect = p->bounces > 0 && ctx->d_reflect;
	/* measured policy: the camera grid is alone on the critical path in front of the primary rays, so
	 * it is built by the one-launch kernel; a light grid only when nothing else (second light, uniform grid) runs beside it */
	ctx->fused_cam_now = ctx->force_fused_grid || ctx->num_faces <= 150000;
	ctx->fused_grid_now = ctx->force_fused_grid || (!reflect && nl <= 1 && ctx->num_faces <= 150000);
	const bool hoist = ctx->hoist_normals && ctx->d_tri_nv; /* shading normals once per triangle instead of once per pixel */
	const bool split_map = parallel && !ctx->direct_shadow && !ctx->use_fused_regroup && !getenv("UG_NO_SPLIT_MAP");
	int k = 0;
	static_assert(sizeof(FrameConst) % 16 == 0, "FrameConst is fetched in 16-byte words");
	ugk_fetch_consts(s, ctx->h_ring, ctx->d_fc, sizeof(FrameConst), ctx->d_fc_seq, UG_FC_SLOTS, ctx->d_tstart);
	L.n += 1;
	unsigned long long *const tl = ctx->d_stamps;
	ugk_stamp(s, tl, 0);
	stage_mark(ctx, k, stage_events); /* 0 */
	if (p->transform && ctx->d_vertex_orig_list) {
		ugk_transform(s, ctx->d_vertexlist, ctx->d_vertex_orig_list, ctx->size_orig_list, ctx->offset_orig_list, ctx->d_fc, sms);
		L.n += 1;
	}
	stage_mark(ctx, k, stage_events); /* 1: transform done */
	if (parallel && nl > 0) {
		UG_CUDA(ctx, cudaEventRecord(ctx->ev_fork, s));
		for (int l = 0; l < nl; l++) {
			UG_CUDA(ctx, cudaStreamWaitEvent(ctx->side[l], ctx->ev_fork, 0));
			if (l == 0) ugk_stamp(ctx->side[l], tl, 10);
			if (split_map) { /* the ray -> light cell map gets its own branch: it must not wait for the light grid */
				UG_CUDA(ctx, cudaStreamWaitEvent(ctx->side_map[l], ctx->ev_fork, 0));
				enqueue_light_zero(ctx, ctx->side_map[l], l);
			} else enqueue_light_zero(ctx, ctx->side[l], l);
			enqueue_light_grid(ctx, ctx->side[l], l, L, light_grid_splits(ctx, l) ? 1 : 3);
			if (l == 0 && !light_grid_splits(ctx, l)) ugk_stamp(ctx->side[l], tl, 11);
		}
	}
	const bool gather_strips = ctx->comm && ctx->comm_world > 1 && ctx->strip_count == ctx->comm_world && ctx->strip_index == ctx->comm_rank;
	const bool gather_frames = !gather_strips && ctx->comm && ctx->comm_world > 1 && (p->flags & UG_FLAG_GATHER_FRAMES) && ctx->strip_count == 1;
	if (parallel) { /* the per-pixel reset + ray directions do not depend on the grid: their own branch of the graph */
		UG_CUDA(ctx, cudaEventRecord(ctx->ev_fork2, s));
		if (reflect) { /* so does the world-space uniform grid of row R */
			UG_CUDA(ctx, cudaStreamWaitEvent(ctx->side_ugrid, ctx->ev_fork2, 0));
			ugk_stamp(ctx->side_ugrid, tl, 20);
			ugk_reflect_grid(ctx->side_ugrid, ctx, &L.n);
			ugk_stamp(ctx->side_ugrid, tl, 21);
			UG_CUDA(ctx, cudaEventRecord(ctx->ev_ugrid, ctx->side_ugrid));
		}
		UG_CUDA(ctx, cudaStreamWaitEvent(ctx->side_init, ctx->ev_fork2, 0));
		ugk_init_pixels(ctx->side_init, ctx->d_fc, ctx->d_t, ctx->d_id, ctx->d_shadowed, ctx->d_best, full ? ctx->d_normal : nullptr, ctx->d_dir, ctx->N, sms);
		if (hoist) ugk_shade_normals(ctx->side_init, ctx->d_fc, ctx->d_facelist, ctx->d_vertexlist, ctx->num_faces, ctx->d_tri_nv, sms); /* needed by k_shade only */
		UG_CUDA(ctx, cudaEventRecord(ctx->ev_init, ctx->side_init));
	} else {
		ugk_init_pixels(s, ctx->d_fc, ctx->d_t, ctx->d_id, ctx->d_shadowed, ctx->d_best, full ? ctx->d_normal : nullptr, ctx->d_dir, ctx->N, sms);
		if (hoist) ugk_shade_normals(s, ctx->d_fc, ctx->d_facelist, ctx->d_vertexlist, ctx->num_faces, ctx->d_tri_nv, sms);
	}
	L.n += hoist ? 2 : 1;
	enqueue_grid(ctx, s, &ctx->d_fc->cam, ctx->cam_grid, L);
	ugk_stamp(s, tl, 1);
	if (parallel && nl > 0) { /* held-back sort passes of the light grids: beside the primary rays instead of beside the camera grid */
		bool any = false;
		for (int l = 0; l < nl; l++) any = any || light_grid_splits(ctx, l);
		if (any) {
			UG_CUDA(ctx, cudaEventRecord(ctx->ev_camgrid, s));
			for (int l = 0; l < nl; l++) {
				if (!light_grid_splits(ctx, l)) continue;
				UG_CUDA(ctx, cudaStreamWaitEvent(ctx->side[l], ctx->ev_camgrid, 0));
				enqueue_light_grid(ctx, ctx->side[l], l, L, 2);
				if (l == 0) ugk_stamp(ctx->side[l], tl, 11);
			}
		}
	}
	if (parallel && (gather_strips || gather_frames) && ctx->use_push) {
		/* the gather's "ready" handshake hides behind the rest of the frame.  It starts only now: an NCCL kernel spins on an SM
		 * until the slowest rank arrives, and while it does the cooperative camera-grid build (all blocks co-resident or none)
		 * cannot start - issued at the head of the frame it put every rank's skew on the critical path */
		UG_CUDA(ctx, cudaEventRecord(ctx->ev_cam, s));
		UG_CUDA(ctx, cudaStreamWaitEvent(ctx->side_comm, ctx->ev_cam, 0));
		if (ctx->async_readback && ctx->comm_rank == 0) UG_CUDA(ctx, cudaStreamWaitEvent(ctx->side_comm, ctx->ev_copy_done, cudaEventWaitExternal)); /* "ready" = the last read-back is over as well */
		int rc = ugc_ready(ctx, ctx->side_comm, gather_deferred(ctx, p));
		if (rc) return rc;
		ugk_stamp(ctx->side_comm, tl, 6);
		UG_CUDA(ctx, cudaEventRecord(ctx->ev_ready, ctx->side_comm));
	}
	if (parallel) UG_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_init, 0));
	stage_mark(ctx, k, stage_events); /* 2: camera grid done */
	if (ctx->cam_grid.S > 1) ugk_trace_slabs(s, ctx->d_fc, ctx->cam_grid, ctx->d_dir, ctx->d_t, ctx->d_id, full ? ctx->d_normal : nullptr, ctx->C_cam, sms);
	else ugk_trace(s, ctx->d_fc, ctx->cam_grid, ctx->d_best, ctx->d_cell_done, ctx->d_dir, ctx->d_t, ctx->d_id, full ? ctx->d_normal : nullptr, sms, ctx->trace_small);
	L.n += 1;
	ugk_stamp(s, tl, 2);
	stage_mark(ctx, k, stage_events); /* 3: primary rays done */
	if (parallel && (nl > 0 || reflect)) UG_CUDA(ctx, cudaEventRecord(ctx->ev_trace, s));
	if (parallel && reflect) {
		/* row R needs the primary hits only: the bounces run beside the shadow rays (on the branch that built the
		 * uniform grid, so that is finished), and a short combine pass joins them with the shaded frame */
		UG_CUDA(ctx, cudaStreamWaitEvent(ctx->side_refl, ctx->ev_trace, 0));
		UG_CUDA(ctx, cudaStreamWaitEvent(ctx->side_refl, ctx->ev_ugrid, 0));
		ugk_stamp(ctx->side_refl, tl, 22);
		ugk_reflect_rays(ctx->side_refl, ctx, p->bounces, &L.n);
		ugk_stamp(ctx->side_refl, tl, 23);
		UG_CUDA(ctx, cudaEventRecord(ctx->ev_refl, ctx->side_refl));
	}
	if (parallel && nl > 0) {
		for (int l = 0; l < nl; l++) {
			if (split_map) {
				UG_CUDA(ctx, cudaStreamWaitEvent(ctx->side_map[l], ctx->ev_trace, 0));
				if (l == 0) ugk_stamp(ctx->side_map[l], tl, 12);
				enqueue_light_regroup(ctx, ctx->side_map[l], l, L, 1, true);
				UG_CUDA(ctx, cudaEventRecord(ctx->ev_map[l], ctx->side_map[l]));
				UG_CUDA(ctx, cudaStreamWaitEvent(ctx->side[l], ctx->ev_map[l], 0));
				enqueue_light_regroup(ctx, ctx->side[l], l, L, 6, true);
			} else {
				UG_CUDA(ctx, cudaStreamWaitEvent(ctx->side[l], ctx->ev_trace, 0));
				if (l == 0) ugk_stamp(ctx->side[l], tl, 12);
				enqueue_light_regroup(ctx, ctx->side[l], l, L, 7, true);
			}
			if (l == 0) ugk_stamp(ctx->side[l], tl, 13);
			enqueue_light_shadow(ctx, ctx->side[l], l, L);
			if (l == 0) ugk_stamp(ctx->side[l], tl, 14);
			UG_CUDA(ctx, cudaEventRecord(ctx->ev_join[l], ctx->side[l]));
			UG_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_join[l], 0));
		}
	} else {
		for (int l = 0; l < nl; l++) {
			enqueue_light_zero(ctx, s, l);
			enqueue_light_grid(ctx, s, l, L);
			stage_mark(ctx, k, stage_events); /* 4+3l: light grid */
			enqueue_light_regroup(ctx, s, l, L, 7, true);
			stage_mark(ctx, k, stage_events); /* 5+3l: ray -> light cell map + regroup */
			enqueue_light_shadow(ctx, s, l, L);
			stage_mark(ctx, k, stage_events); /* 6+3l: shadow rays */
		}
	}
	/* UG_FLAG_GATHER_DEFERRED: the frame lands in slot (sequence & 1) of rank 0's buffer and nobody waits for "done" here -
	 * the next frame's "ready" handshake (or ug_comm_flush) is what tells rank 0 that this one is whole */
	const bool deferred = gather_deferred(ctx, p);
	const size_t slot_off = deferred ? (size_t)gather_slot(ctx, p) * (ctx->rgb_full_bytes / 2) : 0;
	size_t g_bytes[UG_MAX_RANKS], g_offs[UG_MAX_RANKS];
	if (gather_strips) {
		size_t acc = 0;
		for (int r = 0; r < ctx->comm_world && r < UG_MAX_RANKS; r++) {
			int lo = 0, hi = 0;
			ug_strip_rows(ctx, r, &lo, &hi);
			g_bytes[r] = 3 * (size_t)ctx->W * (size_t)(hi - lo) * UG_TILE;
			g_offs[r] = acc;
			acc += g_bytes[r];
		}
	} else if (gather_frames) { /* frame-parallel rendering: frame r lands in slot r of the gathered buffer */
		for (int r = 0; r < ctx->comm_world && r < UG_MAX_RANKS; r++) { g_bytes[r] = 3 * (size_t)ctx->N; g_offs[r] = (size_t)r * g_bytes[r]; }
	}
	/* fused compute + collective: with the push transport the shading kernels store every pixel a second time, straight into
	 * rank 0's frame over NVLink (P2P stores), so the exchange overlaps the shading instead of following it as a copy */
	const bool fused_gather = (gather_strips || gather_frames) && ctx->use_push && ctx->fused_gather && parallel;
	ctx->rgb_remote = nullptr;
	if (fused_gather) {
		unsigned char *base = ctx->comm_rank == 0 ? ctx->d_rgb_full : ctx->peer_full;
		ctx->rgb_remote = base + slot_off + g_offs[ctx->comm_rank];
		UG_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_ready, 0)); /* rank 0 is past its read-out of what these stores overwrite */
	}
	/* A frame that is not gathered publishes its counters beside the shading instead of behind it: every grid is final once the
	 * shadow rays have joined, the reflection ray counts once that branch is done (one kernel and one dependent launch less
	 * on the frame's tail).  A gathered frame's publish also carries what the "done" handshake returns, so it stays last. */
	const bool early_publish = parallel && !(gather_strips || gather_frames) && !getenv("UG_LATE_PUBLISH");
	if (early_publish) {
		ctx->overflow_in_frame = false;
		UG_CUDA(ctx, cudaEventRecord(ctx->ev_preshade, s));
		UG_CUDA(ctx, cudaStreamWaitEvent(ctx->side_init, ctx->ev_preshade, 0));
		if (reflect) UG_CUDA(ctx, cudaStreamWaitEvent(ctx->side_init, ctx->ev_refl, 0));
		enqueue_publish(ctx, ctx->side_init, p, false, deferred, L);
		UG_CUDA(ctx, cudaEventRecord(ctx->ev_pub, ctx->side_init));
	}
	/* asynchronous read-back (ug_render_frame_host_async): the copy of the previous image runs on its own stream beside this
	 * frame; only the first kernel that rewrites the image waits for it (an external event-wait node of the graph) */
	if (ctx->async_readback) UG_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_copy_done, parallel ? cudaEventWaitExternal : 0));
	ugk_shade(s, ctx->d_fc, ctx->d_t, ctx->d_dir, ctx->d_id, ctx->d_shadowed, ctx->cam_grid.tri_rec, ctx->d_materiallist_index, ctx->d_materiallist, ctx->num_materials,
	          ctx->d_rgb, ctx->d_matid, ctx->N, 1, sms, (p->flags & UG_FLAG_SPOT_SHADE) != 0, ctx->rgb_remote, (ctx->dir_recompute & 1) != 0, hoist ? ctx->d_tri_nv : nullptr);
	L.n += 1;
	ugk_stamp(s, tl, 3);
	stage_mark(ctx, k, stage_events); /* shade */
	if (reflect) { /* row R: reflection bounces on top of the shaded frame */
		if (parallel) UG_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_refl, 0));
		else {
			ugk_reflect_grid(s, ctx, &L.n);
			stage_mark(ctx, k, stage_events); /* uniform grid */
			ugk_reflect_rays(s, ctx, p->bounces, &L.n);
			stage_mark(ctx, k, stage_events); /* reflection rays: spawn + one DDA kernel per bounce */
		}
		ugk_reflect_combine(s, ctx, &L.n);
		ugk_stamp(s, tl, 4);
		stage_mark(ctx, k, stage_events); /* combine */
		ctx->last_bounces = p->bounces;
	} else {
		ctx->last_bounces = 0;
	}
	if ((gather_strips || gather_frames) && ctx->use_push && !fused_gather) {
		if (parallel) UG_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_ready, 0));
		else { int rc = ugc_ready(ctx, s, gather_deferred(ctx, p)); if (rc) return rc; }
	}
	ctx->overflow_in_frame = false;
	if (gather_strips || gather_frames) {
		/* the one exchange step: RGB8 strips (or whole frames) -> rank 0 over NVLink, inside the frame */
		u32 *ow = ugc_overflow_word(ctx);
		if (ow && !deferred) { /* the "done" handshake also tells every rank whether any pair arena overflowed */
			const u32 *src[UG_MAX_LIGHTS + 3];
			int n = 0;
			src[n++] = ctx->cam_grid.counters;
			for (int l = 0; l < nl; l++) src[n++] = ctx->light_grid[l].counters;
			if (reflect) src[n++] = ctx->refl.grid.counters;
			ugk_overflow_any(s, src, n, ow);
			L.n += 1;
			ctx->overflow_in_frame = true;
		}
		int rc = ugc_gather_strips(ctx, s, ctx->d_rgb, slot_off, g_bytes, g_offs, deferred, fused_gather);
		if (rc) return rc;
		ugk_stamp(s, tl, 5);
		if (!(fused_gather && deferred)) L.n += 1;
	}
	if (early_publish) UG_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_pub, 0)); /* the counters went out beside the shading */
	else enqueue_publish(ctx, s, p, gather_strips || gather_frames, deferred, L);
	if (stage_events) ctx->stage_marks = k;
	UG_CUDA(ctx, cudaGetLastError());
	return UG_OK;
}

/* fill the pinned constant block from the frame parameters (host math only) */
static int prepare_frame(ug_context *ctx, const ug_frame_params *p)
{
	if (!p) return ug_fail(ctx, UG_ERR_ARG, "null frame params");
	if (p->num_lights < 0 || p->num_lights > UG_MAX_LIGHTS) return ug_fail(ctx, UG_ERR_ARG, "num_lights %d not in [0,%d]", p->num_lights, UG_MAX_LIGHTS);
	if (p->bounces < 0 || p->bounces > UG_MAX_BOUNCES) return ug_fail(ctx, UG_ERR_ARG, "bounces %d not in [0,%d]", p->bounces, UG_MAX_BOUNCES);
	int rc = ug_set_screen(ctx, p->width, p->height, p->fovy);
	if (rc) return rc;
	rc = ug_set_slabs(ctx, p->num_slabs > 0 ? p->num_slabs : 1);
	if (rc) return rc;
	rc = ug_set_strip(ctx, p->strip_count > 0 ? p->strip_index : 0, p->strip_count > 0 ? p->strip_count : 1);
	if (rc) return rc;
	rc = ensure_buffers(ctx, p->num_lights, (p->flags & UG_FLAG_FULL_BUFFERS) != 0, ctx->d_reflect ? p->bounces : 0);
	if (rc) return rc;
	FrameConst *fc = ctx->h_fc; /* the master copy: no frame in flight reads it (stage_consts copies it into a ring slot) */
	ctx->lg_now = lg_eligible(ctx, p->num_lights) && ctx->lg_ok && !(p->flags & (UG_FLAG_SPHERICAL_LIGHTS | UG_FLAG_GATHER_DEFERRED)) && !getenv("UG_NO_SHARD_LIGHTS");
	ug_camera cam;
	memset(&cam, 0, sizeof(cam));
	ug_camera_setCameraCenter(&cam, p->eye[0], p->eye[1], p->eye[2]);
	ug_camera_setCameraLookAt(&cam, p->look[0], p->look[1], p->look[2]);
	ug_camera_setCameraUp(&cam, p->up[0], p->up[1], p->up[2]);
	ug_camera_setNearFar(&cam, p->nearPlane, p->farPlane);
	cam.fovy = p->fovy;
	cam.aspect = (float)p->width / (float)p->height; /* camera.h:137 */
	ug_camera_adjustCameraAndPosition(&cam);
	ugh_view_const(&fc->cam, &cam, ctx->W, ctx->H, ctx->nbx, ctx->nby, ctx->gy_lo, ctx->gy_cnt, UG_LIGHT_GRID_PERSPECTIVE, 0.0f, 0.0f, ctx->num_slabs);
	memcpy(fc->shade_mv, cam.modelview_matrix, sizeof(fc->shade_mv));
	memcpy(fc->shade_eye, cam.worldori, sizeof(fc->shade_eye));
	for (int l = 0; l < p->num_lights; l++) {
		ug_camera lc;
		memset(&lc, 0, sizeof(lc));
		ug_camera_setCameraCenter(&lc, p->lights[l].pos[0], p->lights[l].pos[1], p->lights[l].pos[2]);
		ug_camera_setCameraLookAt(&lc, p->lights[l].look[0], p->lights[l].look[1], p->lights[l].look[2]);
		ug_camera_setCameraUp(&lc, p->lights[l].up[0], p->lights[l].up[1], p->lights[l].up[2]);
		ug_camera_setNearFar(&lc, p->nearPlane, p->farPlane);
		lc.fovy = p->fovy;
		lc.aspect = 1.0f;
		ug_camera_adjustCameraAndPosition(&lc);
		/* main.cu:186-187: x_max = y_max = M_PI, narrowed to the float parameters of the two calls that follow */
		int row_lo = 0, row_hi = ctx->nby;
		if (ctx->lg_now) lg_rows(ctx, ctx->comm_rank, &row_lo, &row_hi); /* this rank's share of the light grid (k_lightshare.cu) */
		ugh_view_const(&fc->light[l], &lc, ctx->W, ctx->H, ctx->nbx, ctx->nby, row_lo, row_hi - row_lo,
		               (p->flags & UG_FLAG_SPHERICAL_LIGHTS) ? UG_LIGHT_GRID_SPHERICAL : UG_LIGHT_GRID_PERSPECTIVE, (float)M_PI, (float)M_PI, ctx->num_slabs);
		fc->light[l].shards = ctx->lg_now ? ctx->comm_world : 1;
		for (int k = 0; k < 3; k++) fc->light_pos[3 * l + k] = p->lights[l].pos[k];
		/* shading runs with the LAST light camera's modelview resident in dd_camcoords (main.cu:205-211) */
		memcpy(fc->shade_mv, lc.modelview_matrix, sizeof(fc->shade_mv));
		memcpy(fc->shade_eye, lc.worldori, sizeof(fc->shade_eye));
	}
	for (int l = p->num_lights; l < UG_MAX_LIGHTS; l++)
		for (int k = 0; k < 3; k++) fc->light_pos[3 * l + k] = 0.0f;
	fc->num_lights = p->num_lights;
	ctx->num_lights = p->num_lights;
	set_rotation(ctx, p->rot_factor);
	return UG_OK;
}

/* does this frame end in a gather to rank 0, and is its completion deferred (UG_FLAG_GATHER_DEFERRED)? */
static bool gathers(const ug_context *ctx, const ug_frame_params *p)
{
	if (!ctx->comm || ctx->comm_world <= 1) return false;
	if (ctx->strip_count == ctx->comm_world && ctx->strip_index == ctx->comm_rank) return true;
	return (p->flags & UG_FLAG_GATHER_FRAMES) && ctx->strip_count == 1;
}
static bool gather_deferred(const ug_context *ctx, const ug_frame_params *p) { return gathers(ctx, p) && ctx->use_push && (p->flags & UG_FLAG_GATHER_DEFERRED); }
static int gather_slot(const ug_context *ctx, const ug_frame_params *p) { return gather_deferred(ctx, p) ? (int)(ctx->gather_seq & 1) : 0; }

static unsigned long long frame_key(const ug_context *ctx, const ug_frame_params *p)
{
	unsigned long long k = 1469598103934665603ull;
	auto mix = [&](unsigned long long v) { k = (k ^ v) * 1099511628211ull; };
	mix((unsigned long long)ctx->W); mix((unsigned long long)ctx->H); mix((unsigned long long)ctx->gy_lo); mix((unsigned long long)ctx->gy_cnt);
	for (int b : ctx->strip_bounds) mix((unsigned long long)b); /* the gather's offsets depend on every rank's strip */
	mix((unsigned long long)ctx->num_faces); mix((unsigned long long)p->num_lights); mix((unsigned long long)(p->flags & (UG_FLAG_FULL_BUFFERS | UG_FLAG_GATHER_FRAMES | UG_FLAG_SPOT_SHADE | UG_FLAG_GATHER_DEFERRED)));
	mix((unsigned long long)(p->transform && ctx->d_vertex_orig_list ? 1 : 0)); mix((unsigned long long)ctx->cam_grid.cap);
	for (int l = 0; l < p->num_lights; l++) mix((unsigned long long)ctx->light_grid[l].cap);
	mix((unsigned long long)(p->bounces > 0 && ctx->d_reflect ? p->bounces : 0)); mix((unsigned long long)ctx->refl.grid.cap);
	mix((unsigned long long)ctx->num_slabs); mix((unsigned long long)gather_slot(ctx, p)); mix((unsigned long long)ctx->trace_small); mix((unsigned long long)ctx->async_readback); mix((unsigned long long)ctx->lg_now); mix((unsigned long long)(uintptr_t)ctx->lg_local);
	mix((unsigned long long)(uintptr_t)ctx->comm); mix((unsigned long long)(uintptr_t)ctx->d_rgb_full);
	mix((unsigned long long)(uintptr_t)ctx->peer_full); mix((unsigned long long)ctx->use_push);
	mix((unsigned long long)(uintptr_t)ctx->stream); mix((unsigned long long)(uintptr_t)ctx->d_rgb);
	mix((unsigned long long)pdl_mask_for(ctx, false));
	return k;
}

static int run_frame_once(ug_context *ctx, const ug_frame_params *p)
{
	const bool stage_events = (p->flags & UG_FLAG_STAGE_EVENTS) != 0;
	const bool use_graph = !(p->flags & UG_FLAG_NO_GRAPH) && !stage_events;
	const int gs = gather_slot(ctx, p); /* a deferred gather alternates between two destination slots: one captured graph per slot */
	int fc_slot = 0;
	int rc0 = stage_consts(ctx, &fc_slot);
	if (rc0) return rc0;
	if (ctx->trace_item_forced) ctx->trace_small = ctx->trace_item_forced == 64;
	else if (ctx->frame_valid && ctx->h_counters[0]) {
		/* primary rays: work items of 64 pairs instead of 128 when the camera grid is so small that the items would not fill
		 * the GPU twice (ugk_trace); decided from the pair count the last finished frame published - a heuristic, both
		 * kernels are exact - with some hysteresis so that the frame graph is not re-captured back and forth */
		const double waves = (double)((ctx->h_counters[0] + 127u) / 128u) / ((double)ctx->num_sms * 16.0);
		if (waves < 1.8) ctx->trace_small = true;
		else if (waves > 2.2) ctx->trace_small = false;
	}
	if (use_graph) {
		const unsigned long long key = frame_key(ctx, p);
		if (!ctx->graph_exec[gs] || ctx->graph_key[gs] != key) {
			if (ctx->graph_exec[gs]) { cudaGraphExecDestroy(ctx->graph_exec[gs]); ctx->graph_exec[gs] = nullptr; }
			cudaGraph_t graph = nullptr;
			Launches L;
			UG_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
			int rc = enqueue_frame(ctx, p, true, false, L);
			cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
			if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
			if (e != cudaSuccess) return ug_fail(ctx, UG_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
			e = cudaGraphInstantiate(&ctx->graph_exec[gs], graph, cudaGraphInstantiateFlagUseNodePriority);
			cudaGraphDestroy(graph);
			if (e != cudaSuccess) return ug_fail(ctx, UG_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
			ctx->graph_key[gs] = key;
			ctx->graph_launches[gs] = L.n;
			ctx->graph_overflow[gs] = ctx->overflow_in_frame;
			ctx->graph_own[gs] = ctx->own_in_frame;
		}
		ctx->overflow_in_frame = ctx->graph_overflow[gs]; /* a property of the captured frame, not of the last capture */
		ctx->own_in_frame = ctx->graph_own[gs];
		UG_CUDA(ctx, cudaEventRecord(ctx->ev_t0, ctx->stream));
		UG_CUDA(ctx, cudaGraphLaunch(ctx->graph_exec[gs], ctx->stream));
		UG_CUDA(ctx, cudaEventRecord(ctx->ev_t1, ctx->stream));
		ctx->stats.kernel_launches = ctx->graph_launches[gs];
	} else {
		Launches L;
		UG_CUDA(ctx, cudaEventRecord(ctx->ev_t0, ctx->stream));
		int rc = enqueue_frame(ctx, p, !stage_events, stage_events, L);
		if (rc) return rc;
		UG_CUDA(ctx, cudaEventRecord(ctx->ev_t1, ctx->stream));
		ctx->stats.kernel_launches = L.n;
	}
	rc0 = consts_launched(ctx, fc_slot);
	if (rc0) return rc0;
	if (gather_deferred(ctx, p)) ctx->gather_seq++;
	ctx->frame_valid = true;
	return UG_OK;
}

/* after a synchronise: did any pair arena overflow?  If so grow it (the caller re-runs the frame). */
static int check_overflow(ug_context *ctx, int num_lights, bool *grew, bool reflect = false)
{
	*grew = false;
	GridBuffers *gs[2 + UG_MAX_LIGHTS];
	int n = 0;
	gs[n++] = &ctx->cam_grid;
	for (int l = 0; l < num_lights; l++) gs[n++] = &ctx->light_grid[l];
	if (reflect) gs[n++] = &ctx->refl.grid;
	for (int i = 0; i < n; i++) {
		const u32 *c = ctx->h_counters + 4 * ((reflect && i == n - 1) ? 1 + UG_MAX_LIGHTS : i); /* copied by the frame itself; the caller has synchronised */
		if (c[1]) {
			const size_t want = (size_t)c[2] + (size_t)c[2] / 4 + 1024;
			int rc = alloc_grid(ctx, *gs[i], gs[i]->T, gs[i]->C, want, 0, gs[i] != &ctx->refl.grid);
			if (rc) return rc;
			if (gs[i] == &ctx->refl.grid && (rc = alloc_reflect_records(ctx)) != UG_OK) return rc;
			*grew = true;
		}
	}
	if (ctx->lg_now && ctx->overflow_in_frame && ctx->h_counters[UG_HC_WORDS - 1]) {
		/* somebody's arena overflowed in a frame with sharded light grids (the flag is the same on every rank): the slots of the
		 * shared lists cannot grow, so from now on every rank builds its light grids in full again */
		ctx->lg_disabled = true;
		ctx->lg_now = false;
		ug_drop_graphs(ctx);
	}
	if (*grew) ug_drop_graphs(ctx);
	return UG_OK;
}

extern "C" int ug_render_frame(ug_context *ctx, const ug_frame_params *p)
{
	if (!ctx) return UG_ERR_ARG;
	int rc = prepare_frame(ctx, p);
	if (rc) return rc;
	ctx->stats.overflowed = 0;
	return run_frame_once(ctx, p);
}

extern "C" int ug_render_frame_host(ug_context *ctx, const ug_frame_params *p, uint8_t *host_rgb)
{
	if (!ctx || !host_rgb) return ug_fail(ctx, UG_ERR_ARG, "ug_render_frame_host: null argument");
	/* a deferred gather is whole only after the NEXT frame's handshake (or ug_comm_flush): there is nothing this call could
	 * hand back.  Checked before any collective, and the flag is the same on every rank, so all ranks return together. */
	if (p && (p->flags & UG_FLAG_GATHER_DEFERRED)) return ug_fail(ctx, UG_ERR_ARG, "ug_render_frame_host: UG_FLAG_GATHER_DEFERRED frames are read with ug_comm_flush + ug_read");
	int rc = prepare_frame(ctx, p);
	if (rc) return rc;
	ctx->stats.overflowed = 0;
	/* main.cu:245: image back to the host.  A pinned / registered destination is written by the copy engine
	 * directly; pageable memory goes through the context's pinned staging buffer. */
	cudaPointerAttributes attr;
	bool direct = cudaPointerGetAttributes(&attr, host_rgb) == cudaSuccess && attr.type == cudaMemoryTypeHost;
	cudaGetLastError();
	const bool gathered = ctx->comm && ctx->comm_world > 1 && ctx->strip_count == ctx->comm_world;
	const bool collective = gathers(ctx, p);
	const size_t out_bytes = gathered ? 3 * (size_t)ctx->W * ctx->H : 3 * (size_t)ctx->N;
	unsigned char *dst = direct ? host_rgb : ctx->h_rgb_pinned;
	if (gathered && !direct && ctx->comm_rank == 0) { /* pageable destination of a gathered frame: through a pinned staging block of the full size */
		if (ctx->h_full_bytes < out_bytes) {
			if (ctx->h_full_pinned) cudaFreeHost(ctx->h_full_pinned);
			ctx->h_full_pinned = nullptr; ctx->h_full_bytes = 0;
			UG_CUDA(ctx, cudaMallocHost((void **)&ctx->h_full_pinned, out_bytes));
			ctx->h_full_bytes = out_bytes;
		}
		dst = ctx->h_full_pinned;
	}
	for (int attempt = 0; attempt < 3; attempt++) {
		rc = run_frame_once(ctx, p);
		if (rc) return rc;
		if (gathered) { if (ctx->comm_rank == 0) UG_CUDA(ctx, cudaMemcpyAsync(dst, ctx->d_rgb_full, out_bytes, cudaMemcpyDeviceToHost, ctx->stream)); }
		else UG_CUDA(ctx, cudaMemcpyAsync(dst, ctx->d_rgb, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
		UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		bool grew = false;
		rc = check_overflow(ctx, p->num_lights, &grew, p->bounces > 0 && ctx->d_reflect);
		if (rc) return rc;
		const bool grew_here = grew;
		if (collective) {
			/* every attempt issues the frame's collectives (handshakes, gather): the decision to run it again must be the same
			 * on every rank, or the ranks fall out of step and hang.  One 4-byte max-allreduce after the frame. */
			unsigned any = grew ? 1u : 0u;
			if (ctx->overflow_in_frame) any = ctx->h_counters[UG_HC_WORDS - 1]; /* summed over the ranks by the frame's own "done" handshake */
			else if ((rc = ugc_allreduce_max(ctx, ctx->stream, &any)) != UG_OK) return rc;
			grew = any != 0;
		}
		if (!grew) {
			if (!direct && (!gathered || ctx->comm_rank == 0)) memcpy(host_rgb, dst, out_bytes);
			return UG_OK;
		}
		if (grew_here) ctx->stats.overflowed = 1; /* the other ranks run the frame again too, but their arenas were large enough */
	}
	return ug_fail(ctx, UG_ERR_OVERFLOW, "pair arena still too small after growing");
}

/* Frames in flight: the call returns once the frame and the read-back of its image are enqueued; the copy runs on its own
 * stream, so the next frame (submitted right away) renders while this one's image crosses PCIe.  `host_rgb` must be pinned
 * (cudaHostAlloc / cudaHostRegister) and stay untouched until ug_wait_frame.  An arena overflow cannot be repaired behind the
 * caller's back here: ug_wait_frame reports it (after growing the arena) and the caller submits the frame again. */
extern "C" int ug_render_frame_host_async(ug_context *ctx, const ug_frame_params *p, uint8_t *host_rgb)
{
	if (!ctx || !host_rgb) return ug_fail(ctx, UG_ERR_ARG, "ug_render_frame_host_async: null argument");
	if (p && (p->flags & UG_FLAG_GATHER_DEFERRED)) return ug_fail(ctx, UG_ERR_ARG, "ug_render_frame_host_async: UG_FLAG_GATHER_DEFERRED frames are read with ug_comm_flush + ug_read");
	cudaPointerAttributes attr;
	const bool pinned = cudaPointerGetAttributes(&attr, host_rgb) == cudaSuccess && attr.type == cudaMemoryTypeHost;
	cudaGetLastError();
	if (!pinned) return ug_fail(ctx, UG_ERR_ARG, "ug_render_frame_host_async: the destination must be pinned host memory");
	if (!ctx->async_readback) { ctx->async_readback = true; ug_drop_graphs(ctx); }
	int rc = prepare_frame(ctx, p);
	if (rc) return rc;
	ctx->stats.overflowed = 0;
	rc = run_frame_once(ctx, p);
	if (rc) return rc;
	const bool gathered = ctx->comm && ctx->comm_world > 1 && ctx->strip_count == ctx->comm_world;
	UG_CUDA(ctx, cudaEventRecord(ctx->ev_frame_done, ctx->stream));
	if (!gathered || ctx->comm_rank == 0) {
		const size_t out_bytes = gathered ? 3 * (size_t)ctx->W * ctx->H : 3 * (size_t)ctx->N;
		UG_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_frame_done, 0));
		UG_CUDA(ctx, cudaMemcpyAsync(host_rgb, gathered ? ctx->d_rgb_full : ctx->d_rgb, out_bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
		UG_CUDA(ctx, cudaEventRecord(ctx->ev_copy_done, ctx->copy_stream));
	} else {
		UG_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_frame_done, 0)); /* a rank without an image: its "read-back" is the end of its frame */
	}
	UG_CUDA(ctx, cudaEventRecord(ctx->ev_copy_ring[ctx->async_seq % 4], ctx->copy_stream));
	ctx->async_seq++;
	return UG_OK;
}

/* frames_back = 0: the image of the last ug_render_frame_host_async call is in host memory when this returns, and
 * UG_ERR_OVERFLOW says that a pair arena was too small for that frame (it has been grown: submit the frame again);
 * frames_back = 1..3: the image of the frame submitted that many calls earlier is in host memory (what a loop that keeps
 * frames in flight waits on before it consumes a buffer and hands it out again). */
extern "C" int ug_wait_frame(ug_context *ctx, int frames_back)
{
	if (!ctx || frames_back < 0 || frames_back > 3) return ug_fail(ctx, UG_ERR_ARG, "ug_wait_frame: frames_back must be 0..3");
	if ((unsigned long long)frames_back >= ctx->async_seq) return UG_OK; /* nothing that old was submitted */
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	UG_CUDA(ctx, cudaEventSynchronize(ctx->ev_copy_ring[(ctx->async_seq - 1 - (unsigned long long)frames_back) % 4]));
	if (frames_back > 0) return UG_OK;
	ug_stats st;
	int rc = ug_get_stats(ctx, &st);
	if (rc) return rc;
	if (st.overflowed) return ug_fail(ctx, UG_ERR_OVERFLOW, "a pair arena overflowed in the last frame; it has been grown, render the frame again");
	return UG_OK;
}

/* ------------------------------------------------------------------ per_frame_funcs.h, call by call */
extern "C" int ug_update_light_position(ug_context *ctx, const float *xyz, int num_lights)
{
	if (!ctx || !xyz || num_lights < 0 || num_lights > UG_MAX_LIGHTS) return ug_fail(ctx, UG_ERR_ARG, "updateLightPosition: bad arguments");
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	for (int i = 0; i < 3 * UG_MAX_LIGHTS; i++) ctx->h_fc->light_pos[i] = i < 3 * num_lights ? xyz[i] : 0.0f;
	ctx->h_fc->num_lights = num_lights;
	return upload_consts(ctx);
}

extern "C" int ug_fill_coordinates_data(ug_context *ctx, const ug_camera *cam)
{
	if (!ctx || !cam) return ug_fail(ctx, UG_ERR_ARG, "fillCoordinatesData: bad arguments");
	if (ctx->W <= 0) return ug_fail(ctx, UG_ERR_ARG, "fillCoordinatesData: ug_set_screen first");
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	ugh_view_const(&ctx->h_fc->cam, cam, ctx->W, ctx->H, ctx->nbx, ctx->nby, ctx->gy_lo, ctx->gy_cnt, UG_LIGHT_GRID_PERSPECTIVE, 0.0f, 0.0f, ctx->num_slabs);
	memcpy(ctx->h_fc->shade_mv, cam->modelview_matrix, sizeof(ctx->h_fc->shade_mv));
	memcpy(ctx->h_fc->shade_eye, cam->worldori, sizeof(ctx->h_fc->shade_eye));
	return upload_consts(ctx);
}

extern "C" int ug_build_frustum_grid(ug_context *ctx)
{
	if (!ctx) return UG_ERR_ARG;
	int rc = ensure_buffers(ctx, ctx->num_lights, ctx->full_alloc);
	if (rc) return rc;
	for (int attempt = 0; attempt < 3; attempt++) {
		Launches L;
		enqueue_grid(ctx, ctx->stream, &ctx->d_fc->cam, ctx->cam_grid, L);
		UG_CUDA(ctx, cudaGetLastError());
		UG_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters, ctx->cam_grid.counters, 4 * sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
		UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		bool grew = false;
		rc = check_overflow(ctx, 0, &grew);
		if (rc) return rc;
		if (!grew) return UG_OK;
	}
	return ug_fail(ctx, UG_ERR_OVERFLOW, "pair arena still too small after growing");
}

extern "C" int ug_trace(ug_context *ctx)
{
	if (!ctx || !ctx->cam_grid.tri_rec) return ug_fail(ctx, UG_ERR_ARG, "trace: build_frustum_grid first");
	int rc = ensure_buffers(ctx, ctx->num_lights, true);
	if (rc) return rc;
	ugk_init_pixels(ctx->stream, ctx->d_fc, ctx->d_t, ctx->d_id, ctx->d_shadowed, ctx->d_best, ctx->d_normal, ctx->d_dir, ctx->N, ctx->num_sms);
	if (ctx->cam_grid.S > 1) ugk_trace_slabs(ctx->stream, ctx->d_fc, ctx->cam_grid, ctx->d_dir, ctx->d_t, ctx->d_id, ctx->d_normal, ctx->C_cam, ctx->num_sms);
	else ugk_trace(ctx->stream, ctx->d_fc, ctx->cam_grid, ctx->d_best, ctx->d_cell_done, ctx->d_dir, ctx->d_t, ctx->d_id, ctx->d_normal, ctx->num_sms, ctx->trace_small);
	UG_CUDA(ctx, cudaGetLastError());
	return UG_OK;
}

extern "C" int ug_shadow_pass(ug_context *ctx, int light_id, const ug_camera *light_cam)
{
	if (!ctx || !light_cam || light_id < 0 || light_id >= UG_MAX_LIGHTS) return ug_fail(ctx, UG_ERR_ARG, "shadow pass: bad arguments");
	if (!ctx->d_t) return ug_fail(ctx, UG_ERR_ARG, "shadow pass: trace first");
	ctx->lg_now = false; /* the call-by-call path builds every light grid in full */
	if (ctx->num_lights <= light_id) ctx->num_lights = light_id + 1;
	int rc = ensure_buffers(ctx, ctx->num_lights, ctx->full_alloc);
	if (rc) return rc;
	ugh_view_const(&ctx->h_fc->light[light_id], light_cam, ctx->W, ctx->H, ctx->nbx, ctx->nby, 0, ctx->nby, ctx->light_grid_mode, ctx->light_xM, ctx->light_yM, ctx->num_slabs);
	memcpy(ctx->h_fc->shade_mv, light_cam->modelview_matrix, sizeof(ctx->h_fc->shade_mv)); /* fillCoordinatesData with the light camera */
	memcpy(ctx->h_fc->shade_eye, light_cam->worldori, sizeof(ctx->h_fc->shade_eye));
	rc = upload_consts(ctx);
	if (rc) return rc;
	for (int attempt = 0; attempt < 3; attempt++) {
		Launches L;
		enqueue_light_zero(ctx, ctx->stream, light_id);
		enqueue_light_grid(ctx, ctx->stream, light_id, L);
		UG_CUDA(ctx, cudaGetLastError());
		UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		bool grew = false;
		GridBuffers &g = ctx->light_grid[light_id];
		u32 c[4];
		UG_CUDA(ctx, cudaMemcpy(c, g.counters, sizeof(c), cudaMemcpyDeviceToHost));
		if (c[1]) {
			rc = alloc_grid(ctx, g, g.T, g.C, (size_t)c[2] + (size_t)c[2] / 4 + 1024);
			if (rc) return rc;
			grew = true;
		}
		if (!grew) {
			enqueue_light_rays(ctx, ctx->stream, light_id, L);
			UG_CUDA(ctx, cudaGetLastError());
			return UG_OK;
		}
	}
	return ug_fail(ctx, UG_ERR_OVERFLOW, "pair arena still too small after growing");
}

/* UG_FLAG_GATHER_DEFERRED: make the last gathered frame whole on rank 0 (one handshake + a synchronise); collective */
extern "C" int ug_comm_flush(ug_context *ctx)
{
	if (!ctx) return UG_ERR_ARG;
	if (!ctx->comm || ctx->comm_world <= 1) return UG_OK;
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	int rc = ugc_flush(ctx, ctx->stream);
	if (rc) return rc;
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return UG_OK;
}

extern "C" int ug_set_light_grid_mode(ug_context *ctx, int mode, float xM, float yM)
{
	if (!ctx) return UG_ERR_ARG;
	if (mode != UG_LIGHT_GRID_PERSPECTIVE && mode != UG_LIGHT_GRID_SPHERICAL) return ug_fail(ctx, UG_ERR_ARG, "light grid mode %d unknown", mode);
	ctx->light_grid_mode = mode;
	ctx->light_xM = xM; ctx->light_yM = yM;
	return UG_OK;
}

extern "C" int ug_simple_shade(ug_context *ctx)
{
	if (!ctx || !ctx->d_t) return ug_fail(ctx, UG_ERR_ARG, "simpleShade: trace first");
	ugk_shade(ctx->stream, ctx->d_fc, ctx->d_t, ctx->d_dir, ctx->d_id, ctx->d_shadowed, ctx->cam_grid.tri_rec, ctx->d_materiallist_index, ctx->d_materiallist,
	          ctx->num_materials, ctx->d_rgb, ctx->d_matid, ctx->N, 0, ctx->num_sms);
	UG_CUDA(ctx, cudaGetLastError());
	ctx->frame_valid = true;
	return UG_OK;
}

/* shader.h:88-113 spotlight_shade -> spot_shade (shader_kernel.cu:275-345): what display() switches to from its third frame
 * on (main.cu:205-219).  The reference's host wrapper also copies the two cone angles of every pixel back and printf's
 * them (shader.h:106-111); that debug dump is not reproduced. */
extern "C" int ug_spotlight_shade(ug_context *ctx)
{
	if (!ctx || !ctx->d_t) return ug_fail(ctx, UG_ERR_ARG, "spotlight_shade: trace first");
	ugk_shade(ctx->stream, ctx->d_fc, ctx->d_t, ctx->d_dir, ctx->d_id, ctx->d_shadowed, ctx->cam_grid.tri_rec, ctx->d_materiallist_index, ctx->d_materiallist,
	          ctx->num_materials, ctx->d_rgb, ctx->d_matid, ctx->N, 0, ctx->num_sms, true);
	UG_CUDA(ctx, cudaGetLastError());
	ctx->frame_valid = true;
	return UG_OK;
}

extern "C" int ug_add_shadows(ug_context *ctx)
{
	if (!ctx || !ctx->d_rgb) return ug_fail(ctx, UG_ERR_ARG, "add_shadows: shade first");
	ugk_add_shadows(ctx->stream, ctx->d_shadowed, ctx->d_rgb, ctx->N, ctx->num_sms);
	UG_CUDA(ctx, cudaGetLastError());
	return UG_OK;
}

/* ------------------------------------------------------------------ read-back */
static int buffer_info(ug_context *ctx, int buffer, int index, void **ptr, size_t *bytes)
{
	const size_t N = (size_t)ctx->N;
	u32 c[4] = {0, 0, 0, 0};
	GridBuffers *g = nullptr;
	if (buffer == UG_BUF_CAM_VALUES || buffer == UG_BUF_CAM_SPAN || buffer == UG_BUF_CAM_OFFSET || buffer == UG_BUF_SIZELIST) g = &ctx->cam_grid;
	if (buffer == UG_BUF_LIGHT_VALUES || buffer == UG_BUF_LIGHT_SPAN || buffer == UG_BUF_LIGHT_OFFSET) {
		if (index < 0 || index >= UG_MAX_LIGHTS) return ug_fail(ctx, UG_ERR_ARG, "light index %d", index);
		g = &ctx->light_grid[index];
	}
	if (g && !g->tri_rec) return ug_fail(ctx, UG_ERR_ARG, "buffer %d not built yet", buffer);
	switch (buffer) {
	case UG_BUF_VERTEXLIST: *ptr = ctx->d_vertexlist; *bytes = sizeof(float) * 3 * (size_t)ctx->num_vertices; break;
	case UG_BUF_CAM_VALUES:
	case UG_BUF_LIGHT_VALUES:
		UG_CUDA(ctx, cudaMemcpy(c, g->counters, sizeof(c), cudaMemcpyDeviceToHost));
		*ptr = g->vals[g->final_buf]; *bytes = sizeof(u32) * (size_t)c[0];
		break;
	case UG_BUF_CAM_SPAN:
	case UG_BUF_LIGHT_SPAN: *ptr = g->span; *bytes = sizeof(u32) * (size_t)g->C; break;
	case UG_BUF_CAM_OFFSET:
	case UG_BUF_LIGHT_OFFSET: *ptr = g->offset; *bytes = sizeof(u32) * (size_t)g->C; break;
	case UG_BUF_SIZELIST: *ptr = g->count; *bytes = sizeof(u32) * (size_t)g->T; break;
	case UG_BUF_RGB_FULL: *ptr = ctx->d_rgb_full; *bytes = ctx->rgb_full_bytes; break;
	case UG_BUF_CAM_COUNTERS: *ptr = ctx->cam_grid.counters; *bytes = sizeof(u32) * 32; break;
	case UG_BUF_T: *ptr = ctx->d_t; *bytes = sizeof(float) * N; break;
	case UG_BUF_ID: *ptr = ctx->d_id; *bytes = sizeof(int) * N; break;
	case UG_BUF_NORMAL: *ptr = ctx->d_normal; *bytes = sizeof(float) * 3 * N; break;
	case UG_BUF_DIR: *ptr = ctx->d_dir; *bytes = sizeof(float) * 3 * N; break;
	case UG_BUF_SHADOWED: *ptr = ctx->d_shadowed; *bytes = sizeof(int) * N; break;
	case UG_BUF_RGB: *ptr = ctx->d_rgb; *bytes = 3 * N; break;
	case UG_BUF_LIGHT_CELL:
		if (index < 0 || index >= UG_MAX_LIGHTS) return ug_fail(ctx, UG_ERR_ARG, "light index %d", index);
		*ptr = ctx->rays[index].cell; *bytes = sizeof(u32) * N;
		break;
	default: return ug_fail(ctx, UG_ERR_ARG, "unknown buffer %d", buffer);
	}
	if (!*ptr) return ug_fail(ctx, UG_ERR_ARG, "buffer %d not allocated (flags?)", buffer);
	return UG_OK;
}

extern "C" int ug_device_ptr(ug_context *ctx, int buffer, int index, void **ptr, size_t *bytes)
{
	if (!ctx || !ptr || !bytes) return UG_ERR_ARG;
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return buffer_info(ctx, buffer, index, ptr, bytes);
}

extern "C" int ug_read(ug_context *ctx, int buffer, int index, void *dst, size_t bytes)
{
	if (!ctx || !dst) return UG_ERR_ARG;
	void *ptr = nullptr;
	size_t have = 0;
	int rc = ug_device_ptr(ctx, buffer, index, &ptr, &have);
	if (rc) return rc;
	if (bytes > have) return ug_fail(ctx, UG_ERR_ARG, "buffer %d holds %zu bytes, %zu requested", buffer, have, bytes);
	UG_CUDA(ctx, cudaMemcpy(dst, ptr, bytes, cudaMemcpyDeviceToHost));
	return UG_OK;
}

extern "C" int ug_get_stats(ug_context *ctx, ug_stats *out)
{
	if (!ctx || !out) return UG_ERR_ARG;
	UG_CUDA(ctx, cudaSetDevice(ctx->device));
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	ug_stats &s = ctx->stats;
	s.num_faces = ctx->num_faces; s.num_vertices = ctx->num_vertices; s.num_materials = ctx->num_materials;
	s.width = ctx->W; s.height = ctx->gy_cnt * UG_TILE; s.num_cells = ctx->C_cam; s.num_pixels = ctx->N;
	s.pair_capacity = (unsigned)ctx->cam_grid.cap;
	u32 c[4] = {0, 0, 0, 0};
	if (ctx->cam_grid.counters) {
		UG_CUDA(ctx, cudaMemcpy(c, ctx->cam_grid.counters, sizeof(c), cudaMemcpyDeviceToHost));
		s.cam_pairs = c[0];
		if (c[1]) s.overflowed = 1;
	}
	for (int l = 0; l < UG_MAX_LIGHTS; l++) {
		s.light_pairs[l] = 0;
		if (l < ctx->num_lights && ctx->light_grid[l].counters) {
			UG_CUDA(ctx, cudaMemcpy(c, ctx->light_grid[l].counters, sizeof(c), cudaMemcpyDeviceToHost));
			s.light_pairs[l] = c[0];
			if (c[1]) s.overflowed = 1;
		}
	}
	s.reflect_rays = 0;
	s.reflect_pairs = 0;
	for (int b = 0; b < ctx->last_bounces && b < UG_MAX_BOUNCES; b++) s.reflect_rays += ctx->h_counters[4 * (2 + UG_MAX_LIGHTS) + b];
	if (ctx->last_bounces) s.reflect_pairs = ctx->h_counters[4 * (1 + UG_MAX_LIGHTS)];
	if (ctx->frame_valid) {
		float ms = 0;
		if (cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1) == cudaSuccess) s.ms_frame = ms;
		else cudaGetLastError();
		s.ms_own = ctx->own_in_frame ? (float)ctx->h_counters[UG_HC_OWN] * 1e-6f : s.ms_frame;
	}
	*out = s;
	if (s.overflowed && ctx->frame_valid) {
		/* ug_render_frame is asynchronous and cannot re-run itself: the frame that overflowed is incomplete, but the arenas are
		 * grown here so that rendering it again succeeds (ug_render_frame_host does both on its own) */
		bool grew = false;
		int rc = check_overflow(ctx, ctx->num_lights, &grew, ctx->last_bounces > 0);
		if (rc) return rc;
		if (grew) ctx->frame_valid = false;
	}
	return UG_OK;
}

/* stage timings of the last UG_FLAG_STAGE_EVENTS frame: ms between consecutive marks */
extern "C" int ug_stage_times(ug_context *ctx, float *out, int max)
{
	if (!ctx || !out) return UG_ERR_ARG;
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	int n = 0;
	for (int k = 0; k + 1 < ctx->stage_marks && n < max; k++) {
		float ms = 0;
		if (cudaEventElapsedTime(&ms, ctx->ev_stage[k], ctx->ev_stage[k + 1]) != cudaSuccess) { cudaGetLastError(); break; }
		out[n++] = ms;
	}
	return n;
}

/* the same image as a binary P6, top row first (what image viewers expect) - ~12x smaller than the P3 text
 * and written with one fwrite per row */
/* UG_TIMELINE=1: ns since the frame's first marker for each of the 64 marker slots (0 = not reached) */
extern "C" int ug_timeline(ug_context *ctx, double *ns64)
{
	if (!ctx || !ns64 || !ctx->d_stamps) return ug_fail(ctx, UG_ERR_ARG, "ug_timeline: set UG_TIMELINE=1 before ug_create");
	unsigned long long h[64];
	UG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	UG_CUDA(ctx, cudaMemcpy(h, ctx->d_stamps, sizeof(h), cudaMemcpyDeviceToHost));
	for (int i = 0; i < 64; i++) ns64[i] = h[i] ? (double)((long long)(h[i] - h[0])) : 0.0;
	return UG_OK;
}

/* the frame as a PNG, top row first - what display() gets from writePPM + `convert -flip` (main.cu:255-259)
 * without the external tool.  8-bit RGB, filter 0, zlib stream of stored (uncompressed) deflate blocks: no dependency. */
static unsigned png_crc(const unsigned char *p, size_t n, unsigned crc)
{
	static unsigned table[256];
	if (!table[1])
		for (unsigned i = 0; i < 256; i++) {
			unsigned c = i;
			for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
			table[i] = c;
		}
	for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
	return crc;
}
static void png_be32(unsigned char *p, unsigned v) { p[0] = (unsigned char)(v >> 24); p[1] = (unsigned char)(v >> 16); p[2] = (unsigned char)(v >> 8); p[3] = (unsigned char)v; }
static bool png_chunk(FILE *fp, const char *type, const unsigned char *data, size_t n)
{
	unsigned char head[8], tail[4];
	png_be32(head, (unsigned)n);
	memcpy(head + 4, type, 4);
	unsigned crc = png_crc(head + 4, 4, 0xffffffffu);
	if (n) crc = png_crc(data, n, crc);
	png_be32(tail, crc ^ 0xffffffffu);
	return fwrite(head, 1, 8, fp) == 8 && (!n || fwrite(data, 1, n, fp) == n) && fwrite(tail, 1, 4, fp) == 4;
}
extern "C" int ug_write_png(ug_context *ctx, const char *filename)
{
	if (!ctx || !filename || !ctx->d_rgb) return ug_fail(ctx, UG_ERR_ARG, "write_png: nothing rendered");
	const size_t bytes = 3 * (size_t)ctx->N, row = 3 * (size_t)ctx->W;
	const int rows = ctx->gy_cnt * UG_TILE;
	std::vector<unsigned char> img(bytes);
	int rc = ug_read(ctx, UG_BUF_RGB, 0, img.data(), bytes);
	if (rc) return rc;
	/* scanlines, top row first (pixel row 0 is the bottom one, per_app_funcs.h:48-60), each behind its filter byte */
	std::vector<unsigned char> raw((row + 1) * (size_t)rows);
	for (int y = 0; y < rows; y++) {
		unsigned char *dst = raw.data() + (size_t)y * (row + 1);
		dst[0] = 0;
		memcpy(dst + 1, img.data() + (size_t)(rows - 1 - y) * row, row);
	}
	/* zlib: header, stored blocks of at most 65535 bytes, Adler-32 of the raw data */
	std::vector<unsigned char> z;
	z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
	z.push_back(0x78); z.push_back(0x01);
	unsigned a = 1, b = 0;
	for (size_t pos = 0; pos < raw.size() || pos == 0;) {
		const size_t n = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
		z.push_back(pos + n >= raw.size() ? 1 : 0);
		z.push_back((unsigned char)(n & 0xff)); z.push_back((unsigned char)(n >> 8));
		z.push_back((unsigned char)(~n & 0xff)); z.push_back((unsigned char)((~n >> 8) & 0xff));
		z.insert(z.end(), raw.begin() + (long)pos, raw.begin() + (long)(pos + n));
		for (size_t i = pos; i < pos + n; i++) { a += raw[i]; if (a >= 65521u) a -= 65521u; b += a; if (b >= 65521u) b -= 65521u; }
		pos += n;
		if (!n) break;
	}
	unsigned char ad[4];
	png_be32(ad, (b << 16) | a);
	z.insert(z.end(), ad, ad + 4);
	FILE *fp = fopen(filename, "wb");
	if (!fp) return ug_fail(ctx, UG_ERR_IO, "cannot open %s", filename);
	static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
	unsigned char ihdr[13];
	png_be32(ihdr, (unsigned)ctx->W); png_be32(ihdr + 4, (unsigned)rows);
	ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0; /* 8 bits, truecolour, deflate, adaptive filters, no interlace */
	bool ok = fwrite(sig, 1, 8, fp) == 8 && png_chunk(fp, "IHDR", ihdr, 13) && png_chunk(fp, "IDAT", z.data(), z.size()) && png_chunk(fp, "IEND", nullptr, 0);
	ok = (fclose(fp) == 0) && ok;
	return ok ? UG_OK : ug_fail(ctx, UG_ERR_IO, "short write to %s", filename);
}

extern "C" int ug_write_ppm_binary(ug_context *ctx, const char *filename)
{
	if (!ctx || !filename || !ctx->d_rgb) return ug_fail(ctx, UG_ERR_ARG, "write_ppm_binary: nothing rendered");
	const size_t bytes = 3 * (size_t)ctx->N, row = 3 * (size_t)ctx->W;
	const int rows = ctx->gy_cnt * UG_TILE;
	std::vector<unsigned char> img(bytes);
	int rc = ug_read(ctx, UG_BUF_RGB, 0, img.data(), bytes);
	if (rc) return rc;
	FILE *fp = fopen(filename, "wb");
	if (!fp) return ug_fail(ctx, UG_ERR_IO, "cannot open %s", filename);
	fprintf(fp, "P6\n%d %d\n255\n", ctx->W, rows);
	bool ok = true;
	for (int y = rows - 1; y >= 0 && ok; y--) ok = fwrite(img.data() + row * (size_t)y, 1, row, fp) == row;
	ok = (fclose(fp) == 0) && ok;
	return ok ? UG_OK : ug_fail(ctx, UG_ERR_IO, "short write to %s", filename);
}

extern "C" int ug_write_ppm(ug_context *ctx, const char *filename)
{
	if (!ctx || !filename || !ctx->d_rgb) return ug_fail(ctx, UG_ERR_ARG, "writePPM: nothing rendered");
	const size_t bytes = 3 * (size_t)ctx->N;
	std::vector<unsigned char> img(bytes);
	int rc = ug_read(ctx, UG_BUF_RGB, 0, img.data(), bytes);
	if (rc) return rc;
	FILE *fp = fopen(filename, "w");
	if (!fp) return ug_fail(ctx, UG_ERR_IO, "cannot open %s", filename);
	/* per_app_funcs.h:39-66: ASCII P3, a newline before every image row, buffer order (bottom row first) */
	fprintf(fp, "P3\n%d %d\n%d\n", ctx->W, ctx->gy_cnt * UG_TILE, 255);
	for (size_t i = 0; i < bytes; i++) {
		if (i % (3 * (size_t)ctx->W) == 0) fprintf(fp, "\n");
		fprintf(fp, "%d ", (int)img[i]);
	}
	fprintf(fp, "\n");
	fclose(fp);
	return UG_OK;
}
