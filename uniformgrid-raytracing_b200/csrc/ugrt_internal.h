/* ugrt_internal.h -- private declarations shared by the .cu/.cpp files of libugrt.so */
#ifndef UGRT_INTERNAL_H
#define UGRT_INTERNAL_H

#include "ugrt.h"
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

typedef uint32_t u32;
typedef uint64_t u64;

/* ---------------------------------------------------------------- device-visible constants
 * One ViewConst per "camera" (the eye, and every light treated as a camera - main.cu:148-170).
 * Lives in global memory and is refreshed by ONE pinned H2D copy per frame, so the captured CUDA
 * graph never needs its kernel parameters patched. */
struct ViewConst {
	float cc[64];       /* dd_camcoords layout, per_frame_funcs.h:18-39 */
	float lattice[100]; /* texdir, 5x5 float4 */
	float tan_half;     /* tanf(FOVY*M_PI/360.0f), evaluated on the host (glibc) - misc_kernel.cu:115 */
	float near_plane;
	int nbx, nby;       /* cells of the full grid */
	int gy_lo, gy_cnt;  /* owned cell rows (image strip); lights own everything */
	int W, H;           /* full screen */
	int grid_mode;      /* UG_LIGHT_GRID_*: how triangles and shadow rays are binned into this view's grid (the eye: always perspective) */
	float xM, yM;       /* spherical mode: angular extents (main.cu:186-187 passes pi, pi) */
	int num_slabs;      /* NUM_SLABS, main.cu.h:18 (1 in the committed build): depth slabs of the grid */
	int shards;         /* a light grid built in row shards by the ranks of a strip gather (k_lightshare.cu): how many; else 1 */
	int pad_[3];
};

struct FrameConst {
	ViewConst cam;
	ViewConst light[UG_MAX_LIGHTS];
	float light_pos[3 * UG_MAX_LIGHTS]; /* dd_light_position, main.cu.h:63 */
	float shade_mv[16];                 /* the modelview resident in dd_camcoords when shading runs (last light's) */
	float shade_eye[4];                 /* and that camera's eye, dd_camcoords[0..3] (spot_shade, shader_kernel.cu:289-291) */
	float rot_cos, rot_sin;             /* cosf/sinf(rot_factor) from the host - transformation_kernel.cu:14-15 */
	int num_lights;
	int pad_;
};

/* ---------------------------------------------------------------- device buffers */
struct SortScratch {
	u32 *tile_hist = nullptr; /* [256 * G] digit-major block histograms of one radix pass */
	u32 *tile_base = nullptr; /* their exclusive scan */
	u64 *scan_state = nullptr; /* decoupled look-back descriptors of ugk_exclusive_scan */
	size_t scan_tiles = 0;
};

struct GridBuffers {       /* FrustumGrid, frustum_grid.h:15-68 */
	float4 *tri_rec = nullptr; /* [4*T]: (tvec.xyz,e2q) (e1.xyz,0) (e2.xyz,0) (qvec.xyz,0) for this view's origin */
	uint2 *rect = nullptr;     /* [T]: gx0 | gx1<<16, gy0 | gy1<<16 (owned-row clipped) */
	u32 *count = nullptr;      /* [T] d_sizeList */
	u32 *scan = nullptr;       /* [T+1] exclusive; [T] = P */
	u32 *keys[2] = {nullptr, nullptr};
	u32 *vals[2] = {nullptr, nullptr}; /* vals[final_buf] = d_triangle_value_list */
	u32 *span = nullptr;       /* [C] */
	u32 *offset = nullptr;     /* [C+1] */
	u32 *counters = nullptr;   /* [8]: 0 = P (clamped), 1 = overflow flag, 2 = P wanted */
	SortScratch sort;
	u32 *fused_hist = nullptr, *fused_tot = nullptr; /* scratch of the fused build (k_gridfused.cu) */
	u64 *fused_state_a = nullptr, *fused_state_b = nullptr;
	/* scratch of the three-launch build (k_gridsweep.cu); null when the grid is not eligible (more than UG_SWEEP_MAXB cell rows or columns) */
	u32 *sweep_hist = nullptr, *sweep_base = nullptr, *sweep_ctl = nullptr, *sweep_tile_first = nullptr;
	u64 *sweep_desc_a = nullptr, *sweep_desc_b = nullptr, *sweep_desc = nullptr;
	size_t cap = 0;            /* pair capacity */
	int final_buf = 0;         /* which of vals[] holds the sorted list */
	int key_bits = 0;
	u32 offset_bias = 0;       /* added to every cell offset the sweep build writes (a shard's slot base, k_lightshare.cu) */
	int C = 0, T = 0;          /* C = cells * S: span/offset have one entry per (cell, slab), key = cell*S + slab (grid_kernel.cu:322) */
	/* depth slabs (NUM_SLABS > 1) */
	int S = 1, cells = 0;
	float *projz = nullptr;    /* [T] d_projCoordZ (frustum_grid.h:33): min NDC z (perspective) / min distance (spherical) */
	u32 *zrange = nullptr;     /* [2] zMin bits, order-preserving zMax: the host loop of frustum_grid.h:227-235 as atomics */
	u32 *cspan = nullptr, *coffset = nullptr; /* [cells] a cell's slabs as ONE range of the sorted list (what an any-hit over all slabs walks) */
	const u32 *cell_span() const { return S > 1 ? cspan : span; }
	const u32 *cell_offset() const { return S > 1 ? coffset : offset; }
};

struct LightRays {         /* DecisionData's regroup state (decision_data.h:72-80), one per light */
	u32 *cell = nullptr;   /* [N]   d_map[N..2N): light cell per pixel, C = outside the light frustum */
	u32 *pix = nullptr;    /* [N]   d_map[0..N) after the regroup: pixel ids grouped by light cell */
	u32 *span = nullptr;   /* [C+1] rays per light cell (last = outside) */
	u32 *offset = nullptr; /* [C+2] */
	u32 *rank = nullptr;   /* [N]   position of every ray among its cell's rays (handed out by the histogram's atomics) */
	u32 *item_off = nullptr; /* [C+1] first work item of every cell */
	uint2 *items = nullptr;  /* work items of the shadow pass: (cell, local index) */
	u32 *counters = nullptr; /* [0] number of items, [1] ticket */
	u64 *state_a = nullptr, *state_b = nullptr; /* look-back descriptors (self-resetting) */
	size_t item_cap = 0;
	uint4 *heavy = nullptr;  /* k_shadowdirect.cu: queue of (cell, pixel block, lane mask, list segment) */
	size_t heavy_cap = 0;
};

struct UGridConst;
#define UG_MAX_BOUNCES 31
#define UG_FC_SLOTS 4
#define UG_SWEEP_MAXB 1024 /* k_gridsweep.cu sorts on (cell column, cell row): at most this many of each */
struct ReflectState {      /* row R: world-space uniform grid + per-pixel reflection ray state (k_reflect.cu) */
	GridBuffers grid;      /* count/scan/keys/vals/span/offset of the uniform grid; grid.C = cell arena */
	u32 *box = nullptr;    /* [6] scene bounding box (order-preserving uint encoding) */
	UGridConst *uc = nullptr;
	float4 *geo = nullptr; /* [3T] v0, e1, e2 */
	float4 *tbox = nullptr; /* [2T] grown bounding boxes */
	uint4 *prec = nullptr;  /* [grid.cap] per-pair records (k_ugrid_records) */
	uint4 *cellrec = nullptr; /* [grid.C] (offset, span, union box) */
	size_t prec_cap = 0;
	uint4 *rect = nullptr; /* [T] cell box of every triangle + pair count */
	float *O = nullptr, *r = nullptr, *w = nullptr, *k0 = nullptr; /* per-pixel ray, running weight, first-hit reflectivity */
	float *term = nullptr; /* [term_bounces][3N] what bounce b adds to the pixel; summed in order by the combine pass */
	int term_bounces = 0;
	u32 *list[2] = {nullptr, nullptr}, *list0 = nullptr, *depth = nullptr; /* compacted rays: list0 = bounce 1 (kept for the combine), then ping-pong */
	u32 *ray_counts = nullptr; /* [UG_MAX_BOUNCES + 1] live reflection rays per bounce */
	int N = 0;
};

struct ug_context {
	int device = 0;
	cudaStream_t own_stream = nullptr, stream = nullptr;
	cudaStream_t side[UG_MAX_LIGHTS] = {};
	cudaStream_t side_init = nullptr, side_refl = nullptr, side_ugrid = nullptr, side_comm = nullptr;
	cudaEvent_t ev_ready = nullptr, ev_cam = nullptr;
	cudaEvent_t ev_ugrid = nullptr, ev_camgrid = nullptr;
	cudaStream_t side_map[UG_MAX_LIGHTS] = {}; /* ray -> light cell map of every light: needs the primary hits, not the light grid */
	cudaEvent_t ev_map[UG_MAX_LIGHTS] = {};
	unsigned long long *d_stamps = nullptr; /* [64] UG_TIMELINE markers, null when off */
	bool use_fused_ugrid = false, force_fused_grid = false, fused_grid_now = true, fused_cam_now = true;
	float refl_cells_per_tri = 2.0f; /* row R uniform-grid resolution (env UG_REFL_CELLS_PER_TRI overrides) */
	cudaEvent_t ev_fork2 = nullptr, ev_init = nullptr, ev_refl = nullptr;
	cudaEvent_t ev_preshade = nullptr, ev_pub = nullptr; /* the counters' publish runs beside the shading (frames that are not gathered) */
	cudaEvent_t ev_fork = nullptr, ev_trace = nullptr, ev_join[UG_MAX_LIGHTS] = {}, ev_t0 = nullptr, ev_t1 = nullptr;
	cudaEvent_t ev_stage[40] = {};
	int stage_marks = 0;
	std::string err;
	int num_sms = 148;
	bool use_fused_regroup = true;
	bool use_fused = true;      /* fused cooperative grid build when the device supports it (UG_NO_FUSED=1 disables) */
	bool own_in_frame = false;  /* the last enqueued frame published the ns of this rank's own work (h_counters[UG_HC_OWN]) */
	bool overflow_in_frame = false; /* the last enqueued frame's "done" handshake carried the overflow flags (h_counters[UG_HC_WORDS-1]) */
	bool light_after_cam = false; /* UG_LIGHT_AFTER_CAM=1: strip mode holds a light grid's sort passes back until the camera grid is done */
	int light_build_sms = 0;    /* UG_LIGHT_BLOCKS: SMs' worth of blocks for a light-grid build in strip mode (0: all) */
	bool direct_shadow = false; /* UG_SHADOW_DIRECT=1: shadow rays in screen order, no regroup (k_shadowdirect.cu); default is the regrouped path (k_shadow.cu) */
	int dir_recompute = 0;      /* UG_DIR_RECOMPUTE: bit 0 = k_shade, bit 1 = k_regroup phase A evaluate the primary ray direction again instead of reading d_primary_ray_direction (frames only).  Measured slower on the B200 (config D: k_shade 174 -> 247 us, regroup 288 -> 342 us; the lattice expression with its two double operations costs more than 12 bytes per pixel), so the stored buffer stays the default */
	bool use_sweep = true;      /* three-launch one-sweep grid build (k_gridsweep.cu) where eligible (UG_NO_SWEEP=1 disables) */

	/* Model (scene.h:13-57) */
	std::vector<float> h_vertexlist, h_materiallist;
	std::vector<int> h_facelist, h_materiallist_index;
	float *d_vertexlist = nullptr, *d_materiallist = nullptr, *d_vertex_orig_list = nullptr;
	float *d_reflect = nullptr;  /* [M] reflectivity per material (MTL `r`), row R */
	float4 *d_tri_nv = nullptr;  /* [T] view-space shading normal of every triangle for the frame being rendered (k_shade_normals) */
	bool hoist_normals = true;   /* UG_NO_HOIST_NORMALS=1: k_shade evaluates the normal per pixel again */
	std::vector<float> h_reflect;
	int *d_facelist = nullptr, *d_materiallist_index = nullptr;
	int num_faces = 0, num_vertices = 0, num_materials = 0;
	int size_orig_list = 0, offset_orig_list = 0;

	/* screen (main.cu.h:10-30 made runtime) */
	int W = 0, H = 0, nbx = 0, nby = 0;
	float fovy = 45.0f;
	int strip_index = 0, strip_count = 1, gy_lo = 0, gy_cnt = 0;
	std::vector<int> strip_bounds; /* [strip_count + 1] cell-row boundaries when the strips are not an equal split (ug_set_strip_bounds) */
	int N = 0, C_cam = 0, C_light = 0; /* pixels / cells of this strip; cells of a (full) light grid */

	/* frame constants */
	/* h_fc is the host master copy every entry point edits.  Each upload copies it into the next slot of a small pinned
	 * ring; the frame's first kernel (k_fetch_consts) picks its slot from a device-side sequence counter that the host
	 * mirrors, so a captured graph never needs patching and ug_render_frame never waits for the previous frame. */
	FrameConst *h_fc = nullptr;
	FrameConst *h_ring = nullptr;              /* pinned [UG_FC_SLOTS] */
	cudaEvent_t ev_ring[UG_FC_SLOTS] = {};     /* recorded after the launch that reads the slot */
	bool ring_used[UG_FC_SLOTS] = {};
	unsigned long long *d_tstart = nullptr;    /* %globaltimer of the last frame's first kernel */
	u32 *d_fc_seq = nullptr;                   /* device mirror of fc_seq (advanced by k_fetch_consts) */
	unsigned long long fc_seq = 0;
	FrameConst *d_fc = nullptr;
	float *h_vstage[2] = {nullptr, nullptr}; /* pinned staging of the keyframe loader (ug_model_update_vertices), alternating */
	cudaEvent_t ev_vstage[2] = {nullptr, nullptr};
	size_t vstage_floats = 0;
	int vstage_next = 0;
	float bounds[6] = {0, 0, 0, 0, 0, 0}; /* xMin,yMin,zMin,xMax,yMax,zMax of the last keyframe (scene.h:92-111) */
	u32 *h_counters = nullptr;  /* pinned [1+UG_MAX_LIGHTS][4]: counters of every grid, copied back by the frame */
	int num_lights = 0;
	int last_bounces = 0;
	int num_slabs = 1;           /* NUM_SLABS of every grid (ug_set_slabs / ug_frame_params.num_slabs) */
	int light_grid_mode = UG_LIGHT_GRID_PERSPECTIVE; /* ug_set_light_grid_mode: used by ug_shadow_pass */
	float light_xM = 3.14159274f, light_yM = 3.14159274f;

	/* grids */
	GridBuffers cam_grid;
	GridBuffers light_grid[UG_MAX_LIGHTS];
	LightRays rays[UG_MAX_LIGHTS];
	ReflectState refl;
	int lights_allocated = 0;

	/* DecisionData (decision_data.h:66-82) */
	float *d_t = nullptr;
	int *d_id = nullptr, *d_shadowed = nullptr, *d_matid = nullptr;
	float *d_dir = nullptr;     /* [3N] d_primary_ray_direction, always kept */
	float *d_normal = nullptr;  /* [3N] FULL_BUFFERS only */
	u64 *d_best = nullptr;      /* [N] per-pixel (t bits, pair position) merge word of split cells */
	u32 *d_cell_done = nullptr; /* [C_cam] arrival counters of split cells (self-resetting) */
	unsigned char *d_rgb = nullptr;      /* where frames are written (own buffer or the caller's) */
	unsigned char *d_rgb_own = nullptr, *d_rgb_ext = nullptr;
	size_t rgb_ext_bytes = 0;
	void *gl_resource = nullptr; /* present_gl.cpp: the registered PBO (cudaGraphicsResource_t); mapped = d_rgb_ext is its storage */
	bool gl_mapped = false;
	/* ug_render_frame_host_async: the image read-back runs on its own stream while the next frame renders */
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t ev_frame_done = nullptr, ev_copy_done = nullptr;
	cudaEvent_t ev_copy_ring[4] = {};   /* read-back of async frame n = ring[n % 4]: what ug_wait_frame(ctx, frames_back) waits on */
	unsigned long long async_seq = 0;   /* async frames submitted so far */
	bool async_readback = false; /* frames wait (inside the graph, in front of the first kernel that rewrites the image) for the last read-back */
	unsigned char *h_rgb_pinned = nullptr;
	unsigned char *h_full_pinned = nullptr; /* rank 0: staging of a gathered frame whose destination is pageable */
	size_t h_full_bytes = 0;
	size_t pix_alloc = 0;
	bool full_alloc = false;

	/* multi-GPU strips (comm.cpp) */
	void *comm = nullptr;            /* ncclComm_t */
	int comm_rank = 0, comm_world = 1;
	unsigned char *d_rgb_full = nullptr; /* rank 0: the gathered frame, 3*W*H */
	size_t rgb_full_bytes = 0;
	unsigned char *peer_full = nullptr;  /* ranks > 0: rank 0's gather buffer mapped through CUDA IPC */
	void *d_comm_scratch = nullptr;      /* 256 B: IPC handle, flags, barrier words */
	/* light grids sharded over the ranks of a strip gather (comm.cpp, k_lightshare.cu) */
	unsigned char *lg_local = nullptr;          /* this rank's copy of every light grid's lists (exported through CUDA IPC) */
	unsigned char *lg_peer[UG_MAX_RANKS] = {};  /* every rank's copy ([comm_rank] = lg_local) */
	void *d_lg_handles = nullptr;
	size_t lg_bytes = 0, lg_slot = 0;           /* bytes of the buffer; pair capacity of one rank's slot of a light's value list */
	int lg_lights = 0, lg_cells = 0;
	bool lg_ok = false;                         /* the buffers are mapped everywhere */
	bool lg_disabled = false;                   /* UG_NO_SHARD_LIGHTS=1, or a sharded frame overflowed: every rank builds the light grids in full again */
	bool lg_now = false;                        /* the frame being enqueued shards its light grids */
	u32 *d_lg_seq = nullptr;                    /* [UG_MAX_LIGHTS][2]: builds pushed so far, blocks of the running push that are done */
	bool use_push = false;
	bool flag_handshake = true;          /* push transport, frames that are complete when the call's graph is: ready / done as flags in rank 0's buffer (UG_NCCL_HANDSHAKE=1: two ncclAllReduce) */
	bool fused_gather = true;            /* push transport: the shader (and the reflection combine) store into rank 0's frame themselves;
	                                        UG_NO_FUSED_GATHER=1 keeps the copy-engine node instead */
	unsigned char *rgb_remote = nullptr; /* where this frame's pixels go on rank 0 (its own buffer or the IPC-mapped one), null without a fused gather */

	/* CUDA graph of the frame */
	cudaGraphExec_t graph_exec[2] = {nullptr, nullptr}; /* [1]: the odd frames of a deferred gather (they write the other slot) */
	unsigned long long graph_key[2] = {0, 0};
	int graph_launches[2] = {0, 0};
	bool graph_overflow[2] = {false, false};
	bool graph_own[2] = {false, false};
	int trace_item_forced = 0;  /* UG_TRACE_ITEM=64|128 */
	bool trace_small = false; /* primary rays in work items of 64 pairs (small camera grids), see run_frame_once */ /* that graph's "done" handshake carries the overflow flags */
	unsigned long long gather_seq = 0; /* frames gathered with UG_FLAG_GATHER_DEFERRED so far: frame k lands in slot k & 1 */

	ug_stats stats = {};
	bool frame_valid = false;
};

/* ---------------------------------------------------------------- error helpers */
int ug_fail(ug_context *ctx, int code, const char *fmt, ...);
void ug_drop_graphs(ug_context *ctx);
void ug_strip_rows(const ug_context *ctx, int r, int *lo, int *hi);
#define UG_CUDA(ctx, call)                                                                              \
	do {                                                                                                \
		cudaError_t e_ = (call);                                                                        \
		if (e_ != cudaSuccess) return ug_fail(ctx, UG_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while (0)

/* ---------------------------------------------------------------- kernel launchers (k_*.cu) */
/* K1 transform */
void ugk_fetch_consts(cudaStream_t s, const void *host_ring, void *dev_block, size_t bytes, u32 *seq, int slots, unsigned long long *t_start);
void ugk_publish(cudaStream_t s, const u32 *const *src, const int *words, const int *dst_off, int n, u32 *host_block);
void ugk_overflow_any(cudaStream_t s, const u32 *const *src, int n, u32 *out);
#define UG_HC_WORDS (4 * (2 + UG_MAX_LIGHTS) + UG_MAX_BOUNCES + 1 + 2) /* h_counters: grids, reflection ray counts, [last - 1] = ns of this rank's own work, [last] = ranks that overflowed */
#define UG_HC_OWN (UG_HC_WORDS - 2)
void ugk_stamp(cudaStream_t s, unsigned long long *slots, int slot);
void ugk_comm_ready(cudaStream_t s, void *flags, u32 *seq, int rank);
void ugk_comm_done(cudaStream_t s, void *flags, const u32 *seq, int rank, int world, u32 *over_word, const unsigned long long *t_start, u32 *own_ns);
#define UG_COMM_FLAGS_BYTES 4096 /* page of handshake flags behind rank 0's gather buffer */
void ugk_transform(cudaStream_t s, float *vertexlist, const float *orig, int size, int offset, const FrameConst *fc, int sms);
/* K2 setup+bin, scan, emit */
void ugk_setup_bin(cudaStream_t s, const ViewConst *vc, const int *faces, const float *verts, int T, GridBuffers &g, int sms);
void ugk_scan_counts(cudaStream_t s, GridBuffers &g, int sms);
void ugk_emit_pairs(cudaStream_t s, const ViewConst *vc, GridBuffers &g, int sms);
void ugk_slab_reset(cudaStream_t s, const ViewConst *vc, GridBuffers &g);
void ugk_slab_union(cudaStream_t s, GridBuffers &g, int sms);
void ugk_trace_slabs(cudaStream_t s, const FrameConst *fc, const GridBuffers &g, const float *dir, float *t, int *id, float *normal, int cells, int sms);
void ugk_cell_offsets(cudaStream_t s, GridBuffers &g, int sms);
/* fused single-kernel build (k_gridfused.cu) */
bool ugk_grid_fused_supported(int device);
size_t ugk_grid_fused_state_a(int T);
size_t ugk_grid_fused_state_b(int C);
size_t ugk_grid_fused_hist(int sms);
int ugk_grid_fused(cudaStream_t s, const ViewConst *vc, const int *faces, const float *verts, GridBuffers &g, int sms);
size_t ugk_sweep_desc_a(int T);
size_t ugk_sweep_desc_b(int C);
size_t ugk_sweep_desc(size_t cap, int bins);
size_t ugk_sweep_hist_words();
size_t ugk_sweep_tiles(size_t cap);
int ugk_grid_sweep(cudaStream_t s, const ViewConst *vc, const int *faces, const float *verts, GridBuffers &g, int sms, int *launches, int part = 3);
/* radix sort / scan primitives (k_sort.cu) */
size_t ugk_scan_tiles(size_t n);
/* stable LSD radix sort of (key,val) pairs on `bits` key bits; n read from *n_dev (<= n_max).
 * Returns the index (0/1) of the buffer pair that ends up holding the result. */
int ugk_radix_sort_pairs(cudaStream_t s, u32 *keys[2], u32 *vals[2], const u32 *n_dev, size_t n_max, int bits, SortScratch &sc, int sms, int *launches);
size_t ugk_sort_hist_entries(int sms);
/* out[i] = exclusive_sum(in[0..i)), i in [0,n]; out has n+1 entries */
void ugk_exclusive_scan(cudaStream_t s, const u32 *in, u32 *out, size_t n, u64 *state, size_t state_tiles, int sms);
/* K6 primary */
void ugk_init_pixels(cudaStream_t s, const FrameConst *fc, float *t, int *id, int *shadowed, u64 *best, float *normal, float *dir, int N, int sms);
void ugk_trace(cudaStream_t s, const FrameConst *fc, const GridBuffers &g, u64 *best, u32 *cell_done, const float *dir, float *t, int *id, float *normal,
               int sms, bool small_items);
/* K7/K8 shadow (k_shadow.cu) */
bool ugk_regroup_fused_supported(int device);
size_t ugk_regroup_state_entries(int C);
size_t ugk_shadow_item_cap(int N, int C);
bool ugk_regroup(cudaStream_t s, const FrameConst *fc, int light, const GridBuffers &g, const LightRays &r, const float *t, const float *dir, int N, int C,
                 int sms, bool fused, int *launches, int phases = 7, bool recompute_dir = false);
size_t ugk_shadow_heavy_cap(int N);
void ugk_shadow_direct(cudaStream_t s, const FrameConst *fc, int light, const GridBuffers &g, const LightRays &r, const float *t, const float *dir, int *shadowed,
                       int N, int C, int sms, int *launches);
void ugk_shadow(cudaStream_t s, const FrameConst *fc, int light, const GridBuffers &g, const LightRays &r, const float *t, const float *dir, int *shadowed,
                int N, int C, int sms);
/* K10 shade */
void ugk_shade(cudaStream_t s, const FrameConst *fc, const float *t, const float *dir, const int *id, const int *shadowed, const float4 *rec, const int *mat_idx,
               const float *mat_list, int mat_count, unsigned char *rgb, int *mat_out, int N, int apply_shadow, int sms, bool spot = false,
               unsigned char *rgb_remote = nullptr, bool recompute_dir = false, const float4 *nv = nullptr);
void ugk_shade_normals(cudaStream_t s, const FrameConst *fc, const int *faces, const float *verts, int T, float4 *nv, int sms);
void ugk_add_shadows(cudaStream_t s, const int *shadowed, unsigned char *rgb, int N, int sms);
void ugk_fill_u32(cudaStream_t s, u32 *p, u32 v, size_t n, int sms);
/* row R (k_reflect.cu) */
size_t ugk_ugrid_const_bytes(void);
void ugk_reflect_grid(cudaStream_t s, ug_context *ctx, int *launches);
bool ugk_ugrid_fused_supported(int device);
bool ugk_ugrid_fused(cudaStream_t s, ug_context *ctx);
void ugk_reflect_rays(cudaStream_t s, ug_context *ctx, int bounces, int *launches);
void ugk_reflect_combine(cudaStream_t s, ug_context *ctx, int *launches);

/* comm.cpp */
void ugc_destroy(ug_context *ctx);
int ugc_share_full(ug_context *ctx, size_t bytes);
int ugc_ready(ug_context *ctx, cudaStream_t s, bool deferred);
int ugc_gather_strips(ug_context *ctx, cudaStream_t s, const unsigned char *strip, size_t slot_offset, const size_t *bytes, const size_t *offs, bool deferred,
                      bool already_stored);
int ugc_flush(ug_context *ctx, cudaStream_t s);
int ugc_allreduce_max(ug_context *ctx, cudaStream_t s, unsigned *value); /* host value in/out; synchronises s */
u32 *ugc_overflow_word(ug_context *ctx);
u32 *ugc_own_word(ug_context *ctx);
int ugc_share_light(ug_context *ctx, size_t bytes);
void ug_light_layout(const ug_context *ctx, int l, size_t *span_w, size_t *offset_w, size_t *vals_w);
void ugk_light_push(cudaStream_t s, ug_context *ctx, int l, const GridBuffers &built, int cell_base, int cells, int sms);
void ugk_light_wait(cudaStream_t s, ug_context *ctx, int l);
void ugc_unshare_light(ug_context *ctx);   /* device word the "done" handshake sums over the ranks (push transport), or null */

/* host helpers (camera_host.cpp, obj_loader.cpp) */
void ugh_view_const(ViewConst *vc, const ug_camera *cam, int W, int H, int nbx, int nby, int gy_lo, int gy_cnt, int grid_mode = UG_LIGHT_GRID_PERSPECTIVE,
                    float xM = 0.0f, float yM = 0.0f, int num_slabs = 1);

#endif
