/* ugrt.h -- C-ABI of the B200-native per-frame hot path of sushruta/uniformgrid-raytracing.
 *
 * Plain C: pointers, sizes, PODs.  No torch / C++ types.  Implemented by
 * uniformgrid-raytracing_b200/libugrt.so (hand-written sm_100a CUDA; there is NO CPU fallback - every
 * entry point that computes returns UG_ERR_CUDA when no device is usable).
 *
 * The reference has no FFI: its "API" is the set of free functions and classes that display()
 * (main.cu:59-302) calls on the globals model/fGrid/camera/fTracer/dData/shader (main.cu:26-31).
 * Each function below names the reference interface it replaces (file:line, relative to the
 * reference root).  Two layers are exported:
 *   1. handle-based `ug_*` functions (state explicit, re-entrant, an int status instead of exit());
 *   2. the reference's own names (`init_model`, `build_frustum_grid`, ...) acting on one default
 *      context, so display() can keep its call sequence.
 *
 * Conventions: all buffers are the reference's layouts: pixel id = py*W+px with
 * row 0 the bottom of the view; cell id = gx*NBY+gy (x-major), tile = 8x8 pixels; image = u8[3N] RGB.
 * Every call is asynchronous on the context's stream unless it returns host data.
 */
#ifndef UGRT_H
#define UGRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UG_OK 0
#define UG_ERR_ARG 1      /* bad argument / call order */
#define UG_ERR_CUDA 2     /* CUDA runtime error or no device (ug_last_error has the text) */
#define UG_ERR_IO 3       /* file could not be read / written */
#define UG_ERR_OVERFLOW 4 /* pair arena too small even after growing (frame was re-run and still overflowed) */

#define UG_TILE 8          /* NUM_THREADS_X/Y, main.cu.h:24-25 */
#define UG_MAX_LIGHTS 8    /* reference capacity is 2 (main.cu.h:61,63); only light 0 shades (shader_kernel.cu:52) */
#define UG_MATERIAL_SIZE 6 /* main.cu.h:33 */
#define UG_MAX_RANKS 64    /* largest communicator ug_comm_init accepts */
#define UG_MAX_SLABS 64    /* NUM_SLABS (main.cu.h:18) is 1 in the committed build; the code for more is there and so is ours */

typedef struct ug_context ug_context;

/* ---- camera.h:7-47 `class Camera`, same field names; FOVY / SCREEN_WIDTH:SCREEN_HEIGHT (main.cu.h:10-14,
 * camera.h:137) become the two extra fields.  Pure host math, no OpenGL. */
typedef struct ug_camera {
	float cx, cy, cz;
	float lx, ly, lz;
	float ux, uy, uz;
	float nearPlane, farPlane;
	float worldori[4];
	float modelview_matrix[16];
	float projection_matrix[16];
	float mvp_matrix[16];
	float frustum_plane_eq[6][6];
	float frustumcorner[8][3];
	float fovy;   /* degrees; FOVY */
	float aspect; /* W/H for the eye camera, 1 for light cameras (misc_kernel.cu:115-116 assumes 1) */
} ug_camera;

void ug_camera_setCameraCenter(ug_camera *c, float x, float y, float z);   /* camera.h:65-70 */
void ug_camera_setCameraLookAt(ug_camera *c, float x, float y, float z);   /* camera.h:72-77 */
void ug_camera_setCameraUp(ug_camera *c, float x, float y, float z);       /* camera.h:79-84 */
void ug_camera_setNearFar(ug_camera *c, float nearP, float farP);          /* camera.h:59-63 */
/* camera.h:135-148 adjustCameraAndPosition + :86-113 getGLMatrices + :115-121 getFrustumProperties */
void ug_camera_adjustCameraAndPosition(ug_camera *c);
/* per_frame_funcs.h:18-39: pack eye | 4 near corners | MV | P | MVP into the 64-float block */
void ug_camera_pack(const ug_camera *c, float camcoords[64]);
/* per_frame_funcs.h:161-434 setDirectionTexture: the 5x5 float4 lattice of the near rectangle */
void ug_direction_lattice(const float camcoords[64], float lattice[100]);

typedef struct ug_light {
	float pos[3];  /* light-camera eye == dd_light_position entry (one position for both) */
	float look[3]; /* main.cu:158-164 made data */
	float up[3];
} ug_light;

/* ---- one frame = display() (main.cu:59-302) made data */
typedef struct ug_frame_params {
	int width, height;  /* multiples of UG_TILE; SCREEN_WIDTH/HEIGHT */
	float fovy;         /* FOVY, degrees */
	float nearPlane, farPlane;
	float eye[3], look[3], up[3];
	int num_lights;
	ug_light lights[UG_MAX_LIGHTS];
	int transform;      /* 1: run Model::rotate_bunny(rot_factor) first (main.cu:68) */
	float rot_factor;
	int strip_index, strip_count; /* image-strip sharding: this context renders cell rows
	                                 [strip_index*NBY/strip_count, (strip_index+1)*NBY/strip_count); 0,1 = whole frame */
	int flags;          /* UG_FLAG_* */
	int bounces;        /* row R: reflection bounces after shading (0 = the reference's committed display()); needs ug_model_set_reflect */
	int num_slabs;      /* NUM_SLABS, main.cu.h:18: depth slabs of every grid; 0 or 1 = the committed configuration.  With more, span/offset
	                       buffers hold one entry per (cell, slab) and primary rays follow rckernel_alpha's slab loop (trace_kernel.cu:132-229) */
} ug_frame_params;

#define UG_FLAG_FULL_BUFFERS 1      /* also fill the reference's normal/dir/shadowed buffers (decision_data.h:66-82) */
#define UG_FLAG_NO_GRAPH 2          /* launch kernels directly instead of replaying the captured CUDA graph */
#define UG_FLAG_STAGE_EVENTS 4      /* record one CUDA event per stage (implies NO_GRAPH); see ug_stats.ms_* */
#define UG_FLAG_GATHER_FRAMES 8     /* with ug_comm_init and no strips: every rank renders its own frame, rank 0 collects them (slot r = rank r) */
#define UG_FLAG_SPHERICAL_LIGHTS 16 /* light grids and shadow-ray binning by the spherical map the committed display() uses
                                       (main.cu:189-192 getEffectiveRayGridMapping + buildSphericalGrid with x_max = y_max = M_PI)
                                       instead of the perspective one north-star names (misc_kernel.cu:63-129) */
#define UG_FLAG_SPOT_SHADE 32       /* shade with Shader::spotlight_shade (shader.h:88-113 -> spot_shade, shader_kernel.cu:275-345) instead of
                                       simpleShade: what display() does from its third frame on (main.cu:205-219) */
#define UG_FLAG_GATHER_DEFERRED 64  /* with a communicator: do not wait for the gather at the end of the frame.  Rank 0's buffer holds two
                                       slots; gathered frame k lands in slot k & 1 and is whole once frame k+1 has run on rank 0's stream, or
                                       after ug_comm_flush.  The per-frame "done" handshake leaves the critical path (throughput mode). */

/* how a light's grid is parameterised (ug_set_light_grid_mode, UG_FLAG_SPHERICAL_LIGHTS) */
#define UG_LIGHT_GRID_PERSPECTIVE 0 /* mapSortkernel misc_kernel.cu:63-129 + DSKernel/DSFillkernel grid_kernel.cu:164-332 */
#define UG_LIGHT_GRID_SPHERICAL 1   /* mapSort_Effective_kernel misc_kernel.cu:255-296 + DS_spherical_Kernel/Fillkernel grid_kernel.cu:481-854 */

/* buffers that can be read back for parity checks / interop (ug_read, ug_device_ptr) */
enum ug_buffer {
	UG_BUF_VERTEXLIST = 0,   /* f32[3V]  scene.h:20 */
	UG_BUF_CAM_VALUES,       /* u32[P]   frustum_grid.h:28 d_triangle_value_list (camera grid) */
	UG_BUF_CAM_SPAN,         /* u32[C]   frustum_grid.h:26 (C = cells x num_slabs, entry = cell*num_slabs + slab) */
	UG_BUF_CAM_OFFSET,       /* u32[C]   frustum_grid.h:25 */
	UG_BUF_T,                /* f32[N]   decision_data.h:13 d_primary_ray_t_value (-1 on miss) */
	UG_BUF_ID,               /* i32[N]   decision_data.h:17 d_intersect_id (-2 on miss) */
	UG_BUF_NORMAL,           /* f32[3N]  (UG_FLAG_FULL_BUFFERS) */
	UG_BUF_DIR,              /* f32[3N]  (UG_FLAG_FULL_BUFFERS) */
	UG_BUF_SHADOWED,         /* i32[N]   decision_data.h:16 d_is_shadowed, OR over lights */
	UG_BUF_RGB,              /* u8[3N]   d_image */
	UG_BUF_LIGHT_VALUES,     /* u32[P_l] light grid of light `index` */
	UG_BUF_LIGHT_SPAN,       /* u32[C] */
	UG_BUF_LIGHT_OFFSET,     /* u32[C] */
	UG_BUF_LIGHT_CELL,       /* u32[N]   d_map[N..2N): light cell of every pixel's shadow ray, C = outside */
	UG_BUF_SIZELIST,         /* u32[T]   frustum_grid.h:32 d_sizeList of the camera grid */
	UG_BUF_RGB_FULL,         /* rank 0 of a communicator: the gathered frame(s) - two slots of 3*W*H*world bytes (the second one is only
	                            used by UG_FLAG_GATHER_DEFERRED); strips fill the first 3*W*H bytes of a slot, frames mode all of it */
	UG_BUF_CAM_COUNTERS,     /* u32[32]  debug: pair counters and phase time stamps of the camera-grid build */
	UG_BUF_COUNT_
};

typedef struct ug_stats {
	int num_faces, num_vertices, num_materials;
	int width, height, num_cells, num_pixels; /* of this context's strip */
	unsigned cam_pairs;                   /* P of the camera grid */
	unsigned light_pairs[UG_MAX_LIGHTS];  /* P of each light grid */
	unsigned pair_capacity;
	int overflowed;                       /* 1 if the last frame had to be re-run with a larger arena */
	int kernel_launches;                  /* kernels launched by the last frame (graph nodes count) */
	float ms_frame;                       /* CUDA-event time of the last frame (first kernel -> RGB8 done) */
	float ms_transform, ms_cam_grid, ms_trace, ms_light_grid, ms_regroup, ms_shadow, ms_shade; /* STAGE_EVENTS only */
	unsigned reflect_rays;                /* row R: live reflection rays of the last frame, summed over bounces */
	unsigned reflect_pairs;               /* row R: (cell,triangle) pairs of the uniform grid */
	float ms_own;                         /* gathered frames: first kernel -> this rank's own last store, i.e. ms_frame without the wait for the
	                                         other ranks (what a host balances strip heights with); otherwise = ms_frame */
} ug_stats;

/* ---- lifetime */
int ug_create(ug_context **out, int device);
void ug_destroy(ug_context *ctx);
const char *ug_last_error(const ug_context *ctx);
const char *ug_version(void);
/* host evaluation of the acosf the kernels use for the spherical light grid (a restatement of glibc 2.39's, which is what
 * the host-compiled reference calls): lets a CPU test pin it against libm without a GPU */
void ug_acosf_host(const float *x, float *y, size_t n);
/* host view of how the shadow pass cuts a light cell with `rays` rays and `ntri` triangles into work items:
 * out = {rays per item, list segments, segment length, items}; items <= rays + 16 is what bounds the item table */
void ug_shadow_item_shape(unsigned rays, unsigned ntri, unsigned out[4]);
/* measured FP32 issue rate of this device, FMUL + FADD lane-ops per second with every SM fully occupied (the
 * denominator of the traversal kernels' issue roofline is to be measured in the same harness, not taken from the nominal clock) */
int ug_fp32_issue_peak(ug_context *ctx, double *lane_ops_per_s);
size_t ug_frame_const_bytes(void); /* bytes of per-frame constants ug_render_frame* uploads (pinned H2D) */
int ug_set_stream(ug_context *ctx, void *cuda_stream); /* NULL: the context's own non-blocking stream */
int ug_synchronize(ug_context *ctx);
/* render into a caller-owned device buffer (e.g. the NCCL send buffer of the strip gather) instead of the
 * context's own d_image; NULL restores the internal buffer.  Must hold 3*W*H_strip bytes. */
int ug_set_rgb_buffer(ug_context *ctx, void *device_ptr, size_t bytes);
/* SURVEY 8(f)-4, the CUDA side of the reference's presenter: the frame is rendered straight into an OpenGL pixel buffer
 * object of the calling thread's current GL context (the GL draw itself, main.cu:264-295, stays with the host program).
 * register: main.cu:347 cudaGLRegisterBufferObject(pbo); map: main.cu:139 cudaGLMapBufferObject(&d_image, pbo) - d_image is the
 * PBO's storage until unmap, read it back (ug_read, ug_write_*) while it is mapped; unmap: main.cu:262
 * cudaGLUnmapBufferObject(pbo), ordered behind the frame; unregister restores the context's own d_image. */
int ug_gl_register_pbo(ug_context *ctx, unsigned int pbo);
int ug_gl_map_pbo(ug_context *ctx);
int ug_gl_unmap_pbo(ug_context *ctx);
int ug_gl_unregister_pbo(ug_context *ctx);

/* ---- scene.h `class Model` */
/* per_app_funcs.h:8-37 init_model: argv[1] OBJ, argv[2] material description (scene.h:370-439);
 * OBJ/MTL semantics follow obj_parser/ (first 3 indices of a face, prefix match on usemtl, -1 = none) */
int ug_model_load(ug_context *ctx, const char *obj_path, const char *material_file);
/* same state from host arrays (what load_model leaves in h_vertexlist/h_facelist/..., scene.h:19-23) */
int ug_model_set(ug_context *ctx, const float *vertexlist, int num_vertices, const int *facelist, int num_faces,
                 const int *materiallist_index, const float *materiallist, int num_materials);
/* host-side parse only (no device): two-call protocol, pass NULL arrays to get the counts */
int ug_obj_parse(const char *obj_path, int *num_vertices, int *num_faces, float *vertexlist, int *facelist, int *matidx);
int ug_obj_parse_vertices(const char *obj_path, int *num_vertices, float *vertexlist, int capacity);
int ug_material_parse(const char *file, int *num_materials, float *materiallist);
/* reflectivity per material: the MTL `r` value (obj_parser/obj_parser.cpp:262-266 keeps it in obj_material.reflect; the
 * 6-float material file of scene.h:370-439 has no slot for it).  Row R is an authored spec - see DESIGN.md. */
int ug_model_set_reflect(ug_context *ctx, const float *reflect, int num_materials);
/* keyframed scenes - Model::tmp_model(path, i) (scene.h:70-120): reload the vertices of <path>/f_<i>.obj into the loaded
 * model (same topology, at most the model's vertex count) and stage them to the device asynchronously on the frame's
 * stream through pinned memory; ug_model_update_vertices is the same from a host array; ug_model_bounds returns the
 * xMin..zMax the loader tracks (scene.h:92-111) */
int ug_model_tmp_model(ug_context *ctx, const char *path, int i);
int ug_model_update_vertices(ug_context *ctx, const float *vertexlist, int num_vertices);
int ug_model_bounds(ug_context *ctx, float *min_max6);
int ug_model_init_orig_list(ug_context *ctx, int vertices_size, int vertices_offset); /* scene.h:336-368 */
int ug_model_rotate_bunny(ug_context *ctx, float rot_factor);                         /* scene.h:122-139 */

/* ---- per_frame_funcs.h, one call per reference function (state lives in ctx) */
int ug_set_screen(ug_context *ctx, int width, int height, float fovy);            /* main.cu.h:10-30 made runtime */
int ug_set_strip(ug_context *ctx, int strip_index, int strip_count);
/* the same with explicit boundaries (cell rows, row_bounds[0] = 0 ... row_bounds[strip_count] = height/8): strips of unequal height for
 * load balance; ug_frame_params with the same (strip_index, strip_count) keep them, every rank of a communicator passes the same array */
int ug_set_strip_bounds(ug_context *ctx, const int *row_bounds, int strip_count, int strip_index);
int ug_set_slabs(ug_context *ctx, int num_slabs);                                 /* main.cu.h:18 NUM_SLABS made runtime */
int ug_update_light_position(ug_context *ctx, const float *xyz, int num_lights);  /* per_frame_funcs.h:6-16 */
int ug_fill_coordinates_data(ug_context *ctx, const ug_camera *cam);              /* :18-43 (+ setDirectionTexture) */
int ug_build_frustum_grid(ug_context *ctx);                                       /* :45-55 -> frustum_grid.h:210-366 */
int ug_trace(ug_context *ctx);                                                    /* frustum_tracer.h:35-58 */
int ug_shadow_pass(ug_context *ctx, int light_id, const ug_camera *light_cam);    /* main.cu:148-203: fillCoordinatesData(light)
                                                                                     + getRayGridMapping (:72, perspective variant
                                                                                     misc_kernel.cu:63-129) + build_secondary_frustum_grid
                                                                                     (:57) + processData (:116) + check_for_shadows (:139) */
/* which binning ug_shadow_pass uses from now on: per_frame_funcs.h:97-114 getEffectiveRayGridMapping(xM, yM) and :57-70
 * build_secondary_frustum_grid(xM, yM) -> FrustumGrid::buildSphericalGrid (frustum_grid.h:368-532) when spherical */
int ug_set_light_grid_mode(ug_context *ctx, int mode, float xM, float yM);
int ug_simple_shade(ug_context *ctx);                                             /* shader.h:70-86 */
int ug_spotlight_shade(ug_context *ctx);                                          /* shader.h:88-113 (without its printf dump) */
int ug_add_shadows(ug_context *ctx);                                              /* shader.h:58-68 */

/* ---- the whole frame: every step of display() between "frame_cnt++" and the image being complete,
 * replayed as one CUDA graph, no host synchronisation inside */
int ug_render_frame(ug_context *ctx, const ug_frame_params *p); /* asynchronous: a pair arena that proved too small shows as
                                                                   ug_stats.overflowed; ug_get_stats then grows it - render again */
/* same, end to end with host buffers: uploads the frame constants, renders, copies the RGB8 strip into
 * host_rgb (3*W*H_strip bytes, pinned or pageable) and synchronises - main.cu:245 included */
int ug_render_frame_host(ug_context *ctx, const ug_frame_params *p, uint8_t *host_rgb);
/* frames in flight: returns once the frame and the read-back of its image (on a stream of its own) are enqueued, so the next
 * frame renders while this image crosses PCIe; host_rgb must be pinned.  ug_wait_frame(ctx, 0): the last submitted frame's
 * image is in host memory (UG_ERR_OVERFLOW: an arena was too small, it has been grown, submit the frame again);
 * ug_wait_frame(ctx, k), k = 1..3: the image of the frame submitted k calls earlier is.  display()'s cudaMemcpy of the image
 * (main.cu:245), taken off the frame's critical path. */
int ug_render_frame_host_async(ug_context *ctx, const ug_frame_params *p, uint8_t *pinned_host_rgb);
int ug_wait_frame(ug_context *ctx, int frames_back);

/* ---- read-back / interop */
int ug_get_stats(ug_context *ctx, ug_stats *out);                                        /* synchronises */
int ug_read(ug_context *ctx, int buffer, int index, void *dst, size_t bytes);            /* synchronises; D2H */
int ug_device_ptr(ug_context *ctx, int buffer, int index, void **ptr, size_t *bytes);    /* no copy */
int ug_timeline(ug_context *ctx, double *ns64); /* development aid, needs env UG_TIMELINE=1 at ug_create: marker times of the last frame */
int ug_write_ppm_binary(ug_context *ctx, const char *filename); /* same image as binary P6, top row first */
int ug_write_png(ug_context *ctx, const char *filename);        /* the same image as an 8-bit RGB PNG, top row first: replaces writePPM + `convert -flip` of
                                                                    main.cu:255-259 (stored deflate blocks, no zlib dependency) */
int ug_write_ppm(ug_context *ctx, const char *filename);                                 /* per_app_funcs.h:39-66, ASCII P3 */
/* ms between the stage marks of the last UG_FLAG_STAGE_EVENTS frame, in order: transform, camera grid, primary rays,
 * then per light (light grid, ray->cell map + regroup, shadow rays), then shade; returns the number of intervals
 * written (these are the cutCreateTimer printf's of the reference) */
int ug_stage_times(ug_context *ctx, float *out_ms, int max_entries);

/* ---- multi-GPU: one context per GPU/process, image strips (ug_frame_params.strip_*), scene replicated.  With a communicator
 * every ug_render_frame* ends with the gather of the RGB8 strips to rank 0 over NCCL (part of the frame's CUDA graph);
 * ug_render_frame_host then delivers the FULL frame on rank 0 (other ranks get nothing).  The 128-byte id comes from
 * ug_comm_unique_id on one rank and is handed to the others by the host program (MPI, torch.distributed, a file ...). */
int ug_comm_unique_id(void *out128);
int ug_comm_init(ug_context *ctx, int rank, int world, const void *unique_id128);
int ug_comm_flush(ug_context *ctx); /* collective: completes the last UG_FLAG_GATHER_DEFERRED frame on rank 0 and synchronises */

/* ---- the reference's own entry points on a process-wide default context */
void init_model(int argc, char **argv);                  /* per_app_funcs.h:8 */
void writePPM(char *filename);                           /* per_app_funcs.h:39 */
void updateLightPosition(void);                          /* per_frame_funcs.h:6 */
void fillCoordinatesData(void);                          /* per_frame_funcs.h:18 */
void setDirectionTexture(void);                          /* per_frame_funcs.h:161 */
void build_frustum_grid(void);                           /* per_frame_funcs.h:45 */
void build_secondary_frustum_grid(float xM, float yM);   /* per_frame_funcs.h:57 */
void getRayGridMapping(float *t, float *dir, float *x_list, float *y_list, float *angle, float *cam_pt);   /* :72 */
void getEffectiveRayGridMapping(float *t, float *dir, unsigned int *map, float *cam_pt, float xM, float yM); /* :97 */
void processData(void);                                  /* per_frame_funcs.h:116 */
void check_for_shadows(int light_id);                    /* per_frame_funcs.h:139 */
ug_context *ug_default_context(void);
ug_camera *ug_default_camera(void);                      /* the global `camera` (main.cu:28) */
void ug_default_set_light(int light_id, const ug_light *l);
void ug_default_tmp_model(int i);                        /* model->tmp_model(argv[1], i), scene.h:70-120: keyframe i of the directory init_model was given */
int ug_default_num_frames(void);                         /* Model::num_frames (scene.h:33): argv[3] of a directory scene, else 1 */

#ifdef __cplusplus
}
#endif
#endif /* UGRT_H */
