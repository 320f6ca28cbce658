/* ref_compat.cpp -- the reference's own entry points (per_app_funcs.h, per_frame_funcs.h) on a
 * process-wide default context, so that display() (main.cu:59-302) can keep its call sequence when
 * this library replaces the .cu/.h files of the hot path.  The reference keeps its state in the
 * globals model/fGrid/camera/fTracer/dData/shader (main.cu:26-31); here that state is the default
 * ug_context plus the default ug_camera.  Errors abort like cutilSafeCall does (exit(-1)).
 */
#include "ugrt_internal.h"
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>

static ug_context *g_ctx = nullptr;
static ug_camera g_camera;
static ug_light g_lights[UG_MAX_LIGHTS];
static float g_light_xyz[3 * UG_MAX_LIGHTS] = {10.0f, 12.0f, 6.0f}; /* per_frame_funcs.h:8-10 */
static int g_num_lights = 1;                                        /* main.cu.h:40 h_numLights */
static int g_current_light = 0;
static int g_frames = 1;             /* Model::num_frames (scene.h:33, 59-67): > 1 = keyframed scene */
static std::string g_model_path;     /* argv[1]: the OBJ, or the directory holding f_<i>.obj */

static void die(int rc, const char *what)
{
	if (rc == UG_OK) return;
	fprintf(stderr, "%s failed (%d): %s\n", what, rc, ug_last_error(g_ctx));
	exit(-1);
}

extern "C" {

ug_context *ug_default_context(void)
{
	if (!g_ctx) {
		int dev = 0;
		const char *e = getenv("UG_DEVICE");
		if (e) dev = atoi(e);
		int rc = ug_create(&g_ctx, dev);
		if (rc != UG_OK) { fprintf(stderr, "ug_create failed (%d): no usable CUDA device\n", rc); exit(-1); }
		memset(&g_camera, 0, sizeof(g_camera));
		g_camera.fovy = 45.0f;   /* FOVY, main.cu.h:14 */
		g_camera.aspect = 1.0f;  /* 1024x1024, main.cu.h:10-11 */
		die(ug_set_screen(g_ctx, 1024, 1024, 45.0f), "ug_set_screen");
	}
	return g_ctx;
}

ug_camera *ug_default_camera(void) { ug_default_context(); return &g_camera; }

void ug_default_set_light(int light_id, const ug_light *l)
{
	if (light_id < 0 || light_id >= UG_MAX_LIGHTS || !l) return;
	g_lights[light_id] = *l;
	for (int k = 0; k < 3; k++) g_light_xyz[3 * light_id + k] = l->pos[k];
	if (g_num_lights <= light_id) g_num_lights = light_id + 1;
}

/* per_app_funcs.h:8-37 */
void init_model(int argc, char **argv)
{
	if (argc < 2) {
		fprintf(stderr, "USAGE: ./exec {file/dir_name}.obj [number_frames]\n");
		exit(-1);
	}
	/* per_app_funcs.h:14-30: a directory means a keyframed scene, f_0.obj ... f_<frames-1>.obj, and needs the frame count */
	int frames = 1;
	struct stat sb;
	const bool is_dir = stat(argv[1], &sb) == 0 && S_ISDIR(sb.st_mode);
	if (is_dir) {
		if (argc < 4) {
			fprintf(stderr, "USAGE: ./exec {dir_name}.ply [desc_file] [number_frames]\n");
			exit(-1);
		}
		frames = atoi(argv[3]);
	}
	ug_context *ctx = ug_default_context();
	g_frames = frames;
	g_model_path = argv[1];
	/* Model::Model(frames): dynamic iff frames > 1 (scene.h:59-67); Model::load_model then reads <dir>/f_0.obj once and
	 * keeps its topology (scene.h:142-222), later frames come through Model::tmp_model == ug_default_tmp_model.  Unlike the
	 * reference's dynamic branch, which never fills h_materiallist_index, the first keyframe's usemtl indices are kept. */
	std::string first = g_model_path;
	if (is_dir && frames > 1) first += "/f_0.obj";
	die(ug_model_load(ctx, first.c_str(), argc > 2 ? argv[2] : NULL), "init_model");
}

/* Model::tmp_model(path, i) on the default context's model, path = what init_model was given (scene.h:70-120) */
void ug_default_tmp_model(int i)
{
	die(ug_model_tmp_model(ug_default_context(), g_model_path.c_str(), i), "tmp_model");
}
int ug_default_num_frames(void) { return g_frames; }

/* per_app_funcs.h:39-66 */
void writePPM(char *filename) { die(ug_write_ppm(ug_default_context(), filename), "writePPM"); }

/* per_frame_funcs.h:6-16 (positions come from ug_default_set_light; default = the reference's literal) */
void updateLightPosition(void) { die(ug_update_light_position(ug_default_context(), g_light_xyz, g_num_lights), "updateLightPosition"); }

/* per_frame_funcs.h:18-43: upload the 64-float block of the CURRENT global camera (+ lattice) */
void fillCoordinatesData(void) { die(ug_fill_coordinates_data(ug_default_context(), &g_camera), "fillCoordinatesData"); }

/* per_frame_funcs.h:161-434: folded into fillCoordinatesData (which calls it at :40) */
void setDirectionTexture(void) {}

/* per_frame_funcs.h:45-55 */
void build_frustum_grid(void) { die(ug_build_frustum_grid(ug_default_context()), "build_frustum_grid"); }

/* The reference's per-light sequence (main.cu:172-199) is getRayGridMapping -> getEffectiveRayGridMapping ->
 * build_secondary_frustum_grid -> processData -> check_for_shadows, all on the light camera that
 * fillCoordinatesData made resident.  Here the five calls are one fused pass; it runs when
 * check_for_shadows() is reached, the other four record their arguments. */
void getRayGridMapping(float *, float *, float *, float *, float *, float *) {}             /* result discarded in the reference, main.cu:186-187 */
/* these two are the SPHERICAL variants (per_frame_funcs.h:97-114 -> mapSort_Effective_kernel, :57-70 -> buildSphericalGrid):
 * a display() that calls them gets the spherical light grid with its extents, as in the reference */
void getEffectiveRayGridMapping(float *, float *, unsigned int *, float *, float xM, float yM)
{
	die(ug_set_light_grid_mode(ug_default_context(), UG_LIGHT_GRID_SPHERICAL, xM, yM), "getEffectiveRayGridMapping");
}
void build_secondary_frustum_grid(float xM, float yM)
{
	die(ug_set_light_grid_mode(ug_default_context(), UG_LIGHT_GRID_SPHERICAL, xM, yM), "build_secondary_frustum_grid");
}
void processData(void) {}
void check_for_shadows(int light_id)
{
	g_current_light = light_id;
	die(ug_shadow_pass(ug_default_context(), light_id, &g_camera), "check_for_shadows");
}

} /* extern "C" */
