/* display_dropin.c -- TEST PROGRAM (plain C, links libugrt.so only).
 *
 * The reference's main()/display() (main.cu:59-302, 304-400) with the binding INTEGRATION.md section 2 gives a maintainer:
 * the reference's own entry points (init_model, updateLightPosition, fillCoordinatesData, build_frustum_grid, getRayGridMapping,
 * getEffectiveRayGridMapping, build_secondary_frustum_grid, processData, check_for_shadows, writePPM) keep their names and
 * call order, the classes (fTracer, shader, camera) become ug_* calls on the default context.  tests/test_dropin.py
 * compiles this file, runs it on the GPU and compares every PPM it writes with the oracle's frame.
 *
 *   display_dropin <obj | keyframe dir> <material file> <frames> <params file> <output prefix>
 *
 * params file: width height fovy near far / eye look up (9 floats) / light pos look up (9 floats), whitespace separated.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "ugrt.h"

static int frame_cnt = 0;         /* main.cu:33 */
static int h_numLights = 1;       /* main.cu.h:40 */
static float P[5 + 9 + 9];
static const char *out_prefix;
static int keyframed = 0;

static void display(void)
{
	char imgname[1024];
	ug_camera *camera = ug_default_camera();                    /* the global `camera`, main.cu:28 */
	frame_cnt++;
	if (keyframed) ug_default_tmp_model(frame_cnt - 1);         /* model->tmp_model(path, i), scene.h:70-120 */

	updateLightPosition();                                      /* per_frame_funcs.h:6 */
	ug_camera_setCameraCenter(camera, P[5], P[6], P[7]);        /* main.cu:87-90 */
	ug_camera_setCameraLookAt(camera, P[8], P[9], P[10]);
	ug_camera_setCameraUp(camera, P[11], P[12], P[13]);
	ug_camera_setNearFar(camera, P[3], P[4]);
	camera->fovy = P[2];
	camera->aspect = P[0] / P[1];
	ug_camera_adjustCameraAndPosition(camera);                  /* adjustCameraAndPosition + getGLMatrices + getFrustumProperties */
	fillCoordinatesData();                                      /* per_frame_funcs.h:18 (+ setDirectionTexture :161) */
	setDirectionTexture();
	build_frustum_grid();                                       /* per_frame_funcs.h:45 */
	if (ug_trace(ug_default_context()) != UG_OK) exit(3);       /* fTracer->trace(...), main.cu:142-145 */

	for (int l = 0; l < h_numLights; l++) {                     /* main.cu:148-203 */
		ug_camera_setCameraCenter(camera, P[14], P[15], P[16]);
		ug_camera_setCameraLookAt(camera, P[17], P[18], P[19]);
		ug_camera_setCameraUp(camera, P[20], P[21], P[22]);
		camera->aspect = 1.0f;
		ug_camera_adjustCameraAndPosition(camera);
		getRayGridMapping(0, 0, 0, 0, 0, 0);                    /* main.cu:186-187: result discarded */
		getEffectiveRayGridMapping(0, 0, 0, 0, (float)M_PI, (float)M_PI);   /* main.cu:189 */
		build_secondary_frustum_grid((float)M_PI, (float)M_PI);            /* main.cu:191 */
		processData();                                          /* main.cu:194 */
		check_for_shadows(l);                                   /* main.cu:197 */
	}
	if (frame_cnt < 2) { if (ug_simple_shade(ug_default_context()) != UG_OK) exit(4); }       /* main.cu:205-211 */
	else { if (ug_spotlight_shade(ug_default_context()) != UG_OK) exit(4); }                  /* main.cu:213-218 */
	if (ug_add_shadows(ug_default_context()) != UG_OK) exit(5);                               /* main.cu:223 */
	snprintf(imgname, sizeof(imgname), "%s-%d.ppm", out_prefix, frame_cnt);                   /* main.cu:231-233 */
	writePPM(imgname);
}

int main(int argc, char **argv)
{
	init_model(argc < 4 ? argc : 4, argv);                      /* main.cu:380: usage + exit(-1) without a model */
	if (argc < 6) { fprintf(stderr, "usage: %s obj|dir material frames params out-prefix\n", argv[0]); return 2; }
	FILE *f = fopen(argv[4], "r");
	if (!f) { perror(argv[4]); return 2; }
	for (int i = 0; i < 23; i++) if (fscanf(f, "%f", &P[i]) != 1) { fprintf(stderr, "params: 23 numbers expected\n"); return 2; }
	fclose(f);
	out_prefix = argv[5];
	keyframed = ug_default_num_frames() > 1;
	if (ug_set_screen(ug_default_context(), (int)P[0], (int)P[1], P[2]) != UG_OK) return 2;   /* SCREEN_WIDTH/HEIGHT/FOVY, main.cu.h:10-14 */
	ug_light L;
	for (int k = 0; k < 3; k++) { L.pos[k] = P[14 + k]; L.look[k] = P[17 + k]; L.up[k] = P[20 + k]; }
	ug_default_set_light(0, &L);                                /* dd_light_position, per_frame_funcs.h:8-10 */
	const int frames = atoi(argv[3]);
	for (int i = 0; i < frames; i++) display();                 /* glutMainLoop -> display(), MAXFRAMES times (main.cu:297) */
	return 0;
}
