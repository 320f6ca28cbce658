/* ug_oracle.c -- TEST INFRASTRUCTURE ONLY (see ug_oracle.h for scope and pinning).
 *
 * Plain-C restatement of the reference's per-frame hot path.  Every function cites the reference
 * file:line it follows (paths relative to /root/reference).  Arithmetic contract:
 * IEEE binary32, no FMA contraction (built with -ffp-contract=off on baseline x86-64), evaluation
 * order exactly as written in the reference, the few double sub-expressions kept in double.
 */
#define _GNU_SOURCE
#include "ug_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#define TILE 8 /* NUM_THREADS_X/Y, main.cu.h:24-25 */

static int g_W = 1024, g_H = 1024, g_nbx = 128, g_nby = 128, g_fovy = 45, g_q8 = 0, g_threads = 1;
static int g_slabs = 1;    /* NUM_SLABS, main.cu.h:18 (the committed value is 1) */
static float g_cc[64];     /* dd_camcoords, main.cu.h:62; layout per_frame_funcs.h:18-39 */
static float g_light[6];   /* dd_light_position, main.cu.h:63 */
static float g_lattice[25][4]; /* texdir, main.cu.h:58 */

void orc_configure(int W, int H, int fovy, int tex_q8)
{
	g_W = W; g_H = H; g_nbx = W / TILE; g_nby = H / TILE; g_fovy = fovy; g_q8 = tex_q8;
}
void orc_set_slabs(int n) { g_slabs = n > 0 ? n : 1; }
void orc_set_threads(int n) { g_threads = n > 0 ? n : omp_get_max_threads(); }
int orc_get_threads(void) { return g_threads; }

/* ------------------------------------------------------------------ small vector helpers
 * main.cu.h:44-56: DOT is ((a0*b0 + a1*b1) + a2*b2); CROSS is a*b - c*d per component;
 * NORMALIZE multiplies by 1/sqrtf(len2). */
static inline float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(float *d, const float *a, const float *b)
{
	d[0] = a[1] * b[2] - a[2] * b[1];
	d[1] = a[2] * b[0] - a[0] * b[2];
	d[2] = a[0] * b[1] - a[1] * b[0];
}
static inline void normalize3(float *a)
{
	float l = 1 / sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
	a[0] *= l; a[1] *= l; a[2] *= l;
}
/* grid_kernel.cu:4-11 / trace_kernel.cu:47-54: column-major 4x4 times vec4, left-to-right sums */
static inline void mat4_mul_vec4(float *r, const float *m, const float *v)
{
	for (int i = 0; i < 4; i++) r[i] = m[i] * v[0] + m[4 + i] * v[1] + m[8 + i] * v[2] + m[12 + i] * v[3];
}
/* float -> int as x86 cvttss2si does it (what the host-compiled reference executes) */
static inline int f2i(float f) { return (int)f; }

/* ------------------------------------------------------------------ camera (camera.h, GLU)
 * gluPerspective / gluLookAt restated per libGLU (perspective in double then stored float, look-at in
 * float); the GL matrix stack multiplies float matrices.  The reference read these back from the GL
 * driver (camera.h:86-90): parity unpinned, both sides of every test use this same definition. */
static void mat4_mul(float *out, const float *c, const float *m) /* out = c * m, column-major */
{
	float r[16];
	for (int col = 0; col < 4; col++)
		for (int row = 0; row < 4; row++)
			r[col * 4 + row] = c[0 * 4 + row] * m[col * 4 + 0] + c[1 * 4 + row] * m[col * 4 + 1] +
			                   c[2 * 4 + row] * m[col * 4 + 2] + c[3 * 4 + row] * m[col * 4 + 3];
	memcpy(out, r, sizeof(r));
}
static void identity4(float *m) { for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.0f : 0.0f; }

static void perspective(float *proj, double fovy, double aspect, double zn, double zf)
{
	double rad = fovy / 2 * 3.14159265358979323846 / 180, dz = zf - zn, s = sin(rad), cot = cos(rad) / s;
	float m[16] = {0};
	m[0] = (float)(cot / aspect);
	m[5] = (float)cot;
	m[10] = (float)(-(zf + zn) / dz);
	m[11] = -1.0f;
	m[14] = (float)(-2 * zn * zf / dz);
	identity4(proj);
	mat4_mul(proj, proj, m);
}
static void look_at(float *mv, const float *eye, const float *ctr, const float *upv)
{
	float f[3], up[3] = {upv[0], upv[1], upv[2]}, s[3], u[3];
	/* the reference passes floats, promoted to double; GLU narrows the differences back to float */
	for (int k = 0; k < 3; k++) f[k] = (float)((double)ctr[k] - (double)eye[k]);
	float r = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
	if (r != 0.0f) { f[0] /= r; f[1] /= r; f[2] /= r; }
	cross3(s, f, up);
	r = sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
	if (r != 0.0f) { s[0] /= r; s[1] /= r; s[2] /= r; }
	cross3(u, s, f);
	float m[16] = {s[0], u[0], -f[0], 0, s[1], u[1], -f[1], 0, s[2], u[2], -f[2], 0, 0, 0, 0, 1};
	float t[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, -eye[0], -eye[1], -eye[2], 1};
	identity4(mv);
	mat4_mul(mv, mv, m);
	mat4_mul(mv, mv, t);
}
/* camera.h:218-239 Intersect3Planes */
static void three_planes(float *p, const float *n1, const float *n2, const float *n3)
{
	float n2n3[3], n3n1[3], n1n2[3];
	n1n2[0] = (n1[1] * n2[2] - n2[1] * n1[2]); n1n2[1] = (n1[2] * n2[0] - n1[0] * n2[2]); n1n2[2] = (n1[0] * n2[1] - n2[0] * n1[1]);
	n2n3[0] = (n2[1] * n3[2] - n3[1] * n2[2]); n2n3[1] = (n2[2] * n3[0] - n2[0] * n3[2]); n2n3[2] = (n2[0] * n3[1] - n3[0] * n2[1]);
	n3n1[0] = (n3[1] * n1[2] - n1[1] * n3[2]); n3n1[1] = (n3[2] * n1[0] - n3[0] * n1[2]); n3n1[2] = (n3[0] * n1[1] - n1[0] * n3[1]);
	float den = n1[0] * n2n3[0] + n1[1] * n2n3[1] + n1[2] * n2n3[2];
	for (int k = 0; k < 3; k++) p[k] = -(n1[3] * n2n3[k] + n2[3] * n3n1[k] + n3[3] * n1n2[k]) / den;
}

/* per_frame_funcs.h:161-434 setDirectionTexture: texel(col c,row r) = lerp(lerp(c0,c1,c/4),
 * lerp(c3,c2,c/4), r/4); (b-a) in float, weight*(b-a) and the add in double, stored float;
 * end columns/rows are copies. */
static void build_lattice(const float *cc)
{
	const float *c0 = cc + 4, *c1 = cc + 7, *c2 = cc + 10, *c3 = cc + 13;
	static const double w[5] = {0.0, 0.25, 0.5, 0.75, 1.0};
	for (int r = 0; r < 5; r++)
		for (int c = 0; c < 5; c++) {
			for (int k = 0; k < 3; k++) {
				float lo = c0[k], hi = c3[k];
				if (c == 4) { lo = c1[k]; hi = c2[k]; }
				else if (c != 0) { lo = (float)(c0[k] + w[c] * (c1[k] - c0[k])); hi = (float)(c3[k] + w[c] * (c2[k] - c3[k])); }
				float val = lo;
				if (r == 4) val = hi;
				else if (r != 0) val = (float)(lo + w[r] * (hi - lo));
				g_lattice[r * 5 + c][k] = val;
			}
			g_lattice[r * 5 + c][3] = 0.0f;
		}
}

/* main.cu:72-128 + camera.h:135-253 + per_frame_funcs.h:18-43 */
void orc_camera_setup(const float *pose9, float nearP, float farP, int square_aspect, float *out)
{
	float mv[16], proj[16], mvp[16], plane[6][4], corner[4][3];
	float aspect = square_aspect ? 1.0f : (float)g_W / (float)g_H; /* camera.h:137 */
	perspective(proj, g_fovy, aspect, nearP, farP);
	look_at(mv, pose9, pose9 + 3, pose9 + 6);
	/* camera.h:150-165 getMVPMatrix: mvp[i*4+k] = sum_j mv[i*4+j]*proj[j*4+k], starting from 0 */
	for (int i = 0; i < 4; i++)
		for (int k = 0; k < 4; k++) {
			float acc = 0;
			for (int j = 0; j < 4; j++) acc += mv[i * 4 + j] * proj[j * 4 + k];
			mvp[i * 4 + k] = acc;
		}
	/* camera.h:167-216 planes: 0: row3-row0, 1: row3+row0, 2: row3+row1, 3: row3-row1, 4: row3+row2, 5: row3-row2 */
	static const int prow[6] = {0, 0, 1, 1, 2, 2};
	static const float psign[6] = {-1, 1, 1, -1, 1, -1};
	for (int p = 0; p < 6; p++) {
		float e[4];
		for (int k = 0; k < 4; k++)
			e[k] = psign[p] > 0 ? mvp[4 * k + 3] + mvp[4 * k + prow[p]] : mvp[4 * k + 3] - mvp[4 * k + prow[p]];
		float len = sqrtf(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
		for (int k = 0; k < 4; k++) plane[p][k] = e[k] / len;
	}
	/* camera.h:241-246 near corners */
	three_planes(corner[0], plane[0], plane[2], plane[4]);
	three_planes(corner[1], plane[1], plane[2], plane[4]);
	three_planes(corner[2], plane[1], plane[3], plane[4]);
	three_planes(corner[3], plane[0], plane[3], plane[4]);
	/* per_frame_funcs.h:18-39 */
	g_cc[0] = pose9[0]; g_cc[1] = pose9[1]; g_cc[2] = pose9[2]; g_cc[3] = 1.0f;
	for (int i = 0; i < 4; i++)
		for (int k = 0; k < 3; k++) g_cc[4 + 3 * i + k] = corner[i][k];
	memcpy(g_cc + 16, mv, sizeof(mv));
	memcpy(g_cc + 32, proj, sizeof(proj));
	memcpy(g_cc + 48, mvp, sizeof(mvp));
	build_lattice(g_cc);
	if (out) memcpy(out, g_cc, sizeof(g_cc));
}
void orc_set_camcoords(const float *cc) { memcpy(g_cc, cc, sizeof(g_cc)); build_lattice(g_cc); }
void orc_get_dirtex(float *out100) { memcpy(out100, g_lattice, sizeof(g_lattice)); }
void orc_set_lights(const float *lights, int n) { for (int i = 0; i < 6; i++) g_light[i] = (i < 3 * n) ? lights[i] : 0.0f; }

/* ------------------------------------------------------------------ a1: vertex transform
 * transformation_kernel.cu:4-18.  cosf/sinf are evaluated per vertex there; same value every time. */
void orc_transform(float *vertexlist, float *orig, int size, int offset, float rot)
{
	const float c = cosf(rot), s = sinf(rot);
#pragma omp parallel for num_threads(g_threads)
	for (int v = 0; v < size; v++) {
		float x = (orig[v * 3 + 0] - 12.0f) / 12.0f;
		float y = (orig[v * 3 + 1] - 11.0f) / 12.0f;
		float z = (orig[v * 3 + 2] - 4.5f) / 12.0f;
		float *o = vertexlist + (size_t)(offset + v) * 3;
		o[0] = (x * c - y * s) * 9.0f + 14.5f;
		o[1] = (x * s + y * c) * 9.0f + 13.0f;
		o[2] = z * 9.0f + 4.0f;
	}
}

/* ------------------------------------------------------------------ a2-a7: perspective grid
 * grid_kernel.cu:13-36 getTransformedVertex: MV, divide by w, P, divide by w */
static void project_vertex(const float *v, float *ndc)
{
	float p[4] = {v[0], v[1], v[2], 1.0f}, q[4];
	mat4_mul_vec4(q, g_cc + 16, p);
	p[0] = q[0] / q[3]; p[1] = q[1] / q[3]; p[2] = q[2] / q[3]; p[3] = 1.0f;
	mat4_mul_vec4(q, g_cc + 32, p);
	ndc[0] = q[0] / q[3]; ndc[1] = q[1] / q[3]; ndc[2] = q[2] / q[3];
}
/* grid_kernel.cu:132-146: NaN-propagating ternaries, not fminf/fmaxf */
static inline float min3(float a, float b, float c) { return (a < b) ? ((a < c) ? a : c) : ((b < c) ? b : c); }
static inline float max3(float a, float b, float c) { return (a > b) ? ((a > c) ? a : c) : ((b > c) ? b : c); }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* grid_kernel.cu:164-243 (count) and :245-332 (fill) share this: cell rectangle of one triangle */
static void tri_cells(const int *faces, const float *verts, int t, int *gx0, int *gx1, int *gy0, int *gy1, float *zmin)
{
	float n[3][3];
	for (int k = 0; k < 3; k++) project_vertex(verts + 3 * (size_t)faces[3 * t + k], n[k]);
	float xmin = min3(n[0][0], n[1][0], n[2][0]), ymin = min3(n[0][1], n[1][1], n[2][1]);
	float xmax = max3(n[0][0], n[1][0], n[2][0]), ymax = max3(n[0][1], n[1][1], n[2][1]);
	*zmin = min3(n[0][2], n[1][2], n[2][2]);
	*gx0 = clampi(f2i(floorf(((xmin + 1.0f) / 2.0f) * g_nbx)), 0, g_nbx - 1);
	*gy0 = clampi(f2i(floorf(((ymin + 1.0f) / 2.0f) * g_nby)), 0, g_nby - 1);
	*gx1 = clampi(f2i(floorf(((xmax + 1.0f) / 2.0f) * g_nbx)), 0, g_nbx - 1);
	*gy1 = clampi(f2i(floorf(((ymax + 1.0f) / 2.0f) * g_nby)), 0, g_nby - 1);
}

/* ------------------------------------------------------------------ a14/a15: spherical ("effective") light grid
 * grid_kernel.cu:395-421 getEffective_x: angle between the forward axis and the direction with its up component
 * removed, binned as NBX/2 +- (int)((angle/max)*(NBX/2)); the unsigned result is read back as int by the callers.
 * g_cc[16..] = modelview of the light camera: rows right (16,20,24), up (17,21,25), -forward (18,22,26).
 * acosf / sqrtf are libm's here, as in the host-compiled reference. */
static int effective_x(const float *vec, float maxa)
{
	const float *cc = g_cc;
	float upDot = vec[0] * cc[16 + 1] + vec[1] * cc[16 + 5] + vec[2] * cc[16 + 9];
	float tmp[3] = {vec[0] - upDot * cc[16 + 1], vec[1] - upDot * cc[16 + 5], vec[2] - upDot * cc[16 + 9]};
	float val = sqrtf(tmp[0] * tmp[0] + tmp[1] * tmp[1] + tmp[2] * tmp[2]); /* getMagnitude :355-363 */
	tmp[0] /= val; tmp[1] /= val; tmp[2] /= val;
	float fwd = tmp[0] * cc[16 + 2] + tmp[1] * cc[16 + 6] + tmp[2] * cc[16 + 10];
	float angle = acosf(fwd);
	float rightDot = tmp[0] * cc[16 + 0] + tmp[1] * cc[16 + 4] + tmp[2] * cc[16 + 8];
	unsigned half = (unsigned)(g_nbx / 2), k = (unsigned)f2i((angle / maxa) * (g_nbx / 2)); /* wraps like the int sum does on x86 */
	return (int)(rightDot > 0 ? half + k : half - k);
}
/* grid_kernel.cu:452-479 getEffective_y: same with right removed - but the forward dot product is written
 * `a + b*c*d` (:468) and the bin is computed in float and converted float -> unsigned (:475-478). */
static int effective_y(const float *vec, float maxa)
{
	const float *cc = g_cc;
	float rightDot = vec[0] * cc[16 + 0] + vec[1] * cc[16 + 4] + vec[2] * cc[16 + 8];
	float tmp[3] = {vec[0] - rightDot * cc[16 + 0], vec[1] - rightDot * cc[16 + 4], vec[2] - rightDot * cc[16 + 8]};
	float val = sqrtf(tmp[0] * tmp[0] + tmp[1] * tmp[1] + tmp[2] * tmp[2]);
	tmp[0] /= val; tmp[1] /= val; tmp[2] /= val;
	float upDot = tmp[0] * cc[16 + 1] + tmp[1] * cc[16 + 5] + tmp[2] * cc[16 + 9];
	float fwd = tmp[0] * cc[16 + 2] + tmp[1] * cc[16 + 6] * tmp[2] * cc[16 + 10];
	float angle = acosf(fwd);
	unsigned bly;
	if (upDot > 0) bly = (unsigned)(long long)(g_nby / 2 + (angle / maxa) * (g_nby / 2)); /* cvttss2si r64, low half (NaN -> 0) */
	else bly = (unsigned)(long long)(g_nby / 2 - (angle / maxa) * (g_nby / 2));
	return (int)bly;
}
static inline int imin3(int a, int b, int c) { return (a < b) ? ((a < c) ? a : c) : ((b < c) ? b : c); }
static inline int imax3(int a, int b, int c) { return (a > b) ? ((a > c) ? a : c) : ((b > c) ? b : c); }

/* grid_kernel.cu:481-659 DS_spherical_Kernel (count) and :661-854 DS_spherical_Fillkernel share this */
static void tri_cells_sph(const int *faces, const float *verts, int t, float xM, float yM, int *gx0, int *gx1, int *gy0, int *gy1, float *zmin)
{
	int bx[3], by[3];
	float radius[3];
	for (int k = 0; k < 3; k++) {
		const float *v = verts + 3 * (size_t)faces[3 * t + k];
		float p[3] = {v[0] - g_cc[0], v[1] - g_cc[1], v[2] - g_cc[2]};
		radius[k] = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
		p[0] /= radius[k]; p[1] /= radius[k]; p[2] /= radius[k];
		bx[k] = effective_x(p, xM);
		by[k] = effective_y(p, yM);
	}
	*zmin = min3(radius[0], radius[1], radius[2]); /* :613 */
	*gx0 = clampi(imin3(bx[0], bx[1], bx[2]), 0, g_nbx - 1);
	*gy0 = clampi(imin3(by[0], by[1], by[2]), 0, g_nby - 1);
	*gx1 = clampi(imax3(bx[0], bx[1], bx[2]), 0, g_nbx - 1);
	*gy1 = clampi(imax3(by[0], by[1], by[2]), 0, g_nby - 1);
	/* :627-639: the "wrap" branch of size_x / size_y equals gmax - gmin + 1 */
}

/* mode 0: perspective rectangle, 1: spherical */
static void tri_rect(int mode, float xM, float yM, const int *faces, const float *verts, int t, int *r, float *z)
{
	if (mode) tri_cells_sph(faces, verts, t, xM, yM, &r[0], &r[1], &r[2], &r[3], z);
	else tri_cells(faces, verts, t, &r[0], &r[1], &r[2], &r[3], z);
}

static void grid_count(int mode, float xM, float yM, int *faces, float *verts, int T, unsigned *sizeList, float *projZ)
{
#pragma omp parallel for num_threads(g_threads)
	for (int t = 0; t < T; t++) {
		int r[4];
		tri_rect(mode, xM, yM, faces, verts, t, r, &projZ[t]);
		sizeList[t] = (unsigned)((r[1] - r[0] + 1) * (r[3] - r[2] + 1)); /* grid_kernel.cu:231 / :641 */
	}
}
void orc_grid_count(int *faces, float *verts, int T, unsigned *sizeList, float *projZ) { grid_count(0, 0, 0, faces, verts, T, sizeList, projZ); }
void orc_grid_count_spherical(int *faces, float *verts, int T, unsigned *sizeList, float *projZ, float xM, float yM)
{
	grid_count(1, xM, yM, faces, verts, T, sizeList, projZ);
}

/* frustum_grid.h:210-366: count -> inclusive scan -> emit (x-major rectangle order, key =
 * gx*NBY+gy, grid_kernel.cu:318-325) -> stable sort by key (cudppSort = LSD radix) -> per-cell span
 * -> exclusive scan over ALL cells. Implemented as a counting sort, which is what a stable sort by
 * a small integer key is. */
static long build_grid(int mode, float xM, float yM, int *faces, float *verts, int T, unsigned *span, unsigned *offset, unsigned *key_list,
                       unsigned *value_list, long cap)
{
	const int S = g_slabs, C = g_nbx * g_nby * S;
	int *rect = (int *)malloc(sizeof(int) * 4 * (size_t)(T > 0 ? T : 1));
	float *projZ = (float *)malloc(sizeof(float) * (size_t)(T > 0 ? T : 1));
	unsigned *slab = (unsigned *)calloc((size_t)(T > 0 ? T : 1), sizeof(unsigned)); /* d_zSlabs; the reference never initialises it: 0 here */
	long P = 0;
#pragma omp parallel for num_threads(g_threads) reduction(+ : P)
	for (int t = 0; t < T; t++) {
		tri_rect(mode, xM, yM, faces, verts, t, &rect[4 * t], &projZ[t]);
		P += (unsigned)((rect[4 * t + 1] - rect[4 * t] + 1) * (rect[4 * t + 3] - rect[4 * t + 2] + 1));
	}
	if (P > cap) { free(rect); free(projZ); free(slab); return P; }
	{
		/* frustum_grid.h:221-241 (buildGrid) / :385-398 (buildSphericalGrid): host min/max of projCoordZ, then
		 * grid_kernel.cu:334-352 SlabKernel; DSFillkernel clamps the slab again (:300-311) */
		float zMin = mode ? 9999.9f : +2.0f, zMax = mode ? -9999.9f : -2.0f;
		for (int i = 0; i < T; i++) {
			if (zMin > projZ[i] && projZ[i] >= 0.0f) zMin = projZ[i];
			if (zMax < projZ[i]) zMax = projZ[i];
		}
		for (int t = 0; t < T; t++) {
			float pCoord = projZ[t];
			if (pCoord >= 0.0f) {
				unsigned binID = (unsigned)(long long)(S * (float)((pCoord - zMin) / (zMax - zMin))); /* cvttss2si r64, low half */
				if (binID >= (unsigned)S) binID = (unsigned)S - 1;
				slab[t] = binID;
			}
			slab[t] = (unsigned)clampi((int)slab[t], 0, S - 1);
		}
	}
	memset(span, 0, sizeof(unsigned) * C);
	for (int t = 0; t < T; t++)
		for (int gx = rect[4 * t]; gx <= rect[4 * t + 1]; gx++)
			for (int gy = rect[4 * t + 2]; gy <= rect[4 * t + 3]; gy++) span[(gx * g_nby + gy) * S + (int)slab[t]]++;
	unsigned acc = 0;
	for (int c = 0; c < C; c++) { offset[c] = acc; acc += span[c]; }
	unsigned *cursor = (unsigned *)malloc(sizeof(unsigned) * C);
	memcpy(cursor, offset, sizeof(unsigned) * C);
	for (int t = 0; t < T; t++) /* ascending t == emission order; stability keeps it inside a cell */
		for (int gx = rect[4 * t]; gx <= rect[4 * t + 1]; gx++)
			for (int gy = rect[4 * t + 2]; gy <= rect[4 * t + 3]; gy++) {
				unsigned key = (unsigned)((gx * g_nby + gy) * S) + slab[t], at = cursor[key]++; /* grid_kernel.cu:322 */
				key_list[at] = key;
				value_list[at] = (unsigned)t;
			}
	free(cursor);
	free(rect);
	free(projZ);
	free(slab);
	return P;
}
long orc_build_grid(int *faces, float *verts, int T, unsigned *span, unsigned *offset, unsigned *key_list, unsigned *value_list, long cap)
{
	return build_grid(0, 0, 0, faces, verts, T, span, offset, key_list, value_list, cap);
}
/* frustum_grid.h:368-532 buildSphericalGrid: the same sequence around the spherical count / fill kernels */
long orc_build_grid_spherical(int *faces, float *verts, int T, unsigned *span, unsigned *offset, unsigned *key_list, unsigned *value_list, long cap,
                              float xM, float yM)
{
	return build_grid(1, xM, yM, faces, verts, T, span, offset, key_list, value_list, cap);
}

/* ------------------------------------------------------------------ a9: primary ray through a pixel
 * trace_kernel.cu:96-114 with the software texture fetch (CUDA Programming Guide linear filtering:
 * xB = s*N - 0.5, i = floor(xB), alpha = frac(xB), clamp addressing; exact float weights unless q8). */
static void lattice_fetch(float s, float t, float *out)
{
	const int N = 5;
	float xB = s * N - 0.5f, yB = t * N - 0.5f;
	float fi = floorf(xB), fj = floorf(yB);
	float a = xB - fi, b = yB - fj;
	if (g_q8) { a = floorf(a * 256.0f + 0.5f) / 256.0f; b = floorf(b * 256.0f + 0.5f) / 256.0f; }
	int i0 = clampi((int)fi, 0, N - 1), i1 = clampi((int)fi + 1, 0, N - 1);
	int j0 = clampi((int)fj, 0, N - 1), j1 = clampi((int)fj + 1, 0, N - 1);
	float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
	for (int k = 0; k < 3; k++)
		out[k] = w00 * g_lattice[j0 * N + i0][k] + w10 * g_lattice[j0 * N + i1][k] + w01 * g_lattice[j1 * N + i0][k] +
		         w11 * g_lattice[j1 * N + i1][k];
}
static void primary_dir(int px, int py, float *d)
{
	float ftx = (float)px / (float)(g_nbx * TILE);
	float fty = (float)py / (float)(g_nby * TILE);
	ftx = 1 - ftx;
	ftx = (float)(ftx * (0.9 - 0.1) + 0.1); /* double, trace_kernel.cu:102-103 */
	fty = (float)(fty * (0.9 - 0.1) + 0.1);
	lattice_fetch(ftx, fty, d);
	d[0] -= g_cc[0]; d[1] -= g_cc[1]; d[2] -= g_cc[2];
	normalize3(d);
}

/* trace_kernel.cu:153-174 / light_kernel.cu:128-148: (origin - v0, v1 - v0, v2 - v0) */
static void stage_triangle(const float *verts, const int *tris, int tri, const float *origin, float *tv, float *e1, float *e2)
{
	const float *v0 = verts + 3 * (size_t)tris[3 * tri], *v1 = verts + 3 * (size_t)tris[3 * tri + 1], *v2 = verts + 3 * (size_t)tris[3 * tri + 2];
	for (int k = 0; k < 3; k++) { e1[k] = v1[k] - v0[k]; e2[k] = v2[k] - v0[k]; tv[k] = origin[k] - v0[k]; }
}

/* Moeller-Trumbore core shared by trace_kernel.cu:4-45 and light_kernel.cu:13-50; returns 0 when the
 * ray misses, else 1 with the signed t. */
static int moller_trumbore(const float *tv, const float *e1, const float *e2, const float *dir, float *t_out)
{
	float pvec[3], qvec[3];
	cross3(pvec, dir, e2);
	float det = dot3(e1, pvec);
	if (det > -1e-21 && det < 1e-21) return 0; /* EPSILON, main.cu.h:42 (double literal) */
	float inv_det = (float)(1.0 / det);
	float u = dot3(tv, pvec) * inv_det;
	if (u < 0.0 || u > 1.0) return 0;
	cross3(qvec, tv, e1);
	float v = dot3(dir, qvec) * inv_det;
	if (v < 0.0 || u + v > 1.0) return 0;
	*t_out = dot3(e2, qvec) * inv_det;
	return 1;
}

/* trace_kernel.cu:84-270 rckernel_alpha, per pixel (a11) */
void orc_trace(unsigned *value_list, unsigned *span, unsigned *offset, float *normal, float *t_out, float *dir_out,
               int *shadowed, int *id, float *verts, int *tris)
{
	const int S = g_slabs;
#pragma omp parallel for schedule(dynamic, 16) num_threads(g_threads)
	for (int cell = 0; cell < g_nbx * g_nby; cell++) {
		const int bx = cell / g_nby, by = cell % g_nby; /* cell = gx*NBY + gy, trace_kernel.cu:138 */
		/* per-ray state of the tile's 64 threads (trace_kernel.cu:86,124-130): rayDone 0 = nothing yet, 1 = a hit whose depth
		 * slab is not the one being walked, 2 = done */
		float d[TILE * TILE][3], best[TILE * TILE], be1[TILE * TILE][3], be2[TILE * TILE][3];
		int hit[TILE * TILE], rayDone[TILE * TILE];
		for (int sub = 0; sub < TILE * TILE; sub++) {
			primary_dir(bx * TILE + sub % TILE, by * TILE + sub / TILE, d[sub]);
			best[sub] = 99999999.9f; hit[sub] = -1; rayDone[sub] = 0;
			memset(be1[sub], 0, sizeof(be1[sub])); memset(be2[sub], 0, sizeof(be2[sub]));
		}
		int beamDone = 0; /* metadata[3] */
		for (int slab = 0; slab < S; slab++) { /* :132 */
			if (beamDone) continue;            /* :134 */
			const unsigned n = span[cell * S + slab], off = offset[cell * S + slab]; /* :138-139 */
			int all2 = 1;
			for (int sub = 0; sub < TILE * TILE; sub++) {
				if (rayDone[sub] != 2) /* :179 */
					for (unsigned k = 0; k < n; k++) {
						float tv[3], e1[3], e2[3], t;
						stage_triangle(verts, tris, (int)value_list[off + k], g_cc, tv, e1, e2);
						if (!moller_trumbore(tv, e1, e2, d[sub], &t)) continue;
						if (t < 0) t *= -1;                /* trace_kernel.cu:35: |t| */
						if (t < best[sub] && t > 0) {      /* :38 strict: first in list order wins ties */
							best[sub] = t; hit[sub] = (int)(off + k); rayDone[sub] = 1;
							memcpy(be1[sub], e1, sizeof(e1)); memcpy(be2[sub], e2, sizeof(e2));
						}
					}
				/* isWithin, trace_kernel.cu:56-82: a ray that was done (2) or has nothing (0) becomes 0; a ray with a hit is
				 * done iff floor(ndc_z * NUM_SLABS) is the slab being walked */
				if (rayDone[sub] == 0 || rayDone[sub] == 2) rayDone[sub] = 0;
				else {
					float p[4] = {g_cc[0] + best[sub] * d[sub][0], g_cc[1] + best[sub] * d[sub][1], g_cc[2] + best[sub] * d[sub][2], 1.0f}, q[4];
					mat4_mul_vec4(q, g_cc + 48, p);
					q[2] /= q[3];
					rayDone[sub] = (f2i(floorf(q[2] * S)) == slab) ? 2 : 1;
				}
				if (rayDone[sub] != 2) all2 = 0;
			}
			beamDone = all2; /* :219-228: the tile stops walking slabs once all 64 rays are done */
		}
		for (int sub = 0; sub < TILE * TILE; sub++) {
			const int px = bx * TILE + sub % TILE, py = by * TILE + sub / TILE, pix = py * g_W + px;
			float *nrm = normal + 3 * (size_t)pix;
			if (rayDone[sub] == 2) { /* :232-253 */
				float nn[3];
				normalize3(be1[sub]); normalize3(be2[sub]);
				cross3(nn, be1[sub], be2[sub]);
				normalize3(nn);
				for (int k = 0; k < 3; k++) nrm[k] = (nn[k] < 0) ? nn[k] * -1 : nn[k];
				t_out[pix] = best[sub]; id[pix] = (int)value_list[hit[sub]];
			} else {    /* :255-263 */
				nrm[0] = nrm[1] = nrm[2] = -1.0f;
				t_out[pix] = -1.0f; id[pix] = -2;
			}
			shadowed[pix] = 0;
			memcpy(dir_out + 3 * (size_t)pix, d[sub], sizeof(d[sub]));
		}
	}
}

/* ------------------------------------------------------------------ a13: shadow-ray -> light cell
 * misc_kernel.cu:63-129 mapSortkernel (g_cc holds the light camera; cmPt is the primary camera eye) */
void orc_map_rays(float *t_list, float *dir_list, unsigned *map, float *cmPt, float nearPlane)
{
	const int N = g_W * g_H;
	const float tanHalf = tanf(g_fovy * M_PI / 360.0f); /* misc_kernel.cu:115: tanf of a double constant */
#pragma omp parallel for num_threads(g_threads)
	for (int pix = 0; pix < N; pix++) {
		float tv = t_list[pix], P[3], L[3], pl[4];
		for (int k = 0; k < 3; k++) { P[k] = cmPt[k] + tv * dir_list[3 * (size_t)pix + k]; L[k] = P[k] - g_cc[k]; }
		normalize3(L);
		for (int k = 0; k < 4; k++) pl[k] = g_cc[48 + 4 * k + 3] + g_cc[48 + 4 * k + 2]; /* un-normalised near plane */
		float s = -1.0f * (dot3(pl, g_cc) + pl[3]) / (dot3(pl, L));
		float x = s * L[0], y = s * L[1], z = s * L[2];
		float right = x * g_cc[16 + 0] + y * g_cc[16 + 4] + z * g_cc[16 + 8];
		float upd = x * g_cc[16 + 1] + y * g_cc[16 + 5] + z * g_cc[16 + 9];
		/* :115-116: float numerator / double denominator, times the int cell count, floor in double */
		int blx = (int)floor(((right + nearPlane * tanHalf) / (2.0 * nearPlane * tanHalf)) * g_nbx);
		int bly = (int)floor(((upd + nearPlane * tanHalf) / (2.0 * nearPlane * tanHalf)) * g_nby);
		unsigned cell = (unsigned)(g_nbx * g_nby);
		if (blx >= 0 && blx <= g_nbx - 1 && bly >= 0 && bly <= g_nby - 1) cell = (unsigned)(blx * g_nby + bly);
		map[pix] = (unsigned)pix;
		map[N + pix] = cell;
	}
}

/* libm's acosf over an array: what the host-compiled reference gets from glibc; tests pin the product's restatement on it */
void orc_acosf_array(const float *x, float *y, long n)
{
	for (long i = 0; i < n; i++) y[i] = acosf(x[i]);
}

/* a14: misc_kernel.cu:255-296 mapSort_Effective_kernel - the map display() really calls (main.cu:189) */
void orc_map_rays_spherical(float *t_list, float *dir_list, unsigned *map, float *cmPt, float xM, float yM)
{
	const int N = g_W * g_H;
#pragma omp parallel for num_threads(g_threads)
	for (int pix = 0; pix < N; pix++) {
		float tv = t_list[pix], P[3], L[3];
		for (int k = 0; k < 3; k++) { P[k] = cmPt[k] + tv * dir_list[3 * (size_t)pix + k]; L[k] = P[k] - g_cc[k]; }
		normalize3(L);
		int blx = effective_x(L, xM), bly = effective_y(L, yM);
		unsigned cell = (unsigned)(g_nbx * g_nby);
		if (blx >= 0 && blx < g_nbx && bly >= 0 && bly < g_nby) cell = (unsigned)(blx * g_nby + bly);
		map[pix] = (unsigned)pix;
		map[N + pix] = cell;
	}
}

/* ------------------------------------------------------------------ a17/a18: shadow any-hit
 * light_kernel.cu:52-270 per pixel; light_kernel.cu:1-11 isSmaller; sentinel cell => unshadowed. */
void orc_shadow(unsigned *value_list, float *verts, int *tris, unsigned *span, unsigned *offset, float *t_list,
                float *dir_list, int *is_shadowed, unsigned *map, float *cmPt)
{
	const int N = g_W * g_H;
#pragma omp parallel for schedule(dynamic, 256) num_threads(g_threads)
	for (int i = 0; i < N; i++) {
		const unsigned cell = map[N + i];
		if (cell >= (unsigned)(g_nbx * g_nby)) continue;
		const int pix = (int)map[i];
		float tv = t_list[pix], P[3], L[3];
		for (int k = 0; k < 3; k++) { P[k] = cmPt[k] + tv * dir_list[3 * (size_t)pix + k]; L[k] = P[k] - g_cc[k]; }
		normalize3(L);
		const float dx = P[0] - g_cc[0], dy = P[1] - g_cc[1], dz = P[2] - g_cc[2];
		const float dist_b = sqrtf(dx * dx + dy * dy + dz * dz);
		/* light_kernel.cu:105-115: the cell's lists of all depth slabs, one after the other (an OR: the order is immaterial);
		 * sorted keys make them one contiguous range */
		const unsigned l_off = offset[cell * g_slabs], l_end = offset[cell * g_slabs + g_slabs - 1] + span[cell * g_slabs + g_slabs - 1];
		for (unsigned k = 0; k < l_end - l_off; k++) {
			float o[3], e1[3], e2[3], t;
			stage_triangle(verts, tris, (int)value_list[l_off + k], g_cc, o, e1, e2);
			if (!moller_trumbore(o, e1, e2, L, &t)) continue;
			if (!(t < 999999.9f) || t == 0) continue; /* light_kernel.cu:45-47 + `if (value)` :191 */
			float q[3] = {g_cc[0] + t * L[0], g_cc[1] + t * L[1], g_cc[2] + t * L[2]};
			const float ax = q[0] - g_cc[0], ay = q[1] - g_cc[1], az = q[2] - g_cc[2];
			const float dist_a = sqrtf(ax * ax + ay * ay + az * az);
			if (dist_a + 1e-03f < dist_b) { is_shadowed[pix] = 1; break; }
		}
	}
}

/* ------------------------------------------------------------------ a19/a20: shading
 * shader_kernel.cu:165-221 lambertian_shade + :46-86 lambert_color_pixel.  Miss pixels (id < 0) are
 * black (the reference reads mat_idx[-2] there - SURVEY A.3); `id` is overwritten with the material
 * index as the reference does (:176). */
/* shader_kernel.cu:223-273 get_along_x / get_along_y: the signed versions of getEffective_x/y's angles */
static float along_x(const float *vec)
{
	const float *cc = g_cc;
	float upDot = vec[0] * cc[16 + 1] + vec[1] * cc[16 + 5] + vec[2] * cc[16 + 9];
	float tmp[3] = {vec[0] - upDot * cc[16 + 1], vec[1] - upDot * cc[16 + 5], vec[2] - upDot * cc[16 + 9]};
	float val = sqrtf(tmp[0] * tmp[0] + tmp[1] * tmp[1] + tmp[2] * tmp[2]);
	tmp[0] /= val; tmp[1] /= val; tmp[2] /= val;
	float fwd = tmp[0] * cc[16 + 2] + tmp[1] * cc[16 + 6] + tmp[2] * cc[16 + 10];
	float angle = acosf(fwd);
	float rightDot = tmp[0] * cc[16 + 0] + tmp[1] * cc[16 + 4] + tmp[2] * cc[16 + 8];
	return rightDot > 0 ? angle : (float)(-1.0 * angle);
}
static float along_y(const float *vec)
{
	const float *cc = g_cc;
	float rightDot = vec[0] * cc[16 + 0] + vec[1] * cc[16 + 4] + vec[2] * cc[16 + 8];
	float tmp[3] = {vec[0] - rightDot * cc[16 + 0], vec[1] - rightDot * cc[16 + 4], vec[2] - rightDot * cc[16 + 8]};
	float val = sqrtf(tmp[0] * tmp[0] + tmp[1] * tmp[1] + tmp[2] * tmp[2]);
	tmp[0] /= val; tmp[1] /= val; tmp[2] /= val;
	float upDot = tmp[0] * cc[16 + 1] + tmp[1] * cc[16 + 5] + tmp[2] * cc[16 + 9];
	float fwd = tmp[0] * cc[16 + 2] + tmp[1] * cc[16 + 6] * tmp[2] * cc[16 + 10]; /* :262, as written */
	float angle = acosf(fwd);
	return upDot > 0 ? angle : (float)(-1.0 * angle);
}

static void shade(int spot, unsigned char *img, float *normal, float *t, float *dir, int *id, float *cam_pos, int *mat_idx, float *mat_list, int mat_count,
                  float *dump);
void orc_shade(unsigned char *img, float *normal, float *t, float *dir, int *id, float *cam_pos, int *mat_idx,
               float *mat_list, int mat_count)
{
	shade(0, img, normal, t, dir, id, cam_pos, mat_idx, mat_list, mat_count, NULL);
}
/* shader_kernel.cu:275-345 spot_shade + :88-128 lambert_color_drop_off_pixel (Shader::spotlight_shade, shader.h:88-113 - what
 * display() shades with from its third frame on, main.cu:205-219).  g_cc holds the resident (last light's) camera: the
 * cone is about ITS axes, from ITS eye.  `dump` (2N floats, may be NULL) receives the two angles like the reference's
 * debug buffer; miss pixels get NaN-free zeros here because the reference computes them from t = -1 garbage. */
void orc_spot_shade(unsigned char *img, float *normal, float *t, float *dir, int *id, float *cam_pos, int *mat_idx,
                    float *mat_list, int mat_count, float *dump)
{
	shade(1, img, normal, t, dir, id, cam_pos, mat_idx, mat_list, mat_count, dump);
}
static void shade(int spot, unsigned char *img, float *normal, float *t, float *dir, int *id, float *cam_pos, int *mat_idx, float *mat_list, int mat_count,
                  float *dump)
{
	const int N = g_W * g_H;
	const float *R = g_cc + 16;
#pragma omp parallel for num_threads(g_threads)
	for (int pix = 0; pix < N; pix++) {
		float color[3] = {0.0f, 0.0f, 0.0f};
		const int tri = id[pix], idx = (tri >= 0) ? mat_idx[tri] : -1;
		id[pix] = idx;
		if (idx >= 0 && idx < mat_count && t[pix] > 0) {
			const float *kd = mat_list + 6 * idx + 3; /* ambient colour := diffuse entries, :180-186 */
			float P[3], Lv[3], Pv[3], Nv[3], ld[3];
			const float *n = normal + 3 * (size_t)pix;
			for (int k = 0; k < 3; k++) P[k] = cam_pos[k] + t[pix] * dir[3 * (size_t)pix + k];
			float drop = 1.0f;
			if (spot) { /* :283-301 */
				float L[3] = {P[0] - g_cc[0], P[1] - g_cc[1], P[2] - g_cc[2]};
				normalize3(L);
				const float x = along_x(L), y = along_y(L);
				if (dump) { dump[2 * (size_t)pix] = x; dump[2 * (size_t)pix + 1] = y; }
				drop = (x < M_PI / 4 && x > -M_PI / 4 && y < M_PI / 4 && y > -M_PI / 4) ? 1.0f : 0.25f;
			}
			for (int k = 0; k < 3; k++) { /* rotation part of the resident modelview, :54-67 */
				Lv[k] = R[k] * g_light[0] + R[4 + k] * g_light[1] + R[8 + k] * g_light[2];
				Pv[k] = R[k] * P[0] + R[4 + k] * P[1] + R[8 + k] * P[2];
				Nv[k] = R[k] * n[0] + R[4 + k] * n[1] + R[8 + k] * n[2];
			}
			normalize3(Nv);
			for (int k = 0; k < 3; k++) ld[k] = Pv[k] - Lv[k];
			normalize3(ld);
			for (int k = 0; k < 3; k++) color[k] += spot ? kd[k] * 0.5f * drop : kd[k] * 0.5f;
			float dd = dot3(ld, Nv);
			if (!(dd > 0)) dd *= -1;
			if (dd > 0)
				for (int k = 0; k < 3; k++) color[k] += spot ? kd[k] * 1.0f * dd * drop : kd[k] * 1.0f * dd;
			for (int k = 0; k < 3; k++) if (color[k] > 1.0f) color[k] = 1.0f;
		}
		for (int k = 0; k < 3; k++) img[3 * (size_t)pix + k] = (unsigned char)(color[k] * 255);
	}
}
/* shader_kernel.cu:347-359 */
void orc_add_shadows(unsigned char *img, int *is_shadowed)
{
	const int N = g_W * g_H;
	for (int pix = 0; pix < N; pix++)
		if (is_shadowed[pix] == 1)
			for (int k = 0; k < 3; k++) img[3 * (size_t)pix + k] /= 3;
}

/* ------------------------------------------------------------------ row R: reflection (AUTHORED SPEC, parity unpinned)
 * The reference has no reflection / uniform-grid / DDA code, so this is the specification the CUDA
 * path is tested against, written with the reference's own arithmetic (intersectTri, light_kernel.cu:13-50, and
 * lambert_color_pixel, shader_kernel.cu:46-86).  It is grid independent: the nearest hit is the minimum over ALL
 * triangles, ties to the lower triangle id, so any conservative acceleration structure must reproduce it.
 *
 *   local0 = the shaded + shadowed RGB8 of the primary hit, as floats;  k0 = reflectivity of its material
 *   if bounces == 0, the pixel is a miss, or !(k0 > 0): the pixel is left as it is
 *   acc = (1-k0)*local0;  w = k0;  ray = reflect(P0, d0, tri0)
 *   for b = 1..bounces:  hit = nearest(ray) or stop;  local = RGB8 of the lambert shade of the hit (no shadow test);
 *        k = (b < bounces) ? reflectivity(hit) : 0;   if !(k > 0) { acc += w*local; stop }
 *        acc += (w*(1-k))*local;  w *= k;  ray = reflect(P, r, tri)
 *   pixel = (uchar)acc
 *   reflect(P, d, tri): n = normalize(cross(e1,e2)), flipped to face the incoming ray (dot(n,d) > 0 => n = -n);
 *                       origin O = P + 1e-3*n;  direction r = d - (2*dot(d,n))*n  (not renormalised)
 *   nearest(O, r): triangles in ascending id, tvec = O - v0, Moeller-Trumbore as intersectTri; accept signed t with
 *                  t > 0 and t < best (best starts at 99999999.9f, strict '<': the lower id wins ties) */
static void lambert_u8(const float *P, const float *v0, const float *v1, const float *v2, const float *kd, float *out)
{
	const float *R = g_cc + 16;
	float e1[3], e2[3], n[3], Lv[3], Pv[3], Nv[3], ld[3], color[3] = {0.0f, 0.0f, 0.0f};
	for (int k = 0; k < 3; k++) { e1[k] = v1[k] - v0[k]; e2[k] = v2[k] - v0[k]; }
	normalize3(e1); normalize3(e2);
	cross3(n, e1, e2);
	normalize3(n);
	for (int k = 0; k < 3; k++) n[k] = (n[k] < 0) ? n[k] * -1 : n[k]; /* trace_kernel.cu:241-243 */
	for (int k = 0; k < 3; k++) {
		Lv[k] = R[k] * g_light[0] + R[4 + k] * g_light[1] + R[8 + k] * g_light[2];
		Pv[k] = R[k] * P[0] + R[4 + k] * P[1] + R[8 + k] * P[2];
		Nv[k] = R[k] * n[0] + R[4 + k] * n[1] + R[8 + k] * n[2];
	}
	normalize3(Nv);
	for (int k = 0; k < 3; k++) ld[k] = Pv[k] - Lv[k];
	normalize3(ld);
	for (int k = 0; k < 3; k++) color[k] += kd[k] * 0.5f;
	float dd = dot3(ld, Nv);
	if (!(dd > 0)) dd *= -1;
	if (dd > 0)
		for (int k = 0; k < 3; k++) color[k] += kd[k] * 1.0f * dd;
	for (int k = 0; k < 3; k++) {
		if (color[k] > 1.0f) color[k] = 1.0f;
		out[k] = (float)(unsigned char)(color[k] * 255);
	}
}

static void reflect_ray(const float *P, const float *d, const float *v0, const float *v1, const float *v2, float *O, float *r)
{
	float e1[3], e2[3], n[3];
	for (int k = 0; k < 3; k++) { e1[k] = v1[k] - v0[k]; e2[k] = v2[k] - v0[k]; }
	cross3(n, e1, e2);
	normalize3(n);
	if (dot3(n, d) > 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
	const float s = 2.0f * dot3(d, n);
	for (int k = 0; k < 3; k++) { O[k] = P[k] + 1e-3f * n[k]; r[k] = d[k] - s * n[k]; }
}

void orc_reflect(float *verts, int *tris, int T, float *t, float *dir, int *id, unsigned char *rgb, float *cam_pos, int *mat_idx, float *mat_list,
                 int mat_count, float *mat_reflect, int bounces)
{
	const int N = g_W * g_H;
	if (bounces <= 0) return;
#pragma omp parallel for schedule(dynamic, 64) num_threads(g_threads)
	for (int pix = 0; pix < N; pix++) {
		int tri = id[pix];
		if (tri < 0 || !(t[pix] > 0)) continue;
		int m = mat_idx[tri];
		float k = (m >= 0 && m < mat_count) ? mat_reflect[m] : 0.0f;
		if (!(k > 0)) continue;
		float P[3], d[3], O[3], r[3], acc[3], w = k;
		for (int c = 0; c < 3; c++) {
			d[c] = dir[3 * (size_t)pix + c];
			P[c] = cam_pos[c] + t[pix] * d[c];
			acc[c] = (1.0f - k) * (float)rgb[3 * (size_t)pix + c];
		}
		const float *v0 = verts + 3 * (size_t)tris[3 * tri], *v1 = verts + 3 * (size_t)tris[3 * tri + 1], *v2 = verts + 3 * (size_t)tris[3 * tri + 2];
		reflect_ray(P, d, v0, v1, v2, O, r);
		for (int b = 1; b <= bounces; b++) {
			float best = 99999999.9f;
			int hit = -1;
			for (int q = 0; q < T; q++) {
				float tv[3], e1[3], e2[3], tt;
				stage_triangle(verts, tris, q, O, tv, e1, e2);
				if (!moller_trumbore(tv, e1, e2, r, &tt)) continue;
				if (tt > 0 && tt < best) { best = tt; hit = q; }
			}
			if (hit < 0) break;
			float local[3];
			for (int c = 0; c < 3; c++) { P[c] = O[c] + best * r[c]; d[c] = r[c]; }
			v0 = verts + 3 * (size_t)tris[3 * hit]; v1 = verts + 3 * (size_t)tris[3 * hit + 1]; v2 = verts + 3 * (size_t)tris[3 * hit + 2];
			m = mat_idx[hit];
			if (m >= 0 && m < mat_count) lambert_u8(P, v0, v1, v2, mat_list + 6 * m + 3, local);
			else local[0] = local[1] = local[2] = 0.0f;
			k = (b < bounces && m >= 0 && m < mat_count) ? mat_reflect[m] : 0.0f;
			if (!(k > 0)) {
				for (int c = 0; c < 3; c++) acc[c] = acc[c] + w * local[c];
				break;
			}
			for (int c = 0; c < 3; c++) acc[c] = acc[c] + (w * (1.0f - k)) * local[c];
			w = w * k;
			reflect_ray(P, d, v0, v1, v2, O, r);
		}
		for (int c = 0; c < 3; c++) rgb[3 * (size_t)pix + c] = (unsigned char)acc[c];
	}
}

/* ------------------------------------------------------------------ a22: loader semantics
 * obj_parser/obj_parser.cpp:296-421 (OBJ), :177-294 (MTL), obj_parser/list.cpp:115-126 (prefix
 * match on usemtl), scene.h:225-331 (flatten: first three indices of each face, 0-based). */
typedef struct { char (*name)[256]; int n, cap; } name_list;
static void read_mtl_names(const char *file, name_list *L)
{
	FILE *fp = fopen(file, "r");
	char line[500];
	if (!fp) { fprintf(stderr, "Error reading file: %s\n", file); return; }
	while (fgets(line, sizeof(line), fp)) {
		char *tok = strtok(line, " \t\n\r");
		if (!tok || strcmp(tok, "newmtl") != 0) continue;
		char *nm = strtok(NULL, " \t"); /* keeps the trailing newline, obj_parser.cpp:216 */
		if (!nm) continue;
		if (L->n == L->cap) { L->cap = L->cap ? 2 * L->cap : 16; L->name = realloc(L->name, (size_t)L->cap * 256); }
		strncpy(L->name[L->n], nm, 255);
		L->name[L->n][255] = 0;
		L->n++;
	}
	fclose(fp);
}
static int find_material(const name_list *L, const char *want)
{
	const size_t n = strlen(want);
	for (int i = 0; i < L->n; i++)
		if (strncmp(L->name[i], want, n) == 0) return i; /* first entry having `want` as a prefix */
	return -1;
}
static int obj_index(int count, int idx) { return idx == 0 ? -1 : (idx < 0 ? count + idx : idx - 1); } /* obj_parser.cpp:16-25 */

int orc_load_obj(char *path, int *num_vertices, int *num_faces, float *vertexlist, int *facelist, int *matidx)
{
	FILE *fp = fopen(path, "r");
	if (!fp) return -1;
	name_list mats = {0, 0, 0};
	char line[500];
	int nv = 0, nf = 0, cur = -1;
	const int fill = vertexlist && facelist && matidx;
	while (fgets(line, sizeof(line), fp)) {
		char *tok = strtok(line, " \t\n\r");
		if (!tok || tok[0] == '#') continue;
		if (strcmp(tok, "v") == 0) {
			double e[3];
			for (int k = 0; k < 3; k++) { char *s = strtok(NULL, " \t\n\r"); e[k] = s ? atof(s) : 0.0; }
			if (fill) for (int k = 0; k < 3; k++) vertexlist[3 * (size_t)nv + k] = (float)e[k];
			nv++;
		} else if (strcmp(tok, "f") == 0) {
			int vi[4] = {0, 0, 0, 0}, n = 0;
			char *s;
			while ((s = strtok(NULL, " \t\n\r")) != NULL) { if (n < 4) vi[n] = atoi(s); n++; }
			if (fill) {
				for (int k = 0; k < 3; k++) facelist[3 * (size_t)nf + k] = obj_index(nv, vi[k]);
				matidx[nf] = cur;
			}
			nf++;
		} else if (strcmp(tok, "usemtl") == 0) {
			char *s = strtok(NULL, " \t\n\r");
			cur = s ? find_material(&mats, s) : -1;
		} else if (strcmp(tok, "mtllib") == 0) {
			char *s = strtok(NULL, " \t\n\r");
			if (s) { char fn[500]; strncpy(fn, s, 499); fn[499] = 0; read_mtl_names(fn, &mats); }
		}
	}
	fclose(fp);
	free(mats.name);
	*num_vertices = nv;
	*num_faces = nf;
	return 0;
}

/* scene.h:370-439 some_material: positional tokens */
int orc_load_materials(char *file, int *num_materials, float *materiallist)
{
	FILE *fp = fopen(file, "r");
	if (!fp) return -1;
	char w[512];
	int nm = 0;
	while (fscanf(fp, "%511s", w) != EOF)
		if (strcmp(w, "newmtl") == 0) nm++;
	*num_materials = nm;
	if (materiallist) {
		rewind(fp);
		for (int m = 0; m < nm; m++) {
			float *o = materiallist + 6 * m;
			if (fscanf(fp, "%511s %511s %511s %f %f %f %511s %f %f %f", w, w, w, o, o + 1, o + 2, w, o + 3, o + 4, o + 5) != 10) break;
			for (int i = 0; i < 12; i++)
				if (fscanf(fp, "%511s", w) != 1) break;
		}
	}
	fclose(fp);
	return 0;
}
