/* camera_host.cpp -- host-side mirror of camera.h (`class Camera`) and of the host halves of
 * per_frame_funcs.h (fillCoordinatesData :18-43, setDirectionTexture :161-434).  Same field names as
 * the reference class; OpenGL is replaced by in-house matrix code that follows libGLU's
 * gluPerspective (double math, float storage) and gluLookAt (float math) and a float matrix stack.
 * This is per-frame host setup in the reference too (a few hundred flops), not device work.
 * Built with -ffp-contract=off: the values must be reproducible bit for bit.
 */
#include "ugrt_internal.h"
#include <math.h>
#include <string.h>

namespace {

struct Mat4 { float m[16]; }; /* column-major like OpenGL */

Mat4 identity()
{
	Mat4 r;
	for (int i = 0; i < 16; i++) r.m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
	return r;
}

/* glMultMatrixf: current = current * m */
Mat4 mul(const Mat4 &c, const Mat4 &m)
{
	Mat4 r;
	for (int col = 0; col < 4; col++)
		for (int row = 0; row < 4; row++) {
			float acc = c.m[row] * m.m[col * 4];
			acc = acc + c.m[4 + row] * m.m[col * 4 + 1];
			acc = acc + c.m[8 + row] * m.m[col * 4 + 2];
			acc = acc + c.m[12 + row] * m.m[col * 4 + 3];
			r.m[col * 4 + row] = acc;
		}
	return r;
}

float len3(const float *v) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

void cross(float *d, const float *a, const float *b)
{
	d[0] = a[1] * b[2] - a[2] * b[1];
	d[1] = a[2] * b[0] - a[0] * b[2];
	d[2] = a[0] * b[1] - a[1] * b[0];
}

/* camera.h:218-239 Intersect3Planes */
void intersect3(float *p, const float *a, const float *b, const float *c)
{
	float ab[3], bc[3], ca[3];
	ab[0] = a[1] * b[2] - b[1] * a[2]; ab[1] = a[2] * b[0] - a[0] * b[2]; ab[2] = a[0] * b[1] - b[0] * a[1];
	bc[0] = b[1] * c[2] - c[1] * b[2]; bc[1] = b[2] * c[0] - b[0] * c[2]; bc[2] = b[0] * c[1] - c[0] * b[1];
	ca[0] = c[1] * a[2] - a[1] * c[2]; ca[1] = c[2] * a[0] - c[0] * a[2]; ca[2] = c[0] * a[1] - a[0] * c[1];
	const float den = a[0] * bc[0] + a[1] * bc[1] + a[2] * bc[2];
	for (int k = 0; k < 3; k++) p[k] = -(a[3] * bc[k] + b[3] * ca[k] + c[3] * ab[k]) / den;
}

} // namespace

extern "C" {

void ug_camera_setCameraCenter(ug_camera *c, float x, float y, float z) { c->cx = x; c->cy = y; c->cz = z; }
void ug_camera_setCameraLookAt(ug_camera *c, float x, float y, float z) { c->lx = x; c->ly = y; c->lz = z; }
void ug_camera_setCameraUp(ug_camera *c, float x, float y, float z) { c->ux = x; c->uy = y; c->uz = z; }
void ug_camera_setNearFar(ug_camera *c, float nearP, float farP) { c->nearPlane = nearP; c->farPlane = farP; }

void ug_camera_adjustCameraAndPosition(ug_camera *c)
{
	/* gluPerspective(FOVY, aspect, near, far) on an identity GL_PROJECTION (main.cu:74-76, camera.h:137) */
	const double fovy = c->fovy, aspect = c->aspect, zn = c->nearPlane, zf = c->farPlane;
	const double half = fovy / 2 * 3.14159265358979323846 / 180;
	const double cot = cos(half) / sin(half), dz = zf - zn;
	Mat4 P;
	memset(&P, 0, sizeof(P));
	P.m[0] = (float)(cot / aspect);
	P.m[5] = (float)cot;
	P.m[10] = (float)(-(zf + zn) / dz);
	P.m[11] = -1.0f;
	P.m[14] = (float)(-2 * zn * zf / dz);
	P = mul(identity(), P);

	/* gluLookAt on an identity GL_MODELVIEW (camera.h:138-141) */
	float f[3] = {(float)((double)c->lx - (double)c->cx), (float)((double)c->ly - (double)c->cy), (float)((double)c->lz - (double)c->cz)};
	float up[3] = {c->ux, c->uy, c->uz}, s[3], u[3];
	float r = len3(f);
	if (r != 0.0f) { f[0] /= r; f[1] /= r; f[2] /= r; }
	cross(s, f, up);
	r = len3(s);
	if (r != 0.0f) { s[0] /= r; s[1] /= r; s[2] /= r; }
	cross(u, s, f);
	Mat4 R = identity(), Tr = identity();
	R.m[0] = s[0]; R.m[4] = s[1]; R.m[8] = s[2];
	R.m[1] = u[0]; R.m[5] = u[1]; R.m[9] = u[2];
	R.m[2] = -f[0]; R.m[6] = -f[1]; R.m[10] = -f[2];
	Tr.m[12] = -c->cx; Tr.m[13] = -c->cy; Tr.m[14] = -c->cz;
	Mat4 MV = mul(mul(identity(), R), Tr);

	c->worldori[0] = c->cx; c->worldori[1] = c->cy; c->worldori[2] = c->cz; c->worldori[3] = 1.0f; /* camera.h:142-145 */
	memcpy(c->modelview_matrix, MV.m, sizeof(MV.m));
	memcpy(c->projection_matrix, P.m, sizeof(P.m));

	/* camera.h:150-165 getMVPMatrix */
	for (int i = 0; i < 4; i++)
		for (int k = 0; k < 4; k++) {
			float acc = 0;
			for (int j = 0; j < 4; j++) acc += c->modelview_matrix[i * 4 + j] * c->projection_matrix[j * 4 + k];
			c->mvp_matrix[i * 4 + k] = acc;
		}

	/* camera.h:167-216 getFrustumPlanes (Gribb-Hartmann): row3 -/+ row0, +/- row1, +/- row2, normalised */
	const float *m = c->mvp_matrix;
	for (int p = 0; p < 6; p++) {
		const int row = p / 2;
		const bool plus = (p == 1 || p == 2 || p == 4);
		float *e = c->frustum_plane_eq[p];
		for (int k = 0; k < 4; k++) e[k] = plus ? m[4 * k + 3] + m[4 * k + row] : m[4 * k + 3] - m[4 * k + row];
		e[4] = sqrtf(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
		e[5] = 0.0f;
	}
	for (int p = 0; p < 6; p++)
		for (int k = 0; k < 4; k++) c->frustum_plane_eq[p][k] /= c->frustum_plane_eq[p][4];

	/* camera.h:241-253 getFrustumCorners */
	static const int tri[8][3] = {{0, 2, 4}, {1, 2, 4}, {1, 3, 4}, {0, 3, 4}, {0, 2, 5}, {1, 2, 5}, {1, 3, 5}, {0, 3, 5}};
	for (int i = 0; i < 8; i++)
		intersect3(c->frustumcorner[i], c->frustum_plane_eq[tri[i][0]], c->frustum_plane_eq[tri[i][1]], c->frustum_plane_eq[tri[i][2]]);
}

void ug_camera_pack(const ug_camera *c, float cc[64])
{
	for (int i = 0; i < 4; i++) cc[i] = c->worldori[i];
	for (int i = 0; i < 4; i++)
		for (int k = 0; k < 3; k++) cc[4 + i * 3 + k] = c->frustumcorner[i][k];
	memcpy(cc + 16, c->modelview_matrix, 16 * sizeof(float));
	memcpy(cc + 32, c->projection_matrix, 16 * sizeof(float));
	memcpy(cc + 48, c->mvp_matrix, 16 * sizeof(float));
}

void ug_direction_lattice(const float cc[64], float lattice[100])
{
	/* texel(col, row) = mix(mix(corner0, corner1, col/4), mix(corner3, corner2, col/4), row/4).  The
	 * reference evaluates a + w*(b-a) with (b-a) in float and the rest in double, then stores float;
	 * the border rows/columns are plain copies (per_frame_funcs.h:167-417). */
	const float *k0 = cc + 4, *k1 = cc + 7, *k2 = cc + 10, *k3 = cc + 13;
	for (int row = 0; row < 5; row++)
		for (int col = 0; col < 5; col++) {
			const double wc = 0.25 * col, wr = 0.25 * row;
			float *out = lattice + 4 * (row * 5 + col);
			for (int k = 0; k < 3; k++) {
				float bottom, top;
				if (col == 0) { bottom = k0[k]; top = k3[k]; }
				else if (col == 4) { bottom = k1[k]; top = k2[k]; }
				else {
					bottom = (float)(k0[k] + wc * (k1[k] - k0[k]));
					top = (float)(k3[k] + wc * (k2[k] - k3[k]));
				}
				if (row == 0) out[k] = bottom;
				else if (row == 4) out[k] = top;
				else out[k] = (float)(bottom + wr * (top - bottom));
			}
			out[3] = 0.0f;
		}
}

} // extern "C"

void ugh_view_const(ViewConst *vc, const ug_camera *cam, int W, int H, int nbx, int nby, int gy_lo, int gy_cnt, int grid_mode, float xM, float yM, int num_slabs)
{
	vc->grid_mode = grid_mode; vc->xM = xM; vc->yM = yM; vc->num_slabs = num_slabs > 0 ? num_slabs : 1;
	ug_camera_pack(cam, vc->cc);
	ug_direction_lattice(vc->cc, vc->lattice);
	/* misc_kernel.cu:115: tanf(FOVY*M_PI/360.0f) - int*double/float -> double, narrowed to tanf's float argument */
	vc->tan_half = tanf((float)((double)cam->fovy * M_PI / 360.0f));
	vc->near_plane = cam->nearPlane;
	vc->nbx = nbx; vc->nby = nby; vc->gy_lo = gy_lo; vc->gy_cnt = gy_cnt;
	vc->W = W; vc->H = H;
	vc->shards = 1;
	vc->pad_[0] = vc->pad_[1] = vc->pad_[2] = 0;
}
