/* ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (oracle/).  Builds oracle/_ref/libugref.so.
 *
 * Includes the reference's kernel files UNMODIFIED from /root/reference (never copied into this repo)
 * behind ref_shim.h and runs them on the host by index emulation:
 *   - kernels without cross-thread dependence run verbatim, one emulated thread at a time
 *     (copy_data_transform, DSKernel, DSFillkernel, do_scan_dump, set_as_zero, create_histogram,
 *      mapSortkernel, lambertian_shade, shadow_kernel);
 *   - the two block-cooperative kernels (rckernel_alpha trace_kernel.cu:84-270, mod_light_rckernel
 *     light_kernel.cu:52-270) are driven per pixel here, calling the reference's own
 *     intersectTriUV / isWithin / intersectTri / isSmaller;
 *   - host logic containing <<<>>> launches / CUDPP / GL (frustum_grid.h, per_frame_funcs.h,
 *     scene.h) is restated with the file:line it follows; CUDPP scan/compact/sort are replaced by
 *     running sums and std::stable_sort (cudppSort is an LSD radix sort => stable).
 *
 * It is the "reference" the C restatement oracle/ug_oracle.c is pinned against (tests/test_oracle_vs_ref.py)
 * and the `--impl reference` arm of bench.py.  OpenMP parallelises over emulated blocks / pixels.
 */
#include "ref_shim.h"
#include <algorithm>
#include <numeric>
#include <vector>
#include <omp.h>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
int g_ref_W = 1024, g_ref_H = 1024, g_ref_NBX = 128, g_ref_NBY = 128, g_ref_FOVY = 45, g_ref_SLABS = 1;
float dd_camcoords[16 * 4];
float dd_light_position[6];
ref_texture texdir;
int g_ref_tex_q8 = 0;
int sharedmem[1];

/* ---- software texture unit (CUDA Programming Guide, "Texture Fetching": linear filtering) ---- */
float4 tex2D(const ref_texture &tex, float s, float t)
{
	const int N = 5;
	float xB = s * N - 0.5f, yB = t * N - 0.5f;
	float fi = floorf(xB), fj = floorf(yB);
	float a = xB - fi, b = yB - fj;
	if (g_ref_tex_q8) { a = floorf(a * 256.0f + 0.5f) / 256.0f; b = floorf(b * 256.0f + 0.5f) / 256.0f; }
	int i0 = (int)fi, j0 = (int)fj, i1 = i0 + 1, j1 = j0 + 1;
	i0 = std::min(std::max(i0, 0), N - 1); i1 = std::min(std::max(i1, 0), N - 1);
	j0 = std::min(std::max(j0, 0), N - 1); j1 = std::min(std::max(j1, 0), N - 1);
	const float4 &T00 = tex.texel[j0 * N + i0], &T10 = tex.texel[j0 * N + i1];
	const float4 &T01 = tex.texel[j1 * N + i0], &T11 = tex.texel[j1 * N + i1];
	float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
	float4 r;
	r.x = w00 * T00.x + w10 * T10.x + w01 * T01.x + w11 * T11.x;
	r.y = w00 * T00.y + w10 * T10.y + w01 * T01.y + w11 * T11.y;
	r.z = w00 * T00.z + w10 * T10.z + w01 * T01.z + w11 * T11.z;
	r.w = 0.0f;
	return r;
}

/* ---- GL matrix stack stand-in (GL keeps float matrices; GLU computes as libGLU does) ---- */
static float gl_mat[2][16];
static int gl_mode = 0; /* 0 modelview, 1 projection */
static void gl_mult(const float *m)
{
	float *c = gl_mat[gl_mode], r[16];
	for (int col = 0; col < 4; col++)
		for (int row = 0; row < 4; row++)
			r[col * 4 + row] = c[0 * 4 + row] * m[col * 4 + 0] + c[1 * 4 + row] * m[col * 4 + 1] +
			                   c[2 * 4 + row] * m[col * 4 + 2] + c[3 * 4 + row] * m[col * 4 + 3];
	memcpy(c, r, sizeof(r));
}
void glMatrixMode(int mode) { gl_mode = (mode == GL_PROJECTION) ? 1 : 0; }
void glLoadIdentity(void)
{
	float *c = gl_mat[gl_mode];
	for (int i = 0; i < 16; i++) c[i] = (i % 5 == 0) ? 1.0f : 0.0f;
}
void glGetFloatv(int which, float *out) { memcpy(out, gl_mat[which == GL_PROJECTION_MATRIX ? 1 : 0], 16 * sizeof(float)); }
void gluPerspective(double fovy, double aspect, double zNear, double zFar)
{
	double radians = fovy / 2 * 3.14159265358979323846 / 180;
	double deltaZ = zFar - zNear, sine = sin(radians);
	if (deltaZ == 0 || sine == 0 || aspect == 0) return;
	double cotangent = cos(radians) / sine;
	float m[16] = {0};
	m[0] = (float)(cotangent / aspect);
	m[5] = (float)cotangent;
	m[10] = (float)(-(zFar + zNear) / deltaZ);
	m[11] = -1.0f;
	m[14] = (float)(-2 * zNear * zFar / deltaZ);
	gl_mult(m);
}
void gluLookAt(double ex, double ey, double ez, double cx, double cy, double cz, double ux, double uy, double uz)
{
	float f[3] = {(float)(cx - ex), (float)(cy - ey), (float)(cz - ez)}, up[3] = {(float)ux, (float)uy, (float)uz}, s[3], u[3];
	float r = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
	if (r != 0.0f) { f[0] /= r; f[1] /= r; f[2] /= r; }
	s[0] = f[1] * up[2] - f[2] * up[1]; s[1] = f[2] * up[0] - f[0] * up[2]; s[2] = f[0] * up[1] - f[1] * up[0];
	r = sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
	if (r != 0.0f) { s[0] /= r; s[1] /= r; s[2] /= r; }
	u[0] = s[1] * f[2] - s[2] * f[1]; u[1] = s[2] * f[0] - s[0] * f[2]; u[2] = s[0] * f[1] - s[1] * f[0];
	float m[16] = {s[0], u[0], -f[0], 0, s[1], u[1], -f[1], 0, s[2], u[2], -f[2], 0, 0, 0, 0, 1};
	gl_mult(m);
	float t[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, (float)-ex, (float)-ey, (float)-ez, 1};
	gl_mult(t);
}

/* ---- the reference, verbatim ---- */
#include "grid_kernel.cu"
#include "misc_kernel.cu"
#include "trace_kernel.cu"
#include "light_kernel.cu"
#include "shader_kernel.cu"
#include "transformation_kernel.cu"
#include "camera.h"
#include "obj_parser/objLoader.h"

static int g_threads = 1;

/* emulate <<<grid, block>>>: every (block, thread) runs the kernel body once */
template <class F> static void launch(int gx, int gy, int bx, int by, F body)
{
#pragma omp parallel for collapse(2) schedule(dynamic, 4) num_threads(g_threads)
	for (int Bx = 0; Bx < gx; Bx++)
		for (int By = 0; By < gy; By++) {
			gridDim.x = gx; gridDim.y = gy; gridDim.z = 1;
			blockDim.x = bx; blockDim.y = by; blockDim.z = 1;
			blockIdx.x = Bx; blockIdx.y = By; blockIdx.z = 0;
			for (int ty = 0; ty < by; ty++)
				for (int tx = 0; tx < bx; tx++) {
					threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = 0;
					body();
				}
		}
}

extern "C" {

void ref_configure(int W, int H, int fovy, int tex_q8)
{
	g_ref_W = W; g_ref_H = H; g_ref_NBX = W / NUM_THREADS_X; g_ref_NBY = H / NUM_THREADS_Y;
	g_ref_FOVY = fovy; g_ref_tex_q8 = tex_q8;
}
void ref_set_slabs(int n) { g_ref_SLABS = n > 0 ? n : 1; }
void ref_set_threads(int n) { g_threads = n > 0 ? n : omp_get_max_threads(); }
int ref_get_threads(void) { return g_threads; }

/* per_frame_funcs.h:161-434 setDirectionTexture: 5x5 lattice of the near rectangle; texel (col c,
 * row r) = lerp(lerp(corner0,corner1,c/4), lerp(corner3,corner2,c/4), r/4) with the reference's
 * float/double mix: (b-a) in float, w*(..) and the sum in double, stored as float. */
static void set_direction_texture(const float *cc)
{
	const float *c0 = cc + 4, *c1 = cc + 7, *c2 = cc + 10, *c3 = cc + 13;
	const double w[5] = {0.0, 0.25, 0.5, 0.75, 1.0};
	for (int r = 0; r < 5; r++)
		for (int c = 0; c < 5; c++) {
			float out[3];
			for (int k = 0; k < 3; k++) {
				/* endpoints are copied, not interpolated (per_frame_funcs.h:169-172,193-196 ...) */
				float bot = (c == 0) ? c0[k] : (c == 4) ? c1[k] : (float)(c0[k] + w[c] * (c1[k] - c0[k]));
				float top = (c == 0) ? c3[k] : (c == 4) ? c2[k] : (float)(c3[k] + w[c] * (c2[k] - c3[k]));
				out[k] = (r == 0) ? bot : (r == 4) ? top : (float)(bot + w[r] * (top - bot));
			}
			texdir.texel[r * 5 + c] = float4{out[0], out[1], out[2], 0.0f};
		}
}

/* main.cu:72-125 + per_frame_funcs.h:18-43: pose -> Camera (verbatim class) -> dd_camcoords + texture.
 * square_aspect=1 is used for light cameras: mapSortkernel (misc_kernel.cu:115-116) assumes aspect 1. */
void ref_camera_setup(const float *pose9, float nearP, float farP, int square_aspect, float *out_camcoords)
{
	static Camera *camera = new Camera();
	glMatrixMode(GL_PROJECTION);
	glLoadIdentity();
	camera->setCameraCenter(pose9[0], pose9[1], pose9[2]);
	camera->setCameraLookAt(pose9[3], pose9[4], pose9[5]);
	camera->setCameraUp(pose9[6], pose9[7], pose9[8]);
	camera->setNearFar(nearP, farP);
	int saveH = g_ref_H;
	if (square_aspect) g_ref_H = g_ref_W;
	camera->adjustCameraAndPosition();
	g_ref_H = saveH;
	camera->getGLMatrices();
	camera->getFrustumProperties();
	/* fillCoordinatesData, per_frame_funcs.h:18-39 */
	float *cc = dd_camcoords;
	for (int i = 0; i < 4; i++) cc[i] = camera->worldori[i];
	for (int i = 0; i < 4; i++)
		for (int k = 0; k < 3; k++) cc[i * 3 + k + 4] = camera->frustumcorner[i][k];
	for (int i = 0; i < 16; i++) cc[16 + i] = camera->modelview_matrix[i];
	for (int i = 0; i < 16; i++) cc[32 + i] = camera->projection_matrix[i];
	for (int i = 0; i < 16; i++) cc[48 + i] = camera->mvp_matrix[i];
	set_direction_texture(cc);
	if (out_camcoords) memcpy(out_camcoords, cc, sizeof(dd_camcoords));
}

void ref_set_camcoords(const float *cc)
{
	memcpy(dd_camcoords, cc, sizeof(dd_camcoords));
	set_direction_texture(dd_camcoords);
}
void ref_get_dirtex(float *out100)
{
	for (int i = 0; i < 25; i++) { out100[i * 4 + 0] = texdir.texel[i].x; out100[i * 4 + 1] = texdir.texel[i].y; out100[i * 4 + 2] = texdir.texel[i].z; out100[i * 4 + 3] = texdir.texel[i].w; }
}

/* per_frame_funcs.h:6-16 */
void ref_set_lights(const float *lights, int n)
{
	for (int i = 0; i < 6; i++) dd_light_position[i] = (i < 3 * n) ? lights[i] : 0.0f;
}

/* scene.h:122-139 rotate_bunny -> transformation_kernel.cu:4-18 */
void ref_transform(float *vertexlist, float *orig_list, int size, int offset, float rot)
{
	launch(size / 256 + 1, 1, 256, 1, [&] { copy_data_transform(vertexlist, orig_list, size, offset, rot); });
}

/* frustum_grid.h:213-218: DSKernel */
void ref_grid_count(int *faces, float *verts, int T, unsigned *sizeList, float *projZ)
{
	int mp[1] = {T};
	launch(T / NUMTHREADSDS + 1, 1, NUMTHREADSDS, 1, [&] { DSKernel(sizeList, projZ, faces, verts, mp); });
}

/* frustum_grid.h:210-366 buildGrid.  Returns P (number of (cell,triangle) pairs); fills span/offset
 * (C entries each) and, if cap >= P, key_list/value_list. */
long ref_build_grid(int *faces, float *verts, int T, unsigned *span, unsigned *offset,
                    unsigned *key_list, unsigned *value_list, long cap)
{
	const int C = NUM_BLOCKS_X * NUM_BLOCKS_Y * NUM_SLABS;
	int mp[1] = {T};
	std::vector<unsigned> sizeList(T), scanned(T), zSlabs(T, 0);
	std::vector<float> projZ(T);
	const int nb = T / NUMTHREADSDS + 1;
	launch(nb, 1, NUMTHREADSDS, 1, [&] { DSKernel(sizeList.data(), projZ.data(), faces, verts, mp); });
	/* :221-241 host zMin/zMax (only feeds SlabKernel; inert at NUM_SLABS 1) */
	float zMin = +2.0f, zMax = -2.0f;
	for (int i = 0; i < T; i++) {
		if (zMin > projZ[i] && projZ[i] >= 0.0f) zMin = projZ[i];
		if (zMax < projZ[i]) zMax = projZ[i];
	}
	/* :249 cudppScan inclusive */
	unsigned run = 0;
	for (int i = 0; i < T; i++) { run += sizeList[i]; scanned[i] = run; }
	const long P = T ? (long)scanned[T - 1] : 0;
	if (P > cap) {
		/* caller only wanted the count / must re-call with room */
		return P;
	}
	std::vector<unsigned> keys(P), vals(P);
	launch(nb, 1, NUMTHREADSDS, 1, [&] { SlabKernel(zSlabs.data(), projZ.data(), mp, zMin, zMax); });
	launch(nb, 1, NUMTHREADSDS, 1, [&] { DSFillkernel(keys.data(), vals.data(), scanned.data(), zSlabs.data(), faces, verts, mp); });
	/* :298 cudppSort (stable LSD radix) */
	std::vector<unsigned> perm(P);
	std::iota(perm.begin(), perm.end(), 0u);
	std::stable_sort(perm.begin(), perm.end(), [&](unsigned a, unsigned b) { return keys[a] < keys[b]; });
	for (long i = 0; i < P; i++) { key_list[i] = keys[perm[i]]; value_list[i] = vals[perm[i]]; }
	/* :304-337 do_scan_dump + cudppCompact */
	std::vector<unsigned> pos(P), head(P), compacted;
	launch((int)(P / 256) + 1, 1, 256, 1, [&] { do_scan_dump(key_list, pos.data(), head.data(), (int)P); });
	for (long i = 0; i < P; i++) if (head[i]) compacted.push_back(pos[i]);
	const int ncells = (int)compacted.size();
	/* :340-357 set_as_zero + create_histogram */
	launch(C / 256 + 1, 1, 256, 1, [&] { set_as_zero(span); });
	launch(ncells / 256 + 1, 1, 256, 1, [&] { create_histogram(compacted.data(), ncells, (int)P, key_list, span); });
	/* :361 cudppScan exclusive over ALL cells */
	unsigned acc = 0;
	for (int c = 0; c < C; c++) { offset[c] = acc; acc += span[c]; }
	return P;
}

/* frustum_grid.h:377: DS_spherical_Kernel */
void ref_grid_count_spherical(int *faces, float *verts, int T, unsigned *sizeList, float *projZ, float xM, float yM)
{
	int mp[1] = {T};
	launch(T / NUMTHREADSDS + 1, 1, NUMTHREADSDS, 1, [&] { DS_spherical_Kernel(sizeList, projZ, faces, verts, mp, xM, yM); });
}

/* frustum_grid.h:368-532 buildSphericalGrid: the sequence of buildGrid around the spherical kernels */
long ref_build_grid_spherical(int *faces, float *verts, int T, unsigned *span, unsigned *offset, unsigned *key_list, unsigned *value_list, long cap,
                              float xM, float yM)
{
	const int C = NUM_BLOCKS_X * NUM_BLOCKS_Y * NUM_SLABS;
	int mp[1] = {T};
	std::vector<unsigned> sizeList(T), scanned(T), zSlabs(T, 0);
	std::vector<float> projZ(T);
	const int nb = T / NUMTHREADSDS + 1;
	launch(nb, 1, NUMTHREADSDS, 1, [&] { DS_spherical_Kernel(sizeList.data(), projZ.data(), faces, verts, mp, xM, yM); });
	float zMin = 9999.9f, zMax = -9999.9f; /* :385-398 */
	for (int i = 0; i < T; i++) {
		if (zMin > projZ[i] && projZ[i] >= 0.0f) zMin = projZ[i];
		if (zMax < projZ[i]) zMax = projZ[i];
	}
	unsigned run = 0; /* :414 cudppScan inclusive */
	for (int i = 0; i < T; i++) { run += sizeList[i]; scanned[i] = run; }
	const long P = T ? (long)scanned[T - 1] : 0;
	if (P > cap) return P;
	std::vector<unsigned> keys(P), vals(P);
	launch(nb, 1, NUMTHREADSDS, 1, [&] { SlabKernel(zSlabs.data(), projZ.data(), mp, zMin, zMax); });
	launch(nb, 1, NUMTHREADSDS, 1, [&] { DS_spherical_Fillkernel(keys.data(), vals.data(), scanned.data(), zSlabs.data(), faces, verts, mp, xM, yM); });
	std::vector<unsigned> perm(P); /* :464 cudppSort (stable LSD radix) */
	std::iota(perm.begin(), perm.end(), 0u);
	std::stable_sort(perm.begin(), perm.end(), [&](unsigned a, unsigned b) { return keys[a] < keys[b]; });
	for (long i = 0; i < P; i++) { key_list[i] = keys[perm[i]]; value_list[i] = vals[perm[i]]; }
	std::vector<unsigned> pos(P), head(P), compacted; /* :470-503 do_scan_dump + cudppCompact */
	launch((int)(P / 256) + 1, 1, 256, 1, [&] { do_scan_dump(key_list, pos.data(), head.data(), (int)P); });
	for (long i = 0; i < P; i++) if (head[i]) compacted.push_back(pos[i]);
	const int ncells = (int)compacted.size();
	launch(C / 256 + 1, 1, 256, 1, [&] { set_as_zero(span); }); /* :506-523 */
	launch(ncells / 256 + 1, 1, 256, 1, [&] { create_histogram(compacted.data(), ncells, (int)P, key_list, span); });
	unsigned acc = 0; /* :527 cudppScan exclusive over ALL cells */
	for (int c = 0; c < C; c++) { offset[c] = acc; acc += span[c]; }
	return P;
}

/* misc_kernel.cu:255-296 mapSort_Effective_kernel, verbatim (dd_camcoords must hold the LIGHT camera) */
void ref_map_rays_spherical(float *t_list, float *dir_list, unsigned *map, float *cmPt, float xM, float yM)
{
	launch(NUM_BLOCKS_X, NUM_BLOCKS_Y, NUM_THREADS_X, NUM_THREADS_Y, [&] { mapSort_Effective_kernel(t_list, dir_list, map, cmPt, xM, yM); });
}

/* trace_kernel.cu:84-270 rckernel_alpha, one emulated CTA (8x8 tile) at a time: the per-ray state of its 64 threads is
 * kept across the slab loop (:132-229) because the "beam done" flag metadata[3] couples them; within a slab every ray
 * walks the cell's list calling the reference's own intersectTriUV / isWithin.  Buffers as decision_data.h:66-82. */
void ref_trace(unsigned *value_list, unsigned *span, unsigned *offset, float *dd_normal, float *dd_t, float *dd_dir,
               int *dd_shadowed, int *dd_id, float *verts, int *tris)
{
	const int W = SCREEN_WIDTH, H = SCREEN_HEIGHT, NBY = NUM_BLOCKS_Y;
#pragma omp parallel for schedule(dynamic, 64) num_threads(g_threads)
	for (int cellrow = 0; cellrow < (W / 8) * (H / 8); cellrow++) {
		const int bx = cellrow / NBY, by = cellrow % NBY;
		float ray_direction[64][3], oldt[64], e1[64][3], e2[64][3];
		int rayDone[64], tri_intersected[64];
		for (int th = 0; th < 64; th++) {
			const int tx = th % 8, ty = th / 8, px = bx * 8 + tx, py = by * 8 + ty;
			/* :96-114 */
			float ftx = (float)px / (float)(NUM_BLOCKS_X * NUM_THREADS_X);
			float fty = (float)py / (float)(NUM_BLOCKS_Y * NUM_THREADS_Y);
			ftx = 1 - ftx;
			ftx = ftx * (0.9 - 0.1) + 0.1;
			fty = fty * (0.9 - 0.1) + 0.1;
			float4 rayDir = tex2D(texdir, ftx, fty);
			rayDir.x -= dd_camcoords[0]; rayDir.y -= dd_camcoords[1]; rayDir.z -= dd_camcoords[2];
			ray_direction[th][0] = rayDir.x; ray_direction[th][1] = rayDir.y; ray_direction[th][2] = rayDir.z;
			NORMALIZE(ray_direction[th]);
			rayDone[th] = 0; oldt[th] = 99999999.9f; tri_intersected[th] = -1;
			for (int k = 0; k < 3; k++) e1[th][k] = e2[th][k] = 0.0f;
		}
		int beam_done = 0; /* metadata[3], :131 */
		for (int slab = 0; slab < NUM_SLABS; slab++) {
			if (beam_done) continue; /* :134 */
			const unsigned cell = (bx * NBY + by) * NUM_SLABS + slab; /* :138 (blockIdx.x*gridDim.y+blockIdx.y)*NUM_SLABS+slab */
			const unsigned n = span[cell], off = offset[cell];
			int all_done = 1;
			for (int th = 0; th < 64; th++) {
				if (rayDone[th] != 2) /* :179 */
					for (unsigned k = 0; k < n; k++) {
						/* staging :153-174 */
						const int curface = value_list[off + k];
						const int f1 = 3 * tris[curface * 3 + 0], f2 = 3 * tris[curface * 3 + 1], f3 = 3 * tris[curface * 3 + 2];
						float vd[9];
						vd[0] = verts[f1 + 0]; vd[1] = verts[f1 + 1]; vd[2] = verts[f1 + 2];
						vd[3] = verts[f2 + 0] - vd[0]; vd[4] = verts[f2 + 1] - vd[1]; vd[5] = verts[f2 + 2] - vd[2];
						vd[6] = verts[f3 + 0] - vd[0]; vd[7] = verts[f3 + 1] - vd[1]; vd[8] = verts[f3 + 2] - vd[2];
						vd[0] = dd_camcoords[0] - vd[0]; vd[1] = dd_camcoords[1] - vd[1]; vd[2] = dd_camcoords[2] - vd[2];
						/* :181-203 */
						float tmp_u, tmp_v;
						float value = intersectTriUV(&vd[0], &vd[3], &vd[6], ray_direction[th], oldt[th], &tmp_u, &tmp_v);
						if (value) {
							oldt[th] = value; rayDone[th] = 1; tri_intersected[th] = off + k;
							e1[th][0] = vd[3]; e1[th][1] = vd[4]; e1[th][2] = vd[5];
							e2[th][0] = vd[6]; e2[th][1] = vd[7]; e2[th][2] = vd[8];
						}
					}
				rayDone[th] = isWithin(rayDone[th], oldt[th], ray_direction[th], slab); /* :217 */
				if (rayDone[th] != 2) all_done = 0;                                       /* :219-228 */
			}
			beam_done = all_done;
		}
		for (int th = 0; th < 64; th++) {
			const int tx = th % 8, ty = th / 8, px = bx * 8 + tx, py = by * 8 + ty;
			const int pixelID = py * W + px; /* :91 */
			float *d_normals = &dd_normal[pixelID * 3];
			if (rayDone[th] == 2) { /* :232-253 */
				NORMALIZE(e1[th]); NORMALIZE(e2[th]);
				float normal[3];
				CROSS(normal, e1[th], e2[th]);
				NORMALIZE(normal);
				(normal[0] < 0) ? normal[0] *= -1 : normal[0] = normal[0];
				(normal[1] < 0) ? normal[1] *= -1 : normal[1] = normal[1];
				(normal[2] < 0) ? normal[2] *= -1 : normal[2] = normal[2];
				dd_t[pixelID] = oldt[th]; dd_shadowed[pixelID] = 0; dd_id[pixelID] = value_list[tri_intersected[th]];
				d_normals[0] = normal[0]; d_normals[1] = normal[1]; d_normals[2] = normal[2];
			} else { /* :255-263 */
				dd_t[pixelID] = -1.0f; dd_shadowed[pixelID] = 0; dd_id[pixelID] = -2;
				d_normals[0] = -1.0f; d_normals[1] = -1.0f; d_normals[2] = -1.0f;
			}
			dd_dir[pixelID * 3 + 0] = ray_direction[th][0]; dd_dir[pixelID * 3 + 1] = ray_direction[th][1]; dd_dir[pixelID * 3 + 2] = ray_direction[th][2];
		}
	}
}

/* misc_kernel.cu:63-129 mapSortkernel, verbatim (dd_camcoords must hold the LIGHT camera) */
void ref_map_rays(float *t_list, float *dir_list, unsigned *map, float *cmPt, float nearPlane)
{
	launch(NUM_BLOCKS_X, NUM_BLOCKS_Y, NUM_THREADS_X, NUM_THREADS_Y, [&] { mapSortkernel(t_list, dir_list, map, cmPt, nearPlane); });
}

/* light_kernel.cu:52-270 mod_light_rckernel, per pixel; sentinel cell => unshadowed.
 * is_shadowed is OR-ed (never cleared here; ref_trace clears it). */
void ref_shadow(unsigned *curflist, float *verts, int *tris, unsigned *blockcnt, unsigned *blockcntscan, float *t_list,
                float *dir_list, int *is_shadowed, unsigned *d_map, float *cmPt)
{
	const int N = IMAGE_SIZE;
	const unsigned sentinel = NUM_BLOCKS_X * NUM_BLOCKS_Y;
#pragma omp parallel for schedule(dynamic, 256) num_threads(g_threads)
	for (int p = 0; p < N; p++) {
		const unsigned cell = d_map[N + p];
		if (cell >= sentinel) continue;
		const int pseudoPixelId = d_map[p];
		/* :171-184 */
		float rayDirection[3], ptIntersection[3];
		float tVal = t_list[pseudoPixelId];
		rayDirection[0] = cmPt[0] + tVal * dir_list[pseudoPixelId * 3 + 0];
		rayDirection[1] = cmPt[1] + tVal * dir_list[pseudoPixelId * 3 + 1];
		rayDirection[2] = cmPt[2] + tVal * dir_list[pseudoPixelId * 3 + 2];
		ptIntersection[0] = rayDirection[0]; ptIntersection[1] = rayDirection[1]; ptIntersection[2] = rayDirection[2];
		rayDirection[0] -= dd_camcoords[0]; rayDirection[1] -= dd_camcoords[1]; rayDirection[2] -= dd_camcoords[2];
		NORMALIZE(rayDirection);
		int done = 0;
		for (int slab = 0; slab < NUM_SLABS && !done; slab++) { /* :105-115: the cell's list of every depth slab in turn */
		const unsigned n = blockcnt[cell * NUM_SLABS + slab], off = blockcntscan[cell * NUM_SLABS + slab];
		for (unsigned k = 0; k < n && !done; k++) {
			/* staging :128-148 */
			const int curface = curflist[off + k];
			const int f1 = 3 * tris[curface * 3 + 0], f2 = 3 * tris[curface * 3 + 1], f3 = 3 * tris[curface * 3 + 2];
			float vd[9];
			vd[0] = verts[f1 + 0]; vd[1] = verts[f1 + 1]; vd[2] = verts[f1 + 2];
			vd[3] = verts[f2 + 0] - vd[0]; vd[4] = verts[f2 + 1] - vd[1]; vd[5] = verts[f2 + 2] - vd[2];
			vd[6] = verts[f3 + 0] - vd[0]; vd[7] = verts[f3 + 1] - vd[1]; vd[8] = verts[f3 + 2] - vd[2];
			vd[0] = dd_camcoords[0] - vd[0]; vd[1] = dd_camcoords[1] - vd[1]; vd[2] = dd_camcoords[2] - vd[2];
			/* :190-202 */
			float value = intersectTri(&vd[0], &vd[3], &vd[6], rayDirection, 999999.9f);
			if (value) {
				float pt[3];
				pt[0] = dd_camcoords[0] + value * rayDirection[0];
				pt[1] = dd_camcoords[1] + value * rayDirection[1];
				pt[2] = dd_camcoords[2] + value * rayDirection[2];
				if (isSmaller(pt, ptIntersection, dd_camcoords)) { done = 1; is_shadowed[pseudoPixelId] = 1; }
			}
		}
		}
	}
}

/* shader.h:70-86 simpleShade -> shader_kernel.cu:165-221.  mat_idx must be readable at index -2
 * (the reference reads mat_idx[-2] on miss pixels, shader_kernel.cu:169-170): callers pass a pointer
 * two ints into a buffer whose first two entries are -1. */
void ref_shade(unsigned char *img, float *normal, float *t, float *dir, int *id, float *cam_pos, int *mat_idx, float *mat_list, int mat_count)
{
	launch(NUM_BLOCKS_X, NUM_BLOCKS_Y, NUM_THREADS_X, NUM_THREADS_Y, [&] { lambertian_shade(img, normal, t, dir, id, cam_pos, mat_idx, mat_list, mat_count); });
}
/* shader.h:88-113 spotlight_shade -> shader_kernel.cu:275-345 spot_shade, verbatim; `dump` = its 2N-float debug buffer */
void ref_spot_shade(unsigned char *img, float *normal, float *t, float *dir, int *id, float *cam_pos, int *mat_idx, float *mat_list, int mat_count, float *dump)
{
	launch(NUM_BLOCKS_X, NUM_BLOCKS_Y, NUM_THREADS_X, NUM_THREADS_Y, [&] { spot_shade(img, normal, t, dir, id, cam_pos, mat_idx, mat_list, mat_count, dump); });
}
/* shader.h:58-68 add_shadows -> shader_kernel.cu:347-359 */
void ref_add_shadows(unsigned char *img, int *is_shadowed)
{
	launch(NUM_BLOCKS_X, NUM_BLOCKS_Y, NUM_THREADS_X, NUM_THREADS_Y, [&] { shadow_kernel(img, is_shadowed); });
}

/* scene.h:225-331 load_model (static branch) + scene.h:370-439 some_material, on the reference's own
 * obj_parser.  Two-call protocol: call with NULL arrays to get the counts, then again with storage. */
int ref_load_obj(char *path, int *num_vertices, int *num_faces, float *vertexlist, int *facelist, int *matidx)
{
	objLoader *objData = new objLoader();
	if (!objData->load(path)) { delete objData; return -1; }
	*num_faces = objData->faceCount;
	*num_vertices = objData->vertexCount;
	if (vertexlist && facelist && matidx) {
		for (int f = 0; f < objData->faceCount; f++) {
			facelist[f * 3 + 0] = objData->faceList[f]->vertex_index[0];
			facelist[f * 3 + 1] = objData->faceList[f]->vertex_index[1];
			facelist[f * 3 + 2] = objData->faceList[f]->vertex_index[2];
			matidx[f] = objData->faceList[f]->material_index;
		}
		for (int v = 0; v < objData->vertexCount; v++)
			for (int k = 0; k < 3; k++) vertexlist[v * 3 + k] = objData->vertexList[v]->e[k];
	}
	delete objData;
	return 0;
}
int ref_load_materials(char *file, int *num_materials, float *materiallist)
{
	FILE *fp = fopen(file, "r");
	if (!fp) return -1;
	char cjunk[512];
	int nm = 0;
	while (fscanf(fp, "%511s", cjunk) != EOF)
		if (strcmp(cjunk, "newmtl") == 0) nm++;
	fclose(fp);
	*num_materials = nm;
	if (!materiallist) return 0;
	fp = fopen(file, "r");
	for (int mt = 0; mt < nm; mt++) {
		for (int i = 0; i < 3; i++) fscanf(fp, "%511s", cjunk);
		for (int i = 0; i < 3; i++) fscanf(fp, "%f", &materiallist[mt * MATERIAL_SIZE + i]);
		fscanf(fp, "%511s", cjunk);
		for (int i = 0; i < 3; i++) fscanf(fp, "%f", &materiallist[mt * MATERIAL_SIZE + 3 + i]);
		for (int i = 0; i < 11; i++) fscanf(fp, "%511s", cjunk);
		fscanf(fp, "%511s", cjunk);
	}
	fclose(fp);
	return 0;
}

} /* extern "C" */
