/* k_shadow_dev.cuh -- device functions shared by the two shadow-ray paths (k_shadow.cu: rays regrouped by light cell;
 * k_shadowdirect.cu: rays taken in screen order): the ray -> light cell map and the shadow ray itself. */
#ifndef UG_K_SHADOW_DEV_CUH
#define UG_K_SHADOW_DEV_CUH

#include "k_grid_dev.cuh"

/* phase A body: misc_kernel.cu:63-129 for one pixel */
struct LightMap {
	float lx, ly, lz, p0, p1, p2, num, half, r0, r1, r2, u0, u1, u2;
	double den;
	float inv_den;  /* (float)(1/den): only used to decide whether the double expression can be skipped */
	int nbx, nby, shards;
};
static __device__ __forceinline__ LightMap make_light_map(const ViewConst &lv)
{
	LightMap m;
	const float *cc = lv.cc;
	m.lx = cc[0]; m.ly = cc[1]; m.lz = cc[2];
	/* misc_kernel.cu:97-101: un-normalised near plane = MVP row 3 + row 2 */
	m.p0 = cc[48 + 3] + cc[48 + 2]; m.p1 = cc[48 + 7] + cc[48 + 6]; m.p2 = cc[48 + 11] + cc[48 + 10];
	const float p3 = cc[48 + 15] + cc[48 + 14];
	m.num = ug_dot(m.p0, m.p1, m.p2, m.lx, m.ly, m.lz) + p3;
	m.half = lv.near_plane * lv.tan_half;                                                   /* float */
	m.den = __dmul_rn(__dmul_rn(2.0, (double)lv.near_plane), (double)lv.tan_half);          /* 2.0 * near * tan: double */
	m.inv_den = (float)(1.0 / m.den);
	m.r0 = cc[16 + 0]; m.r1 = cc[16 + 4]; m.r2 = cc[16 + 8];
	m.u0 = cc[16 + 1]; m.u1 = cc[16 + 5]; m.u2 = cc[16 + 9];
	m.nbx = lv.nbx; m.nby = lv.nby; m.shards = lv.shards;
	return m;
}
/* floor(((double)a / den) * nb) as an int, the reference's double expression.  A float estimate of the same value is within
 * 2e-7 relative of it (three roundings); wherever the estimate is further than that from an integer, its floor IS the floor of
 * the double expression, and the two double divides per ray (the most expensive instructions of the map on this part) are
 * skipped.  Rays whose estimate is close to a cell border, out of range or not finite take the exact path. */
static __device__ __forceinline__ int light_bin(float a, const LightMap &m, int nb)
{
	const float vf = (a * m.inv_den) * (float)nb;
	const float fl = floorf(vf), fr = vf - fl;
	const float tol = fabsf(vf) * 4e-7f + 1e-6f;
	if (fr > tol && fr < 1.0f - tol && fabsf(vf) < 1.0e6f) return (int)fl;
	return ug_d2i(floor(__dmul_rn(__ddiv_rn((double)a, m.den), (double)nb)));
}

static __device__ __forceinline__ u32 light_cell_of(const LightMap &m, float ex, float ey, float ez, float tv, float dx, float dy, float dz)
{
	/* misc_kernel.cu:74-88: P = eye + t*dir (misses have t = -1 and still shoot), L = normalize(P - light) */
	const float Px = ex + tv * dx, Py = ey + tv * dy, Pz = ez + tv * dz;
	float Lx = Px - m.lx, Ly = Py - m.ly, Lz = Pz - m.lz;
	ug_normalize(Lx, Ly, Lz);
	const float s = -1.0f * m.num / ug_dot(m.p0, m.p1, m.p2, Lx, Ly, Lz);
	const float X = s * Lx, Y = s * Ly, Z = s * Lz;
	const float right = X * m.r0 + Y * m.r1 + Z * m.r2;
	const float up = X * m.u0 + Y * m.u1 + Z * m.u2;
	/* misc_kernel.cu:115-116: (float)/(double) -> double, * int, floor, -> int */
	const int blx = light_bin(right + m.half, m, m.nbx);
	const int bly = light_bin(up + m.half, m, m.nby);
	u32 cell = (u32)(m.nbx * m.nby);
	if (blx >= 0 && blx <= m.nbx - 1 && bly >= 0 && bly <= m.nby - 1) {
		if (m.shards > 1) { /* the grid was built in row shards (k_lightshare.cu): cells are numbered shard-major */
			const int q = ((bly + 1) * m.shards - 1) / m.nby, b0 = m.nby * q / m.shards, b1 = m.nby * (q + 1) / m.shards;
			cell = (u32)(m.nbx * b0 + blx * (b1 - b0) + (bly - b0));
		} else cell = (u32)(blx * m.nby + bly);
	}
	return cell;
}

/* a14: misc_kernel.cu:255-296 mapSort_Effective_kernel - the spherical variant of the same map (the one display()
 * actually calls, main.cu:189): angular bins of the normalised light->point direction */
static __device__ __noinline__ u32 light_cell_sph(const ViewConst &lv, float ex, float ey, float ez, float tv, float dx, float dy, float dz)
{
	const float *cc = lv.cc;
	const float Px = ex + tv * dx, Py = ey + tv * dy, Pz = ez + tv * dz;
	float Lx = Px - cc[0], Ly = Py - cc[1], Lz = Pz - cc[2];
	ug_normalize(Lx, Ly, Lz);
	const int blx = ug_effective_x(cc, Lx, Ly, Lz, lv.xM, lv.nbx);
	const int bly = ug_effective_y(cc, Lx, Ly, Lz, lv.yM, lv.nby);
	u32 cell = (u32)(lv.nbx * lv.nby);
	if (blx >= 0 && blx < lv.nbx && bly >= 0 && bly < lv.nby) cell = (u32)(blx * lv.nby + bly);
	return cell;
}

struct ShadowRay { float Lx, Ly, Lz, dist_b; int pix; };

/* light_kernel.cu:164-184: rebuild the shadow ray of one pixel from (t, dir) */
static __device__ __forceinline__ ShadowRay make_shadow_ray(const float *__restrict__ dir, const float *__restrict__ t, int pix, float ex, float ey, float ez,
                                                            float lx, float ly, float lz)
{
	ShadowRay r;
	const float tv = t[pix], dx = dir[3 * (size_t)pix], dy = dir[3 * (size_t)pix + 1], dz = dir[3 * (size_t)pix + 2];
	const float Px = ex + tv * dx, Py = ey + tv * dy, Pz = ez + tv * dz;
	r.Lx = Px - lx; r.Ly = Py - ly; r.Lz = Pz - lz;
	const float bx = r.Lx, by = r.Ly, bz = r.Lz;
	ug_normalize(r.Lx, r.Ly, r.Lz);
	r.dist_b = sqrtf(bx * bx + by * by + bz * bz); /* isSmaller's distance_b, light_kernel.cu:5 */
	r.pix = pix;
	return r;
}


#endif
