/* k_shade_dev.cuh -- the Lambert shade of one surface point, shared by k_shade (primary hits) and the
 * reflection kernels (k_reflect.cu). */
#ifndef UG_K_SHADE_DEV_CUH
#define UG_K_SHADE_DEV_CUH

#include "k_common.cuh"

/* lambert_color_pixel (shader_kernel.cu:46-86) at world point P of triangle `tri`, normal as rckernel_alpha leaves it
 * (abs(normalize(cross(normalize(e1), normalize(e2)))), trace_kernel.cu:234-243), clamp and RGB8 truncation of
 * lambertian_shade (:212-220).  Returns r | g<<8 | b<<16. */
/* SPOT: lambert_color_drop_off_pixel (shader_kernel.cu:88-128) - the same with every term scaled by `drop` last */
/* the part of it that depends on the triangle and the resident modelview only: Nv = normalize(R * abs(normalize(cross(
 * normalize(e1), normalize(e2))))).  k_shade_normals evaluates it once per triangle and frame with exactly these
 * operations, so a pixel that reads the stored value gets the bits it would have computed itself. */
static __device__ __forceinline__ void ug_shade_normal(const float *R, float4 r1, float4 r2, float &Nvx, float &Nvy, float &Nvz)
{
	float ax = r1.x, ay = r1.y, az = r1.z, bx = r2.x, by = r2.y, bz = r2.z, nx, ny, nz;
	ug_normalize(ax, ay, az);
	ug_normalize(bx, by, bz);
	ug_cross(nx, ny, nz, ax, ay, az, bx, by, bz);
	ug_normalize(nx, ny, nz);
	nx = (nx < 0) ? nx * -1.0f : nx; ny = (ny < 0) ? ny * -1.0f : ny; nz = (nz < 0) ? nz * -1.0f : nz;
	Nvx = R[0] * nx + R[4] * ny + R[8] * nz; Nvy = R[1] * nx + R[5] * ny + R[9] * nz; Nvz = R[2] * nx + R[6] * ny + R[10] * nz;
	ug_normalize(Nvx, Nvy, Nvz);
}
/* ... and the part that depends on the point */
template <bool SPOT>
static __device__ __forceinline__ u32 ug_lambert_rgb8_n(const float *R, float Lvx, float Lvy, float Lvz, float Px, float Py, float Pz, float Nvx, float Nvy, float Nvz,
                                                        float kr, float kg, float kb, float drop)
{
	const float Pvx = R[0] * Px + R[4] * Py + R[8] * Pz, Pvy = R[1] * Px + R[5] * Py + R[9] * Pz, Pvz = R[2] * Px + R[6] * Py + R[10] * Pz;
	float ldx = Pvx - Lvx, ldy = Pvy - Lvy, ldz = Pvz - Lvz;
	ug_normalize(ldx, ldy, ldz);
	float cr = 0.0f, cg = 0.0f, cb = 0.0f;
	if (SPOT) { cr += kr * 0.5f * drop; cg += kg * 0.5f * drop; cb += kb * 0.5f * drop; }
	else { cr += kr * 0.5f; cg += kg * 0.5f; cb += kb * 0.5f; } /* ambient term uses the DIFFUSE colour, :180-186 */
	float dd = ug_dot(ldx, ldy, ldz, Nvx, Nvy, Nvz);
	if (!(dd > 0)) dd *= -1.0f;
	if (dd > 0) {
		if (SPOT) { cr += kr * 1.0f * dd * drop; cg += kg * 1.0f * dd * drop; cb += kb * 1.0f * dd * drop; }
		else { cr += kr * 1.0f * dd; cg += kg * 1.0f * dd; cb += kb * 1.0f * dd; }
	}
	cr = cr > 1.0f ? 1.0f : cr; cg = cg > 1.0f ? 1.0f : cg; cb = cb > 1.0f ? 1.0f : cb;
	const unsigned char r8 = (unsigned char)(cr * 255.0f), g8 = (unsigned char)(cg * 255.0f), b8 = (unsigned char)(cb * 255.0f);
	return (u32)r8 | ((u32)g8 << 8) | ((u32)b8 << 16);
}
template <bool SPOT>
static __device__ __forceinline__ u32 ug_lambert_rgb8_t(const float *R, float Lvx, float Lvy, float Lvz, float Px, float Py, float Pz, float4 r1, float4 r2,
                                                        float kr, float kg, float kb, float drop)
{
	float Nvx, Nvy, Nvz;
	ug_shade_normal(R, r1, r2, Nvx, Nvy, Nvz);
	return ug_lambert_rgb8_n<SPOT>(R, Lvx, Lvy, Lvz, Px, Py, Pz, Nvx, Nvy, Nvz, kr, kg, kb, drop);
}

static __device__ __forceinline__ u32 ug_lambert_rgb8(const float *R, float Lvx, float Lvy, float Lvz, float Px, float Py, float Pz, float4 r1, float4 r2,
                                                      float kr, float kg, float kb)
{
	return ug_lambert_rgb8_t<false>(R, Lvx, Lvy, Lvz, Px, Py, Pz, r1, r2, kr, kg, kb, 1.0f);
}

/* spot_shade's cone test (shader_kernel.cu:223-273 get_along_x/y, :291-301): signed angles of the direction from the
 * resident (light) camera's eye to the point, about its up and right axes; full intensity inside +-pi/4 on both, a
 * quarter outside.  R = the resident modelview (rows right / up / -forward); get_along_y keeps the `a + b*c*d` dot
 * product of getEffective_y (:262).  The comparisons are against the DOUBLE M_PI/4. */
static __device__ __forceinline__ float ug_spot_drop_off(const float *R, float lex, float ley, float lez, float Px, float Py, float Pz)
{
	float vx = Px - lex, vy = Py - ley, vz = Pz - lez;
	ug_normalize(vx, vy, vz);
	float x, y;
	{
		const float upd = vx * R[1] + vy * R[5] + vz * R[9];
		float tx = vx - upd * R[1], ty = vy - upd * R[5], tz = vz - upd * R[9];
		const float val = sqrtf(tx * tx + ty * ty + tz * tz);
		tx /= val; ty /= val; tz /= val;
		const float fwd = tx * R[2] + ty * R[6] + tz * R[10];
		const float angle = ug_acosf(fwd);
		const float rightd = tx * R[0] + ty * R[4] + tz * R[8];
		x = rightd > 0.0f ? angle : -angle;
	}
	{
		const float rd = vx * R[0] + vy * R[4] + vz * R[8];
		float tx = vx - rd * R[0], ty = vy - rd * R[4], tz = vz - rd * R[8];
		const float val = sqrtf(tx * tx + ty * ty + tz * tz);
		tx /= val; ty /= val; tz /= val;
		const float upd = tx * R[1] + ty * R[5] + tz * R[9];
		const float fwd = tx * R[2] + ty * R[6] * tz * R[10];
		const float angle = ug_acosf(fwd);
		y = upd > 0.0f ? angle : -angle;
	}
	const double q = 0.78539816339744830962; /* M_PI/4 */
	return ((double)x < q && (double)x > -q && (double)y < q && (double)y > -q) ? 1.0f : 0.25f;
}

/* one pixel: shader_kernel.cu:165-221 lambertian_shade, or :275-345 spot_shade when SPOT (+ :347-359 when apply_shadow).
 * nv != nullptr: the per-triangle view-space normals k_shade_normals made for this frame (one 16-byte load instead of the
 * two record rows and four of the pixel's five normalisations). */
template <bool SPOT, bool HOIST = false>
static __device__ __forceinline__ u32 shade_pixel(const float *R, const float *leye, float ex, float ey, float ez, float Lvx, float Lvy, float Lvz, float tv, float dx, float dy,
                                                  float dz, int tri, int is_shadowed, const float4 *__restrict__ rec, const int *__restrict__ mat_idx,
                                                  const float *__restrict__ mat_list, int mat_count, int apply_shadow, int &idx_out, const float4 *__restrict__ nv = nullptr)
{
	const int idx = (tri >= 0) ? mat_idx[tri] : -1; /* the reference reads mat_idx[-2] on a miss; ours: black */
	idx_out = idx;
	u32 c = 0;
	if (idx >= 0 && idx < mat_count && tv > 0.0f) {
		const float Px = ex + tv * dx, Py = ey + tv * dy, Pz = ez + tv * dz;
		const float drop = SPOT ? ug_spot_drop_off(R, leye[0], leye[1], leye[2], Px, Py, Pz) : 1.0f;
		const float kr = mat_list[idx * UG_MATERIAL_SIZE + 3], kg = mat_list[idx * UG_MATERIAL_SIZE + 4], kb = mat_list[idx * UG_MATERIAL_SIZE + 5];
		if (HOIST) {
			const float4 n4 = nv[tri];
			c = ug_lambert_rgb8_n<SPOT>(R, Lvx, Lvy, Lvz, Px, Py, Pz, n4.x, n4.y, n4.z, kr, kg, kb, drop);
		} else {
			const float4 r1 = rec[4 * (size_t)tri + 1], r2 = rec[4 * (size_t)tri + 2];
			c = ug_lambert_rgb8_t<SPOT>(R, Lvx, Lvy, Lvz, Px, Py, Pz, r1, r2, kr, kg, kb, drop);
		}
	}
	if (apply_shadow && is_shadowed == 1) /* shader_kernel.cu:351-356: integer division per channel */
		c = ((c & 255u) / 3u) | ((((c >> 8) & 255u) / 3u) << 8) | ((((c >> 16) & 255u) / 3u) << 16);
	return c;
}

#endif
