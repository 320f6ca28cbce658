/* k_shade_dev.cuh -- the Lambert shade of one surface point, shared by k_shade (primary hits) and the
 * reflection kernels (k_reflect.cu). */
#ifndef UG_K_SHADE_DEV_CUH
#define UG_K_SHADE_DEV_CUH

#include "k_common.cuh"

/* lambert_color_pixel (shader_kernel.cu:46-86) at world point P of triangle `tri`, normal as rckernel_alpha leaves it
 * (abs(normalize(cross(normalize(e1), normalize(e2)))), trace_kernel.cu:234-243), clamp and RGB8 truncation of
 * lambertian_shade (:212-220).  Returns r | g<<8 | b<<16. */
static __device__ __forceinline__ u32 ug_lambert_rgb8(const float *R, float Lvx, float Lvy, float Lvz, float Px, float Py, float Pz, float4 r1, float4 r2,
                                                      float kr, float kg, float kb)
{
	float ax = r1.x, ay = r1.y, az = r1.z, bx = r2.x, by = r2.y, bz = r2.z, nx, ny, nz;
	ug_normalize(ax, ay, az);
	ug_normalize(bx, by, bz);
	ug_cross(nx, ny, nz, ax, ay, az, bx, by, bz);
	ug_normalize(nx, ny, nz);
	nx = (nx < 0) ? nx * -1.0f : nx; ny = (ny < 0) ? ny * -1.0f : ny; nz = (nz < 0) ? nz * -1.0f : nz;
	const float Pvx = R[0] * Px + R[4] * Py + R[8] * Pz, Pvy = R[1] * Px + R[5] * Py + R[9] * Pz, Pvz = R[2] * Px + R[6] * Py + R[10] * Pz;
	float Nvx = R[0] * nx + R[4] * ny + R[8] * nz, Nvy = R[1] * nx + R[5] * ny + R[9] * nz, Nvz = R[2] * nx + R[6] * ny + R[10] * nz;
	ug_normalize(Nvx, Nvy, Nvz);
	float ldx = Pvx - Lvx, ldy = Pvy - Lvy, ldz = Pvz - Lvz;
	ug_normalize(ldx, ldy, ldz);
	float cr = 0.0f, cg = 0.0f, cb = 0.0f;
	cr += kr * 0.5f; cg += kg * 0.5f; cb += kb * 0.5f;  /* ambient term uses the DIFFUSE colour, :180-186 */
	float dd = ug_dot(ldx, ldy, ldz, Nvx, Nvy, Nvz);
	if (!(dd > 0)) dd *= -1.0f;
	if (dd > 0) { cr += kr * 1.0f * dd; cg += kg * 1.0f * dd; cb += kb * 1.0f * dd; }
	cr = cr > 1.0f ? 1.0f : cr; cg = cg > 1.0f ? 1.0f : cg; cb = cb > 1.0f ? 1.0f : cb;
	const unsigned char r8 = (unsigned char)(cr * 255.0f), g8 = (unsigned char)(cg * 255.0f), b8 = (unsigned char)(cb * 255.0f);
	return (u32)r8 | ((u32)g8 << 8) | ((u32)b8 << 16);
}

/* one pixel: shader_kernel.cu:165-221 (+ :347-359 when apply_shadow) */
static __device__ __forceinline__ u32 shade_pixel(const float *R, float ex, float ey, float ez, float Lvx, float Lvy, float Lvz, float tv, float dx, float dy,
                                                  float dz, int tri, int is_shadowed, const float4 *__restrict__ rec, const int *__restrict__ mat_idx,
                                                  const float *__restrict__ mat_list, int mat_count, int apply_shadow, int &idx_out)
{
	const int idx = (tri >= 0) ? mat_idx[tri] : -1; /* the reference reads mat_idx[-2] on a miss; ours: black */
	idx_out = idx;
	u32 c = 0;
	if (idx >= 0 && idx < mat_count && tv > 0.0f) {
		const float Px = ex + tv * dx, Py = ey + tv * dy, Pz = ez + tv * dz;
		const float4 r1 = rec[4 * (size_t)tri + 1], r2 = rec[4 * (size_t)tri + 2];
		c = ug_lambert_rgb8(R, Lvx, Lvy, Lvz, Px, Py, Pz, r1, r2, mat_list[idx * UG_MATERIAL_SIZE + 3], mat_list[idx * UG_MATERIAL_SIZE + 4],
		                    mat_list[idx * UG_MATERIAL_SIZE + 5]);
	}
	if (apply_shadow && is_shadowed == 1) /* shader_kernel.cu:351-356: integer division per channel */
		c = ((c & 255u) / 3u) | ((((c >> 8) & 255u) / 3u) << 8) | ((((c >> 16) & 255u) / 3u) << 16);
	return c;
}

#endif
