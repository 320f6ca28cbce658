/* k_common.cuh -- device helpers shared by the kernels.
 *
 * Arithmetic contract: results must be bit-identical to the reference's
 * device code compiled as host C (x86-64 SSE2, no FMA).  The whole library is compiled with
 * -fmad=false (no FFMA/DFMA contraction), IEEE division and square root (nvcc defaults
 * -prec-div=true -prec-sqrt=true -ftz=false); evaluation order below is the reference's.
 */
#ifndef UG_K_COMMON_CUH
#define UG_K_COMMON_CUH

#include "ugrt_internal.h"

#define UG_FULL_MASK 0xffffffffu

/* main.cu.h:44-56 */
__device__ __forceinline__ float ug_dot(float ax, float ay, float az, float bx, float by, float bz)
{
	return ax * bx + ay * by + az * bz; /* ((ax*bx + ay*by) + az*bz), no contraction */
}
__device__ __forceinline__ void ug_cross(float &dx, float &dy, float &dz, float ax, float ay, float az, float bx, float by, float bz)
{
	dx = ay * bz - az * by;
	dy = az * bx - ax * bz;
	dz = ax * by - ay * bx;
}
__device__ __forceinline__ void ug_normalize(float &x, float &y, float &z)
{
	const float l = 1.0f / sqrtf(x * x + y * y + z * z); /* NORMALIZE: 1/sqrtf, then three multiplies */
	x *= l; y *= l; z *= l;
}

/* float -> int the way the host-compiled reference does it (x86 cvttss2si: NaN, +-inf and
 * out-of-range give INT_MIN; PTX cvt.rzi saturates instead) - grid_kernel.cu:208-215 */
__device__ __forceinline__ int ug_f2i(float f)
{
	return (f >= 2147483648.0f || f < -2147483648.0f || f != f) ? (int)0x80000000 : (int)f;
}
__device__ __forceinline__ int ug_d2i(double f)
{
	return (f >= 2147483648.0 || f < -2147483648.0 || f != f) ? (int)0x80000000 : (int)f;
}

/* float -> unsigned the way gcc does it on x86-64 (grid_kernel.cu:475-478 assigns a float expression to an
 * `unsigned int`): cvttss2si into a 64-bit register, low 32 bits kept; NaN / out of range give 0x8000000000000000 -> 0 */
__device__ __forceinline__ u32 ug_f2u(float f)
{
	return (f >= 9223372036854775808.0f || f < -9223372036854775808.0f || f != f) ? 0u : (u32)(unsigned long long)__float2ll_rz(f);
}

/* acosf exactly as the reference's host build gets it from glibc 2.39 (sysdeps/ieee754/flt-32/e_acosf.c, the
 * fdlibm float algorithm: only float + - * / and sqrtf, so -fmad=false / IEEE div+sqrt reproduce it bit for bit;
 * tests/test_abi.py checks this restatement against libm through ug_acosf_host).  CUDA's own acosf differs in the
 * last ulp, which would move vertices and rays across cell borders of the spherical light grid (a14/a15). */
__host__ __device__ __forceinline__ float ug_acosf(float x)
{
	const float one = 1.0f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f,
	            pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f,
	            pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f,
	            qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
#ifdef __CUDA_ARCH__
	const int hx = __float_as_int(x);
#else
	int hx; memcpy(&hx, &x, 4);
#endif
	const int ix = hx & 0x7fffffff;
	if (ix == 0x3f800000) return hx > 0 ? 0.0f : pi + 2.0f * pio2_lo; /* |x| == 1 */
	if (ix > 0x3f800000) return (x - x) / (x - x);                    /* |x| > 1 or NaN: NaN */
	if (ix < 0x3f000000) {                                            /* |x| < 0.5 */
		if (ix <= 0x23000000) return pio2_hi + pio2_lo;
		const float z = x * x;
		const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
		const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
		const float r = p / q;
		return pio2_hi - (x - (pio2_lo - x * r));
	}
	if (hx < 0) {                                                     /* x < -0.5 */
		const float z = (one + x) * 0.5f;
		const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
		const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
		const float s = sqrtf(z);
		const float r = p / q;
		const float w = r * s - pio2_lo;
		return pi - 2.0f * (s + w);
	}
	const float z = (one - x) * 0.5f;                                 /* x > 0.5 */
	const float s = sqrtf(z);
#ifdef __CUDA_ARCH__
	const float df = __int_as_float(__float_as_int(s) & 0xfffff000);
#else
	int is_; memcpy(&is_, &s, 4); is_ &= 0xfffff000; float df; memcpy(&df, &is_, 4);
#endif
	const float c = (z - df * df) / (s + df);
	const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
	const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
	const float r = p / q;
	const float w = r * s + c;
	return 2.0f * (df + w);
}

/* grid_kernel.cu:4-11: one row of a column-major 4x4 times (x,y,z,w), summed left to right */
__device__ __forceinline__ float ug_row(const float *m, int r, float x, float y, float z, float w)
{
	return m[r] * x + m[4 + r] * y + m[8 + r] * z + m[12 + r] * w;
}

/* a9: primary ray through pixel (px,py) - trace_kernel.cu:96-114 with the texture fetch done in
 * software: CUDA linear filtering (xB = s*5-0.5, i = floor, alpha = frac, clamp), exact float weights.
 * `lat` = the 5x5 float4 lattice, `eye` = cc[0..2]. */
__device__ __forceinline__ void ug_primary_dir(const float *lat, float ex, float ey, float ez, int px, int py, int Wfull, int Hfull,
                                               float &dx, float &dy, float &dz)
{
	float ftx = (float)px / (float)Wfull;
	float fty = (float)py / (float)Hfull;
	ftx = 1.0f - ftx;
	ftx = (float)__dadd_rn(__dmul_rn((double)ftx, 0.9 - 0.1), 0.1); /* double, trace_kernel.cu:102-103 */
	fty = (float)__dadd_rn(__dmul_rn((double)fty, 0.9 - 0.1), 0.1);
	const float xB = ftx * 5.0f - 0.5f, yB = fty * 5.0f - 0.5f;
	const float fi = floorf(xB), fj = floorf(yB);
	const float a = xB - fi, b = yB - fj;
	int i0 = (int)fi, j0 = (int)fj, i1 = i0 + 1, j1 = j0 + 1;
	i0 = min(max(i0, 0), 4); i1 = min(max(i1, 0), 4);
	j0 = min(max(j0, 0), 4); j1 = min(max(j1, 0), 4);
	const float *T00 = lat + 4 * (j0 * 5 + i0), *T10 = lat + 4 * (j0 * 5 + i1);
	const float *T01 = lat + 4 * (j1 * 5 + i0), *T11 = lat + 4 * (j1 * 5 + i1);
	const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
	dx = w00 * T00[0] + w10 * T10[0] + w01 * T01[0] + w11 * T11[0];
	dy = w00 * T00[1] + w10 * T10[1] + w01 * T01[1] + w11 * T11[1];
	dz = w00 * T00[2] + w10 * T10[2] + w01 * T01[2] + w11 * T11[2];
	dx -= ex; dy -= ey; dz -= ez;
	ug_normalize(dx, dy, dz);
}

/* Per-triangle record for one ray origin O (the eye or a light): everything of Moeller-Trumbore that
 * does not depend on the ray.  r0 = (O-v0, dot(e2,q)), r1 = (e1, 0), r2 = (e2, 0), r3 = (q = cross(O-v0,e1), 0).
 * trace_kernel.cu:4-45 / light_kernel.cu:13-50 evaluate q and dot(e2,q) per ray from the same inputs
 * with the same operations, so hoisting them is bit-exact. */
struct TriRec { float4 r0, r1, r2, r3; };

/* ray-dependent part; returns false on a miss, else t (signed) */
__device__ __forceinline__ bool ug_intersect(const float4 &r0, const float4 &r1, const float4 &r2, const float4 &r3,
                                             float dx, float dy, float dz, float &t)
{
	float px, py, pz;
	ug_cross(px, py, pz, dx, dy, dz, r2.x, r2.y, r2.z);     /* pvec = dir x e2 */
	const float det = ug_dot(r1.x, r1.y, r1.z, px, py, pz); /* e1 . pvec */
	/* `det > -EPSILON && det < EPSILON` compares against the DOUBLE 1e-21 (main.cu.h:42).  float(1e-21)
	 * = 0x1c971da0 lies just below it, so the double test is true exactly when |det| <= float(1e-21). */
	if (fabsf(det) <= 1e-21f) return false;
	const float inv = 1.0f / det;                           /* (float)(1.0/det): one correctly rounded divide */
	const float u = ug_dot(r0.x, r0.y, r0.z, px, py, pz) * inv;
	if (u < 0.0f || u > 1.0f) return false;
	const float v = ug_dot(dx, dy, dz, r3.x, r3.y, r3.z) * inv;
	if (v < 0.0f || u + v > 1.0f) return false;
	t = r0.w * inv;
	return true;
}

/* same test with the record read from shared memory by 32-bit shared address (broadcast LDS.128);
 * the loads are issued only when the value is needed, so a ray rejected by the u test never reads qvec */
__device__ __forceinline__ float4 ug_lds128(u32 addr)
{
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
	return v;
}
__device__ __forceinline__ bool ug_intersect_lds(u32 addr, float dx, float dy, float dz, float &t)
{
	const float4 r2 = ug_lds128(addr + 32), r1 = ug_lds128(addr + 16);
	float px, py, pz;
	ug_cross(px, py, pz, dx, dy, dz, r2.x, r2.y, r2.z);
	const float det = ug_dot(r1.x, r1.y, r1.z, px, py, pz);
	if (fabsf(det) <= 1e-21f) return false;
	const float inv = 1.0f / det;
	const float4 r0 = ug_lds128(addr);
	const float u = ug_dot(r0.x, r0.y, r0.z, px, py, pz) * inv;
	if (u < 0.0f || u > 1.0f) return false;
	const float4 r3 = ug_lds128(addr + 48);
	const float v = ug_dot(dx, dy, dz, r3.x, r3.y, r3.z) * inv;
	if (v < 0.0f || u + v > 1.0f) return false;
	t = r0.w * inv;
	return true;
}

/* The same test cut in two, so that a caller can keep TWO triangles in flight per ray: the first half (pvec, det, 1/det, u)
 * is evaluated for both without a branch in between - two independent dependency chains instead of one, which is what the
 * issue slots of a warp were waiting for - and only a ray that passes a u test goes on to v and t.  Operations and their
 * order are those of ug_intersect, so the results are bit-identical; a |det| <= EPSILON triangle gets inf/NaN here and is
 * masked by the returned flag instead of leaving early. */
struct TriU { float inv, u; };
__device__ __forceinline__ bool ug_isect_u(const float4 &r0, const float4 &r1, const float4 &r2, float dx, float dy, float dz, TriU &o)
{
	float px, py, pz;
	ug_cross(px, py, pz, dx, dy, dz, r2.x, r2.y, r2.z);
	const float det = ug_dot(r1.x, r1.y, r1.z, px, py, pz);
	o.inv = 1.0f / det;
	o.u = ug_dot(r0.x, r0.y, r0.z, px, py, pz) * o.inv;
	return !(fabsf(det) <= 1e-21f) && !(o.u < 0.0f || o.u > 1.0f);
}
__device__ __forceinline__ bool ug_isect_vt(const float4 &r0, const float4 &r3, float dx, float dy, float dz, const TriU &o, float &t)
{
	const float v = ug_dot(dx, dy, dz, r3.x, r3.y, r3.z) * o.inv;
	if (v < 0.0f || o.u + v > 1.0f) return false;
	t = r0.w * o.inv;
	return true;
}

__device__ __forceinline__ u32 ug_lane() { return threadIdx.x & 31u; }

/* ---------------------------------------------------------------- programmatic dependent launch (UG_PDL builds)
 * The frame is a chain of short dependent kernels (a dozen boundaries on its critical path at 2-3 us each).  With
 * -DUG_PDL=1 the kernels of that chain are launched with cudaLaunchAttributeProgrammaticStreamSerialization, which
 * becomes a programmatic edge of the captured frame graph: the next kernel's blocks can be dispatched before this grid
 * has been retired and sit in griddepcontrol.wait until every prerequisite grid has finished and its stores are
 * visible.  Rules that keep this exactly as safe as a plain launch:
 *   - ug_pdl_enter<group>() is the FIRST statement of every kernel that is ever launched through ug_launch: nothing of
 *     a predecessor is read and nothing is written before it (completion is transitive along the chain because every
 *     kernel of the chain passes its own wait before it ends);
 *   - in a kernel launched the plain way the instructions do nothing.
 * When the dependents go is a per-group choice (UG_PDL_EARLY, a mask of UG_PDL_* bits): a kernel of a group in the mask
 * lets them go right behind its own wait (launch_dependents), so they become resident as soon as all of this grid's
 * blocks have started; the others let them go when their last block has exited.  Early is not free: blocks that wait
 * hold their threads and registers, and beside a multi-wave kernel of another branch of the frame graph they keep that
 * kernel's next wave out (measured on config B: plain launches 0.351 ms, nothing early 0.346, everything early 0.401 -
 * it is the sweep passes and the primary rays whose early dependents cost, 0.366 each; the table is at UG_PDL_EARLY).
 * -DUG_PDL=0 compiles neither instruction and ug_launch is the <<<>>> launch (SASS identical to the round's earlier builds). */
#ifndef UG_PDL
#define UG_PDL 1
#endif
enum { UG_PDL_SWEEP = 1, UG_PDL_TRACE = 2, UG_PDL_REGROUP = 4, UG_PDL_SHADOW = 8, UG_PDL_SHADE = 16, UG_PDL_REFLECT = 32, UG_PDL_PUBLISH = 64, UG_PDL_ALL = 127 };
#ifndef UG_PDL_EARLY
#define UG_PDL_EARLY UG_PDL_SHADOW /* measured on the B200 (config B, ms/frame; plain launches 0.3513): 0 = 0.3464, SHADOW 0.3460, SHADE 0.3460, REFLECT 0.3467, PUBLISH 0.3463,
                                    * REGROUP 0.3474, TRACE 0.3660, SWEEP 0.3668, all 0.4016 - the gain is the attribute, not the early trigger; SHADOW is the build the
                                    * parity tests of that run were made on, so it is the default */
#endif
template <unsigned GROUP>
__device__ __forceinline__ void ug_pdl_enter()
{
#if UG_PDL
	asm volatile("griddepcontrol.wait;" ::: "memory");
	if ((unsigned)(UG_PDL_EARLY) & GROUP) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

#ifdef __CUDACC__
/* which kernels of the chain get the attribute: one bit per group, set per enqueued frame by context.cu (UG_PDL builds; 0
 * = plain launches: a frame that is gathered over a communicator, UG_NO_PDL=1; UG_PDL_MASK picks groups) */
extern thread_local unsigned ugk_pdl_mask;
template <typename... P, typename... A>
static inline void ug_launch(unsigned group, void (*kernel)(P...), dim3 grid, dim3 block, cudaStream_t s, A &&...args)
{
#if UG_PDL
	if (ugk_pdl_mask & group) {
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = grid;
		cfg.blockDim = block;
		cfg.dynamicSmemBytes = 0;
		cfg.stream = s;
		cudaLaunchAttribute at[1];
		at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		at[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = at;
		cfg.numAttrs = 1;
		cudaLaunchKernelEx(&cfg, kernel, static_cast<A &&>(args)...); /* an error stays in cudaGetLastError, like <<<>>> */
		return;
	}
#endif
	(void)group;
	kernel<<<grid, block, 0, s>>>(static_cast<A &&>(args)...);
}
#endif

#endif
