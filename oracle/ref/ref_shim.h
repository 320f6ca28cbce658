/* ref_shim.h -- TEST INFRASTRUCTURE ONLY (oracle/): host-compilation shim for the reference's CUDA files.
 *
 * The reference (/root/reference) is one CUDA-2.x translation unit that cannot be built any more
 * (cutil, cudpp, GLUT, sm_12, texture references).  Its *arithmetic* is plain C though: with the few
 * definitions below the six kernel files (grid_kernel.cu, misc_kernel.cu, trace_kernel.cu,
 * light_kernel.cu, shader_kernel.cu, transformation_kernel.cu) and camera.h compile UNMODIFIED with
 * g++ from where they lie.  oracle/ref/ref_driver.cpp includes them and exports `ref_*` C functions
 * that run the kernels by index emulation; the result is oracle/_ref/libugref.so.
 *
 * Nothing here is shipped or measured as product; it validates oracle/ug_oracle.c and is the
 * "reference" CPU baseline of bench.py.
 *
 * Compile-time configuration macros of main.cu.h:10-42 become runtime variables (the reference is
 * hard-wired to 1024x1024 / 128x128 cells; BASELINE configs need other sizes).  The macro NAMES are
 * kept so the reference sources see what they expect.
 */
#ifndef UG_REF_SHIM_H
#define UG_REF_SHIM_H

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define __device__
#define __global__
#define __constant__
#define __shared__
#define __host__

struct ref_uint3 { unsigned int x, y, z; };
typedef ref_uint3 uint3;
typedef ref_uint3 dim3;
extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
static inline void __syncthreads() {}

struct float4 { float x, y, z, w; };

/* main.cu.h:10-42 as runtime values */
extern int g_ref_W, g_ref_H, g_ref_NBX, g_ref_NBY, g_ref_FOVY, g_ref_SLABS;
#define SCREEN_WIDTH    g_ref_W
#define SCREEN_HEIGHT   g_ref_H
#define IMAGE_SIZE      (SCREEN_WIDTH * SCREEN_HEIGHT)
#define FOVY            g_ref_FOVY
#define NUM_BLOCKS_X    g_ref_NBX
#define NUM_BLOCKS_Y    g_ref_NBY
#define NUM_SLABS       g_ref_SLABS /* main.cu.h:18 compiles 1; ref_set_slabs exercises the code paths for more */
#define NUM_THREADS_X   8
#define NUM_THREADS_Y   8
#define MAX_TRIANGLES   (NUM_THREADS_X * NUM_THREADS_Y)
#define TOTAL_THREADS   (NUM_THREADS_X * NUM_THREADS_Y)
#define MAX_RAYS_PER_BLOCK 64
#define MATERIAL_SIZE   6
#define NUMTHREADSDS    64
#define EPSILON 0.000000000000000000001

/* main.cu.h:44-56, verbatim semantics (these four macros ARE the arithmetic contract) */
#define CROSS(dest,v1,v2) \
	dest[0]=v1[1]*v2[2]-v1[2]*v2[1]; \
	dest[1]=v1[2]*v2[0]-v1[0]*v2[2]; \
	dest[2]=v1[0]*v2[1]-v1[1]*v2[0];
#define DOT(v1,v2) (v1[0]*v2[0]+v1[1]*v2[1]+v1[2]*v2[2])
#define SUB(dest,v1,v2) \
	dest[0]=v1[0]-v2[0]; \
	dest[1]=v1[1]-v2[1]; \
	dest[2]=v1[2]-v2[2];
#define NORMALIZE(A)    {float l=1/sqrtf(A[0]*A[0]+A[1]*A[1]+A[2]*A[2]);A[0]*=l;A[1]*=l;A[2]*=l;}

/* main.cu.h:60-63 */
extern float dd_camcoords[16 * 4];
extern float dd_light_position[6];

/* main.cu.h:58 + per_frame_funcs.h:427-433: 5x5 float4 direction texture, normalised coordinates,
 * clamp addressing, linear filtering.  tex2D follows the CUDA Programming Guide's linear-filter
 * formula (xB = x*N - 0.5, i = floor(xB), alpha = frac(xB)); `g_ref_tex_q8` selects the hardware's
 * 8-bit fixed-point weights (1) or exact float weights (0, default - see DESIGN.md "direction lattice"). */
struct ref_texture { float4 texel[5 * 5]; };
extern ref_texture texdir;
extern int g_ref_tex_q8;
float4 tex2D(const ref_texture &tex, float s, float t);

extern int sharedmem[];

/* cutil / runtime stubs used by camera.h */
#define cutilSafeCall(x) (x)
static inline int cudaMalloc(void **p, size_t n) { *p = malloc(n); return 0; }
static inline int cudaFree(void *p) { free(p); return 0; }

/* GL/GLU stubs used by camera.h:86-148 (restated per the GL/GLU specification; the values the
 * reference got came from the GL driver - "parity unpinned" for these two matrices). */
#define GL_MODELVIEW_MATRIX 0x0BA6
#define GL_PROJECTION_MATRIX 0x0BA7
#define GL_MODELVIEW 0x1700
#define GL_PROJECTION 0x1701
void glGetFloatv(int which, float *out);
void glMatrixMode(int mode);
void glLoadIdentity(void);
void gluPerspective(double fovy, double aspect, double zNear, double zFar);
void gluLookAt(double ex, double ey, double ez, double cx, double cy, double cz, double ux, double uy, double uz);

#endif
