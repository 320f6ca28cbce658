/* ug_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C11) of the per-frame hot path of sushruta/uniformgrid-raytracing, used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER.  Never linked into
 * or called from the product library (libugrt.so); the product fails loudly without its CUDA code.
 *
 * Pinning: the reference holds no golden vectors or tests for this path, so the oracle
 * is pinned against the reference's own kernel sources compiled as host C++ (oracle/_ref/libugref.so,
 * recipe oracle/Makefile) on seeded scenes - tests/test_oracle_vs_ref.py - and against fixtures that
 * library generated (tests/golden/, script tests/golden/make_golden.py).
 * Row R (reflection) has no reference implementation: parity unpinned, see DESIGN.md.
 *
 * The API mirrors oracle/ref/ref_driver.cpp one to one (prefix orc_ instead of ref_) so the Python
 * harness drives both through the same code.  State is global like the reference's (main.cu.h:58-63).
 */
#ifndef UG_ORACLE_H
#define UG_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

void orc_configure(int W, int H, int fovy, int tex_q8);
void orc_set_slabs(int n); /* NUM_SLABS (main.cu.h:18), default 1: depth slabs of every grid, SURVEY 8f-3 */
void orc_set_threads(int n);
int orc_get_threads(void);
void orc_camera_setup(const float *pose9, float nearP, float farP, int square_aspect, float *out_camcoords);
void orc_set_camcoords(const float *cc);
void orc_get_dirtex(float *out100);
void orc_set_lights(const float *lights, int n);
void orc_transform(float *vertexlist, float *orig_list, int size, int offset, float rot);
void orc_grid_count(int *faces, float *verts, int T, unsigned *sizeList, float *projZ);
long orc_build_grid(int *faces, float *verts, int T, unsigned *span, unsigned *offset, unsigned *key_list,
                    unsigned *value_list, long cap);
/* a14/a15: the spherical light grid display() uses (grid_kernel.cu:395-479,481-854; misc_kernel.cu:255-296) */
void orc_grid_count_spherical(int *faces, float *verts, int T, unsigned *sizeList, float *projZ, float xM, float yM);
long orc_build_grid_spherical(int *faces, float *verts, int T, unsigned *span, unsigned *offset, unsigned *key_list,
                              unsigned *value_list, long cap, float xM, float yM);
void orc_map_rays_spherical(float *t_list, float *dir_list, unsigned *map, float *cmPt, float xM, float yM);
void orc_acosf_array(const float *x, float *y, long n); /* libm acosf, the reference's */
void orc_trace(unsigned *value_list, unsigned *span, unsigned *offset, float *normal, float *t, float *dir,
               int *shadowed, int *id, float *verts, int *tris);
void orc_map_rays(float *t_list, float *dir_list, unsigned *map, float *cmPt, float nearPlane);
void orc_shadow(unsigned *value_list, float *verts, int *tris, unsigned *span, unsigned *offset, float *t_list,
                float *dir_list, int *is_shadowed, unsigned *map, float *cmPt);
void orc_shade(unsigned char *img, float *normal, float *t, float *dir, int *id, float *cam_pos, int *mat_idx,
               float *mat_list, int mat_count);
void orc_spot_shade(unsigned char *img, float *normal, float *t, float *dir, int *id, float *cam_pos, int *mat_idx,
                    float *mat_list, int mat_count, float *dump);
void orc_add_shadows(unsigned char *img, int *is_shadowed);
/* row R (reflection): authored spec, parity unpinned - see ug_oracle.c.  `rgb` holds the shaded + shadowed frame on
 * entry and the final frame on return; g_cc / g_light must be what was resident when orc_shade ran. */
void orc_reflect(float *verts, int *tris, int T, float *t, float *dir, int *id, unsigned char *rgb, float *cam_pos, int *mat_idx, float *mat_list,
                 int mat_count, float *mat_reflect, int bounces);
int orc_load_obj(char *path, int *num_vertices, int *num_faces, float *vertexlist, int *facelist, int *matidx);
int orc_load_materials(char *file, int *num_materials, float *materiallist);

#ifdef __cplusplus
}
#endif
#endif
