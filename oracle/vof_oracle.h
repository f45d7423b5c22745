/* TEST INFRASTRUCTURE ONLY (oracle/).  Not part of the product: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load libvoforacle.so.
 *
 * Plain-C restatement of dune-vof's per-timestep hot path on an axis-aligned Cartesian
 * grid.  Every function cites the reference file:line it restates.  Pinned bit-for-bit against the reference's
 * own templates (oracle/_ref/libvofref.so, built from /root/reference) by tests/test_oracle_vs_reference.py and
 * against the golden vectors in tests/golden.  The only third-party arithmetic (dune-common FieldMatrix::solve,
 * dune-spgrid cell coordinates) is restated from the published closed forms: parity is UNPINNED at those two
 * boundaries.
 *
 * Conventions: cell index x-fastest; cell lower corner = origin + i*h; centre = lower + 0.5*h;
 * faces ordered x-,x+,y-,y+,z-,z+; half-space stored as (n[dim], d), fluid side n.x + d > 0;
 * flags 0 empty, 1 mixed, 2 mixedfull, 3 full, 99 nan (flagset.hh:18-24).
 */
#ifndef VOF_ORACLE_H
#define VOF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
  int dim;
  int n[3];
  double origin[3];
  double h[3];
} vo_grid;

enum { VO_RECON_YOUNGS = 0, VO_RECON_HF = 1, VO_RECON_SWARTZ = 2 };
enum { VO_PROB_ROTATING_CIRCLE = 0, VO_PROB_SFLOW = 1, VO_PROB_SLOTTED_CYLINDER = 2, VO_PROB_LINEAR_WALL = 3, VO_PROB_LEVEQUE = 4 };

/* operators */
int vo_flag ( const vo_grid *g, const double *u, double eps, uint8_t *flags );
int vo_reconstruct ( const vo_grid *g, int kind, const double *u, const uint8_t *flags, double *hs, int max_iterations );
int vo_evolve ( const vo_grid *g, const double *hs, const uint8_t *flags, int problem, double time, double dt, double *update, double *dt_est );
int vo_curvature ( const vo_grid *g, const double *u, const double *hs, const uint8_t *flags, double *kappa, uint8_t *valid );
int vo_curvature_swartz ( const vo_grid *g, const double *hs, const uint8_t *flags, double *kappa );   /* 2D; curvature/swartzcurvature.hh:41-173 */
int vo_init ( const vo_grid *g, int problem, double time, double *u );
int vo_circle_interpolation ( const vo_grid *g, const double *center, double radius, double *u );   /* interpolation/circleinterpolation.hh:140-153 */
int vo_l1error_circle ( const vo_grid *g, const double *hs, const uint8_t *flags, const double *center, double radius, double *error );   /* test/errors.hh:62-95 */
int vo_face_velocity ( const vo_grid *g, int problem, double time, int64_t cell, int face, double *v );
int vo_run ( const vo_grid *g, int kind, int problem, double eps, double cfl, int max_iterations, int passes,
             double *u, double *time_io, double *dt_io, double *phase_seconds, int64_t *mixed_cell_steps );

/* geometry primitives on one axis-aligned cell */
int vo_locate ( int dim, const double *lo, const double *h, const double *normal, double fraction, double *out );
int vo_volume_fraction ( int dim, const double *lo, const double *h, const double *hs, double *out );
int vo_flux ( int dim, const double *lo, const double *h, int face, const double *v, const double *hs, double *out );
int vo_interface ( int dim, const double *lo, const double *h, const double *hs, int *nverts, double *verts, double *centroid, double *measure );
double vo_brent_cubic ( double A, double B, double C, double target, double a, double b, double tol );
double vo_det_cospi ( double s );

/* ---- unstructured meshes as flattened arrays: triangles and quadrilaterals (dim 2), tetrahedra and hexahedra (dim 3);
 *      the layout of include/vofx.h's vofx_mesh_desc.  All geometry of the grid (centres, volumes, normals) is DATA
 *      supplied by the grid side; D below is dim. ---- */
typedef struct
{
  int64_t n_cells;
  const int32_t *corner_offsets;     /* [n_cells+1] */
  const double *corners;             /* [..][D] cell corners in DUNE corner order */
  const double *centers;             /* [n_cells][D] */
  const double *volumes;             /* [n_cells] */
  const int32_t *face_offsets;       /* [n_cells+1] intersections in iteration order */
  const int32_t *face_neighbor;      /* outside cell or -1 */
  const double *face_corners;        /* [..][2][2]; dim 3: 3 or 4 corners of D doubles per face */
  const double *face_centers;        /* [..][2] */
  const double *face_unit_normals;   /* [..][2] centerUnitOuterNormal */
  const double *face_int_normals;    /* [..][2] integrationOuterNormal at the centre */
  const double *face_volumes;        /* [..] */
  const int32_t *stencil_offsets;    /* [n_cells+1] CSR vertex-neighbour stencil, */
  const int32_t *stencil;            /*             ascending cell index, self excluded */
  int32_t dim;                       /* 0 or 2: two dimensions; 3: three */
  const int32_t *face_corner_offsets;/* dim 3: [n_faces+1] positions (in corners) in face_corners; dim 2: unused (2 each) */
} vo_mesh;

int vo_mesh_flag ( const vo_mesh *m, const double *u, double eps, uint8_t *flags );
int vo_mesh_youngs ( const vo_mesh *m, const double *u, const uint8_t *flags, double *hs );
int vo_mesh_swartz ( const vo_mesh *m, const double *u, const uint8_t *flags, double *hs, int max_iterations );   /* 2D, modifiedswartz.hh:70-153 */
int vo_mesh_evolve ( const vo_mesh *m, const double *hs, const uint8_t *flags, const double *face_velocity, double dt, double *update, double *dt_est );
int vo_mesh_locate ( const vo_mesh *m, int64_t cell, const double *normal, double fraction, double *out );
/* 3D primitives on explicit corner lists: cell with 4 (tetrahedron) or 8 (hexahedron) corners; intersection with 3 or 4 */
int vo_locate_corners ( int nc, const double *corners, const double *normal, double fraction, double *out );
int vo_flux_corners ( int nc, const double *corners, const double *v, const double *hs, double *out );

#ifdef __cplusplus
}
#endif

#endif /* VOF_ORACLE_H */
