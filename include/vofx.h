/* vofx -- C ABI of the B200-native volume-of-fluid step.
 *
 * This is the drop-in boundary for dune-vof's per-timestep hot path.  Each entry point
 * names the reference interface it replaces (paths relative to the reference's dune/vof/).  The header adaptor in
 * include/dune/vofx/ wraps these calls in classes with the reference's operator names and signatures;
 * INTEGRATION.md shows the binding a dune-vof maintainer would add.
 *
 * Plain C: pointers and sizes only, no C++/torch types.  All functions return 0 on success and a non-zero status
 * otherwise (never throw across the boundary); vofx_last_error() describes the last failure of the calling thread.
 * There is NO CPU fallback: every entry point needs a CUDA device and fails with VOFX_ERR_CUDA without one.
 *
 * Grid convention (defined here so that host and device generate bit-identical coordinates):
 *   structured Cartesian grid, global cell multi-index (i,j,k), cell index x-fastest;
 *   cell lower corner x_lo[d] = origin[d] + i[d]*h[d]; upper corner x_lo[d] + h[d]; centre x_lo[d] + 0.5*h[d];
 *   faces are numbered x-,x+,y-,y+,z-,z+ (dune-grid's cube numbering).
 * Slab decomposition (multi-GPU): a context owns the cells [lo, lo+n_local) and stores `halo` extra layers on
 *   either side along the LAST axis only; every per-cell array passed to a context has
 *   n_local[0] * n_local[1] * (n_local[2] + 2*halo) entries in 3D (n_local[0] * (n_local[1] + 2*halo) in 2D).
 *   With halo == 0 and lo == 0, n_local == n_global this is the plain single-GPU layout.
 *
 * Data layout:
 *   colour u       double[cells]            (dune-vof ColorFunction, colorfunction.hh:20-52)
 *   flags          uint8_t[cells]           values of enum Flag (flagset.hh:18-24): 0 empty, 1 mixed, 2 mixedfull, 3 full, 99 nan
 *   half-spaces    double[cells*(dim+1)]    per cell (n[0..dim), d); fluid side n.x + d > 0 (geometry/halfspace.hh:32-112);
 *                                           the zero normal means "no reconstruction"
 *   curvature      double[cells], validity uint8_t[cells]
 */
#ifndef VOFX_H
#define VOFX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VOFX_VERSION 1

/* Tunables read from the environment at context creation / launch (defaults are the measured optima on B200):
 *   VOFX_LANES_PER_CELL = 4 | 8   lanes per mixed cell of the 3D Young's / height-function kernel (default 4)
 *   VOFX_FLUX_LANES     = 8 | 4   lanes per mixed cell of the 3D flux kernel (default 8)
 *   VOFX_SWARTZ_LANES   = 4 | 8   lanes per task of the 3D Swartz kernels (default 4)
 *   VOFX_SWARTZ_MODE    = batched (default) | cell | serial   formulation of the 3D Swartz reconstruction
 *   VOFX_YOUNGS_BLOCKS_PER_SM = n   grid of the 3D plane-constant kernel in 64-thread blocks per SM (default 60)
 *   VOFX_FLUX_BLOCKS_PER_SM, VOFX_UPDATE_BLOCKS_PER_SM = n   grids of the 3D flux / update kernels (default 8)
 *   VOFX_SWZ_BLOCKS_PER_SM, VOFX_SWZ_OTHER_BLOCKS_PER_SM = n   grids of the batched Swartz kernels (default 80 / 48)
 *   VOFX_SWZ_NO_SORT    = 1   batched 3D Swartz: do not sort the (cell, neighbour) tasks by last iteration's bracket
 *   VOFX_NO_COST_ORDER  = 1   process the band in list order instead of last step's expensive cells first
 *   VOFX_BAND_PARTS     = 1 .. 4   chains ( plane-constant solve -> root search -> fluxes ) a 3D Young's / height-function pass runs
 *                                  on as many streams (default 2; joined before the update)
 *   VOFX_PDL            = 1   launch the kernels of a pass with programmatic dependent launch (measured slower: off by default)
 *   VOFX_NO_GRAPH       = 1   vofx_mesh_run issues its kernels one by one instead of replaying a CUDA graph
 * None of them changes a result. */

enum vofx_status
{
  VOFX_OK = 0,
  VOFX_ERR_ARGUMENT = 1,   /* bad descriptor / null pointer / unsupported combination */
  VOFX_ERR_CUDA = 2,       /* CUDA runtime error, or no device */
  VOFX_ERR_STATE = 3       /* call sequence error (e.g. step without velocity) */
};

enum vofx_flag_value { VOFX_FLAG_EMPTY = 0, VOFX_FLAG_MIXED = 1, VOFX_FLAG_MIXEDFULL = 2, VOFX_FLAG_FULL = 3, VOFX_FLAG_NAN = 99 };

/* reconstruction operators of the reference */
enum vofx_reconstruction
{
  VOFX_RECON_YOUNGS = 0,   /* ModifiedYoungsReconstruction,              reconstruction/modifiedyoungs.hh:63-134 */
  VOFX_RECON_HF = 1,       /* HeightFunctionReconstruction< Youngs >,    reconstruction/heightfunction.hh:70-259 (the default factory, reconstruction.hh:29-37) */
  VOFX_RECON_SWARTZ = 2    /* ModifiedSwartzReconstruction< Youngs >,    reconstruction/modifiedswartz.hh:70-243 */
};

typedef struct
{
  int32_t dim;             /* 2 or 3 */
  int32_t n_global[3];     /* cells per axis of the whole grid */
  double origin[3];        /* lower corner of the whole grid */
  double h[3];             /* mesh widths */
  int32_t lo[3];           /* global multi-index of the first owned cell */
  int32_t n_local[3];      /* owned cells per axis */
  int32_t halo;            /* halo layers stored along the last axis (0 on a single GPU) */
} vofx_grid_desc;

typedef struct vofx_ctx vofx_ctx;

const char *vofx_last_error ( void );

/* replaces: construction of the operators + VertexNeighborsStencil in Algorithm's constructor (algorithm.hh:46-51).
 * device < 0 selects the current device.  The context owns device scratch (mixed-cell list, flux records, tables);
 * callers own all per-cell arrays. */
int vofx_create ( const vofx_grid_desc *grid, int device, vofx_ctx **out );
void vofx_destroy ( vofx_ctx *ctx );

/* run all work of this context on the given cudaStream_t (default: the legacy default stream) */
int vofx_set_stream ( vofx_ctx *ctx, void *cuda_stream );

/* number of cells of a per-cell array of this context (including halo layers) */
int64_t vofx_cells ( const vofx_ctx *ctx );

/* ---- velocity (replaces Velocity< Problem, GridView >, velocity.hh:5-33, evaluated at face centres) ------------
 * Separable description: component a at a face centre x and time t is
 *     v_a = coef_a * c_a(t) * sum_{r < nterms_a} ( f_{a,r,0}(x_0) * f_{a,r,1}(x_1) ) * f_{a,r,2}(x_2)
 * (2D: the product of two factors) with the products and sums associated exactly as written, left to right.
 * Every factor is a table over the cells of its axis, sampled at the three coordinates a face centre can have
 * in that axis when seen from cell i: the cell centre, its lower face and its upper face (n_global[axis] entries
 * each, HOST pointers, copied).  c_a(t) is 1 (VOFX_TIME_CONST) or cos(pi t/period) (VOFX_TIME_COSPI) evaluated by
 * the fixed polynomial documented in DESIGN.md, so that time can live on the device.
 * All dune-vof test problems (rigid rotations, SFlow, LinearWall) and the LeVeque fields are of this form.       */
enum vofx_time_factor { VOFX_TIME_CONST = 0, VOFX_TIME_COSPI = 1 };

int vofx_velocity_clear ( vofx_ctx *ctx );
int vofx_velocity_set_component ( vofx_ctx *ctx, int component, int nterms, double coef, int time_factor, double period );
int vofx_velocity_set_factor ( vofx_ctx *ctx, int component, int term, int axis,
                               const double *at_centre, const double *at_lower_face, const double *at_upper_face );

/* general fallback for an arbitrary host velocity functor: the caller samples it at every face centre (as
 * Velocity::operator() would, velocity.hh:15-20, frozen at the start-of-step time) and passes
 * v_host[ ( cell * 2*dim + face ) * dim + component ] over all cells of the context (HOST pointer, copied).
 * While set it overrides the separable description; vofx_velocity_clear() removes it. */
int vofx_velocity_set_faces ( vofx_ctx *ctx, const double *v_host );

/* built-in problems (tables generated on the host with libm from the coordinates defined above):
 * 0 RotatingCircle (test/problems/rotatingcircle.hh:45-52,92-101), 1 SFlow (sflow.hh:35-42,73-82),
 * 2 SlottedCylinder (slottedcylinder.hh:47-55, 2D), 3 LinearWall (linearwall.hh:27-31), 4 LeVeque deformation field */
enum vofx_problem { VOFX_PROB_ROTATING_CIRCLE = 0, VOFX_PROB_SFLOW = 1, VOFX_PROB_SLOTTED_CYLINDER = 2, VOFX_PROB_LINEAR_WALL = 3, VOFX_PROB_LEVEQUE = 4 };
int vofx_velocity_builtin ( vofx_ctx *ctx, int problem );

/* velocity at the centre of `face` of the cell with LOCAL storage index `cell` at time t (for tests) */
int vofx_face_velocity ( vofx_ctx *ctx, double time, int64_t cell, int face, double *v_host );

/* ---- operators on DEVICE arrays ---------------------------------------------------------------------------------*/

/* replaces FlagOperator::operator()( color, flags ), flagging.hh:35-70 */
int vofx_flag ( vofx_ctx *ctx, const double *u, double eps, uint8_t *flags );

/* replaces {ModifiedYoungs,HeightFunction,ModifiedSwartz}Reconstruction::operator()( color, reconstructions, flags ):
 * clears `halfspaces`, then reconstructs every mixed cell.  max_iterations: Swartz only (reference default 10). */
int vofx_reconstruct ( vofx_ctx *ctx, int kind, const double *u, const uint8_t *flags, double *halfspaces, int max_iterations );

/* replaces Evolution::operator()( reconstructions, flags, velocity, deltaT, update ), evolution/evolution.hh:58-76:
 * clears `update`, accumulates the geometric fluxes, returns the time-step estimate (min over mixed cells) in *dt_est
 * (host pointer). */
int vofx_evolve ( vofx_ctx *ctx, const double *halfspaces, const uint8_t *flags, double time, double dt, double *update, double *dt_est );

/* replaces ColorFunction::axpy( a, x ), colorfunction.hh:31-37 */
int vofx_axpy ( vofx_ctx *ctx, double *u, double a, const double *x );

/* replaces CartesianHeightFunctionCurvature::operator()( color, reconstructions, flags, curvature ),
 * curvature/cartesianheightfunctioncurvature.hh:54-87.  `valid` (may be NULL) receives satisfiesConstraint_. */
int vofx_curvature ( vofx_ctx *ctx, const double *u, const double *halfspaces, const uint8_t *flags, double *kappa, uint8_t *valid );

/* ---- fused time loop on a DEVICE-resident colour function (replaces the loop body of Algorithm::operator(),
 * algorithm.hh:68-95, without l1error/output) -------------------------------------------------------------------
 * State (time, dt) lives on the device; no host synchronisation per step.  vofx_step_begin sets (time, dt = 0):
 * as in the reference the first step is a dry run that only produces the time-step estimate. */
typedef struct
{
  int32_t reconstruction;   /* enum vofx_reconstruction */
  int32_t max_iterations;   /* Swartz */
  double eps;               /* flag tolerance (scheme.eps) */
  double cfl;               /* dt = cfl * dtEst */
  int32_t with_curvature;   /* also evaluate the height-function curvature each step (into kappa) */
} vofx_step_params;

int vofx_step_begin ( vofx_ctx *ctx, double time, double dt );
/* Incremental flagging for loops on a colour function that ONLY the loop passes of this context change between two
 * passes (plus the halo layers, refilled by the caller's exchange): flags can then only change within two face steps
 * of the previous pass's mixed cells (flagging.hh:46-63 reads a cell and its face neighbours; the update writes the
 * mixed cells and their face neighbours), so the dense pass over all cells is replaced by a re-flag of that
 * neighbourhood.  The first pass after vofx_step_begin / vofx_color_upload / an operator-level call, or on other
 * arrays than the previous pass, is dense.  Flags and band list are identical to the dense pass's (as sets; the band
 * list has no defined order in either).  vofx_step_resident enables it by itself; vofx_step_host never uses it.
 * Costs five bytes per 32 cells of device memory. */
int vofx_set_incremental_flagging ( vofx_ctx *ctx, int enable );
/* one loop pass: flag -> reconstruct -> evolve -> u += update -> time += dt -> dt = cfl*dtEst.
 * flags / halfspaces / kappa are outputs (kappa may be NULL unless with_curvature). */
int vofx_step ( vofx_ctx *ctx, const vofx_step_params *params, double *u, uint8_t *flags, double *halfspaces, double *kappa );
/* copy the device-resident (time, dt, last dtEst, number of mixed cells of the last step) to the host; synchronises */
int vofx_step_state ( vofx_ctx *ctx, double *time, double *dt, double *dt_est, int64_t *mixed_cells );
/* multi-GPU: address of the device scalar holding the last local dtEst, for a MIN all-reduce between
 * vofx_step_local() and vofx_step_finish() (the "nccl" transport of dune-vof_b200/decomposition.py) */
int vofx_step_local ( vofx_ctx *ctx, const vofx_step_params *params, double *u, uint8_t *flags, double *halfspaces, double *kappa );
int vofx_step_finish ( vofx_ctx *ctx, const vofx_step_params *params );
/* vofx_step_local in three phases, so that the halo exchange (during phase 0) and the MIN-reduction of the time-step
 * estimate (during phase 2) overlap with the kernels: 0 flags of the interior layers, 1 flags next to the halos +
 * reconstruction + fluxes, 2 update.  Phases 0,1,2 in this order are equivalent to vofx_step_local. */
int vofx_step_phase ( vofx_ctx *ctx, const vofx_step_params *params, double *u, uint8_t *flags, double *halfspaces, double *kappa, int phase );
double *vofx_dt_est_device ( vofx_ctx *ctx );

/* ---- HOST-buffer variants: what the DUNE-side containers (std::vector based DataSet, dataset.hh:26-87) bind to.
 * They stage through pinned memory, copy to the device, run the operator and copy the results back. */
int vofx_flag_host ( vofx_ctx *ctx, const double *u, double eps, uint8_t *flags );
int vofx_reconstruct_host ( vofx_ctx *ctx, int kind, const double *u, const uint8_t *flags, double *halfspaces, int max_iterations );
int vofx_evolve_host ( vofx_ctx *ctx, const double *halfspaces, const uint8_t *flags, double time, double dt, double *update, double *dt_est );
int vofx_curvature_host ( vofx_ctx *ctx, const double *u, const double *halfspaces, const uint8_t *flags, double *kappa, uint8_t *valid );
/* replaces: SwartzCurvature< GV, StS >( stencils )( color, reconstructions, flags, curvature ), 2D only like the reference
 * (curvature/swartzcurvature.hh:41-173); kappa is zero outside the mixed cells */
int vofx_curvature_swartz ( vofx_ctx *ctx, const double *halfspaces, const uint8_t *flags, double *kappa );
int vofx_curvature_swartz_host ( vofx_ctx *ctx, const double *halfspaces, const uint8_t *flags, double *kappa );
/* one full loop pass on a host colour function: H2D u, step, D2H u (and flags if non-NULL) */
int vofx_step_host ( vofx_ctx *ctx, const vofx_step_params *params, double *u, uint8_t *flags );
/* The same pass for a driver that keeps its colour function on the HOST and says what it did to it: `touched` lists the
 * n_touched cells (storage indices) the caller has changed in `u` since the previous vofx_step_host / vofx_step_host_sparse
 * call on the same array (none for a plain time loop; boundary or source cells otherwise).  The device mirror left by that
 * call is then current elsewhere: only the listed values cross PCIe host -> device, and, as in vofx_step_host, only the
 * (cell, new value) pairs of the cells the update changed come back.  The first call on an array, or a call after the
 * mirror was used by another entry point, uploads the whole array.  With n_touched == 0 flags are maintained incrementally
 * (vofx_set_incremental_flagging); otherwise the pass flags densely. */
int vofx_step_host_sparse ( vofx_ctx *ctx, const vofx_step_params *params, double *u, uint8_t *flags, const int32_t *touched, int64_t n_touched );
/* bytes the last vofx_step_host call moved across PCIe: the whole colour function host->device; device->host only the
 * (cell, new value) pairs of the cells the update touched (mixed cells and their face neighbours) plus the loop state */
int vofx_host_transfer_bytes ( const vofx_ctx *ctx, int64_t *h2d_bytes, int64_t *d2h_bytes );

/* sparse write-back for drivers that run the pass themselves on DEVICE arrays (slab decomposition): after
 * vofx_track_changes( ctx, 1 ) every update records the (cell, new value) pairs it writes; vofx_changed_fetch copies
 * the pairs of the last finished pass into a HOST colour function: u_host[ storage index + index_offset ] = value. */
int vofx_track_changes ( vofx_ctx *ctx, int enable );
int vofx_changed_fetch ( vofx_ctx *ctx, double *u_host, int64_t index_offset, int64_t *n_changed );

/* context-resident colour function (what a drop-in for Algorithm::operator() uses, algorithm.hh:54-103): upload the
 * host colour function once, run loop passes on the device mirror, download when the data writer wants it */
int vofx_color_upload ( vofx_ctx *ctx, const double *u_host );
int vofx_step_resident ( vofx_ctx *ctx, const vofx_step_params *params );
int vofx_color_download ( vofx_ctx *ctx, double *u_host, uint8_t *flags_host /* may be NULL */ );

/* ---- checkpoint / restart of the context-resident colour function in the byte format of the reference's BinaryWriter
 * (example/binarywriter.hh:36-60: `binaryStream << time; uh.write( binaryStream )`, colorfunction.hh:40-51): the time, then
 * one double per OWNED cell in x-fastest order (the iteration order of a structured grid view), raw little-endian doubles.
 * bin2vtk / a restarted example/vof.cc read these files.  vofx_checkpoint_name builds the writer's file name
 * "s<size:04>-p<rank:04>-<prefix>-<level>-<step:05>.bin" (binarywriter.hh:46-51).  After a read the halos of a distributed
 * run must be refilled (vofx_halo_push) and the loop restarted with vofx_step_begin( time, dt ). */
int vofx_checkpoint_name ( char *out, int64_t bytes, const char *prefix, int level, int step, int rank, int world );
int vofx_checkpoint_write ( vofx_ctx *ctx, const char *path, double time );
int vofx_checkpoint_read ( vofx_ctx *ctx, const char *path, double *time );

/* ---- multi-GPU: the grid is cut into slabs along the last axis, one context per GPU (one per process -- one per MPI
 * rank of a DUNE host -- or several in one process).  Replaces the reference's exchanges over GridView::communicate
 * (dataset.hh:69-81; call sites flagging.hh:69, modifiedyoungs.hh:75, evolution/evolution.hh:73) and gridView().comm().min
 * (evolution.hh:75) by stores into the neighbours' memory over NVLink, issued by the kernels themselves:
 *   - the colour function is context-resident (vofx_color_device / vofx_color_upload) inside a "window" whose CUDA IPC
 *     handle is handed to the other ranks; it only ever changes in the update kernel, which also writes every new value
 *     of a boundary layer into the neighbour's halo copy -- there is no halo exchange after the start of a loop;
 *   - flags and reconstructions of the first halo layer are recomputed locally (owner-computes);
 *   - the time-step estimates travel as one double per rank into every rank's window; passes are ordered by epoch
 *     counters in the same windows (a wait that times out after ~10 s makes the next vofx_step_state fail).
 * Call sequence on every rank: vofx_create (lo / n_local / halo of the slab; halo = 2 Young's, 4 height functions or
 * curvature, 3 Swartz) -> vofx_comm_export -> [the host all-gathers the handles: MPI_Allgather, torch.distributed, a
 * pipe] -> vofx_comm_connect -> velocity, initial data (vofx_color_upload or vofx_init on vofx_color_device) ->
 * vofx_step_begin -> [host barrier] -> vofx_halo_push -> [host barrier] -> vofx_step_distributed ... (no host
 * communication per pass; the pass is graph-capturable) -> vofx_color_download.
 * Results are bit-identical to the single-context run: coordinates come from global indices, MIN is exact. */
#define VOFX_COMM_HANDLE_BYTES 80
/* allocates the window (must precede any use of the context-resident colour function) and returns its handle: the
 * cudaIpcMemHandle_t, the exporting process's id and the window's address there (ranks that are threads of one process
 * are connected through the address; peer access between their devices must be possible) */
int vofx_comm_export ( vofx_ctx *ctx, void *handle /* VOFX_COMM_HANDLE_BYTES */ );
/* layers[ r ] = owned layers of rank r along the last axis; handles = world * VOFX_COMM_HANDLE_BYTES bytes in rank order
 * (the entry of `rank` itself is ignored) */
int vofx_comm_connect ( vofx_ctx *ctx, int rank, int world, const int32_t *layers, const void *handles );
/* the same for `world` contexts of ONE process in slab order (same or different devices; peer access is enabled) */
int vofx_comm_connect_local ( vofx_ctx **contexts, int world );
/* unmaps the other ranks' windows.  Every rank calls it, then the host synchronises the ranks (barrier), THEN the contexts
 * are destroyed: an exported window must not be freed while another process still maps it */
int vofx_comm_disconnect ( vofx_ctx *ctx );
/* device addresses of the context-resident colour function / flags / half-spaces (storage size incl. halos) */
double *vofx_color_device ( vofx_ctx *ctx );
uint8_t *vofx_flags_device ( vofx_ctx *ctx );
double *vofx_halfspaces_device ( vofx_ctx *ctx );
/* copies the owned boundary layers into the neighbours' halos (after the initial data is in place); synchronises */
int vofx_halo_push ( vofx_ctx *ctx );
/* one loop pass of the distributed run on the context-resident colour function (vofx_step_resident when not connected) */
int vofx_step_distributed ( vofx_ctx *ctx, const vofx_step_params *params );
/* the same pass in two calls: phase 0 = flags and band list, phase 1 = everything else.  In between a driver whose
 * velocity is an arbitrary HOST functor (Velocity< Problem, GridView >, velocity.hh:5-33) samples it on the faces of the
 * band only: vofx_band_list_fetch_in_flight returns the storage indices of the pass's mixed cells, and
 * vofx_velocity_set_band_faces takes v_host[ ( i * 2*dim + face ) * dim + component ] for the i-th of them (HOST
 * pointer, copied) -- 6 x 3 doubles per mixed cell instead of per cell of the grid (vofx_velocity_set_faces). */
int vofx_step_distributed_phase ( vofx_ctx *ctx, const vofx_step_params *params, int phase );
int vofx_band_list_fetch_in_flight ( vofx_ctx *ctx, int32_t *cells_host, int64_t capacity, int64_t *n_cells );
int vofx_velocity_set_band_faces ( vofx_ctx *ctx, int64_t n_cells, const double *v_host );

/* ---- interface extraction (replaces getInterfaceVertices, utility.hh:19-48, and MixedCellMapper::update,
 * mixedcellmapper.hh:82-86): the owned mixed cells in ascending cell index, and for the i-th of them the vertices of
 * interface( entity, reconstructions ) (geometry/polytope.hh:69-75: a segment in 2D, a polygon with 3..6 vertices in
 * 3D) as a CSR ( offsets[ i ] .. offsets[ i+1 ] ).  vofx_interface works on DEVICE arrays and returns the sizes;
 * vofx_interface_fetch copies cells[ n_mixed ], offsets[ n_mixed + 1 ] and vertices[ n_vertices * dim ] to the HOST
 * (the InterfaceGrid / bin2vtk consumers, interfacegrid/dataset.hh:114-120, bin2vtk.cc:264-281, live there). */
int vofx_interface ( vofx_ctx *ctx, const double *halfspaces, const uint8_t *flags, int64_t *n_mixed, int64_t *n_vertices );
int vofx_interface_host ( vofx_ctx *ctx, const double *halfspaces_host, const uint8_t *flags_host, int64_t *n_mixed, int64_t *n_vertices );
int vofx_interface_fetch ( vofx_ctx *ctx, int32_t *cells_host, int64_t *offsets_host, double *vertices_host );

/* ---- diagnostics of the colour function over the OWNED cells, reduced on the device (keeps a driver's conservation /
 * EOC reporting without a full-field copy per step): out_host[0] = sum of the cell values, [1] = min, [2] = max,
 * [3] = cellwiseL1error (test/errors.hh:138-151: sum |uh - uhExact| * volume; 0 when u_exact is NULL),
 * [4] = mass = sum uh * volume.  u, u_exact: DEVICE arrays of the context's storage size; out_host: 5 HOST doubles
 * (the call synchronises the context's stream).  Fixed summation tree: reproducible, equal to the reference's
 * sequential sums up to round-off (1e-12 relative in the tests), min / max exact. */
int vofx_diagnostics ( vofx_ctx *ctx, const double *u, const double *u_exact /* may be NULL */, double *out_host );

/* ---- initial data on the device (replaces Average< Problem >, test/average.hh:28-54, i.e.
 * RecursiveInterpolationCube depth 3, interpolation/recursiveinterpolationcube.hh:33-82, for the built-in problems) */
int vofx_init ( vofx_ctx *ctx, int problem, double time, double *u );
/* circleInterpolation( center, radius, uh ), interpolation/circleinterpolation.hh:75-153 (2D): the exact area fraction of the
 * disc in every cell -- the initialiser of RotatingCircle< double, 2 > (test/average.hh:82-101, configs[0]).  asin / sin are
 * CUDA's: cells cut by the circle agree with the host's glibc to ~1e-16, all others exactly. */
int vofx_init_circle ( vofx_ctx *ctx, const double *center /* [2], host */, double radius, double *u );
/* l1error( gridView, reconstructions, flags, RotatingCircle< double, 2 >, time ), test/errors.hh:62-95, summed over the owned
 * cells on the device (fixed summation tree): |exact disc /\ cell  symmetric-difference  reconstructed fluid part|.
 * vofx_l1error_circle reads the flags / reconstructions the last pass of the context-resident loop left behind (what
 * Algorithm::operator() accumulates as dt * l1error, algorithm.hh:85-89); the _device variant takes DEVICE arrays. */
int vofx_l1error_circle ( vofx_ctx *ctx, const double *center, double radius, double *error_host );
int vofx_l1error_circle_device ( vofx_ctx *ctx, const double *halfspaces, const uint8_t *flags, const double *center, double radius, double *error_host );

/* ---- geometry primitives on single cells (DEVICE work on small batches; for parity tests) ---------------------
 * arrays are HOST pointers of `count` items: lo[count*dim], h[count*dim], ... */
/* measured fp64 instruction rates of the current GPU, in 1e9 operations/s: rates[0] fused multiply-add (one operation
 * each), rates[1] separate multiply and add (the parity contract forbids FMA contraction); the denominators of the
 * band kernels' fp64 rooflines */
int vofx_test_fp64_rate ( double *rates );
/* the device's straight-line IEEE division / square root (vofx_math.cuh: vdiv, vdiv3 / vdiv2, vsqrt) against the native
 * operators on `count` pseudo-random operand sets: mismatches[0] vdiv, [1] vdiv3 / vdiv2, [2] vsqrt -- all must be 0 */
int vofx_test_arith ( int64_t count, uint64_t seed, int64_t *mismatches );
int vofx_test_locate ( int dim, int64_t count, const double *lo, const double *h, const double *normal, const double *fraction, double *halfspaces );
int vofx_test_flux ( int dim, int64_t count, const double *lo, const double *h, const int32_t *face, const double *v, const double *halfspaces, double *flux );
int vofx_test_interface ( int dim, int64_t count, const double *lo, const double *h, const double *halfspaces, int32_t *nverts, double *centroid, double *measure );

/* the band list of the last finished loop pass (storage indices of the mixed cells the pass reconstructed, in no defined
 * order): at most `capacity` entries are copied to cells_host, *n_cells receives the length of the list; synchronises */
int vofx_band_list_fetch ( vofx_ctx *ctx, int32_t *cells_host, int64_t capacity, int64_t *n_cells );

/* work counters of the step in flight (between vofx_step_phase( 1 ) and vofx_step_finish; synchronises): out[0] cells in the
 * band list, out[1] face clips left to the serial kernel (a node of the swept box on the plane), out[2] cells whose
 * plane-constant solve is left to the serial kernel (tied node heights outside the table), out[3] cells with a root search */
int vofx_step_counters ( vofx_ctx *ctx, int64_t *out );

/* counters of kernel launches issued by this context since creation (bench.py reports them as gpu_launches) */
int64_t vofx_launch_count ( const vofx_ctx *ctx );

/* per-kernel device timing for the roofline report: while enabled, every kernel launch of the context is bracketed
 * by a CUDA event pair on the launching stream.  vofx_profile_read synchronises and returns, per kernel name
 * (names: max_entries rows of 48 chars), the summed duration in ms and the number of launches, then clears. */
int vofx_profile_enable ( vofx_ctx *ctx, int enable );
int vofx_profile_read ( vofx_ctx *ctx, int max_entries, char *names, double *total_ms, int64_t *launches, int *n_entries );

/* ==== unstructured 2D grids (triangles and quadrilaterals): CSR vertex-neighbour stencils ==========================
 * The grid side (the DUNE adaptor) flattens `elements( gridView )` / `intersections( gridView, entity )` into the
 * arrays below, in leaf-index order.  All geometry the reference asks the grid for is DATA here and is used verbatim
 * (never recomputed): geometry().corner/center/volume of cells and intersections, centerUnitOuterNormal,
 * integrationOuterNormal at the intersection centre.  Arrays are HOST pointers, copied at creation.
 *   corners        DUNE corner order; the library forms makePolygon( geometry ) (geometry/2d/polygon.hh:230-241:
 *                  triangle 0,1,2; quadrilateral 0,1,3,2)
 *   faces          the intersections of a cell in the grid's iteration order; face_neighbor = index of
 *                  intersection.outside() or -1 without neighbour
 *   stencil        VertexNeighborsStencil (stencil/vertexneighborsstencil.hh:52-94): all cells sharing a vertex with the
 *                  cell, ascending index, self excluded
 * Operators (same semantics and bits as the structured ones): flagging.hh:35-70, reconstruction/modifiedyoungs.hh:63-134
 * (height functions and HF curvature are Cartesian-only in the reference too), evolution/evolution.hh:58-145 with
 * 2d/upwindpolygon.hh:24-30.  The velocity is SAMPLED: one vector per intersection = what the reference's Velocity
 * returns for refElement.position( 0, 0 ) (velocity.hh:15-20, evolution.hh:107-108).
 * dim 3 (tetrahedra and hexahedra; "[2]" below reads "[3]"): the library forms makePolyhedron( geometry ) from the 4 or 8
 * corners (3d/polyhedron.hh:366-404), locates with geometry/algorithm.hh:119-236 and clips upwindPolygon3d of the 3 or 4
 * intersection corners (3d/upwindpolygon.hh:24-74: prism / cube topology); face_corner_offsets is then required.  The
 * Swartz reconstruction on unstructured grids is 2D only. */
typedef struct vofx_mesh_desc
{
  int32_t dim;                       /* 2 or 3 */
  int64_t n_cells;
  const int32_t *corner_offsets;     /* [n_cells + 1] */
  const double *corners;             /* [corner_offsets[n_cells]][2] */
  const double *centers;             /* [n_cells][2] */
  const double *volumes;             /* [n_cells] */
  const int32_t *face_offsets;       /* [n_cells + 1] */
  const int32_t *face_neighbor;      /* [n_faces], n_faces = face_offsets[n_cells] */
  const double *face_corners;        /* [n_faces][2][2]  intersection.geometry().corner( 0 ), corner( 1 ) */
  const double *face_centers;        /* [n_faces][2] */
  const double *face_unit_normals;   /* [n_faces][2] */
  const double *face_int_normals;    /* [n_faces][2] */
  const double *face_volumes;        /* [n_faces] */
  const int32_t *stencil_offsets;    /* [n_cells + 1] */
  const int32_t *stencil;            /* [stencil_offsets[n_cells]] */
  const int32_t *face_corner_offsets;/* dim 3: [n_faces + 1], intersection s has the corners face_corners[ face_corner_offsets[s] ... ][3];
                                        dim 2: ignored (may be NULL), two corners per intersection */
} vofx_mesh_desc;

typedef struct vofx_mesh vofx_mesh;

int vofx_mesh_create ( const vofx_mesh_desc *desc, int device /* -1: current */, vofx_mesh **out );
void vofx_mesh_destroy ( vofx_mesh *mesh );
/* launch on the caller's CUDA stream (cudaStream_t; NULL: the mesh's own non-blocking stream), like vofx_set_stream */
int vofx_mesh_set_stream ( vofx_mesh *mesh, void *stream );
int64_t vofx_mesh_cells ( const vofx_mesh *mesh );
int64_t vofx_mesh_faces ( const vofx_mesh *mesh );
int64_t vofx_mesh_launch_count ( const vofx_mesh *mesh );

/* operators on DEVICE arrays: u[n_cells], flags[n_cells], halfspaces[n_cells][3] = (n, d), face_velocity[n_faces][2],
 * update[n_cells]; dt_est is a HOST scalar (the call synchronises the mesh's stream to return it) */
int vofx_mesh_flag ( vofx_mesh *mesh, const double *u, double eps, uint8_t *flags );
int vofx_mesh_reconstruct ( vofx_mesh *mesh, const double *u, const uint8_t *flags, double *halfspaces );
/* ModifiedSwartzReconstruction in 2D (reconstruction/modifiedswartz.hh:70-153): Young's first, then the fixed-point iteration */
int vofx_mesh_reconstruct_swartz ( vofx_mesh *mesh, const double *u, const uint8_t *flags, double *halfspaces, int max_iterations );
int vofx_mesh_evolve ( vofx_mesh *mesh, const double *halfspaces, const uint8_t *flags, const double *face_velocity, double dt,
                       double *update, double *dt_est );
/* the same on HOST arrays (staged through device scratch of the mesh object) */
int vofx_mesh_flag_host ( vofx_mesh *mesh, const double *u, double eps, uint8_t *flags );
int vofx_mesh_reconstruct_host ( vofx_mesh *mesh, const double *u, const uint8_t *flags, double *halfspaces );
int vofx_mesh_reconstruct_swartz_host ( vofx_mesh *mesh, const double *u, const uint8_t *flags, double *halfspaces, int max_iterations );
int vofx_mesh_evolve_host ( vofx_mesh *mesh, const double *halfspaces, const uint8_t *flags, const double *face_velocity, double dt,
                            double *update, double *dt_est );
/* the time loop of algorithm.hh:68-95 on a DEVICE-RESIDENT colour function: upload once, run `passes` loop passes
 * (flag, Young's, evolve, u += update, time += dt if dt > 0, dt = cfl * dtEst) without any host synchronisation, read
 * back.  time_io / dt_io are HOST scalars: in = state before the first pass, out = state after the last. */
/* reconstruction used by vofx_mesh_run: VOFX_RECON_YOUNGS (default) or VOFX_RECON_SWARTZ with max_iterations (reference default 10) */
int vofx_mesh_set_reconstruction ( vofx_mesh *mesh, int kind, int max_iterations );
int vofx_mesh_color_upload ( vofx_mesh *mesh, const double *u_host );
int vofx_mesh_velocity_upload ( vofx_mesh *mesh, const double *face_velocity_host );
int vofx_mesh_run ( vofx_mesh *mesh, double eps, double cfl, int passes, double *time_io, double *dt_io );
int vofx_mesh_color_download ( vofx_mesh *mesh, double *u_host, uint8_t *flags_host /* may be NULL */ );

#ifdef __cplusplus
}
#endif

#endif /* VOFX_H */
