// vofx: implementation of the C ABI in include/vofx.h (host side: context, velocity tables, launches, staging).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
//   -fmad=false: the parity contract of vofx_math.cuh (no FMA contraction in the device arithmetic).
#include "../../include/vofx.h"

#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vofx_launch.h"

using namespace vofx;

namespace
{

  thread_local std::string g_lastError;

  int fail ( int status, const std::string &msg )
  {
    g_lastError = msg;
    return status;
  }

#define VOFX_CUDA( call )                                                                                       \
  do                                                                                                            \
  {                                                                                                             \
    const cudaError_t vofx_err_ = ( call );                                                                     \
    if( vofx_err_ != cudaSuccess )                                                                              \
      return fail( VOFX_ERR_CUDA, std::string( #call ) + ": " + cudaGetErrorString( vofx_err_ ) );              \
  } while( 0 )

  constexpr int kSMs = 148;   // B200: grids are sized in multiples of the SM count

} // namespace

namespace vofx
{
  // the error text of this thread, shared with the other translation units of the C ABI (vofx_mesh.cu)
  int set_error ( int status, const std::string &msg ) { return fail( status, msg ); }
}

struct vofx_ctx
{
  vofx_grid_desc desc;
  Grid grid;
  int device = 0;
  cudaStream_t stream = nullptr;
  int numSMs = kSMs;

  // scratch owned by the context
  int capacity = 0;
  int *mixedList = nullptr;
  int *slot = nullptr;
  FluxRecord *records = nullptr;
  int *tasks = nullptr;             // 3D clip tasks ( mixed-list slot * 8 + face ), at most 6 per mixed cell
  int taskCapacity = 0;
  void *sweepTable = nullptr;       // coop::SweepTable
  int *serialCells = nullptr;       // mixed cells whose plane constant is left to the serial walk
  // batched 3D Swartz: per-cell and per-(cell, neighbour) records, grow-only
  unsigned char *swzCells = nullptr, *swzTasks = nullptr;
  size_t swzCellCap = 0, swzTaskCap = 0;
  size_t swzCellCount = 0;          // cells the buffers hold
  size_t swzHint = 0;               // band length of the last pass the host has seen (readState)
  int *swzCounter = nullptr;
  int *swzOrder = nullptr, *swzSort = nullptr;   // per iteration: active tasks sorted by class, and the sort's counters (64 ints)
  size_t swzOrderCap = 0;
  int *swzCellList = nullptr, *swzListCount = nullptr;   // the cells still iterating: two lists of swzCellCount entries, and their lengths
  size_t swzCellListCap = 0;
  bool swzSortTasks = true;                      // VOFX_SWZ_NO_SORT=1 switches the per-iteration task sort off
  // interface extraction (vofx_interface): grow-only scratch
  int *ifBlockSums = nullptr, *ifCells = nullptr, *ifNverts = nullptr, *ifOffsets = nullptr;
  long long *ifTotal = nullptr;
  double *ifVerts = nullptr, *ifCsr = nullptr;
  size_t ifBlockCap = 0, ifCellCap = 0, ifNvertsCap = 0, ifOffsetsCap = 0, ifVertsCap = 0, ifCsrCap = 0;
  long long ifMixed = 0, ifVertices = 0;
  // curvature entries written by the previous step of the time loop (sparse clear instead of a dense memset)
  int *curvList = nullptr, *curvListCount = nullptr;
  const double *curvListOwner = nullptr;
  // sparse write-back of the host-buffer step: (cell, new value) of every cell the update touched
  int *changedIdx = nullptr;
  double *changedVal = nullptr;
  int changedCapacity = 0;
  bool trackChanges = false;
  long long lastH2D = 0, lastD2H = 0;   // bytes moved by the last vofx_step_host call
  const double *hostMirrorOf = nullptr; // host array whose current content the device mirror dU holds (vofx_step_host_sparse)
  int *touchedIdx = nullptr;            // cells the caller changed on the host since the last call, and their values
  double *touchedVal = nullptr;
  size_t touchedCap = 0, touchedValCap = 0;
  // incremental flagging (vofx_set_incremental_flagging): one mark per aligned 32-cell row segment, and the arrays the
  // last finished pass ran on (the flags / band list are only reusable for exactly those)
  unsigned char *marks = nullptr;
  int *segList = nullptr;
  cudaStream_t sideStream = nullptr;      // side branch of a pass: k_mark_band + k_scan_marks next to the band kernels
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  bool sidePending = false;               // the side branch of the pass in flight has not been joined yet
  // band parts (BandPart, vofx_types.cuh): the chains plane-constant solve -> root search -> fluxes of a pass on own streams
  int bandParts = 2;                      // VOFX_BAND_PARTS, read at creation; measured at 512^3 (ms per step, 20 / 200 steps): 1 chain 0.536 / 0.606, 2 chains 0.531 / 0.599, 3 and 4 chains slower (0.585 / 0.672 before the kernels shrank)
  cudaStream_t partStream[ kMaxBandParts ] = { nullptr, nullptr, nullptr, nullptr };
  cudaEvent_t evBandFork = nullptr, evBandJoin[ kMaxBandParts ] = { nullptr, nullptr, nullptr, nullptr };
  int segsPerRow = 0;
  long long nSegments = 0;
  bool incremental = false;
  const double *primedU = nullptr;
  const unsigned char *primedFlags = nullptr;
  // distributed step (vofx_comm_*): one cudaMalloc block = resident colour function (incl. halos) + CommSlots, exported
  // to the other ranks; their blocks mapped here
  void *window = nullptr;
  size_t windowBytes = 0, slotsOffset = 0;
  CommPeers peers;
  bool commConnected = false;
  cudaStream_t ownStream = nullptr;   // a distributed context never runs on the legacy default stream (see ensureWindow)
  std::vector< void * > ipcMapped;
  unsigned char *costClass = nullptr;   // per cell: how expensive its plane-constant solve was in the previous step (k_cost_order)
  int *order = nullptr;                 // processing order of the band list for k_youngs_coop3
  bool costOrder = true;
  void *diagScratch = nullptr;     // block partials + results of vofx_diagnostics
  void *rootTasks = nullptr;        // coop::RootTask per mixed cell: brackets found by k_youngs_coop3, roots by k_locate_root3
  int fluxLanes = 8;               // lanes per mixed cell of the 3D flux kernel; measured at 512^3: 0.36 ms (8) vs 0.51 ms (4)
  int fluxBlocksPerSM = 8, updateBlocksPerSM = 8;
  int youngsBlocksPerSM = 60;      // 64-thread blocks of the plane-constant kernel per SM (grid-stride over the band); measured at
                                   // 512^3 (ms): 10 0.322, 16 0.312, 20 0.300, 30 0.271, 40 0.266, 60 0.266, 120 0.268 - 10 blocks are
                                   // resident, a grid of 16 left the second wave 40 % empty
  int swartzLanes = 4, swartzMode = 0;   // VOFX_SWARTZ_LANES / VOFX_SWARTZ_MODE (0 batched, 1 cell, 2 serial), read at creation
  int lanesPerCell = 4;            // measured on B200 at 512^3: 0.47 ms (4 lanes per cell) vs 0.64 ms (8)
  void *clipTable = nullptr;        // coop::ClipCase[ 256 ] by inner mask, then [ 6561 ] by ( outer / inner / on the plane ) pattern
  StepState *state = nullptr;
  StepState *stateHost = nullptr;   // pinned mirror for read-backs
  double *curv1 = nullptr;
  unsigned char *ok1 = nullptr;
  double *scalars = nullptr;        // [0] time override, [1] dt override (operator-level evolve)

  // velocity
  Velocity velHost;                 // device pointers inside
  Velocity *velDev = nullptr;
  std::vector< double * > velTables;
  double *velFaces = nullptr;
  double *velBandFaces = nullptr;   // vofx_velocity_set_band_faces: grow-only
  size_t velBandCap = 0;
  bool velocitySet = false;

  // staging for the host-buffer variants
  void *pinned = nullptr;
  size_t pinnedBytes = 0;
  double *dU = nullptr, *dHs = nullptr, *dUpd = nullptr, *dKappa = nullptr;
  unsigned char *dFlags = nullptr, *dValid = nullptr;

  long long launches = 0;

  // optional per-kernel timing (vofx_profile_*): one CUDA event pair per launch, on the launching stream
  bool profiling = false;
  std::vector< cudaEvent_t > eventPool;
  size_t eventsUsed = 0;
  struct ProfEntry { const char *name; cudaEvent_t begin, end; };
  std::vector< ProfEntry > profEntries;
  cudaEvent_t pendingBegin = nullptr;
};

namespace
{

  int setDevice ( const vofx_ctx *ctx )
  {
    VOFX_CUDA( cudaSetDevice( ctx->device ) );
    return VOFX_OK;
  }

  int gridBlocks ( const vofx_ctx *ctx, long long items, int threads, int blocksPerSM )
  {
    long long need = ( items + threads - 1 ) / threads;
    const long long cap = (long long) ctx->numSMs * blocksPerSM;
    if( need > cap )
      need = cap;
    if( need < 1 )
      need = 1;
    return (int) need;
  }

  cudaEvent_t takeEvent ( vofx_ctx *ctx )
  {
    if( ctx->eventsUsed == ctx->eventPool.size() )
    {
      cudaEvent_t e = nullptr;
      if( cudaEventCreate( &e ) != cudaSuccess )
        return nullptr;
      ctx->eventPool.push_back( e );
    }
    return ctx->eventPool[ ctx->eventsUsed++ ];
  }

  // call directly before a kernel launch; checkLaunch() closes the pair
  void profBegin ( vofx_ctx *ctx )
  {
    ctx->pendingBegin = nullptr;
    if( !ctx->profiling )
      return;
    ctx->pendingBegin = takeEvent( ctx );
    if( ctx->pendingBegin )
      cudaEventRecord( ctx->pendingBegin, ctx->stream );
  }

  int checkLaunch ( vofx_ctx *ctx, const char *what )
  {
    ctx->launches++;
    if( ctx->profiling && ctx->pendingBegin )
    {
      cudaEvent_t end = takeEvent( ctx );
      if( end )
      {
        cudaEventRecord( end, ctx->stream );
        ctx->profEntries.push_back( { what, ctx->pendingBegin, end } );
      }
      ctx->pendingBegin = nullptr;
    }
    const cudaError_t e = cudaGetLastError();
    if( e != cudaSuccess )
      return fail( VOFX_ERR_CUDA, std::string( what ) + ": " + cudaGetErrorString( e ) );
    return VOFX_OK;
  }

  template< class T >
  int ensureDevice ( T *&p, size_t count )
  {
    if( p )
      return VOFX_OK;
    VOFX_CUDA( cudaMalloc( (void **) &p, count * sizeof( T ) ) );
    return VOFX_OK;
  }

  template< class T >
  int growDevice ( T *&p, size_t &cap, size_t count )
  {
    if( count <= cap )
      return VOFX_OK;
    cudaFree( p );
    p = nullptr;
    cap = 0;
    const size_t want = count + count / 4 + 1024;
    VOFX_CUDA( cudaMalloc( (void **) &p, want * sizeof( T ) ) );
    cap = want;
    return VOFX_OK;
  }

  int ensurePinned ( vofx_ctx *ctx, size_t bytes )
  {
    if( ctx->pinnedBytes >= bytes )
      return VOFX_OK;
    if( ctx->pinned )
      cudaFreeHost( ctx->pinned );
    ctx->pinned = nullptr;
    ctx->pinnedBytes = 0;
    VOFX_CUDA( cudaMallocHost( &ctx->pinned, bytes ) );
    ctx->pinnedBytes = bytes;
    return VOFX_OK;
  }

  // ---- launches -------------------------------------------------------------------------------------------------

  bool flagByRows ( const vofx_ctx *ctx, const double *u, const unsigned char *flags )
  {
    return ( ctx->grid.ns[ 0 ] & 3 ) == 0 && ( reinterpret_cast< uintptr_t >( u ) & 15 ) == 0 && ( reinterpret_cast< uintptr_t >( flags ) & 3 ) == 0;
  }

  // flags (+ mixed list) of the storage layers [ layer0, layer1 ) along the last axis; the whole box by default
  int launchFlag ( vofx_ctx *ctx, const double *u, double eps, unsigned char *flags, bool compact, int layer0 = 0, int layer1 = -1 )
  {
    const Grid &g = ctx->grid;
    const int la = g.dim - 1;
    const int nLayers = g.ns[ la ];
    if( layer1 < 0 )
      layer1 = nLayers;
    const bool whole = ( layer0 == 0 && layer1 == nLayers );
    if( flagByRows( ctx, u, flags ) )
    {
      // storage layers outside the global domain (halos of the first / last slab) hold no cells: they are never read,
      // so a slab's arrays may alias a window of one contiguous global array (PipelinedAlgorithm, decomposition.py)
      const int dom0 = ( -g.g0[ la ] > 0 ) ? -g.g0[ la ] : 0;
      const int dom1 = ( g.ng[ la ] - g.g0[ la ] < nLayers ) ? g.ng[ la ] - g.g0[ la ] : nLayers;
      layer0 = layer0 > dom0 ? layer0 : dom0;
      layer1 = layer1 < dom1 ? layer1 : dom1;
    }
    if( layer0 >= layer1 )
      return VOFX_OK;
    profBegin( ctx );
    if( flagByRows( ctx, u, flags ) )
    {
      // row kernel: blocks of 128 threads walk 4 rows each (grid >> 148 SMs x resident blocks on every production size)
      const int rowsPerLayer = ( g.dim == 3 ) ? g.ns[ 1 ] : 1;
      const int rowBegin = layer0 * rowsPerLayer, rowEnd = layer1 * rowsPerLayer;
      const int nRows = rowEnd - rowBegin;
      // measured at 512^3 (ms): 1 row 0.234, 2 0.230, 4 0.226, 8 0.222, 16 0.229, 32 0.248; a software-pipelined variant
      // (loads of the next row issued before the current one is processed) was slower (0.259)
      const int rowsPerBlock = ( nRows >= 8 * ctx->numSMs * 16 ) ? 8 : 1;
      const int blocks = ( nRows + rowsPerBlock - 1 ) / rowsPerBlock;
      launch::flag_rows( g.dim, blocks, ctx->stream, g, u, eps, flags, ctx->mixedList, ctx->slot, ctx->capacity, ctx->state, compact ? 1 : 0, rowsPerBlock, rowBegin, rowEnd );
      return checkLaunch( ctx, "k_flag_compact" );
    }
    if( !whole )
      return fail( VOFX_ERR_ARGUMENT, "partial flagging needs the row kernel (row length a multiple of 4, aligned arrays)" );
    const long long quads = ( g.cells + 3 ) / 4;
    const int blocks = gridBlocks( ctx, quads, 256, 8 );
    launch::flag_compact( g.dim, blocks, ctx->stream, g, u, eps, flags, ctx->mixedList, ctx->slot, ctx->capacity, ctx->state, compact ? 1 : 0 );
    return checkLaunch( ctx, "k_flag_compact" );
  }

  bool capturing ( const vofx_ctx *ctx )
  {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if( cudaStreamIsCapturing( ctx->stream, &st ) != cudaSuccess )
    {
      cudaGetLastError();
      return false;
    }
    return st != cudaStreamCaptureStatusNone;
  }

  // F1 inside the time loop.  part 0: the layers whose flags depend on owned colour values only (all layers without
  // halos); part 1: the halo layers and the first / last owned layer (always the dense kernel: their colour values arrive
  // from another rank).  With incremental flagging part 0 is k_reflag once the previous pass ran on the same arrays.
  int launchStepFlags ( vofx_ctx *ctx, const vofx_step_params *p, const double *u, unsigned char *flags, int part )
  {
    const Grid &g = ctx->grid;
    const int nLayers = g.ns[ g.dim - 1 ];
    const bool split = ctx->desc.halo > 0 && flagByRows( ctx, u, flags ) && g.own1 - g.own0 > 2;
    const int in0 = split ? g.own0 + 1 : 0, in1 = split ? g.own1 - 1 : nLayers;
    if( ctx->desc.halo > 0 && !split )
      return part == 1 ? launchFlag( ctx, u, p->eps, flags, true ) : VOFX_OK;     // thin slab: one dense pass after the halos arrived
    if( part == 1 )
    {
      if( !split )
        return VOFX_OK;
      int rc = launchFlag( ctx, u, p->eps, flags, true, 0, in0 );
      return rc ? rc : launchFlag( ctx, u, p->eps, flags, true, in1, nLayers );
    }
    const bool inc = ctx->incremental && ctx->marks && ctx->primedU == u && ctx->primedFlags == flags;
    if( !inc )
      return launchFlag( ctx, u, p->eps, flags, true, in0, in1 );
    profBegin( ctx );
    // the previous pass's side branch marked the row segments within two face steps of its band and listed them
    launch::reflag( g.dim, ctx->numSMs * 8, ctx->stream, g, u, p->eps, flags, ctx->mixedList, ctx->slot, ctx->capacity, ctx->state, ctx->segList, ctx->segsPerRow );
    return checkLaunch( ctx, "k_reflag" );
  }

  // Side branch of a pass (incremental flagging): once the band list of the pass is complete, k_mark_band and k_scan_marks
  // prepare the NEXT pass's segment list on a second stream, next to reconstruction, fluxes and update; joinSide() brings
  // the branch back before the pass ends (inside a stream capture this is a fork / join of the graph).
  int forkSide ( vofx_ctx *ctx, const double *u, const unsigned char *flags )
  {
    if( !ctx->incremental || !ctx->marks )
      return VOFX_OK;
    const Grid &g = ctx->grid;
    if( ctx->desc.halo > 0 && !( flagByRows( ctx, u, flags ) && g.own1 - g.own0 > 2 ) )
      return VOFX_OK;                        // thin slab: always flagged densely
    if( !ctx->sideStream )
    {
      VOFX_CUDA( cudaStreamCreateWithFlags( &ctx->sideStream, cudaStreamNonBlocking ) );
      VOFX_CUDA( cudaEventCreateWithFlags( &ctx->evFork, cudaEventDisableTiming ) );
      VOFX_CUDA( cudaEventCreateWithFlags( &ctx->evJoin, cudaEventDisableTiming ) );
    }
    const bool split = ctx->desc.halo > 0;
    const int in0 = split ? g.own0 + 1 : 0, in1 = split ? g.own1 - 1 : g.ns[ g.dim - 1 ];
    VOFX_CUDA( cudaEventRecord( ctx->evFork, ctx->stream ) );
    VOFX_CUDA( cudaStreamWaitEvent( ctx->sideStream, ctx->evFork, 0 ) );
    launch::mark_scan( g.dim, ctx->numSMs * 4, ctx->sideStream, g, ctx->mixedList, ctx->capacity, ctx->state, ctx->marks, ctx->segList, ctx->segsPerRow,
                       ctx->nSegments, in0, in1 );
    ctx->launches += 2;
    VOFX_CUDA( cudaGetLastError() );
    VOFX_CUDA( cudaEventRecord( ctx->evJoin, ctx->sideStream ) );
    ctx->sidePending = true;
    return VOFX_OK;
  }

  int joinSide ( vofx_ctx *ctx )
  {
    if( !ctx->sidePending )
      return VOFX_OK;
    VOFX_CUDA( cudaStreamWaitEvent( ctx->stream, ctx->evJoin, 0 ) );
    ctx->sidePending = false;
    return VOFX_OK;
  }

  // the pass that was just issued leaves flags / band list / previous-band list valid for ( u, flags )
  void notePassIssued ( vofx_ctx *ctx, const double *u, const unsigned char *flags )
  {
    if( !ctx->incremental || capturing( ctx ) )
      return;
    ctx->primedU = u;
    ctx->primedFlags = flags;
  }

  int launchResetCounters ( vofx_ctx *ctx )
  {
    // operator-level calls rebuild the band list of the context: the next loop pass must not reuse it
    ctx->primedU = nullptr;
    ctx->primedFlags = nullptr;
    profBegin( ctx );
    launch::reset_counters( ctx->stream, ctx->state );
    return checkLaunch( ctx, "k_reset_counters" );
  }

  int launchCompactFlags ( vofx_ctx *ctx, const unsigned char *flags )
  {
    const Grid &g = ctx->grid;
    const int blocks = gridBlocks( ctx, g.cells, 256, 8 );
    profBegin( ctx );
    launch::compact_flags( g.dim, blocks, ctx->stream, g, flags, ctx->mixedList, ctx->slot, ctx->capacity, ctx->state );
    return checkLaunch( ctx, "k_compact_flags" );
  }

  int readState ( vofx_ctx *ctx, StepState &s );

  int launchReconstruct ( vofx_ctx *ctx, int kind, const double *u, const unsigned char *flags, double *hs, int maxIterations )
  {
    const Grid &g = ctx->grid;
    const int blocks = ctx->numSMs * 8;
    if( g.dim == 2 && kind == VOFX_RECON_HF )
    {
      // Young's normal + height-function normal + one plane-constant solve per cell, in one pass
      profBegin( ctx );
      launch::hf_fused2( blocks, ctx->stream, g, u, ctx->mixedList, ctx->state, ctx->capacity, hs );
      return checkLaunch( ctx, "k_hf_fused" );
    }
    profBegin( ctx );
    if( g.dim == 2 )
      launch::youngs2( blocks, ctx->stream, g, u, ctx->mixedList, ctx->state, ctx->capacity, hs );
    else
    {
      // 3D height functions ride in the same kernel (Young's normal -> HF normal -> one plane-constant solve)
      const bool hf3 = ( kind == VOFX_RECON_HF );
      const bool ordered = ctx->costOrder && ctx->order && ctx->costClass;
      if( ordered )
      {
        launch::cost_order( ctx->numSMs * 2, ctx->stream, ctx->mixedList, ctx->state, ctx->capacity, ctx->costClass, ctx->order );
        ctx->launches++;       // k_cost_hist + k_cost_order
        int rco = checkLaunch( ctx, "k_cost_order" );
        if( rco )
          return rco;
        profBegin( ctx );
      }
      launch::youngs_coop3( ctx->lanesPerCell, ctx->numSMs * ctx->youngsBlocksPerSM / 2, ctx->stream, g, u, ctx->mixedList, ctx->state, ctx->capacity, hs, ctx->sweepTable, ctx->serialCells, ctx->rootTasks, hf3 ? 1 : 0,
                            ordered ? ctx->order : nullptr, ctx->costClass );
      int rc3 = checkLaunch( ctx, hf3 ? "k_hf_coop" : "k_youngs" );
      if( rc3 )
        return rc3;
      profBegin( ctx );
      launch::locate_root3( ctx->numSMs * 8, ctx->stream, ctx->state, ctx->capacity, ctx->rootTasks, hs );
      rc3 = checkLaunch( ctx, "k_locate_root" );
      if( rc3 )
        return rc3;
      // cells whose walk hits a state the reference leaves undefined (never seen on the BASELINE configs): serial kernel
      profBegin( ctx );
      launch::locate_serial3( ctx->stream, g, u, ctx->state, hs, ctx->serialCells, ctx->capacity );
    }
    int rc = checkLaunch( ctx, g.dim == 2 ? "k_youngs" : "k_locate_serial" );
    if( rc )
      return rc;
    if( kind == VOFX_RECON_HF )
    {
      // 2D: k_hf_fused above; 3D: done inside k_youngs_coop3< L, true >
    }
    else if( kind == VOFX_RECON_SWARTZ )
    {
      profBegin( ctx );
      if( g.dim == 2 )
        launch::swartz2( ctx->numSMs * 16, ctx->stream, g, u, flags, ctx->mixedList, ctx->state, ctx->capacity, hs, maxIterations );
      else
      {
        const int lanes = ctx->swartzLanes;   // measured at 256^3 per step: batched 10.2 ms (4 lanes) / 12.7 ms (8); per-cell kernel 38.6 ms; serial kernel 261 ms
        if( ctx->swartzMode == 2 )
          launch::swartz3( ctx->numSMs * 16, ctx->stream, g, u, flags, ctx->mixedList, ctx->state, ctx->capacity, hs, maxIterations );
        else if( ctx->swartzMode == 1 )
          launch::swartz_coop3( lanes, ctx->numSMs * 16, ctx->stream, g, u, flags, ctx->mixedList, ctx->state, ctx->capacity, hs, maxIterations, ctx->sweepTable );
        else
        {
          // batched over the band.  The buffers are sized on the host from the band of an EARLIER pass (the first pass reads
          // its own count back once; afterwards every read of the loop state refreshes the hint), so a pass neither blocks
          // the stream nor allocates, and it can be captured into a CUDA graph; k_swz_setup guards the capacities on the device
          if( !capturing( ctx ) )
          {
            size_t want = ctx->swzHint;
            if( !ctx->swzCells || want == 0 )
            {
              StepState now;
              int rcs = readState( ctx, now );
              if( rcs )
                return rcs;
              want = (size_t) ( now.mixedCount < ctx->capacity ? now.mixedCount : ctx->capacity );
            }
            const size_t cellsWanted = want + want / 2 + 4096;        // 50 % headroom: the band grows slowly
            if( cellsWanted > ctx->swzCellCount )
            {
              int rcs = growDevice( ctx->swzCells, ctx->swzCellCap, cellsWanted * launch::swz_cell_bytes() + 16 );
              rcs = rcs ? rcs : growDevice( ctx->swzTasks, ctx->swzTaskCap, cellsWanted * 27 * launch::swz_task_bytes() + 16 );
              rcs = rcs ? rcs : ensureDevice( ctx->swzCounter, (size_t) 1 );
              rcs = rcs ? rcs : growDevice( ctx->swzOrder, ctx->swzOrderCap, cellsWanted * 27 + 16 );
              rcs = rcs ? rcs : ensureDevice( ctx->swzSort, (size_t) 64 );
              rcs = rcs ? rcs : growDevice( ctx->swzCellList, ctx->swzCellListCap, 2 * cellsWanted + 16 );
              rcs = rcs ? rcs : ensureDevice( ctx->swzListCount, (size_t) 4 );
              if( rcs )
                return rcs;
              ctx->swzCellCount = cellsWanted;
            }
          }
          if( !ctx->swzCells )
            return fail( VOFX_ERR_STATE, "the batched 3D Swartz pass must run once eagerly before it is captured (it sizes its buffers)" );
          VOFX_CUDA( cudaMemsetAsync( ctx->swzCounter, 0, sizeof( int ), ctx->stream ) );
          profBegin( ctx );
          launch::swartz_batched3( lanes, ctx->numSMs, ctx->stream, g, u, flags, ctx->mixedList, ctx->state, ctx->capacity, hs, maxIterations, ctx->sweepTable,
                                   ctx->swzCells, ctx->swzTasks, ctx->swzCounter, (int) ctx->swzCellCount, (int) ( ctx->swzCellCount * 27 ),
                                   ctx->swzSortTasks ? ctx->swzOrder : nullptr, ctx->swzSortTasks ? ctx->swzSort : nullptr,
                                   ctx->swzSortTasks ? ctx->swzCellList : nullptr, ctx->swzSortTasks ? ctx->swzListCount : nullptr );
          ctx->launches += ( ctx->swzSortTasks ? 6 : 4 ) * maxIterations;
        }
      }
      rc = checkLaunch( ctx, "k_swartz" );
    }
    else if( kind != VOFX_RECON_YOUNGS )
      return fail( VOFX_ERR_ARGUMENT, "unknown reconstruction kind" );
    return rc;
  }

  int launchFlux ( vofx_ctx *ctx, const double *hs, const unsigned char *flags, const double *timeOverride, const double *dtOverride )
  {
    const Grid &g = ctx->grid;
    const int blocks = ctx->numSMs * 8;
    profBegin( ctx );
    if( g.dim == 2 )
    {
      launch::flux2( blocks, ctx->stream, g, ctx->velDev, hs, flags, ctx->mixedList, ctx->state, ctx->capacity, ctx->records, timeOverride, dtOverride );
      return checkLaunch( ctx, "k_flux" );
    }
    launch::flux_cell3( ctx->fluxLanes, ctx->numSMs * ctx->fluxBlocksPerSM, ctx->stream, g, ctx->velDev, ctx->velHost, hs, flags, ctx->mixedList, ctx->state, ctx->capacity, ctx->records, ctx->tasks, ctx->taskCapacity,
                        ctx->clipTable, timeOverride, dtOverride );
    ctx->launches++;     // k_flux_clip_serial3 rides along (a handful of clips per step)
    return checkLaunch( ctx, "k_flux" );
  }

  int launchCurvature ( vofx_ctx *ctx, const double *u, const double *hs, const unsigned char *flags, double *kappa, unsigned char *valid, bool sparseClear );

  // Reconstruction and fluxes of a loop pass.  3D Young's / height functions without curvature: `bandParts` chains
  // ( plane-constant solve -> root search -> serial leftovers -> fluxes ) on as many streams, joined before the update
  // (BandPart, vofx_types.cuh); everything else: one kernel after the other on the context's stream.
  int launchBand ( vofx_ctx *ctx, const vofx_step_params *p, const double *u, const unsigned char *flags, double *hs, double *kappa )
  {
    const Grid &g = ctx->grid;
    const int kind = p->reconstruction;
    const int parts = ctx->bandParts;
    const bool chains = parts > 1 && g.dim == 3 && ( kind == VOFX_RECON_YOUNGS || kind == VOFX_RECON_HF ) && !p->with_curvature
                        && !ctx->profiling;      // the per-kernel event pairs of vofx_profile_* time one kernel after the other
    if( !chains )
    {
      int rc = launchReconstruct( ctx, kind, u, flags, hs, p->max_iterations );
      if( !rc && p->with_curvature )
        rc = launchCurvature( ctx, u, hs, flags, kappa, nullptr, true );
      return rc ? rc : launchFlux( ctx, hs, flags, nullptr, nullptr );
    }
    if( !ctx->evBandFork )
    {
      int least = 0, greatest = 0;
      VOFX_CUDA( cudaDeviceGetStreamPriorityRange( &least, &greatest ) );
      for( int q = 0; q < parts; ++q )
      {
        // the chain that started first runs at the highest priority: its successors fill what it leaves free
        const int prio = ( greatest + q < least ) ? greatest + q : least;
        VOFX_CUDA( cudaStreamCreateWithPriority( &ctx->partStream[ q ], cudaStreamNonBlocking, prio ) );
        VOFX_CUDA( cudaEventCreateWithFlags( &ctx->evBandJoin[ q ], cudaEventDisableTiming ) );
      }
      VOFX_CUDA( cudaEventCreateWithFlags( &ctx->evBandFork, cudaEventDisableTiming ) );
    }
    const bool hf3 = ( kind == VOFX_RECON_HF );
    const bool ordered = ctx->costOrder && ctx->order && ctx->costClass;
    if( ordered )
    {
      launch::cost_order( ctx->numSMs * 2, ctx->stream, ctx->mixedList, ctx->state, ctx->capacity, ctx->costClass, ctx->order, parts );
      ctx->launches++;         // k_cost_hist + k_cost_order
      int rc = checkLaunch( ctx, "k_cost_order" );
      if( rc )
        return rc;
    }
    const int *order = ordered ? ctx->order : nullptr;
    VOFX_CUDA( cudaEventRecord( ctx->evBandFork, ctx->stream ) );
    for( int q = 0; q < parts; ++q )
    {
      cudaStream_t s = ctx->partStream[ q ];
      BandPart bp;
      bp.part = q;
      bp.parts = parts;
      VOFX_CUDA( cudaStreamWaitEvent( s, ctx->evBandFork, 0 ) );
      const int yBlocks = ( ctx->numSMs * ctx->youngsBlocksPerSM / 2 + parts - 1 ) / parts;
      launch::youngs_coop3( ctx->lanesPerCell, yBlocks, s, g, u, ctx->mixedList, ctx->state, ctx->capacity, hs, ctx->sweepTable, ctx->serialCells, ctx->rootTasks,
                            hf3 ? 1 : 0, order, ctx->costClass, bp );
      launch::locate_root3( ( ctx->numSMs * 8 + parts - 1 ) / parts, s, ctx->state, ctx->capacity, ctx->rootTasks, hs, bp );
      launch::locate_serial3( s, g, u, ctx->state, hs, ctx->serialCells, ctx->capacity, bp );
      launch::flux_cell3( ctx->fluxLanes, ( ctx->numSMs * ctx->fluxBlocksPerSM + parts - 1 ) / parts, s, g, ctx->velDev, ctx->velHost, hs, flags, ctx->mixedList, ctx->state,
                          ctx->capacity, ctx->records, ctx->tasks, ctx->taskCapacity, ctx->clipTable, nullptr, nullptr, bp, 0 );
      ctx->launches += 4;
      VOFX_CUDA( cudaGetLastError() );
      VOFX_CUDA( cudaEventRecord( ctx->evBandJoin[ q ], s ) );
    }
    for( int q = 0; q < parts; ++q )
      VOFX_CUDA( cudaStreamWaitEvent( ctx->stream, ctx->evBandJoin[ q ], 0 ) );
    // the clips no chain could take from its table (a handful per step)
    launch::flux_clip_serial3( ctx->stream, g, ctx->velDev, hs, flags, ctx->mixedList, ctx->state, ctx->records, ctx->tasks, ctx->taskCapacity, nullptr, nullptr );
    return checkLaunch( ctx, "k_flux_clip_serial" );
  }

  int launchUpdate ( vofx_ctx *ctx, const unsigned char *flags, double *target, bool accumulate )
  {
    const Grid &g = ctx->grid;
    const int blocks = ctx->numSMs * ctx->updateBlocksPerSM;
    profBegin( ctx );
    const bool track = accumulate && ctx->trackChanges && ctx->changedIdx;
    launch::update( g.dim, blocks, ctx->stream, g, flags, ctx->mixedList, ctx->slot, ctx->state, ctx->capacity, ctx->records, target, accumulate ? 1 : 0,
                    ctx->state, track ? ctx->changedIdx : nullptr, track ? ctx->changedVal : nullptr, ctx->changedCapacity,
                    nullptr, ctx->segsPerRow,
                    ( accumulate && ctx->commConnected && target == ctx->dU ) ? &ctx->peers : nullptr );
    return checkLaunch( ctx, "k_update" );
  }

  int launchCurvature ( vofx_ctx *ctx, const double *u, const double *hs, const unsigned char *flags, double *kappa, unsigned char *valid,
                        bool sparseClear )
  {
    const Grid &g = ctx->grid;
    const int blocks = ctx->numSMs * 8;
    // curvature[ entity ] = 0 and satisfiesConstraint_[ entity ] = 0 for all cells, cartesianheightfunctioncurvature.hh:57-60.
    // Inside the time loop the array is the one written by the previous step: only those entries are non-zero.
    if( sparseClear && ctx->curvList && ctx->curvListOwner == kappa )
    {
      launch::clear_listed( ctx->numSMs * 2, ctx->stream, ctx->curvList, ctx->curvListCount, kappa );
      ctx->launches++;
    }
    else
      VOFX_CUDA( cudaMemsetAsync( kappa, 0, sizeof( double ) * (size_t) g.cells, ctx->stream ) );
    if( valid )
      VOFX_CUDA( cudaMemsetAsync( valid, 0, (size_t) g.cells, ctx->stream ) );
    profBegin( ctx );
    launch::curvature_local( g.dim, blocks, ctx->stream, g, u, hs, ctx->mixedList, ctx->state, ctx->capacity, ctx->curv1, ctx->ok1 );
    int rc = checkLaunch( ctx, "k_curvature_local" );
    if( rc )
      return rc;
    profBegin( ctx );
    launch::curvature_average( g.dim, blocks, ctx->stream, g, flags, ctx->mixedList, ctx->slot, ctx->state, ctx->capacity, ctx->curv1, ctx->ok1, kappa, valid );
    rc = checkLaunch( ctx, "k_curvature_average" );
    if( rc || !sparseClear )
    {
      ctx->curvListOwner = nullptr;
      return rc;
    }
    // remember which entries are non-zero now
    if( !ctx->curvList )
    {
      rc = ensureDevice( ctx->curvList, (size_t) ctx->capacity );
      rc = rc ? rc : ensureDevice( ctx->curvListCount, (size_t) 1 );
      if( rc )
        return rc;
    }
    launch::copy_list( ctx->numSMs * 2, ctx->stream, ctx->mixedList, ctx->state, ctx->capacity, ctx->curvList, ctx->curvListCount );
    ctx->launches++;
    ctx->curvListOwner = kappa;
    return VOFX_OK;
  }

  int checkOverflow ( vofx_ctx *ctx, const StepState &s )
  {
    if( s.commError )
      return fail( VOFX_ERR_STATE, "distributed step: the wait for " + std::string( ( s.commError & 15 ) == 1 ? "the halo pushes" : ( ( s.commError & 15 ) == 2 ? "the fluxes" : "the time-step estimate" ) )
                   + " of rank " + std::to_string( ( s.commError >> 4 ) & 15 ) + " in pass " + std::to_string( s.commError >> 8 )
                   + " timed out (a rank died, or the ranks' call sequences differ)" );
    if( s.overflow == 2 || s.overflowLast == 2 )
    {
      ctx->swzHint = ctx->swzHint * 2 + 4096;      // the next eager pass enlarges the buffers
      return fail( VOFX_ERR_STATE, "the band outgrew the buffers of the batched Swartz reconstruction (restart the loop from the last state: vofx_step_begin)" );
    }
    if( s.overflow || s.overflowLast )
      return fail( VOFX_ERR_STATE, "mixed-cell list overflow: more than " + std::to_string( ctx->capacity ) + " mixed cells" );
    return VOFX_OK;
  }

  int readState ( vofx_ctx *ctx, StepState &s )
  {
    // through PINNED memory: an asynchronous copy into pageable memory is staged by the runtime, which waits for the stream
    // inside the call -- with several contexts driven by threads of one process that wait can block the other threads' launches
    // (the stream may hold a kernel that waits for exactly those launches)
    if( !ctx->stateHost )
      VOFX_CUDA( cudaMallocHost( (void **) &ctx->stateHost, sizeof( StepState ) ) );
    VOFX_CUDA( cudaMemcpyAsync( ctx->stateHost, ctx->state, sizeof( StepState ), cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    s = *ctx->stateHost;
    const int band = s.mixedCount > s.mixedCountLast ? s.mixedCount : s.mixedCountLast;
    if( band > 0 )
      ctx->swzHint = (size_t) band;
    return checkOverflow( ctx, s );
  }

  // ---- built-in velocity tables ---------------------------------------------------------------------------------

  typedef double ( *Factor ) ( double );
  double fOne ( double ) { return 1.0; }
  double fMinusHalf ( double x ) { return x - 0.5; }                 // c = x - 0.5
  double fHalfMinus ( double x ) { return 0.5 - x; }
  double fNegMinusHalf ( double x ) { return -( x - 0.5 ); }
  double fSflowLin ( double x ) { return 4 * x - 2; }
  double fSflowCube ( double x ) { const double y = 4 * x - 2; return y * y * y; }
  double fSinSq ( double x ) { const double s = std::sin( M_PI * x ); return s * s; }
  double fTwoSinSq ( double x ) { return 2.0 * fSinSq( x ); }
  double fSinTwo ( double x ) { return std::sin( 2.0 * M_PI * x ); }
  double fNegSinTwo ( double x ) { return -std::sin( 2.0 * M_PI * x ); }

  int setFactor ( vofx_ctx *ctx, int component, int term, int axis, Factor f )
  {
    const vofx_grid_desc &d = ctx->desc;
    const int n = d.n_global[ axis ];
    std::vector< double > centre( n ), lower( n ), upper( n );
    for( int i = 0; i < n; ++i )
    {
      const double lo = d.origin[ axis ] + i * d.h[ axis ];
      lower[ i ] = f( lo );
      upper[ i ] = f( lo + d.h[ axis ] );
      centre[ i ] = f( lo + 0.5 * d.h[ axis ] );
    }
    return vofx_velocity_set_factor( ctx, component, term, axis, centre.data(), lower.data(), upper.data() );
  }

  int setTerm ( vofx_ctx *ctx, int component, int term, Factor fx, Factor fy, Factor fz )
  {
    int rc = setFactor( ctx, component, term, 0, fx );
    rc = rc ? rc : setFactor( ctx, component, term, 1, fy );
    if( ctx->desc.dim == 3 )
      rc = rc ? rc : setFactor( ctx, component, term, 2, fz );
    return rc;
  }

  int uploadVelocity ( vofx_ctx *ctx )
  {
    VOFX_CUDA( cudaMemcpyAsync( ctx->velDev, &ctx->velHost, sizeof( Velocity ), cudaMemcpyHostToDevice, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    return VOFX_OK;
  }

  int requireVelocity ( vofx_ctx *ctx )
  {
    if( !ctx->velocitySet )
      return fail( VOFX_ERR_STATE, "no velocity set (vofx_velocity_builtin / vofx_velocity_set_*)" );
    return VOFX_OK;
  }

} // namespace


extern "C"
{

  const char *vofx_last_error ( void ) { return g_lastError.c_str(); }

  int vofx_create ( const vofx_grid_desc *desc, int device, vofx_ctx **out )
  {
    if( !desc || !out )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    if( desc->dim != 2 && desc->dim != 3 )
      return fail( VOFX_ERR_ARGUMENT, "dim must be 2 or 3" );
    if( desc->halo < 0 )
      return fail( VOFX_ERR_ARGUMENT, "negative halo" );
    for( int d = 0; d < desc->dim; ++d )
    {
      if( desc->n_global[ d ] < 1 || desc->n_local[ d ] < 1 || desc->lo[ d ] < 0 || desc->lo[ d ] + desc->n_local[ d ] > desc->n_global[ d ] )
        return fail( VOFX_ERR_ARGUMENT, "inconsistent grid extents" );
      if( !( desc->h[ d ] > 0.0 ) )
        return fail( VOFX_ERR_ARGUMENT, "mesh width must be positive" );
      if( d != desc->dim - 1 && ( desc->lo[ d ] != 0 || desc->n_local[ d ] != desc->n_global[ d ] ) )
        return fail( VOFX_ERR_ARGUMENT, "only the last axis may be decomposed (slabs)" );
    }

    int count = 0;
    if( cudaGetDeviceCount( &count ) != cudaSuccess || count == 0 )
      return fail( VOFX_ERR_CUDA, "no CUDA device: vofx has no CPU fallback" );
    if( device < 0 )
      VOFX_CUDA( cudaGetDevice( &device ) );
    if( device >= count )
      return fail( VOFX_ERR_ARGUMENT, "device index out of range" );
    VOFX_CUDA( cudaSetDevice( device ) );

    vofx_ctx *ctx = new vofx_ctx();
    ctx->desc = *desc;
    ctx->device = device;
    cudaDeviceProp prop;
    if( cudaGetDeviceProperties( &prop, device ) == cudaSuccess )
      ctx->numSMs = prop.multiProcessorCount;

    Grid &g = ctx->grid;
    g.dim = desc->dim;
    const int la = desc->dim - 1;
    g.cells = 1;
    for( int d = 0; d < 3; ++d )
    {
      const bool used = d < desc->dim;
      g.ns[ d ] = used ? desc->n_local[ d ] + ( d == la ? 2 * desc->halo : 0 ) : 1;
      g.g0[ d ] = used ? desc->lo[ d ] - ( d == la ? desc->halo : 0 ) : 0;
      g.ng[ d ] = used ? desc->n_global[ d ] : 1;
      g.origin[ d ] = used ? desc->origin[ d ] : 0.0;
      g.h[ d ] = used ? desc->h[ d ] : 1.0;
      g.cells *= g.ns[ d ];
    }
    if( g.cells >= ( 1LL << 31 ) )
    {
      delete ctx;
      return fail( VOFX_ERR_ARGUMENT, "more than 2^31 cells per context" );
    }
    g.own0 = desc->halo;
    g.own1 = desc->halo + desc->n_local[ la ];
    // mixed cells one layer beyond the owned box are reconstructed too (their fluxes enter owned cells)
    const int dom0 = ( -g.g0[ la ] > 0 ) ? -g.g0[ la ] : 0;
    const int dom1 = ( g.ng[ la ] - g.g0[ la ] < g.ns[ la ] ) ? g.ng[ la ] - g.g0[ la ] : g.ns[ la ];
    g.list0 = ( g.own0 - 1 > dom0 ) ? g.own0 - 1 : dom0;
    g.list1 = ( g.own1 + 1 < dom1 ) ? g.own1 + 1 : dom1;

    // capacity of the mixed-cell list: every cell on small grids, an eighth of the cells (>= 2^20) on large ones
    long long cap = g.cells / 8;
    if( cap < ( 1LL << 20 ) )
      cap = 1LL << 20;
    if( cap > g.cells )
      cap = g.cells;
    ctx->capacity = (int) cap;

    int rc = VOFX_OK;
    auto alloc = [ & ] ( void **p, size_t bytes ) {
      if( rc == VOFX_OK && cudaMalloc( p, bytes ) != cudaSuccess )
        rc = fail( VOFX_ERR_CUDA, "cudaMalloc of context scratch failed" );
    };
    alloc( (void **) &ctx->mixedList, sizeof( int ) * (size_t) ctx->capacity );
    alloc( (void **) &ctx->slot, sizeof( int ) * (size_t) g.cells );
    alloc( (void **) &ctx->records, sizeof( FluxRecord ) * (size_t) ctx->capacity );
    alloc( (void **) &ctx->state, sizeof( StepState ) );
    if( desc->dim == 3 )
    {
      // clips left to the serial kernel (a node of the swept box exactly on the plane), stored from the end
      const long long tc = 6LL * ctx->capacity;
      ctx->taskCapacity = (int) ( tc < ( 1LL << 30 ) ? tc : ( 1LL << 30 ) );
      alloc( (void **) &ctx->tasks, sizeof( int ) * (size_t) ctx->taskCapacity );
      alloc( (void **) &ctx->clipTable, launch::clip_table_bytes() );
      alloc( (void **) &ctx->sweepTable, launch::sweep_table_bytes() );
      alloc( (void **) &ctx->serialCells, sizeof( int ) * ( (size_t) ctx->capacity + 2 * kMaxBandParts ) );   // one region per band part
      alloc( (void **) &ctx->rootTasks, launch::root_task_bytes() * (size_t) ctx->capacity );
      alloc( (void **) &ctx->costClass, (size_t) g.cells );
      alloc( (void **) &ctx->order, sizeof( int ) * (size_t) ctx->capacity );
      if( rc == VOFX_OK && cudaMemset( ctx->costClass, 0, (size_t) g.cells ) != cudaSuccess )
        rc = fail( VOFX_ERR_CUDA, "cudaMemset of context scratch failed" );
    }
    alloc( (void **) &ctx->curv1, sizeof( double ) * (size_t) ctx->capacity );
    alloc( (void **) &ctx->ok1, (size_t) ctx->capacity );
    alloc( (void **) &ctx->scalars, sizeof( double ) * 4 );
    alloc( (void **) &ctx->velDev, sizeof( Velocity ) );
    if( rc != VOFX_OK )
    {
      vofx_destroy( ctx );
      return rc;
    }
    StepState init;
    std::memset( &init, 0, sizeof( init ) );
    init.dtEst = DBL_MAX;
    init.epoch = 1;
    cudaMemcpy( ctx->state, &init, sizeof( init ), cudaMemcpyHostToDevice );
    // consumers index records[] / curv1[] with slot[ cell ]: never with uninitialised bits
    cudaMemset( ctx->slot, 0, sizeof( int ) * (size_t) g.cells );
    if( ctx->sweepTable )
    {
      std::vector< unsigned char > table( launch::sweep_table_bytes() );
      if( launch::make_sweep_table( table.data(), table.size() ) != 96 )
      {
        vofx_destroy( ctx );
        return fail( VOFX_ERR_STATE, "sweep table generation failed" );
      }
      cudaMemcpy( ctx->sweepTable, table.data(), table.size(), cudaMemcpyHostToDevice );
      if( const char *env = std::getenv( "VOFX_FLUX_LANES" ) )
        ctx->fluxLanes = ( std::atoi( env ) == 4 ) ? 4 : 8;
      if( const char *env = std::getenv( "VOFX_BAND_PARTS" ) )
        ctx->bandParts = std::atoi( env ) < 1 ? 1 : ( std::atoi( env ) > kMaxBandParts ? kMaxBandParts : std::atoi( env ) );
      if( std::getenv( "VOFX_SWZ_NO_SORT" ) )
        ctx->swzSortTasks = false;
      if( std::getenv( "VOFX_NO_COST_ORDER" ) )
        ctx->costOrder = false;
      if( const char *env = std::getenv( "VOFX_FLUX_BLOCKS_PER_SM" ) )
        ctx->fluxBlocksPerSM = ( std::atoi( env ) >= 1 ) ? std::atoi( env ) : 8;
      if( const char *env = std::getenv( "VOFX_UPDATE_BLOCKS_PER_SM" ) )
        ctx->updateBlocksPerSM = ( std::atoi( env ) >= 1 ) ? std::atoi( env ) : 8;
      if( const char *env = std::getenv( "VOFX_YOUNGS_BLOCKS_PER_SM" ) )
        ctx->youngsBlocksPerSM = ( std::atoi( env ) >= 2 ) ? std::atoi( env ) : 60;
      if( const char *env = std::getenv( "VOFX_LANES_PER_CELL" ) )
        ctx->lanesPerCell = ( std::atoi( env ) == 8 ) ? 8 : 4;
      if( const char *env = std::getenv( "VOFX_SWARTZ_LANES" ) )
        ctx->swartzLanes = ( std::atoi( env ) == 8 ) ? 8 : 4;
      if( const char *env = std::getenv( "VOFX_SWARTZ_MODE" ) )
        ctx->swartzMode = std::strcmp( env, "serial" ) == 0 ? 2 : ( std::strcmp( env, "cell" ) == 0 ? 1 : 0 );
    }
    if( ctx->clipTable )
    {
      std::vector< unsigned char > table( launch::clip_table_bytes() );
      launch::make_clip_table( table.data() );
      cudaMemcpy( ctx->clipTable, table.data(), table.size(), cudaMemcpyHostToDevice );
    }
    std::memset( &ctx->velHost, 0, sizeof( Velocity ) );
    *out = ctx;
    return VOFX_OK;
  }

  void vofx_destroy ( vofx_ctx *ctx )
  {
    if( !ctx )
      return;
    cudaSetDevice( ctx->device );
    cudaFree( ctx->mixedList );
    cudaFree( ctx->slot );
    cudaFree( ctx->records );
    cudaFree( ctx->tasks );
    cudaFree( ctx->clipTable );
    cudaFree( ctx->sweepTable );
    cudaFree( ctx->serialCells );
    cudaFree( ctx->touchedIdx );
    cudaFree( ctx->touchedVal );
    cudaFree( ctx->rootTasks );
    cudaFree( ctx->diagScratch );
    cudaFree( ctx->costClass );
    cudaFree( ctx->order );
    cudaFree( ctx->marks );
    cudaFree( ctx->segList );
    if( ctx->sideStream )
      cudaStreamDestroy( ctx->sideStream );
    if( ctx->evFork )
      cudaEventDestroy( ctx->evFork );
    if( ctx->evJoin )
      cudaEventDestroy( ctx->evJoin );
    for( int q = 0; q < kMaxBandParts; ++q )
    {
      if( ctx->partStream[ q ] )
        cudaStreamDestroy( ctx->partStream[ q ] );
      if( ctx->evBandJoin[ q ] )
        cudaEventDestroy( ctx->evBandJoin[ q ] );
    }
    if( ctx->evBandFork )
      cudaEventDestroy( ctx->evBandFork );
    cudaFree( ctx->swzCells );
    cudaFree( ctx->swzTasks );
    cudaFree( ctx->swzCounter );
    cudaFree( ctx->swzOrder );
    cudaFree( ctx->swzCellList );
    cudaFree( ctx->swzListCount );
    cudaFree( ctx->swzSort );
    cudaFree( ctx->ifBlockSums );
    cudaFree( ctx->ifCells );
    cudaFree( ctx->ifNverts );
    cudaFree( ctx->ifOffsets );
    cudaFree( ctx->ifTotal );
    cudaFree( ctx->ifVerts );
    cudaFree( ctx->ifCsr );
    cudaFree( ctx->curvList );
    cudaFree( ctx->curvListCount );
    cudaFree( ctx->changedIdx );
    cudaFree( ctx->changedVal );
    cudaFree( ctx->state );
    cudaFree( ctx->curv1 );
    cudaFree( ctx->ok1 );
    cudaFree( ctx->scalars );
    cudaFree( ctx->velDev );
    for( double *t : ctx->velTables )
      cudaFree( t );
    cudaFree( ctx->velFaces );
    cudaFree( ctx->velBandFaces );
    for( void *m : ctx->ipcMapped )
      cudaIpcCloseMemHandle( m );
    if( ctx->ownStream )
      cudaStreamDestroy( ctx->ownStream );
    if( ctx->window )
      cudaFree( ctx->window );        // the resident colour function lives inside the window
    else
      cudaFree( ctx->dU );
    cudaFree( ctx->dHs );
    cudaFree( ctx->dUpd );
    cudaFree( ctx->dKappa );
    cudaFree( ctx->dFlags );
    cudaFree( ctx->dValid );
    if( ctx->pinned )
      cudaFreeHost( ctx->pinned );
    if( ctx->stateHost )
      cudaFreeHost( ctx->stateHost );
    for( cudaEvent_t e : ctx->eventPool )
      cudaEventDestroy( e );
    delete ctx;
  }

  int vofx_set_stream ( vofx_ctx *ctx, void *cuda_stream )
  {
    if( !ctx )
      return fail( VOFX_ERR_ARGUMENT, "null context" );
    ctx->stream = (cudaStream_t) cuda_stream;
    if( !ctx->stream && ctx->ownStream )
      ctx->stream = ctx->ownStream;      // distributed contexts keep off the legacy default stream
    return VOFX_OK;
  }

  int64_t vofx_cells ( const vofx_ctx *ctx ) { return ctx ? (int64_t) ctx->grid.cells : 0; }

  int64_t vofx_launch_count ( const vofx_ctx *ctx ) { return ctx ? (int64_t) ctx->launches : 0; }

  int vofx_profile_enable ( vofx_ctx *ctx, int enable )
  {
    if( !ctx )
      return fail( VOFX_ERR_ARGUMENT, "null context" );
    ctx->profiling = enable != 0;
    ctx->profEntries.clear();
    ctx->eventsUsed = 0;
    return VOFX_OK;
  }

  int vofx_profile_read ( vofx_ctx *ctx, int max_entries, char *names, double *total_ms, int64_t *launches, int *n_entries )
  {
    if( !ctx || !names || !total_ms || !launches || !n_entries || max_entries < 1 )
      return fail( VOFX_ERR_ARGUMENT, "bad argument" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    int n = 0;
    for( const auto &entry : ctx->profEntries )
    {
      float ms = 0.0f;
      VOFX_CUDA( cudaEventElapsedTime( &ms, entry.begin, entry.end ) );
      int k = 0;
      for( ; k < n; ++k )
        if( std::strncmp( names + 48 * k, entry.name, 47 ) == 0 )
          break;
      if( k == n )
      {
        if( n == max_entries )
          continue;
        std::strncpy( names + 48 * k, entry.name, 47 );
        names[ 48 * k + 47 ] = 0;
        total_ms[ k ] = 0.0;
        launches[ k ] = 0;
        n++;
      }
      total_ms[ k ] += ms;
      launches[ k ]++;
    }
    *n_entries = n;
    ctx->profEntries.clear();
    ctx->eventsUsed = 0;
    return VOFX_OK;
  }

  // ---- velocity -----------------------------------------------------------------------------------------------------

  int vofx_velocity_clear ( vofx_ctx *ctx )
  {
    if( !ctx )
      return fail( VOFX_ERR_ARGUMENT, "null context" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    for( double *t : ctx->velTables )
      cudaFree( t );
    ctx->velTables.clear();
    std::memset( &ctx->velHost, 0, sizeof( Velocity ) );
    ctx->velocitySet = false;
    return uploadVelocity( ctx );
  }

  int vofx_velocity_set_component ( vofx_ctx *ctx, int component, int nterms, double coef, int time_factor, double period )
  {
    if( !ctx || component < 0 || component >= ctx->desc.dim || nterms < 0 || nterms > 2 )
      return fail( VOFX_ERR_ARGUMENT, "bad velocity component description" );
    if( time_factor != VOFX_TIME_CONST && time_factor != VOFX_TIME_COSPI )
      return fail( VOFX_ERR_ARGUMENT, "unknown time factor" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    ctx->velHost.nterms[ component ] = nterms;
    ctx->velHost.coef[ component ] = coef;
    ctx->velHost.timeKind[ component ] = time_factor;
    ctx->velHost.period[ component ] = period;
    ctx->velocitySet = true;
    return uploadVelocity( ctx );
  }

  int vofx_velocity_set_factor ( vofx_ctx *ctx, int component, int term, int axis, const double *at_centre, const double *at_lower_face, const double *at_upper_face )
  {
    if( !ctx || component < 0 || component >= ctx->desc.dim || term < 0 || term > 1 || axis < 0 || axis >= ctx->desc.dim
        || !at_centre || !at_lower_face || !at_upper_face )
      return fail( VOFX_ERR_ARGUMENT, "bad velocity factor description" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    const size_t n = (size_t) ctx->desc.n_global[ axis ];
    const double *src[ 3 ] = { at_centre, at_lower_face, at_upper_face };
    for( int k = 0; k < 3; ++k )
    {
      double *dev = nullptr;
      VOFX_CUDA( cudaMalloc( (void **) &dev, n * sizeof( double ) ) );
      ctx->velTables.push_back( dev );
      VOFX_CUDA( cudaMemcpy( dev, src[ k ], n * sizeof( double ), cudaMemcpyHostToDevice ) );
      ctx->velHost.tab[ component ][ term ][ axis ][ k ] = dev;
    }
    return uploadVelocity( ctx );
  }

  int vofx_velocity_set_faces ( vofx_ctx *ctx, const double *v_host )
  {
    if( !ctx || !v_host )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    const size_t count = (size_t) ctx->grid.cells * 2 * ctx->desc.dim * ctx->desc.dim;
    rc = ensureDevice( ctx->velFaces, count );
    if( rc )
      return rc;
    VOFX_CUDA( cudaMemcpyAsync( ctx->velFaces, v_host, count * sizeof( double ), cudaMemcpyHostToDevice, ctx->stream ) );
    ctx->velHost.faces = ctx->velFaces;
    ctx->velHost.bandFaces = nullptr;
    ctx->velHost.bandSlot = nullptr;
    ctx->velocitySet = true;
    return uploadVelocity( ctx );
  }

  int vofx_velocity_builtin ( vofx_ctx *ctx, int problem )
  {
    if( !ctx )
      return fail( VOFX_ERR_ARGUMENT, "null context" );
    int rc = vofx_velocity_clear( ctx );
    if( rc )
      return rc;
    const int dim = ctx->desc.dim;
    const double omega = 2 * M_PI / 10;
    switch( problem )
    {
    case VOFX_PROB_ROTATING_CIRCLE:
      // 2D rotatingcircle.hh:45-52: r = ( x1 - c1, c0 - x0 ) * omega ; 3D :92-101: rot = ( c1, -c0, 0 ) * omega with c = x - 0.5
      rc = vofx_velocity_set_component( ctx, 0, 1, omega, VOFX_TIME_CONST, 1.0 );
      rc = rc ? rc : setTerm( ctx, 0, 0, fOne, fMinusHalf, fOne );
      rc = rc ? rc : vofx_velocity_set_component( ctx, 1, 1, omega, VOFX_TIME_CONST, 1.0 );
      rc = rc ? rc : setTerm( ctx, 1, 0, dim == 2 ? fHalfMinus : fNegMinusHalf, fOne, fOne );
      if( dim == 3 )
        rc = rc ? rc : vofx_velocity_set_component( ctx, 2, 0, 0.0, VOFX_TIME_CONST, 1.0 );
      return rc;
    case VOFX_PROB_SFLOW:
      // sflow.hh:35-42,73-82: v = ( 0.25*( y0 + y1^3 ), -0.25*( y1 + y0^3 ) [, 0] ), y = 4x - 2
      rc = vofx_velocity_set_component( ctx, 0, 2, 0.25, VOFX_TIME_CONST, 1.0 );
      rc = rc ? rc : setTerm( ctx, 0, 0, fSflowLin, fOne, fOne );
      rc = rc ? rc : setTerm( ctx, 0, 1, fOne, fSflowCube, fOne );
      rc = rc ? rc : vofx_velocity_set_component( ctx, 1, 2, -0.25, VOFX_TIME_CONST, 1.0 );
      rc = rc ? rc : setTerm( ctx, 1, 0, fOne, fSflowLin, fOne );
      rc = rc ? rc : setTerm( ctx, 1, 1, fSflowCube, fOne, fOne );
      if( dim == 3 )
        rc = rc ? rc : vofx_velocity_set_component( ctx, 2, 0, 0.0, VOFX_TIME_CONST, 1.0 );
      return rc;
    case VOFX_PROB_SLOTTED_CYLINDER:
      if( dim != 2 )
        return fail( VOFX_ERR_ARGUMENT, "SlottedCylinder is a 2D problem" );
      // slottedcylinder.hh:47-55: rot = ( x1 - 0.5, -( x0 - 0.5 ) ) * omega
      rc = vofx_velocity_set_component( ctx, 0, 1, omega, VOFX_TIME_CONST, 1.0 );
      rc = rc ? rc : setTerm( ctx, 0, 0, fOne, fMinusHalf, fOne );
      rc = rc ? rc : vofx_velocity_set_component( ctx, 1, 1, omega, VOFX_TIME_CONST, 1.0 );
      rc = rc ? rc : setTerm( ctx, 1, 0, fNegMinusHalf, fOne, fOne );
      return rc;
    case VOFX_PROB_LINEAR_WALL:
      // linearwall.hh:27-31: ( 0.02, 0 [, 0] )
      rc = vofx_velocity_set_component( ctx, 0, 1, 0.02, VOFX_TIME_CONST, 1.0 );
      rc = rc ? rc : setTerm( ctx, 0, 0, fOne, fOne, fOne );
      for( int a = 1; a < dim; ++a )
        rc = rc ? rc : vofx_velocity_set_component( ctx, a, 0, 0.0, VOFX_TIME_CONST, 1.0 );
      return rc;
    case VOFX_PROB_LEVEQUE:
      if( dim == 2 )
      {
        // single vortex: ( sin^2(pi x) sin(2 pi y), -sin(2 pi x) sin^2(pi y) ) cos(pi t/8)
        rc = vofx_velocity_set_component( ctx, 0, 1, 1.0, VOFX_TIME_COSPI, 8.0 );
        rc = rc ? rc : setTerm( ctx, 0, 0, fSinSq, fSinTwo, fOne );
        rc = rc ? rc : vofx_velocity_set_component( ctx, 1, 1, 1.0, VOFX_TIME_COSPI, 8.0 );
        rc = rc ? rc : setTerm( ctx, 1, 0, fNegSinTwo, fSinSq, fOne );
      }
      else
      {
        // deformation field: ( 2 sin^2(pi x) sin(2 pi y) sin(2 pi z), -sin(2 pi x) sin^2(pi y) sin(2 pi z),
        //                      -sin(2 pi x) sin(2 pi y) sin^2(pi z) ) cos(pi t/3)
        rc = vofx_velocity_set_component( ctx, 0, 1, 1.0, VOFX_TIME_COSPI, 3.0 );
        rc = rc ? rc : setTerm( ctx, 0, 0, fTwoSinSq, fSinTwo, fSinTwo );
        rc = rc ? rc : vofx_velocity_set_component( ctx, 1, 1, 1.0, VOFX_TIME_COSPI, 3.0 );
        rc = rc ? rc : setTerm( ctx, 1, 0, fNegSinTwo, fSinSq, fSinTwo );
        rc = rc ? rc : vofx_velocity_set_component( ctx, 2, 1, 1.0, VOFX_TIME_COSPI, 3.0 );
        rc = rc ? rc : setTerm( ctx, 2, 0, fNegSinTwo, fSinTwo, fSinSq );
      }
      return rc;
    }
    return fail( VOFX_ERR_ARGUMENT, "unknown built-in problem" );
  }

  int vofx_face_velocity ( vofx_ctx *ctx, double time, int64_t cell, int face, double *v_host )
  {
    if( !ctx || !v_host || cell < 0 || cell >= ctx->grid.cells || face < 0 || face >= 2 * ctx->desc.dim )
      return fail( VOFX_ERR_ARGUMENT, "bad argument" );
    int rc = setDevice( ctx );
    rc = rc ? rc : requireVelocity( ctx );
    if( rc )
      return rc;
    profBegin( ctx );
    launch::test_face_velocity( ctx->stream, ctx->grid, ctx->velDev, time, cell, face, ctx->scalars );
    rc = checkLaunch( ctx, "k_test_face_velocity" );
    if( rc )
      return rc;
    double v[ 3 ];
    VOFX_CUDA( cudaMemcpyAsync( v, ctx->scalars, sizeof( v ), cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    for( int d = 0; d < ctx->desc.dim; ++d )
      v_host[ d ] = v[ d ];
    return VOFX_OK;
  }

  // ---- operators on device arrays -------------------------------------------------------------------------------------

  int vofx_flag ( vofx_ctx *ctx, const double *u, double eps, uint8_t *flags )
  {
    if( !ctx || !u || !flags )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = setDevice( ctx );
    if( flags == ctx->primedFlags || u == ctx->primedU )
      ctx->primedU = nullptr, ctx->primedFlags = nullptr;
    return rc ? rc : launchFlag( ctx, u, eps, flags, false );
  }

  int vofx_reconstruct ( vofx_ctx *ctx, int kind, const double *u, const uint8_t *flags, double *halfspaces, int max_iterations )
  {
    if( !ctx || !u || !flags || !halfspaces )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    // reconstructions.clear(), modifiedyoungs.hh:65
    VOFX_CUDA( cudaMemsetAsync( halfspaces, 0, sizeof( double ) * (size_t) ctx->grid.cells * ( ctx->desc.dim + 1 ), ctx->stream ) );
    rc = launchResetCounters( ctx );
    rc = rc ? rc : launchCompactFlags( ctx, flags );
    rc = rc ? rc : launchReconstruct( ctx, kind, u, flags, halfspaces, max_iterations );
    if( rc )
      return rc;
    StepState s;
    return readState( ctx, s );
  }

  int vofx_evolve ( vofx_ctx *ctx, const double *halfspaces, const uint8_t *flags, double time, double dt, double *update, double *dt_est )
  {
    if( !ctx || !halfspaces || !flags || !update || !dt_est )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = setDevice( ctx );
    rc = rc ? rc : requireVelocity( ctx );
    if( rc )
      return rc;
    const double td[ 2 ] = { time, dt };
    VOFX_CUDA( cudaMemcpyAsync( ctx->scalars, td, sizeof( td ), cudaMemcpyHostToDevice, ctx->stream ) );
    // update.clear(), evolution/evolution.hh:62
    VOFX_CUDA( cudaMemsetAsync( update, 0, sizeof( double ) * (size_t) ctx->grid.cells, ctx->stream ) );
    rc = launchResetCounters( ctx );
    rc = rc ? rc : launchCompactFlags( ctx, flags );
    rc = rc ? rc : launchFlux( ctx, halfspaces, flags, ctx->scalars, ctx->scalars + 1 );
    rc = rc ? rc : launchUpdate( ctx, flags, update, false );
    if( rc )
      return rc;
    StepState s;
    rc = readState( ctx, s );
    *dt_est = s.dtEst;
    return rc;
  }

  int vofx_axpy ( vofx_ctx *ctx, double *u, double a, const double *x )
  {
    if( !ctx || !u || !x )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    profBegin( ctx );
    launch::axpy( gridBlocks( ctx, ctx->grid.cells, 256, 8 ), ctx->stream, ctx->grid.cells, u, a, x );
    return checkLaunch( ctx, "k_axpy" );
  }

  int vofx_diagnostics ( vofx_ctx *ctx, const double *u, const double *u_exact, double *out_host )
  {
    if( !ctx || !u || !out_host )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    if( !ctx->diagScratch )
      VOFX_CUDA( cudaMalloc( &ctx->diagScratch, launch::diag_scratch_bytes() ) );
    const Grid &g = ctx->grid;
    const long long layer = g.cells / g.ns[ g.dim - 1 ];
    const long long first = layer * g.own0, n = layer * ( g.own1 - g.own0 );
    double volume = 1.0;
    for( int d = 0; d < g.dim; ++d )
      volume *= g.h[ d ];
    profBegin( ctx );
    const double *out = launch::diagnostics( ctx->stream, n, u + first, u_exact ? u_exact + first : nullptr, volume, ctx->diagScratch );
    ctx->launches += 1;
    rc = checkLaunch( ctx, "k_diagnostics" );
    if( rc )
      return rc;
    VOFX_CUDA( cudaMemcpyAsync( out_host, out, 5 * sizeof( double ), cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    return VOFX_OK;
  }

  // circleInterpolation on the device (2D), interpolation/circleinterpolation.hh:140-153
  int vofx_init_circle ( vofx_ctx *ctx, const double *center, double radius, double *u )
  {
    if( !ctx || !center || !u )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    if( ctx->desc.dim != 2 )
      return fail( VOFX_ERR_ARGUMENT, "circleInterpolation is a 2D initialiser (as in the reference)" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    if( u == ctx->primedU )
      ctx->primedU = nullptr, ctx->primedFlags = nullptr;
    profBegin( ctx );
    launch::circle_init( gridBlocks( ctx, ctx->grid.cells, 128, 16 ), ctx->stream, ctx->grid, center[ 0 ], center[ 1 ], radius, u );
    return checkLaunch( ctx, "k_circle_init" );
  }

  // l1error( gridView, reconstructions, flags, RotatingCircle< double, 2 >, time ), test/errors.hh:62-95, over the owned cells
  int vofx_l1error_circle_device ( vofx_ctx *ctx, const double *halfspaces, const uint8_t *flags, const double *center, double radius, double *error_host )
  {
    if( !ctx || !halfspaces || !flags || !center || !error_host )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    if( ctx->desc.dim != 2 )
      return fail( VOFX_ERR_ARGUMENT, "the circle error is defined in 2D (as in the reference)" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    if( !ctx->diagScratch )
      VOFX_CUDA( cudaMalloc( &ctx->diagScratch, launch::diag_scratch_bytes() ) );
    profBegin( ctx );
    const double *out = launch::l1error_circle( ctx->stream, ctx->grid, flags, halfspaces, center[ 0 ], center[ 1 ], radius, ctx->diagScratch );
    ctx->launches += 1;
    rc = checkLaunch( ctx, "k_l1error_circle" );
    if( rc )
      return rc;
    VOFX_CUDA( cudaMemcpyAsync( error_host, out, sizeof( double ), cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    return VOFX_OK;
  }

  int vofx_l1error_circle ( vofx_ctx *ctx, const double *center, double radius, double *error_host )
  {
    if( !ctx || !ctx->dHs || !ctx->dFlags )
      return fail( VOFX_ERR_STATE, "vofx_l1error_circle reads the flags and reconstructions of the last pass of the context-resident loop" );
    return vofx_l1error_circle_device( ctx, ctx->dHs, ctx->dFlags, center, radius, error_host );
  }

  int vofx_curvature ( vofx_ctx *ctx, const double *u, const double *halfspaces, const uint8_t *flags, double *kappa, uint8_t *valid )
  {
    if( !ctx || !u || !halfspaces || !flags || !kappa )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = setDevice( ctx );
    rc = rc ? rc : launchResetCounters( ctx );
    rc = rc ? rc : launchCompactFlags( ctx, flags );
    rc = rc ? rc : launchCurvature( ctx, u, halfspaces, flags, kappa, valid, false );
    if( rc )
      return rc;
    StepState s;
    return readState( ctx, s );
  }

  // SwartzCurvature (2D), curvature/swartzcurvature.hh:41-173
  int vofx_curvature_swartz ( vofx_ctx *ctx, const double *halfspaces, const uint8_t *flags, double *kappa )
  {
    if( !ctx || !halfspaces || !flags || !kappa )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    if( ctx->desc.dim != 2 )
      return fail( VOFX_ERR_ARGUMENT, "SwartzCurvature is implemented in 2D only (as in the reference)" );
    int rc = setDevice( ctx );
    rc = rc ? rc : launchResetCounters( ctx );
    rc = rc ? rc : launchCompactFlags( ctx, flags );
    if( rc )
      return rc;
    VOFX_CUDA( cudaMemsetAsync( kappa, 0, sizeof( double ) * (size_t) ctx->grid.cells, ctx->stream ) );
    profBegin( ctx );
    launch::swartz_curvature2( ctx->numSMs * 8, ctx->stream, ctx->grid, halfspaces, flags, ctx->mixedList, ctx->slot, ctx->state, ctx->capacity, ctx->curv1, kappa );
    ctx->launches++;
    rc = checkLaunch( ctx, "k_swartz_curvature" );
    if( rc )
      return rc;
    StepState s;
    return readState( ctx, s );
  }

  // ---- fused time loop ---------------------------------------------------------------------------------------------------

  int vofx_step_begin ( vofx_ctx *ctx, double time, double dt )
  {
    if( !ctx )
      return fail( VOFX_ERR_ARGUMENT, "null context" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    StepState init;
    std::memset( &init, 0, sizeof( init ) );
    init.time = time;
    init.dt = dt;
    init.dtEst = DBL_MAX;
    init.epoch = 1;
    VOFX_CUDA( cudaMemcpyAsync( ctx->state, &init, sizeof( init ), cudaMemcpyHostToDevice, ctx->stream ) );
    if( ctx->window )
      VOFX_CUDA( cudaMemsetAsync( static_cast< char * >( ctx->window ) + ctx->slotsOffset, 0, sizeof( CommSlots ), ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    ctx->primedU = nullptr;       // the first pass of a loop flags densely
    ctx->primedFlags = nullptr;
    return VOFX_OK;
  }

  int vofx_set_incremental_flagging ( vofx_ctx *ctx, int enable )
  {
    if( !ctx )
      return fail( VOFX_ERR_ARGUMENT, "null context" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    ctx->primedU = nullptr;
    ctx->primedFlags = nullptr;
    if( enable && !ctx->marks )
    {
      const Grid &g = ctx->grid;
      ctx->segsPerRow = ( g.ns[ 0 ] + 31 ) / 32;
      ctx->nSegments = (long long) ctx->segsPerRow * g.ns[ 1 ] * g.ns[ 2 ];
      if( ctx->nSegments >= ( 1LL << 31 ) )
        return fail( VOFX_ERR_ARGUMENT, "too many row segments" );
      rc = ensureDevice( ctx->marks, (size_t) ctx->nSegments + 16 );     // k_scan_marks reads 16 marks per thread
      rc = rc ? rc : ensureDevice( ctx->segList, (size_t) ctx->nSegments );
      if( rc )
        return rc;
      VOFX_CUDA( cudaMemsetAsync( ctx->marks, 0, (size_t) ctx->nSegments + 16, ctx->stream ) );
    }
    ctx->incremental = enable != 0;
    return VOFX_OK;
  }

  // one loop pass in three phases so that a multi-GPU driver can overlap its communication (decomposition.py):
  //   0  flags of the interior layers (no halo value is read)            -- while the colour halos are in flight
  //   1  flags of the layers next to / inside the halos, reconstruction, [curvature,] fluxes (time-step estimate)
  //   2  update                                                         -- while the time-step estimate is MIN-reduced
  int vofx_step_phase ( vofx_ctx *ctx, const vofx_step_params *p, double *u, uint8_t *flags, double *halfspaces, double *kappa, int phase )
  {
    if( !ctx || !p || !u || !flags || !halfspaces || ( p->with_curvature && !kappa ) || phase < 0 || phase > 2 )
      return fail( VOFX_ERR_ARGUMENT, "bad argument" );
    int rc = setDevice( ctx );
    rc = rc ? rc : requireVelocity( ctx );
    if( rc )
      return rc;
    if( phase == 0 )
      return ctx->desc.halo > 0 ? launchStepFlags( ctx, p, u, flags, 0 ) : VOFX_OK;
    if( phase == 1 )
    {
      // algorithm.hh:75-81: flagOperator, reconstructionOperator, evolutionOperator
      if( ctx->desc.halo == 0 )
        rc = launchStepFlags( ctx, p, u, flags, 0 );
      rc = rc ? rc : launchStepFlags( ctx, p, u, flags, 1 );
      rc = rc ? rc : forkSide( ctx, u, flags );
      return rc ? rc : launchBand( ctx, p, u, flags, halfspaces, kappa );
    }
    rc = launchUpdate( ctx, flags, u, true );   // uh.axpy( 1.0, update ), algorithm.hh:83
    if( !rc )
      notePassIssued( ctx, u, flags );
    return rc;
  }

  int vofx_step_local ( vofx_ctx *ctx, const vofx_step_params *p, double *u, uint8_t *flags, double *halfspaces, double *kappa )
  {
    if( !ctx || !p || !u || !flags || !halfspaces || ( p->with_curvature && !kappa ) )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = setDevice( ctx );
    rc = rc ? rc : requireVelocity( ctx );
    // algorithm.hh:75-83: flagOperator, reconstructionOperator, evolutionOperator, uh.axpy( 1.0, update )
    rc = rc ? rc : launchStepFlags( ctx, p, u, flags, 0 );
    rc = rc ? rc : launchStepFlags( ctx, p, u, flags, 1 );
    rc = rc ? rc : forkSide( ctx, u, flags );
    rc = rc ? rc : launchBand( ctx, p, u, flags, halfspaces, kappa );
    rc = rc ? rc : launchUpdate( ctx, flags, u, true );
    if( !rc )
      notePassIssued( ctx, u, flags );
    return rc;
  }

  int vofx_step_finish ( vofx_ctx *ctx, const vofx_step_params *p )
  {
    if( !ctx || !p )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = setDevice( ctx );
    rc = rc ? rc : joinSide( ctx );
    if( rc )
      return rc;
    profBegin( ctx );
    launch::finalize( ctx->stream, ctx->state, p->cfl );
    return checkLaunch( ctx, "k_finalize" );
  }

  int vofx_step ( vofx_ctx *ctx, const vofx_step_params *p, double *u, uint8_t *flags, double *halfspaces, double *kappa )
  {
    int rc = vofx_step_local( ctx, p, u, flags, halfspaces, kappa );
    return rc ? rc : vofx_step_finish( ctx, p );
  }

  double *vofx_dt_est_device ( vofx_ctx *ctx ) { return ctx ? &ctx->state->dtEst : nullptr; }

  int vofx_step_state ( vofx_ctx *ctx, double *time, double *dt, double *dt_est, int64_t *mixed_cells )
  {
    if( !ctx )
      return fail( VOFX_ERR_ARGUMENT, "null context" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    StepState s;
    rc = readState( ctx, s );
    if( time )
      *time = s.time;
    if( dt )
      *dt = s.dt;
    if( dt_est )
      *dt_est = s.dtEstLast;
    if( mixed_cells )
      *mixed_cells = s.mixedCountLast;
    return rc;
  }

  int vofx_band_list_fetch ( vofx_ctx *ctx, int32_t *cells_host, int64_t capacity, int64_t *n_cells )
  {
    if( !ctx || !n_cells || ( capacity > 0 && !cells_host ) )
      return fail( VOFX_ERR_ARGUMENT, "bad argument" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    StepState s;
    rc = readState( ctx, s );
    if( rc )
      return rc;
    *n_cells = s.mixedCountLast;
    const size_t n = (size_t) ( s.mixedCountLast < capacity ? s.mixedCountLast : capacity );
    if( n > 0 )
    {
      VOFX_CUDA( cudaMemcpyAsync( cells_host, ctx->mixedList, n * sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream ) );
      VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    }
    return VOFX_OK;
  }

  int vofx_band_list_fetch_in_flight ( vofx_ctx *ctx, int32_t *cells_host, int64_t capacity, int64_t *n_cells )
  {
    if( !ctx || !n_cells || ( capacity > 0 && !cells_host ) )
      return fail( VOFX_ERR_ARGUMENT, "bad argument" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    StepState s;
    rc = readState( ctx, s );
    if( rc )
      return rc;
    *n_cells = s.mixedCount;
    const size_t n = (size_t) ( s.mixedCount < capacity ? s.mixedCount : capacity );
    if( n > 0 )
    {
      VOFX_CUDA( cudaMemcpyAsync( cells_host, ctx->mixedList, n * sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream ) );
      VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    }
    return VOFX_OK;
  }

  int vofx_velocity_set_band_faces ( vofx_ctx *ctx, int64_t n_cells, const double *v_host )
  {
    if( !ctx || n_cells < 0 || ( n_cells > 0 && !v_host ) )
      return fail( VOFX_ERR_ARGUMENT, "bad argument" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    const size_t count = (size_t) n_cells * 2 * ctx->desc.dim * ctx->desc.dim;
    rc = growDevice( ctx->velBandFaces, ctx->velBandCap, count + 1 );
    if( rc )
      return rc;
    if( count )
      VOFX_CUDA( cudaMemcpyAsync( ctx->velBandFaces, v_host, count * sizeof( double ), cudaMemcpyHostToDevice, ctx->stream ) );
    ctx->velHost.bandFaces = ctx->velBandFaces;
    ctx->velHost.bandSlot = ctx->slot;
    ctx->velocitySet = true;
    return uploadVelocity( ctx );
  }

  int vofx_step_counters ( vofx_ctx *ctx, int64_t *out )
  {
    if( !ctx || !out )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    StepState s;
    rc = readState( ctx, s );
    out[ 0 ] = s.mixedCount;
    out[ 1 ] = s.serialCount;
    out[ 2 ] = s.serialLocate + s.serialLocateMore[ 0 ] + s.serialLocateMore[ 1 ] + s.serialLocateMore[ 2 ];
    out[ 3 ] = s.rootCount;
    return rc;
  }

  // ---- host-buffer variants ---------------------------------------------------------------------------------------------

  static int stageIn ( vofx_ctx *ctx, void *dev, const void *host, size_t bytes )
  {
    int rc = ensurePinned( ctx, bytes );
    if( rc )
      return rc;
    std::memcpy( ctx->pinned, host, bytes );
    VOFX_CUDA( cudaMemcpyAsync( dev, ctx->pinned, bytes, cudaMemcpyHostToDevice, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    return VOFX_OK;
  }

  static int stageOut ( vofx_ctx *ctx, void *host, const void *dev, size_t bytes )
  {
    int rc = ensurePinned( ctx, bytes );
    if( rc )
      return rc;
    VOFX_CUDA( cudaMemcpyAsync( ctx->pinned, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    std::memcpy( host, ctx->pinned, bytes );
    return VOFX_OK;
  }

  static int ensureMirrors ( vofx_ctx *ctx )
  {
    const size_t N = (size_t) ctx->grid.cells;
    int rc = setDevice( ctx );
    rc = rc ? rc : ensureDevice( ctx->dU, N );
    rc = rc ? rc : ensureDevice( ctx->dFlags, N );
    rc = rc ? rc : ensureDevice( ctx->dHs, N * ( ctx->desc.dim + 1 ) );
    return rc;
  }

  int vofx_flag_host ( vofx_ctx *ctx, const double *u, double eps, uint8_t *flags )
  {
    if( !ctx || !u || !flags )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    const size_t N = (size_t) ctx->grid.cells;
    int rc = ensureMirrors( ctx );
    ctx->hostMirrorOf = nullptr;
    rc = rc ? rc : stageIn( ctx, ctx->dU, u, N * sizeof( double ) );
    rc = rc ? rc : vofx_flag( ctx, ctx->dU, eps, ctx->dFlags );
    return rc ? rc : stageOut( ctx, flags, ctx->dFlags, N );
  }

  int vofx_reconstruct_host ( vofx_ctx *ctx, int kind, const double *u, const uint8_t *flags, double *halfspaces, int max_iterations )
  {
    if( !ctx || !u || !flags || !halfspaces )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    const size_t N = (size_t) ctx->grid.cells;
    int rc = ensureMirrors( ctx );
    ctx->hostMirrorOf = nullptr;
    rc = rc ? rc : stageIn( ctx, ctx->dU, u, N * sizeof( double ) );
    rc = rc ? rc : stageIn( ctx, ctx->dFlags, flags, N );
    rc = rc ? rc : vofx_reconstruct( ctx, kind, ctx->dU, ctx->dFlags, ctx->dHs, max_iterations );
    return rc ? rc : stageOut( ctx, halfspaces, ctx->dHs, N * ( ctx->desc.dim + 1 ) * sizeof( double ) );
  }

  int vofx_evolve_host ( vofx_ctx *ctx, const double *halfspaces, const uint8_t *flags, double time, double dt, double *update, double *dt_est )
  {
    if( !ctx || !halfspaces || !flags || !update || !dt_est )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    const size_t N = (size_t) ctx->grid.cells;
    int rc = ensureMirrors( ctx );
    rc = rc ? rc : ensureDevice( ctx->dUpd, N );
    rc = rc ? rc : stageIn( ctx, ctx->dHs, halfspaces, N * ( ctx->desc.dim + 1 ) * sizeof( double ) );
    rc = rc ? rc : stageIn( ctx, ctx->dFlags, flags, N );
    rc = rc ? rc : vofx_evolve( ctx, ctx->dHs, ctx->dFlags, time, dt, ctx->dUpd, dt_est );
    return rc ? rc : stageOut( ctx, update, ctx->dUpd, N * sizeof( double ) );
  }

  int vofx_curvature_host ( vofx_ctx *ctx, const double *u, const double *halfspaces, const uint8_t *flags, double *kappa, uint8_t *valid )
  {
    if( !ctx || !u || !halfspaces || !flags || !kappa )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    const size_t N = (size_t) ctx->grid.cells;
    int rc = ensureMirrors( ctx );
    rc = rc ? rc : ensureDevice( ctx->dKappa, N );
    rc = rc ? rc : ensureDevice( ctx->dValid, N );
    ctx->hostMirrorOf = nullptr;
    rc = rc ? rc : stageIn( ctx, ctx->dU, u, N * sizeof( double ) );
    rc = rc ? rc : stageIn( ctx, ctx->dHs, halfspaces, N * ( ctx->desc.dim + 1 ) * sizeof( double ) );
    rc = rc ? rc : stageIn( ctx, ctx->dFlags, flags, N );
    rc = rc ? rc : vofx_curvature( ctx, ctx->dU, ctx->dHs, ctx->dFlags, ctx->dKappa, ctx->dValid );
    rc = rc ? rc : stageOut( ctx, kappa, ctx->dKappa, N * sizeof( double ) );
    if( !rc && valid )
      rc = stageOut( ctx, valid, ctx->dValid, N );
    return rc;
  }

  int vofx_curvature_swartz_host ( vofx_ctx *ctx, const double *halfspaces, const uint8_t *flags, double *kappa )
  {
    if( !ctx || !halfspaces || !flags || !kappa )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    const size_t N = (size_t) ctx->grid.cells;
    int rc = ensureMirrors( ctx );
    rc = rc ? rc : ensureDevice( ctx->dKappa, N );
    rc = rc ? rc : stageIn( ctx, ctx->dHs, halfspaces, N * ( ctx->desc.dim + 1 ) * sizeof( double ) );
    rc = rc ? rc : stageIn( ctx, ctx->dFlags, flags, N );
    rc = rc ? rc : vofx_curvature_swartz( ctx, ctx->dHs, ctx->dFlags, ctx->dKappa );
    return rc ? rc : stageOut( ctx, kappa, ctx->dKappa, N * sizeof( double ) );
  }

  // !declared: the whole host colour function goes host -> device.  Otherwise the device mirror of the previous
  // vofx_step_host* call on this array is current except for the nTouched cells the caller lists (it changed them on the
  // host since that call): only those values are uploaded, and flags are maintained incrementally
  static int stepHostImpl ( vofx_ctx *ctx, const vofx_step_params *p, double *u, uint8_t *flags, bool declared, const int32_t *touched, int64_t nTouched )
  {
    if( !ctx || !p || !u || ( nTouched > 0 && !touched ) || nTouched < 0 )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    const size_t N = (size_t) ctx->grid.cells;
    const bool sparse = declared && ctx->hostMirrorOf == u;     // else: first call on this array, or the mirror was used otherwise
    int rc = ensureMirrors( ctx );
    if( !rc && p->with_curvature )
      rc = ensureDevice( ctx->dKappa, N );
    // the update touches the mixed cells and their face neighbours only: the host colour function is brought up to
    // date from the list of (cell, new value) pairs instead of a second full-size copy across PCIe
    if( !rc && !ctx->changedIdx )
    {
      const long long cap = 7LL * ctx->capacity < (long long) N ? 7LL * ctx->capacity : (long long) N;
      ctx->changedCapacity = (int) ( cap < ( 1LL << 30 ) ? cap : ( 1LL << 30 ) );
      rc = ensureDevice( ctx->changedIdx, (size_t) ctx->changedCapacity );
      rc = rc ? rc : ensureDevice( ctx->changedVal, (size_t) ctx->changedCapacity );
    }
    if( rc )
      return rc;
    if( !sparse )
    {
      // u may be caller-pinned memory: copy directly, no staging copy
      VOFX_CUDA( cudaMemcpyAsync( ctx->dU, u, N * sizeof( double ), cudaMemcpyHostToDevice, ctx->stream ) );
      ctx->lastH2D = (long long) ( N * sizeof( double ) );
      ctx->primedU = nullptr;       // the caller's colour function was copied in again: dense flags
      ctx->primedFlags = nullptr;
    }
    else
    {
      ctx->lastH2D = 0;
      if( nTouched > 0 )
      {
        // ( index, value ) pairs through the pinned staging buffer, scattered by a kernel; a cell the caller changed may
        // change flags anywhere near it: the next pass flags densely (incremental flagging only covers the band's own writes)
        const size_t bytes = (size_t) nTouched * ( sizeof( int ) + sizeof( double ) ) + 16;
        rc = ensurePinned( ctx, bytes );
        rc = rc ? rc : growDevice( ctx->touchedIdx, ctx->touchedCap, (size_t) nTouched );
        rc = rc ? rc : growDevice( ctx->touchedVal, ctx->touchedValCap, (size_t) nTouched );
        if( rc )
          return rc;
        VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );     // the staging buffer may still feed the previous call's copies
        double *hostVal = static_cast< double * >( ctx->pinned );
        int *hostIdx = reinterpret_cast< int * >( hostVal + nTouched );
        for( int64_t i = 0; i < nTouched; ++i )
        {
          if( touched[ i ] < 0 || (size_t) touched[ i ] >= N )
            return fail( VOFX_ERR_ARGUMENT, "vofx_step_host_sparse: touched cell out of range" );
          hostIdx[ i ] = touched[ i ];
          hostVal[ i ] = u[ touched[ i ] ];
        }
        VOFX_CUDA( cudaMemcpyAsync( ctx->touchedVal, hostVal, (size_t) nTouched * sizeof( double ), cudaMemcpyHostToDevice, ctx->stream ) );
        VOFX_CUDA( cudaMemcpyAsync( ctx->touchedIdx, hostIdx, (size_t) nTouched * sizeof( int ), cudaMemcpyHostToDevice, ctx->stream ) );
        launch::scatter_values( gridBlocks( ctx, nTouched, 256, 8 ), ctx->stream, nTouched, ctx->touchedIdx, ctx->touchedVal, ctx->dU );
        ctx->launches++;
        VOFX_CUDA( cudaGetLastError() );
        ctx->lastH2D = (long long) ( nTouched * ( sizeof( int ) + sizeof( double ) ) );
        ctx->primedU = nullptr;
        ctx->primedFlags = nullptr;
      }
      else if( !ctx->incremental )
      {
        rc = vofx_set_incremental_flagging( ctx, 1 );       // nothing but the passes changes the mirror: re-flag the band's neighbourhood only
        if( rc )
          return rc;
      }
    }
    ctx->lastD2H = (long long) sizeof( StepState ) + ( flags ? (long long) N : 0 );
    const bool tracking = ctx->trackChanges;
    ctx->trackChanges = true;
    ctx->hostMirrorOf = nullptr;  // until this call has brought the host array up to date
    rc = vofx_step( ctx, p, ctx->dU, ctx->dFlags, ctx->dHs, ctx->dKappa );
    ctx->trackChanges = tracking;
    if( rc )
      return rc;
    if( flags )
      VOFX_CUDA( cudaMemcpyAsync( flags, ctx->dFlags, N, cudaMemcpyDeviceToHost, ctx->stream ) );
    StepState s;
    rc = readState( ctx, s );
    if( rc )
      return rc;
    const size_t n = (size_t) s.changedCountLast;
    if( n > (size_t) ctx->changedCapacity )
    {
      VOFX_CUDA( cudaMemcpyAsync( u, ctx->dU, N * sizeof( double ), cudaMemcpyDeviceToHost, ctx->stream ) );
      VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
      ctx->lastD2H += (long long) ( N * sizeof( double ) );
      ctx->hostMirrorOf = u;
      return VOFX_OK;
    }
    ctx->hostMirrorOf = u;
    if( n == 0 )
      return VOFX_OK;
    ctx->lastD2H += (long long) ( n * ( sizeof( int ) + sizeof( double ) ) );
    // grow the staging buffer geometrically: re-pinning host memory every step (the band grows) costs milliseconds
    size_t want = n * ( sizeof( int ) + sizeof( double ) ) + 16;
    if( want > ctx->pinnedBytes )
      want = ( 2 * want > ( size_t( 64 ) << 20 ) ) ? 2 * want : ( size_t( 64 ) << 20 );
    rc = ensurePinned( ctx, want );
    if( rc )
      return rc;
    double *hostVal = static_cast< double * >( ctx->pinned );
    int *hostIdx = reinterpret_cast< int * >( hostVal + n );
    VOFX_CUDA( cudaMemcpyAsync( hostVal, ctx->changedVal, n * sizeof( double ), cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaMemcpyAsync( hostIdx, ctx->changedIdx, n * sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    for( size_t i = 0; i < n; ++i )
      u[ hostIdx[ i ] ] = hostVal[ i ];
    return VOFX_OK;
  }

  int vofx_step_host ( vofx_ctx *ctx, const vofx_step_params *p, double *u, uint8_t *flags )
  {
    return stepHostImpl( ctx, p, u, flags, false, nullptr, 0 );
  }

  int vofx_step_host_sparse ( vofx_ctx *ctx, const vofx_step_params *p, double *u, uint8_t *flags, const int32_t *touched, int64_t n_touched )
  {
    return stepHostImpl( ctx, p, u, flags, true, touched, n_touched );
  }

  // Sparse write-back for drivers that run the pass themselves (slab decomposition): vofx_track_changes( 1 ) makes every
  // following update record the (cell, new value) pairs it writes; vofx_changed_fetch copies the pairs of the LAST
  // finished pass into the caller's host colour function (indices are storage indices of this context, shifted by
  // `index_offset` cells) and returns their number.
  int vofx_track_changes ( vofx_ctx *ctx, int enable )
  {
    if( !ctx )
      return fail( VOFX_ERR_ARGUMENT, "null context" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    if( enable && !ctx->changedIdx )
    {
      const size_t N = (size_t) ctx->grid.cells;
      const long long cap = 7LL * ctx->capacity < (long long) N ? 7LL * ctx->capacity : (long long) N;
      ctx->changedCapacity = (int) ( cap < ( 1LL << 30 ) ? cap : ( 1LL << 30 ) );
      rc = ensureDevice( ctx->changedIdx, (size_t) ctx->changedCapacity );
      rc = rc ? rc : ensureDevice( ctx->changedVal, (size_t) ctx->changedCapacity );
      if( rc )
        return rc;
    }
    ctx->trackChanges = enable != 0;
    return VOFX_OK;
  }

  int vofx_changed_fetch ( vofx_ctx *ctx, double *u_host, int64_t index_offset, int64_t *n_changed )
  {
    if( !ctx || !u_host || !n_changed )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    if( !ctx->changedIdx )
      return fail( VOFX_ERR_STATE, "vofx_track_changes has not been enabled" );
    StepState s;
    int rc = readState( ctx, s );
    if( rc )
      return rc;
    const size_t n = (size_t) s.changedCountLast;
    *n_changed = (int64_t) n;
    if( n > (size_t) ctx->changedCapacity )
      return fail( VOFX_ERR_STATE, "changed-cell list overflow" );
    ctx->lastD2H = (long long) sizeof( StepState ) + (long long) ( n * ( sizeof( int ) + sizeof( double ) ) );
    if( n == 0 )
      return VOFX_OK;
    size_t want = n * ( sizeof( int ) + sizeof( double ) ) + 16;
    if( want > ctx->pinnedBytes )
      want = ( 2 * want > ( size_t( 64 ) << 20 ) ) ? 2 * want : ( size_t( 64 ) << 20 );
    rc = ensurePinned( ctx, want );
    if( rc )
      return rc;
    double *hostVal = static_cast< double * >( ctx->pinned );
    int *hostIdx = reinterpret_cast< int * >( hostVal + n );
    VOFX_CUDA( cudaMemcpyAsync( hostVal, ctx->changedVal, n * sizeof( double ), cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaMemcpyAsync( hostIdx, ctx->changedIdx, n * sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    for( size_t i = 0; i < n; ++i )
      u_host[ (int64_t) hostIdx[ i ] + index_offset ] = hostVal[ i ];
    return VOFX_OK;
  }

  int vofx_host_transfer_bytes ( const vofx_ctx *ctx, int64_t *h2d, int64_t *d2h )
  {
    if( !ctx || !h2d || !d2h )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    *h2d = ctx->lastH2D;
    *d2h = ctx->lastD2H;
    return VOFX_OK;
  }

  // context-resident colour function: upload once, step on the device, download when the host needs it
  int vofx_color_upload ( vofx_ctx *ctx, const double *u )
  {
    if( !ctx || !u )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = ensureMirrors( ctx );
    if( rc )
      return rc;
    ctx->hostMirrorOf = nullptr;
    VOFX_CUDA( cudaMemcpyAsync( ctx->dU, u, (size_t) ctx->grid.cells * sizeof( double ), cudaMemcpyHostToDevice, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    ctx->primedU = nullptr;
    ctx->primedFlags = nullptr;
    return VOFX_OK;
  }

  int vofx_color_download ( vofx_ctx *ctx, double *u, uint8_t *flags )
  {
    if( !ctx || !u )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = ensureMirrors( ctx );
    if( rc )
      return rc;
    VOFX_CUDA( cudaMemcpyAsync( u, ctx->dU, (size_t) ctx->grid.cells * sizeof( double ), cudaMemcpyDeviceToHost, ctx->stream ) );
    if( flags )
      VOFX_CUDA( cudaMemcpyAsync( flags, ctx->dFlags, (size_t) ctx->grid.cells, cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    return VOFX_OK;
  }

  // ---- checkpoint / restart in the byte format of the reference's BinaryWriter (example/binarywriter.hh:36-60) ------------------
  // binaryStream << time; uh.write( binaryStream ) (colorfunction.hh:40-44): the time and then one double per cell in the grid
  // view's iteration order, raw little-endian IEEE doubles (Dune::Fem::BinaryFileOutStream writes primitives verbatim; dune-fem
  // is a third-party dependency that is not vendored: that layout is restated from its public documentation).  A rank writes
  // its OWNED cells (x fastest), as a DUNE rank writes the cells of its grid view.
  static void checkpoint_name ( char *out, size_t bytes, const char *prefix, int level, int step, int rank, int world )
  {
    // "s" setw(4) size "-p" setw(4) rank "-" prefix "-" level "-" setw(5) step ".bin", fill '0' (binarywriter.hh:46-51)
    std::snprintf( out, bytes, "s%04d-p%04d-%s-%d-%05d.bin", world, rank, prefix, level, step );
  }

  int vofx_checkpoint_name ( char *out, int64_t bytes, const char *prefix, int level, int step, int rank, int world )
  {
    if( !out || bytes < 32 || !prefix )
      return fail( VOFX_ERR_ARGUMENT, "bad argument" );
    checkpoint_name( out, (size_t) bytes, prefix, level, step, rank, world );
    return VOFX_OK;
  }

  int vofx_checkpoint_write ( vofx_ctx *ctx, const char *path, double time )
  {
    if( !ctx || !path )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = ensureMirrors( ctx );
    if( rc )
      return rc;
    const Grid &g = ctx->grid;
    const long long layer = g.cells / g.ns[ g.dim - 1 ];
    const size_t first = (size_t) ( layer * g.own0 ), n = (size_t) ( layer * ( g.own1 - g.own0 ) );
    rc = ensurePinned( ctx, n * sizeof( double ) );
    if( rc )
      return rc;
    VOFX_CUDA( cudaMemcpyAsync( ctx->pinned, ctx->dU + first, n * sizeof( double ), cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    FILE *f = std::fopen( path, "wb" );
    if( !f )
      return fail( VOFX_ERR_ARGUMENT, std::string( "cannot open " ) + path + " for writing" );
    const bool ok = std::fwrite( &time, sizeof( double ), 1, f ) == 1 && std::fwrite( ctx->pinned, sizeof( double ), n, f ) == n;
    if( std::fclose( f ) != 0 || !ok )
      return fail( VOFX_ERR_ARGUMENT, std::string( "short write to " ) + path );
    return VOFX_OK;
  }

  int vofx_checkpoint_read ( vofx_ctx *ctx, const char *path, double *time )
  {
    if( !ctx || !path || !time )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = ensureMirrors( ctx );
    if( rc )
      return rc;
    const Grid &g = ctx->grid;
    const long long layer = g.cells / g.ns[ g.dim - 1 ];
    const size_t first = (size_t) ( layer * g.own0 ), n = (size_t) ( layer * ( g.own1 - g.own0 ) );
    rc = ensurePinned( ctx, n * sizeof( double ) );
    if( rc )
      return rc;
    FILE *f = std::fopen( path, "rb" );
    if( !f )
      return fail( VOFX_ERR_ARGUMENT, std::string( "cannot open " ) + path );
    const bool ok = std::fread( time, sizeof( double ), 1, f ) == 1 && std::fread( ctx->pinned, sizeof( double ), n, f ) == n;
    char extra;
    const bool tooLong = ok && std::fread( &extra, 1, 1, f ) == 1;
    std::fclose( f );
    if( !ok || tooLong )
      return fail( VOFX_ERR_ARGUMENT, std::string( path ) + " does not hold a time and one double per owned cell of this context" );
    ctx->hostMirrorOf = nullptr;
    VOFX_CUDA( cudaMemcpyAsync( ctx->dU + first, ctx->pinned, n * sizeof( double ), cudaMemcpyHostToDevice, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    ctx->primedU = nullptr;       // new colour values: the next pass flags densely
    ctx->primedFlags = nullptr;
    return VOFX_OK;
  }

  int vofx_step_resident ( vofx_ctx *ctx, const vofx_step_params *p )
  {
    if( !ctx || !p )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = ensureMirrors( ctx );
    if( !rc && p->with_curvature )
      rc = ensureDevice( ctx->dKappa, (size_t) ctx->grid.cells );
    // the mirror is only written by the loop passes and by vofx_color_upload: flags are maintained incrementally
    if( !rc && !ctx->incremental )
      rc = vofx_set_incremental_flagging( ctx, 1 );
    ctx->hostMirrorOf = nullptr;
    return rc ? rc : vofx_step( ctx, p, ctx->dU, ctx->dFlags, ctx->dHs, ctx->dKappa );
  }

  // ---- distributed step over peer-mapped memory ------------------------------------------------------------------------------

  static int ensureWindow ( vofx_ctx *ctx )
  {
    if( ctx->window )
      return VOFX_OK;
    if( ctx->dU )
      return fail( VOFX_ERR_STATE, "vofx_comm_export must come before the first use of the context-resident colour function" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    const size_t uBytes = (size_t) ctx->grid.cells * sizeof( double );
    ctx->slotsOffset = ( uBytes + 255 ) & ~size_t( 255 );
    ctx->windowBytes = ctx->slotsOffset + ( ( sizeof( CommSlots ) + 255 ) & ~size_t( 255 ) );
    VOFX_CUDA( cudaMalloc( &ctx->window, ctx->windowBytes ) );
    VOFX_CUDA( cudaMemset( ctx->window, 0, ctx->windowBytes ) );
    ctx->dU = static_cast< double * >( ctx->window );
    // the kernels of a distributed pass WAIT for kernels of other contexts: on the legacy default stream, which every
    // context of a process shares (and which synchronises with all blocking streams), they would wait for ever
    if( !ctx->stream )
    {
      VOFX_CUDA( cudaStreamCreateWithFlags( &ctx->ownStream, cudaStreamNonBlocking ) );
      ctx->stream = ctx->ownStream;
    }
    return VOFX_OK;
  }

  static int fillPeers ( vofx_ctx *ctx, int rank, int world, const int32_t *layers, void *const *windows )
  {
    if( world < 1 || world > kMaxRanks || rank < 0 || rank >= world )
      return fail( VOFX_ERR_ARGUMENT, "bad rank / world (at most 16 ranks)" );
    const Grid &g = ctx->grid;
    const int H = ctx->desc.halo;
    const int nMine = g.own1 - g.own0;
    if( layers[ rank ] != nMine )
      return fail( VOFX_ERR_ARGUMENT, "layers[ rank ] differs from the context's owned layers" );
    if( world > 1 && H < 1 )
      return fail( VOFX_ERR_ARGUMENT, "a distributed context needs halo layers" );
    for( int r = 0; r < world; ++r )
      if( world > 1 && layers[ r ] < H )
        return fail( VOFX_ERR_ARGUMENT, "a slab is thinner than the halo" );
    CommPeers &P = ctx->peers;
    std::memset( &P, 0, sizeof( P ) );
    P.rank = rank;
    P.world = world;
    const long long layerCells = g.cells / g.ns[ g.dim - 1 ];
    for( int r = 0; r < world; ++r )
    {
      // every context's window has the same layout only up to the colour function's size: the slots sit at its own offset
      const size_t uBytes = (size_t) layerCells * ( layers[ r ] + 2 * H ) * sizeof( double );
      const size_t off = ( uBytes + 255 ) & ~size_t( 255 );
      P.slots[ r ] = reinterpret_cast< CommSlots * >( static_cast< char * >( windows[ r ] ) + off );
    }
    if( rank > 0 )
    {
      P.uLo = static_cast< double * >( windows[ rank - 1 ] );
      P.offLo = (long long) layers[ rank - 1 ] * layerCells;
    }
    if( rank + 1 < world )
    {
      P.uHi = static_cast< double * >( windows[ rank + 1 ] );
      P.offHi = -(long long) nMine * layerCells;
    }
    P.pushLo1 = g.own0 + H;
    P.pushHi0 = g.own1 - H;
    ctx->commConnected = world > 1;
    // everything a pass needs is allocated now: a pass of the distributed run must never block the host (cudaMalloc /
    // cudaFree may wait for the device, and a kernel of ANOTHER context of this process may be spinning on a signal that
    // only this context's next launches can send)
    int rc = ensureMirrors( ctx );
    rc = rc ? rc : ensureDevice( ctx->dKappa, (size_t) g.cells );
    rc = rc ? rc : vofx_set_incremental_flagging( ctx, 1 );
    if( rc )
      return rc;
    // ... and every kernel is loaded now (CUDA loads kernels lazily at their first launch, which may wait for the device)
    launch::preload_dense();
    launch::preload_band2d();
    launch::preload_recon3d();
    launch::preload_flux3d();
    launch::preload_swartz3d();
    VOFX_CUDA( cudaGetLastError() );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    return VOFX_OK;
  }

  int vofx_comm_export ( vofx_ctx *ctx, void *handle )
  {
    if( !ctx || !handle )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = ensureWindow( ctx );
    if( rc )
      return rc;
    // the handle: cudaIpcMemHandle_t (64 bytes) + process id + the window's address in that process; a rank of the SAME
    // process (several contexts driven by threads of one process) is connected through the plain pointer -- CUDA IPC
    // handles cannot be opened by the process that exported them
    static_assert( sizeof( cudaIpcMemHandle_t ) + 16 == VOFX_COMM_HANDLE_BYTES, "handle size" );
    cudaIpcMemHandle_t h;
    VOFX_CUDA( cudaIpcGetMemHandle( &h, ctx->window ) );
    char *out = static_cast< char * >( handle );
    std::memcpy( out, &h, sizeof( h ) );
    const long long pid = (long long) getpid();
    const unsigned long long address = (unsigned long long) reinterpret_cast< uintptr_t >( ctx->window );
    std::memcpy( out + sizeof( h ), &pid, 8 );
    std::memcpy( out + sizeof( h ) + 8, &address, 8 );
    return VOFX_OK;
  }

  int vofx_comm_connect ( vofx_ctx *ctx, int rank, int world, const int32_t *layers, const void *handles )
  {
    if( !ctx || !layers || !handles )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = ensureWindow( ctx );
    if( rc )
      return rc;
    if( world < 1 || world > kMaxRanks || rank < 0 || rank >= world )
      return fail( VOFX_ERR_ARGUMENT, "bad rank / world (at most 16 ranks)" );
    std::vector< void * > windows( world, nullptr );
    for( int r = 0; r < world; ++r )
    {
      if( r == rank )
      {
        windows[ r ] = ctx->window;
        continue;
      }
      const char *entry = static_cast< const char * >( handles ) + (size_t) r * VOFX_COMM_HANDLE_BYTES;
      cudaIpcMemHandle_t h;
      long long pid = 0;
      unsigned long long address = 0;
      std::memcpy( &h, entry, sizeof( h ) );
      std::memcpy( &pid, entry + sizeof( h ), 8 );
      std::memcpy( &address, entry + sizeof( h ) + 8, 8 );
      if( pid == (long long) getpid() )
      {
        windows[ r ] = reinterpret_cast< void * >( (uintptr_t) address );     // a context of this process
        cudaPointerAttributes attr;
        if( cudaPointerGetAttributes( &attr, windows[ r ] ) == cudaSuccess && attr.device != ctx->device )
        {
          const cudaError_t e = cudaDeviceEnablePeerAccess( attr.device, 0 );
          if( e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled )
            return fail( VOFX_ERR_CUDA, std::string( "cudaDeviceEnablePeerAccess: " ) + cudaGetErrorString( e ) );
        }
        cudaGetLastError();
        continue;
      }
      void *p = nullptr;
      VOFX_CUDA( cudaIpcOpenMemHandle( &p, h, cudaIpcMemLazyEnablePeerAccess ) );
      ctx->ipcMapped.push_back( p );
      windows[ r ] = p;
    }
    return fillPeers( ctx, rank, world, layers, windows.data() );
  }

  int vofx_comm_connect_local ( vofx_ctx **contexts, int world )
  {
    if( !contexts || world < 1 || world > kMaxRanks )
      return fail( VOFX_ERR_ARGUMENT, "bad argument (at most 16 contexts)" );
    std::vector< void * > windows( world, nullptr );
    std::vector< int32_t > layers( world, 0 );
    for( int r = 0; r < world; ++r )
    {
      if( !contexts[ r ] )
        return fail( VOFX_ERR_ARGUMENT, "null context" );
      int rc = ensureWindow( contexts[ r ] );
      if( rc )
        return rc;
      windows[ r ] = contexts[ r ]->window;
      layers[ r ] = contexts[ r ]->grid.own1 - contexts[ r ]->grid.own0;
    }
    for( int r = 0; r < world; ++r )
    {
      // contexts on different devices of one process: plain peer access
      for( int q = 0; q < world; ++q )
        if( contexts[ q ]->device != contexts[ r ]->device )
        {
          cudaSetDevice( contexts[ r ]->device );
          const cudaError_t e = cudaDeviceEnablePeerAccess( contexts[ q ]->device, 0 );
          if( e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled )
            return fail( VOFX_ERR_CUDA, std::string( "cudaDeviceEnablePeerAccess: " ) + cudaGetErrorString( e ) );
          cudaGetLastError();
        }
      int rc = fillPeers( contexts[ r ], r, world, layers.data(), windows.data() );
      if( rc )
        return rc;
    }
    return VOFX_OK;
  }

  int vofx_comm_disconnect ( vofx_ctx *ctx )
  {
    if( !ctx )
      return fail( VOFX_ERR_ARGUMENT, "null context" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    for( void *m : ctx->ipcMapped )
      cudaIpcCloseMemHandle( m );
    ctx->ipcMapped.clear();
    ctx->commConnected = false;
    std::memset( &ctx->peers, 0, sizeof( ctx->peers ) );
    ctx->peers.world = 1;
    return VOFX_OK;
  }

  double *vofx_color_device ( vofx_ctx *ctx )
  {
    if( !ctx || ensureMirrors( ctx ) )
      return nullptr;
    ctx->hostMirrorOf = nullptr;     // the caller may write through the pointer
    return ctx->dU;
  }

  uint8_t *vofx_flags_device ( vofx_ctx *ctx )
  {
    if( !ctx || ensureMirrors( ctx ) )
      return nullptr;
    return ctx->dFlags;
  }

  double *vofx_halfspaces_device ( vofx_ctx *ctx )
  {
    if( !ctx || ensureMirrors( ctx ) )
      return nullptr;
    return ctx->dHs;
  }

  int vofx_halo_push ( vofx_ctx *ctx )
  {
    if( !ctx )
      return fail( VOFX_ERR_ARGUMENT, "null context" );
    if( !ctx->commConnected )
      return VOFX_OK;
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    const Grid &g = ctx->grid;
    const long long layerCells = g.cells / g.ns[ g.dim - 1 ];
    profBegin( ctx );
    launch::halo_push( ctx->numSMs * 4, ctx->stream, ctx->peers, ctx->dU, layerCells, g.own0, g.own1 );
    rc = checkLaunch( ctx, "k_halo_push" );
    if( rc )
      return rc;
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    return VOFX_OK;
  }

  int vofx_step_distributed_phase ( vofx_ctx *ctx, const vofx_step_params *p, int phase )
  {
    if( !ctx || !p || phase < 0 || phase > 1 )
      return fail( VOFX_ERR_ARGUMENT, "bad argument" );
    int rc = ensureMirrors( ctx );
    if( !rc && p->with_curvature )
      rc = ensureDevice( ctx->dKappa, (size_t) ctx->grid.cells );
    if( !rc && !ctx->incremental )
      rc = vofx_set_incremental_flagging( ctx, 1 );
    if( rc )
      return rc;
    ctx->hostMirrorOf = nullptr;
    double *u = ctx->dU;
    unsigned char *flags = ctx->dFlags;
    if( phase == 0 )
    {
      // flags of the layers that depend on owned colour values only, while the neighbours' last pushes are still in flight
      rc = launchStepFlags( ctx, p, u, flags, 0 );
      if( rc )
        return rc;
      if( ctx->commConnected )
      {
        profBegin( ctx );
        launch::comm_wait_halo( ctx->stream, ctx->peers, ctx->state );
        rc = checkLaunch( ctx, "k_comm_wait_halo" );
      }
      rc = rc ? rc : launchStepFlags( ctx, p, u, flags, 1 );
      return rc ? rc : forkSide( ctx, u, flags );
    }
    rc = requireVelocity( ctx );
    rc = rc ? rc : launchBand( ctx, p, u, flags, ctx->dHs, ctx->dKappa );
    if( rc )
      return rc;
    if( ctx->commConnected )
    {
      profBegin( ctx );
      launch::comm_flux_done( ctx->stream, ctx->peers, ctx->state );
      rc = checkLaunch( ctx, "k_comm_flux_done" );
    }
    rc = rc ? rc : launchUpdate( ctx, flags, u, true );
    if( rc )
      return rc;
    notePassIssued( ctx, u, flags );
    rc = joinSide( ctx );
    if( rc )
      return rc;
    profBegin( ctx );
    if( ctx->commConnected )
    {
      launch::finalize_dist( ctx->stream, ctx->peers, ctx->state, p->cfl );
      return checkLaunch( ctx, "k_finalize_dist" );
    }
    launch::finalize( ctx->stream, ctx->state, p->cfl );
    return checkLaunch( ctx, "k_finalize" );
  }

  int vofx_step_distributed ( vofx_ctx *ctx, const vofx_step_params *p )
  {
    int rc = vofx_step_distributed_phase( ctx, p, 0 );
    return rc ? rc : vofx_step_distributed_phase( ctx, p, 1 );
  }

  // ---- interface extraction ------------------------------------------------------------------------------------------------

  int vofx_interface ( vofx_ctx *ctx, const double *halfspaces, const uint8_t *flags, int64_t *n_mixed, int64_t *n_vertices )
  {
    if( !ctx || !halfspaces || !flags || !n_mixed || !n_vertices )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    const Grid &g = ctx->grid;
    const int dim = g.dim;
    const int stride = ( dim == 2 ) ? 2 : 12;
    rc = growDevice( ctx->ifBlockSums, ctx->ifBlockCap, (size_t) launch::scan_blocks( g.cells ) );
    rc = rc ? rc : ensureDevice( ctx->ifTotal, (size_t) 2 );
    if( rc )
      return rc;
    // ordered list of the owned mixed cells (MixedCellMapper numbering)
    launch::scan( dim, ctx->stream, g, g.cells, flags, nullptr, ctx->ifBlockSums, ctx->ifTotal, nullptr, 0 );
    long long total = 0;
    VOFX_CUDA( cudaMemcpyAsync( &total, ctx->ifTotal, sizeof( total ), cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    ctx->launches += 2;
    ctx->ifMixed = total;
    ctx->ifVertices = 0;
    *n_mixed = total;
    *n_vertices = 0;
    if( total == 0 )
      return VOFX_OK;
    if( total > ctx->capacity )
      return fail( VOFX_ERR_STATE, "more mixed cells than the context's capacity" );
    rc = growDevice( ctx->ifCells, ctx->ifCellCap, (size_t) total );
    rc = rc ? rc : growDevice( ctx->ifNverts, ctx->ifNvertsCap, (size_t) total );
    rc = rc ? rc : growDevice( ctx->ifOffsets, ctx->ifOffsetsCap, (size_t) total );
    rc = rc ? rc : growDevice( ctx->ifVerts, ctx->ifVertsCap, (size_t) total * stride * dim );
    if( rc )
      return rc;
    launch::scan( dim, ctx->stream, g, g.cells, flags, nullptr, ctx->ifBlockSums, ctx->ifTotal, ctx->ifCells, 1 );
    const int blocks = gridBlocks( ctx, total, 128, 8 );
    launch::interface_cut( dim, blocks, ctx->stream, g, halfspaces, ctx->ifCells, total, ctx->ifNverts, ctx->ifVerts );
    // CSR offsets = exclusive scan of the vertex counts
    launch::scan( dim, ctx->stream, g, total, nullptr, ctx->ifNverts, ctx->ifBlockSums, ctx->ifTotal + 1, nullptr, 0 );
    long long nv = 0;
    VOFX_CUDA( cudaMemcpyAsync( &nv, ctx->ifTotal + 1, sizeof( nv ), cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    launch::scan( dim, ctx->stream, g, total, nullptr, ctx->ifNverts, ctx->ifBlockSums, ctx->ifTotal + 1, ctx->ifOffsets, 1 );
    rc = growDevice( ctx->ifCsr, ctx->ifCsrCap, (size_t) ( nv > 0 ? nv : 1 ) * dim );
    if( rc )
      return rc;
    launch::interface_gather( dim, blocks, ctx->stream, total, ctx->ifNverts, ctx->ifOffsets, ctx->ifVerts, ctx->ifCsr );
    ctx->launches += 6;
    VOFX_CUDA( cudaGetLastError() );
    ctx->ifVertices = nv;
    *n_vertices = nv;
    return VOFX_OK;
  }

  int vofx_interface_host ( vofx_ctx *ctx, const double *halfspaces, const uint8_t *flags, int64_t *n_mixed, int64_t *n_vertices )
  {
    if( !ctx || !halfspaces || !flags || !n_mixed || !n_vertices )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    const size_t N = (size_t) ctx->grid.cells;
    int rc = ensureMirrors( ctx );
    rc = rc ? rc : stageIn( ctx, ctx->dHs, halfspaces, N * ( ctx->desc.dim + 1 ) * sizeof( double ) );
    rc = rc ? rc : stageIn( ctx, ctx->dFlags, flags, N );
    return rc ? rc : vofx_interface( ctx, ctx->dHs, ctx->dFlags, n_mixed, n_vertices );
  }

  int vofx_interface_fetch ( vofx_ctx *ctx, int32_t *cells, int64_t *offsets, double *vertices )
  {
    if( !ctx || !cells || !offsets || !vertices )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    const size_t n = (size_t) ctx->ifMixed;
    offsets[ 0 ] = 0;
    if( n == 0 )
      return VOFX_OK;
    std::vector< int > off32( n );
    VOFX_CUDA( cudaMemcpyAsync( cells, ctx->ifCells, n * sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaMemcpyAsync( off32.data(), ctx->ifOffsets, n * sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaMemcpyAsync( vertices, ctx->ifCsr, (size_t) ctx->ifVertices * ctx->desc.dim * sizeof( double ), cudaMemcpyDeviceToHost, ctx->stream ) );
    VOFX_CUDA( cudaStreamSynchronize( ctx->stream ) );
    for( size_t i = 0; i < n; ++i )
      offsets[ i ] = off32[ i ];
    offsets[ n ] = ctx->ifVertices;
    return VOFX_OK;
  }

  // ---- initial data -------------------------------------------------------------------------------------------------------

  int vofx_init ( vofx_ctx *ctx, int problem, double time, double *u )
  {
    if( !ctx || !u )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = setDevice( ctx );
    if( rc )
      return rc;
    const int dim = ctx->desc.dim;
    Indicator I;
    std::memset( &I, 0, sizeof( I ) );
    I.dim = dim;
    switch( problem )
    {
    case VOFX_PROB_ROTATING_CIRCLE:
      if( dim != 3 )
        return fail( VOFX_ERR_ARGUMENT, "the 2D RotatingCircle is initialised by the exact circle average on the host (test/average.hh:82-101)" );
      {
        // rotatingcircle.hh:81-90
        double c[ 3 ] = { 0.5, 0.5, 0.5 };
        const double a[ 3 ] = { 0.0, 0.25, 0.0 }, b[ 3 ] = { 0.25, 0.0, 0.0 }, z[ 3 ] = { 0.0, 0.0, 0.0 };
        const double cs = std::cos( ( 2 * M_PI / 10 ) * time ), sn = std::sin( ( 2 * M_PI / 10 ) * time ), msn = -std::sin( ( 2 * M_PI / 10 ) * time );
        for( int k = 0; k < 3; ++k )
        {
          c[ k ] += cs * a[ k ];
          c[ k ] += sn * b[ k ];
          c[ k ] += msn * z[ k ];
          I.c[ k ] = c[ k ];
        }
        I.kind = 0;
        I.r = 0.15;
      }
      break;
    case VOFX_PROB_SFLOW:
      I.kind = 1;
      I.c[ 0 ] = I.c[ 1 ] = I.c[ 2 ] = 0.5;
      I.r = 0.25;
      break;
    case VOFX_PROB_SLOTTED_CYLINDER:
      if( dim != 2 )
        return fail( VOFX_ERR_ARGUMENT, "SlottedCylinder is a 2D problem" );
      I.kind = 2;
      I.cs = std::cos( ( 2.0 * M_PI / 10.0 ) * time );
      I.sn = std::sin( ( 2.0 * M_PI / 10.0 ) * time );
      break;
    case VOFX_PROB_LINEAR_WALL:
      I.kind = 3;
      I.c[ 0 ] = 0.3 + time * 0.02;
      break;
    case VOFX_PROB_LEVEQUE:
      I.kind = 0;
      I.r = 0.15;
      if( dim == 2 )
      {
        I.c[ 0 ] = 0.5;
        I.c[ 1 ] = 0.75;
      }
      else
        I.c[ 0 ] = I.c[ 1 ] = I.c[ 2 ] = 0.35;
      break;
    default:
      return fail( VOFX_ERR_ARGUMENT, "unknown built-in problem" );
    }
    const int blocks = gridBlocks( ctx, ctx->grid.cells, 128, 16 );
    if( u == ctx->primedU )
      ctx->primedU = nullptr, ctx->primedFlags = nullptr;
    profBegin( ctx );
    launch::init( dim, blocks, ctx->stream, ctx->grid, I, u );
    return checkLaunch( ctx, "k_init" );
  }

  // ---- primitives (tests) ----------------------------------------------------------------------------------------------------

  namespace
  {
    struct DevBuf
    {
      void *p = nullptr;
      ~DevBuf () { cudaFree( p ); }
      int alloc ( size_t bytes )
      {
        VOFX_CUDA( cudaMalloc( &p, bytes ? bytes : 1 ) );
        return VOFX_OK;
      }
      int upload ( const void *src, size_t bytes )
      {
        int rc = alloc( bytes );
        if( rc )
          return rc;
        VOFX_CUDA( cudaMemcpy( p, src, bytes, cudaMemcpyHostToDevice ) );
        return VOFX_OK;
      }
    };

    int requireDevice ()
    {
      int count = 0;
      if( cudaGetDeviceCount( &count ) != cudaSuccess || count == 0 )
        return fail( VOFX_ERR_CUDA, "no CUDA device: vofx has no CPU fallback" );
      return VOFX_OK;
    }
  } // namespace

  // measured fp64 rates of this GPU in G operations per second (an FMA counts as ONE operation here):
  // rates[0] fused multiply-add, rates[1] separate multiply + add (two operations per pair)
  int vofx_test_fp64_rate ( double *rates )
  {
    if( !rates )
      return fail( VOFX_ERR_ARGUMENT, "null argument" );
    int rc = requireDevice();
    if( rc )
      return rc;
    cudaDeviceProp prop;
    VOFX_CUDA( cudaGetDeviceProperties( &prop, 0 ) );
    DevBuf scratch;
    rc = scratch.alloc( 64 );
    if( rc )
      return rc;
    const int blocks = prop.multiProcessorCount * 8, iters = 4096, reps = 10;
    for( int mode = 0; mode < 2; ++mode )
    {
      const float ms = launch::fp64_rate( mode, blocks, iters, reps, (double *) scratch.p );
      const double ops = (double) blocks * 256.0 * iters * 8.0 * reps * ( mode == 0 ? 1.0 : 2.0 );
      rates[ mode ] = ops / ( ms * 1e-3 ) / 1e9;
    }
    VOFX_CUDA( cudaGetLastError() );
    return VOFX_OK;
  }

  int vofx_test_arith ( int64_t count, uint64_t seed, int64_t *mismatches )
  {
    if( count < 0 || !mismatches )
      return fail( VOFX_ERR_ARGUMENT, "bad argument" );
    int rc = requireDevice();
    if( rc )
      return rc;
    DevBuf d;
    rc = d.alloc( 3 * sizeof( unsigned long long ) );
    if( rc )
      return rc;
    VOFX_CUDA( cudaMemset( d.p, 0, 3 * sizeof( unsigned long long ) ) );
    launch::test_arith( count, seed, (unsigned long long *) d.p );
    VOFX_CUDA( cudaGetLastError() );
    unsigned long long out[ 3 ];
    VOFX_CUDA( cudaMemcpy( out, d.p, sizeof( out ), cudaMemcpyDeviceToHost ) );
    for( int k = 0; k < 3; ++k )
      mismatches[ k ] = (int64_t) out[ k ];
    return VOFX_OK;
  }

  int vofx_test_locate ( int dim, int64_t count, const double *lo, const double *h, const double *normal, const double *fraction, double *halfspaces )
  {
    if( ( dim != 2 && dim != 3 ) || count < 0 || !lo || !h || !normal || !fraction || !halfspaces )
      return fail( VOFX_ERR_ARGUMENT, "bad argument" );
    int rc = requireDevice();
    if( rc )
      return rc;
    DevBuf dLo, dH, dN, dF, dOut;
    const size_t n = (size_t) count;
    rc = dLo.upload( lo, n * dim * 8 );
    rc = rc ? rc : dH.upload( h, n * dim * 8 );
    rc = rc ? rc : dN.upload( normal, n * dim * 8 );
    rc = rc ? rc : dF.upload( fraction, n * 8 );
    rc = rc ? rc : dOut.alloc( n * ( dim + 1 ) * 8 );
    if( rc )
      return rc;
    const int blocks = (int) ( ( n + 63 ) / 64 ) > 0 ? (int) ( ( n + 63 ) / 64 ) : 1;
    if( dim == 2 )
      launch::test_locate2( blocks, count, (double *) dLo.p, (double *) dH.p, (double *) dN.p, (double *) dF.p, (double *) dOut.p );
    else
      launch::test_locate3( blocks, count, (double *) dLo.p, (double *) dH.p, (double *) dN.p, (double *) dF.p, (double *) dOut.p );
    VOFX_CUDA( cudaGetLastError() );
    VOFX_CUDA( cudaMemcpy( halfspaces, dOut.p, n * ( dim + 1 ) * 8, cudaMemcpyDeviceToHost ) );
    return VOFX_OK;
  }

  int vofx_test_flux ( int dim, int64_t count, const double *lo, const double *h, const int32_t *face, const double *v, const double *halfspaces, double *flux )
  {
    if( ( dim != 2 && dim != 3 ) || count < 0 || !lo || !h || !face || !v || !halfspaces || !flux )
      return fail( VOFX_ERR_ARGUMENT, "bad argument" );
    int rc = requireDevice();
    if( rc )
      return rc;
    DevBuf dLo, dH, dFace, dV, dHs, dOut;
    const size_t n = (size_t) count;
    rc = dLo.upload( lo, n * dim * 8 );
    rc = rc ? rc : dH.upload( h, n * dim * 8 );
    rc = rc ? rc : dFace.upload( face, n * 4 );
    rc = rc ? rc : dV.upload( v, n * dim * 8 );
    rc = rc ? rc : dHs.upload( halfspaces, n * ( dim + 1 ) * 8 );
    rc = rc ? rc : dOut.alloc( n * 8 );
    if( rc )
      return rc;
    const int blocks = (int) ( ( n + 63 ) / 64 ) > 0 ? (int) ( ( n + 63 ) / 64 ) : 1;
    if( dim == 2 )
      launch::test_flux2( blocks, count, (double *) dLo.p, (double *) dH.p, (int *) dFace.p, (double *) dV.p, (double *) dHs.p, (double *) dOut.p );
    else
      launch::test_flux3( blocks, count, (double *) dLo.p, (double *) dH.p, (int *) dFace.p, (double *) dV.p, (double *) dHs.p, (double *) dOut.p );
    VOFX_CUDA( cudaGetLastError() );
    VOFX_CUDA( cudaMemcpy( flux, dOut.p, n * 8, cudaMemcpyDeviceToHost ) );
    return VOFX_OK;
  }

  int vofx_test_interface ( int dim, int64_t count, const double *lo, const double *h, const double *halfspaces, int32_t *nverts, double *centroid, double *measure )
  {
    if( ( dim != 2 && dim != 3 ) || count < 0 || !lo || !h || !halfspaces || !nverts || !centroid || !measure )
      return fail( VOFX_ERR_ARGUMENT, "bad argument" );
    int rc = requireDevice();
    if( rc )
      return rc;
    DevBuf dLo, dH, dHs, dNv, dCen, dMeas;
    const size_t n = (size_t) count;
    rc = dLo.upload( lo, n * dim * 8 );
    rc = rc ? rc : dH.upload( h, n * dim * 8 );
    rc = rc ? rc : dHs.upload( halfspaces, n * ( dim + 1 ) * 8 );
    rc = rc ? rc : dNv.alloc( n * 4 );
    rc = rc ? rc : dCen.alloc( n * dim * 8 );
    rc = rc ? rc : dMeas.alloc( n * 8 );
    if( rc )
      return rc;
    const int blocks = (int) ( ( n + 63 ) / 64 ) > 0 ? (int) ( ( n + 63 ) / 64 ) : 1;
    if( dim == 2 )
      launch::test_interface2( blocks, count, (double *) dLo.p, (double *) dH.p, (double *) dHs.p, (int *) dNv.p, (double *) dCen.p, (double *) dMeas.p );
    else
      launch::test_interface3( blocks, count, (double *) dLo.p, (double *) dH.p, (double *) dHs.p, (int *) dNv.p, (double *) dCen.p, (double *) dMeas.p );
    VOFX_CUDA( cudaGetLastError() );
    VOFX_CUDA( cudaMemcpy( nverts, dNv.p, n * 4, cudaMemcpyDeviceToHost ) );
    VOFX_CUDA( cudaMemcpy( centroid, dCen.p, n * dim * 8, cudaMemcpyDeviceToHost ) );
    VOFX_CUDA( cudaMemcpy( measure, dMeas.p, n * 8, cudaMemcpyDeviceToHost ) );
    return VOFX_OK;
  }

} // extern "C"
