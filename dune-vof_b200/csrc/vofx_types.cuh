// vofx: plain-data types shared by the kernels and the host side of the C ABI.
#pragma once

#include "vofx_grid.cuh"

namespace vofx
{

  // device-resident loop state (Algorithm::operator() locals, algorithm.hh:61)
  struct StepState
  {
    double time;
    double dt;
    double dtEst;               // running MIN of the current step (bits compared as unsigned: values are >= 0)
    double dtEstLast;           // estimate of the last finished step
    int mixedCount;             // entries in the mixed list
    int mixedCountLast;
    int overflow;               // mixed list capacity exceeded
    int taskCount;              // entries in the clip-task list of the current step (k_flux_prepare -> k_flux_clip)
    int serialCount;            // clip tasks left to the serial kernel (stored from the end of the task array)
    int serialLocate;           // mixed cells whose plane-constant solve is left to the serial kernel
    int rootCount;              // mixed cells whose bracket is found and whose root search is left to k_locate_root3
    int changedCount;           // entries of the changed-cell list of the current step (host-buffer variant: sparse write-back)
    int changedCountLast;
    int unused0, unused1;
    int overflowLast;           // the last finished pass overflowed the mixed list: it was a no-op (time and u unchanged)
    int segCount;               // incremental flagging: marked row segments of this pass (k_scan_marks -> k_flag_segments)
    int epoch;                  // number of the pass in flight (1, 2, ...): the stamp of the peer-memory signals of the distributed step
    int commError;              // a wait on a peer's signal timed out
    int pad;
    int serialLocateMore[ 3 ];  // serialLocate of the band parts 1 .. kMaxBandParts - 1 (BandPart below; part 0 counts in serialLocate)
    int pad2;
    int costHist[ 32 ];         // k_cost_hist: band cells per ( band part, cost class of the previous step: 0 .. 6 = bracket the solve ended in, 7 = walked )
    int costFill[ 32 ];         // k_cost_order: entries of each bucket placed so far
  };

  // The band kernels of a pass (plane-constant solve, root search, fluxes) can run as `parts` independent chains on as many
  // streams: chain p takes the positions p, p + parts, p + 2 parts, ... of the processing order.  A cell's fluxes need its own
  // reconstruction only (owner-computes gather), so the chains meet again only before the update; the
  // tail of one chain's kernel (latency-bound, a cell is ~70 k cycles of dependent work) is filled by the other chains.
  constexpr int kMaxBandParts = 4;
  struct BandPart
  {
    int part = 0, parts = 1;
  };
  static_assert( kMaxBandParts == 4, "StepState::serialLocateMore holds kMaxBandParts - 1 counters" );

  // ---- distributed step over peer-mapped memory (vofx_comm_*, include/vofx.h) ----------------------------------------
  constexpr int kMaxRanks = 16;

  // written by the OTHER ranks (and by the rank itself for its own entry) with plain stores over NVLink
  struct CommSlots
  {
    int fluxDone[ kMaxRanks ];          // [ r ]: last pass whose fluxes rank r has finished (it no longer reads its colour halos)
    int updateDone[ kMaxRanks ];        // [ r ]: last pass whose update rank r has finished (its pushes into my halos have landed)
    double dtEst[ 2 ][ kMaxRanks ];     // [ pass & 1 ][ r ]: rank r's local time-step estimate of that pass
  };

  struct CommPeers
  {
    int rank, world;
    CommSlots *slots[ kMaxRanks ];      // every rank's slots as mapped into this process
    // my owned layers [ own0, pushLo1 ) are the upper halo of the lower neighbour: its cell = my storage index + offLo;
    // my owned layers [ pushHi0, own1 ) are the lower halo of the upper neighbour
    double *uLo, *uHi;
    long long offLo, offHi;
    int pushLo1, pushHi0;
  };

  // per mixed cell: what it adds to its own update and to each face neighbour's update (Appendix C of SURVEY.md)
  struct FluxRecord
  {
    double self[ 6 ];
    double nb[ 6 ];
    unsigned char selfMask, nbMask;
    unsigned char pad[ 6 ];
  };

  struct Indicator
  {
    int kind;          // 0 ball |x-c| < r, 1 ball |x-c|^2 <= r^2, 2 slotted cylinder, 3 half space x0 <= pos
    int dim;
    double c[ 3 ];     // centre (kinds 0,1) / wall position in c[0] (kind 3)
    double r;
    double cs, sn;     // slotted cylinder: cos/sin of the rotation angle
  };

} // namespace vofx
