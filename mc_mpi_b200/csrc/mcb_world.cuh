// mcb_world.cuh -- device-side structures of the PERSISTENT exchange ("world") kernel.
//
// The reference's workers (src/worker_sync.cpp:24-135, src/worker_async.cpp:19-105,
// src/worker_rma.cpp:15-68) alternate  simulate -> exchange -> termination vote  on the
// host.  Here one resident kernel per GPU does all three for the whole run:
//
//   * the rank's sub-slab is cut into V WINDOWS; `cpw` CTAs of the launch serve each window
//     (V = 1 when the CTA-private tally of the whole sub-slab fits shared memory);
//   * every directed link window -> neighbouring window is a set of single-producer /
//     single-consumer RINGS, one per warp ("stripe"): warp w of the producing window stores
//     its escapees as 24-byte wire records into stripe w of the consuming window's memory;
//     warp w of the consuming window polls the next slots of that stripe in its OWN memory,
//     takes the records that have arrived into idle lanes and returns credits.  Every 8-byte
//     word of a slot carries the slot's lap parity in a bit the record never uses, so a word
//     validates itself: no fence, no published count, no atomics on the path, no host.
//     Between the last window of rank r and the first window of rank r+1 the consumer's
//     memory is peer-mapped over NVLink: the escapee stores ARE the communication, exactly
//     RmaComm's MPI_Put into the neighbour's window (src/rma_comm.cpp:133-186) with the
//     occupancy word replaced by per-slot parity + a credit counter;
//   * a ring that is full blocks the sending lanes (back-pressure); a warp with blocked lanes
//     drains its own inbound stripes into its CTA's BANK (a multi-producer / multi-consumer
//     queue in local memory, cursors in shared memory) so that two neighbours can never wait
//     on each other; idle lanes refill from rings, then bank, then -- in the source window -- by
//     giving birth to source particles in place (rnd_seed chain by LCG jump-ahead,
//     src/layer.cpp:101-120);
//   * termination (StateComm, src/state_comm.cpp:35-65; MPI_Allreduce,
//     src/worker_sync.cpp:112-120): every CTA adds the histories it disabled to ONE 64-bit
//     device-side counter on the home rank (system-scope red over NVLink); the home rank's
//     idle warps compare it with nb_particles and raise every rank's `done` flag;
//   * flow control: a warp only ever feeds "its" stripe of the neighbouring windows, so stripe s
//     of every window of every rank forms a CHAIN.  Every warp also reports the histories it
//     disables to a per-chain counter on the home rank, and a source warp gives birth only
//     while its own chain holds less than its share of the histories in flight: no chain can
//     hoard the global budget in the queue of a slow warp while the others run dry.
#pragma once
#include "mcb_kernels.cuh"

namespace mcb {

constexpr int kWorldMaxRanks = 64;
constexpr int kWorldMaxWarps = 32;   // warps per CTA
constexpr int kWorldMaxWindows = 640; // windows per rank (their bounds travel as kernel parameters)

// what a window PRODUCES into (side 0 = towards lower cells, 1 = towards higher cells)
struct LinkOut {
  unsigned long long *rec;   // consumer's ring memory [stripes][cap][3 words]
  const unsigned *credit;    // LOCAL: records the consumer has taken [stripes] (it stores them)
  int mode;                  // 0 = global border: absorb (src/layer.cpp:350-360); 1 = ring
  int outer;                 // 1 = the link leaves the rank (statistics only)
};
// what a window CONSUMES from (side 0 = from the lower neighbour, 1 = from the higher one)
struct LinkIn {
  const unsigned long long *rec;  // LOCAL ring memory [stripes][cap][3 words]
  unsigned *credit;               // producer's credit array [stripes]
  int present;
  int pad;
};
// the window's banks, one per CTA serving it: records waiting for a free lane (overflow of
// the rings).  Pushed and popped by the warps of that CTA only, so the cursors live in its
// shared memory during a run; `ht` keeps them between runs.
struct BankQ {
  unsigned long long *rec;   // [cpw][cap][3 words], same self-validating slots as the rings
  unsigned *ht;              // [cpw][2]: head (next to pop), tail (next to push)
  unsigned cap;              // records per CTA, power of two
  unsigned log2cap;
};
// (the cells of window v are [win_lo[v], win_lo[v + 1]) of WorldParams: kernel parameters sit
// in the constant bank, so everything derived from them is warp-uniform for the compiler)
struct WindowDesc {
  LinkOut out[2];
  LinkIn in[2];
  BankQ bank;
};

// per rank, at offset 0 of the rank's exported exchange block (peers map it)
struct WorldCtrl {
  unsigned long long disabled_global;  // HOME rank: histories disabled anywhere in the world
  unsigned long long born;             // HOME rank: source histories handed out so far
  unsigned done;                       // raised by the home rank's kernel: the run is over
  unsigned error;                      // -MCB200_ERR_* raised by this rank's kernel
  unsigned pad[58];
};
static_assert(sizeof(WorldCtrl) == 256, "WorldCtrl layout");

struct WorldCounters {
  unsigned long long events, scatters;
  unsigned long long n_cls[3];       // absorbed at the global left / right border, dead
  unsigned long long sent[2];        // records pushed into rings, per side (all windows)
  unsigned long long sent_outer[2];  // of which across the rank boundary (NVLink)
  unsigned long long births;
  unsigned long long idle_polls, blocked_passes, bank_pushes, bank_pops;
  unsigned long long lane_slots;     // 32 x event iterations of all warps: events / lane_slots = utilisation
  unsigned long long idle_ns;        // summed over warps: time without a single live history
  unsigned acc_range;
  unsigned bank_full;                // a bank had no room for a blocked warp's records (they stayed in the ring)
};

struct WorldParams {
  const WindowDesc *win;
  int V, cpw;                        // windows of this launch, CTAs per window
  int rank_lo;                       // first global cell of the rank
  int pad0;
  const CellXs *xs;                  // the rank's cell constants (cell c at xs[c - rank_lo])
  float dx, minw;
  int retire_batch;
  unsigned ring_cap, ring_log2;      // records per stripe, power of two >= 32
  // the source (src_window < 0 on ranks that do not hold x_ini)
  int src_window;
  int src_index;                     // (int)(x_ini / dx), src/layer.cpp:106
  unsigned long long src_total;
  unsigned long long chain_state;    // Layer::seed before the first birth (src/layer.cpp:36)
  int rng;                           // 0 = LCG (parity mode), 1 = Philox2x32-10
  unsigned rng_key;                  // Philox key of the run
  float x_ini, wmc;
  unsigned long long inflight_limit; // births stop while born - disabled exceeds this
  // termination
  unsigned long long total;          // nb_particles of the run
  WorldCtrl *ctrl;                   // this rank's
  unsigned long long *home_disabled; // &home->disabled_global (peer-mapped unless home)
  unsigned *home_chain_disabled;     // home rank's per-stripe counts of disabled histories [stripes]
  unsigned chain_limit;              // births of a source warp pause above this many live histories
                                     // of ITS chain (stripe s of every window of every rank)
  unsigned pad1;
  unsigned *const *done_ptrs;        // home rank: every rank's &ctrl->done [n_ranks]
  int n_ranks, is_home;
  unsigned long long max_run_ns;     // idle warps end the run (error) after this long; 0 = never
  // results
  unsigned long long *acc;           // the rank's tally u64[2][ncell_rank] (gacc layout)
  int ncell_rank;                    // cells of the rank + kAccExtra
  WorldCounters *ctr;
  int win_lo[kWorldMaxWindows + 1];  // window v = global cells [win_lo[v], win_lo[v + 1])
};

struct WorldLaunch {
  int block, grid;
  size_t smem;
  int xs_smem;   // cell constants in shared memory (1) or read from global memory / L2 (0)
};

size_t world_smem_bytes(int m_max, int block, bool xs_smem);
cudaError_t world_configure(int device, int m_max, int block, bool xs_smem, WorldLaunch *out,
                            int *max_ctas_per_sm);
cudaError_t world_upload_jump_table(const JumpTable &jt);   // current device
cudaError_t launch_world(const WorldParams &p, const WorldLaunch &cfg, cudaStream_t stream);

}  // namespace mcb
