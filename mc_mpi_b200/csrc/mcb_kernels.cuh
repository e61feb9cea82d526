// mcb_kernels.cuh -- device-side data structures + launch wrappers of the
// tracking path (implemented in mcb_kernels.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "mcb_math.cuh"

namespace mcb {

// ---- the tally: an exact long accumulator ---------------------------------
// weights_absorbed[cell] (include/layer/layer.hpp:92) is kept per cell as a
// 128-bit two's-complement fixed-point number, least significant bit 2^-120,
// in four 32-bit digits.  A float deposit is its 24-bit significand shifted
// to its exponent: an INTEGER add.  Integer adds are associative, so the
// tally is the exact sum of the per-event floats (floats below 2^-97 lose the
// bits under 2^-120) whatever the order, the CTA shape, the number of
// launches or the number of GPUs -- unlike the reference's float tally, which
// moves by 3e-4 per cell with the OpenMP thread count (SURVEY hard part 1).
// The digits are 32-bit because sm_100a has native shared-memory atomics only
// for 32-bit integers (64-bit and float shared atomics are CAS loops).
constexpr int kAccDigits = 4;
constexpr int kAccLsbLog2 = -120;
// three extra "cells" after the m real ones collect the weight carried by
// histories classified left / right / dead
constexpr int kAccExtra = 3;

// Escapees leave the tracking kernel through per-CTA "stripes": CTA b appends to its own
// region [b*stripe_cap, (b+1)*stripe_cap) of the launch's scratch outbox with a SHARED-memory
// counter (no global atomic on the path; a stripe that fills up spills into one common
// overflow segment).  gather_stripes then packs the stripes into the layer's contiguous
// outbox of 24-byte wire records.  kStripes bounds the grid of the tracking kernel.
constexpr int kStripes = 2048;
// bank slots are handed to warps in chunks of kWorkChunk (one global atomic per chunk)
constexpr int kWorkChunk = 64;

// Device counters of one layer; zeroed per tracking launch, read back after.
struct DevCounters {
  unsigned long long cursor;      // next unclaimed slot of the launch's bank range
  unsigned long long out_total[2];  // escapees of this launch per side (written by gather_stripes)
  unsigned long long n_cls[3];    // histories classified left / right / dead
  unsigned long long events;
  unsigned long long scatters;
  unsigned int overflow;          // an outbox was too small (host sizes them so it cannot be)
  unsigned int acc_range;         // a deposit did not fit the accumulator (weight >= 2^8)
};

// per-cell constants of the event, precomputed once per layer from the public
// sigs / absorption_rates vectors exactly as src/layer.cpp:131-133 does
//   .x = sig_a = sigs*a          .y = sig_i = sigs*(float)(1.0 - a)
//   .z = a float just BELOW 1/sig_i (by 2^-20; +inf when sig_i <= EPS), used only by the
//        "certain crossing" test of the event (see track_kernel); .w unused
typedef float4 CellXs;

struct TrackParams {
  // particle bank, structure of arrays of vectors: seed[] (8 B) and
  // st[] = {x, mu, wmc, bits(index)} (16 B); every access is one coalesced
  // 64-bit / 128-bit transaction per lane
  const unsigned long long *bank_seed;
  const float4 *bank_st;
  long long take_base;   // first bank slot of this launch
  long long take_count;  // number of particles to track
  // the sub-slab
  const CellXs *xs;      // m entries (global memory; staged to smem when it fits)
  int idx_lo;            // Layer::index_start
  int m;
  float dx;
  float minw;            // particle_min_weight
  // outputs
  unsigned *acc;         // the global tally: u64[2][m + kAccExtra] (low halves, then high
                         // halves of the 128-bit accumulators), see gacc_add
  // striped outbox per side (24-byte records, 3 words each).  The memory is either this
  // layer's own scratch (packed afterwards by gather_stripes) or -- direct peer exchange --
  // the NEIGHBOUR GPU's inbox, mapped over NVLink: then the escapee stores of the tracking
  // kernel ARE the communication.
  unsigned long long *out_rec[2];
  unsigned *fills[2];    // per side: fill of stripe b at [b]; overflow segment's at [ovf_slot]
  int ovf_slot[2];       // index of the overflow counter in fills[side]
  int stripe_cap[2];     // records per stripe
  long long ovf_base[2]; // first record of the overflow segment (= nstripes * stripe_cap)
  long long ovf_cap[2];  // its capacity
  int write_side[2];     // 0: global border, escapees are only counted
  int retire_batch;      // retire / refill once this many lanes of a warp are without a live
                         // history (>= 1): amortises the bookkeeping over short segments
  unsigned rng_key;      // Philox key (rng mode 1)
  DevCounters *ctr;
};

enum TallyMode { kTallyShared = 1, kTallyGlobal = 2 };

struct TrackLaunch {
  int rng;          // 0 = LCG (parity mode), 1 = Philox2x32-10
  int tally_mode;   // TallyMode
  int block;        // threads per CTA
  int grid;         // CTAs
  size_t smem;      // dynamic shared memory bytes
};

size_t track_smem_bytes(int tally_mode, int m);
cudaError_t track_configure(int device, int m, int want_mode, int want_block,
                            int want_blocks_per_sm, int rng, TrackLaunch *out);
// grid actually used for `take` particles (<= cfg.grid <= kStripes)
int track_grid(const TrackLaunch &cfg, long long take);
cudaError_t launch_track(const TrackParams &p, const TrackLaunch &cfg,
                         cudaStream_t stream);
// pack the stripes (+ overflow segment) of one side behind `settled_n` records of the
// layer's contiguous outbox; adds the number of records to *out_total
cudaError_t launch_gather_stripes(const unsigned long long *scratch, const unsigned *stripe_n,
                                  int nstripes, int stripe_cap, long long ovf_base, long long ovf_cap,
                                  unsigned long long *settled, long long settled_n,
                                  unsigned long long *out_total, cudaStream_t stream);

// same packing, but from an INBOX (stripes filled by a neighbour's tracking kernel) straight
// into the bank layout behind `bank_n` particles; *out_total = number of particles added
cudaError_t launch_gather_stripes_to_bank(const unsigned long long *scratch,
                                          const unsigned *fills, int nstripes, int stripe_cap,
                                          long long ovf_base, long long ovf_cap,
                                          unsigned long long *bank_seed, float4 *bank_st,
                                          long long bank_n, unsigned long long *out_total,
                                          cudaStream_t stream);

// device-side birth (src/layer.cpp:101-120): particle i of the batch gets
// seed = rnd_seed^(i+1)(chain_state), mu from the first rnd_real draw
cudaError_t launch_birth(long long n, unsigned long long chain_state,
                         const JumpTable &seed_jump, float x_ini, float wmc,
                         int index, unsigned long long *seed_out, float4 *st_out,
                         cudaStream_t stream);
// Philox mode: history first_id + i starts at counter (first_id + i) << kPhiloxEventBits; its
// event 0 is the birth (mu from word 0), so it is banked with event number 1
cudaError_t launch_birth_philox(long long n, unsigned long long first_id, unsigned key,
                                float x_ini, float wmc, int index,
                                unsigned long long *seed_out, float4 *st_out,
                                cudaStream_t stream);
// known-answer-test: out[i] = philox2x32_10(c0[i], c1[i], key[i])
cudaError_t launch_test_philox(long long n, const unsigned *c0, const unsigned *c1,
                               const unsigned *key, uint2 *out, cudaStream_t stream);

// 24-byte AoS records (include/types/particle.hpp) <-> bank layout
cudaError_t launch_aos_to_soa(long long n, const void *aos,
                              unsigned long long *seed, float4 *st,
                              cudaStream_t stream);
cudaError_t launch_soa_to_aos(long long n, const unsigned long long *seed,
                              const float4 *st, void *aos, cudaStream_t stream);

// known-answer-test kernels
cudaError_t launch_test_rnd_real(long long n, unsigned long long *seeds,
                                 float *out, cudaStream_t stream);
cudaError_t launch_test_math(int which, long long n, const float *in, float *out,
                             cudaStream_t stream);
cudaError_t launch_test_div(long long n, const float *a, const float *b, float *out,
                            cudaStream_t stream);
// exact accumulation of n floats into ONE accumulator (kAccDigits words)
cudaError_t launch_test_accumulate(long long n, const float *in, unsigned *acc4,
                                   cudaStream_t stream);

}  // namespace mcb
