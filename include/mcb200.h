/*
 * mcb200.h -- C ABI of the B200-native particle-tracking path (libmcb200.so).
 *
 * Drop-in boundary for mc-mpi's `Layer` hot path.  Plain C: pointers, sizes,
 * PODs; no C++ / torch types.  Every entry point names the reference
 * interface it stands in for (file:line in lkskstlr/mc-mpi).  All functions
 * return MCB200_OK (0) or a negative error code and never call exit();
 * mcb200_last_error() gives the message (thread-local).  The reference's own
 * convention (print + exit, include/gpu_errcheck/gpu_errcheck.hpp:10-18,
 * src/layer.cpp:276-277) is re-created one level up, in the `cusimulate`
 * shim and the C++ `Layer` facade (include/mcb200/layer.hpp).
 *
 * There is NO CPU fallback: without a CUDA device every compute entry point
 * fails with MCB200_ERR_CUDA.
 */
#ifndef MCB200_H
#define MCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCB200_ABI_VERSION 1

enum {
  MCB200_OK = 0,
  MCB200_ERR_INVALID = -1,  /* bad argument / descriptor */
  MCB200_ERR_CUDA = -2,     /* CUDA runtime error (message has the string) */
  MCB200_ERR_NOMEM = -3,    /* host or device allocation failed */
  MCB200_ERR_CAPACITY = -4, /* caller buffer too small */
  MCB200_ERR_RANGE = -5,    /* a weight / deposit outside (-2^7, 2^7): not a Monte-Carlo weight */
  MCB200_ERR_TIMEOUT = -6   /* a multi-GPU run made no progress (a peer died?) and was stopped */
};

/* include/types/particle.hpp:7-18 -- the 24-byte POD the workers ship over
 * MPI (offsets 0/8/12/16/20).  Same layout, so a `Particle*` can be passed. */
typedef struct mcb200_particle {
  uint64_t seed; /* seed_t: state of the particle's own LCG stream */
  float x;       /* absolute position */
  float mu;      /* direction cosine */
  float wmc;     /* Monte-Carlo weight */
  int32_t index; /* GLOBAL cell index */
} mcb200_particle;

/* Construction parameters of one layer (= one contiguous sub-slab on one
 * GPU).  Mirrors Layer::Layer, src/layer.cpp:44-69, plus what the reference
 * hard-codes (cross-sections) or cannot express (device, tally resolution). */
typedef struct mcb200_layer_desc {
  int32_t abi_version;        /* MCB200_ABI_VERSION */
  int32_t device;             /* CUDA device ordinal */
  float x_min, x_max;         /* sub-slab bounds (layer.hpp:85) */
  int32_t index_start;        /* first global cell of the sub-slab (layer.hpp:89) */
  int32_t m;                  /* number of cells (layer.hpp:88) */
  float dx;                   /* cell width used for edges; <= 0 -> (x_max-x_min)/m
                                 like layer.cpp:47.  Multi-GPU runs pass the ONE
                                 global dx so that N GPUs == 1 GPU bit for bit. */
  float particle_min_weight;  /* layer.hpp:105 */
  int32_t left_border;        /* 1/0, or -1 -> |x_min| < 1e-4      (layer.cpp:47) */
  int32_t right_border;       /* 1/0, or -1 -> |x_max - 1| < 1e-4  (layer.cpp:48) */
  const float *sigs;              /* m floats, or NULL -> expf(-x_mid) (layer.cpp:53-60) */
  const float *absorption_rates;  /* m floats, or NULL -> 0.5       (layer.cpp:63) */
  int32_t keep_border;        /* 1: particles absorbed at a global border are
                                 ALSO written to the outboxes (tests); 0: only
                                 counted, like layer.cpp:350-360 */
} mcb200_layer_desc;

typedef struct mcb200_layer mcb200_layer;

/* Cumulative counters of a layer (all 64-bit: SURVEY hard part 7). */
typedef struct mcb200_counts {
  int64_t nb_disabled;  /* Layer::nb_disabled (layer.hpp:97): dead + absorbed at borders */
  int64_t nb_active;    /* Layer::nb_active() (layer.cpp:84-87): bank + unborn */
  int64_t n_bank;       /* particles resident in the device bank */
  int64_t n_unborn;     /* Layer::nb_particles_create (layer.hpp:109) */
  int64_t n_outbox_left;   /* = particles_left.size()  (layer.hpp:94) */
  int64_t n_outbox_right;  /* = particles_right.size() (layer.hpp:95) */
  int64_t events;       /* particle_step executions (layer.cpp:123) */
  int64_t scatters;     /* events that took the di < di_edge branch (layer.cpp:160) */
  int64_t n_left;       /* histories classified -1 (layer.cpp:202-205,217) */
  int64_t n_right;      /* histories classified +1 (layer.cpp:207-210) */
  int64_t n_dead;       /* histories classified 0  (layer.cpp:212-215) */
  double w_left;        /* weight carried out to the left / right / by dead */
  double w_right;
  double w_dead;
  int64_t launches;     /* tracking-kernel launches so far */
  double track_ms;      /* summed device time of those launches (CUDA events) */
  int64_t gpu_launches; /* ALL kernels this layer launched (tracking, birth, transposes) */
} mcb200_counts;

/* ---- life cycle ------------------------------------------------------- */
/* Layer::Layer, src/layer.cpp:44-69 */
int mcb200_layer_create(const mcb200_layer_desc *desc, mcb200_layer **out);
void mcb200_layer_destroy(mcb200_layer *l);
/* deep copy (Layer is copy-constructed / returned by value in the reference:
 * src/layer.cpp:41, include/mcmpi/worker.hpp:60) */
int mcb200_layer_clone(mcb200_layer *src, mcb200_layer **out);
/* overwrite the public, mutable `sigs` / `absorption_rates` vectors
 * (include/layer/layer.hpp:103-104); either pointer may be NULL = keep */
int mcb200_layer_set_cross_sections(mcb200_layer *l, const float *sigs,
                                    const float *absorption_rates);
int mcb200_layer_get_cross_sections(mcb200_layer *l, float *sigs_out,
                                    float *absorption_rates_out);

/* the cross-sections Layer::Layer hard-codes (src/layer.cpp:53-63) for a layer of
 * m cells spanning [x_min, x_max]: sigs[i] = expf(-x_mid_i), absorption = 0.5.
 * Host-only helper (no GPU needed).  A K-GPU run that must equal the 1-GPU run
 * slices the table of the WHOLE slab instead of letting every sub-slab recompute
 * its own from rounded bounds (which moves ~20-30 % of the entries by 1 ulp). */
int mcb200_default_cross_sections(float x_min, float x_max, int32_t m,
                                  float *sigs_out, float *absorption_rates_out);

/* ---- particle sources ------------------------------------------------- */
/* Layer::create_particles(x_ini, wmc, n, seed), src/layer.cpp:71-82: registers
 * n unborn particles; they are born ON THE DEVICE (rnd_seed chain via LCG
 * jump-ahead, first rnd_real draw -> mu, src/layer.cpp:101-120) when
 * simulate() needs them.  No-op when x_ini is outside (x_min, x_max). */
int mcb200_layer_create_particles(mcb200_layer *l, float x_ini, float wmc,
                                  int64_t n, uint64_t seed);
/* append n host particles to the bank -- what the workers do on receive
 * (src/worker_sync.cpp:47-108, src/async_comm.cpp:140-143, src/rma_comm.cpp:210) */
int mcb200_layer_push(mcb200_layer *l, const mcb200_particle *aos, int64_t n);
/* same, from DEVICE memory on the layer's GPU (24-byte records).  Asynchronous: the
 * source must stay untouched until the next simulate() / pop on this layer. */
int mcb200_layer_push_device(mcb200_layer *l, const void *dev_aos, int64_t n);

/* ---- the hot path ----------------------------------------------------- */
/* Layer::simulate(nb_particles), src/layer.cpp:239-361: tracks the last
 * `nb_particles` of the bank (births included) until each has left the
 * sub-slab or dropped below particle_min_weight; -1 = until nothing is left
 * (simulate_helper, :220-237).  Blocking.  `counts` may be NULL. */
int mcb200_layer_simulate(mcb200_layer *l, int64_t nb_particles,
                          mcb200_counts *counts);

/* The same for particles that sit in HOST memory, what Layer::simulate(.., use_gpu) /
 * cusimulate do with `particles` (src/layer.cpp:257-263, src/culayer.cu:41-92): appends the n
 * records and tracks exactly them, in chunks -- the host-to-device copy of chunk k+1 runs under
 * the tracking of chunk k (pinned host memory makes the copies asynchronous; pageable memory
 * works, without the overlap).  Equivalent to push(aos, n) + simulate(n). */
int mcb200_layer_simulate_host(mcb200_layer *l, const mcb200_particle *aos, int64_t n,
                               mcb200_counts *counts);

/* ---- results ---------------------------------------------------------- */
int mcb200_layer_counts(mcb200_layer *l, mcb200_counts *out);
/* particles_left / particles_right (layer.hpp:94-95): copy up to `cap`
 * escapees to the host and clear the outbox (what the workers do after
 * sending: worker_sync.cpp:63,75).  Order within an outbox is unspecified. */
int mcb200_layer_pop_left(mcb200_layer *l, mcb200_particle *aos, int64_t cap,
                          int64_t *n_out);
int mcb200_layer_pop_right(mcb200_layer *l, mcb200_particle *aos, int64_t cap,
                           int64_t *n_out);
/* same, into DEVICE memory on the layer's GPU (24-byte records) -- the
 * buffer a neighbour exchange (NCCL send / P2P copy) ships */
int mcb200_layer_pop_left_device(mcb200_layer *l, void *dev_aos, int64_t cap,
                                 int64_t *n_out);
int mcb200_layer_pop_right_device(mcb200_layer *l, void *dev_aos, int64_t cap,
                                  int64_t *n_out);
/* zero-copy variant for a device-side exchange: the outbox itself (side 0 = left,
 * 1 = right) as `*n_out` contiguous 24-byte records in device memory, valid until the
 * next simulate() on this layer; mcb200_layer_outbox_clear() empties it once shipped. */
int mcb200_layer_outbox_device(mcb200_layer *l, int32_t side, void **dev_aos_out,
                               int64_t *n_out);
int mcb200_layer_outbox_clear(mcb200_layer *l, int32_t side);
/* ---- direct peer exchange over NVLink ------------------------------------ */
/* The MPI workers move escapees with Sendrecv / Isend / MPI_Put
 * (src/worker_sync.cpp:47-108, src/async_comm.cpp, src/rma_comm.cpp:133-186).  Here a layer
 * can own an INBOX that its neighbours' tracking kernels store into directly (peer-mapped
 * device memory: CUDA IPC between processes, a plain pointer inside one process): the escapee
 * stores of the kernel are the communication, like RmaComm's MPI_Put into the neighbour's
 * window.  The inbox is double-buffered by a parity bit so that cycle c+1 can be written
 * while cycle c is being ingested.  Protocol per cycle, on every rank:
 *   set_exchange_parity(c & 1); simulate(..);  <barrier / all-gather across ranks>;
 *   ingest_inbox(from left, c & 1); ingest_inbox(from right, c & 1)                          */
#define MCB200_IPC_HANDLE_BYTES 64
typedef struct mcb200_inbox_geom {
  int32_t nstripes;     /* stripes per slot = max CTAs of a sender's launch */
  int32_t stripe_cap;   /* records per stripe */
  int64_t ovf_cap;      /* records of the overflow segment */
  int64_t max_take;     /* max particles a sender may track per launch */
  int64_t slot_bytes;   /* bytes of one (side, parity) slot; 4 slots in the block */
  int64_t fills_offset; /* byte offset of the fill counters inside a slot */
} mcb200_inbox_geom;
/* allocate this layer's inbox for senders that track at most `max_take` particles per
 * launch; returns the IPC handle other PROCESSES open and the geometry they need */
int mcb200_layer_inbox_create(mcb200_layer *l, int64_t max_take,
                              uint8_t handle_out[MCB200_IPC_HANDLE_BYTES],
                              mcb200_inbox_geom *geom_out);
/* route this layer's escapees on `side` (0 left, 1 right) into a neighbour's inbox:
 * _peer: the neighbour lives in another process (handle + geometry from its inbox_create);
 * _local: the neighbour is `other`, in this process (any device with peer access)        */
int mcb200_layer_connect_peer(mcb200_layer *l, int32_t side,
                              const uint8_t handle[MCB200_IPC_HANDLE_BYTES],
                              const mcb200_inbox_geom *geom);
int mcb200_layer_connect_local(mcb200_layer *l, int32_t side, mcb200_layer *other);
/* unmap the neighbours' inboxes (before THEY are destroyed; escapees go to the local
 * outboxes again) */
int mcb200_layer_disconnect_peers(mcb200_layer *l);
int mcb200_layer_set_exchange_parity(mcb200_layer *l, int32_t parity);
/* append what the neighbour on `from_side` stored in this layer's inbox (given parity) to
 * the bank and reset that slot; *n_out = particles received */
int mcb200_layer_ingest_inbox(mcb200_layer *l, int32_t from_side, int32_t parity,
                              int64_t *n_out);


/* ---- the world: one whole run on N GPUs, driven from C ------------------------------ */
/* Stands in for Worker::spin (src/worker_sync.cpp:24-135, src/worker_async.cpp:19-105,
 * src/worker_rma.cpp:15-68) and Worker::gather_weights_absorbed (src/worker.cpp:183-216).
 * One `mcb200_world` = one rank = one contiguous sub-slab on one GPU (decompose_domain,
 * src/layer.cpp:17-42, but with the ONE global dx and slices of the ONE global
 * cross-section table, so that K ranks reproduce the single-rank result bit for bit).
 * A run is ONE resident kernel per rank: escapees go straight into the neighbour GPU's
 * memory (peer-mapped rings, the MPI_Put of src/rma_comm.cpp:133-186), source particles
 * are born in the kernel (src/layer.cpp:101-120), and the run ends when a device-side
 * global count of disabled histories reaches nb_particles (StateComm,
 * src/state_comm.cpp:35-65; MPI_Allreduce, src/worker_sync.cpp:112-120).  The host only
 * launches and waits.  A sub-slab too wide for a CTA-private tally is cut into windows
 * served by groups of CTAs of the same launch, exchanging through the same rings.
 *
 * Ranks in ONE process (one host thread drives N GPUs):
 *   create x N; connect_local(all pairs); mcb200_world_run(worlds, N, ...)
 * Ranks in N processes (one per GPU, e.g. under torchrun / mpirun):
 *   create; export -> ship handle + geom to the other ranks (any transport) -> connect_peer;
 *   per run: <barrier> prepare; <barrier>; launch; wait.  The first barrier orders "every
 *   rank's previous wait() has returned" before anybody's prepare (which wipes the rings a
 *   neighbour's finishing kernel may still return credits into); the second orders every
 *   prepare before anybody's launch.                                                         */
typedef struct mcb200_world mcb200_world;

typedef struct mcb200_world_desc {
  int32_t abi_version;       /* MCB200_ABI_VERSION */
  int32_t device;            /* CUDA device ordinal of this rank */
  int32_t rank, world_size;
  float x_min, x_max, x_ini; /* config.yaml keys (src/worker.cpp:315-333) */
  int32_t nb_cells;
  float particle_min_weight;
  const int32_t *cuts;       /* world_size + 1 ascending cell boundaries, cuts[0] = 0,
                                cuts[world_size] = nb_cells; NULL = the reference's equal
                                split (src/layer.cpp:24-27).  Results do not depend on them. */
  const float *sigs;              /* GLOBAL tables, nb_cells entries; NULL = src/layer.cpp:53-63 */
  const float *absorption_rates;
  int32_t windows;           /* windows per rank; 0 = as few as fit shared memory */
  int32_t block;             /* threads per CTA; 0 = auto */
  int32_t max_ctas;          /* cap on the CTAs of the launch, counted as 256-thread CTAs (ranks that
                                share one GPU must all be resident at once); 0 = fill the GPU */
  int32_t ring_cap;          /* records per ring (power of two >= 32); 0 = auto */
  int32_t retire_batch;      /* see mcb200_layer_set_option; 0 = auto */
  int32_t xs_global;         /* 1: keep the cell constants in global memory / L2 instead of shared
                                memory (halves the shared memory a cell costs); 0 = only when needed */
  int64_t bank_cap;          /* records of one CTA's overflow bank (power of two); 0 = auto */
  int64_t inflight_limit;    /* source births pause above this many live histories; 0 = auto */
} mcb200_world_desc;

/* what a peer must know about a rank's exchange block to store into it */
typedef struct mcb200_world_geom {
  int32_t rank, world_size;
  int32_t stripes;           /* rings per link = warps serving one window */
  int32_t ring_cap;
  int64_t block_bytes;
  int64_t off_rec[2];        /* rings filled by the left [0] / right [1] neighbour */
  int64_t off_credit[2];     /* credits for this rank's own sends to the left / right */
  int64_t off_chain;         /* per-stripe counts of disabled histories (used on the home rank) */
} mcb200_world_geom;

/* one rank's share of a run (counters are for THIS run; the tally is cumulative) */
typedef struct mcb200_world_result {
  int64_t events, scatters;        /* particle_step executions / scatter branches on this rank */
  int64_t n_left, n_right, n_dead; /* histories disabled here: absorbed at the GLOBAL left /
                                      right border (src/layer.cpp:350-360), below min weight */
  int64_t births;                  /* source histories born here */
  int64_t sent_left, sent_right;   /* records shipped to the neighbour ranks (NVLink) */
  int64_t window_crossings;        /* records exchanged between windows inside this rank */
  int64_t idle_polls, blocked_passes, bank_pushes, bank_pops;   /* exchange diagnostics */
  int64_t lane_slots;              /* lane x event-iteration slots offered: events / lane_slots =
                                      fraction of lanes that carried a live history */
  int64_t idle_warp_ns;            /* summed over the warps of the launch: time a warp spent without
                                      a single live history (waiting for neighbours / the end) */
  double w_left, w_right, w_dead;  /* cumulative weight absorbed at the borders / by the dead */
  double kernel_ms;                /* device time of the resident kernel (CUDA events) */
  int32_t windows, ctas, block, stripes, ring_cap;
  int32_t error;                   /* MCB200_OK or the MCB200_ERR_* the kernel raised */
} mcb200_world_result;

int mcb200_world_create(const mcb200_world_desc *desc, mcb200_world **out);
void mcb200_world_destroy(mcb200_world *w);
/* the IPC handle + geometry of this rank's exchange block, for ranks in other processes */
int mcb200_world_export(mcb200_world *w, uint8_t handle_out[MCB200_IPC_HANDLE_BYTES],
                        mcb200_world_geom *geom_out);
/* map another rank's exchange block: every rank connects to every other rank (neighbours for
 * the rings, the home rank -- the one holding x_ini -- for the global count and `done`) */
int mcb200_world_connect_peer(mcb200_world *w, int32_t peer_rank,
                              const uint8_t handle[MCB200_IPC_HANDLE_BYTES],
                              const mcb200_world_geom *geom);
int mcb200_world_connect_local(mcb200_world *w, mcb200_world *peer);
int mcb200_world_disconnect(mcb200_world *w);
/* a run of nb_particles source histories (seed chain from `seed`, src/layer.cpp:36):
 * <barrier>, prepare on EVERY rank, <barrier>, launch, wait (see above). */
int mcb200_world_prepare(mcb200_world *w, int64_t nb_particles, uint64_t seed);
int mcb200_world_launch(mcb200_world *w);
int mcb200_world_wait(mcb200_world *w, mcb200_world_result *out);
/* all ranks live in this process: prepare / launch / wait for each (results may be NULL) */
int mcb200_world_run(mcb200_world *const *worlds, int32_t n, int64_t nb_particles,
                     uint64_t seed, mcb200_world_result *results);
/* this rank's cells [*lo, *lo + *m) */
int mcb200_world_cells(mcb200_world *w, int32_t *lo, int32_t *m);
/* this rank's slice of weights_absorbed (m entries; see mcb200_layer_weights_absorbed*) */
int mcb200_world_tally(mcb200_world *w, float *out_m);
int mcb200_world_tally_f64(mcb200_world *w, double *out_m);
int mcb200_world_tally_exact(mcb200_world *w, uint32_t *out_4m, int32_t *lsb_log2);
int mcb200_world_reset_tally(mcb200_world *w);
/* Worker::gather_weights_absorbed (src/worker.cpp:183-216) for ranks of one process: the
 * disjoint slices concatenated, nb_cells entries */
int mcb200_world_gather_tally_f64(mcb200_world *const *worlds, int32_t n, double *out_nb_cells);
/* knobs: "max_run_ms" (device-side cap on a run, 0 = none), "stall_ms" (the host stops a run
 * whose global counters have not moved for this long; default 15000), "rng" (0 = LCG parity
 * mode, 1 = Philox2x32-10, see mcb200_layer_set_option), "retire_batch", "inflight_limit" */
int mcb200_world_set_option(mcb200_world *w, const char *key, int64_t value);
void *mcb200_world_stream(mcb200_world *w);

/* weights_absorbed (layer.hpp:92), m entries.  The device keeps every cell as
 * an EXACT 128-bit fixed-point sum of the per-event float deposits (order-,
 * launch- and GPU-count-independent).  _exact returns it: four little-endian
 * 32-bit digits per cell, out[4*c + j], two's complement, value = integer *
 * 2^lsb_log2 (lsb_log2 = -120).  _f64 / the float version the reference
 * exposes are that number rounded once. */
int mcb200_layer_weights_absorbed(mcb200_layer *l, float *out_m);
int mcb200_layer_weights_absorbed_f64(mcb200_layer *l, double *out_m);
int mcb200_layer_weights_absorbed_exact(mcb200_layer *l, uint32_t *out_4m,
                                        int32_t *lsb_log2);
/* zero weights_absorbed and the class weights (the reference's callers do
 * `std::fill(weights_absorbed...)` on the public vector between runs) */
int mcb200_layer_reset_tally(mcb200_layer *l);
/* Layer::dump_WA(), src/layer.cpp:363-380, same "%.4e %.3e\n" text; path NULL
 * -> "WA.out" in the current directory like the reference */
int mcb200_layer_dump_WA(mcb200_layer *l, const char *path);

/* ---- plumbing --------------------------------------------------------- */
/* the cudaStream_t all of this layer's work is queued on (for CUDA-event
 * timing by the caller) */
void *mcb200_layer_stream(mcb200_layer *l);
/* tuning knobs; key/value, returns MCB200_ERR_INVALID for unknown keys:
 *   "tally_mode"   0 auto, 1 CTA-private shared-memory tally, 2 global (L2) tally
 *   "block"        threads per CTA,  "blocks_per_sm" CTAs per SM
 *   "retire_batch" lanes of a warp without a live history before retire/refill runs (0 auto)
 *   "rng"          0 = the reference's per-particle LCG stream (src/random.cpp:12-16): results
 *                  equal the CPU reference's bit for bit (the default, the parity mode);
 *                  1 = counter-based Philox2x32-10 on (history id, event number), keyed by the
 *                  seed given to create_particles: no seed chain, statistically equivalent
 *                  results (tests/test_gpu_philox.py states the acceptance bounds)
 *   "birth_chunk"  max particles born per launch
 *   "host_chunk"   max particles per chunk of mcb200_layer_simulate_host */
int mcb200_layer_set_option(mcb200_layer *l, const char *key, int64_t value);

const char *mcb200_last_error(void);
int mcb200_abi_version(void);
int mcb200_device_count(void);

/* ---- primitives exposed for known-answer tests ------------------------- */
/* rnd_real on the device, one draw per element: seeds[i] advanced in place,
 * out[i] = the float (src/curandom.cu:7-14, checked like src/test_curandom.cu) */
int mcb200_test_rnd_real(int device, uint64_t *seeds_host, float *out_host,
                         int64_t n);
/* Philox2x32-10 on the device (the "rng" = 1 generator): out[2i], out[2i+1] = the two output
 * words for counter (c0[i], c1[i]) and key[i]; checked against Random123's known answers */
int mcb200_test_philox(int device, const uint32_t *c0_host, const uint32_t *c1_host,
                       const uint32_t *key_host, uint32_t *out2_host, int64_t n);
/* device logf / expf as used by the tracking kernel, element-wise */
int mcb200_test_logf(int device, const float *in_host, float *out_host, int64_t n);
int mcb200_test_expf(int device, const float *in_host, float *out_host, int64_t n);
/* di_edge = (x_edge - x) / mu (src/layer.cpp:154-158) as the kernel computes it: IEEE
 * division with the reciprocal of mu hoisted out of the event loop; out = FLT_MAX where
 * |mu| <= 1e-4 */
int mcb200_test_edge_distance(int device, const float *a_host, const float *mu_host,
                              float *out_host, int64_t n);
/* the exact accumulator the tally uses: sum of n floats -> 4 digits (+ rounded double) */
int mcb200_test_accumulate(int device, const float *in_host, int64_t n,
                           uint32_t *out4, double *out_f64);
/* device-side birth only: first n particles of create_particles(x_ini,wmc,.,seed) */
int mcb200_test_birth(int device, float x_ini, float wmc, float dx, int64_t n,
                      uint64_t seed, mcb200_particle *out_host);

#ifdef __cplusplus
}
#endif
#endif /* MCB200_H */
