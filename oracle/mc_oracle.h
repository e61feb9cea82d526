/*
 * mc_oracle.h -- CPU restatement of mc-mpi's particle-tracking path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may build, load or call it.  The product path (mc_mpi_b200/) never
 * touches it and fails loudly when its CUDA library is missing.
 *
 * Parity status: PINNED.  tests/test_oracle_pin.py checks this restatement
 *   - bit-for-bit (float tally, every final particle state, every counter)
 *     against the unmodified reference sources compiled into oracle/_ref
 *     (oracle/Makefile, recipe only -- no reference source is copied), and
 *   - against the reference's own golden file data/test_layer_target_WA.out
 *     (byte-exact, src/test_layer.cpp:58-68) and the known answers of
 *     SURVEY.md Appendix B, committed under tests/golden/.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference repository root).
 */
#ifndef MC_ORACLE_H
#define MC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* include/types/particle.hpp:7-18 -- 24-byte POD, offsets 0/8/12/16/20 */
typedef struct orc_particle {
  uint64_t seed;
  float x;
  float mu;
  float wmc;
  int32_t index;
} orc_particle;

typedef struct orc_layer orc_layer;

/* src/random.cpp:12-16 and :24-29 */
float orc_rnd_real(uint64_t *seed);
uint64_t orc_rnd_seed(uint64_t *seed);

/* src/layer.cpp:44-69 (ctor) and :17-42 (decompose_domain) */
orc_layer *orc_layer_new(float x_min, float x_max, int index_start, int m,
                         float particle_min_weight);
orc_layer *orc_decompose_domain(float x_min, float x_max, float x_ini,
                                int world_size, int world_rank, int nb_cells,
                                int nb_particles, float particle_min_weight);
void orc_layer_free(orc_layer *l);

/* src/layer.cpp:71-82 */
void orc_create_particles(orc_layer *l, float x_ini, float wmc, int n,
                          uint64_t seed);
/* src/layer.cpp:239-361 (CPU branch), sequential == reference nthread=1 */
void orc_simulate(orc_layer *l, int nb_particles);
/* same semantics, OpenMP over particles; the float tally then depends on the
 * thread count exactly like the reference's does, the double / fixed-point
 * tallies and all particle states do not. */
void orc_simulate_mt(orc_layer *l, int nb_particles, int nthread);
/* src/layer.cpp:84-87 */
int orc_nb_active(const orc_layer *l);
/* src/layer.cpp:363-380, writes to `path` instead of ./WA.out */
int orc_dump_WA(const orc_layer *l, const char *path);

/* append particles to the bank (what the workers do when receiving:
 * src/worker_sync.cpp:47-108, src/async_comm.cpp:140-143) */
void orc_push(orc_layer *l, const orc_particle *p, int n);

/* public data of Layer (include/layer/layer.hpp:85-109) */
int orc_m(const orc_layer *l);
int orc_index_start(const orc_layer *l);
float orc_dx(const orc_layer *l);
float orc_x_min(const orc_layer *l);
float orc_x_max(const orc_layer *l);
int orc_left_border(const orc_layer *l);
int orc_right_border(const orc_layer *l);
int orc_nb_disabled(const orc_layer *l);
int orc_nb_particles_create(const orc_layer *l);
float *orc_sigs(orc_layer *l);               /* mutable, m entries */
float *orc_absorption_rates(orc_layer *l);   /* mutable, m entries */
float *orc_weights_absorbed(orc_layer *l);   /* float tally, m entries */
int orc_particles_size(const orc_layer *l);
orc_particle *orc_particles(orc_layer *l);
int orc_particles_left_size(const orc_layer *l);
orc_particle *orc_particles_left(orc_layer *l);
int orc_particles_right_size(const orc_layer *l);
orc_particle *orc_particles_right(orc_layer *l);
void orc_clear_left(orc_layer *l);
void orc_clear_right(orc_layer *l);

/*
 * Extra instrumentation the reference does not have (checker-side only).
 *  - double tally: the same per-event float dw summed in double.
 *  - exact tally: every float dw as an integer multiple of 2^-120, summed in
 *    128-bit two's complement (unsigned __int128 per cell, 16 little-endian
 *    bytes).  Integer adds are associative, so this is what the CUDA path's
 *    long accumulator is compared to bit-for-bit.
 *  - "keep_border": when non-zero, particles escaping through a global border
 *    are ALSO appended to absorbed_left / absorbed_right before the reference
 *    semantics (count as disabled, drop) are applied, so tests can compare
 *    their final states.
 */
double *orc_tally_f64(orc_layer *l);
void *orc_tally_exact(orc_layer *l);                 /* m x 16 bytes */
int orc_tally_exact_lsb_log2(void);
void orc_tally_exact_f64(const orc_layer *l, double *out_m);
void orc_class_weights_exact(const orc_layer *l, double out3[3]);
double orc_accumulate_exact(const float *in, int64_t n, void *out16);
void orc_set_keep_border(orc_layer *l, int keep);
int orc_absorbed_left_size(const orc_layer *l);
orc_particle *orc_absorbed_left(orc_layer *l);
int orc_absorbed_right_size(const orc_layer *l);
orc_particle *orc_absorbed_right(orc_layer *l);
int orc_dead_size(const orc_layer *l);
orc_particle *orc_dead(orc_layer *l);

typedef struct orc_stats {
  int64_t events;      /* particle_step calls */
  int64_t scatters;    /* events that took the di < di_edge branch */
  int64_t n_left;      /* simulate_particle returned -1 */
  int64_t n_right;     /* returned +1 */
  int64_t n_dead;      /* returned 0 */
  double w_left;       /* sum of wmc of particles that went left */
  double w_right;
  double w_dead;
} orc_stats;
void orc_get_stats(const orc_layer *l, orc_stats *out);

/* restated glibc 2.39 logf / expf (ARM optimized-routines algorithm); used by
 * tests to show that the published algorithm the CUDA path implements equals
 * the libm the reference links against.  The oracle itself calls libm. */
float orc_logf_restated(float x);
float orc_expf_restated(float x);
/* vector forms: kind 0 = libm, 1 = restated */
void orc_logf_v(int kind, const float *in, float *out, int64_t n);
void orc_expf_v(int kind, const float *in, float *out, int64_t n);
void orc_rnd_real_v(uint64_t *seed, float *out, int64_t n);
void orc_rnd_seed_v(uint64_t *seed, uint64_t *out, int64_t n);

#ifdef __cplusplus
}
#endif
#endif
