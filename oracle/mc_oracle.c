/*
 * mc_oracle.c -- CPU restatement of mc-mpi's particle-tracking path.
 *
 * TEST INFRASTRUCTURE ONLY (see mc_oracle.h).  Parity status: PINNED against
 * oracle/_ref (the unmodified reference sources) and the reference's golden
 * file by tests/test_oracle_pin.py.
 *
 * Written from the behaviour of the reference (citations are file:line in the
 * reference repository); plain C, float arithmetic evaluated exactly as the
 * reference's x86-64 Release build does: no FMA contraction, libm logf/expf.
 * Build flags (oracle/Makefile): -O2 -ffp-contract=off, no -ffast-math, no
 * -march.
 */
#include "mc_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* include/types/types.hpp:14-16 */
#define ORC_EPS 1e-4F
#define ORC_MAXREAL FLT_MAX
/* src/layer.cpp:15 */
#define ORC_MAX_PARTICLES_VECTOR 1000000

/* ------------------------------------------------------------------ RNG -- */

/* src/random.cpp:8-16: 64-bit LCG modulo 2^63, float = state * 2^-63 (the
 * u64 -> f32 conversion rounds to nearest, so 1.0f is reachable). */
float orc_rnd_real(uint64_t *seed) {
  const uint64_t g = 6364136223846793005ull, c = 1442695040888963407ull;
  const uint64_t p = (uint64_t)1 << 63;
  const float inv_p = (float)1 / (float)p;
  *seed = (g * *seed + c) % p;
  return (float)(*seed) * inv_p;
}

/* src/random.cpp:20-29: the per-particle seed chain (other multiplier). */
uint64_t orc_rnd_seed(uint64_t *seed) {
  const uint64_t g = 5177284530976225183ull, c = 2096348467109453893ull;
  const uint64_t p = (uint64_t)1 << 63;
  *seed = (g * *seed + c) % p;
  return *seed;
}

/* ------------------------------------------------------- small vectors -- */

typedef struct pvec {
  orc_particle *d;
  int n, cap;
} pvec;

static void pvec_reserve(pvec *v, int cap) {
  if (cap <= v->cap) return;
  int nc = v->cap ? v->cap : 1024;
  while (nc < cap) nc *= 2;
  v->d = (orc_particle *)realloc(v->d, (size_t)nc * sizeof(orc_particle));
  if (!v->d) {
    fprintf(stderr, "mc_oracle: out of memory\n");
    exit(EXIT_FAILURE);
  }
  v->cap = nc;
}
static void pvec_push(pvec *v, const orc_particle *p) {
  pvec_reserve(v, v->n + 1);
  v->d[v->n++] = *p;
}

/* --------------------------------------------------------------- layer -- */

struct orc_layer {
  /* include/layer/layer.hpp:85-109 */
  float x_min, x_max;
  int m, index_start;
  float dx;
  int left_border, right_border;
  float particle_min_weight;
  float *sigs, *absorption_rates, *weights_absorbed;
  pvec particles, particles_left, particles_right;
  int nb_disabled;
  uint64_t seed;
  float x_ini, wmc;
  int nb_particles_create;
  /* checker-side instrumentation */
  double *tally_f64;
  unsigned __int128 *tally_x; /* exact: integer * 2^ORC_ACC_LSB_LOG2, two's complement */
  unsigned __int128 cls_x[3]; /* exact weight carried left / right / by the dead */
  int keep_border;
  pvec absorbed_left, absorbed_right, dead;
  orc_stats st;
};

/* src/layer.cpp:44-69 */
orc_layer *orc_layer_new(float x_min, float x_max, int index_start, int m,
                         float particle_min_weight) {
  orc_layer *l = (orc_layer *)calloc(1, sizeof(orc_layer));
  l->x_min = x_min;
  l->x_max = x_max;
  l->m = m;
  l->index_start = index_start;
  l->dx = (x_max - x_min) / m; /* :47 float / int */
  l->left_border = fabs((double)x_min) < (double)ORC_EPS;          /* :47 */
  l->right_border = fabs((double)x_max - 1.0) < (double)ORC_EPS;   /* :48 */
  l->particle_min_weight = particle_min_weight;
  l->sigs = (float *)malloc(sizeof(float) * (size_t)(m > 0 ? m : 1));
  l->absorption_rates = (float *)malloc(sizeof(float) * (size_t)(m > 0 ? m : 1));
  l->weights_absorbed = (float *)calloc((size_t)(m > 0 ? m : 1), sizeof(float));
  l->tally_f64 = (double *)calloc((size_t)(m > 0 ? m : 1), sizeof(double));
  l->tally_x = (unsigned __int128 *)calloc((size_t)(m > 0 ? m : 1), sizeof(unsigned __int128));
  for (int i = 0; i < m; ++i) {
    /* :58 -- (x_min + i*dx) in float, "+ 0.5*dx" in double, narrowed to float */
    float x_mid = (float)((double)(x_min + (i * l->dx)) + 0.5 * (double)l->dx);
    l->sigs[i] = expf(-x_mid); /* :59, exp(float) resolves to expf in C++ */
    l->absorption_rates[i] = 0.5f; /* :63 */
  }
  pvec_reserve(&l->particles, 10000); /* :51,68 */
  return l;
}

void orc_layer_free(orc_layer *l) {
  if (!l) return;
  free(l->sigs);
  free(l->absorption_rates);
  free(l->weights_absorbed);
  free(l->tally_f64);
  free(l->tally_x);
  free(l->particles.d);
  free(l->particles_left.d);
  free(l->particles_right.d);
  free(l->absorbed_left.d);
  free(l->absorbed_right.d);
  free(l->dead.d);
  free(l);
}

/* src/layer.cpp:89-121 */
static void create_particles_n(orc_layer *l, int n) {
  if (l->nb_particles_create <= 0) return;
  int best = ORC_MAX_PARTICLES_VECTOR - l->particles.n; /* :95 */
  if (best > n) n = best;                                /* :96 */
  if (l->nb_particles_create < n) n = l->nb_particles_create; /* :98 */
  l->nb_particles_create -= n;

  orc_particle p;
  memset(&p, 0, sizeof(p));
  p.x = l->x_ini;
  p.wmc = l->wmc;
  p.index = (int)(l->x_ini / l->dx); /* :106, ignores x_min */
  pvec_reserve(&l->particles, l->particles.n + n);
  for (int i = 0; i < n; ++i) {
    p.seed = orc_rnd_seed(&l->seed);           /* :111 */
    p.mu = 2 * orc_rnd_real(&p.seed) - 1;      /* :112, draw #1 of the stream */
    pvec_push(&l->particles, &p);
  }
}

/* src/layer.cpp:71-82 */
void orc_create_particles(orc_layer *l, float x_ini, float wmc, int n,
                          uint64_t seed) {
  if (x_ini > l->x_min && x_ini < l->x_max) {
    l->x_ini = x_ini;
    l->wmc = wmc;
    l->nb_particles_create = n;
    l->seed = seed;
    create_particles_n(l, ORC_MAX_PARTICLES_VECTOR);
  }
}

/* src/layer.cpp:17-42 */
orc_layer *orc_decompose_domain(float x_min, float x_max, float x_ini,
                                int world_size, int world_rank, int nb_cells,
                                int nb_particles, float particle_min_weight) {
  int cells_per_layer = nb_cells / world_size;
  int num_with_extra = nb_cells % world_size;
  int nb_my_cells = cells_per_layer + (world_rank < num_with_extra);
  int start_index = world_rank * cells_per_layer +
                    (world_rank < num_with_extra ? world_rank : num_with_extra);
  float dx = (x_max - x_min) / ((float)nb_cells);
  int cell_ini = (int)((x_ini - x_min) / dx);
  orc_layer *l = orc_layer_new(x_min + start_index * dx,
                               x_min + (start_index + nb_my_cells) * dx,
                               start_index, nb_my_cells, particle_min_weight);
  if (cell_ini >= start_index && cell_ini < start_index + nb_my_cells) {
    uint64_t seed = 5127801; /* :36 */
    orc_create_particles(l, x_ini, (float)(1.0 / nb_particles), nb_particles,
                         seed);
  }
  return l;
}

/* src/layer.cpp:84-87 */
int orc_nb_active(const orc_layer *l) {
  return l->particles.n + l->nb_particles_create;
}

/*
 * Exact accumulation (checker side of the CUDA path's long accumulator): a
 * float is its 24-bit significand times a power of two, so with a fixed
 * least significant bit of 2^ORC_ACC_LSB_LOG2 every deposit is an integer and
 * the 128-bit two's-complement sum is exact and order-independent.  Bits below
 * the LSB (floats under 2^-97) are dropped toward zero, like on the device.
 */
#define ORC_ACC_LSB_LOG2 (-120)
static inline void acc_add_exact(unsigned __int128 *acc, float v) {
  uint32_t b;
  memcpy(&b, &v, 4);
  uint32_t e = (b >> 23) & 0xffu;
  uint32_t mant = (b & 0x7fffffu) | (e ? 0x800000u : 0u);
  int pos = (int)(e ? e : 1u) - (150 + ORC_ACC_LSB_LOG2);
  if (pos < 0) {
    mant = pos > -24 ? mant >> (-pos) : 0u;
    pos = 0;
  }
  if (mant == 0u || pos > 103) return; /* zero, or not a weight (>= 2^7, inf, nan) */
  unsigned __int128 d = (unsigned __int128)mant << pos;
  if (b >> 31) *acc -= d; else *acc += d;
}
static inline double acc_to_double(unsigned __int128 a) {
  return ldexp((double)(__int128)a, ORC_ACC_LSB_LOG2);
}

/* accumulators one worker thread tallies into */
typedef struct tally_ctx {
  float *wl;     /* thread-private float tally, like :314-315 */
  double *wd;
  unsigned __int128 *wx;
  int64_t events, scatters;
} tally_ctx;

/* src/layer.cpp:123-190 -- one event */
static inline void particle_step(const orc_layer *l, orc_particle *p,
                                 tally_ctx *t) {
  const int il = p->index - l->index_start;
  const float a = l->absorption_rates[il];
  const float interaction_rate = (float)(1.0 - (double)a); /* :131 */
  const float sig_a = l->sigs[il] * a;                     /* :132 */
  const float sig_i = l->sigs[il] * interaction_rate;      /* :133 */

  const float h = orc_rnd_real(&p->seed);                       /* :136 */
  float di = sig_i > ORC_EPS ? -logf(h) / sig_i : ORC_MAXREAL;  /* :137 */

  int index_new;
  float x_new_edge;
  if (p->mu < 0) { /* :143-152 */
    index_new = p->index - 1;
    x_new_edge = p->index * l->dx;
  } else {
    index_new = p->index + 1;
    x_new_edge = (p->index + 1) * l->dx;
  }

  float di_edge = ORC_MAXREAL; /* :154-158 */
  if (p->mu < -ORC_EPS || ORC_EPS < p->mu) {
    di_edge = (x_new_edge - p->x) / p->mu;
  }

  if (di < di_edge) { /* :160-166 scatter inside the cell */
    index_new = p->index;
    const float step = di * p->mu;
    p->x += step;
    p->mu = 2 * orc_rnd_real(&p->seed) - 1;
    t->scatters++;
  } else { /* :167-172 move onto the edge */
    di = di_edge;
    p->x = x_new_edge;
  }

  const float dw = (1 - expf(-sig_a * di)) * p->wmc; /* :175 */
  p->wmc -= dw;                                      /* :178 */
  t->wl[il] += dw;                                   /* :179 */
  t->wd[il] += (double)dw;
  acc_add_exact(&t->wx[il], dw);
  p->index = index_new;                              /* :181 */
  t->events++;
}

/* src/layer.cpp:192-218 */
static inline int simulate_particle(const orc_layer *l, orc_particle *p,
                                    tally_ctx *t) {
  while ((p->wmc >= l->particle_min_weight) &&
         (p->index < l->index_start + l->m) && (p->index >= l->index_start)) {
    particle_step(l, p, t);
  }
  if (p->index == l->index_start - 1) return -1;
  if (p->index == l->index_start + l->m) return +1;
  if (p->wmc < l->particle_min_weight) return 0;
  return -1;
}

static void simulate_impl(orc_layer *l, int nb_particles, int nthread);

/* src/layer.cpp:220-237 */
static void simulate_helper(orc_layer *l, int nb_particles, int nthread) {
  if (nb_particles == -1) {
    while ((l->particles.n > 0) || (l->nb_particles_create > 0)) {
      simulate_impl(l, ORC_MAX_PARTICLES_VECTOR, nthread);
    }
  }
  while ((nb_particles > 0) &&
         (l->particles.n > 0 || l->nb_particles_create > 0)) {
    int this_call = nb_particles < ORC_MAX_PARTICLES_VECTOR
                        ? nb_particles
                        : ORC_MAX_PARTICLES_VECTOR;
    simulate_impl(l, this_call, nthread);
    nb_particles -= this_call;
  }
}

/* src/layer.cpp:239-253, :303-361 (the CPU branch) */
static void simulate_impl(orc_layer *l, int nb_particles, int nthread) {
  if ((nb_particles == -1) || (nb_particles > ORC_MAX_PARTICLES_VECTOR))
    simulate_helper(l, nb_particles, nthread); /* :241-242, then falls through */

  if ((l->particles.n < nb_particles) && l->nb_particles_create > 0)
    create_particles_n(l, nb_particles); /* :244-245 */

  if (l->particles.n < nb_particles) nb_particles = l->particles.n; /* :247 */
  if (nb_particles <= 0) return;

  int *result = (int *)malloc(sizeof(int) * (size_t)nb_particles);
  const int particles_size = l->particles.n;
  const int m = l->m;
  int64_t ev = 0, sc = 0;

  if (nthread <= 1) {
    /* one thread: thread-private tally zeroed per call, merged after (:314-329) */
    tally_ctx t;
    t.wl = (float *)calloc((size_t)m, sizeof(float));
    t.wd = l->tally_f64;
    t.wx = l->tally_x;
    t.events = t.scatters = 0;
    for (int i = 0; i < nb_particles; i++) /* :317-321, bank consumed backwards */
      result[i] = simulate_particle(l, &l->particles.d[particles_size - 1 - i], &t);
    for (int j = 0; j < m; j++) l->weights_absorbed[j] += t.wl[j];
    free(t.wl);
    ev = t.events;
    sc = t.scatters;
  } else {
#ifdef _OPENMP
    omp_set_num_threads(nthread); /* :306-309 */
#endif
    float **wls = (float **)calloc((size_t)nthread, sizeof(float *));
    double **wds = (double **)calloc((size_t)nthread, sizeof(double *));
    unsigned __int128 **wxs =
        (unsigned __int128 **)calloc((size_t)nthread, sizeof(unsigned __int128 *));
    int64_t *evs = (int64_t *)calloc((size_t)nthread * 2, sizeof(int64_t));
    int used = 1;
#pragma omp parallel
    {
#ifdef _OPENMP
      int tid = omp_get_thread_num();
#pragma omp single
      used = omp_get_num_threads();
#else
      int tid = 0;
#endif
      tally_ctx t;
      t.wl = (float *)calloc((size_t)m, sizeof(float));
      t.wd = (double *)calloc((size_t)m, sizeof(double));
      t.wx = (unsigned __int128 *)calloc((size_t)m, sizeof(unsigned __int128));
      t.events = t.scatters = 0;
#pragma omp for schedule(static)
      for (int i = 0; i < nb_particles; i++)
        result[i] = simulate_particle(l, &l->particles.d[particles_size - 1 - i], &t);
      wls[tid] = t.wl;
      wds[tid] = t.wd;
      wxs[tid] = t.wx;
      evs[2 * tid] = t.events;
      evs[2 * tid + 1] = t.scatters;
    }
    /* the reference merges under `omp critical` in arrival order (:323-329);
     * here in thread-id order so that a run is at least self-reproducible */
    for (int k = 0; k < used; k++) {
      if (!wls[k]) continue;
      for (int j = 0; j < m; j++) {
        l->weights_absorbed[j] += wls[k][j];
        l->tally_f64[j] += wds[k][j];
        l->tally_x[j] += wxs[k][j];
      }
      ev += evs[2 * k];
      sc += evs[2 * k + 1];
      free(wls[k]);
      free(wds[k]);
      free(wxs[k]);
    }
    free(wls);
    free(wds);
    free(wxs);
    free(evs);
  }
  l->st.events += ev;
  l->st.scatters += sc;

  for (int i = 0; i < nb_particles; i++) { /* :332-346 */
    const orc_particle *p = &l->particles.d[particles_size - 1 - i];
    switch (result[i]) {
    case -1:
      pvec_push(&l->particles_left, p);
      l->st.n_left++;
      l->st.w_left += (double)p->wmc;
      acc_add_exact(&l->cls_x[0], p->wmc);
      break;
    case 1:
      pvec_push(&l->particles_right, p);
      l->st.n_right++;
      l->st.w_right += (double)p->wmc;
      acc_add_exact(&l->cls_x[1], p->wmc);
      break;
    case 0:
      l->nb_disabled++;
      l->st.n_dead++;
      l->st.w_dead += (double)p->wmc;
      acc_add_exact(&l->cls_x[2], p->wmc);
      if (l->keep_border) pvec_push(&l->dead, p);
      break;
    }
  }
  free(result);
  l->particles.n -= nb_particles; /* :348 */

  if (l->left_border) { /* :350-354 */
    if (l->keep_border)
      for (int i = 0; i < l->particles_left.n; i++)
        pvec_push(&l->absorbed_left, &l->particles_left.d[i]);
    l->nb_disabled += l->particles_left.n;
    l->particles_left.n = 0;
  }
  if (l->right_border) { /* :356-360 */
    if (l->keep_border)
      for (int i = 0; i < l->particles_right.n; i++)
        pvec_push(&l->absorbed_right, &l->particles_right.d[i]);
    l->nb_disabled += l->particles_right.n;
    l->particles_right.n = 0;
  }
}

void orc_simulate(orc_layer *l, int nb_particles) {
  simulate_impl(l, nb_particles, 1);
}
void orc_simulate_mt(orc_layer *l, int nb_particles, int nthread) {
  simulate_impl(l, nb_particles, nthread);
}

/* src/layer.cpp:363-380 */
int orc_dump_WA(const orc_layer *l, const char *path) {
  FILE *f = fopen(path, "w");
  if (!f) return -1;
  for (int i = 0; i < l->m; ++i) {
    fprintf(f, "%.4e %.3e\n",
            (double)(l->x_min + (i * l->dx)) + 0.5 * (double)l->dx,
            (double)(l->weights_absorbed[i] / l->dx));
  }
  fclose(f);
  return 0;
}

void orc_push(orc_layer *l, const orc_particle *p, int n) {
  pvec_reserve(&l->particles, l->particles.n + n);
  memcpy(l->particles.d + l->particles.n, p, (size_t)n * sizeof(orc_particle));
  l->particles.n += n;
}

/* ----------------------------------------------------------- accessors -- */
int orc_m(const orc_layer *l) { return l->m; }
int orc_index_start(const orc_layer *l) { return l->index_start; }
float orc_dx(const orc_layer *l) { return l->dx; }
float orc_x_min(const orc_layer *l) { return l->x_min; }
float orc_x_max(const orc_layer *l) { return l->x_max; }
int orc_left_border(const orc_layer *l) { return l->left_border; }
int orc_right_border(const orc_layer *l) { return l->right_border; }
int orc_nb_disabled(const orc_layer *l) { return l->nb_disabled; }
int orc_nb_particles_create(const orc_layer *l) { return l->nb_particles_create; }
float *orc_sigs(orc_layer *l) { return l->sigs; }
float *orc_absorption_rates(orc_layer *l) { return l->absorption_rates; }
float *orc_weights_absorbed(orc_layer *l) { return l->weights_absorbed; }
int orc_particles_size(const orc_layer *l) { return l->particles.n; }
orc_particle *orc_particles(orc_layer *l) { return l->particles.d; }
int orc_particles_left_size(const orc_layer *l) { return l->particles_left.n; }
orc_particle *orc_particles_left(orc_layer *l) { return l->particles_left.d; }
int orc_particles_right_size(const orc_layer *l) { return l->particles_right.n; }
orc_particle *orc_particles_right(orc_layer *l) { return l->particles_right.d; }
void orc_clear_left(orc_layer *l) { l->particles_left.n = 0; }
void orc_clear_right(orc_layer *l) { l->particles_right.n = 0; }
double *orc_tally_f64(orc_layer *l) { return l->tally_f64; }
void *orc_tally_exact(orc_layer *l) { return l->tally_x; }
int orc_tally_exact_lsb_log2(void) { return ORC_ACC_LSB_LOG2; }
void orc_tally_exact_f64(const orc_layer *l, double *out_m) {
  for (int i = 0; i < l->m; i++) out_m[i] = acc_to_double(l->tally_x[i]);
}
/* exact class weights (left, right, dead) rounded once to double */
void orc_class_weights_exact(const orc_layer *l, double out3[3]) {
  for (int k = 0; k < 3; k++) out3[k] = acc_to_double(l->cls_x[k]);
}
/* exact sum of n floats: 16 bytes little-endian + the rounded double */
double orc_accumulate_exact(const float *in, int64_t n, void *out16) {
  unsigned __int128 a = 0;
  for (int64_t i = 0; i < n; i++) acc_add_exact(&a, in[i]);
  if (out16) memcpy(out16, &a, 16);
  return acc_to_double(a);
}
void orc_set_keep_border(orc_layer *l, int keep) { l->keep_border = keep; }
int orc_absorbed_left_size(const orc_layer *l) { return l->absorbed_left.n; }
orc_particle *orc_absorbed_left(orc_layer *l) { return l->absorbed_left.d; }
int orc_absorbed_right_size(const orc_layer *l) { return l->absorbed_right.n; }
orc_particle *orc_absorbed_right(orc_layer *l) { return l->absorbed_right.d; }
int orc_dead_size(const orc_layer *l) { return l->dead.n; }
orc_particle *orc_dead(orc_layer *l) { return l->dead.d; }
void orc_get_stats(const orc_layer *l, orc_stats *out) { *out = l->st; }

/* ------------------------------------------- restated glibc logf / expf -- */
/*
 * The only third-party arithmetic on the path is libm's logf / expf
 * (SURVEY.md 8c).  The reference links the system glibc (2.39 here), whose
 * single-precision log/exp are the ARM "optimized-routines" algorithms
 * (sysdeps/ieee754/flt-32/e_logf.c, e_expf.c, e_logf_data.c, e_exp2f_data.c;
 * both evaluate in double and round once).  They are restated here so that a
 * CPU test can show restated == libm over the path's whole input domain
 * (tests/test_oracle_pin.py; the exhaustive sweep is oracle/sweep_libm.c), and
 * the CUDA path implements the same published algorithm with DFMA/DMUL/DADD.
 */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint64_t d2u(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }
static inline double u2d(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }

static const double orc_log_tab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2},
    {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},
    {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3},
    {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4},
    {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5},
    {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},
    {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},
    {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},
    {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};

/* valid for x in {+0} U [2^-126, +inf): the path only feeds h in [0, 1] */
float orc_logf_restated(float x) {
  const double ln2 = 0x1.62e42fefa39efp-1;
  const double a0 = -0x1.00ea348b88334p-2, a1 = 0x1.5575b0be00b6ap-2,
               a2 = -0x1.ffffef20a4123p-2;
  uint32_t ix = f2u(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix == 0) return -INFINITY;
  uint32_t tmp = ix - 0x3f330000u;
  int i = (tmp >> 19) & 15;
  int k = (int32_t)tmp >> 23;
  uint32_t iz = ix - (tmp & 0xff800000u);
  double z = (double)u2f(iz);
  double r = fma(z, orc_log_tab[i][0], -1.0);
  double y0 = fma((double)k, ln2, orc_log_tab[i][1]);
  double r2 = r * r;
  double y = fma(a1, r, a2);
  y = fma(a0, r2, y);
  y = fma(y, r2, y0 + r);
  return (float)y;
}

/* 2^(i/32) correctly rounded, minus (i << 47): the exp2f_data table */
static const uint64_t orc_exp_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full,
    0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull, 0x3fef54873168b9aaull,
    0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull,
    0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull,
    0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull,
    0x3feea11473eb0187ull, 0x3feea589994cce13ull, 0x3feeace5422aa0dbull,
    0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull,
    0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full,
    0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

/* valid for x <= 88 (the path only feeds x in [-inf, +0]) */
float orc_expf_restated(float x) {
  const double shift = 0x1.8p+52, inv_ln2_n = 0x1.71547652b82fep+0 * 32;
  const double c0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32,
               c1 = 0x1.ebfce50fac4f3p-3 / 32 / 32,
               c2 = 0x1.62e42ff0c52d6p-1 / 32;
  if (!(x > -104.0f)) return 0.0f; /* underflow to +0 below -103.97, -inf */
  double z = inv_ln2_n * (double)x;
  double kd = z + shift;
  uint64_t ki = d2u(kd);
  kd -= shift;
  double r = z - kd;
  uint64_t t = orc_exp_tab[ki & 31] + (ki << 47);
  double s = u2d(t);
  double zz = fma(c0, r, c1);
  double r2 = r * r;
  double y = fma(c2, r, 1.0);
  y = fma(zz, r2, y);
  y = y * s;
  return (float)y;
}

/* vector forms for the tests: kind 0 = libm, 1 = restated */
void orc_logf_v(int kind, const float *in, float *out, int64_t n) {
  for (int64_t i = 0; i < n; i++) out[i] = kind ? orc_logf_restated(in[i]) : logf(in[i]);
}
void orc_expf_v(int kind, const float *in, float *out, int64_t n) {
  for (int64_t i = 0; i < n; i++) out[i] = kind ? orc_expf_restated(in[i]) : expf(in[i]);
}
/* n successive rnd_real draws of one stream (src/random.cpp:12-16) */
void orc_rnd_real_v(uint64_t *seed, float *out, int64_t n) {
  for (int64_t i = 0; i < n; i++) out[i] = orc_rnd_real(seed);
}
/* the per-particle seed chain (src/layer.cpp:111): out[i] = i-th rnd_seed */
void orc_rnd_seed_v(uint64_t *seed, uint64_t *out, int64_t n) {
  for (int64_t i = 0; i < n; i++) out[i] = orc_rnd_seed(seed);
}
