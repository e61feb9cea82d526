/*
 * ref_shim.cpp -- extern "C" window onto the UNMODIFIED reference Layer.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is ours; it is compiled together with
 * the reference's own src/layer.cpp + src/random.cpp taken from where they
 * lie under $(REF) (oracle/Makefile) into oracle/_ref/libmcref.so.  No
 * reference source is copied into this repository.  It is used (a) to pin
 * oracle/mc_oracle.c and (b) as the "reference" CPU baseline of bench.py.
 */
#include "layer.hpp"   /* $(REF)/include/layer/layer.hpp */
#include "random.hpp"  /* $(REF)/include/random/random.hpp */

#include <cstring>
#include <new>

extern "C" {

void *ref_decompose_domain(float x_min, float x_max, float x_ini, int world_size,
                           int world_rank, int nb_cells, int nb_particles,
                           float particle_min_weight) {
  return new Layer(decompose_domain(x_min, x_max, x_ini, world_size, world_rank,
                                    nb_cells, nb_particles, particle_min_weight));
}
void *ref_layer_new(float x_min, float x_max, int index_start, int m,
                    float particle_min_weight) {
  return new Layer(x_min, x_max, index_start, m, particle_min_weight);
}
void ref_layer_free(void *l) { delete static_cast<Layer *>(l); }

void ref_create_particles(void *l, float x_ini, float wmc, int n,
                          unsigned long long seed) {
  static_cast<Layer *>(l)->create_particles(x_ini, wmc, n, seed);
}
void ref_simulate(void *l, int nb_particles, int nthread) {
  static_cast<Layer *>(l)->simulate(nb_particles, nthread);
}
int ref_nb_active(void *l) { return static_cast<Layer *>(l)->nb_active(); }
void ref_dump_WA(void *l) { static_cast<Layer *>(l)->dump_WA(); }

float ref_dx(void *l) { return static_cast<Layer *>(l)->dx; }
float ref_x_min(void *l) { return static_cast<Layer *>(l)->x_min; }
float ref_x_max(void *l) { return static_cast<Layer *>(l)->x_max; }
int ref_left_border(void *l) { return static_cast<Layer *>(l)->left_border; }
int ref_right_border(void *l) { return static_cast<Layer *>(l)->right_border; }
int ref_nb_disabled(void *l) { return static_cast<Layer *>(l)->nb_disabled; }
int ref_nb_particles_create(void *l) {
  return static_cast<Layer *>(l)->nb_particles_create;
}
int ref_m(void *l) { return (int)static_cast<Layer *>(l)->weights_absorbed.size(); }
float *ref_sigs(void *l) { return static_cast<Layer *>(l)->sigs.data(); }
float *ref_absorption_rates(void *l) {
  return static_cast<Layer *>(l)->absorption_rates.data();
}
float *ref_weights_absorbed(void *l) {
  return static_cast<Layer *>(l)->weights_absorbed.data();
}
int ref_particles_size(void *l) {
  return (int)static_cast<Layer *>(l)->particles.size();
}
Particle *ref_particles(void *l) {
  return static_cast<Layer *>(l)->particles.data();
}
int ref_particles_left_size(void *l) {
  return (int)static_cast<Layer *>(l)->particles_left.size();
}
Particle *ref_particles_left(void *l) {
  return static_cast<Layer *>(l)->particles_left.data();
}
int ref_particles_right_size(void *l) {
  return (int)static_cast<Layer *>(l)->particles_right.size();
}
Particle *ref_particles_right(void *l) {
  return static_cast<Layer *>(l)->particles_right.data();
}
void ref_clear_left(void *l) { static_cast<Layer *>(l)->particles_left.clear(); }
void ref_clear_right(void *l) { static_cast<Layer *>(l)->particles_right.clear(); }
void ref_push(void *l, const Particle *p, int n) {
  Layer *L = static_cast<Layer *>(l);
  L->particles.insert(L->particles.end(), p, p + n);
}

/* one event / one history through the reference's own public methods, with a
 * caller-supplied tally array of m floats (what simulate() passes them) */
int ref_particle_step(void *l, Particle *p, float *tally_m) {
  Layer *L = static_cast<Layer *>(l);
  std::vector<real_t> w(tally_m, tally_m + L->weights_absorbed.size());
  int r = L->particle_step(*p, w);
  std::memcpy(tally_m, w.data(), w.size() * sizeof(float));
  return r;
}

float ref_rnd_real(unsigned long long *seed) { return rnd_real(seed); }
unsigned long long ref_rnd_seed(unsigned long long *seed) { return rnd_seed(seed); }
int ref_sizeof_particle(void) { return (int)sizeof(Particle); }

} /* extern "C" */
