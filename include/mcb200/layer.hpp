// mcb200/layer.hpp -- C++ facade over the C ABI (mcb200.h) with the public
// surface of the reference's Layer (include/layer/layer.hpp:14-114), so that
// code written against the reference -- worker_sync / worker_async /
// worker_rma (src/worker*.cpp), src/test_layer.cpp, src/test_layer_perf.cpp --
// compiles unchanged with -Iinclude/mcb200/compat and links libmcb200.so.
//
// "Staging mode": the host vectors `particles`, `particles_left`,
// `particles_right` stay the interface the workers send from / receive into;
// simulate() pushes the particles it consumes to the GPU, tracks them there
// and pops the escapees back.  Source particles are born on the device and
// never cross PCIe.  There is no CPU tracking path: simulate() ALWAYS runs on
// the GPU (`nthread` and `use_gpu` are accepted and ignored), and the
// per-particle CPU members simulate_particle / particle_step are not provided.
// Errors follow the reference convention: message on stderr + exit
// (src/layer.cpp:276-277, include/gpu_errcheck/gpu_errcheck.hpp:10-18).
#ifndef MCB200_LAYER_HPP
#define MCB200_LAYER_HPP

#include <limits>
#include <vector>

#include "../mcb200.h"

// include/types/types.hpp:6-16
typedef unsigned long long seed_t;
typedef float real_t;
#ifndef MCMPI_REAL_T
#define MCMPI_REAL_T MPI_FLOAT
#endif
constexpr real_t MAXREAL = std::numeric_limits<real_t>::max();
#ifndef EPS_PRECISION
#define EPS_PRECISION 1e-4F
#endif

// include/types/particle.hpp:7-18 (same tag, same layout: it is the MPI wire format)
typedef struct particle_tag {
public:
  seed_t seed;
  real_t x;
  real_t mu;
  real_t wmc;
  int index;
} Particle;
static_assert(sizeof(Particle) == 24 && sizeof(Particle) == sizeof(mcb200_particle),
              "Particle must stay the 24-byte wire format");

class Layer {
public:
  // include/layer/layer.hpp:26
  Layer(real_t x_min, real_t x_max, int index_start, int m, real_t particle_min_weight);
  Layer(const Layer &other);
  Layer(Layer &&other) noexcept;
  ~Layer();
  Layer &operator=(const Layer &) = delete;  // const members, as in the reference

  // include/layer/layer.hpp:40 -- the particles are born on the device
  void create_particles(real_t x_ini, real_t wmc, int n, seed_t seed);
  // include/layer/layer.hpp:56 -- always the GPU path
  void simulate(int nb_particles, int nthread = -1, bool use_gpu = true);
  // include/layer/layer.hpp:65
  void dump_WA();
  // include/layer/layer.hpp:74
  int nb_active() const;

  // -- Data -- (include/layer/layer.hpp:85-109)
  const real_t x_min, x_max;

private:
  const int m;
  const int index_start;

public:
  std::vector<real_t> weights_absorbed;
  std::vector<Particle> particles;
  std::vector<Particle> particles_left;
  std::vector<Particle> particles_right;
  const real_t dx;
  int nb_disabled = 0;

  const bool left_border, right_border;

public:
  // -- physical properties -- public and mutable like the reference's; a
  // change is picked up by the next simulate()
  std::vector<real_t> sigs;
  std::vector<real_t> absorption_rates;
  const real_t particle_min_weight;

  seed_t seed = 0;
  real_t x_ini = 0, wmc = 0;
  int nb_particles_create = 0;

  // -- extensions (not in the reference) --
  // cell width used for the edges; decompose_domain_global_dx passes the ONE
  // global dx so that K layers reproduce the single-layer trajectories
  void set_edge_dx(real_t w);
  mcb200_layer *handle();           // the C-ABI object (created on first use)
  mcb200_counts counts();           // events, scatters, escape weights, timings
  std::vector<double> weights_absorbed_f64();

private:
  void ensure_device();
  void sync_cross_sections();
  mcb200_layer *h_ = nullptr;
  real_t edge_dx_ = 0;
  std::vector<real_t> sigs_uploaded_, abs_uploaded_;
};

// include/layer/layer.hpp:112-114, src/layer.cpp:17-42
Layer decompose_domain(real_t x_min, real_t x_max, real_t x_ini, int world_size, int world_rank,
                       int nb_cells, int nb_particles, real_t particle_min_weight);
// same decomposition, but every layer tracks with the global cell width
// (SURVEY hard part 3): N layers == 1 layer bit for bit
Layer decompose_domain_global_dx(real_t x_min, real_t x_max, real_t x_ini, int world_size,
                                 int world_rank, int nb_cells, int nb_particles,
                                 real_t particle_min_weight);

#endif  // MCB200_LAYER_HPP
