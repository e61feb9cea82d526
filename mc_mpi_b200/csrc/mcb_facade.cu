// mcb_facade.cu -- the reference-shaped C++ surface above the C ABI:
//   * class Layer          (include/mcb200/layer.hpp  <-> include/layer/layer.hpp)
//   * decompose_domain     (src/layer.cpp:17-42)
//   * cusimulate           (include/mcb200/culayer.hpp <-> include/culayer/culayer.hpp)
// Host-only code; the physics stays in the kernels behind mcb200.h.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <mutex>
#include <utility>
#include <vector>

#include "../../include/mcb200/culayer.hpp"
#include "../../include/mcb200/layer.hpp"

namespace {

// the reference's error convention: message + exit (src/layer.cpp:276-277,
// include/gpu_errcheck/gpu_errcheck.hpp:10-18)
void die_on(int rc, const char *where) {
  if (rc == MCB200_OK) return;
  std::fprintf(stderr, "mcb200: %s failed (%d): %s\n", where, rc, mcb200_last_error());
  std::exit(EXIT_FAILURE);
}

int default_device() {
  const char *e = std::getenv("MCB200_DEVICE");
  if (e && *e) return std::atoi(e);
  e = std::getenv("LOCAL_RANK");  // one process per GPU under mpirun / torchrun launchers
  if (e && *e) {
    const int n = mcb200_device_count();
    if (n > 0) return std::atoi(e) % n;
  }
  return 0;
}

}  // namespace

// src/layer.cpp:44-49
Layer::Layer(real_t x_min_, real_t x_max_, int index_start_, int m_, real_t particle_min_weight_)
    : x_min(x_min_), x_max(x_max_), m(m_), index_start(index_start_),
      dx((x_max_ - x_min_) / m_), left_border(std::fabs((double)x_min_) < EPS_PRECISION),
      right_border(std::fabs((double)x_max_ - 1.0) < EPS_PRECISION),
      particle_min_weight(particle_min_weight_) {
  // cross-sections as the reference hard-codes them, src/layer.cpp:53-63
  sigs.reserve((size_t)m);
  for (int i = 0; i < m; ++i) {
    real_t x_mid = x_min + (i * dx) + 0.5 * dx;
    sigs.push_back(expf(-x_mid));
  }
  absorption_rates = std::vector<real_t>((size_t)m, 0.5);
  weights_absorbed = std::vector<real_t>((size_t)m, 0.0);
  particles.reserve(10000);
}

Layer::Layer(const Layer &o)
    : x_min(o.x_min), x_max(o.x_max), m(o.m), index_start(o.index_start),
      weights_absorbed(o.weights_absorbed), particles(o.particles),
      particles_left(o.particles_left), particles_right(o.particles_right), dx(o.dx),
      nb_disabled(o.nb_disabled), left_border(o.left_border), right_border(o.right_border),
      sigs(o.sigs), absorption_rates(o.absorption_rates),
      particle_min_weight(o.particle_min_weight), seed(o.seed), x_ini(o.x_ini), wmc(o.wmc),
      nb_particles_create(o.nb_particles_create), edge_dx_(o.edge_dx_),
      sigs_uploaded_(o.sigs_uploaded_), abs_uploaded_(o.abs_uploaded_) {
  if (o.h_) die_on(mcb200_layer_clone(o.h_, &h_), "Layer copy");
}

Layer::Layer(Layer &&o) noexcept
    : x_min(o.x_min), x_max(o.x_max), m(o.m), index_start(o.index_start),
      weights_absorbed(std::move(o.weights_absorbed)), particles(std::move(o.particles)),
      particles_left(std::move(o.particles_left)), particles_right(std::move(o.particles_right)),
      dx(o.dx), nb_disabled(o.nb_disabled), left_border(o.left_border),
      right_border(o.right_border), sigs(std::move(o.sigs)),
      absorption_rates(std::move(o.absorption_rates)),
      particle_min_weight(o.particle_min_weight), seed(o.seed), x_ini(o.x_ini), wmc(o.wmc),
      nb_particles_create(o.nb_particles_create), h_(o.h_),
      edge_dx_(o.edge_dx_), sigs_uploaded_(std::move(o.sigs_uploaded_)),
      abs_uploaded_(std::move(o.abs_uploaded_)) {
  o.h_ = nullptr;
}

Layer::~Layer() {
  if (h_) mcb200_layer_destroy(h_);
}

void Layer::set_edge_dx(real_t w) {
  if (h_) {
    std::fprintf(stderr, "mcb200: set_edge_dx after the device layer exists\n");
    std::exit(EXIT_FAILURE);
  }
  edge_dx_ = w;
}

void Layer::ensure_device() {
  if (h_) return;
  mcb200_layer_desc d;
  std::memset(&d, 0, sizeof d);
  d.abi_version = MCB200_ABI_VERSION;
  d.device = default_device();
  d.x_min = x_min;
  d.x_max = x_max;
  d.index_start = index_start;
  d.m = m;
  d.dx = edge_dx_ > 0 ? edge_dx_ : dx;
  d.particle_min_weight = particle_min_weight;
  d.left_border = left_border;
  d.right_border = right_border;
  d.sigs = sigs.data();
  d.absorption_rates = absorption_rates.data();
  d.keep_border = 0;
  die_on(mcb200_layer_create(&d, &h_), "Layer (device side)");
  sigs_uploaded_ = sigs;
  abs_uploaded_ = absorption_rates;
  if (nb_particles_create > 0)
    die_on(mcb200_layer_create_particles(h_, x_ini, wmc, nb_particles_create, seed),
           "create_particles");
}

void Layer::sync_cross_sections() {
  if (sigs != sigs_uploaded_ || absorption_rates != abs_uploaded_) {
    if ((int)sigs.size() != m || (int)absorption_rates.size() != m) {
      std::fprintf(stderr, "mcb200: sigs / absorption_rates must keep %d entries\n", m);
      std::exit(EXIT_FAILURE);
    }
    die_on(mcb200_layer_set_cross_sections(h_, sigs.data(), absorption_rates.data()),
           "set_cross_sections");
    sigs_uploaded_ = sigs;
    abs_uploaded_ = absorption_rates;
  }
}

// src/layer.cpp:71-82
void Layer::create_particles(real_t x_ini_, real_t wmc_, int n, seed_t seed_) {
  if (x_ini_ > x_min && x_ini_ < x_max) {
    this->x_ini = x_ini_;
    this->wmc = wmc_;
    this->nb_particles_create = n;
    this->seed = seed_;
    if (h_) die_on(mcb200_layer_create_particles(h_, x_ini, wmc, n, seed), "create_particles");
  }
}

// src/layer.cpp:84-87
int Layer::nb_active() const { return (int)particles.size() + nb_particles_create; }

mcb200_layer *Layer::handle() {
  ensure_device();
  return h_;
}

mcb200_counts Layer::counts() {
  ensure_device();
  mcb200_counts c;
  die_on(mcb200_layer_counts(h_, &c), "counts");
  return c;
}

std::vector<double> Layer::weights_absorbed_f64() {
  ensure_device();
  std::vector<double> w((size_t)m);
  die_on(mcb200_layer_weights_absorbed_f64(h_, w.data()), "weights_absorbed_f64");
  return w;
}

// src/layer.cpp:239-361.  Which particles one call consumes follows the
// reference: births top the bank up only when it holds fewer than asked
// (:244-245), and the bank is consumed from the back (:319).
void Layer::simulate(int nb_particles, int /*nthread*/, bool /*use_gpu*/) {
  ensure_device();
  sync_cross_sections();
  const long long avail = (long long)particles.size() + nb_particles_create;
  long long want = nb_particles < 0 ? avail : nb_particles;
  if (want > avail) want = avail;
  if (want <= 0) return;
  long long n_birth = 0;
  if ((long long)particles.size() < want) {
    n_birth = nb_particles_create < want ? nb_particles_create : want;
  }
  const long long n_host = want - n_birth;
  mcb200_counts c;
  if (n_host > 0) {
    // the tail of `particles` (src/layer.cpp:319), copied in chunks under the tracking
    const Particle *tail = particles.data() + (particles.size() - (size_t)n_host);
    die_on(mcb200_layer_simulate_host(h_, reinterpret_cast<const mcb200_particle *>(tail), n_host, &c),
           "simulate_host");
    particles.resize(particles.size() - (size_t)n_host);
  }
  if (n_birth > 0 || n_host == 0) die_on(mcb200_layer_simulate(h_, n_birth, &c), "simulate");
  // the reference's counters are `int` (include/layer/layer.hpp:97,109): refuse to wrap
  if (c.n_unborn > INT_MAX || c.nb_disabled > INT_MAX) {
    std::fprintf(stderr, "Layer: more than INT_MAX particles in nb_disabled / nb_particles_create; "
                         "use the C ABI (64-bit counters) for runs of this size\n");
    std::exit(EXIT_FAILURE);
  }
  nb_particles_create = (int)c.n_unborn;
  nb_disabled = (int)c.nb_disabled;

  // escapees back into the vectors the workers send from (:332-346); the
  // device already absorbed the ones crossing a global border (:350-360)
  std::vector<Particle> *dst[2] = {&particles_left, &particles_right};
  const int64_t n_out[2] = {c.n_outbox_left, c.n_outbox_right};
  for (int s = 0; s < 2; ++s) {
    if (n_out[s] <= 0) continue;
    const size_t old = dst[s]->size();
    dst[s]->resize(old + (size_t)n_out[s]);
    int64_t got = 0;
    mcb200_particle *p = reinterpret_cast<mcb200_particle *>(dst[s]->data() + old);
    die_on(s == 0 ? mcb200_layer_pop_left(h_, p, n_out[s], &got)
                  : mcb200_layer_pop_right(h_, p, n_out[s], &got),
           "pop");
    dst[s]->resize(old + (size_t)got);
  }
  die_on(mcb200_layer_weights_absorbed(h_, weights_absorbed.data()), "weights_absorbed");
}

// src/layer.cpp:363-380
void Layer::dump_WA() {
  ensure_device();
  if (mcb200_layer_dump_WA(h_, nullptr) != MCB200_OK) {
    std::fprintf(stderr, "Couldn't open file WA.out for writing.\n");
    std::exit(1);
  }
}

static Layer decompose_impl(real_t x_min, real_t x_max, real_t x_ini, int world_size,
                            int world_rank, int nb_cells, int nb_particles,
                            real_t particle_min_weight, bool global_dx) {
  // src/layer.cpp:24-33
  int cells_per_layer = nb_cells / world_size;
  int num_with_extra = nb_cells % world_size;
  int nb_my_cells = cells_per_layer + (world_rank < num_with_extra);
  int start_index =
      world_rank * cells_per_layer + (world_rank < num_with_extra ? world_rank : num_with_extra);
  real_t dx = (x_max - x_min) / ((float)nb_cells);
  int cell_ini = (int)((x_ini - x_min) / dx);

  Layer layer(x_min + start_index * dx, x_min + (start_index + nb_my_cells) * dx, start_index,
              nb_my_cells, particle_min_weight);
  if (global_dx) {
    // one dx and one cross-section table for the whole slab, sliced per layer, so that K
    // layers reproduce the single-layer trajectories and tallies bit for bit
    layer.set_edge_dx(dx);
    std::vector<real_t> s((size_t)nb_cells), a((size_t)nb_cells);
    die_on(mcb200_default_cross_sections(x_min, x_max, nb_cells, s.data(), a.data()),
           "default_cross_sections");
    layer.sigs.assign(s.begin() + start_index, s.begin() + start_index + nb_my_cells);
    layer.absorption_rates.assign(a.begin() + start_index,
                                  a.begin() + start_index + nb_my_cells);
  }
  if ((cell_ini >= start_index) && (cell_ini < start_index + nb_my_cells)) {
    seed_t seed = 5127801;  // :36
    layer.create_particles(x_ini, 1.0 / nb_particles, nb_particles, seed);
  }
  return layer;
}

Layer decompose_domain(real_t x_min, real_t x_max, real_t x_ini, int world_size, int world_rank,
                       int nb_cells, int nb_particles, real_t particle_min_weight) {
  return decompose_impl(x_min, x_max, x_ini, world_size, world_rank, nb_cells, nb_particles,
                        particle_min_weight, false);
}

Layer decompose_domain_global_dx(real_t x_min, real_t x_max, real_t x_ini, int world_size,
                                 int world_rank, int nb_cells, int nb_particles,
                                 real_t particle_min_weight) {
  return decompose_impl(x_min, x_max, x_ini, world_size, world_rank, nb_cells, nb_particles,
                        particle_min_weight, true);
}

// include/culayer/culayer.hpp:6-13, src/culayer.cu:41-92
void cusimulate(int n, Particle *particles, float const *const sigs,
                float const *const absorption_rates, float *const weights_absorbed, int min_index,
                int max_index, float dx) {
  if (n <= 0) return;
  const int n_cells = max_index - min_index;
  mcb200_layer_desc d;
  std::memset(&d, 0, sizeof d);
  d.abi_version = MCB200_ABI_VERSION;
  d.device = default_device();
  d.x_min = min_index * dx;
  d.x_max = max_index * dx;
  d.index_start = min_index;
  d.m = n_cells;
  d.dx = dx;
  d.particle_min_weight = 0.0f;  // the reference kernel has no cut-off (culayer_kernel.cu:53)
  d.left_border = 0;             // every particle comes back to the caller,
  d.right_border = 0;            // who does the border bookkeeping (layer.cpp:264-298)
  d.sigs = sigs;
  d.absorption_rates = absorption_rates;
  // one device layer is kept between calls (stream, events, tally, outboxes: nine
  // allocations) and only rebuilt when the geometry changes; the reference allocates and frees
  // its buffers per call (src/culayer.cu:57-92)
  static std::mutex mu;
  static mcb200_layer *cached = nullptr;
  static mcb200_layer_desc cached_desc;
  std::lock_guard<std::mutex> lock(mu);
  const bool same = cached && cached_desc.device == d.device && cached_desc.index_start == d.index_start &&
                    cached_desc.m == d.m && cached_desc.dx == d.dx;
  if (!same) {
    if (cached) mcb200_layer_destroy(cached);
    cached = nullptr;
    die_on(mcb200_layer_create(&d, &cached), "cusimulate: layer");
    cached_desc = d;
  } else {
    die_on(mcb200_layer_set_cross_sections(cached, sigs, absorption_rates), "cusimulate: tables");
    die_on(mcb200_layer_reset_tally(cached), "cusimulate: reset");
  }
  mcb200_layer *h = cached;
  mcb200_counts c;
  die_on(mcb200_layer_simulate_host(h, reinterpret_cast<const mcb200_particle *>(particles), n, &c),
         "cusimulate: simulate");
  int64_t nl = 0, nr = 0;
  mcb200_particle *out = reinterpret_cast<mcb200_particle *>(particles);
  die_on(mcb200_layer_pop_left(h, out, n, &nl), "cusimulate: pop_left");
  die_on(mcb200_layer_pop_right(h, out + nl, n - nl, &nr), "cusimulate: pop_right");
  if (nl + nr != n) {
    // with no weight cut-off nothing can die; mirrors src/layer.cpp:276-277
    std::fprintf(stderr, "There was a particle which was not disabled nor transported");
    std::exit(EXIT_FAILURE);
  }
  std::vector<float> w((size_t)n_cells);
  die_on(mcb200_layer_weights_absorbed(h, w.data()), "cusimulate: tally");
  for (int j = 0; j < n_cells; ++j) weights_absorbed[j] += w[(size_t)j];
}
