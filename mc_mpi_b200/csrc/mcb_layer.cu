// mcb_layer.cu -- host side of one layer (= one sub-slab on one GPU) and the
// C ABI of include/mcb200.h.  Mirrors the state machine of the reference's
// Layer (include/layer/layer.hpp, src/layer.cpp) with the bank, the outboxes
// and the tally resident in HBM:
//
//   bank      seed[cap] (u64) + st[cap] (float4 {x, mu, wmc, bits(index)})
//   outboxes  contiguous 24-byte wire records per side (packed from the tracking kernel's
//             per-CTA stripes by gather_stripes after every launch)
//   tally     u64[2][m+3] exact 128-bit long accumulators (mcb_kernels.cuh): low halves,
//             then high halves; the CTA-private copies in shared memory use 32-bit digits
//   xs        float4[m] {sig_a, sig_i, ~1/sig_i, -}
//
// Nothing here computes physics on the CPU: there is no fallback path.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

#include "../../include/mcb200.h"
#include "mcb_host.hpp"
#include "mcb_kernels.cuh"

namespace mcb {

static thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}

CellXs host_cell_xs(float sig, float a) {
  const float interaction_rate = (float)(1.0 - (double)a);
  const float sig_a = sig * a;
  const float sig_i = sig * interaction_rate;
  // a float safely below 1/sig_i for the kernel's "certain crossing" test: the whole margin of
  // the test (2^-18, against <= 2^-22 of accumulated rounding) is taken here, once per cell
  const float inv_lb = sig_i > MCB_EPS
                           ? (float)((1.0 / (double)sig_i) * (1.0 - 0x1p-18 - 0x1p-19))
                           : std::numeric_limits<float>::infinity();
  return make_float4(sig_a, sig_i, inv_lb, 0.0f);
}

double acc_to_double(const unsigned d[kAccDigits]) {
  unsigned __int128 u = 0;
  for (int j = kAccDigits - 1; j >= 0; --j) u = (u << 32) | d[j];
  return std::ldexp((double)(__int128)u, kAccLsbLog2);
}

void acc_halves_to_digits(const unsigned *raw, int ncell, int m, unsigned *out) {
  // 32-bit digit j of cell c is word (j & 1) of half j / 2 (little endian)
  const size_t nc = (size_t)ncell;
  for (int c = 0; c < m; ++c)
    for (int j = 0; j < kAccDigits; ++j)
      out[(size_t)c * kAccDigits + j] = raw[(size_t)(j / 2) * 2 * nc + 2 * (size_t)c + (size_t)(j & 1)];
}

const char *last_error_cstr() { return g_last_error.c_str(); }
void set_last_error(const std::string &m) { g_last_error = m; }
std::string get_last_error() { return g_last_error; }

}  // namespace mcb

namespace {

using mcb::DeviceGuard;
using mcb::acc_to_double;
using mcb::fail;

// scratch device allocation of the known-answer-test entry points: freed on every return path
template <typename T>
struct DevScratch {
  T *p = nullptr;
  ~DevScratch() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t count) { return cudaMalloc(&p, count * sizeof(T)); }
};

// growable pair of device arrays in the bank layout
struct SoaBuf {
  unsigned long long *seed = nullptr;
  float4 *st = nullptr;
  long long cap = 0;
};

}  // namespace

struct mcb200_layer {
  // --- what Layer carries (include/layer/layer.hpp:85-109)
  int device = 0;
  float x_min = 0, x_max = 0;
  int index_start = 0, m = 0;
  float dx = 0;
  float particle_min_weight = 0;
  bool left_border = false, right_border = false;
  bool keep_border = false;
  std::vector<float> sigs, absorption_rates;
  long long nb_disabled = 0;
  // unborn source particles (layer.hpp:107-109)
  unsigned long long chain_state = 0;
  float x_ini = 0, wmc = 0;
  long long n_unborn = 0;
  // --- device state
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  SoaBuf bank;
  long long n_bank = 0;
  unsigned long long *d_out[2] = {nullptr, nullptr};      // settled outboxes, 3 words / record
  long long out_cap[2] = {0, 0};
  long long n_out[2] = {0, 0};
  unsigned long long *d_scratch[2] = {nullptr, nullptr};  // per-launch stripes + overflow
  long long scratch_cap[2] = {0, 0};
  unsigned *d_stripe_n = nullptr;                         // [2][kStripes + 1]
  // direct peer exchange: this layer's inbox (4 slots: from-side x parity) and the
  // neighbours' inboxes this layer's kernel stores into
  unsigned char *d_inbox = nullptr;
  mcb200_inbox_geom inbox_geom{};
  struct Peer {
    unsigned char *base = nullptr;  // the neighbour's inbox block, mapped into this process
    bool ipc = false;               // opened with cudaIpcOpenMemHandle (close on destroy)
    mcb200_inbox_geom geom{};
  } peer[2];
  int parity = 0;
  mcb::CellXs *d_xs = nullptr;
  unsigned *d_acc = nullptr;          // u64[2][m + kAccExtra], see mcb_kernels.cu gacc_add
  mcb::DevCounters *d_ctr = nullptr;
  mcb::DevCounters *h_ctr = nullptr;  // pinned
  unsigned long long *h_cls = nullptr;  // pinned: [2][3] halves of the 3 class accumulators
  void *d_stage = nullptr;            // AoS staging for push / pop
  long long stage_cap = 0;            // in particles
  // pipelined host path (simulate_host): a copy stream and two staging buffers, so that the
  // H2D copy of chunk k+1 runs under the tracking of chunk k
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr};
  void *d_pipe[2] = {nullptr, nullptr};
  long long pipe_cap = 0;             // particles per staging buffer
  long long opt_host_chunk = 1ll << 25;
  bool xs_dirty = true;
  mcb::JumpTable seed_jump;
  // --- knobs / cumulative stats
  int opt_tally_mode = 0, opt_block = 0, opt_bps = 0, opt_retire_batch = 0;
  int opt_rng = 0;                    // 0 = LCG (parity), 1 = Philox2x32-10
  unsigned long long philox_next_id = 0;   // histories handed out so far in Philox mode
  long long opt_birth_chunk = 1ll << 26;
  bool cfg_dirty = true;
  mcb::TrackLaunch cfg{};
  long long events = 0, scatters = 0, n_cls[3] = {0, 0, 0};
  double w_cls[3] = {0, 0, 0};
  long long launches = 0, gpu_launches = 0;
  double track_ms = 0;

  int ncell() const { return m + mcb::kAccExtra; }
  size_t acc_words() const { return (size_t)mcb::kAccDigits * (size_t)ncell(); }
};

namespace {

int soa_reserve(mcb200_layer *l, SoaBuf *b, long long used, long long want) {
  if (want <= b->cap) return MCB200_OK;
  long long cap = b->cap > 0 ? b->cap : 4096;
  while (cap < want) cap += cap / 2 + 1;
  unsigned long long *ns = nullptr;
  float4 *nt = nullptr;
  MCB_CUDA(cudaMalloc(&ns, (size_t)cap * sizeof(unsigned long long)));
  cudaError_t e = cudaMalloc(&nt, (size_t)cap * sizeof(float4));
  if (e != cudaSuccess) {
    cudaFree(ns);
    return fail(MCB200_ERR_NOMEM, std::string("cudaMalloc bank: ") + cudaGetErrorString(e));
  }
  if (used > 0) {
    e = cudaMemcpyAsync(ns, b->seed, (size_t)used * sizeof(unsigned long long),
                        cudaMemcpyDeviceToDevice, l->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(nt, b->st, (size_t)used * sizeof(float4), cudaMemcpyDeviceToDevice,
                          l->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(l->stream);
    if (e != cudaSuccess) {
      cudaFree(ns);
      cudaFree(nt);
      return fail(MCB200_ERR_CUDA, std::string("growing the bank: ") + cudaGetErrorString(e));
    }
  }
  cudaFree(b->seed);
  cudaFree(b->st);
  b->seed = ns;
  b->st = nt;
  b->cap = cap;
  return MCB200_OK;
}

// growable buffer of 24-byte records; the first `used` records survive a growth
int rec_reserve(mcb200_layer *l, unsigned long long **buf, long long *cap, long long used,
                long long want) {
  if (want <= *cap) return MCB200_OK;
  long long nc = *cap > 0 ? *cap : 4096;
  while (nc < want) nc += nc / 2 + 1;
  unsigned long long *nb = nullptr;
  MCB_CUDA(cudaMalloc(&nb, (size_t)nc * sizeof(mcb200_particle)));
  if (used > 0) {
    cudaError_t e = cudaMemcpyAsync(nb, *buf, (size_t)used * sizeof(mcb200_particle),
                                    cudaMemcpyDeviceToDevice, l->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(l->stream);
    if (e != cudaSuccess) {
      cudaFree(nb);
      return fail(MCB200_ERR_CUDA, std::string("growing an outbox: ") + cudaGetErrorString(e));
    }
  }
  cudaFree(*buf);
  *buf = nb;
  *cap = nc;
  return MCB200_OK;
}

int stage_reserve(mcb200_layer *l, long long n) {
  if (n <= l->stage_cap) return MCB200_OK;
  long long cap = l->stage_cap > 0 ? l->stage_cap : 4096;
  while (cap < n) cap += cap / 2 + 1;
  if (l->d_stage) cudaFree(l->d_stage);
  l->d_stage = nullptr;
  l->stage_cap = 0;
  MCB_CUDA(cudaMalloc(&l->d_stage, (size_t)cap * sizeof(mcb200_particle)));
  l->stage_cap = cap;
  return MCB200_OK;
}

// per-cell event constants, src/layer.cpp:131-133 (host code is built with
// -ffp-contract=off and no -mfma, like the reference: plain IEEE float ops)
int upload_xs(mcb200_layer *l) {
  std::vector<mcb::CellXs> xs((size_t)l->m);
  for (int i = 0; i < l->m; ++i)
    xs[(size_t)i] = mcb::host_cell_xs(l->sigs[(size_t)i], l->absorption_rates[(size_t)i]);
  MCB_CUDA(cudaMemcpyAsync(l->d_xs, xs.data(), xs.size() * sizeof(mcb::CellXs),
                           cudaMemcpyHostToDevice, l->stream));
  MCB_CUDA(cudaStreamSynchronize(l->stream));
  l->xs_dirty = false;
  return MCB200_OK;
}

// the m cell accumulators as cell-major digits: out[4*c + j]
int fetch_tally(mcb200_layer *l, std::vector<unsigned> *out) {
  DeviceGuard g(l->device);
  std::vector<unsigned> raw(l->acc_words());
  MCB_CUDA(cudaMemcpyAsync(raw.data(), l->d_acc, raw.size() * sizeof(unsigned),
                           cudaMemcpyDeviceToHost, l->stream));
  MCB_CUDA(cudaStreamSynchronize(l->stream));
  out->resize((size_t)l->m * mcb::kAccDigits);
  mcb::acc_halves_to_digits(raw.data(), l->ncell(), l->m, out->data());
  return MCB200_OK;
}

// Layer::create_particles(int n), src/layer.cpp:89-121, on the device
int birth(mcb200_layer *l, long long n) {
  if (n <= 0) return MCB200_OK;
  int rc = soa_reserve(l, &l->bank, l->n_bank, l->n_bank + n);
  if (rc) return rc;
  const float cell = l->x_ini / l->dx;  // :106, ignores x_min
  const int index = (int)cell;
  if (l->opt_rng == 0) {
    MCB_CUDA(mcb::launch_birth(n, l->chain_state, l->seed_jump, l->x_ini, l->wmc, index,
                               l->bank.seed + l->n_bank, l->bank.st + l->n_bank, l->stream));
    l->chain_state = mcb::jump_state(l->seed_jump, (unsigned long long)n, l->chain_state);
  } else {
    // counter-based: the seed given to create_particles is the KEY, history ids just count up
    MCB_CUDA(mcb::launch_birth_philox(n, l->philox_next_id, mcb::philox_key(l->chain_state), l->x_ini,
                                      l->wmc, index, l->bank.seed + l->n_bank,
                                      l->bank.st + l->n_bank, l->stream));
    l->philox_next_id += (unsigned long long)n;
  }
  l->gpu_launches++;
  l->n_bank += n;
  l->n_unborn -= n;
  return MCB200_OK;
}

int track(mcb200_layer *l, long long take) {
  if (take <= 0) return MCB200_OK;
  if (l->xs_dirty) {
    int rc = upload_xs(l);
    if (rc) return rc;
  }
  if (l->cfg_dirty) {
    MCB_CUDA(mcb::track_configure(l->device, l->m, l->opt_tally_mode, l->opt_block, l->opt_bps,
                                  l->opt_rng, &l->cfg));
    l->cfg_dirty = false;
  }
  const bool write_side[2] = {!l->left_border || l->keep_border,
                              !l->right_border || l->keep_border};
  // per-CTA stripes: twice the fair share each (dynamic work distribution keeps CTAs within
  // a few percent of each other), plus one overflow segment that can take everything
  const int grid = mcb::track_grid(l->cfg, take);
  const int stripe_cap = (int)(2 * ((take + grid - 1) / grid) + 64);
  const long long ovf_base = (long long)grid * stripe_cap;
  for (int s = 0; s < 2; ++s) {
    if (!write_side[s]) continue;
    if (l->peer[s].base) {
      // direct peer exchange: the stripes live in the neighbour's inbox, with ITS geometry
      if (grid > l->peer[s].geom.nstripes || take > l->peer[s].geom.max_take)
        return fail(MCB200_ERR_CAPACITY,
                    "track: this launch (" + std::to_string(grid) + " CTAs, " + std::to_string(take) +
                        " particles) does not fit the neighbour's inbox (" +
                        std::to_string(l->peer[s].geom.nstripes) + " stripes, max_take " +
                        std::to_string(l->peer[s].geom.max_take) +
                        "): inboxes are sized for the default launch shape (4 CTAs of 256 threads "
                        "per SM); do not combine \"block\" / \"blocks_per_sm\" options with a "
                        "connected neighbour");
      continue;
    }
    int rc = rec_reserve(l, &l->d_scratch[s], &l->scratch_cap[s], 0, ovf_base + take);
    if (rc) return rc;
    rc = rec_reserve(l, &l->d_out[s], &l->out_cap[s], l->n_out[s], l->n_out[s] + take);
    if (rc) return rc;
  }
  std::memset(l->h_ctr, 0, sizeof(mcb::DevCounters));
  MCB_CUDA(cudaMemcpyAsync(l->d_ctr, l->h_ctr, sizeof(mcb::DevCounters), cudaMemcpyHostToDevice,
                           l->stream));
  MCB_CUDA(cudaMemsetAsync(l->d_stripe_n, 0, 2 * (mcb::kStripes + 1) * sizeof(unsigned),
                           l->stream));
  mcb::TrackParams p{};
  p.bank_seed = l->bank.seed;
  p.bank_st = l->bank.st;
  p.take_base = l->n_bank - take;  // the LAST `take` of the bank, src/layer.cpp:319
  p.take_count = take;
  p.xs = l->d_xs;
  p.idx_lo = l->index_start;
  p.m = l->m;
  p.dx = l->dx;
  p.minw = l->particle_min_weight;
  p.acc = l->d_acc;
  for (int s = 0; s < 2; ++s) {
    p.write_side[s] = write_side[s] ? 1 : 0;
    if (write_side[s] && l->peer[s].base) {
      // escapees on side s land in the neighbour's inbox slot "from side 1-s", this parity
      const mcb200_inbox_geom &g = l->peer[s].geom;
      unsigned char *slot = l->peer[s].base + (size_t)((1 - s) * 2 + l->parity) * (size_t)g.slot_bytes;
      p.out_rec[s] = reinterpret_cast<unsigned long long *>(slot);
      p.fills[s] = reinterpret_cast<unsigned *>(slot + g.fills_offset);
      p.ovf_slot[s] = g.nstripes;
      p.stripe_cap[s] = g.stripe_cap;
      p.ovf_base[s] = (long long)g.nstripes * g.stripe_cap;
      p.ovf_cap[s] = g.ovf_cap;
    } else {
      p.out_rec[s] = l->d_scratch[s];
      p.fills[s] = l->d_stripe_n + s * (mcb::kStripes + 1);
      p.ovf_slot[s] = grid;
      p.stripe_cap[s] = stripe_cap;
      p.ovf_base[s] = ovf_base;
      p.ovf_cap[s] = take;
    }
  }
  p.ctr = l->d_ctr;
  p.rng_key = mcb::philox_key(l->chain_state);
  // short segments (thin sub-slabs of a multi-GPU run) retire often: batch the bookkeeping
  p.retire_batch = l->opt_retire_batch > 0 ? l->opt_retire_batch : (l->m < 512 ? 4 : 2);
  if (p.retire_batch > 32) p.retire_batch = 32;
  MCB_CUDA(cudaEventRecord(l->ev0, l->stream));
  MCB_CUDA(mcb::launch_track(p, l->cfg, l->stream));
  MCB_CUDA(cudaEventRecord(l->ev1, l->stream));
  for (int s = 0; s < 2; ++s) {
    if (!write_side[s] || l->peer[s].base) continue;  // peer sides: already delivered
    MCB_CUDA(mcb::launch_gather_stripes(l->d_scratch[s],
                                        l->d_stripe_n + s * (mcb::kStripes + 1), grid, stripe_cap,
                                        ovf_base, take, l->d_out[s], l->n_out[s],
                                        &l->d_ctr->out_total[s], l->stream));
    l->gpu_launches++;
  }
  MCB_CUDA(cudaMemcpyAsync(l->h_ctr, l->d_ctr, sizeof(mcb::DevCounters), cudaMemcpyDeviceToHost,
                           l->stream));
  // digits of the three class accumulators (weight carried left / right / by the dead)
  MCB_CUDA(cudaMemcpy2DAsync(l->h_cls, mcb::kAccExtra * sizeof(unsigned long long),
                             reinterpret_cast<unsigned long long *>(l->d_acc) + l->m,
                             (size_t)l->ncell() * sizeof(unsigned long long),
                             mcb::kAccExtra * sizeof(unsigned long long), 2,
                             cudaMemcpyDeviceToHost, l->stream));
  MCB_CUDA(cudaStreamSynchronize(l->stream));
  float ms = 0.f;
  MCB_CUDA(cudaEventElapsedTime(&ms, l->ev0, l->ev1));
  l->track_ms += ms;
  l->launches++;
  l->gpu_launches++;
  if (l->h_ctr->overflow) return fail(MCB200_ERR_CAPACITY, "internal: outbox overflow");
  if (l->h_ctr->acc_range)
    return fail(MCB200_ERR_RANGE, "a particle weight or deposit is outside (-2^7, 2^7)");

  const mcb::DevCounters &c = *l->h_ctr;
  l->n_bank -= take;
  l->events += (long long)c.events;
  l->scatters += (long long)c.scatters;
  for (int k = 0; k < 3; ++k) {
    l->n_cls[k] += (long long)c.n_cls[k];
    unsigned d[mcb::kAccDigits];
    for (int j = 0; j < mcb::kAccDigits; ++j)
      d[j] = (unsigned)(l->h_cls[(j / 2) * mcb::kAccExtra + k] >> (32 * (j & 1)));
    l->w_cls[k] = acc_to_double(d);  // the device accumulators are cumulative
  }
  l->n_out[0] += (long long)c.out_total[0];
  l->n_out[1] += (long long)c.out_total[1];
  // src/layer.cpp:343 (dead) and :350-360 (global borders absorb)
  l->nb_disabled += (long long)c.n_cls[2];
  if (l->left_border) l->nb_disabled += (long long)c.n_cls[0];
  if (l->right_border) l->nb_disabled += (long long)c.n_cls[1];
  return MCB200_OK;
}

int fill_counts(mcb200_layer *l, mcb200_counts *o) {
  if (!o) return MCB200_OK;
  o->nb_disabled = l->nb_disabled;
  o->nb_active = l->n_bank + l->n_unborn;
  o->n_bank = l->n_bank;
  o->n_unborn = l->n_unborn;
  o->n_outbox_left = l->n_out[0];
  o->n_outbox_right = l->n_out[1];
  o->events = l->events;
  o->scatters = l->scatters;
  o->n_left = l->n_cls[0];
  o->n_right = l->n_cls[1];
  o->n_dead = l->n_cls[2];
  o->w_left = l->w_cls[0];
  o->w_right = l->w_cls[1];
  o->w_dead = l->w_cls[2];
  o->launches = l->launches;
  o->track_ms = l->track_ms;
  o->gpu_launches = l->gpu_launches;
  return MCB200_OK;
}

int pop_side(mcb200_layer *l, int side, void *dst, bool dst_is_device, long long cap,
             long long *n_out) {
  if (!l || !n_out || (cap > 0 && !dst)) return fail(MCB200_ERR_INVALID, "pop: null argument");
  DeviceGuard g(l->device);
  const long long n = l->n_out[side];
  if (n > cap) {
    *n_out = n;
    return fail(MCB200_ERR_CAPACITY, "pop: caller buffer too small");
  }
  if (n > 0) {  // the outbox already holds wire records: one copy
    MCB_CUDA(cudaMemcpyAsync(dst, l->d_out[side], (size_t)n * sizeof(mcb200_particle),
                             dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                             l->stream));
    MCB_CUDA(cudaStreamSynchronize(l->stream));
  }
  l->n_out[side] = 0;
  *n_out = n;
  return MCB200_OK;
}

int push_any(mcb200_layer *l, const void *src, bool src_is_device, long long n) {
  if (!l || n < 0 || (n > 0 && !src)) return fail(MCB200_ERR_INVALID, "push: bad argument");
  if (n == 0) return MCB200_OK;
  DeviceGuard g(l->device);
  int rc = soa_reserve(l, &l->bank, l->n_bank, l->n_bank + n);
  if (rc) return rc;
  const void *aos = src;
  if (!src_is_device) {
    rc = stage_reserve(l, n);
    if (rc) return rc;
    MCB_CUDA(cudaMemcpyAsync(l->d_stage, src, (size_t)n * sizeof(mcb200_particle),
                             cudaMemcpyHostToDevice, l->stream));
    aos = l->d_stage;
  }
  MCB_CUDA(mcb::launch_aos_to_soa(n, aos, l->bank.seed + l->n_bank, l->bank.st + l->n_bank,
                                  l->stream));
  l->gpu_launches++;
  // a host source (and the staging buffer) may be reused as soon as we return; a DEVICE
  // source only has to stay untouched until the next blocking call on this layer
  // (simulate / pop / counts), which lets an exchange queue both neighbours' records
  if (!src_is_device) MCB_CUDA(cudaStreamSynchronize(l->stream));
  l->n_bank += n;
  return MCB200_OK;
}

}  // namespace

extern "C" {

const char *mcb200_last_error(void) { return mcb::last_error_cstr(); }
int mcb200_abi_version(void) { return MCB200_ABI_VERSION; }
int mcb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int mcb200_default_cross_sections(float x_min, float x_max, int32_t m, float *sigs_out,
                                  float *absorption_rates_out) {
  if (m <= 0) return fail(MCB200_ERR_INVALID, "default_cross_sections: m must be positive");
  const float ctor_dx = (x_max - x_min) / m;  // Layer::dx, src/layer.cpp:47
  for (int i = 0; i < m; ++i) {
    const float base = x_min + (i * ctor_dx);
    const float x_mid = (float)((double)base + 0.5 * (double)ctor_dx);  // :58
    if (sigs_out) sigs_out[i] = expf(-x_mid);                            // :59
    if (absorption_rates_out) absorption_rates_out[i] = 0.5f;            // :63
  }
  return MCB200_OK;
}

int mcb200_layer_create(const mcb200_layer_desc *d, mcb200_layer **out) {
  if (!d || !out) return fail(MCB200_ERR_INVALID, "create: null argument");
  *out = nullptr;
  if (d->abi_version != MCB200_ABI_VERSION)
    return fail(MCB200_ERR_INVALID, "create: abi_version mismatch");
  if (d->m <= 0) return fail(MCB200_ERR_INVALID, "create: m must be positive");
  if (!(d->x_max > d->x_min) || !std::isfinite(d->x_min) || !std::isfinite(d->x_max))
    return fail(MCB200_ERR_INVALID, "create: need finite x_min < x_max");
  int ndev = 0;
  MCB_CUDA(cudaGetDeviceCount(&ndev));
  if (d->device < 0 || d->device >= ndev)
    return fail(MCB200_ERR_INVALID, "create: no such CUDA device");

  mcb200_layer *l = new (std::nothrow) mcb200_layer();
  if (!l) return fail(MCB200_ERR_NOMEM, "create: out of host memory");
  l->device = d->device;
  l->x_min = d->x_min;
  l->x_max = d->x_max;
  l->index_start = d->index_start;
  l->m = d->m;
  const float ctor_dx = (d->x_max - d->x_min) / d->m;  // Layer::dx, src/layer.cpp:47
  l->dx = d->dx > 0.0f ? d->dx : ctor_dx;
  l->particle_min_weight = d->particle_min_weight;
  l->left_border = d->left_border < 0 ? std::fabs((double)d->x_min) < (double)MCB_EPS
                                      : d->left_border != 0;
  l->right_border = d->right_border < 0 ? std::fabs((double)d->x_max - 1.0) < (double)MCB_EPS
                                        : d->right_border != 0;
  l->keep_border = d->keep_border != 0;
  l->sigs.resize((size_t)l->m);
  l->absorption_rates.resize((size_t)l->m);
  // cross-sections default to the reference's hard-coded ones (src/layer.cpp:53-63)
  mcb200_default_cross_sections(d->x_min, d->x_max, d->m, l->sigs.data(),
                                l->absorption_rates.data());
  if (d->sigs) l->sigs.assign(d->sigs, d->sigs + l->m);
  if (d->absorption_rates)
    l->absorption_rates.assign(d->absorption_rates, d->absorption_rates + l->m);
  l->seed_jump = mcb::make_jump_table(mcb::kSeedG, mcb::kSeedC);

  DeviceGuard g(l->device);
  int rc = MCB200_OK;
  auto cuda_ok = [&](cudaError_t e, const char *what) {
    if (e != cudaSuccess && rc == MCB200_OK)
      rc = fail(e == cudaErrorMemoryAllocation ? MCB200_ERR_NOMEM : MCB200_ERR_CUDA,
                std::string(what) + ": " + cudaGetErrorString(e));
    return e == cudaSuccess;
  };
  cuda_ok(cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking), "cudaStreamCreate");
  cuda_ok(cudaEventCreate(&l->ev0), "cudaEventCreate");
  cuda_ok(cudaEventCreate(&l->ev1), "cudaEventCreate");
  cuda_ok(cudaMalloc(&l->d_xs, (size_t)l->m * sizeof(mcb::CellXs)), "cudaMalloc xs");
  cuda_ok(cudaMalloc(&l->d_acc, l->acc_words() * sizeof(unsigned)), "cudaMalloc tally");
  cuda_ok(cudaMalloc(&l->d_ctr, sizeof(mcb::DevCounters)), "cudaMalloc counters");
  cuda_ok(cudaMalloc(&l->d_stripe_n, 2 * (mcb::kStripes + 1) * sizeof(unsigned)),
          "cudaMalloc stripe fills");
  cuda_ok(cudaMallocHost(&l->h_ctr, sizeof(mcb::DevCounters)), "cudaMallocHost counters");
  cuda_ok(cudaMallocHost(&l->h_cls, 2 * mcb::kAccExtra * sizeof(unsigned long long)),
          "cudaMallocHost class weights");
  if (rc == MCB200_OK)
    cuda_ok(cudaMemsetAsync(l->d_acc, 0, l->acc_words() * sizeof(unsigned), l->stream),
            "cudaMemset tally");
  if (rc == MCB200_OK) cuda_ok(cudaStreamSynchronize(l->stream), "cudaStreamSynchronize");
  if (rc != MCB200_OK) {
    std::string keep = mcb::get_last_error();
    mcb200_layer_destroy(l);
    mcb::set_last_error(keep);
    return rc;
  }
  *out = l;
  return MCB200_OK;
}

void mcb200_layer_destroy(mcb200_layer *l) {
  if (!l) return;
  {
    DeviceGuard g(l->device);
    if (l->stream) cudaStreamSynchronize(l->stream);
    cudaFree(l->bank.seed);
    cudaFree(l->bank.st);
    for (int s = 0; s < 2; ++s) {
      cudaFree(l->d_out[s]);
      cudaFree(l->d_scratch[s]);
    }
    cudaFree(l->d_stripe_n);
    for (int s = 0; s < 2; ++s)
      if (l->peer[s].base && l->peer[s].ipc) cudaIpcCloseMemHandle(l->peer[s].base);
    cudaFree(l->d_inbox);
    cudaFree(l->d_xs);
    cudaFree(l->d_acc);
    cudaFree(l->d_ctr);
    cudaFree(l->d_stage);
    for (int b = 0; b < 2; ++b) {
      cudaFree(l->d_pipe[b]);
      if (l->ev_copied[b]) cudaEventDestroy(l->ev_copied[b]);
    }
    if (l->copy_stream) cudaStreamDestroy(l->copy_stream);
    if (l->h_ctr) cudaFreeHost(l->h_ctr);
    if (l->h_cls) cudaFreeHost(l->h_cls);
    if (l->ev0) cudaEventDestroy(l->ev0);
    if (l->ev1) cudaEventDestroy(l->ev1);
    if (l->stream) cudaStreamDestroy(l->stream);
    cudaGetLastError();
  }
  delete l;
}

int mcb200_layer_clone(mcb200_layer *src, mcb200_layer **out) {
  if (!src || !out) return fail(MCB200_ERR_INVALID, "clone: null argument");
  *out = nullptr;
  mcb200_layer_desc d{};
  d.abi_version = MCB200_ABI_VERSION;
  d.device = src->device;
  d.x_min = src->x_min;
  d.x_max = src->x_max;
  d.index_start = src->index_start;
  d.m = src->m;
  d.dx = src->dx;
  d.particle_min_weight = src->particle_min_weight;
  d.left_border = src->left_border;
  d.right_border = src->right_border;
  d.sigs = src->sigs.data();
  d.absorption_rates = src->absorption_rates.data();
  d.keep_border = src->keep_border;
  mcb200_layer *l = nullptr;
  int rc = mcb200_layer_create(&d, &l);
  if (rc) return rc;
  DeviceGuard g(l->device);
  auto bail = [&](int code) {
    std::string keep = mcb::get_last_error();
    mcb200_layer_destroy(l);
    mcb::set_last_error(keep);
    return code;
  };
  cudaStreamSynchronize(src->stream);
  if (src->n_bank > 0) {
    rc = soa_reserve(l, &l->bank, 0, src->n_bank);
    if (rc) return bail(rc);
    if (cudaMemcpyAsync(l->bank.seed, src->bank.seed, (size_t)src->n_bank * 8,
                        cudaMemcpyDeviceToDevice, l->stream) != cudaSuccess ||
        cudaMemcpyAsync(l->bank.st, src->bank.st, (size_t)src->n_bank * 16,
                        cudaMemcpyDeviceToDevice, l->stream) != cudaSuccess)
      return bail(fail(MCB200_ERR_CUDA, "clone: device copy failed"));
  }
  for (int k = 0; k < 2; ++k) {
    if (src->n_out[k] <= 0) continue;
    rc = rec_reserve(l, &l->d_out[k], &l->out_cap[k], 0, src->n_out[k]);
    if (rc) return bail(rc);
    if (cudaMemcpyAsync(l->d_out[k], src->d_out[k], (size_t)src->n_out[k] * sizeof(mcb200_particle),
                        cudaMemcpyDeviceToDevice, l->stream) != cudaSuccess)
      return bail(fail(MCB200_ERR_CUDA, "clone: device copy failed"));
  }
  if (cudaMemcpyAsync(l->d_acc, src->d_acc, l->acc_words() * sizeof(unsigned),
                      cudaMemcpyDeviceToDevice, l->stream) != cudaSuccess ||
      cudaStreamSynchronize(l->stream) != cudaSuccess)
    return bail(fail(MCB200_ERR_CUDA, "clone: device copy failed"));
  l->n_bank = src->n_bank;
  l->n_out[0] = src->n_out[0];
  l->n_out[1] = src->n_out[1];
  l->nb_disabled = src->nb_disabled;
  l->chain_state = src->chain_state;
  l->x_ini = src->x_ini;
  l->wmc = src->wmc;
  l->n_unborn = src->n_unborn;
  l->opt_tally_mode = src->opt_tally_mode;
  l->opt_block = src->opt_block;
  l->opt_bps = src->opt_bps;
  l->opt_retire_batch = src->opt_retire_batch;
  l->opt_birth_chunk = src->opt_birth_chunk;
  l->opt_rng = src->opt_rng;
  l->philox_next_id = src->philox_next_id;
  l->events = src->events;
  l->scatters = src->scatters;
  for (int k = 0; k < 3; ++k) {
    l->n_cls[k] = src->n_cls[k];
    l->w_cls[k] = src->w_cls[k];
  }
  l->launches = src->launches;
  l->gpu_launches = src->gpu_launches;
  l->track_ms = src->track_ms;
  *out = l;
  return MCB200_OK;
}

int mcb200_layer_set_cross_sections(mcb200_layer *l, const float *sigs,
                                    const float *absorption_rates) {
  if (!l) return fail(MCB200_ERR_INVALID, "set_cross_sections: null layer");
  if (sigs) l->sigs.assign(sigs, sigs + l->m);
  if (absorption_rates) l->absorption_rates.assign(absorption_rates, absorption_rates + l->m);
  l->xs_dirty = true;
  return MCB200_OK;
}

int mcb200_layer_get_cross_sections(mcb200_layer *l, float *sigs_out, float *absorption_rates_out) {
  if (!l) return fail(MCB200_ERR_INVALID, "get_cross_sections: null layer");
  if (sigs_out) std::memcpy(sigs_out, l->sigs.data(), (size_t)l->m * sizeof(float));
  if (absorption_rates_out)
    std::memcpy(absorption_rates_out, l->absorption_rates.data(), (size_t)l->m * sizeof(float));
  return MCB200_OK;
}

int mcb200_layer_create_particles(mcb200_layer *l, float x_ini, float wmc, int64_t n,
                                  uint64_t seed) {
  if (!l || n < 0) return fail(MCB200_ERR_INVALID, "create_particles: bad argument");
  if (!(x_ini > l->x_min && x_ini < l->x_max)) return MCB200_OK;  // src/layer.cpp:73
  l->x_ini = x_ini;  // :75-78
  l->wmc = wmc;
  l->n_unborn = n;
  l->chain_state = seed;
  return MCB200_OK;
}

int mcb200_layer_push(mcb200_layer *l, const mcb200_particle *aos, int64_t n) {
  return push_any(l, aos, false, n);
}
int mcb200_layer_push_device(mcb200_layer *l, const void *dev_aos, int64_t n) {
  return push_any(l, dev_aos, true, n);
}

int mcb200_layer_simulate(mcb200_layer *l, int64_t nb_particles, mcb200_counts *counts) {
  if (!l) return fail(MCB200_ERR_INVALID, "simulate: null layer");
  DeviceGuard g(l->device);
  if (!g.ok) return fail(MCB200_ERR_CUDA, "simulate: cudaSetDevice failed");
  long long want = nb_particles < 0 ? l->n_bank + l->n_unborn : (long long)nb_particles;
  if (want > l->n_bank + l->n_unborn) want = l->n_bank + l->n_unborn;
  // with a neighbour's inbox connected, one call = ONE launch (a second launch of the same
  // parity would overwrite the stripes of the first): the rest stays banked for the next cycle
  for (int s = 0; s < 2; ++s)
    if (l->peer[s].base) {
      if (want > l->peer[s].geom.max_take) want = l->peer[s].geom.max_take;
      if (want > l->opt_birth_chunk) want = l->opt_birth_chunk;
    }
  while (want > 0) {
    // src/layer.cpp:244-245: births top the bank up when it holds fewer than asked
    if (l->n_bank < want && l->n_unborn > 0) {
      long long nb = want - l->n_bank;
      if (nb > l->n_unborn) nb = l->n_unborn;
      if (nb > l->opt_birth_chunk) nb = l->opt_birth_chunk;
      int rc = birth(l, nb);
      if (rc) return rc;
    }
    long long take = want < l->n_bank ? want : l->n_bank;
    // keep launches bounded by the birth chunk so outboxes stay bounded too
    if (take > l->opt_birth_chunk) take = l->opt_birth_chunk;
    for (int s = 0; s < 2; ++s)  // a connected neighbour's inbox bounds one launch
      if (l->peer[s].base && take > l->peer[s].geom.max_take) take = l->peer[s].geom.max_take;
    if (take <= 0) break;
    int rc = track(l, take);
    if (rc) return rc;
    want -= take;
  }
  return fill_counts(l, counts);
}

int mcb200_layer_simulate_host(mcb200_layer *l, const mcb200_particle *aos, int64_t n,
                               mcb200_counts *counts) {
  if (!l || n < 0 || (n > 0 && !aos)) return fail(MCB200_ERR_INVALID, "simulate_host: bad argument");
  DeviceGuard g(l->device);
  if (!g.ok) return fail(MCB200_ERR_CUDA, "simulate_host: cudaSetDevice failed");
  if (l->peer[0].base || l->peer[1].base) {
    // a connected neighbour's inbox bounds a launch: take the plain path
    int rc = push_any(l, aos, false, n);
    if (rc) return rc;
    return mcb200_layer_simulate(l, n, counts);
  }
  // chunk sizes grow (x4 from 1M) up to `host_chunk`: only the first, small copy is exposed,
  // every later one runs under the tracking of the chunk before it; few launches, because
  // each one ends with a tail of its longest histories
  const long long cmax = l->opt_host_chunk < l->opt_birth_chunk ? l->opt_host_chunk : l->opt_birth_chunk;
  std::vector<long long> offs, cnts;
  for (long long off = 0, c = (1ll << 20) < cmax ? (1ll << 20) : cmax; off < n;
       c = c * 4 < cmax ? c * 4 : cmax) {
    const long long cnt = (n - off) < c ? (n - off) : c;
    offs.push_back(off);
    cnts.push_back(cnt);
    off += cnt;
  }
  const long long nchunks = (long long)offs.size();
  if (n > 0) {
    if (!l->copy_stream) {
      MCB_CUDA(cudaStreamCreateWithFlags(&l->copy_stream, cudaStreamNonBlocking));
      for (int b = 0; b < 2; ++b) MCB_CUDA(cudaEventCreateWithFlags(&l->ev_copied[b], cudaEventDisableTiming));
    }
    long long need = 0;
    for (long long c : cnts) need = c > need ? c : need;
    if (l->pipe_cap < need) {
      for (int b = 0; b < 2; ++b) {
        cudaFree(l->d_pipe[b]);
        l->d_pipe[b] = nullptr;
      }
      l->pipe_cap = 0;
      for (int b = 0; b < 2; ++b)
        MCB_CUDA(cudaMalloc(&l->d_pipe[b], (size_t)need * sizeof(mcb200_particle)));
      l->pipe_cap = need;
    }
  }
  auto copy_chunk = [&](long long k) -> int {
    const int b = (int)(k & 1);
    MCB_CUDA(cudaMemcpyAsync(l->d_pipe[b], aos + offs[(size_t)k],
                             (size_t)cnts[(size_t)k] * sizeof(mcb200_particle),
                             cudaMemcpyHostToDevice, l->copy_stream));
    MCB_CUDA(cudaEventRecord(l->ev_copied[b], l->copy_stream));
    return MCB200_OK;
  };
  if (nchunks > 0) {
    int rc = copy_chunk(0);
    if (rc) return rc;
  }
  for (long long k = 0; k < nchunks; ++k) {
    const long long cnt = cnts[(size_t)k];
    const int b = (int)(k & 1);
    // the other staging buffer is free: track() of chunk k-1 synchronised the tracking stream
    // after its transpose.  Its copy runs on the copy engine while chunk k is tracked.
    if (k + 1 < nchunks) {
      int rc = copy_chunk(k + 1);
      if (rc) return rc;
    }
    int rc = soa_reserve(l, &l->bank, l->n_bank, l->n_bank + cnt);
    if (rc) return rc;
    MCB_CUDA(cudaStreamWaitEvent(l->stream, l->ev_copied[b], 0));
    MCB_CUDA(mcb::launch_aos_to_soa(cnt, l->d_pipe[b], l->bank.seed + l->n_bank,
                                    l->bank.st + l->n_bank, l->stream));
    l->gpu_launches++;
    l->n_bank += cnt;
    rc = track(l, cnt);   // blocking: returns when chunk k is tracked
    if (rc) return rc;
  }
  return fill_counts(l, counts);
}

int mcb200_layer_reset_tally(mcb200_layer *l) {
  if (!l) return fail(MCB200_ERR_INVALID, "reset_tally: null layer");
  DeviceGuard g(l->device);
  MCB_CUDA(cudaMemsetAsync(l->d_acc, 0, l->acc_words() * sizeof(unsigned), l->stream));
  MCB_CUDA(cudaStreamSynchronize(l->stream));
  for (int k = 0; k < 3; ++k) l->w_cls[k] = 0.0;
  return MCB200_OK;
}

int mcb200_layer_counts(mcb200_layer *l, mcb200_counts *out) {
  if (!l || !out) return fail(MCB200_ERR_INVALID, "counts: null argument");
  return fill_counts(l, out);
}

int mcb200_layer_pop_left(mcb200_layer *l, mcb200_particle *aos, int64_t cap, int64_t *n_out) {
  long long n = 0;
  int rc = pop_side(l, 0, aos, false, cap, &n);
  if (n_out) *n_out = n;
  return rc;
}
int mcb200_layer_pop_right(mcb200_layer *l, mcb200_particle *aos, int64_t cap, int64_t *n_out) {
  long long n = 0;
  int rc = pop_side(l, 1, aos, false, cap, &n);
  if (n_out) *n_out = n;
  return rc;
}
int mcb200_layer_pop_left_device(mcb200_layer *l, void *dev_aos, int64_t cap, int64_t *n_out) {
  long long n = 0;
  int rc = pop_side(l, 0, dev_aos, true, cap, &n);
  if (n_out) *n_out = n;
  return rc;
}
int mcb200_layer_pop_right_device(mcb200_layer *l, void *dev_aos, int64_t cap, int64_t *n_out) {
  long long n = 0;
  int rc = pop_side(l, 1, dev_aos, true, cap, &n);
  if (n_out) *n_out = n;
  return rc;
}

int mcb200_layer_inbox_create(mcb200_layer *l, int64_t max_take,
                              uint8_t handle_out[MCB200_IPC_HANDLE_BYTES],
                              mcb200_inbox_geom *geom_out) {
  if (!l || max_take <= 0 || !geom_out)
    return fail(MCB200_ERR_INVALID, "inbox_create: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == MCB200_IPC_HANDLE_BYTES, "IPC handle size");
  DeviceGuard g(l->device);
  if (l->d_inbox) return fail(MCB200_ERR_INVALID, "inbox_create: the layer already has an inbox");
  cudaDeviceProp prop;
  MCB_CUDA(cudaGetDeviceProperties(&prop, l->device));
  mcb200_inbox_geom q{};
  // senders run the same kernel: at most 4 CTAs of 256 threads per SM (track_configure)
  q.nstripes = prop.multiProcessorCount * 4;
  if (q.nstripes > mcb::kStripes) q.nstripes = mcb::kStripes;
  q.stripe_cap = (int32_t)(2 * ((max_take + q.nstripes - 1) / q.nstripes) + 64);
  q.ovf_cap = max_take;
  q.max_take = max_take;
  const size_t rec_bytes =
      ((size_t)q.nstripes * (size_t)q.stripe_cap + (size_t)q.ovf_cap) * sizeof(mcb200_particle);
  q.fills_offset = (int64_t)((rec_bytes + 255) / 256 * 256);
  q.slot_bytes =
      q.fills_offset + (int64_t)(((size_t)(q.nstripes + 1) * sizeof(unsigned) + 255) / 256 * 256);
  MCB_CUDA(cudaMalloc(&l->d_inbox, 4 * (size_t)q.slot_bytes));
  l->inbox_geom = q;
  // only the fill counters need to start at zero
  for (int s = 0; s < 4; ++s)
    MCB_CUDA(cudaMemsetAsync(l->d_inbox + (size_t)s * (size_t)q.slot_bytes + q.fills_offset, 0,
                             (size_t)(q.nstripes + 1) * sizeof(unsigned), l->stream));
  MCB_CUDA(cudaStreamSynchronize(l->stream));
  *geom_out = q;
  if (handle_out) {
    cudaIpcMemHandle_t h;
    MCB_CUDA(cudaIpcGetMemHandle(&h, l->d_inbox));
    std::memcpy(handle_out, &h, sizeof h);
  }
  return MCB200_OK;
}

int mcb200_layer_connect_peer(mcb200_layer *l, int32_t side,
                              const uint8_t handle[MCB200_IPC_HANDLE_BYTES],
                              const mcb200_inbox_geom *geom) {
  if (!l || side < 0 || side > 1 || !handle || !geom)
    return fail(MCB200_ERR_INVALID, "connect_peer: bad argument");
  DeviceGuard g(l->device);
  if (l->peer[side].base) return fail(MCB200_ERR_INVALID, "connect_peer: side already connected");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof h);
  void *p = nullptr;
  MCB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  l->peer[side].base = static_cast<unsigned char *>(p);
  l->peer[side].ipc = true;
  l->peer[side].geom = *geom;
  return MCB200_OK;
}

int mcb200_layer_connect_local(mcb200_layer *l, int32_t side, mcb200_layer *other) {
  if (!l || side < 0 || side > 1 || !other || !other->d_inbox)
    return fail(MCB200_ERR_INVALID, "connect_local: bad argument (the neighbour needs an inbox)");
  if (l->peer[side].base) return fail(MCB200_ERR_INVALID, "connect_local: side already connected");
  if (other->device != l->device) {
    DeviceGuard g(l->device);
    int can = 0;
    MCB_CUDA(cudaDeviceCanAccessPeer(&can, l->device, other->device));
    if (!can) return fail(MCB200_ERR_CUDA, "connect_local: no peer access between the devices");
    cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
      return fail(MCB200_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
    cudaGetLastError();
  }
  l->peer[side].base = other->d_inbox;
  l->peer[side].ipc = false;
  l->peer[side].geom = other->inbox_geom;
  return MCB200_OK;
}

int mcb200_layer_disconnect_peers(mcb200_layer *l) {
  if (!l) return fail(MCB200_ERR_INVALID, "disconnect_peers: null layer");
  DeviceGuard g(l->device);
  if (l->stream) cudaStreamSynchronize(l->stream);
  for (int s = 0; s < 2; ++s) {
    if (l->peer[s].base && l->peer[s].ipc) cudaIpcCloseMemHandle(l->peer[s].base);
    l->peer[s] = mcb200_layer::Peer{};
  }
  return MCB200_OK;
}

int mcb200_layer_set_exchange_parity(mcb200_layer *l, int32_t parity) {
  if (!l || parity < 0 || parity > 1)
    return fail(MCB200_ERR_INVALID, "set_exchange_parity: bad argument");
  l->parity = parity;
  return MCB200_OK;
}

int mcb200_layer_ingest_inbox(mcb200_layer *l, int32_t from_side, int32_t parity, int64_t *n_out) {
  if (!l || from_side < 0 || from_side > 1 || parity < 0 || parity > 1 || !l->d_inbox)
    return fail(MCB200_ERR_INVALID, "ingest_inbox: bad argument (no inbox?)");
  DeviceGuard g(l->device);
  const mcb200_inbox_geom &q = l->inbox_geom;
  int rc = soa_reserve(l, &l->bank, l->n_bank, l->n_bank + q.max_take);
  if (rc) return rc;
  unsigned char *slot = l->d_inbox + (size_t)(from_side * 2 + parity) * (size_t)q.slot_bytes;
  unsigned *fills = reinterpret_cast<unsigned *>(slot + q.fills_offset);
  MCB_CUDA(cudaMemsetAsync(&l->d_ctr->out_total[0], 0, sizeof(unsigned long long), l->stream));
  MCB_CUDA(mcb::launch_gather_stripes_to_bank(
      reinterpret_cast<const unsigned long long *>(slot), fills, q.nstripes, q.stripe_cap,
      (long long)q.nstripes * q.stripe_cap, q.ovf_cap, l->bank.seed, l->bank.st, l->n_bank,
      &l->d_ctr->out_total[0], l->stream));
  l->gpu_launches++;
  // the slot is written again two cycles from now: its fill counters must be zero by then
  MCB_CUDA(cudaMemsetAsync(fills, 0, (size_t)(q.nstripes + 1) * sizeof(unsigned), l->stream));
  unsigned long long n = 0;
  MCB_CUDA(cudaMemcpyAsync(&n, &l->d_ctr->out_total[0], sizeof n, cudaMemcpyDeviceToHost,
                           l->stream));
  MCB_CUDA(cudaStreamSynchronize(l->stream));
  if ((long long)n > q.max_take)
    return fail(MCB200_ERR_CAPACITY, "ingest_inbox: the inbox was over-filled");
  l->n_bank += (long long)n;
  if (n_out) *n_out = (int64_t)n;
  return MCB200_OK;
}

int mcb200_layer_outbox_device(mcb200_layer *l, int32_t side, void **dev_aos_out, int64_t *n_out) {
  if (!l || side < 0 || side > 1 || !dev_aos_out || !n_out)
    return fail(MCB200_ERR_INVALID, "outbox_device: bad argument");
  *dev_aos_out = l->d_out[side];
  *n_out = l->n_out[side];
  return MCB200_OK;
}

int mcb200_layer_outbox_clear(mcb200_layer *l, int32_t side) {
  if (!l || side < 0 || side > 1) return fail(MCB200_ERR_INVALID, "outbox_clear: bad argument");
  l->n_out[side] = 0;
  return MCB200_OK;
}

int mcb200_layer_weights_absorbed_exact(mcb200_layer *l, uint32_t *out_4m, int32_t *lsb_log2) {
  if (!l || !out_4m) return fail(MCB200_ERR_INVALID, "weights_absorbed_exact: null argument");
  std::vector<unsigned> d;
  int rc = fetch_tally(l, &d);
  if (rc) return rc;
  std::memcpy(out_4m, d.data(), d.size() * sizeof(unsigned));
  if (lsb_log2) *lsb_log2 = mcb::kAccLsbLog2;
  return MCB200_OK;
}

int mcb200_layer_weights_absorbed_f64(mcb200_layer *l, double *out_m) {
  if (!l || !out_m) return fail(MCB200_ERR_INVALID, "weights_absorbed_f64: null argument");
  std::vector<unsigned> d;
  int rc = fetch_tally(l, &d);
  if (rc) return rc;
  for (int i = 0; i < l->m; ++i) out_m[i] = acc_to_double(&d[(size_t)i * mcb::kAccDigits]);
  return MCB200_OK;
}

int mcb200_layer_weights_absorbed(mcb200_layer *l, float *out_m) {
  if (!l || !out_m) return fail(MCB200_ERR_INVALID, "weights_absorbed: null argument");
  std::vector<unsigned> d;
  int rc = fetch_tally(l, &d);
  if (rc) return rc;
  for (int i = 0; i < l->m; ++i)
    out_m[i] = (float)acc_to_double(&d[(size_t)i * mcb::kAccDigits]);
  return MCB200_OK;
}

int mcb200_layer_dump_WA(mcb200_layer *l, const char *path) {
  if (!l) return fail(MCB200_ERR_INVALID, "dump_WA: null layer");
  std::vector<float> w((size_t)l->m);
  int rc = mcb200_layer_weights_absorbed(l, w.data());
  if (rc) return rc;
  FILE *f = std::fopen(path ? path : "WA.out", "w");
  if (!f) return fail(MCB200_ERR_INVALID, "Couldn't open file WA.out for writing.");
  const float ldx = (l->x_max - l->x_min) / l->m;  // the reference prints with Layer::dx
  for (int i = 0; i < l->m; ++i) {                 // src/layer.cpp:373-377
    const float base = l->x_min + (i * ldx);
    const float ratio = w[(size_t)i] / ldx;
    std::fprintf(f, "%.4e %.3e\n", (double)base + 0.5 * (double)ldx, (double)ratio);
  }
  std::fclose(f);
  return MCB200_OK;
}

void *mcb200_layer_stream(mcb200_layer *l) { return l ? (void *)l->stream : nullptr; }

int mcb200_layer_set_option(mcb200_layer *l, const char *key, int64_t value) {
  if (!l || !key) return fail(MCB200_ERR_INVALID, "set_option: null argument");
  const std::string k(key);
  if (k == "tally_mode") l->opt_tally_mode = (int)value;
  else if (k == "block") l->opt_block = (int)value;
  else if (k == "blocks_per_sm") l->opt_bps = (int)value;
  else if (k == "retire_batch") l->opt_retire_batch = (int)value;
  else if (k == "rng") {
    if (value != 0 && value != 1) return fail(MCB200_ERR_INVALID, "set_option: rng is 0 (LCG) or 1 (Philox)");
    if (l->n_bank > 0) return fail(MCB200_ERR_INVALID, "set_option: rng cannot change while particles are banked");
    l->opt_rng = (int)value;
    l->cfg_dirty = true;
  } else if (k == "host_chunk") {
    if (value <= 0) return fail(MCB200_ERR_INVALID, "set_option: host_chunk must be positive");
    l->opt_host_chunk = value;
  } else if (k == "birth_chunk") {
    if (value <= 0) return fail(MCB200_ERR_INVALID, "set_option: birth_chunk must be positive");
    l->opt_birth_chunk = value;
  } else
    return fail(MCB200_ERR_INVALID, "set_option: unknown key " + k);
  l->cfg_dirty = true;
  return MCB200_OK;
}

// ---- known-answer-test entry points ------------------------------------

int mcb200_test_rnd_real(int device, uint64_t *seeds_host, float *out_host, int64_t n) {
  if (n < 0 || (n > 0 && (!seeds_host || !out_host)))
    return fail(MCB200_ERR_INVALID, "test_rnd_real: bad argument");
  if (n == 0) return MCB200_OK;
  DeviceGuard g(device);
  if (!g.ok) return fail(MCB200_ERR_CUDA, "test_rnd_real: cudaSetDevice failed");
  DevScratch<unsigned long long> ds;
  DevScratch<float> dout;
  MCB_CUDA(ds.alloc((size_t)n));
  MCB_CUDA(dout.alloc((size_t)n));
  MCB_CUDA(cudaMemcpy(ds.p, seeds_host, (size_t)n * 8, cudaMemcpyHostToDevice));
  MCB_CUDA(mcb::launch_test_rnd_real(n, ds.p, dout.p, nullptr));
  MCB_CUDA(cudaMemcpy(seeds_host, ds.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  MCB_CUDA(cudaMemcpy(out_host, dout.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return MCB200_OK;
}

int mcb200_test_philox(int device, const uint32_t *c0_host, const uint32_t *c1_host,
                       const uint32_t *key_host, uint32_t *out2_host, int64_t n) {
  if (n < 0 || (n > 0 && (!c0_host || !c1_host || !key_host || !out2_host)))
    return fail(MCB200_ERR_INVALID, "test_philox: bad argument");
  if (n == 0) return MCB200_OK;
  DeviceGuard g(device);
  if (!g.ok) return fail(MCB200_ERR_CUDA, "test_philox: cudaSetDevice failed");
  DevScratch<unsigned> d0, d1, dk;
  DevScratch<uint2> dout;
  MCB_CUDA(d0.alloc((size_t)n));
  MCB_CUDA(d1.alloc((size_t)n));
  MCB_CUDA(dk.alloc((size_t)n));
  MCB_CUDA(dout.alloc((size_t)n));
  MCB_CUDA(cudaMemcpy(d0.p, c0_host, (size_t)n * 4, cudaMemcpyHostToDevice));
  MCB_CUDA(cudaMemcpy(d1.p, c1_host, (size_t)n * 4, cudaMemcpyHostToDevice));
  MCB_CUDA(cudaMemcpy(dk.p, key_host, (size_t)n * 4, cudaMemcpyHostToDevice));
  MCB_CUDA(mcb::launch_test_philox(n, d0.p, d1.p, dk.p, dout.p, nullptr));
  MCB_CUDA(cudaMemcpy(out2_host, dout.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return MCB200_OK;
}

static int test_math(int which, int device, const float *in_host, float *out_host, int64_t n) {
  if (n < 0 || (n > 0 && (!in_host || !out_host)))
    return fail(MCB200_ERR_INVALID, "test_math: bad argument");
  if (n == 0) return MCB200_OK;
  DeviceGuard g(device);
  if (!g.ok) return fail(MCB200_ERR_CUDA, "test_math: cudaSetDevice failed");
  DevScratch<float> din, dout;
  MCB_CUDA(din.alloc((size_t)n));
  MCB_CUDA(dout.alloc((size_t)n));
  MCB_CUDA(cudaMemcpy(din.p, in_host, (size_t)n * 4, cudaMemcpyHostToDevice));
  MCB_CUDA(mcb::launch_test_math(which, n, din.p, dout.p, nullptr));
  MCB_CUDA(cudaMemcpy(out_host, dout.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return MCB200_OK;
}
int mcb200_test_logf(int device, const float *in_host, float *out_host, int64_t n) {
  return test_math(0, device, in_host, out_host, n);
}
int mcb200_test_expf(int device, const float *in_host, float *out_host, int64_t n) {
  return test_math(1, device, in_host, out_host, n);
}

int mcb200_test_edge_distance(int device, const float *a_host, const float *mu_host,
                              float *out_host, int64_t n) {
  if (n < 0 || (n > 0 && (!a_host || !mu_host || !out_host)))
    return fail(MCB200_ERR_INVALID, "test_edge_distance: bad argument");
  if (n == 0) return MCB200_OK;
  DeviceGuard g(device);
  if (!g.ok) return fail(MCB200_ERR_CUDA, "test_edge_distance: cudaSetDevice failed");
  DevScratch<float> da, db, dout;
  MCB_CUDA(da.alloc((size_t)n));
  MCB_CUDA(db.alloc((size_t)n));
  MCB_CUDA(dout.alloc((size_t)n));
  MCB_CUDA(cudaMemcpy(da.p, a_host, (size_t)n * 4, cudaMemcpyHostToDevice));
  MCB_CUDA(cudaMemcpy(db.p, mu_host, (size_t)n * 4, cudaMemcpyHostToDevice));
  MCB_CUDA(mcb::launch_test_div(n, da.p, db.p, dout.p, nullptr));
  MCB_CUDA(cudaMemcpy(out_host, dout.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return MCB200_OK;
}

int mcb200_test_accumulate(int device, const float *in_host, int64_t n, uint32_t *out4,
                           double *out_f64) {
  if (n < 0 || (n > 0 && !in_host) || !out4)
    return fail(MCB200_ERR_INVALID, "test_accumulate: bad argument");
  DeviceGuard g(device);
  if (!g.ok) return fail(MCB200_ERR_CUDA, "test_accumulate: cudaSetDevice failed");
  DevScratch<float> din;
  DevScratch<unsigned> dacc;
  MCB_CUDA(din.alloc((size_t)(n > 0 ? n : 1)));
  MCB_CUDA(dacc.alloc(mcb::kAccDigits + 1));
  MCB_CUDA(cudaMemset(dacc.p, 0, (mcb::kAccDigits + 1) * sizeof(unsigned)));
  if (n > 0) MCB_CUDA(cudaMemcpy(din.p, in_host, (size_t)n * 4, cudaMemcpyHostToDevice));
  MCB_CUDA(mcb::launch_test_accumulate(n, din.p, dacc.p, nullptr));
  unsigned h[mcb::kAccDigits + 1];
  MCB_CUDA(cudaMemcpy(h, dacc.p, sizeof h, cudaMemcpyDeviceToHost));
  for (int j = 0; j < mcb::kAccDigits; ++j) out4[j] = h[j];
  if (out_f64) *out_f64 = acc_to_double(h);
  if (h[mcb::kAccDigits]) return fail(MCB200_ERR_RANGE, "test_accumulate: value outside (-2^7, 2^7)");
  return MCB200_OK;
}

int mcb200_test_birth(int device, float x_ini, float wmc, float dx, int64_t n, uint64_t seed,
                      mcb200_particle *out_host) {
  if (n < 0 || (n > 0 && !out_host)) return fail(MCB200_ERR_INVALID, "test_birth: bad argument");
  if (n == 0) return MCB200_OK;
  DeviceGuard g(device);
  if (!g.ok) return fail(MCB200_ERR_CUDA, "test_birth: cudaSetDevice failed");
  DevScratch<unsigned long long> ds;
  DevScratch<float4> dst;
  DevScratch<mcb200_particle> daos;
  MCB_CUDA(ds.alloc((size_t)n));
  MCB_CUDA(dst.alloc((size_t)n));
  MCB_CUDA(daos.alloc((size_t)n));
  const mcb::JumpTable jt = mcb::make_jump_table(mcb::kSeedG, mcb::kSeedC);
  const float cell = x_ini / dx;
  MCB_CUDA(mcb::launch_birth(n, seed, jt, x_ini, wmc, (int)cell, ds.p, dst.p, nullptr));
  MCB_CUDA(mcb::launch_soa_to_aos(n, ds.p, dst.p, daos.p, nullptr));
  MCB_CUDA(cudaMemcpy(out_host, daos.p, (size_t)n * 24, cudaMemcpyDeviceToHost));
  return MCB200_OK;
}

}  // extern "C"
