// mcb_world.cu -- host side of the persistent multi-GPU run ("world") and its C ABI
// (include/mcb200.h, mcb200_world_*).  Stands in for Worker::Worker / Worker::spin /
// Worker::gather_weights_absorbed (src/worker.cpp:16-34,183-216, src/worker_sync.cpp:24-135):
// one rank = one contiguous sub-slab on one GPU; a run is ONE resident kernel per rank
// (mcb_world_kernel.cu) -- the host launches it and waits.  Nothing here computes physics.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "mcb_host.hpp"
#include "mcb_world.cuh"

using mcb::DeviceGuard;
using mcb::fail;

struct mcb200_world {
  // --- the decomposition (decompose_domain, src/layer.cpp:17-42)
  int device = 0, rank = 0, K = 1;
  float x_min = 0, x_max = 0, x_ini = 0, dx = 0, minw = 0;
  int nb_cells = 0;
  std::vector<int> cuts;      // K + 1 cell boundaries
  int lo = 0, M = 0;          // this rank's cells [lo, lo + M)
  int home_rank = 0;          // the rank whose cells contain the source (births + global count)
  int src_cell = -1;          // (int)((x_ini - x_min) / dx), :30; -1 = x_ini outside the slab
  int src_index = 0;          // (int)(x_ini / dx), :106
  std::vector<float> sigs, absorption_rates;   // this rank's slices
  // --- launch shape
  int V = 1, cpw = 1, S = 0;  // windows, CTAs per window, stripes (= warps per window)
  std::vector<int> win_lo;    // V + 1 boundaries (global cells)
  mcb::WorldLaunch cfg{};
  unsigned ring_cap = 0;
  unsigned bank_cap = 0, bank_log2 = 0;
  unsigned long long inflight_limit = 0;
  int retire_batch = 0;
  long long max_run_ms = 0, stall_ms = 15000;
  int rng = 0;                // 0 = LCG (parity mode), 1 = Philox2x32-10
  // --- device state
  cudaStream_t stream = nullptr, side = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  mcb::CellXs *d_xs = nullptr;
  unsigned long long *d_acc = nullptr;       // u64[2][M + kAccExtra]
  mcb::WindowDesc *d_win = nullptr;
  mcb::WorldCounters *d_ctr = nullptr, *h_ctr = nullptr;
  unsigned long long *h_prog = nullptr;      // pinned: {disabled_global, born} of the home rank
  unsigned **d_done_ptrs = nullptr;
  unsigned char *d_xblock = nullptr;         // exported exchange block (ctrl + outer rings)
  unsigned char *d_inner = nullptr;          // inner links + banks
  size_t inner_bytes = 0;
  mcb200_world_geom geom{};
  struct Peer {
    unsigned char *base = nullptr;
    bool ipc = false;
    mcb200_world_geom geom{};
  };
  std::vector<Peer> peers;                   // indexed by rank (own entry: base = d_xblock)
  bool table_dirty = true;
  // --- the run in flight
  long long nb_particles = 0;
  unsigned long long seed0 = 0;
  bool prepared = false, launched = false;
  bool bank_dirty = false;    // the last run failed: banks may hold stale records
  long long gpu_launches = 0;

  int ncell() const { return M + mcb::kAccExtra; }
};

namespace {

size_t align256(size_t b) { return (b + 255) / 256 * 256; }

unsigned pow2_floor(unsigned long long v) {
  unsigned p = 1;
  while ((unsigned long long)p * 2 <= v && p < (1u << 30)) p *= 2;
  return p;
}
unsigned pow2_ceil(unsigned long long v) {
  unsigned p = 1;
  while (p < v && p < (1u << 30)) p *= 2;
  return p;
}
unsigned ilog2(unsigned v) {
  unsigned l = 0;
  while ((1u << l) < v) ++l;
  return l;
}

// layout of the inner-link area of one rank: per inner boundary b (between windows b, b+1)
// two rings (right-going, left-going), each followed by its credit counters; then the banks
struct InnerLayout {
  size_t ring_bytes, cnt_bytes, link_bytes, bank_rec_bytes, bank_bytes, total;
  size_t banks_off;
};
InnerLayout inner_layout(int V, int S, int cpw, unsigned ring_cap, unsigned bank_cap) {
  InnerLayout q{};
  q.ring_bytes = align256((size_t)S * ring_cap * sizeof(mcb200_particle));
  q.cnt_bytes = align256((size_t)S * sizeof(unsigned));
  q.link_bytes = 2 * (q.ring_bytes + q.cnt_bytes);
  q.banks_off = (size_t)(V > 1 ? V - 1 : 0) * q.link_bytes;
  // one bank per CTA of the window (bank_cap records each), then their cursors
  q.bank_rec_bytes = align256((size_t)cpw * bank_cap * sizeof(mcb200_particle));
  q.bank_bytes = q.bank_rec_bytes + align256((size_t)cpw * 2 * sizeof(unsigned));
  q.total = q.banks_off + (size_t)V * q.bank_bytes;
  return q;
}

// pick (block, CTAs per SM, windows): as few windows as let every CTA keep its window's cell
// constants + private tally in shared memory at full occupancy; wide slabs get more windows
int choose_shape(mcb200_world *w, const mcb200_world_desc *d, int m_max_all) {
  cudaDeviceProp prop;
  MCB_CUDA(cudaGetDeviceProperties(&prop, w->device));
  const int blocks[3] = {256, 512, 1024};
  // Candidates: cell constants next to the tally in shared memory, or left in global memory /
  // L2 (half the shared memory per cell: fewer, wider windows for sub-slabs of ~1e6 cells) x
  // CTA size.  Taken: the shape that keeps the most lanes busy (windows x CTAs per window x
  // threads); ties go to constants in shared memory, then to the smaller CTA.
  long long best_lanes = -1;
  mcb::WorldLaunch best_cfg{};
  int best_V = 0, best_cpw = 0;
  for (int xs = d->xs_global ? 0 : 1; xs >= 0; --xs) {
    const size_t per_cell = (xs ? sizeof(mcb::CellXs) : 0) + mcb::kAccDigits * sizeof(unsigned);
    for (int bi = 0; bi < 3; ++bi) {
      const int block = d->block > 0 ? d->block : blocks[bi];
      if (block % 32 || block < 32 || block > 1024)
        return fail(MCB200_ERR_INVALID, "world: block must be a multiple of 32 in [32, 1024]");
      if (d->block > 0 && bi > 0) break;
      const int bps = 1024 / block > 0 ? 1024 / block : 1;
      const size_t budget = prop.sharedMemPerMultiprocessor / (size_t)bps - 1024;
      const size_t fixed = mcb::world_smem_bytes(0, block, xs != 0);
      if (budget <= fixed + 64) continue;
      const int mw_max = (int)((budget - fixed) / per_cell);
      int V = d->windows > 0 ? d->windows : (m_max_all + mw_max - 1) / mw_max;
      if (V < 1) V = 1;
      if (V > m_max_all) V = m_max_all;
      const int mw = (m_max_all + V - 1) / V;
      if (mcb::world_smem_bytes(mw, block, xs != 0) > prop.sharedMemPerBlockOptin) continue;
      mcb::WorldLaunch cfg{};
      int per_sm = 0;
      MCB_CUDA(mcb::world_configure(w->device, mw, block, xs != 0, &cfg, &per_sm));
      int capacity = cfg.grid;
      if (d->max_ctas > 0) {   // the caller's cap counts 256-thread CTAs
        const int cap_b = d->max_ctas * 256 / block > 0 ? d->max_ctas * 256 / block : 1;
        if (cap_b < capacity) capacity = cap_b;
      }
      if (V > capacity || V > mcb::kWorldMaxWindows) continue;   // too many windows for this CTA size
      const int cpw = capacity / V;
      if (d->windows <= 0 && V > 1) {
        // more, narrower windows than shared memory demands if that fills the CTA slots the
        // division left over (e.g. 320 windows needed, 592 slots: 592 windows of one CTA)
        int Vf = capacity / cpw;
        if (Vf > mcb::kWorldMaxWindows) Vf = mcb::kWorldMaxWindows;
        if (Vf > m_max_all) Vf = m_max_all;
        if (Vf > V) V = Vf;
      }
      const long long lanes = (long long)V * cpw * block;
      if (lanes > best_lanes) {
        best_lanes = lanes;
        best_cfg = cfg;
        best_V = V;
        best_cpw = cpw;
      }
    }
  }
  if (best_lanes < 0)
    return fail(MCB200_ERR_INVALID, "world: the sub-slab does not fit the GPU in windows "
                                    "(too many cells for the CTAs available)");
  w->cfg = best_cfg;
  w->V = best_V;
  w->cpw = best_cpw;
  w->cfg.grid = w->V * w->cpw;
  w->S = w->cpw * (w->cfg.block / 32);
  // (re)apply the launch attributes of the chosen variant
  int per_sm = 0;
  const int mw = (m_max_all + w->V - 1) / w->V;
  mcb::WorldLaunch again{};
  MCB_CUDA(mcb::world_configure(w->device, mw, w->cfg.block, w->cfg.xs_smem != 0, &again, &per_sm));
  w->cfg.smem = again.smem;
  return MCB200_OK;
}

void free_device(mcb200_world *w) {
  DeviceGuard g(w->device);
  if (w->stream) cudaStreamSynchronize(w->stream);
  for (auto &p : w->peers)
    if (p.base && p.ipc) cudaIpcCloseMemHandle(p.base);
  w->peers.clear();
  cudaFree(w->d_xs);
  cudaFree(w->d_acc);
  cudaFree(w->d_win);
  cudaFree(w->d_ctr);
  cudaFree(w->d_done_ptrs);
  cudaFree(w->d_xblock);
  cudaFree(w->d_inner);
  if (w->h_ctr) cudaFreeHost(w->h_ctr);
  if (w->h_prog) cudaFreeHost(w->h_prog);
  if (w->ev0) cudaEventDestroy(w->ev0);
  if (w->ev1) cudaEventDestroy(w->ev1);
  if (w->stream) cudaStreamDestroy(w->stream);
  if (w->side) cudaStreamDestroy(w->side);
  cudaGetLastError();
}

int alloc_device(mcb200_world *w) {
  DeviceGuard g(w->device);
  if (!g.ok) return fail(MCB200_ERR_CUDA, "world: cudaSetDevice failed");
  MCB_CUDA(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
  MCB_CUDA(cudaStreamCreateWithFlags(&w->side, cudaStreamNonBlocking));
  MCB_CUDA(cudaEventCreate(&w->ev0));
  MCB_CUDA(cudaEventCreate(&w->ev1));
  MCB_CUDA(cudaMalloc(&w->d_xs, (size_t)w->M * sizeof(mcb::CellXs)));
  MCB_CUDA(cudaMalloc(&w->d_acc, 2 * (size_t)w->ncell() * sizeof(unsigned long long)));
  MCB_CUDA(cudaMalloc(&w->d_win, (size_t)w->V * sizeof(mcb::WindowDesc)));
  MCB_CUDA(cudaMalloc(&w->d_ctr, sizeof(mcb::WorldCounters)));
  MCB_CUDA(cudaMalloc(&w->d_done_ptrs, (size_t)w->K * sizeof(unsigned *)));
  MCB_CUDA(cudaMallocHost(&w->h_ctr, sizeof(mcb::WorldCounters)));
  MCB_CUDA(cudaMallocHost(&w->h_prog, 2 * sizeof(unsigned long long)));
  // the exported block: ctrl | rings filled by the left / right neighbour | credits for my sends
  mcb200_world_geom &q = w->geom;
  q.rank = w->rank;
  q.world_size = w->K;
  q.stripes = w->S;
  q.ring_cap = (int32_t)w->ring_cap;
  size_t off = align256(sizeof(mcb::WorldCtrl));
  const size_t ring_bytes = align256((size_t)w->S * w->ring_cap * sizeof(mcb200_particle));
  const size_t cnt_bytes = align256((size_t)w->S * sizeof(unsigned));
  for (int s = 0; s < 2; ++s) {
    q.off_rec[s] = (int64_t)off;
    off += ring_bytes;
  }
  for (int s = 0; s < 2; ++s) {
    q.off_credit[s] = (int64_t)off;
    off += cnt_bytes;
  }
  q.off_chain = (int64_t)off;   // one counter per stripe: histories of that chain disabled anywhere
  off += cnt_bytes;
  q.block_bytes = (int64_t)off;
  MCB_CUDA(cudaMalloc(&w->d_xblock, off));
  MCB_CUDA(cudaMemsetAsync(w->d_xblock, 0, off, w->stream));
  const InnerLayout il = inner_layout(w->V, w->S, w->cpw, w->ring_cap, w->bank_cap);
  w->inner_bytes = il.total;
  MCB_CUDA(cudaMalloc(&w->d_inner, il.total));
  MCB_CUDA(cudaMemsetAsync(w->d_inner, 0, il.total, w->stream));
  MCB_CUDA(cudaMemsetAsync(w->d_acc, 0, 2 * (size_t)w->ncell() * sizeof(unsigned long long),
                           w->stream));
  // cell constants of the rank's slice
  std::vector<mcb::CellXs> xs((size_t)w->M);
  for (int i = 0; i < w->M; ++i)
    xs[(size_t)i] = mcb::host_cell_xs(w->sigs[(size_t)i], w->absorption_rates[(size_t)i]);
  MCB_CUDA(cudaMemcpyAsync(w->d_xs, xs.data(), xs.size() * sizeof(mcb::CellXs),
                           cudaMemcpyHostToDevice, w->stream));
  const mcb::JumpTable jt = mcb::make_jump_table(mcb::kSeedG, mcb::kSeedC);
  MCB_CUDA(mcb::world_upload_jump_table(jt));
  MCB_CUDA(cudaStreamSynchronize(w->stream));
  w->peers.assign((size_t)w->K, mcb200_world::Peer{});
  w->peers[(size_t)w->rank].base = w->d_xblock;
  w->peers[(size_t)w->rank].geom = w->geom;
  return MCB200_OK;
}

// (re)build the device table of windows from the current peer mappings
int upload_windows(mcb200_world *w) {
  const int r = w->rank, K = w->K, V = w->V;
  for (int nb = r - 1; nb <= r + 1; nb += 2)
    if (nb >= 0 && nb < K && !w->peers[(size_t)nb].base)
      return fail(MCB200_ERR_INVALID, "world: neighbour rank " + std::to_string(nb) + " is not connected");
  if (!w->peers[(size_t)w->home_rank].base)
    return fail(MCB200_ERR_INVALID, "world: the home rank is not connected");
  if (r == w->home_rank)
    for (int k = 0; k < K; ++k)
      if (!w->peers[(size_t)k].base)
        return fail(MCB200_ERR_INVALID, "world: the home rank must be connected to every rank (missing " +
                                            std::to_string(k) + ")");
  const InnerLayout il = inner_layout(V, w->S, w->cpw, w->ring_cap, w->bank_cap);
  auto inner_ring = [&](int b, int dir) {   // dir 0 = right-going (b -> b+1), 1 = left-going
    return w->d_inner + (size_t)b * il.link_bytes + (size_t)dir * (il.ring_bytes + il.cnt_bytes);
  };
  std::vector<mcb::WindowDesc> tab((size_t)V);
  for (int v = 0; v < V; ++v) {
    mcb::WindowDesc &d = tab[(size_t)v];
    std::memset(&d, 0, sizeof d);
    for (int s = 0; s < 2; ++s) {
      const int nv = s == 0 ? v - 1 : v + 1;        // neighbouring window inside the rank
      const int nr = s == 0 ? r - 1 : r + 1;        // neighbouring rank beyond the edge
      mcb::LinkOut &o = d.out[s];
      mcb::LinkIn &in = d.in[s];
      if (nv >= 0 && nv < V) {
        const int b = s == 0 ? v - 1 : v;           // inner boundary between b and b + 1
        // this window sends over boundary b in direction `s == 1 ? right : left`
        unsigned char *snd = inner_ring(b, s == 1 ? 0 : 1);
        unsigned char *rcv = inner_ring(b, s == 1 ? 1 : 0);
        o.rec = reinterpret_cast<unsigned long long *>(snd);
        o.credit = reinterpret_cast<const unsigned *>(snd + il.ring_bytes);
        o.mode = 1;
        o.outer = 0;
        in.rec = reinterpret_cast<const unsigned long long *>(rcv);
        in.credit = reinterpret_cast<unsigned *>(rcv + il.ring_bytes);
        in.present = 1;
      } else if (nr >= 0 && nr < K) {
        // the neighbour rank's exchange block: I store into ITS "from side 1-s" ring, it
        // stores into MY "from side s" ring; credits live with the producer
        const mcb200_world::Peer &pr = w->peers[(size_t)nr];
        o.rec = reinterpret_cast<unsigned long long *>(pr.base + pr.geom.off_rec[1 - s]);
        o.credit = reinterpret_cast<const unsigned *>(w->d_xblock + w->geom.off_credit[s]);
        o.mode = 1;
        o.outer = 1;
        in.rec = reinterpret_cast<const unsigned long long *>(w->d_xblock + w->geom.off_rec[s]);
        in.credit = reinterpret_cast<unsigned *>(pr.base + pr.geom.off_credit[1 - s]);
        in.present = 1;
      } else {
        o.mode = 0;   // global border (rank 0's left / rank K-1's right): absorb
        in.present = 0;
      }
    }
    unsigned char *bank = w->d_inner + il.banks_off + (size_t)v * il.bank_bytes;
    d.bank.rec = reinterpret_cast<unsigned long long *>(bank);
    d.bank.ht = reinterpret_cast<unsigned *>(bank + il.bank_rec_bytes);
    d.bank.cap = w->bank_cap;
    d.bank.log2cap = w->bank_log2;
  }
  MCB_CUDA(cudaMemcpyAsync(w->d_win, tab.data(), tab.size() * sizeof(mcb::WindowDesc),
                           cudaMemcpyHostToDevice, w->stream));
  std::vector<unsigned *> done((size_t)K, nullptr);
  for (int k = 0; k < K; ++k)
    if (w->peers[(size_t)k].base)
      done[(size_t)k] = &reinterpret_cast<mcb::WorldCtrl *>(w->peers[(size_t)k].base)->done;
  if (r != w->home_rank)   // only the home rank raises `done` flags
    for (int k = 0; k < K; ++k) done[(size_t)k] = &reinterpret_cast<mcb::WorldCtrl *>(w->d_xblock)->done;
  MCB_CUDA(cudaMemcpyAsync(w->d_done_ptrs, done.data(), done.size() * sizeof(unsigned *),
                           cudaMemcpyHostToDevice, w->stream));
  MCB_CUDA(cudaStreamSynchronize(w->stream));
  w->table_dirty = false;
  return MCB200_OK;
}

int fetch_world_tally(mcb200_world *w, std::vector<unsigned> *digits, double w_cls[3]) {
  DeviceGuard g(w->device);
  const size_t words = 4 * (size_t)w->ncell();
  std::vector<unsigned> raw(words);
  MCB_CUDA(cudaMemcpyAsync(raw.data(), w->d_acc, words * sizeof(unsigned), cudaMemcpyDeviceToHost,
                           w->stream));
  MCB_CUDA(cudaStreamSynchronize(w->stream));
  std::vector<unsigned> all((size_t)w->ncell() * mcb::kAccDigits);
  mcb::acc_halves_to_digits(raw.data(), w->ncell(), w->ncell(), all.data());
  if (digits) digits->assign(all.begin(), all.begin() + (size_t)w->M * mcb::kAccDigits);
  if (w_cls)
    for (int k = 0; k < 3; ++k)
      w_cls[k] = mcb::acc_to_double(&all[((size_t)w->M + (size_t)k) * mcb::kAccDigits]);
  return MCB200_OK;
}

}  // namespace

extern "C" {

int mcb200_world_create(const mcb200_world_desc *d, mcb200_world **out) {
  if (!d || !out) return fail(MCB200_ERR_INVALID, "world_create: null argument");
  *out = nullptr;
  if (d->abi_version != MCB200_ABI_VERSION)
    return fail(MCB200_ERR_INVALID, "world_create: abi_version mismatch");
  if (d->world_size < 1 || d->world_size > mcb::kWorldMaxRanks || d->rank < 0 ||
      d->rank >= d->world_size)
    return fail(MCB200_ERR_INVALID, "world_create: bad rank / world_size");
  if (d->nb_cells < d->world_size)
    return fail(MCB200_ERR_INVALID, "world_create: fewer cells than ranks");
  if (!(d->x_max > d->x_min) || !std::isfinite(d->x_min) || !std::isfinite(d->x_max))
    return fail(MCB200_ERR_INVALID, "world_create: need finite x_min < x_max");
  int ndev = 0;
  MCB_CUDA(cudaGetDeviceCount(&ndev));
  if (d->device < 0 || d->device >= ndev)
    return fail(MCB200_ERR_INVALID, "world_create: no such CUDA device");

  mcb200_world *w = new (std::nothrow) mcb200_world();
  if (!w) return fail(MCB200_ERR_NOMEM, "world_create: out of host memory");
  auto bail = [&](int rc) {
    const std::string keep = mcb::get_last_error();
    free_device(w);
    delete w;
    mcb::set_last_error(keep);
    return rc;
  };
  w->device = d->device;
  w->rank = d->rank;
  w->K = d->world_size;
  w->x_min = d->x_min;
  w->x_max = d->x_max;
  w->x_ini = d->x_ini;
  w->minw = d->particle_min_weight;
  w->nb_cells = d->nb_cells;
  const int K = w->K;
  // the sub-slabs: caller's cuts, or the reference's split (src/layer.cpp:24-27)
  w->cuts.resize((size_t)K + 1);
  if (d->cuts) {
    for (int k = 0; k <= K; ++k) w->cuts[(size_t)k] = d->cuts[k];
    bool ok = w->cuts[0] == 0 && w->cuts[(size_t)K] == d->nb_cells;
    for (int k = 0; k < K; ++k) ok = ok && w->cuts[(size_t)k] < w->cuts[(size_t)k + 1];
    if (!ok) return bail(fail(MCB200_ERR_INVALID, "world_create: cuts must ascend from 0 to nb_cells"));
  } else {
    const int cells_per_layer = d->nb_cells / K, num_with_extra = d->nb_cells % K;
    for (int k = 0; k <= K; ++k)
      w->cuts[(size_t)k] = k * cells_per_layer + (k < num_with_extra ? k : num_with_extra);
  }
  w->lo = w->cuts[(size_t)w->rank];
  w->M = w->cuts[(size_t)w->rank + 1] - w->lo;
  int m_max_all = 0;
  for (int k = 0; k < K; ++k) {
    const int mk = w->cuts[(size_t)k + 1] - w->cuts[(size_t)k];
    if (mk > m_max_all) m_max_all = mk;
  }
  // ONE dx for the whole slab (src/layer.cpp:29), the source cell (:30) and its owner (:34)
  w->dx = (d->x_max - d->x_min) / ((float)d->nb_cells);
  w->home_rank = 0;
  w->src_cell = -1;
  if (d->x_ini > d->x_min && d->x_ini < d->x_max) {
    const float rel = (d->x_ini - d->x_min) / w->dx;
    const int cell_ini = (int)rel;
    if (cell_ini >= 0 && cell_ini < d->nb_cells) {
      w->src_cell = cell_ini;
      for (int k = 0; k < K; ++k)
        if (cell_ini >= w->cuts[(size_t)k] && cell_ini < w->cuts[(size_t)k + 1]) w->home_rank = k;
    }
  }
  {
    const float cell = d->x_ini / w->dx;   // src/layer.cpp:106, ignores x_min
    w->src_index = (int)cell;
  }
  // cross-sections: slices of the ONE global table
  {
    std::vector<float> s((size_t)d->nb_cells), a((size_t)d->nb_cells);
    mcb200_default_cross_sections(d->x_min, d->x_max, d->nb_cells, s.data(), a.data());
    if (d->sigs) s.assign(d->sigs, d->sigs + d->nb_cells);
    if (d->absorption_rates) a.assign(d->absorption_rates, d->absorption_rates + d->nb_cells);
    w->sigs.assign(s.begin() + w->lo, s.begin() + w->lo + w->M);
    w->absorption_rates.assign(a.begin() + w->lo, a.begin() + w->lo + w->M);
  }
  {
    DeviceGuard g(w->device);
    int rc = choose_shape(w, d, m_max_all);
    if (rc) return bail(rc);
  }
  // windows of THIS rank: equal cell counts
  w->win_lo.resize((size_t)w->V + 1);
  if (w->V > w->M) return bail(fail(MCB200_ERR_INVALID, "world_create: more windows than cells"));
  if (w->V > mcb::kWorldMaxWindows)
    return bail(fail(MCB200_ERR_INVALID, "world_create: too many windows (" + std::to_string(w->V) + ")"));
  for (int v = 0; v <= w->V; ++v)
    w->win_lo[(size_t)v] = w->lo + (int)(((long long)w->M * v) / w->V);
  {
    int mw = 0;
    for (int v = 0; v < w->V; ++v)
      if (w->win_lo[(size_t)v + 1] - w->win_lo[(size_t)v] > mw) mw = w->win_lo[(size_t)v + 1] - w->win_lo[(size_t)v];
    if (mcb::world_smem_bytes(mw, w->cfg.block, w->cfg.xs_smem != 0) > w->cfg.smem)
      return bail(fail(MCB200_ERR_INVALID, "world_create: internal: window larger than planned"));
  }
  // histories in flight: a few per lane of the whole world keeps every GPU fed
  w->inflight_limit = d->inflight_limit > 0
                          ? (unsigned long long)d->inflight_limit
                          : 4ull * (unsigned long long)K * (unsigned long long)w->cfg.grid *
                                (unsigned long long)w->cfg.block;
  // rings: a stripe holds what one chain of warps (stripe s of every window of every rank) can
  // have in flight -- the records pile up in front of the slowest window, and a full ring costs
  // the sender its lanes -- within [64, 8192] records, at least ~2M records per link
  if (d->ring_cap > 0) {
    if (d->ring_cap < 32 || (d->ring_cap & (d->ring_cap - 1)))
      return bail(fail(MCB200_ERR_INVALID, "world_create: ring_cap must be a power of two >= 32"));
    w->ring_cap = (unsigned)d->ring_cap;
  } else {
    // twice a chain's share of the histories in flight (births are throttled per chain, so a
    // ring cannot be filled by its own chain: senders practically never lose lanes to a full ring)
    const unsigned chain = pow2_ceil(2ull * (w->inflight_limit / (unsigned long long)w->S) + 1ull);
    const unsigned bulk = pow2_ceil((2u << 20) / (unsigned)w->S);
    unsigned c = chain > bulk ? chain : bulk;
    w->ring_cap = c < 64 ? 64 : c > 8192 ? 8192 : c;
  }
  // banks: 512 MB per rank in total, per CTA a power of two in [1024, 1M] records
  if (d->bank_cap > 0) {
    if (d->bank_cap < 32 || (d->bank_cap & (d->bank_cap - 1)) || d->bank_cap > (1ll << 30))
      return bail(fail(MCB200_ERR_INVALID, "world_create: bank_cap must be a power of two >= 32"));
    w->bank_cap = (unsigned)d->bank_cap;
  } else {
    unsigned c = pow2_floor((512ull << 20) / (24ull * (unsigned long long)w->cfg.grid));
    w->bank_cap = c < 1024 ? 1024 : c > (1u << 20) ? (1u << 20) : c;
  }
  if (w->bank_cap < 4u * (unsigned)w->cfg.block)
    return bail(fail(MCB200_ERR_INVALID, "world_create: bank_cap must be at least 4 x the CTA size"));
  w->bank_log2 = ilog2(w->bank_cap);
  w->retire_batch = d->retire_batch > 0 ? (d->retire_batch > 32 ? 32 : d->retire_batch)
                                        : (m_max_all / w->V < 256 ? 6 : m_max_all / w->V < 512 ? 4 : 2);
  int rc = alloc_device(w);
  if (rc) return bail(rc);
  *out = w;
  return MCB200_OK;
}

void mcb200_world_destroy(mcb200_world *w) {
  if (!w) return;
  free_device(w);
  delete w;
}

int mcb200_world_export(mcb200_world *w, uint8_t handle_out[MCB200_IPC_HANDLE_BYTES],
                        mcb200_world_geom *geom_out) {
  if (!w || !geom_out) return fail(MCB200_ERR_INVALID, "world_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == MCB200_IPC_HANDLE_BYTES, "IPC handle size");
  DeviceGuard g(w->device);
  *geom_out = w->geom;
  if (handle_out) {
    cudaIpcMemHandle_t h;
    MCB_CUDA(cudaIpcGetMemHandle(&h, w->d_xblock));
    std::memcpy(handle_out, &h, sizeof h);
  }
  return MCB200_OK;
}

static int check_geom(mcb200_world *w, const mcb200_world_geom *q) {
  if (q->world_size != w->K || q->rank < 0 || q->rank >= w->K || q->rank == w->rank)
    return fail(MCB200_ERR_INVALID, "world_connect: the peer belongs to another world");
  if (q->stripes != w->S || q->ring_cap != (int32_t)w->ring_cap)
    return fail(MCB200_ERR_INVALID,
                "world_connect: ring geometry differs between ranks (stripes " +
                    std::to_string(q->stripes) + " vs " + std::to_string(w->S) + ", ring_cap " +
                    std::to_string(q->ring_cap) + " vs " + std::to_string(w->ring_cap) + ")");
  if (w->peers[(size_t)q->rank].base)
    return fail(MCB200_ERR_INVALID, "world_connect: rank already connected");
  return MCB200_OK;
}

int mcb200_world_connect_peer(mcb200_world *w, int32_t peer_rank,
                              const uint8_t handle[MCB200_IPC_HANDLE_BYTES],
                              const mcb200_world_geom *geom) {
  if (!w || !handle || !geom || geom->rank != peer_rank)
    return fail(MCB200_ERR_INVALID, "world_connect_peer: bad argument");
  int rc = check_geom(w, geom);
  if (rc) return rc;
  DeviceGuard g(w->device);
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof h);
  void *p = nullptr;
  MCB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  mcb200_world::Peer &pr = w->peers[(size_t)peer_rank];
  pr.base = static_cast<unsigned char *>(p);
  pr.ipc = true;
  pr.geom = *geom;
  w->table_dirty = true;
  return MCB200_OK;
}

int mcb200_world_connect_local(mcb200_world *w, mcb200_world *peer) {
  if (!w || !peer || w == peer) return fail(MCB200_ERR_INVALID, "world_connect_local: bad argument");
  int rc = check_geom(w, &peer->geom);
  if (rc) return rc;
  if (peer->device != w->device) {
    DeviceGuard g(w->device);
    int can = 0;
    MCB_CUDA(cudaDeviceCanAccessPeer(&can, w->device, peer->device));
    if (!can) return fail(MCB200_ERR_CUDA, "world_connect_local: no peer access between the devices");
    cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
      return fail(MCB200_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
    cudaGetLastError();
  }
  mcb200_world::Peer &pr = w->peers[(size_t)peer->rank];
  pr.base = peer->d_xblock;
  pr.ipc = false;
  pr.geom = peer->geom;
  w->table_dirty = true;
  return MCB200_OK;
}

int mcb200_world_disconnect(mcb200_world *w) {
  if (!w) return fail(MCB200_ERR_INVALID, "world_disconnect: null world");
  DeviceGuard g(w->device);
  if (w->stream) cudaStreamSynchronize(w->stream);
  for (int k = 0; k < w->K; ++k) {
    if (k == w->rank) continue;
    mcb200_world::Peer &pr = w->peers[(size_t)k];
    if (pr.base && pr.ipc) cudaIpcCloseMemHandle(pr.base);
    pr = mcb200_world::Peer{};
  }
  w->table_dirty = true;
  return MCB200_OK;
}

int mcb200_world_prepare(mcb200_world *w, int64_t nb_particles, uint64_t seed) {
  if (!w || nb_particles < 0) return fail(MCB200_ERR_INVALID, "world_prepare: bad argument");
  if (w->launched) return fail(MCB200_ERR_INVALID, "world_prepare: a run is in flight");
  DeviceGuard g(w->device);
  if (w->table_dirty) {
    int rc = upload_windows(w);
    if (rc) return rc;
  }
  // A run starts with zeroed rings and credits (a slot's lap parity is only meaningful
  // from a zeroed ring, and the kernels count from zero); the rings are empty between runs,
  // so nothing is lost.  All ranks do this BEFORE the barrier that precedes the launches.
  MCB_CUDA(cudaMemsetAsync(w->d_xblock, 0, (size_t)w->geom.block_bytes, w->stream));
  const InnerLayout il = inner_layout(w->V, w->S, w->cpw, w->ring_cap, w->bank_cap);
  if (il.banks_off > 0) MCB_CUDA(cudaMemsetAsync(w->d_inner, 0, il.banks_off, w->stream));
  // banks: head == tail between runs and the lap parity of every slot is consistent with
  // them, so they carry over; after a failed run everything is wiped
  if (w->bank_dirty) {
    MCB_CUDA(cudaMemsetAsync(w->d_inner + il.banks_off, 0, (size_t)w->V * il.bank_bytes, w->stream));
    w->bank_dirty = false;
  }
  MCB_CUDA(cudaMemsetAsync(w->d_ctr, 0, sizeof(mcb::WorldCounters), w->stream));
  MCB_CUDA(cudaStreamSynchronize(w->stream));
  w->nb_particles = nb_particles;
  w->seed0 = seed;
  w->prepared = true;
  return MCB200_OK;
}

int mcb200_world_launch(mcb200_world *w) {
  if (!w) return fail(MCB200_ERR_INVALID, "world_launch: null world");
  if (!w->prepared || w->launched)
    return fail(MCB200_ERR_INVALID, "world_launch: call prepare first (once per run)");
  DeviceGuard g(w->device);
  mcb::WorldParams p{};
  p.win = w->d_win;
  p.V = w->V;
  p.cpw = w->cpw;
  p.rank_lo = w->lo;
  p.xs = w->d_xs;
  for (int v = 0; v <= w->V; ++v) p.win_lo[v] = w->win_lo[(size_t)v];
  p.dx = w->dx;
  p.minw = w->minw;
  p.retire_batch = w->retire_batch;
  p.ring_cap = w->ring_cap;
  p.ring_log2 = ilog2(w->ring_cap);
  const bool is_home = w->rank == w->home_rank;
  p.src_window = -1;
  if (is_home && w->src_cell >= 0)
    for (int v = 0; v < w->V; ++v)
      if (w->src_cell >= w->win_lo[(size_t)v] && w->src_cell < w->win_lo[(size_t)v + 1]) p.src_window = v;
  p.src_index = w->src_index;
  p.src_total = p.src_window >= 0 ? (unsigned long long)w->nb_particles : 0ull;
  p.chain_state = w->seed0;
  p.rng = w->rng;
  p.rng_key = mcb::philox_key(w->seed0);
  p.x_ini = w->x_ini;
  p.wmc = (float)(1.0 / (double)w->nb_particles);   // src/layer.cpp:38
  p.inflight_limit = w->inflight_limit;
  // with x_ini outside the slab nothing is ever born (src/layer.cpp:73): the run is empty
  p.total = w->src_cell >= 0 ? (unsigned long long)w->nb_particles : 0ull;
  p.ctrl = reinterpret_cast<mcb::WorldCtrl *>(w->d_xblock);
  p.home_disabled =
      &reinterpret_cast<mcb::WorldCtrl *>(w->peers[(size_t)w->home_rank].base)->disabled_global;
  p.home_chain_disabled = reinterpret_cast<unsigned *>(w->peers[(size_t)w->home_rank].base +
                                                      w->peers[(size_t)w->home_rank].geom.off_chain);
  // a chain (stripe s of every window of every rank) gets its share of the histories in flight
  {
    const unsigned long long per = w->inflight_limit / (unsigned long long)w->S;
    p.chain_limit = per < 64ull ? 64u : per > 0x7fffffffull ? 0x7fffffffu : (unsigned)per;
  }
  p.done_ptrs = w->d_done_ptrs;
  p.n_ranks = w->K;
  p.is_home = is_home ? 1 : 0;
  p.max_run_ns = (unsigned long long)w->max_run_ms * 1000000ull;
  p.acc = w->d_acc;
  p.ncell_rank = w->ncell();
  p.ctr = w->d_ctr;
  MCB_CUDA(cudaEventRecord(w->ev0, w->stream));
  MCB_CUDA(mcb::launch_world(p, w->cfg, w->stream));
  MCB_CUDA(cudaEventRecord(w->ev1, w->stream));
  w->gpu_launches++;
  w->launched = true;
  w->prepared = false;
  return MCB200_OK;
}

int mcb200_world_wait(mcb200_world *w, mcb200_world_result *out) {
  if (!w) return fail(MCB200_ERR_INVALID, "world_wait: null world");
  if (!w->launched) return fail(MCB200_ERR_INVALID, "world_wait: nothing was launched");
  DeviceGuard g(w->device);
  // The kernel ends by itself when the global count reaches nb_particles.  The host only
  // watches: if the home rank's counters stop moving for `stall_ms` (a peer process died,
  // a kernel never became resident ...) it raises this rank's `done` so the GPU is freed.
  using clk = std::chrono::steady_clock;
  auto last_move = clk::now(), last_look = clk::now();
  unsigned long long seen[2] = {~0ull, ~0ull};
  bool stalled = false;
  mcb::WorldCtrl *home = reinterpret_cast<mcb::WorldCtrl *>(w->peers[(size_t)w->home_rank].base);
  mcb::WorldCtrl *mine = reinterpret_cast<mcb::WorldCtrl *>(w->d_xblock);
  for (;;) {
    cudaError_t q = cudaEventQuery(w->ev1);
    if (q == cudaSuccess) break;
    if (q != cudaErrorNotReady) {
      w->launched = false;
      return fail(MCB200_ERR_CUDA, std::string("world_wait: ") + cudaGetErrorString(q));
    }
    const auto now = clk::now();
    if (w->stall_ms > 0 && now - last_look > std::chrono::milliseconds(50)) {
      last_look = now;
      if (cudaMemcpyAsync(w->h_prog, home, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                          w->side) == cudaSuccess &&
          cudaStreamSynchronize(w->side) == cudaSuccess) {
        if (w->h_prog[0] != seen[0] || w->h_prog[1] != seen[1]) {
          seen[0] = w->h_prog[0];
          seen[1] = w->h_prog[1];
          last_move = now;
        } else if (!stalled && now - last_move > std::chrono::milliseconds(w->stall_ms)) {
          // stop this rank's kernel; an error it raised itself (a bank overflow ...) is kept
          unsigned flags[2] = {1u, 0u};   // done, error
          cudaMemcpyAsync(&flags[1], &mine->error, sizeof(unsigned), cudaMemcpyDeviceToHost, w->side);
          cudaStreamSynchronize(w->side);
          if (flags[1] == 0u) flags[1] = (unsigned)(-MCB200_ERR_TIMEOUT);
          cudaMemcpyAsync(&mine->done, flags, sizeof flags, cudaMemcpyHostToDevice, w->side);
          cudaStreamSynchronize(w->side);
          stalled = true;
        }
      } else {
        cudaGetLastError();
      }
    }
    std::this_thread::sleep_for(std::chrono::microseconds(100));
  }
  w->launched = false;
  float ms = 0.f;
  MCB_CUDA(cudaEventElapsedTime(&ms, w->ev0, w->ev1));
  mcb::WorldCtrl ctrl;
  MCB_CUDA(cudaMemcpyAsync(w->h_ctr, w->d_ctr, sizeof(mcb::WorldCounters), cudaMemcpyDeviceToHost,
                           w->stream));
  MCB_CUDA(cudaMemcpyAsync(&ctrl, mine, sizeof ctrl, cudaMemcpyDeviceToHost, w->stream));
  MCB_CUDA(cudaStreamSynchronize(w->stream));
  const mcb::WorldCounters &c = *w->h_ctr;
  int err = MCB200_OK;
  if (ctrl.error) err = -(int)ctrl.error;
  else if (c.acc_range) err = MCB200_ERR_RANGE;
  if (err) w->bank_dirty = true;
  if (out) {
    std::memset(out, 0, sizeof *out);
    out->events = (int64_t)c.events;
    out->scatters = (int64_t)c.scatters;
    out->n_left = (int64_t)c.n_cls[0];
    out->n_right = (int64_t)c.n_cls[1];
    out->n_dead = (int64_t)c.n_cls[2];
    out->births = (int64_t)c.births;
    out->sent_left = (int64_t)c.sent_outer[0];
    out->sent_right = (int64_t)c.sent_outer[1];
    out->window_crossings = (int64_t)(c.sent[0] + c.sent[1] - c.sent_outer[0] - c.sent_outer[1]);
    out->idle_polls = (int64_t)c.idle_polls;
    out->blocked_passes = (int64_t)c.blocked_passes;
    out->bank_pushes = (int64_t)c.bank_pushes;
    out->bank_pops = (int64_t)c.bank_pops;
    out->lane_slots = (int64_t)c.lane_slots;
    out->idle_warp_ns = (int64_t)c.idle_ns;
    double wc[3] = {0, 0, 0};
    int rc = fetch_world_tally(w, nullptr, wc);
    if (rc) return rc;
    out->w_left = wc[0];
    out->w_right = wc[1];
    out->w_dead = wc[2];
    out->kernel_ms = (double)ms;
    out->windows = w->V;
    out->ctas = w->cfg.grid;
    out->block = w->cfg.block;
    out->stripes = w->S;
    out->ring_cap = (int32_t)w->ring_cap;
    out->error = err;
  }
  if (err == MCB200_ERR_TIMEOUT)
    return fail(err, stalled ? "world_wait: the run made no progress for stall_ms and was stopped by the "
                               "host (a peer rank died, the kernels of the ranks never ran at the same "
                               "time, or a rank was prepared while a neighbour's previous run was still "
                               "ending)"
                             : "world_wait: the kernel hit the max_run_ms cap and stopped itself");
  if (err == MCB200_ERR_CAPACITY) return fail(err, "world_wait: a bank was corrupted (internal error)");
  if (err == MCB200_ERR_TIMEOUT && c.bank_full)
    return fail(err, "world_wait: the run stopped making progress with a full bank (raise bank_cap)");
  if (err == MCB200_ERR_RANGE) return fail(err, "a particle weight or deposit is outside (-2^7, 2^7)");
  if (err) return fail(err, "world_wait: the kernel reported an error");
  return MCB200_OK;
}

int mcb200_world_run(mcb200_world *const *worlds, int32_t n, int64_t nb_particles, uint64_t seed,
                     mcb200_world_result *results) {
  if (!worlds || n < 1) return fail(MCB200_ERR_INVALID, "world_run: bad argument");
  for (int i = 0; i < n; ++i)
    if (!worlds[i] || worlds[i]->K != n)
      return fail(MCB200_ERR_INVALID, "world_run: pass every rank of the world, in one process");
  for (int i = 0; i < n; ++i) {
    int rc = mcb200_world_prepare(worlds[i], nb_particles, seed);
    if (rc) return rc;
  }
  int first = MCB200_OK;
  std::string msg;
  int launched = 0;
  for (int i = 0; i < n; ++i) {
    int rc = mcb200_world_launch(worlds[i]);
    if (rc) {
      first = rc;
      msg = mcb::get_last_error();
      break;
    }
    ++launched;
  }
  if (first != MCB200_OK) {
    // free the kernels already resident: raise their `done`
    for (int i = 0; i < launched; ++i) {
      DeviceGuard g(worlds[i]->device);
      const unsigned one = 1u;
      cudaMemcpyAsync(&reinterpret_cast<mcb::WorldCtrl *>(worlds[i]->d_xblock)->done, &one, sizeof one,
                      cudaMemcpyHostToDevice, worlds[i]->side);
      cudaStreamSynchronize(worlds[i]->side);
    }
  }
  for (int i = 0; i < launched; ++i) {
    int rc = mcb200_world_wait(worlds[i], results ? &results[i] : nullptr);
    if (rc && first == MCB200_OK) {
      first = rc;
      msg = mcb::get_last_error();
    }
  }
  if (first != MCB200_OK) return fail(first, msg);
  return MCB200_OK;
}

int mcb200_world_cells(mcb200_world *w, int32_t *lo, int32_t *m) {
  if (!w) return fail(MCB200_ERR_INVALID, "world_cells: null world");
  if (lo) *lo = w->lo;
  if (m) *m = w->M;
  return MCB200_OK;
}

int mcb200_world_tally_exact(mcb200_world *w, uint32_t *out_4m, int32_t *lsb_log2) {
  if (!w || !out_4m) return fail(MCB200_ERR_INVALID, "world_tally_exact: null argument");
  std::vector<unsigned> d;
  int rc = fetch_world_tally(w, &d, nullptr);
  if (rc) return rc;
  std::memcpy(out_4m, d.data(), d.size() * sizeof(unsigned));
  if (lsb_log2) *lsb_log2 = mcb::kAccLsbLog2;
  return MCB200_OK;
}

int mcb200_world_tally_f64(mcb200_world *w, double *out_m) {
  if (!w || !out_m) return fail(MCB200_ERR_INVALID, "world_tally_f64: null argument");
  std::vector<unsigned> d;
  int rc = fetch_world_tally(w, &d, nullptr);
  if (rc) return rc;
  for (int i = 0; i < w->M; ++i) out_m[i] = mcb::acc_to_double(&d[(size_t)i * mcb::kAccDigits]);
  return MCB200_OK;
}

int mcb200_world_tally(mcb200_world *w, float *out_m) {
  if (!w || !out_m) return fail(MCB200_ERR_INVALID, "world_tally: null argument");
  std::vector<unsigned> d;
  int rc = fetch_world_tally(w, &d, nullptr);
  if (rc) return rc;
  for (int i = 0; i < w->M; ++i)
    out_m[i] = (float)mcb::acc_to_double(&d[(size_t)i * mcb::kAccDigits]);
  return MCB200_OK;
}

int mcb200_world_reset_tally(mcb200_world *w) {
  if (!w) return fail(MCB200_ERR_INVALID, "world_reset_tally: null world");
  if (w->launched) return fail(MCB200_ERR_INVALID, "world_reset_tally: a run is in flight");
  DeviceGuard g(w->device);
  MCB_CUDA(cudaMemsetAsync(w->d_acc, 0, 2 * (size_t)w->ncell() * sizeof(unsigned long long), w->stream));
  MCB_CUDA(cudaStreamSynchronize(w->stream));
  return MCB200_OK;
}

int mcb200_world_gather_tally_f64(mcb200_world *const *worlds, int32_t n, double *out) {
  if (!worlds || n < 1 || !out) return fail(MCB200_ERR_INVALID, "world_gather_tally: bad argument");
  for (int i = 0; i < n; ++i) {
    if (!worlds[i]) return fail(MCB200_ERR_INVALID, "world_gather_tally: null world");
    int rc = mcb200_world_tally_f64(worlds[i], out + worlds[i]->lo);
    if (rc) return rc;
  }
  return MCB200_OK;
}

int mcb200_world_set_option(mcb200_world *w, const char *key, int64_t value) {
  if (!w || !key) return fail(MCB200_ERR_INVALID, "world_set_option: null argument");
  const std::string k(key);
  if (k == "max_run_ms") w->max_run_ms = value;
  else if (k == "stall_ms") w->stall_ms = value;
  else if (k == "rng" && (value == 0 || value == 1)) w->rng = (int)value;
  else if (k == "retire_batch") w->retire_batch = value < 1 ? 1 : value > 32 ? 32 : (int)value;
  else if (k == "inflight_limit" && value > 0) w->inflight_limit = (unsigned long long)value;
  else return fail(MCB200_ERR_INVALID, "world_set_option: unknown key " + k);
  return MCB200_OK;
}

void *mcb200_world_stream(mcb200_world *w) { return w ? (void *)w->stream : nullptr; }

}  // extern "C"
