// mcb_kernels.cu -- hand-written sm_100a kernels of the particle-tracking path.
//
// track_kernel<SHARED, MAXB> is the product: a persistent, history-based
// tracking kernel.  Each lane owns one particle in registers and runs the
// reference's event loop (src/layer.cpp:123-218) until the particle leaves the
// sub-slab or drops below particle_min_weight; finished lanes are retired and
// refilled from the bank inside the loop, so a warp keeps 32 live histories
// regardless of how unequal their lengths are (293 vs 708 events on the
// default slab).  Per event:
//    * 1-2 steps of the per-particle 64-bit LCG (registers only),
//    * glibc-exact logf / expf in FP64, two IEEE fp32 divides,
//    * one EXACT deposit into the CTA-private per-cell tally in shared memory
//      (128-bit long accumulator in 32-bit digits: 1-2 native ATOMS.ADD plus
//      rare carries), merged into the global tally ONCE per CTA,
//    * escapees compacted with ballot/popc prefix sums into per-CTA outbox stripes
//      (one SHARED-memory atomic per warp per retire event, no global atomic), packed
//      into contiguous left / right send buffers by gather_stripes_kernel.
// Nothing on this path is a contraction: no tensor cores by design.
#include "mcb_kernels.cuh"
#include "mcb_event.cuh"

namespace mcb {

struct TrackSmem {
  MathTables math;
  unsigned int n_cls[3];
  unsigned int pad;
  unsigned int out_n[2];   // fill of this CTA's outbox stripes
  unsigned int pad2[2];
};

// ----------------------------------------------------------- the hot path --

// MAXB = largest CTA this instantiation is launched with.  The 256-thread one is the
// workhorse: 4 CTAs per SM (1024 resident threads, <= 64 registers).  Measured on B200
// (tools/quick_variants.py): a 40-register cap (6 CTAs/SM) makes ptxas re-materialise
// addresses and constants inside the event loop and is 5 % slower (1.80e11 vs 1.89e11
// events/s); the kernel is issue-bound with ~5 eligible warps per issue slot, so the extra
// occupancy buys nothing.  The 1024-thread one exists for sub-slabs whose CTA-private tally
// only fits once per SM.
template <bool SHARED, int MAXB, int RNG>
__global__ void __launch_bounds__(MAXB, MAXB <= 256 ? 4 : 1) track_kernel(const TrackParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TrackSmem *sm = reinterpret_cast<TrackSmem *>(smem_raw);
  CellXs *s_xs = reinterpret_cast<CellXs *>(smem_raw + sizeof(TrackSmem));
  const int ncell = p.m + kAccExtra;
  // CTA-private copy of the tally (digit-major) when it fits in shared memory,
  // else the L2-resident global one
  unsigned *acc = reinterpret_cast<unsigned *>(s_xs + p.m);   // only touched when SHARED
  unsigned long long *gacc = reinterpret_cast<unsigned long long *>(p.acc);

  load_math_tables(&sm->math);
  if (threadIdx.x < 3) sm->n_cls[threadIdx.x] = 0u;
  if (threadIdx.x < 2) sm->out_n[threadIdx.x] = 0u;
  if (SHARED) {
    for (int c = threadIdx.x; c < p.m; c += blockDim.x) s_xs[c] = p.xs[c];
    for (int c = threadIdx.x; c < kAccDigits * ncell; c += blockDim.x) acc[c] = 0u;
  }
  __syncthreads();

  // 32-bit shared-window addresses of the three shared structures the event touches
  const unsigned tb_s = smem_addr(&sm->math);
  const unsigned xs_s = smem_addr(s_xs);
  const unsigned acc_s = smem_addr(s_xs + p.m);
  const unsigned acc_stride = (unsigned)ncell * 4u;  // bytes between digits

  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int lo = p.idx_lo;
  const int hi = p.idx_lo + p.m;
  const float dx = p.dx, minw = p.minw;
  const int retire_batch = p.retire_batch;
  const unsigned long long take = (unsigned long long)p.take_count;

  // particle state, include/types/particle.hpp:7-18, one history per lane
  unsigned long long seed = 0;
  float x = 0.f, mu = 0.f, wmc = 0.f;
  float rmu = 0.f;  // recip_for_div(mu), refreshed whenever mu changes
  int step = 1;     // dir_step(mu), likewise
  int idx = 0;
  bool active = false;
  bool exhausted = false;  // warp-uniform: the bank range has been handed out
  bool drained = false;    // warp-uniform: the global cursor has passed the end of the range
  unsigned long long w_next = 0ull, w_end = 0ull;  // this warp's chunk of bank slots
  unsigned n_ev = 0, n_sc = 0;

  for (;;) {
    // ---- liveness: loop condition of simulate_particle, src/layer.cpp:195-197
    const bool alive = active && (wmc >= minw) && ((unsigned)(idx - lo) < (unsigned)p.m);
    // steady state: (almost) all 32 lanes carry a live history -> one vote, straight to the
    // event.  Retire / refill / exit run only once `retire_batch` lanes are without a live
    // history (or, at the tail, on every iteration), so that their cost is shared.
    const unsigned nolive = __ballot_sync(MCB_FULL, !alive);
    if (nolive != 0u && (__popc(nolive) >= retire_batch || exhausted)) {
    const bool fin = active && !alive;
    const unsigned fm = __ballot_sync(MCB_FULL, fin);
    if (fm) {
      // ---- retire: classification of src/layer.cpp:202-217, routing of
      // :332-346, global-border absorption of :350-360
      const int cls = (idx == lo - 1) ? 0 : (idx == hi) ? 1 : (wmc < minw) ? 2 : 0;
      if (fin) {
        if (SHARED) acc_add_smem(acc_s + (unsigned)(p.m + cls) * 4u, acc_stride, wmc,
                                 &p.ctr->acc_range);
        else gacc_add(&gacc[p.m + cls], ncell, wmc, &p.ctr->acc_range);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const bool mine = fin && cls == c;
        const unsigned cm = __ballot_sync(MCB_FULL, mine);
        if (cm == 0u) continue;
        const int cnt = __popc(cm);
        const int leader = __ffs(cm) - 1;
        if (lane == leader) atomicAdd(&sm->n_cls[c], (unsigned)cnt);
        if (c < 2 && p.write_side[c]) {
          // warp-aggregated append to this CTA's stripe: ONE shared-memory atomic per warp,
          // ranks by popc; the records are the 24-byte wire format
          unsigned base = 0u;
          if (lane == leader) base = atomicAdd(&sm->out_n[c], (unsigned)cnt);
          base = __shfl_sync(MCB_FULL, base, leader);
          if (mine) {
            const unsigned my = base + (unsigned)__popc(cm & lt_mask);
            long long slot = -1;
            if (my < (unsigned)p.stripe_cap[c]) {
              slot = (long long)blockIdx.x * p.stripe_cap[c] + my;
            } else {  // stripe full (unbalanced CTAs): the common overflow segment
              const unsigned o = atomicAdd(&p.fills[c][p.ovf_slot[c]], 1u);
              if ((long long)o < p.ovf_cap[c]) slot = p.ovf_base[c] + o;
              else atomicExch(&p.ctr->overflow, 1u);
            }
            if (slot >= 0) {
              // with a connected peer this is a store into the neighbour GPU's memory
              unsigned long long *rec = p.out_rec[c] + 3 * slot;
              rec[0] = seed;
              rec[1] = (unsigned long long)__float_as_uint(x) |
                       ((unsigned long long)__float_as_uint(mu) << 32);
              rec[2] = (unsigned long long)__float_as_uint(wmc) |
                       ((unsigned long long)(unsigned)idx << 32);
            }
          }
        }
      }
      active = active && !fin;
    }

    // ---- refill idle lanes from the bank (coalesced 8 B + 16 B per lane)
    const unsigned im = __ballot_sync(MCB_FULL, !active);
    if (im && !exhausted) {
      // bank slots come in warp-private chunks: one global atomic per kWorkChunk particles
      if (w_next == w_end && !drained) {
        unsigned long long base = 0ull;
        if (lane == 0) base = atomicAdd(&p.ctr->cursor, (unsigned long long)kWorkChunk);
        base = __shfl_sync(MCB_FULL, base, 0);
        w_next = base < take ? base : take;
        w_end = base + kWorkChunk < take ? base + kWorkChunk : take;
        drained = base + kWorkChunk >= take;
      }
      const unsigned long long avail = w_end - w_next;
      const unsigned long long cnt = (unsigned long long)__popc(im);
      if (!active) {
        const unsigned long long r = (unsigned long long)__popc(im & lt_mask);
        if (r < avail) {
          const long long slot = p.take_base + (long long)(w_next + r);
          seed = __ldcs(&p.bank_seed[slot]);
          const float4 st = __ldcs(&p.bank_st[slot]);
          x = st.x;
          mu = st.y;
          rmu = recip_for_div(mu);
          step = dir_step(mu);
          wmc = st.z;
          idx = __float_as_int(st.w);
          active = true;
        }
      }
      w_next += cnt < avail ? cnt : avail;
      if (w_next == w_end && drained) exhausted = true;
      continue;  // fresh lanes go through the liveness test first
    }
    if (im == MCB_FULL) break;  // bank handed out and every lane retired
    }

    // ---- one event per live lane: Layer::particle_step, src/layer.cpp:123-190
    if (alive) {
      event_step<SHARED, SHARED, RNG>(seed, x, mu, wmc, rmu, step, idx, n_sc, lo, dx, tb_s, xs_s, acc_s,
                                      acc_stride, p.xs, gacc, ncell, &p.ctr->acc_range, p.rng_key);
      ++n_ev;
      // a second event under the same vote for the lanes that are still live: the loop top
      // (vote, count, branches) is shared by two events; a lane that finished on the first one
      // waits one slot longer for its retirement
      if ((wmc >= minw) && ((unsigned)(idx - lo) < (unsigned)p.m)) {
        event_step<SHARED, SHARED, RNG>(seed, x, mu, wmc, rmu, step, idx, n_sc, lo, dx, tb_s, xs_s,
                                        acc_s, acc_stride, p.xs, gacc, ncell, &p.ctr->acc_range,
                                        p.rng_key);
        ++n_ev;
      }
    }
  }

  // ---- per-CTA flush: counters once, the CTA-private tally merged digit-wise
  unsigned long long ev = n_ev, sc = n_sc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ev += __shfl_xor_sync(MCB_FULL, ev, o);
    sc += __shfl_xor_sync(MCB_FULL, sc, o);
  }
  if (lane == 0) {
    if (ev) atomicAdd(&p.ctr->events, ev);
    if (sc) atomicAdd(&p.ctr->scatters, sc);
  }
  __syncthreads();
  if (SHARED) {
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) {
      unsigned d[kAccDigits];
      unsigned any = 0u;
#pragma unroll
      for (int j = 0; j < kAccDigits; ++j) {
        d[j] = acc[j * ncell + c];
        any |= d[j];
      }
      if (any) gacc_merge(&gacc[c], ncell, d);
    }
  }
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    if (sm->n_cls[c]) atomicAdd(&p.ctr->n_cls[c], (unsigned long long)sm->n_cls[c]);
  }
  if (threadIdx.x < 2 && p.write_side[threadIdx.x]) {
    const int c = threadIdx.x;
    const unsigned n = sm->out_n[c];
    p.fills[c][blockIdx.x] = n < (unsigned)p.stripe_cap[c] ? n : (unsigned)p.stripe_cap[c];
  }
}

size_t track_smem_bytes(int tally_mode, int m) {
  size_t b = sizeof(TrackSmem);
  if (tally_mode == kTallyShared)
    b += (size_t)m * sizeof(CellXs) + (size_t)(m + kAccExtra) * kAccDigits * sizeof(unsigned);
  return b;
}

typedef void (*TrackFn)(const TrackParams);
static TrackFn track_fn(int mode, int block, int rng) {
  if (rng == 0) {
    if (block <= 256) return mode == kTallyShared ? track_kernel<true, 256, 0> : track_kernel<false, 256, 0>;
    return mode == kTallyShared ? track_kernel<true, 1024, 0> : track_kernel<false, 1024, 0>;
  }
  if (block <= 256) return mode == kTallyShared ? track_kernel<true, 256, 1> : track_kernel<false, 256, 1>;
  return mode == kTallyShared ? track_kernel<true, 1024, 1> : track_kernel<false, 1024, 1>;
}

cudaError_t track_configure(int device, int m, int want_mode, int want_block,
                            int want_blocks_per_sm, int rng, TrackLaunch *out) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return e;
  const size_t per_sm = prop.sharedMemPerMultiprocessor;   // 228 KB on B200
  const size_t per_cta_max = prop.sharedMemPerBlockOptin;  // 227 KB
  const size_t reserve = 1024;                             // driver reserve per CTA

  int mode = want_mode;
  if (mode != kTallyShared && mode != kTallyGlobal)
    mode = track_smem_bytes(kTallyShared, m) <= per_cta_max ? kTallyShared : kTallyGlobal;
  if (mode == kTallyShared && track_smem_bytes(kTallyShared, m) > per_cta_max)
    return cudaErrorInvalidValue;
  const size_t smem = track_smem_bytes(mode, m);

  // CTA shape: as many resident threads as the register file allows, in CTAs
  // small enough that their private tally copies fit side by side
  int block = want_block > 0 ? want_block : 256;
  int bps = want_blocks_per_sm > 0 ? want_blocks_per_sm : 1024 / block;
  if (want_block <= 0 && want_blocks_per_sm <= 0) {
    while (bps > 1 && (smem + reserve) * (size_t)bps > per_sm) {
      bps /= 2;
      if (block < 1024) block *= 2;
    }
  } else {
    while (bps > 1 && (smem + reserve) * (size_t)bps > per_sm) --bps;
  }
  if (block > 1024 || block % 32) return cudaErrorInvalidValue;

  out->rng = rng ? 1 : 0;
  out->tally_mode = mode;
  out->block = block;
  out->grid = prop.multiProcessorCount * bps;
  if (out->grid > kStripes) out->grid = kStripes;
  out->smem = smem;
  return cudaFuncSetAttribute(track_fn(mode, block, out->rng),
                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

int track_grid(const TrackLaunch &cfg, long long take) {
  long long need = (take + cfg.block - 1) / cfg.block;
  if (need < 1) need = 1;
  return (int)(need < (long long)cfg.grid ? need : (long long)cfg.grid);
}

cudaError_t launch_track(const TrackParams &p, const TrackLaunch &cfg, cudaStream_t stream) {
  if (p.take_count <= 0) return cudaSuccess;
  track_fn(cfg.tally_mode, cfg.block, cfg.rng)<<<track_grid(cfg, p.take_count), cfg.block, cfg.smem,
                                                 stream>>>(p);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- gather --

// CTA s packs segment s (stripe s, or the overflow segment for s == nstripes) behind the
// segments before it: an exclusive prefix over <= kStripes + 1 fills (block reduction),
// then a linear, coalesced copy of 8-byte words.
// `fills[s]` = records in stripe s, `fills[nstripes]` = records in the overflow segment.
// TO_BANK = false: copy the wire records as they are (outbox); true: unpack them into the
// bank layout (inbox of a direct peer exchange).
template <bool TO_BANK>
__global__ void __launch_bounds__(256) gather_stripes_kernel(
    const unsigned long long *__restrict__ scratch, const unsigned *__restrict__ fills,
    int nstripes, int stripe_cap, long long ovf_base, long long ovf_cap,
    unsigned long long *__restrict__ dst_rec, float4 *__restrict__ dst_st, long long dst_n,
    unsigned long long *out_total) {
  __shared__ unsigned long long s_part[8];
  const int seg = blockIdx.x;
  // a fill counter is never trusted beyond the segment it counts (a sender with another
  // geometry, a corrupted table): what is copied always lies inside the stripes
  const unsigned cap_ovf = ovf_cap < 0xffffffffll ? (unsigned)ovf_cap : 0xffffffffu;
  unsigned long long before = 0ull;
  for (int t = threadIdx.x; t < seg; t += blockDim.x) before += min(fills[t], (unsigned)stripe_cap);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(MCB_FULL, before, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = before;
  __syncthreads();
  before = 0ull;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) before += s_part[w];
  const unsigned n = min(fills[seg], seg == nstripes ? cap_ovf : (unsigned)stripe_cap);
  const long long src = seg == nstripes ? ovf_base : (long long)seg * stripe_cap;
  const unsigned long long *from = scratch + 3 * src;
  if (!TO_BANK) {
    unsigned long long *to = dst_rec + 3 * (dst_n + (long long)before);
    for (unsigned i = threadIdx.x; i < 3u * n; i += blockDim.x) to[i] = from[i];
  } else {
    const long long base = dst_n + (long long)before;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long a = from[3 * i], b = from[3 * i + 1], c = from[3 * i + 2];
      dst_rec[base + i] = a;  // the bank's seed array
      dst_st[base + i] =
          make_float4(__uint_as_float((unsigned)b), __uint_as_float((unsigned)(b >> 32)),
                      __uint_as_float((unsigned)c), __uint_as_float((unsigned)(c >> 32)));
    }
  }
  if (seg == nstripes && threadIdx.x == 0) *out_total = before + n;
}

cudaError_t launch_gather_stripes(const unsigned long long *scratch, const unsigned *fills,
                                  int nstripes, int stripe_cap, long long ovf_base, long long ovf_cap,
                                  unsigned long long *settled, long long settled_n,
                                  unsigned long long *out_total, cudaStream_t stream) {
  gather_stripes_kernel<false><<<nstripes + 1, 256, 0, stream>>>(
      scratch, fills, nstripes, stripe_cap, ovf_base, ovf_cap, settled, nullptr, settled_n, out_total);
  return cudaGetLastError();
}

cudaError_t launch_gather_stripes_to_bank(const unsigned long long *scratch,
                                          const unsigned *fills, int nstripes, int stripe_cap,
                                          long long ovf_base, long long ovf_cap,
                                          unsigned long long *bank_seed, float4 *bank_st,
                                          long long bank_n, unsigned long long *out_total,
                                          cudaStream_t stream) {
  gather_stripes_kernel<true><<<nstripes + 1, 256, 0, stream>>>(
      scratch, fills, nstripes, stripe_cap, ovf_base, ovf_cap, bank_seed, bank_st, bank_n, out_total);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ birth --

// Layer::create_particles(int n), src/layer.cpp:89-121, in parallel: the seed
// chain state after k steps is an affine function of the start state, so
// thread t jumps straight to particle t (63-entry table of squarings) and then
// strides by the grid size with one precomputed multiply-add.
__global__ void __launch_bounds__(256) birth_kernel(long long n, unsigned long long chain_state,
                                                    const JumpTable jt, Affine stride_map,
                                                    float x_ini, float wmc, int index,
                                                    unsigned long long *__restrict__ seed_out,
                                                    float4 *__restrict__ st_out) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  if (tid >= n) return;
  // particle `tid` carries rnd_seed^(tid+1)(chain_state)   (:111)
  unsigned long long s = jump_state(jt, (unsigned long long)tid + 1ull, chain_state);
  for (long long i = tid; i < n; i += nthreads) {
    unsigned long long ps = lcg_next(s);                               // :112 draw #1
    const float mu = __fsub_rn(__fmul_rn(2.0f, lcg_to_real(ps)), 1.0f);
    __stcs(&seed_out[i], ps);
    __stcs(&st_out[i], make_float4(x_ini, mu, wmc, __int_as_float(index)));
    s = affine_apply(stride_map, s);
  }
}

cudaError_t launch_birth(long long n, unsigned long long chain_state, const JumpTable &seed_jump,
                         float x_ini, float wmc, int index, unsigned long long *seed_out,
                         float4 *st_out, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int block = 256;
  long long want = (n + block - 1) / block;
  const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
  const Affine stride = jump_map(seed_jump, (unsigned long long)grid * block);
  birth_kernel<<<grid, block, 0, stream>>>(n, chain_state, seed_jump, stride, x_ini, wmc, index,
                                           seed_out, st_out);
  return cudaGetLastError();
}

// Philox mode: no chain -- history `first_id + i` IS its counter; event 0 draws the direction
__global__ void __launch_bounds__(256) birth_philox_kernel(long long n, unsigned long long first_id,
                                                           unsigned key, float x_ini, float wmc,
                                                           int index,
                                                           unsigned long long *__restrict__ seed_out,
                                                           float4 *__restrict__ st_out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long c0 = (first_id + (unsigned long long)i) << kPhiloxEventBits;
    const uint2 w = philox_draws(c0, key);
    const float mu = __fsub_rn(__fmul_rn(2.0f, u32_to_real(w.x)), 1.0f);
    __stcs(&seed_out[i], c0 + 1ull);
    __stcs(&st_out[i], make_float4(x_ini, mu, wmc, __int_as_float(index)));
  }
}

cudaError_t launch_birth_philox(long long n, unsigned long long first_id, unsigned key, float x_ini,
                                float wmc, int index, unsigned long long *seed_out, float4 *st_out,
                                cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int block = 256;
  long long want = (n + block - 1) / block;
  const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
  birth_philox_kernel<<<grid, block, 0, stream>>>(n, first_id, key, x_ini, wmc, index, seed_out, st_out);
  return cudaGetLastError();
}

__global__ void test_philox_kernel(long long n, const unsigned *c0, const unsigned *c1,
                                   const unsigned *key, uint2 *out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = philox2x32_10(c0[i], c1[i], key[i]);
}
cudaError_t launch_test_philox(long long n, const unsigned *c0, const unsigned *c1,
                               const unsigned *key, uint2 *out, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  test_philox_kernel<<<(int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024), 256, 0, stream>>>(n, c0, c1, key, out);
  return cudaGetLastError();
}

// ------------------------------------------------- wire format <-> bank --

// include/types/particle.hpp:7-18: {u64 seed; f32 x, mu, wmc; i32 index} = three
// 8-byte words per record.
__global__ void __launch_bounds__(256) aos_to_soa_kernel(long long n,
                                                         const unsigned long long *__restrict__ aos,
                                                         unsigned long long *__restrict__ seed,
                                                         float4 *__restrict__ st) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long a = __ldcs(&aos[3 * i]);
    const unsigned long long b = __ldcs(&aos[3 * i + 1]);
    const unsigned long long c = __ldcs(&aos[3 * i + 2]);
    seed[i] = a;
    st[i] = make_float4(__uint_as_float((unsigned)b), __uint_as_float((unsigned)(b >> 32)),
                        __uint_as_float((unsigned)c), __uint_as_float((unsigned)(c >> 32)));
  }
}

__global__ void __launch_bounds__(256) soa_to_aos_kernel(long long n,
                                                         const unsigned long long *__restrict__ seed,
                                                         const float4 *__restrict__ st,
                                                         unsigned long long *__restrict__ aos) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long a = seed[i];
    const float4 s = st[i];
    aos[3 * i] = a;
    aos[3 * i + 1] = (unsigned long long)__float_as_uint(s.x) |
                     ((unsigned long long)__float_as_uint(s.y) << 32);
    aos[3 * i + 2] = (unsigned long long)__float_as_uint(s.z) |
                     ((unsigned long long)__float_as_uint(s.w) << 32);
  }
}

static int stream_grid(long long n, int block) {
  long long want = (n + block - 1) / block;
  return (int)(want < 148 * 8 ? (want > 0 ? want : 1) : 148 * 8);
}

cudaError_t launch_aos_to_soa(long long n, const void *aos, unsigned long long *seed, float4 *st,
                              cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  aos_to_soa_kernel<<<stream_grid(n, 256), 256, 0, stream>>>(
      n, static_cast<const unsigned long long *>(aos), seed, st);
  return cudaGetLastError();
}

cudaError_t launch_soa_to_aos(long long n, const unsigned long long *seed, const float4 *st,
                              void *aos, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  soa_to_aos_kernel<<<stream_grid(n, 256), 256, 0, stream>>>(
      n, seed, st, static_cast<unsigned long long *>(aos));
  return cudaGetLastError();
}

// ---------------------------------------------------- known-answer kernels --

// one rnd_real draw per element: the device counterpart of the reference's
// rnd_real_kernel (src/curandom.cu:7-14)
__global__ void test_rnd_real_kernel(long long n, unsigned long long *seeds, float *out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long s = lcg_next(seeds[i]);
    seeds[i] = s;
    out[i] = lcg_to_real(s);
  }
}

cudaError_t launch_test_rnd_real(long long n, unsigned long long *seeds, float *out,
                                 cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  test_rnd_real_kernel<<<stream_grid(n, 256), 256, 0, stream>>>(n, seeds, out);
  return cudaGetLastError();
}

__global__ void test_math_kernel(int which, long long n, const float *in, float *out) {
  __shared__ MathTables tb_smem;
  load_math_tables(&tb_smem);
  __syncthreads();
  const unsigned tb = smem_addr(&tb_smem);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = which == 0 ? logf_glibc(in[i], tb) : expf_glibc_nonpos(in[i], tb);
}

// the hoisted-reciprocal division exactly as the event uses it (guards and fallback included)
__global__ void test_div_kernel(long long n, const float *a, const float *b, float *out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float r = recip_for_div(b[i]);
    float q = MCB_MAXREAL;
    if (r != 0.0f && fabsf(a[i]) >= 0x1p-100f && fabsf(a[i]) < 0x1p100f) q = div_by_recip(a[i], b[i], r);
    else if (b[i] < -MCB_EPS || MCB_EPS < b[i]) q = __fdiv_rn(a[i], b[i]);
    out[i] = q;
  }
}

cudaError_t launch_test_div(long long n, const float *a, const float *b, float *out,
                            cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  test_div_kernel<<<stream_grid(n, 256), 256, 0, stream>>>(n, a, b, out);
  return cudaGetLastError();
}

cudaError_t launch_test_math(int which, long long n, const float *in, float *out,
                             cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  test_math_kernel<<<stream_grid(n, 256), 256, 0, stream>>>(which, n, in, out);
  return cudaGetLastError();
}

// exact accumulation of n floats into one accumulator (test hook of acc_add):
// a CTA-private copy in shared memory, merged like the tracking kernel does
__global__ void test_accumulate_kernel(long long n, const float *in, unsigned *acc4,
                                       unsigned *range_flag) {
  __shared__ unsigned s_acc[kAccDigits];
  if (threadIdx.x < kAccDigits) s_acc[threadIdx.x] = 0u;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    acc_add_smem(smem_addr(s_acc), 4u, in[i], range_flag);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned d[kAccDigits];
    for (int j = 0; j < kAccDigits; ++j) d[j] = s_acc[j];
    gacc_merge(reinterpret_cast<unsigned long long *>(acc4), 1, d);
  }
}

cudaError_t launch_test_accumulate(long long n, const float *in, unsigned *acc4,
                                   cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  test_accumulate_kernel<<<stream_grid(n, 256), 256, 0, stream>>>(n, in, acc4,
                                                                 acc4 + kAccDigits);
  return cudaGetLastError();
}

}  // namespace mcb
