// mcb_world_kernel.cu -- the persistent exchange kernel (see mcb_world.cuh for the design).
//
// One launch per rank per run.  The event is the one of track_kernel (mcb_event.cuh:
// Layer::particle_step, src/layer.cpp:123-190, bit for bit); what differs is where lanes get
// their histories from and where finished ones go:
//
//   retire   left / right escapee -> the neighbouring window's ring (a global border absorbs,
//            src/layer.cpp:350-360), dead -> counted; every disabled history is added to the
//            home rank's device-side counter (per CTA, batched)
//   refill   inbound rings -> the window's bank -> births in the source window
//   idle     flush counts, home rank: counter == nb_particles -> raise every rank's `done`
#include "mcb_event.cuh"
#include "mcb_world.cuh"

#include "../../include/mcb200.h"

namespace mcb {

// rnd_seed jump-ahead table (src/random.cpp:20-29); uniform index -> constant-bank reads
__constant__ JumpTable c_seed_jump;

// ---- system-scope memory operations ---------------------------------------------------
// Ring counters, records and flags are read and written by kernels on DIFFERENT GPUs (peer
// mappings over NVLink) or different CTAs: every access is explicit about scope and order.
__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned *p, unsigned v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_add_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// per-warp exchange state (warp-uniform, so it lives in shared memory, not in registers)
struct WarpXchg {
  unsigned wr[2];       // records this warp has stored into its outbound stripe, per side
  unsigned cred[2];     // last credit value seen for that stripe
  unsigned rd[2];       // records this warp has taken from its inbound stripe, per side
  unsigned rd_pub[2];   // of which already returned to the producer as credit
};

struct WorldSmem {
  MathTables math;
  WindowDesc win;               // this CTA's window
  unsigned n_cls[3];            // absorbed left / right, dead
  unsigned pend;                // disabled histories not yet added to the home counter
  unsigned sent[2];
  unsigned sent_outer[2];
  unsigned births, idle_polls, blocked, bank_pushes, bank_pops, pad;
  unsigned long long busy_iters;
  unsigned long long t_start;   // globaltimer at kernel start (for the run-time cap)
  WarpXchg wx[kWorldMaxWarps];
};

// 24-byte wire record (include/types/particle.hpp:7-18) <-> lane registers
__device__ __forceinline__ void record_store(unsigned long long *rec, unsigned long long w0,
                                             float x, float mu, float wmc, int idx) {
  st_relaxed_sys(rec + 1, (unsigned long long)__float_as_uint(x) |
                              ((unsigned long long)__float_as_uint(mu) << 32));
  st_relaxed_sys(rec + 2, (unsigned long long)__float_as_uint(wmc) |
                              ((unsigned long long)(unsigned)idx << 32));
  st_relaxed_sys(rec, w0);
}

// every history this CTA disables is owed to the home rank's global counter; pay in batches
__device__ __forceinline__ void note_disabled(WorldSmem *sm, const WorldParams &p, unsigned cnt) {
  const unsigned old = atomicAdd(&sm->pend, cnt);
  if (((old + cnt) ^ old) >> 8) {
    const unsigned v = atomicExch(&sm->pend, 0u);
    if (v) red_add_sys(p.home_disabled, (unsigned long long)v);
  }
}

template <int MAXB>
__global__ void __launch_bounds__(MAXB, MAXB <= 256 ? 4 : 1) world_kernel(const WorldParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WorldSmem *sm = reinterpret_cast<WorldSmem *>(smem_raw);
  const int v = (int)blockIdx.x / p.cpw;             // window of this CTA
  const int cw = (int)blockIdx.x - v * p.cpw;        // CTA within the window
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wv = cw * (int)(blockDim.x >> 5) + warp;  // stripe of this warp

  load_math_tables(&sm->math);
  {
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&p.win[v]);
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(&sm->win);
    for (int i = threadIdx.x; i < (int)(sizeof(WindowDesc) / 8); i += blockDim.x) dst[i] = src[i];
  }
  if (threadIdx.x < 16) (&sm->n_cls[0])[threadIdx.x] = 0u;   // n_cls .. pad
  if (threadIdx.x == 0) {
    sm->busy_iters = 0ull;
    sm->t_start = global_timer_ns();
  }
  if (threadIdx.x < kWorldMaxWarps) {
    WarpXchg z{};
    sm->wx[threadIdx.x] = z;
  }
  __syncthreads();
  const int m = sm->win.m;
  const int lo = sm->win.idx_lo;
  const int hi = lo + m;
  const int ncell = m + kAccExtra;
  CellXs *s_xs = reinterpret_cast<CellXs *>(smem_raw + sizeof(WorldSmem));
  unsigned *acc = reinterpret_cast<unsigned *>(s_xs + m);
  {
    const CellXs *gx = sm->win.xs;
    for (int c = threadIdx.x; c < m; c += blockDim.x) s_xs[c] = gx[c];
    for (int c = threadIdx.x; c < kAccDigits * ncell; c += blockDim.x) acc[c] = 0u;
  }
  __syncthreads();

  const unsigned tb_s = smem_addr(&sm->math);
  const unsigned xs_s = smem_addr(s_xs);
  const unsigned acc_s = smem_addr(acc);
  const unsigned acc_stride = (unsigned)ncell * 4u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const float dx = p.dx, minw = p.minw;
  const int retire_batch = p.retire_batch;
  const unsigned cap = p.ring_cap;
  const bool is_src = v == p.src_window;
  WarpXchg *wx = &sm->wx[warp];

  // particle state, include/types/particle.hpp:7-18, one history per lane
  unsigned long long seed = 0;
  float x = 0.f, mu = 0.f, wmc = 0.f, rmu = 0.f;
  int idx = 0;
  bool active = false;
  bool src_done = !is_src;                        // warp-uniform: no more births to hand out
  unsigned long long w_next = 0ull, w_end = 0ull;  // this warp's chunk of source particles
  int cool = 0;       // iterations to wait before polling again for work that was not there
  unsigned n_ev = 0, n_sc = 0, n_it = 0;

  for (;;) {
    // ---- liveness: loop condition of simulate_particle, src/layer.cpp:195-197
    const bool alive = active && (wmc >= minw) && ((unsigned)(idx - lo) < (unsigned)m);
    const unsigned nolive = __ballot_sync(MCB_FULL, !alive);
    if (nolive != 0u) {
      const bool fin = active && !alive;
      const unsigned fm = __ballot_sync(MCB_FULL, fin);
      if (nolive == MCB_FULL || __popc(fm) >= retire_batch ||
          (__popc(nolive) >= retire_batch && --cool <= 0)) {
        // ================================================================ the pass ==
        // (0) credits: tell the producers what earlier passes took out of their stripes
        if (lane < 2) {
          const unsigned rd = wx->rd[lane];
          if (rd != wx->rd_pub[lane]) {
            st_relaxed_sys(sm->win.in[lane].credit + wv, rd);
            wx->rd_pub[lane] = rd;
          }
        }
        // (1) retire: classification of src/layer.cpp:202-217, routing of :332-346
        bool blocked = false;
        if (fm) {
          const int cls = (idx == lo - 1) ? 0 : (idx == hi) ? 1 : (wmc < minw) ? 2 : 0;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const bool mine = fin && cls == c;
            const unsigned cm = __ballot_sync(MCB_FULL, mine);
            if (cm == 0u) continue;
            const unsigned cnt = (unsigned)__popc(cm);
            if (sm->win.out[c].mode == 0) {
              // global border: absorbed and counted as disabled, src/layer.cpp:350-360
              if (mine) {
                acc_add_smem(acc_s + (unsigned)(m + c) * 4u, acc_stride, wmc, &p.ctr->acc_range);
                active = false;
              }
              if (lane == 0) {
                atomicAdd(&sm->n_cls[c], cnt);
                note_disabled(sm, p, cnt);
              }
            } else {
              // the neighbouring window's stripe `wv`: all-or-nothing per warp and side
              const unsigned wr = wx->wr[c];
              unsigned cred = wx->cred[c];
              if (wr + cnt - cred > cap) {
                cred = __shfl_sync(MCB_FULL, ld_relaxed_sys(sm->win.out[c].credit + wv), 0);
                if (lane == 0) wx->cred[c] = cred;
              }
              if (wr + cnt - cred > cap) {
                blocked = true;   // ring full: these lanes keep their escapee and retry
                continue;
              }
              if (mine) {
                const unsigned slot = (wr + (unsigned)__popc(cm & lt_mask)) & (cap - 1u);
                record_store(sm->win.out[c].rec + ((size_t)wv * cap + slot) * 3, seed, x, mu, wmc,
                             idx);
                active = false;
              }
              __syncwarp();
              if (lane == 0) {
                wx->wr[c] = wr + cnt;
                // release: the records of ALL lanes (ordered by the __syncwarp above) are
                // visible to whoever acquires this count
                st_release_sys(sm->win.out[c].wr_pub + wv, wr + cnt);
                atomicAdd(&sm->sent[c], cnt);
                if (sm->win.out[c].outer) atomicAdd(&sm->sent_outer[c], cnt);
              }
              __syncwarp();
            }
          }
          {
            const bool mine = fin && cls == 2;
            const unsigned cm = __ballot_sync(MCB_FULL, mine);
            if (cm) {
              if (mine) {
                acc_add_smem(acc_s + (unsigned)(m + 2) * 4u, acc_stride, wmc, &p.ctr->acc_range);
                active = false;
              }
              if (lane == 0) {
                atomicAdd(&sm->n_cls[2], (unsigned)__popc(cm));
                note_disabled(sm, p, (unsigned)__popc(cm));
              }
            }
          }
        }

        // (2) refill idle lanes: inbound rings first, then the bank, then births
        bool got = false;
        unsigned im = __ballot_sync(MCB_FULL, !active);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (im == 0u || !sm->win.in[c].present) continue;
          const unsigned rd = wx->rd[c];
          unsigned w = __shfl_sync(MCB_FULL, ld_relaxed_sys(sm->win.in[c].wr_pub + wv), 0);
          if (w == rd) continue;
          // acquire: pairs with the producer's release of this count
          w = __shfl_sync(MCB_FULL, ld_acquire_sys(sm->win.in[c].wr_pub + wv), 0);
          const unsigned avail = w - rd;
          const unsigned nidle = (unsigned)__popc(im);
          const unsigned take = avail < nidle ? avail : nidle;
          const unsigned r = (unsigned)__popc(im & lt_mask);
          if (!active && r < take) {
            const unsigned long long *rec =
                sm->win.in[c].rec + ((size_t)wv * cap + ((rd + r) & (cap - 1u))) * 3;
            const unsigned long long a = ld_relaxed_sys(rec), b = ld_relaxed_sys(rec + 1),
                                     d = ld_relaxed_sys(rec + 2);
            seed = a;
            x = __uint_as_float((unsigned)b);
            mu = __uint_as_float((unsigned)(b >> 32));
            wmc = __uint_as_float((unsigned)d);
            idx = (int)(unsigned)(d >> 32);
            rmu = recip_for_div(mu);
            active = true;
          }
          __syncwarp();
          if (lane == 0) wx->rd[c] = rd + take;
          __syncwarp();
          got = true;
          im = __ballot_sync(MCB_FULL, !active);
        }
        if (im != 0u) {
          // the window's bank (multi-consumer: claim with a CAS on the head)
          const BankQ bq = sm->win.bank;
          unsigned long long h = 0ull;
          unsigned n = 0u;
          if (lane == 0) {
            const unsigned nidle = (unsigned)__popc(im);
            for (;;) {
              h = ld_relaxed_sys(bq.ht);
              const unsigned long long t = ld_relaxed_sys(bq.ht + 1);
              if (t <= h) break;
              n = t - h < (unsigned long long)nidle ? (unsigned)(t - h) : nidle;
              if (atomicCAS(bq.ht, h, h + n) == h) break;
              n = 0u;
            }
          }
          h = __shfl_sync(MCB_FULL, h, 0);
          n = __shfl_sync(MCB_FULL, n, 0);
          if (n) {
            const unsigned r = (unsigned)__popc(im & lt_mask);
            if (!active && r < n) {
              const unsigned long long q = h + r;
              const unsigned long long *rec = bq.rec + (size_t)(q & (bq.cap - 1u)) * 3;
              const unsigned long long want = ((q >> bq.log2cap) + 1ull) & 1ull;
              unsigned long long a;
              unsigned spins = 0u;
              do {  // the pusher may still be writing this slot (it never waits on anyone)
                a = ld_acquire_sys(rec);
                if (++spins > (1u << 24)) {  // cannot happen unless the queue was corrupted
                  atomicExch(&p.ctrl->error, (unsigned)(-MCB200_ERR_CAPACITY));
                  break;
                }
              } while ((a >> 63) != want);
              const unsigned long long b = ld_relaxed_sys(rec + 1), d = ld_relaxed_sys(rec + 2);
              seed = a & kMask63;
              x = __uint_as_float((unsigned)b);
              mu = __uint_as_float((unsigned)(b >> 32));
              wmc = __uint_as_float((unsigned)d);
              idx = (int)(unsigned)(d >> 32);
              rmu = recip_for_div(mu);
              active = true;
            }
            if (lane == 0) atomicAdd(&sm->bank_pops, n);
            got = true;
            im = __ballot_sync(MCB_FULL, !active);
          }
        }
        if (im != 0u && !src_done) {
          // births, src/layer.cpp:101-120: particle i carries rnd_seed^(i+1)(chain_state) and
          // consumes the first draw of its own stream for mu.  Source particles are handed to
          // warps in chunks of kWorkChunk; a chunk is only claimed while the number of
          // histories in flight (born - disabled, both on this rank) is below the limit.
          if (w_next == w_end) {
            unsigned long long base = ~0ull;
            if (lane == 0) {
              const unsigned long long b = ld_relaxed_sys(&p.ctrl->born);
              if (b >= p.src_total) base = p.src_total;
              else if (b - ld_relaxed_sys(&p.ctrl->disabled_global) < p.inflight_limit)
                base = atomicAdd(&p.ctrl->born, (unsigned long long)kWorkChunk);
            }
            base = __shfl_sync(MCB_FULL, base, 0);
            if (base != ~0ull) {
              if (base >= p.src_total) src_done = true;
              else {
                w_next = base;
                w_end = base + kWorkChunk < p.src_total ? base + kWorkChunk : p.src_total;
              }
            }
          }
          const unsigned long long avail = w_end - w_next;
          const unsigned nidle = (unsigned)__popc(im);
          const unsigned n = avail < (unsigned long long)nidle ? (unsigned)avail : nidle;
          if (n) {
            const unsigned r = (unsigned)__popc(im & lt_mask);
            if (!active && r < n) {
              const unsigned long long s = jump_state(c_seed_jump, w_next + r + 1ull, p.chain_state);
              seed = lcg_next(s);                                              // :112 draw #1
              mu = __fsub_rn(__fmul_rn(2.0f, lcg_to_real(seed)), 1.0f);
              rmu = recip_for_div(mu);
              x = p.x_ini;
              wmc = p.wmc;
              idx = p.src_index;
              active = true;
            }
            w_next += n;
            if (lane == 0) atomicAdd(&sm->births, n);
            got = true;
            im = __ballot_sync(MCB_FULL, !active);
          }
        }

        // (3) blocked senders: make room for the neighbours by moving this warp's inbound
        // stripes into the bank -- two windows can then never wait on each other
        if (blocked) {
          if (lane == 0) atomicAdd(&sm->blocked, 1u);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (!sm->win.in[c].present) continue;
            const unsigned rd = wx->rd[c];
            const unsigned w = __shfl_sync(MCB_FULL, ld_acquire_sys(sm->win.in[c].wr_pub + wv), 0);
            const unsigned avail = w - rd;
            if (avail == 0u) continue;
            const unsigned take = avail < 32u ? avail : 32u;
            const BankQ bq = sm->win.bank;
            unsigned long long t = 0ull;
            if (lane == 0) {
              t = atomicAdd(bq.ht + 1, (unsigned long long)take);
              if (t + take - ld_relaxed_sys(bq.ht) > (unsigned long long)bq.cap)
                atomicExch(&p.ctrl->error, (unsigned)(-MCB200_ERR_CAPACITY));
              atomicAdd(&sm->bank_pushes, take);
            }
            t = __shfl_sync(MCB_FULL, t, 0);
            if ((unsigned)lane < take) {
              const unsigned long long *src =
                  sm->win.in[c].rec + ((size_t)wv * cap + ((rd + (unsigned)lane) & (cap - 1u))) * 3;
              const unsigned long long a = ld_relaxed_sys(src), b = ld_relaxed_sys(src + 1),
                                       d = ld_relaxed_sys(src + 2);
              const unsigned long long q = t + (unsigned long long)lane;
              unsigned long long *dst = bq.rec + (size_t)(q & (bq.cap - 1u)) * 3;
              st_relaxed_sys(dst + 1, b);
              st_relaxed_sys(dst + 2, d);
              // word 0 carries the lap parity in bit 63 (seeds are 63-bit): release = "valid"
              st_release_sys(dst, (a & kMask63) | ((((q >> bq.log2cap) + 1ull) & 1ull) << 63));
            }
            __syncwarp();
            if (lane == 0) wx->rd[c] = rd + take;
            __syncwarp();
          }
        }

        // (4) nothing to track (idle lanes, or escapees waiting for room in a full ring) and
        // nothing to fetch: bookkeeping, termination, back off
        const bool live_now = active && (wmc >= minw) && ((unsigned)(idx - lo) < (unsigned)m);
        if (__ballot_sync(MCB_FULL, live_now) == 0u && !got) {
          unsigned stop = 0u;
          if (lane == 0) {
            const unsigned owe = atomicExch(&sm->pend, 0u);
            if (owe) red_add_sys(p.home_disabled, (unsigned long long)owe);
            atomicAdd(&sm->idle_polls, 1u);
            if (p.is_home && ld_relaxed_sys(&p.ctrl->disabled_global) == p.total)
              for (int r = 0; r < p.n_ranks; ++r) st_relaxed_sys(p.done_ptrs[r], 1u);
            stop = ld_relaxed_sys(&p.ctrl->done);
            if (!stop && p.max_run_ns != 0ull && global_timer_ns() - sm->t_start > p.max_run_ns) {
              // the run-time cap: a peer died or the protocol is broken -- never hang the GPU
              atomicExch(&p.ctrl->error, (unsigned)(-MCB200_ERR_TIMEOUT));
              st_relaxed_sys(&p.ctrl->done, 1u);
              stop = 1u;
            }
          }
          if (__shfl_sync(MCB_FULL, stop, 0)) break;
          __nanosleep(500);
        }
        cool = (!got && __ballot_sync(MCB_FULL, !active) != 0u) ? 4 : 0;
        continue;  // fresh lanes go through the liveness test first
      }
    }

    // ---- one event per live lane: Layer::particle_step, src/layer.cpp:123-190
    if (alive) {
      event_step<true>(seed, x, mu, wmc, rmu, idx, n_sc, lo, dx, tb_s, xs_s, acc_s, acc_stride,
                       nullptr, nullptr, ncell, &p.ctr->acc_range);
      ++n_ev;
    }
    ++n_it;
  }

  // ---- per-CTA flush: counters once, the CTA-private tally merged into the rank's
  unsigned long long ev = n_ev, sc = n_sc, it = n_it;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ev += __shfl_xor_sync(MCB_FULL, ev, o);
    sc += __shfl_xor_sync(MCB_FULL, sc, o);
  }
  if (lane == 0) {
    if (ev) atomicAdd(&p.ctr->events, ev);
    if (sc) atomicAdd(&p.ctr->scatters, sc);
    if (it) atomicAdd(&sm->busy_iters, it);
  }
  __syncthreads();
  {
    unsigned long long *gacc = p.acc + sm->win.acc_off;
    for (int c = threadIdx.x; c < m; c += blockDim.x) {
      unsigned d[kAccDigits];
      unsigned any = 0u;
#pragma unroll
      for (int j = 0; j < kAccDigits; ++j) {
        d[j] = acc[j * ncell + c];
        any |= d[j];
      }
      if (any) gacc_merge(&gacc[c], p.ncell_rank, d);
    }
    // the three class accumulators (weight absorbed left / right, carried by the dead)
    if (threadIdx.x < kAccExtra) {
      const int c = m + threadIdx.x;
      unsigned d[kAccDigits];
      unsigned any = 0u;
#pragma unroll
      for (int j = 0; j < kAccDigits; ++j) {
        d[j] = acc[j * ncell + c];
        any |= d[j];
      }
      if (any) gacc_merge(&p.acc[p.ncell_rank - kAccExtra + threadIdx.x], p.ncell_rank, d);
    }
  }
  if (threadIdx.x == 0) {
    WorldCounters *g = p.ctr;
    for (int c = 0; c < 3; ++c)
      if (sm->n_cls[c]) atomicAdd(&g->n_cls[c], (unsigned long long)sm->n_cls[c]);
    for (int c = 0; c < 2; ++c) {
      if (sm->sent[c]) atomicAdd(&g->sent[c], (unsigned long long)sm->sent[c]);
      if (sm->sent_outer[c]) atomicAdd(&g->sent_outer[c], (unsigned long long)sm->sent_outer[c]);
    }
    if (sm->births) atomicAdd(&g->births, (unsigned long long)sm->births);
    if (sm->idle_polls) atomicAdd(&g->idle_polls, (unsigned long long)sm->idle_polls);
    if (sm->blocked) atomicAdd(&g->blocked_passes, (unsigned long long)sm->blocked);
    if (sm->bank_pushes) atomicAdd(&g->bank_pushes, (unsigned long long)sm->bank_pushes);
    if (sm->bank_pops) atomicAdd(&g->bank_pops, (unsigned long long)sm->bank_pops);
    if (sm->busy_iters) atomicAdd(&g->busy_iters, sm->busy_iters);
    // anything still owed to the home counter (only on an abnormal exit)
    const unsigned owe = atomicExch(&sm->pend, 0u);
    if (owe) red_add_sys(p.home_disabled, (unsigned long long)owe);
  }
}

size_t world_smem_bytes(int m_max) {
  return sizeof(WorldSmem) + (size_t)m_max * sizeof(CellXs) +
         (size_t)(m_max + kAccExtra) * kAccDigits * sizeof(unsigned);
}

typedef void (*WorldFn)(const WorldParams);
static WorldFn world_fn(int block) { return block <= 256 ? world_kernel<256> : world_kernel<1024>; }

cudaError_t world_configure(int device, int m_max, int block, WorldLaunch *out,
                            int *max_ctas_per_sm) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return e;
  const size_t smem = world_smem_bytes(m_max);
  if (smem > prop.sharedMemPerBlockOptin || block % 32 || block > 1024 || block < 32)
    return cudaErrorInvalidValue;
  e = cudaFuncSetAttribute(world_fn(block), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, world_fn(block), block, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidValue;
  out->block = block;
  out->grid = prop.multiProcessorCount * per_sm;   // every CTA resident: the kernel is persistent
  out->smem = smem;
  if (max_ctas_per_sm) *max_ctas_per_sm = per_sm;
  return cudaSuccess;
}

cudaError_t world_upload_jump_table(const JumpTable &jt) {
  return cudaMemcpyToSymbol(c_seed_jump, &jt, sizeof(JumpTable));
}

cudaError_t launch_world(const WorldParams &p, const WorldLaunch &cfg, cudaStream_t stream) {
  world_fn(cfg.block)<<<p.V * p.cpw, cfg.block, cfg.smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace mcb
