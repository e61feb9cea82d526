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
__device__ __forceinline__ void red_add_sys(unsigned *p, unsigned v) {
  asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// per-warp exchange state (warp-uniform, so it lives in shared memory, not in registers)
struct __align__(16) WarpXchg {
  unsigned wr[2];       // records this warp has stored into its outbound stripe, per side
  unsigned cred[2];     // last credit value seen for that stripe
  unsigned rd[2];       // records this warp has taken from its inbound stripe, per side
  // pass-only state kept out of the registers of the event loop
  unsigned w_off, w_cnt;   // this warp's chunk of source particles: next, size
  unsigned backoff;        // iterations to wait after a pass that found nothing (doubles to 16)
  unsigned nap;            // ns an idle warp sleeps between looks (doubles to 2 us)
  unsigned src_done;       // no (more) births to hand out
  unsigned born;           // histories this warp gave birth to (its chain's intake)
  unsigned pad[4];
};

struct WorldSmem {
  MathTables math;
  WindowDesc win;               // this CTA's window
  unsigned n_cls[3];            // absorbed left / right, dead
  unsigned pend;                // disabled histories not yet added to the home counter
  unsigned births, idle_polls, blocked, bank_pushes, bank_pops;
  unsigned bank_head, bank_tail;   // this CTA's bank: next to pop / next to push
  unsigned acc_range;              // a deposit did not fit the accumulator (copied out at exit)
  unsigned long long t_start;   // globaltimer at kernel start (for the run-time cap)
  unsigned long long idle_ns;   // time its warps spent without a single live history
  WarpXchg wx[kWorldMaxWarps];
};

// ---- slots of rings and banks ----------------------------------------------------------
// A slot holds the 24-byte wire record (include/types/particle.hpp:7-18) as three 8-byte
// words, and EACH word carries the lap parity of the slot in a bit the record never uses:
//   word 0 = seed                      bit 63 (seeds are 63-bit, src/random.cpp:14)
//   word 1 = x | mu << 32              bit 62 = bit 30 of mu, clear for |mu| < 2
//   word 2 = wmc | index << 32         bit 63 = sign of the cell index, clear for index >= 0
// A slot written in lap k of its ring carries parity (k + 1) & 1 (memory starts zeroed), so a
// reader that expects lap k accepts a word only once the producer's store of THAT lap has
// landed.  Every word validates itself: no ordering between the three stores is needed, hence
// no fence and no "published count" -- the stores are plain posted writes over NVLink and the
// consumer polls its own memory.  (Credits keep the producer from running a lap ahead.)
__device__ __forceinline__ void slot_store(unsigned long long *rec, unsigned par,
                                           unsigned long long seed, float x, float mu, float wmc,
                                           int idx) {
  st_relaxed_sys(rec, seed | ((unsigned long long)par << 63));
  st_relaxed_sys(rec + 1, (unsigned long long)__float_as_uint(x) |
                              ((unsigned long long)(__float_as_uint(mu) | (par << 30)) << 32));
  st_relaxed_sys(rec + 2, (unsigned long long)__float_as_uint(wmc) |
                              ((unsigned long long)((unsigned)idx | (par << 31)) << 32));
}
// the three words of a slot, loaded together (ONE round trip to the L2); true when all of
// them belong to the expected lap
__device__ __forceinline__ bool slot_load(const unsigned long long *rec, unsigned par,
                                          unsigned long long &a, unsigned long long &b,
                                          unsigned long long &d) {
  a = ld_relaxed_sys(rec);
  b = ld_relaxed_sys(rec + 1);
  d = ld_relaxed_sys(rec + 2);
  return (unsigned)(a >> 63) == par && ((unsigned)(b >> 62) & 1u) == par &&
         (unsigned)(d >> 63) == par;
}
__device__ __forceinline__ void slot_decode(unsigned long long a, unsigned long long b,
                                            unsigned long long d, unsigned long long &seed,
                                            float &x, float &mu, float &wmc, int &idx) {
  seed = a & kMask63;
  x = __uint_as_float((unsigned)b);
  mu = __uint_as_float((unsigned)(b >> 32) & 0xbfffffffu);
  wmc = __uint_as_float((unsigned)d);
  idx = (int)((unsigned)(d >> 32) & 0x7fffffffu);
}

// every history this CTA disables is owed to the home rank's global counter; pay in batches
__device__ __forceinline__ void note_disabled(WorldSmem *sm, const WorldParams &p, unsigned cnt) {
  const unsigned old = atomicAdd(&sm->pend, cnt);
  if (((old + cnt) ^ old) >> 8) {
    const unsigned v = atomicExch(&sm->pend, 0u);
    if (v) red_add_sys(p.home_disabled, (unsigned long long)v);
  }
}

// XS_SMEM: the window's cell constants sit next to its tally in shared memory (false: they
// are read from global memory / L2, which halves the shared memory a cell costs -- the shape
// chosen for sub-slabs of ~1e6 cells)
template <int MAXB, bool XS_SMEM, int RNG>
__global__ void __launch_bounds__(MAXB, MAXB <= 256 ? 4 : 1) world_kernel(const WorldParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WorldSmem *sm = reinterpret_cast<WorldSmem *>(smem_raw);
  // grid = (CTAs per window, windows): no division, so the window's bounds (kernel
  // parameters indexed by blockIdx.y) stay in uniform registers
  const int v = (int)blockIdx.y;                     // window of this CTA
  const int cw = (int)blockIdx.x;                    // CTA within the window
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nwarps = (int)(blockDim.x >> 5);
  const int wv = cw * nwarps + warp;                 // stripe of this warp

  load_math_tables(&sm->math);
  {
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&p.win[v]);
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(&sm->win);
    for (int i = threadIdx.x; i < (int)(sizeof(WindowDesc) / 8); i += blockDim.x) dst[i] = src[i];
  }
  if (threadIdx.x < 9) (&sm->n_cls[0])[threadIdx.x] = 0u;   // n_cls .. bank_pops
  if (threadIdx.x == 0) {
    sm->t_start = global_timer_ns();
    sm->idle_ns = 0ull;
    sm->acc_range = 0u;
  }
  if (threadIdx.x < kWorldMaxWarps) {
    WarpXchg z{};
    z.backoff = 2u;
    z.nap = 256u;
    z.src_done = (int)blockIdx.y != p.src_window ? 1u : 0u;
    sm->wx[threadIdx.x] = z;
  }
  __syncthreads();
  // this CTA's bank carries on where the last run left it (head == tail between runs)
  if (threadIdx.x < 2) (&sm->bank_head)[threadIdx.x] = sm->win.bank.ht[2 * cw + threadIdx.x];
  int lo = p.win_lo[v];
  int m = p.win_lo[v + 1] - lo;
  // opaque to the compiler: kept in registers instead of being re-read from the (indexed)
  // parameter bank in every iteration of the event loop
  asm volatile("" : "+r"(lo), "+r"(m));
  const int hi = lo + m;
  const int ncell = m + kAccExtra;
  // dynamic part: [cell constants] [tally] [source particles of each warp's current chunk] --
  // the two arrays of the event first, at offsets that only depend on kernel parameters
  CellXs *s_xs = reinterpret_cast<CellXs *>(smem_raw + sizeof(WorldSmem));
  unsigned *acc = reinterpret_cast<unsigned *>(s_xs + (XS_SMEM ? m : 0));
  unsigned long long *s_birth =
      reinterpret_cast<unsigned long long *>(acc + (size_t)kAccDigits * ncell) +   // 16 B x ncell
      (size_t)warp * kWorkChunk;
  const CellXs *gxs = p.xs + (lo - p.rank_lo);
  {
    if (XS_SMEM)
      for (int c = threadIdx.x; c < m; c += blockDim.x) s_xs[c] = gxs[c];
    for (int c = threadIdx.x; c < kAccDigits * ncell; c += blockDim.x) acc[c] = 0u;
  }
  __syncthreads();

  unsigned tb_s = smem_addr(&sm->math);
  asm volatile("" : "+r"(tb_s));   // same: the shared-window base is not re-derived per use
  const unsigned xs_s = tb_s + (unsigned)sizeof(WorldSmem);
  const unsigned acc_s = xs_s + (XS_SMEM ? (unsigned)m * (unsigned)sizeof(CellXs) : 0u);
  const unsigned acc_stride = (unsigned)ncell * 4u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const float dx = p.dx, minw = p.minw;
  const int retire_batch = p.retire_batch;
  const unsigned cap = p.ring_cap, ring_log2 = p.ring_log2;
  WarpXchg *wx = &sm->wx[warp];
  // what this warp's stripes look like (warp-uniform; the compiler keeps what fits in uniform
  // registers and re-reads the rest from shared memory)
  // (what only the pass needs -- link modes, bank geometry -- is re-read from shared memory
  // there, so that it does not occupy registers across the event loop)
#define mode0 (sm->win.out[0].mode)
#define mode1 (sm->win.out[1].mode)
#define present0 (sm->win.in[0].present)
#define present1 (sm->win.in[1].present)
#define bank_cap (sm->win.bank.cap)
#define bank_log2 (sm->win.bank.log2cap)
#define bank_rec (sm->win.bank.rec + (size_t)cw * sm->win.bank.cap * 3)

  // particle state, include/types/particle.hpp:7-18, one history per lane
  unsigned long long seed = 0;
  float x = 0.f, mu = 0.f, wmc = 0.f, rmu = 0.f;
  int idx = 0, step = 1;
  bool active = false;
  int cool = 0;        // iterations to wait before looking again for work that was not there
  unsigned n_ev = 0, n_sc = 0, n_it = 0;

  for (;;) {
    // ---- liveness: loop condition of simulate_particle, src/layer.cpp:195-197
    const bool alive = active && (wmc >= minw) && ((unsigned)(idx - lo) < (unsigned)m);
    const unsigned nolive = __ballot_sync(MCB_FULL, !alive);
    // steady state: (almost) all lanes carry a live history -> one vote, one compare, straight
    // to the event.  The pass below runs once `retire_batch` (1..32) lanes are without one (its
    // cost is shared), less often while the last pass found nothing to refill them with, always
    // when none is left.
    if ((unsigned)__popc(nolive) >= (unsigned)retire_batch && (nolive == MCB_FULL || --cool <= 0)) {
      // ================================================================== the pass ==
      const bool fin = active && !alive;
      const unsigned fm = __ballot_sync(MCB_FULL, fin);
      unsigned blocked = 0u;   // bit c: the outbound stripe on side c was full
      // a warp without a single live history: the time until it has one again is idle time
      unsigned long long t_idle0 = 0ull;
      if (nolive == MCB_FULL && fm == 0u) t_idle0 = global_timer_ns();
      // (1) retire: classification of src/layer.cpp:202-217, routing of :332-346
      if (fm) {
        const bool goes[2] = {fin && idx == lo - 1, fin && idx == hi};
        const unsigned m0 = __ballot_sync(MCB_FULL, goes[0]);
        const unsigned m1 = __ballot_sync(MCB_FULL, goes[1]);
        const unsigned md = fm & ~(m0 | m1);   // below min weight inside the window: dead
        // this warp's outbound stripes: records stored so far / credits seen (warp-uniform)
        unsigned wr_c[2] = {wx->wr[0], wx->wr[1]};
        unsigned cred_c[2] = {wx->cred[0], wx->cred[1]};
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const unsigned cm = c ? m1 : m0;
          if (cm == 0u) continue;
          const unsigned cnt = (unsigned)__popc(cm);
          if ((c ? mode1 : mode0) == 0) {
            // global border: absorbed and counted as disabled, src/layer.cpp:350-360
            if (goes[c]) {
              acc_add_smem(acc_s + (unsigned)(m + c) * 4u, acc_stride, wmc, &sm->acc_range);
              active = false;
            }
            if (lane == 0) {
              atomicAdd(&sm->n_cls[c], cnt);
              note_disabled(sm, p, cnt);
              red_add_sys(p.home_chain_disabled + wv, cnt);
            }
          } else {
            // stripe `wv` of the neighbouring window: all-or-nothing per warp and side
            if (wr_c[c] + cnt - cred_c[c] > cap)   // full as far as we know: look again
              cred_c[c] = ld_relaxed_sys(sm->win.out[c].credit + wv);
            if (wr_c[c] + cnt - cred_c[c] > cap) {
              blocked |= 1u << c;   // ring full: these lanes keep their escapee and retry
            } else {
              if (goes[c]) {
                // with a neighbour on another GPU these three stores ARE the communication
                const unsigned q = wr_c[c] + (unsigned)__popc(cm & lt_mask);
                slot_store(sm->win.out[c].rec + ((size_t)wv * cap + (q & (cap - 1u))) * 3,
                           ((q >> ring_log2) + 1u) & 1u, seed, x, mu, wmc, idx);
                active = false;
              }
              wr_c[c] += cnt;
            }
          }
        }
        if (lane == 0) {
          wx->wr[0] = wr_c[0];
          wx->wr[1] = wr_c[1];
          wx->cred[0] = cred_c[0];
          wx->cred[1] = cred_c[1];
        }
        if (md) {
          if (fin && !goes[0] && !goes[1]) {
            acc_add_smem(acc_s + (unsigned)(m + 2) * 4u, acc_stride, wmc, &sm->acc_range);
            active = false;
          }
          if (lane == 0) {
            atomicAdd(&sm->n_cls[2], (unsigned)__popc(md));
            note_disabled(sm, p, (unsigned)__popc(md));
            red_add_sys(p.home_chain_disabled + wv, (unsigned)__popc(md));
          }
        }
        __syncwarp();
      }

      // (2) refill idle lanes: inbound rings, then this CTA's bank, then births
      bool got = false;
      unsigned im = __ballot_sync(MCB_FULL, !active);
      if (im != 0u && (present0 | present1)) {
        // ONE round trip for both sides: the idle lanes are split between the two inbound
        // stripes in proportion to what each side has delivered so far (rd), idle lane number
        // k of a side looks at slot rd + k of it; taken = the leading run of arrived slots
        const unsigned rd0 = wx->rd[0], rd1 = wx->rd[1];
        __syncwarp();
        const unsigned nidle = (unsigned)__popc(im);
        unsigned n1;   // idle lanes that look at side 1 (the first n1 of them)
        if (!present0) n1 = nidle;
        else if (!present1) n1 = 0u;
        else {
          const float f = __fdividef((float)rd1 + 1.0f, (float)rd0 + (float)rd1 + 2.0f);
          n1 = (unsigned)__float2int_rn(f * (float)nidle);
          if (nidle >= 2u) n1 = n1 < 1u ? 1u : (n1 > nidle - 1u ? nidle - 1u : n1);
          else n1 = (rd0 + rd1 + (unsigned)n_it) & 1u;   // a single idle lane alternates
        }
        const unsigned r = (unsigned)__popc(im & lt_mask);
        const unsigned side = r < n1 ? 1u : 0u;
        const unsigned k = side ? r : r - n1;
        unsigned long long a = 0ull, b = 0ull, d = 0ull;
        bool valid = false;
        if (!active) {
          const unsigned q = (side ? rd1 : rd0) + k;
          valid = slot_load(sm->win.in[side].rec + ((size_t)wv * cap + (q & (cap - 1u))) * 3,
                            ((q >> ring_log2) + 1u) & 1u, a, b, d);
        }
        const unsigned vm = __ballot_sync(MCB_FULL, valid);
        const unsigned s1 = __ballot_sync(MCB_FULL, !active && side == 1u);
        const unsigned s0 = im & ~s1;
        const unsigned inv0 = s0 & ~vm, inv1 = s1 & ~vm;
        // lanes below the first lane of the side whose slot has not arrived
        const unsigned ok0 = inv0 ? (inv0 & (0u - inv0)) - 1u : MCB_FULL;
        const unsigned ok1 = inv1 ? (inv1 & (0u - inv1)) - 1u : MCB_FULL;
        const unsigned t0 = (unsigned)__popc(s0 & vm & ok0), t1 = (unsigned)__popc(s1 & vm & ok1);
        if (valid && (((side ? ok1 : ok0) >> lane) & 1u)) {
          slot_decode(a, b, d, seed, x, mu, wmc, idx);
          rmu = recip_for_div(mu);
          step = dir_step(mu);
          active = true;
        }
        if (t0 | t1) {
          // credits straight away: the slots have been read (their values decided the branch)
          if (lane == 0 && t0) {
            wx->rd[0] = rd0 + t0;
            st_relaxed_sys(sm->win.in[0].credit + wv, rd0 + t0);
          }
          if (lane == 1 && t1) {
            wx->rd[1] = rd1 + t1;
            st_relaxed_sys(sm->win.in[1].credit + wv, rd1 + t1);
          }
          got = true;
          im = __ballot_sync(MCB_FULL, !active);
        }
      }
      if (im != 0u && *(volatile unsigned *)&sm->bank_tail != *(volatile unsigned *)&sm->bank_head) {
        // this CTA's bank (its warps push and pop: claim with a CAS on the head)
        unsigned h = 0u, n = 0u;
        if (lane == 0) {
          const unsigned nidle = (unsigned)__popc(im);
          for (;;) {
            h = *(volatile unsigned *)&sm->bank_head;
            const unsigned avail = *(volatile unsigned *)&sm->bank_tail - h;
            if (avail == 0u || avail > bank_cap) break;
            n = avail < nidle ? avail : nidle;
            if (atomicCAS(&sm->bank_head, h, h + n) == h) break;
            n = 0u;
          }
        }
        h = __shfl_sync(MCB_FULL, h, 0);
        n = __shfl_sync(MCB_FULL, n, 0);
        if (n) {
          const unsigned r = (unsigned)__popc(im & lt_mask);
          if (!active && r < n) {
            const unsigned q = h + r;
            const unsigned long long *rec = bank_rec + (size_t)(q & (bank_cap - 1u)) * 3;
            const unsigned par = ((q >> bank_log2) + 1u) & 1u;
            unsigned long long a, b, d;
            unsigned spins = 0u;
            // the pusher of this slot may still be writing it (it never waits on anyone)
            while (!slot_load(rec, par, a, b, d)) {
              if (++spins > (1u << 22)) {   // cannot happen unless the queue was corrupted
                atomicExch(&p.ctrl->error, (unsigned)(-MCB200_ERR_CAPACITY));
                break;
              }
            }
            slot_decode(a, b, d, seed, x, mu, wmc, idx);
            rmu = recip_for_div(mu);
          step = dir_step(mu);
            active = true;
          }
          if (lane == 0) atomicAdd(&sm->bank_pops, n);
          got = true;
          im = __ballot_sync(MCB_FULL, !active);
        }
      }
      if (im != 0u && !wx->src_done) {
        // births, src/layer.cpp:101-120: particle i carries rnd_seed^(i+1)(chain_state) and
        // consumes the first draw of its own stream for mu.  Source particles are handed to
        // warps in chunks of kWorkChunk.  Claiming a chunk computes all its particle seeds at
        // once, lane L those of particles L and L + 32 (one 63-step jump-ahead per lane and
        // chunk), and parks them in shared memory.
        unsigned w_off = wx->w_off, w_cnt = wx->w_cnt;
        const unsigned born_chain = wx->born;
        // Throttle 1: the histories in flight in the whole world (born - disabled, both counted
        // on this rank).  Throttle 2: the histories in flight of THIS warp's chain -- its own
        // births minus what every rank reported back for stripe `wv`: a chain never takes more
        // than its share, so none can hoard the budget in front of a slow warp downstream
        // while the others run dry.  (A record that changed chains through a bank is credited
        // to the other chain; if the world as a whole runs low, throttle 2 steps aside.)
        unsigned long long born_all = 0ull;
        unsigned live_all_lo = 0u, chain_disabled = 0u;
        if (lane == 0) {
          born_all = ld_relaxed_sys(&p.ctrl->born);
          const unsigned long long live = born_all - ld_relaxed_sys(&p.ctrl->disabled_global);
          live_all_lo = live > 0xffffffffull ? 0xffffffffu : (unsigned)live;
          chain_disabled = ld_relaxed_sys(p.home_chain_disabled + wv);
        }
        born_all = __shfl_sync(MCB_FULL, born_all, 0);
        live_all_lo = __shfl_sync(MCB_FULL, live_all_lo, 0);
        chain_disabled = __shfl_sync(MCB_FULL, chain_disabled, 0);
        const bool world_full = (unsigned long long)live_all_lo >= p.inflight_limit;
        const bool world_low = (unsigned long long)live_all_lo < (p.inflight_limit >> 3);
        const int chain_live = (int)(born_chain - chain_disabled);
        unsigned allowed = chain_live < (int)p.chain_limit ? p.chain_limit - (unsigned)max(chain_live, 0) : 0u;
        if (world_low && allowed < 32u) allowed = 32u;
        // (`born` counts the chunks CLAIMED, so throttle 1 gates the claims only: a warp always
        // may hand out what it has claimed)
        // Throttle 3, pacing: no births while one of this warp's outbound stripes is more than
        // half full -- the source follows the rate its neighbours take records at instead of
        // running into a full ring (which would cost it its lanes)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if ((c ? mode1 : mode0) == 0) continue;
          const unsigned wr = wx->wr[c];
          unsigned cred = wx->cred[c];
          if (wr - cred > (cap >> 1)) {
            cred = ld_relaxed_sys(sm->win.out[c].credit + wv);
            __syncwarp();
            if (lane == 0) wx->cred[c] = cred;
            if (wr - cred > (cap >> 1)) allowed = 0u;
          }
        }
        __syncwarp();
        if (w_off == w_cnt && ((allowed != 0u && !world_full) || born_all >= p.src_total)) {
          unsigned long long base = p.src_total;
          if (born_all < p.src_total) {
            if (lane == 0) base = atomicAdd(&p.ctrl->born, (unsigned long long)kWorkChunk);
            base = __shfl_sync(MCB_FULL, base, 0);
          }
          if (base >= p.src_total) {
            if (lane == 0) wx->src_done = 1u;
          } else {
            if (RNG == 0) {
              const unsigned long long s = jump_state(c_seed_jump, base + 1ull + (unsigned)lane,
                                                      p.chain_state);
              s_birth[lane] = lcg_next(s);                                   // :112 draw #1
              s_birth[lane + 32] = lcg_next(affine_apply(c_seed_jump.pow2[5], s));
            } else {
              // counter-based: history `base + j` is its own counter, event 0 = the birth
              s_birth[lane] = (base + (unsigned)lane) << kPhiloxEventBits;
              s_birth[lane + 32] = (base + 32ull + (unsigned)lane) << kPhiloxEventBits;
            }
            __syncwarp();
            w_off = 0u;
            w_cnt = base + kWorkChunk <= p.src_total ? (unsigned)kWorkChunk
                                                     : (unsigned)(p.src_total - base);
          }
        }
        unsigned avail = w_cnt - w_off;
        if (avail > allowed) avail = allowed;
        const unsigned nidle = (unsigned)__popc(im);
        const unsigned n = avail < nidle ? avail : nidle;
        if (n) {
          const unsigned r = (unsigned)__popc(im & lt_mask);
          if (!active && r < n) {
            seed = s_birth[w_off + r];
            if (RNG == 0) {
              mu = __fsub_rn(__fmul_rn(2.0f, lcg_to_real(seed)), 1.0f);
            } else {
              mu = __fsub_rn(__fmul_rn(2.0f, u32_to_real(philox_draws(seed, p.rng_key).x)), 1.0f);
              seed += 1ull;
            }
            rmu = recip_for_div(mu);
            step = dir_step(mu);
            x = p.x_ini;
            wmc = p.wmc;
            idx = p.src_index;
            active = true;
          }
          w_off += n;
          if (lane == 0) atomicAdd(&sm->births, n);
          got = true;
          im = __ballot_sync(MCB_FULL, !active);
        }
        if (lane == 0) {
          wx->w_off = w_off;
          wx->w_cnt = w_cnt;
          wx->born = born_chain + n;
        }
        __syncwarp();
      }

      // (3) a blocked sender makes room for THAT neighbour: the records the neighbour sent this
      // way move from the ring into the CTA's bank, so that two windows can never wait on each
      // other.  Only that side: the other inbound ring keeps filling, which is the back-pressure
      // the windows upstream (and the birth pacing of the source) react to.
      if (blocked) {
        __syncwarp();
        if (lane == 0) atomicAdd(&sm->blocked, 1u);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (!((blocked >> c) & 1u) || !(c ? present1 : present0)) continue;
          const unsigned rd = wx->rd[c];
          const unsigned q = rd + (unsigned)lane;
          unsigned long long a, b, d;
          const bool valid = slot_load(sm->win.in[c].rec + ((size_t)wv * cap + (q & (cap - 1u))) * 3,
                                       ((q >> ring_log2) + 1u) & 1u, a, b, d);
          const unsigned rv = __ballot_sync(MCB_FULL, valid);
          unsigned take = rv == MCB_FULL ? 32u : (unsigned)(__ffs((int)~rv) - 1);
          if (take == 0u) continue;
          // reserve bank slots; never closer than one claim of every warp (32 x warps) to the
          // slots a popper may still be reading, never beyond the capacity (what does not fit
          // stays in the ring)
          unsigned t = 0u;
          if (lane == 0) {
            const unsigned margin = 32u * (unsigned)nwarps;
            for (;;) {
              t = *(volatile unsigned *)&sm->bank_tail;
              const unsigned used = t - *(volatile unsigned *)&sm->bank_head;
              const unsigned room = used + margin < bank_cap ? bank_cap - margin - used : 0u;
              if (room < take) take = room;
              if (take == 0u || atomicCAS(&sm->bank_tail, t, t + take) == t) break;
            }
            if (take) atomicAdd(&sm->bank_pushes, take);
            else atomicExch(&p.ctr->bank_full, 1u);
          }
          t = __shfl_sync(MCB_FULL, t, 0);
          take = __shfl_sync(MCB_FULL, take, 0);
          if (take == 0u) continue;
          if ((unsigned)lane < take) {
            unsigned long long s2;
            float x2, mu2, w2;
            int i2;
            slot_decode(a, b, d, s2, x2, mu2, w2, i2);
            const unsigned qb = t + (unsigned)lane;
            slot_store(bank_rec + (size_t)(qb & (bank_cap - 1u)) * 3,
                       ((qb >> bank_log2) + 1u) & 1u, s2, x2, mu2, w2, i2);
          }
          if (lane == 0) {
            wx->rd[c] = rd + take;
            st_relaxed_sys(sm->win.in[c].credit + wv, rd + take);
          }
          __syncwarp();
        }
      }

      // (4) nothing to track (idle lanes, or escapees waiting for room in a full ring) and
      // nothing to fetch: bookkeeping, termination, back off
      const bool live_now = active && (wmc >= minw) && ((unsigned)(idx - lo) < (unsigned)m);
      const unsigned lm = __ballot_sync(MCB_FULL, live_now);
      if (lm == 0u && !got) {
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
        // an idle warp must not eat the issue slots of the warps still tracking on its SM
        const unsigned nap = wx->nap;
        __syncwarp();
        __nanosleep(nap);
        if (lane == 0) {
          if (nap < 2048u) wx->nap = nap << 1;
          if (t_idle0) atomicAdd(&sm->idle_ns, global_timer_ns() - t_idle0);
        }
      } else if (lane == 0) {
        wx->nap = 256u;
      }
      // lanes still without a live history: look again later, ever less often (a few events)
      if (lm != MCB_FULL && !got) {
        const unsigned backoff = wx->backoff;
        __syncwarp();
        cool = (int)backoff;
        if (lane == 0 && backoff < 16u) wx->backoff = backoff * 2u;
      } else {
        cool = 0;
        if (lane == 0) wx->backoff = 2u;
      }
      __syncwarp();
      continue;  // fresh lanes go through the liveness test first
    }

    // ---- one event per live lane: Layer::particle_step, src/layer.cpp:123-190
    if (alive) {
      event_step<XS_SMEM, true, RNG>(seed, x, mu, wmc, rmu, step, idx, n_sc, lo, dx, tb_s, xs_s, acc_s,
                                acc_stride, gxs, nullptr, ncell, &sm->acc_range, p.rng_key);
      ++n_ev;
      // a second event under the same vote for the lanes that are still live: the loop top is
      // shared by two events; a lane that finished on the first one waits one slot longer
      if ((wmc >= minw) && ((unsigned)(idx - lo) < (unsigned)m)) {
        event_step<XS_SMEM, true, RNG>(seed, x, mu, wmc, rmu, step, idx, n_sc, lo, dx, tb_s, xs_s, acc_s,
                                  acc_stride, gxs, nullptr, ncell, &sm->acc_range, p.rng_key);
        ++n_ev;
      }
    }
    n_it += 2;   // lane slots offered to the event: n_ev / n_it = lane utilisation
  }

  // ---- per-CTA flush: counters once, the CTA-private tally merged into the rank's
  unsigned long long ev = n_ev, sc = n_sc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ev += __shfl_xor_sync(MCB_FULL, ev, o);
    sc += __shfl_xor_sync(MCB_FULL, sc, o);
  }
  if (lane == 0) {
    if (ev) atomicAdd(&p.ctr->events, ev);
    if (sc) atomicAdd(&p.ctr->scatters, sc);
    if (n_it) atomicAdd(&p.ctr->lane_slots, 32ull * n_it);
    // records this warp pushed into its outbound stripes, per side
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const unsigned w = wx->wr[c];
      if (w) {
        atomicAdd(&p.ctr->sent[c], (unsigned long long)w);
        if (sm->win.out[c].outer) atomicAdd(&p.ctr->sent_outer[c], (unsigned long long)w);
      }
    }
  }
  __syncthreads();
  {
    unsigned long long *gacc = p.acc + (lo - p.rank_lo);
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
    if (sm->births) atomicAdd(&g->births, (unsigned long long)sm->births);
    if (sm->idle_polls) atomicAdd(&g->idle_polls, (unsigned long long)sm->idle_polls);
    if (sm->idle_ns) atomicAdd(&g->idle_ns, sm->idle_ns);
    if (sm->blocked) atomicAdd(&g->blocked_passes, (unsigned long long)sm->blocked);
    if (sm->bank_pushes) atomicAdd(&g->bank_pushes, (unsigned long long)sm->bank_pushes);
    if (sm->bank_pops) atomicAdd(&g->bank_pops, (unsigned long long)sm->bank_pops);
    if (sm->acc_range) atomicExch(&g->acc_range, 1u);
    // the bank's cursors for the next run (head == tail unless the run was stopped)
    sm->win.bank.ht[2 * cw] = sm->bank_head;
    sm->win.bank.ht[2 * cw + 1] = sm->bank_tail;
    // anything still owed to the home counter (only on an abnormal exit)
    const unsigned owe = atomicExch(&sm->pend, 0u);
    if (owe) red_add_sys(p.home_disabled, (unsigned long long)owe);
  }
}

#undef mode0
#undef mode1
#undef present0
#undef present1
#undef bank_cap
#undef bank_log2
#undef bank_rec

size_t world_smem_bytes(int m_max, int block, bool xs_smem) {
  return sizeof(WorldSmem) + (size_t)(block / 32) * kWorkChunk * sizeof(unsigned long long) +
         (xs_smem ? (size_t)m_max * sizeof(CellXs) : 0) +
         (size_t)(m_max + kAccExtra) * kAccDigits * sizeof(unsigned) + 8;
}

typedef void (*WorldFn)(const WorldParams);
static WorldFn world_fn(int block, bool xs_smem, int rng) {
  if (rng == 0) {
    if (block <= 256) return xs_smem ? world_kernel<256, true, 0> : world_kernel<256, false, 0>;
    return xs_smem ? world_kernel<1024, true, 0> : world_kernel<1024, false, 0>;
  }
  if (block <= 256) return xs_smem ? world_kernel<256, true, 1> : world_kernel<256, false, 1>;
  return xs_smem ? world_kernel<1024, true, 1> : world_kernel<1024, false, 1>;
}

cudaError_t world_configure(int device, int m_max, int block, bool xs_smem, WorldLaunch *out,
                            int *max_ctas_per_sm) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return e;
  const size_t smem = world_smem_bytes(m_max, block, xs_smem);
  if (smem > prop.sharedMemPerBlockOptin || block % 32 || block > 1024 || block < 32)
    return cudaErrorInvalidValue;
  for (int rng = 0; rng < 2; ++rng) {
    e = cudaFuncSetAttribute(world_fn(block, xs_smem, rng),
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  int per_sm = 0, per_sm1 = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, world_fn(block, xs_smem, 0), block, smem);
  if (e != cudaSuccess) return e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm1, world_fn(block, xs_smem, 1), block, smem);
  if (e != cudaSuccess) return e;
  if (per_sm1 < per_sm) per_sm = per_sm1;   // the launch shape must hold for both generators
  if (per_sm < 1) return cudaErrorInvalidValue;
  out->block = block;
  out->grid = prop.multiProcessorCount * per_sm;   // every CTA resident: the kernel is persistent
  out->smem = smem;
  out->xs_smem = xs_smem ? 1 : 0;
  if (max_ctas_per_sm) *max_ctas_per_sm = per_sm;
  return cudaSuccess;
}

cudaError_t world_upload_jump_table(const JumpTable &jt) {
  return cudaMemcpyToSymbol(c_seed_jump, &jt, sizeof(JumpTable));
}

cudaError_t launch_world(const WorldParams &p, const WorldLaunch &cfg, cudaStream_t stream) {
  world_fn(cfg.block, cfg.xs_smem != 0, p.rng)<<<dim3((unsigned)p.cpw, (unsigned)p.V), cfg.block,
                                                 cfg.smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace mcb
