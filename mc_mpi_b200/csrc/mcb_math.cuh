// mcb_math.cuh -- device arithmetic of the tracking path, bit-compatible with
// the reference's x86-64 build.
//
//  * LCG streams: src/random.cpp:8-29 (rnd_real, rnd_seed) -- counter-free
//    64-bit LCG modulo 2^63 held in two registers; jump-ahead tables make the
//    per-particle seed chain (src/layer.cpp:111) computable in parallel.
//  * logf / expf: the reference calls the system libm (glibc 2.39).  glibc's
//    single-precision log/exp are the ARM optimized-routines algorithms, which
//    evaluate in DOUBLE and round once.  They are implemented here with
//    DFMA/DMUL/DADD (B200 has half-rate FP64), table look-ups from shared
//    memory and integer bit tricks for the conversions, so that the GPU
//    trajectory of a particle is the CPU trajectory bit for bit.  CUDA's own
//    logf/expf (1 ulp, different polynomial) would not be.
//    oracle/sweep_libm.c shows the restated algorithm == libm on every input
//    the path can produce; tests/test_gpu_primitives.py shows device == libm.
//  * All float physics uses __f{add,sub,mul,div}_rn so nvcc can never contract
//    a*b+c into an FMA (the reference build has none: no -march, SSE2 only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mcb {

// src/random.cpp:8-10
constexpr uint64_t kRngG = 6364136223846793005ull;
constexpr uint64_t kRngC = 1442695040888963407ull;
// src/random.cpp:20-22
constexpr uint64_t kSeedG = 5177284530976225183ull;
constexpr uint64_t kSeedC = 2096348467109453893ull;
constexpr uint64_t kMask63 = 0x7fffffffffffffffull;

// include/types/types.hpp:14-16
#define MCB_EPS 1e-4f
#define MCB_MAXREAL 3.402823466e+38f

// glibc e_logf_data.c: {1/c, log(c)} for 16 sub-intervals of [0.7, 1.4)
// glibc e_exp2f_data.c: bits(2^(i/32)) - (i << 47)
struct MathTables {
  double2 log_tab[16];
  unsigned long long exp_tab[32];
};

static __constant__ MathTables c_math_tables = {
    {{0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2},
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
     {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}},
    {0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full,
     0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull, 0x3fef54873168b9aaull,
     0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull,
     0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
     0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull,
     0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull,
     0x3feea11473eb0187ull, 0x3feea589994cce13ull, 0x3feeace5422aa0dbull,
     0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
     0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull,
     0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full,
     0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull}};

// polynomial / reduction constants of the two routines, in the constant bank so that the
// FP64 instructions take them as c[bank][offset] operands instead of re-materialising
// 64-bit immediates in registers on every event
struct MathConsts {
  double ln2, a0, a1, a2;                 // e_logf_data.c: ln2, poly[0..2]
  double inv_ln2_n, shift, c0, c1, c2;    // e_exp2f_data.c: invln2_scaled, shift, poly_scaled
};
static __constant__ MathConsts c_mc = {
    0x1.62e42fefa39efp-1, -0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2,
    0x1.71547652b82fep+0 * 32, 0x1.8p+52, 0x1.c6af84b912394p-5 / 32 / 32 / 32,
    0x1.ebfce50fac4f3p-3 / 32 / 32, 0x1.62e42ff0c52d6p-1 / 32};

// copy the 512-byte tables into shared memory (lane-divergent indices would
// serialise on the constant cache)
__device__ __forceinline__ void load_math_tables(MathTables *s) {
  const int n = (int)(sizeof(MathTables) / sizeof(unsigned long long));
  const unsigned long long *src =
      reinterpret_cast<const unsigned long long *>(&c_math_tables);
  unsigned long long *dst = reinterpret_cast<unsigned long long *>(s);
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// ---- explicit shared-memory accesses ----------------------------------------
// The hot loop addresses shared memory with 32-bit shared-window offsets kept in
// a register and explicit ld.shared / atom.shared: going through generic
// pointers makes ptxas re-derive the window base (S2R SR_CgaCtaId + LEA) at every
// use once registers are capped at 40.
__device__ __forceinline__ unsigned smem_addr(const void *p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ float2 lds_f32x2(unsigned addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f32x4(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned long long lds_u64(unsigned addr) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned atoms_add_u32(unsigned addr, unsigned v) {
  unsigned old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}

// ---- LCG ------------------------------------------------------------------
// one step of the rnd_real stream, src/random.cpp:14 ((G*s + C) % 2^63)
__device__ __forceinline__ uint64_t lcg_next(uint64_t s) {
  return (kRngG * s + kRngC) & kMask63;
}
// src/random.cpp:13,15: (float)seed * (1.0f / 2^63); cvt.rn.f32.u64 rounds to
// nearest like the x86 conversion, the scale is a power of two (exact)
__device__ __forceinline__ float lcg_to_real(uint64_t s) {
  return __fmul_rn(__ull2float_rn(s), 0x1p-63f);
}

// ---- Philox2x32-10: the counter-based "performance mode" generator ---------------------
// (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123's
// philox2x32_10: multiplier 0xD256D193, Weyl key bump 0x9E3779B9, 10 rounds -- pinned to its
// known-answer vectors by tests/test_gpu_primitives.py.)  A particle's 64-bit `seed` field is
// the COUNTER: history id << 23 | event number; the key is the run's seed.  One call per
// event gives both draws an event can consume: word 0 -> the free-flight draw h, word 1 ->
// the scattering angle.  No seed chain, no state to advance but the event number: any
// event of any history can be recomputed from (id, event, key) alone.  The floats are formed
// like rnd_real forms them (src/random.cpp:13-15): integer -> float, round to nearest, times a
// power of two -- so h lies in [0, 1] inclusive here too.
constexpr uint32_t kPhiloxM = 0xD256D193u;
constexpr uint32_t kPhiloxW = 0x9E3779B9u;
constexpr int kPhiloxEventBits = 23;   // events per history before the counter runs into the id
__device__ __forceinline__ uint2 philox2x32_10(uint32_t c0, uint32_t c1, uint32_t key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi = __umulhi(kPhiloxM, c0);
    const uint32_t lo = kPhiloxM * c0;
    c0 = hi ^ key ^ c1;
    c1 = lo;
    key += kPhiloxW;
  }
  return make_uint2(c0, c1);
}
__device__ __forceinline__ uint2 philox_draws(uint64_t counter, uint32_t key) {
  return philox2x32_10((uint32_t)counter, (uint32_t)(counter >> 32), key);
}
__device__ __forceinline__ float u32_to_real(uint32_t w) {
  return __fmul_rn(__uint2float_rn(w), 0x1p-32f);
}

// affine map s -> a*s + c (mod 2^63); powers of the LCG step are such maps
struct Affine {
  uint64_t a, c;
};
__host__ __device__ inline uint64_t affine_apply(Affine f, uint64_t s) {
  return (f.a * s + f.c) & kMask63;
}
// g after f
__host__ __device__ inline Affine affine_compose(Affine g, Affine f) {
  Affine r;
  r.a = (g.a * f.a) & kMask63;
  r.c = (g.a * f.c + g.c) & kMask63;
  return r;
}
// pow2[j] = (LCG step)^(2^j), j < 63
struct JumpTable {
  Affine pow2[63];
};
__host__ inline JumpTable make_jump_table(uint64_t g, uint64_t c) {
  JumpTable t;
  Affine f{g & kMask63, c & kMask63};
  for (int j = 0; j < 63; ++j) {
    t.pow2[j] = f;
    f = affine_compose(f, f);
  }
  return t;
}
__host__ __device__ inline Affine jump_map(const JumpTable &t, uint64_t k) {
  Affine r{1, 0};
  for (int j = 0; j < 63 && (k >> j); ++j)
    if ((k >> j) & 1) r = affine_compose(t.pow2[j], r);
  return r;
}
// state after k steps from s (powers of one map commute: any bit order works)
__host__ __device__ inline uint64_t jump_state(const JumpTable &t, uint64_t k, uint64_t s) {
  for (int j = 0; j < 63 && (k >> j); ++j)
    if ((k >> j) & 1) s = affine_apply(t.pow2[j], s);
  return s;
}

// ---- logf -----------------------------------------------------------------
// glibc sysdeps/ieee754/flt-32/e_logf.c (LOGF_TABLE_BITS 4, POLY_ORDER 4).
// Domain on the path: h = rnd_real() in {+0} U [2^-63, 1] (src/layer.cpp:136).
// `tb` = smem_addr() of the CTA's MathTables copy.
__device__ __forceinline__ float logf_glibc(float x, unsigned tb) {
  const double ln2 = c_mc.ln2, a0 = c_mc.a0, a1 = c_mc.a1, a2 = c_mc.a2;
  const uint32_t ix = __float_as_uint(x);
  const uint32_t tmp = ix - 0x3f330000u;
  const uint32_t i = (tmp >> 19) & 15u;
  const int k = (int)tmp >> 23;
  const uint32_t iz = ix - (tmp & 0xff800000u);
  // (double)asfloat(iz): iz is a positive normal float in [0.7, 1.4), so the
  // widening is an exponent re-bias and a mantissa shift
  const double z = __hiloint2double((int)((iz >> 3) + 0x38000000u), (int)(iz << 29));
  const double2 t = lds_f64x2(tb + i * 16u);  // MathTables::log_tab[i]
  const double r = __fma_rn(z, t.x, -1.0);
  const double y0 = __fma_rn((double)k, ln2, t.y);
  const double r2 = __dmul_rn(r, r);
  double y = __fma_rn(a1, r, a2);
  y = __fma_rn(a0, r2, y);
  y = __fma_rn(y, r2, __dadd_rn(y0, r));
  float res = __double2float_rn(y);
  // logf(+0) = -inf (glibc: __math_divzerof); logf(1) = +0 falls out of the
  // main path (r = 0, y0 = 0)
  return ix == 0u ? __int_as_float(0xff800000) : res;
}

// ---- expf -----------------------------------------------------------------
// glibc sysdeps/ieee754/flt-32/e_expf.c (EXP2F_TABLE_BITS 5).  Domain on the
// path: -sig_a*di in [-inf, +0] (src/layer.cpp:175); the result only enters
// as 1 - expf().
// the double-precision core: valid for x in [-104, 0] (finite); returns y with expf(x) = (float)y
__device__ __forceinline__ double expf_core(float x, unsigned tb) {
  const double shift = c_mc.shift, inv_ln2_n = c_mc.inv_ln2_n;
  const double c0 = c_mc.c0, c1 = c_mc.c1, c2 = c_mc.c2;
  // (double)x: one F2F on the (lightly used) conversion pipe; an integer re-bias would cost
  // five issue slots of a kernel that is bound by instruction issue
  const double xd = (double)x;
  const double z = __dmul_rn(inv_ln2_n, xd);
  double kd = __dadd_rn(z, shift);
  const uint32_t ki = (uint32_t)__double2loint(kd);
  kd = __dsub_rn(kd, shift);
  const double r = __dsub_rn(z, kd);
  // t = T[ki % 32] + (ki << 47): only the high word changes
  const unsigned long long t0 =
      lds_u64(tb + (unsigned)sizeof(double2) * 16u + (ki & 31u) * 8u);  // MathTables::exp_tab
  const double s = __hiloint2double((int)((uint32_t)(t0 >> 32) + (ki << 15)),
                                    (int)(uint32_t)t0);
  const double zz = __fma_rn(c0, r, c1);
  const double r2 = __dmul_rn(r, r);
  double y = __fma_rn(c2, r, 1.0);
  y = __fma_rn(zz, r2, y);
  return __dmul_rn(y, s);
}
__device__ __forceinline__ float expf_glibc_nonpos(float x, unsigned tb) {
  const float res = __double2float_rn(expf_core(x, tb));
  // x < -103.97 underflows to +0 in glibc; covers -inf and the garbage the core makes of huge |x|
  return (x > -104.0f) ? res : 0.0f;
}
// 1 - expf(x) as the event needs it (src/layer.cpp:175).  expf(x) <= 2^-25 gives exactly 1.0f
// whatever its bits, so arguments below -104 (down to -inf, or NaN) are clamped instead of
// tested: one FMNMX replaces a compare and a select
__device__ __forceinline__ float one_minus_expf_nonpos(float x, unsigned tb) {
  return __fsub_rn(1.0f, __double2float_rn(expf_core(fmaxf(x, -104.0f), tb)));
}

}  // namespace mcb
