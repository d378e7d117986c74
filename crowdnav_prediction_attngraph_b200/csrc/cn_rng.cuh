// Legacy numpy RandomState (MT19937) replica, one generator per environment.
//
// The reference re-seeds the process-global legacy `np.random` at every reset
// (crowd_sim/envs/crowd_sim_var_num.py:333-338) and draws from it inside data-dependent
// rejection loops, so bit-identical spawn positions need the exact stream:
//   np.random.seed(int)   -> init_genrand (Knuth LCG 1812433253)
//   random_sample()       -> (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53   (two 32-bit outputs)
//   uniform(lo, hi)       -> lo + (hi - lo) * random_sample()
// State lives in HBM (624 words per environment); it is only touched on reset / goal events.
#pragma once
#include "cn_common.cuh"

struct CnRng {
  uint32_t* key;   // 624 words
  int pos;
  // Deferral of pathological rejection-sampling searches (cn_rejection_sample): a WARP-scope caller gives every
  // search a budget of tries; a search that exhausts it sets `deferred` and the whole event of that environment is
  // redone from scratch by a 512-thread CTA (cn_env_event_heavy_kernel).  budget 0 = unlimited.
  int budget;
  int deferred;
};

// Cooperative execution context.  {lane, 32}: the 32 lanes of a warp execute the same (replicated,
// warp-uniform) control flow and split data-parallel inner loops between them; {0, 1}: a single
// thread (the CPU test harness, where every collective degenerates to the identity); {tid, blockDim, scratch}
// with nlanes > 32: a whole CTA in replicated control flow (collectives become CTA barriers; `scratch` = a few
// ints of shared memory).  Only the RNG / rejection-sampling path supports the CTA scope.
struct CnCoop {
  int lane, nlanes;
  int* scratch;
  float* ftab;       // CTA scope: 5 x CN_FTAB floats of shared memory (fp32 agent table of the rejection sampler)
  // sub-warp groups (nlanes = 16: the two halves of a warp work on different humans at the same time): member mask of
  // the group and the warp lane of its lane 0.  mask == 0 means the whole warp (base 0).
  uint32_t mask;
  int base;
};
CN_HD uint32_t cn_gmask(const CnCoop& c) { return c.mask ? c.mask : 0xffffffffu; }
#define CN_FTAB 132
CN_HD bool cn_any(const CnCoop& c, bool pred) {
#if defined(__CUDA_ARCH__)
  if (c.nlanes > 32) return __syncthreads_or(pred ? 1 : 0) != 0;
  if (c.nlanes > 1) return __any_sync(cn_gmask(c), pred) != 0;
#endif
  return pred;
}
CN_HD void cn_coop_sync(const CnCoop& c) {
#if defined(__CUDA_ARCH__)
  if (c.nlanes > 32) __syncthreads();
  else if (c.nlanes > 1) __syncwarp(cn_gmask(c));
#endif
  (void)c;
}
CN_HD uint32_t cn_ballot(const CnCoop& c, bool pred) {
#if defined(__CUDA_ARCH__)
  if (c.nlanes > 1) return (__ballot_sync(cn_gmask(c), pred) & cn_gmask(c)) >> c.base;      // group-relative bits
#endif
  return pred ? 1u : 0u;
}
CN_HD float cn_bcast_f(const CnCoop& c, float v, int src) {
#if defined(__CUDA_ARCH__)
  if (c.nlanes > 1) return __shfl_sync(cn_gmask(c), v, c.base + src);
#endif
  (void)src;
  return v;
}
CN_HD int cn_bcast_i(const CnCoop& c, int v, int src) {
#if defined(__CUDA_ARCH__)
  if (c.nlanes > 1) return __shfl_sync(cn_gmask(c), v, c.base + src);
#endif
  (void)src;
  return v;
}
// warp-wide min / max (exact, order independent for non-NaN inputs).  On the device the float is
// mapped to a monotonically ordered int32 and reduced with one REDUX instruction.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ int cn_f2ord(float f) { const int b = __float_as_int(f); return b ^ ((b >> 31) & 0x7fffffff); }
__device__ __forceinline__ float cn_ord2f(int o) { return __int_as_float(o ^ ((o >> 31) & 0x7fffffff)); }
#endif
CN_HD float cn_warp_min(const CnCoop& c, float v) {
#if defined(__CUDA_ARCH__)
  if (c.nlanes > 1) return cn_ord2f(__reduce_min_sync(cn_gmask(c), cn_f2ord(v)));
#endif
  (void)c;
  return v;
}
CN_HD float cn_warp_max(const CnCoop& c, float v) {
#if defined(__CUDA_ARCH__)
  if (c.nlanes > 1) return cn_ord2f(__reduce_max_sync(cn_gmask(c), cn_f2ord(v)));
#endif
  (void)c;
  return v;
}
CN_HD int cn_popc_below(const CnCoop& c, uint32_t mask) {   // set bits of `mask` at lane positions below this lane
#if defined(__CUDA_ARCH__)
  if (c.nlanes > 1) return __popc(mask & ((1u << c.lane) - 1u));
#endif
  (void)mask;
  return 0;
}
CN_HD int cn_popc(uint32_t m) {
#if defined(__CUDA_ARCH__)
  return __popc(m);
#else
  return __builtin_popcount(m);
#endif
}
CN_HD int cn_ffs(uint32_t m) {    // index of the lowest set bit (m != 0)
#if defined(__CUDA_ARCH__)
  return __ffs(m) - 1;
#else
  return __builtin_ffs((int)m) - 1;
#endif
}

CN_HD void cn_rng_seed(CnRng& r, uint32_t seed, const CnCoop& c) {
  // inherently serial recurrence: lane 0 only
  if (c.lane == 0) {
    for (int i = 0; i < 624; ++i) {
      r.key[i] = seed;
      seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)(i + 1);
    }
  }
  cn_coop_sync(c);
  r.pos = 624;
}

CN_HD uint32_t cn_rng_mix(uint32_t a, uint32_t b, uint32_t m) {
  const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return m ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// genrand twist.  Sequential semantics: key[i] <- f(key[i], key[i+1], key[(i+397) % 624]) for i = 0..623
// in order.  Chunks of up to 227 consecutive i are independent (every operand a chunk reads is either
// not yet overwritten or was produced >= 227 positions earlier), so each chunk is read by all lanes,
// synchronised, then written.  (A CTA-scope context uses chunks of 224; its other lanes only take the barriers.)
CN_HD void cn_rng_twist(CnRng& r, const CnCoop& c) {
  const int nl = c.nlanes < 224 ? c.nlanes : 224;
  for (int base = 0; base < 624; base += nl) {
    const int i = base + c.lane;
    const bool mine = c.lane < nl && i < 624;
    uint32_t v = 0;
    if (mine) {
      const int i1 = (i + 1 == 624) ? 0 : i + 1;
      const int im = (i + 397 < 624) ? i + 397 : i + 397 - 624;
      v = cn_rng_mix(r.key[i], r.key[i1], r.key[im]);
    }
    cn_coop_sync(c);
    if (mine) r.key[i] = v;
    cn_coop_sync(c);
  }
  r.pos = 0;
}

CN_HD uint32_t cn_rng_u32(CnRng& r, const CnCoop& c) {
  if (r.pos >= 624) cn_rng_twist(r, c);
  uint32_t y = r.key[r.pos++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

// tempered output / random_sample() at a word offset AHEAD of the stream position, without consuming
// (caller guarantees pos + off + 1 < 624: no twist inside the window)
CN_HD uint32_t cn_rng_peek_u32(const CnRng& r, int off) {
  uint32_t y = r.key[r.pos + off];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
CN_HD double cn_rng_peek_double(const CnRng& r, int off) {
  const uint32_t a = cn_rng_peek_u32(r, off) >> 5, b = cn_rng_peek_u32(r, off + 1) >> 6;
  return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}
CN_HD double cn_bcast_d(const CnCoop& c, double v, int src) {
#if defined(__CUDA_ARCH__)
  if (c.nlanes > 1) return __shfl_sync(cn_gmask(c), v, c.base + src);
#endif
  (void)src;
  return v;
}

CN_HD double cn_rng_double(CnRng& r, const CnCoop& c) {
  const uint32_t a = cn_rng_u32(r, c) >> 5, b = cn_rng_u32(r, c) >> 6;
  return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}

// np.random.uniform(lo, hi): lo + (hi - lo) * random_sample()   (two roundings, no fma)
CN_HD double cn_rng_uniform(CnRng& r, const CnCoop& c, double lo, double hi) {
  const double scale = hi - lo;
  const double u = cn_rng_double(r, c);
#if defined(__CUDA_ARCH__)
  return __dadd_rn(lo, __dmul_rn(scale, u));
#else
  return lo + scale * u;
#endif
}

// np.random.randint(low, high) of the legacy RandomState (int64 path, range < 2^32): rng = high - 1 - low;
// rng == 0 consumes NO draw; otherwise masked rejection on 32-bit outputs (smallest 2^k - 1 >= rng).
CN_HD int cn_rng_randint(CnRng& r, const CnCoop& c, int low, int high) {
  const uint32_t rng = (uint32_t)(high - 1 - low);
  if (rng == 0) return low;
  uint32_t mask = rng;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  uint32_t v;
  do { v = cn_rng_u32(r, c) & mask; } while (v > rng);
  return low + (int)v;
}
