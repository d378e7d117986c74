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
};

CN_HD void cn_rng_seed(CnRng& r, uint32_t seed) {
  for (int i = 0; i < 624; ++i) {
    r.key[i] = seed;
    seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)(i + 1);
  }
  r.pos = 624;
}

CN_HD void cn_rng_twist(CnRng& r) {
  const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
  int i;
  uint32_t y;
  for (i = 0; i < 624 - 397; ++i) {
    y = (r.key[i] & UPPER) | (r.key[i + 1] & LOWER);
    r.key[i] = r.key[i + 397] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  }
  for (; i < 623; ++i) {
    y = (r.key[i] & UPPER) | (r.key[i + 1] & LOWER);
    r.key[i] = r.key[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  }
  y = (r.key[623] & UPPER) | (r.key[0] & LOWER);
  r.key[623] = r.key[396] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  r.pos = 0;
}

CN_HD uint32_t cn_rng_u32(CnRng& r) {
  if (r.pos >= 624) cn_rng_twist(r);
  uint32_t y = r.key[r.pos++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

CN_HD double cn_rng_double(CnRng& r) {
  const uint32_t a = cn_rng_u32(r) >> 5, b = cn_rng_u32(r) >> 6;
  return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}

// np.random.uniform(lo, hi): lo + (hi - lo) * random_sample()   (two roundings, no fma)
CN_HD double cn_rng_uniform(CnRng& r, double lo, double hi) {
  const double scale = hi - lo;
  const double u = cn_rng_double(r);
#if defined(__CUDA_ARCH__)
  return __dadd_rn(lo, __dmul_rn(scale, u));
#else
  return lo + scale * u;
#endif
}
