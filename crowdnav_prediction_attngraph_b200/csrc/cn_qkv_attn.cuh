// Fused  QKV projection -> human-human attention  for the rollout policy (sm_100a only).
//
// replaces (reference): q/k/v_linear + nn.MultiheadAttention's in_proj and attention product of SpatialEdgeSelfAttn
// (rl/networks/selfAttn_srnn_temp_node.py:63-91), i.e. the two stages "qkv_gemm" + "hh_attention" of cn_policy_act.
//
// The unfused path writes the [Mc, 1536] fp32 projections to HBM (110 MB per step at 4096 envs) only for a second
// kernel to read them back.  Here one output tile holds everything one attention head needs:
//
//   tile (t, h):  rows  = the compacted human rows [m0_t, m0_t + cnt_t) of a run of WHOLE environments (cnt_t <= 128;
//                         the tile table is built on the device every step by cn_qkv_tiles_kernel, TMA loads may
//                         start at any row),
//                 cols  = [Q_h | K_h | V_h], 3 x 64 columns of the folded projection, whose weight rows are stored
//                         head-major (192 consecutive B rows per head) at finalize time,
//
// so the tcgen05 accumulator (128 lanes x 192 TMEM columns, double buffered) is drained straight into the attention:
// K_h and V_h of the tile's 128 rows go to shared memory (64 KB, 16-byte chunks XOR-swizzled by row), Q_h stays in
// the registers of the thread that owns the row (tcgen05.ld: thread = row), which runs the soft-max over the keys of
// its own environment and writes the 64 output columns of head h as the fp16 (hi, lo) A operand of the output
// projection.  Neither Q, K nor V ever reach global memory.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM allocator + tcgen05.mma issuer, warps 2..9 epilogue
// (two per TMEM lane quadrant: the first takes Q and the low half of K, the second the rest of K and V; the first
// then runs the attention of its 32 rows while the MMAs of the next tile fill the other accumulator).
#pragma once
#include "cn_gemm_tc.cuh"

#define QA_BN 192
// K blocks of 32 fp16 (64-byte rows, SWIZZLE_64B) in a 4-deep ring: the fp32 K | V staging leaves 160 KB for operands,
// which is only TWO 80 KB stages of the usual 64-wide blocks -- one block in flight while the other is consumed, and
// the mainloop then waits on L2 latency (measured: 111 us for the fused kernel against 74 us of GEMM at the same 2-stage
// depth).  Half-width blocks keep three loads in flight with the same bytes.
#define QA_BK 32
#define QA_STAGES 4
#define QA_A_TILE_BYTES (TC_BM * QA_BK * 2)                              // 8 KB
#define QA_B_TILE_BYTES (QA_BN * QA_BK * 2)                              // 12 KB
#define QA_STAGE_BYTES (2 * QA_A_TILE_BYTES + 2 * QA_B_TILE_BYTES)       // 40 KB
#define QA_KV_BYTES (TC_BM * 128 * 4)                                    // K | V of 128 rows, fp32: 64 KB
#define QA_MISC_OFF (QA_STAGES * QA_STAGE_BYTES)                         // barriers (256 B) + bias (768 B)
#define QA_KV_OFF (QA_MISC_OFF + 1024)
#define QA_SMEM_BYTES (QA_KV_OFF + QA_KV_BYTES + 1024 /* alignment slack */)
#define QA_THREADS 320
#define QA_MAX_ENVS 8192

// K-major SWIZZLE_64B shared-memory matrix descriptor: rows of 64 bytes, 8-row groups 512 bytes apart (SBO), layout 4
__device__ __forceinline__ uint64_t qa_desc64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}

__device__ __forceinline__ float4 qa_lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// head-major copy of the folded QKV weight / bias: dst row h * 192 + s * 64 + d <- src row s * 512 + h * 64 + d
__global__ void cn_head_major_kernel(const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ wh,
                                     float* __restrict__ bh) {
  const int ro = blockIdx.x, h = ro / 192, s = (ro % 192) / 64, d = ro % 64;
  const int ri = s * 512 + h * 64 + d;
  for (int k = threadIdx.x; k < 512; k += blockDim.x) wh[(size_t)ro * 512 + k] = w[(size_t)ri * 512 + k];
  if (threadIdx.x == 0) bh[ro] = b[ri];
}

// Tile table of the fused kernel: tab[0] = number of row tiles, tab[2 + 2 t] = first row, tab[3 + 2 t] = rows of tile
// t.  A tile is a maximal run of whole environments with at most 128 rows (every environment has 1 <= n_e <= H <= 128
// rows).  One CTA: row_start -> shared memory, every thread finds how many environments fit behind "its" environment
// (binary search), one thread then walks the chain (~N n / 128 dependent shared-memory loads).  Runs on the side
// stream next to the robot branch, off the critical path.
__global__ void __launch_bounds__(1024) cn_qkv_tiles_kernel(const int* __restrict__ row_start, int N, int* __restrict__ tab) {
  __shared__ int rs[QA_MAX_ENVS + 1];
  __shared__ unsigned char step[QA_MAX_ENVS];
  for (int i = threadIdx.x; i <= N; i += blockDim.x) rs[i] = row_start[i];
  __syncthreads();
  for (int e = threadIdx.x; e < N; e += blockDim.x) {
    int lo = e + 1, hi = e + 128 < N ? e + 128 : N;            // largest e2 in [lo, hi] with rs[e2] - rs[e] <= 128
    const int lim = rs[e] + 128;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (rs[mid] <= lim) lo = mid; else hi = mid - 1;
    }
    step[e] = (unsigned char)(lo - e);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int e = 0; e < N; ++t) {
      const int e2 = e + step[e];
      tab[2 + 2 * t] = rs[e];
      tab[3 + 2 * t] = rs[e2] - rs[e];
      e = e2;
    }
    tab[0] = t;
  }
}

__global__ void __launch_bounds__(QA_THREADS, 1)
cn_qkv_attn_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                   const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                   const float* __restrict__ bias /* [8][192] head-major */, float inv_scale, const int* tile_tab,
                   const int* row_start, const int* row_env, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                   int dbg /* diagnostics: 1 = skip the attention, 2 = skip the K / V staging, 4 = skip the key loop, 8 = skip the stores */) {
  cn_pdl_trigger();
  constexpr int K = 512, NUM_KB = K / QA_BK;
  constexpr uint32_t TMEM_COLS = 512;                           // two 192-column accumulators at column 0 and 256
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t raw = tc::smem_u32(tc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = tc_smem_raw + (base - raw);
  const uint32_t bar_base = base + QA_MISC_OFF;
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 32, bar_tfull = bar_base + 64, bar_tempty = bar_base + 80;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(base_ptr + QA_MISC_OFF + 96);
  float* bias_s = reinterpret_cast<float*>(base_ptr + QA_MISC_OFF + 256);
  float* kv = reinterpret_cast<float*>(base_ptr + QA_KV_OFF);   // [128][32 chunks of 16 B]: K chunks 0..15, V 16..31
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < QA_STAGES; ++s) {
      tc::mbar_init(bar_full + 8 * s, 1);
      tc::mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(bar_tfull + 8 * b, 1);
      tc::mbar_init(bar_tempty + 8 * b, 8);                     // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ahi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_alo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_blo) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_ptr_smem)),
                 "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  cn_pdl_wait();                                                // tile table, operands: written by earlier kernels
  const int n_rt = cn_ld_after_wait(tile_tab);                  // row tiles
  const int n_tiles = n_rt * 8;                                 // x 8 heads; head fastest: the 8 CTAs running one row
                                                                // tile at the same time share its A rows in L2
  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = cn_ld_after_wait(tile_tab + 2 + 2 * (tile >> 3)), n0 = (tile & 7) * QA_BN;
        for (int kb = 0; kb < NUM_KB; ++kb, ++it) {
          const uint32_t s = it % QA_STAGES, ph = (it / QA_STAGES) & 1u;
          tc::mbar_wait(bar_empty + 8 * s, ph ^ 1u);
          const uint32_t full = bar_full + 8 * s;
          tc::mbar_expect_tx(full, QA_STAGE_BYTES);
          const uint32_t st = base + s * QA_STAGE_BYTES;
          tc::tma_load_2d(st, &map_ahi, full, kb * QA_BK, m0);
          tc::tma_load_2d(st + QA_A_TILE_BYTES, &map_alo, full, kb * QA_BK, m0);
          tc::tma_load_2d(st + 2 * QA_A_TILE_BYTES, &map_bhi, full, kb * QA_BK, n0);
          tc::tma_load_2d(st + 2 * QA_A_TILE_BYTES + QA_B_TILE_BYTES, &map_blo, full, kb * QA_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(QA_BN);
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
        const uint32_t ab = ti & 1u, aph = (ti >> 1) & 1u;
        tc::mbar_wait(bar_tempty + 8 * ab, aph ^ 1u);
        tc::tcgen05_fence_after();
        const uint32_t tmem_acc = tmem_base + ab * 256u;
        for (int kb = 0; kb < NUM_KB; ++kb, ++it) {
          const uint32_t s = it % QA_STAGES, ph = (it / QA_STAGES) & 1u;
          tc::mbar_wait(bar_full + 8 * s, ph);
          tc::tcgen05_fence_after();
          const uint32_t st = base + s * QA_STAGE_BYTES;
          const uint32_t a_hi = st, a_lo = st + QA_A_TILE_BYTES, b_hi = st + 2 * QA_A_TILE_BYTES,
                         b_lo = st + 2 * QA_A_TILE_BYTES + QA_B_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < QA_BK / 16; ++k) {
            const uint32_t koff = k * 32;                         // 16 fp16 = 32 bytes inside the 64-byte swizzle atom
            const uint64_t dah = qa_desc64(a_hi + koff), dal = qa_desc64(a_lo + koff);
            const uint64_t dbh = qa_desc64(b_hi + koff), dbl = qa_desc64(b_lo + koff);
            tc::mma_f16(tmem_acc, dah, dbh, idesc, (kb != 0 || k != 0) ? 1u : 0u);
            tc::mma_f16(tmem_acc, dah, dbl, idesc, 1u);
            tc::mma_f16(tmem_acc, dal, dbh, idesc, 1u);
          }
          tc::mma_commit(bar_empty + 8 * s);
        }
        tc::mma_commit(bar_tfull + 8 * ab);
      }
    }
  } else {
    // ===================== epilogue + attention (warps 2..9) =====================
    const int q = warp & 3;                                     // TMEM lane quadrant of this warp
    const int role = (warp - 2) >> 2;                           // 0: Q + K[0..31] and the attention, 1: K[32..63] + V
    const int et = threadIdx.x - 64;
    const int r = q * 32 + lane;                                // row of the tile owned by this thread
    const uint32_t kv_s = tc::smem_u32(kv);
    const uint32_t my_row = kv_s + (uint32_t)r * 512u;
    const int rx = r & 15;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const int rt = tile >> 3, head = tile & 7;
      const int m0 = cn_ld_after_wait(tile_tab + 2 + 2 * rt), cnt = cn_ld_after_wait(tile_tab + 3 + 2 * rt);
      const uint32_t ab = ti & 1u, aph = (ti >> 1) & 1u;
      tc::mbar_wait(bar_tfull + 8 * ab, aph);
      tc::tcgen05_fence_after();
      const uint32_t tmem_acc = tmem_base + ab * 256u + ((uint32_t)(q * 32) << 16);
      asm volatile("bar.sync 1, 256;" ::: "memory");            // the previous tile's attention has left K / V / bias
      if (et < QA_BN) bias_s[et] = __ldg(bias + head * QA_BN + et);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // both warps of a quadrant take Q_h of their row (scaled by 1/sqrt(64)): they split the keys of the attention
      float qv[64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tc::tmem_ld32(tmem_acc + (uint32_t)(c * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) qv[c * 32 + j] = fmaf(__uint_as_float(v[j]), inv_scale, bias_s[c * 32 + j]) * 0.125f;
      }
      {
        // role 0: columns 64..95 (K 0..31); role 1: columns 96..127 (K 32..63), 128..159 and 160..191 (V)
        const int nchunk = role == 0 ? 1 : 3, cbase = role == 0 ? 64 : 96;
#pragma unroll 1
        for (int cc = 0; cc < ((dbg & 2) ? 0 : nchunk); ++cc) {
          const int col0 = cbase + 32 * cc;                     // accumulator column of this 32-wide chunk
          uint32_t v[32];
          tc::tmem_ld32(tmem_acc + (uint32_t)col0, v);
          const int ch0 = (col0 - 64) >> 2;                     // first 16-byte chunk inside the K | V row (0..31)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int ch = ch0 + j;
            const uint32_t pos = (uint32_t)((ch & 16) | ((ch & 15) ^ rx));
            tc::st_shared_v4(my_row + (pos << 4),
                             __float_as_uint(fmaf(__uint_as_float(v[4 * j]), inv_scale, bias_s[col0 + 4 * j])),
                             __float_as_uint(fmaf(__uint_as_float(v[4 * j + 1]), inv_scale, bias_s[col0 + 4 * j + 1])),
                             __float_as_uint(fmaf(__uint_as_float(v[4 * j + 2]), inv_scale, bias_s[col0 + 4 * j + 2])),
                             __float_as_uint(fmaf(__uint_as_float(v[4 * j + 3]), inv_scale, bias_s[col0 + 4 * j + 3])));
          }
        }
      }
      // the accumulator is drained: hand it back to the MMA issuer before the attention starts
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tempty + 8 * ab) : "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");            // K, V of all 128 rows are in shared memory
      // ---- attention of row r over the keys of its environment.  The two warps of a quadrant sit on the same
      // scheduler (they hide each other's shared-memory latency): both compute the scores and the soft-max weights
      // of their row, each accumulates and writes one half (32) of the head's 64 output columns.
      if (r < cnt && !(dbg & 1)) {
        const int e = row_env[m0 + r];
        const int j0 = row_start[e] - m0, j1 = row_start[e + 1] - m0;      // keys: the rows of the same environment
        float mx = -INFINITY, l = 0.0f, o[32];
#pragma unroll
        for (int d = 0; d < 32; ++d) o[d] = 0.0f;
        for (int j = j0; j < ((dbg & 4) ? j0 : j1); ++j) {
          const uint32_t krow = kv_s + (uint32_t)j * 512u;
          const uint32_t jx = (uint32_t)(j & 15);
          float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {                       // 8 loads in flight, then their 32 FMAs
            float4 kk[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) kk[c] = qa_lds128(krow + ((((uint32_t)(hb * 8 + c)) ^ jx) << 4));
            // consumed LAST-loaded first: the (ordered, volatile) loads must all be issued before the first FMA can
            // go, so ptxas keeps the eight of them in flight instead of recycling four temporaries (seen in SASS)
#pragma unroll
            for (int c = 7; c >= 0; --c) {
              const int d = 4 * (hb * 8 + c);
              s0 = fmaf(qv[d], kk[c].x, s0); s1 = fmaf(qv[d + 1], kk[c].y, s1);
              s2 = fmaf(qv[d + 2], kk[c].z, s2); s3 = fmaf(qv[d + 3], kk[c].w, s3);
            }
          }
          float4 vv[8];                                         // this warp's half of V_j, requested before the exp
#pragma unroll
          for (int c = 0; c < 8; ++c) vv[c] = qa_lds128(krow + 256u + ((((uint32_t)(role * 8 + c)) ^ jx) << 4));
          const float sc = (s0 + s1) + (s2 + s3);
          if (sc > mx) {                                        // online soft-max: rescale only when the maximum moves
            const float f = expf(mx - sc);                      // exp(-inf) = 0 on the first key
            l *= f;
#pragma unroll
            for (int d = 0; d < 32; ++d) o[d] *= f;
            mx = sc;
          }
          const float pw = expf(sc - mx);
          l += pw;
#pragma unroll
          for (int c = 7; c >= 0; --c) {
            o[4 * c] = fmaf(pw, vv[c].x, o[4 * c]); o[4 * c + 1] = fmaf(pw, vv[c].y, o[4 * c + 1]);
            o[4 * c + 2] = fmaf(pw, vv[c].z, o[4 * c + 2]); o[4 * c + 3] = fmaf(pw, vv[c].w, o[4 * c + 3]);
          }
        }
        const float inv = 1.0f / l;
        const size_t ob = (size_t)(m0 + r) * 512 + head * 64 + role * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t ph[4], pl[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float x0 = fminf(fmaxf(o[8 * c + 2 * t] * inv, -65504.0f), 65504.0f);
            const float x1 = fminf(fmaxf(o[8 * c + 2 * t + 1] * inv, -65504.0f), 65504.0f);
            const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
            const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
            ph[t] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
            pl[t] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
          }
          if (!(dbg & 8)) {
            *reinterpret_cast<uint4*>(out_hi + ob + 8 * c) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
            *reinterpret_cast<uint4*>(out_lo + ob + 8 * c) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
          }
        }
      }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}
