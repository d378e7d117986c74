// tcgen05 / TMEM / TMA GEMM for the policy's 512-wide projections (sm_100a only).
//
//   C[M,N] = act( (A_hi + A_lo)[M,K] . (B_hi + B_lo)[N,K]^T * (1/scale_b) + bias[N] )
//
// "3xFP16" error-compensated product: every fp32 operand is split into two fp16 pieces
// (hi = rn(x), lo = rn(x - hi): 22 significand bits) and the three significant partial
// products  A_hi.B_hi + A_hi.B_lo + A_lo.B_hi  are accumulated by the 5th-gen tensor cores in
// an fp32 TMEM accumulator.  Measured against the fp32 reference policy this keeps the action
// mean / value within 2e-5 (same as a plain fp32 CUDA-core GEMM), where single-pass TF32/BF16
// would break the 1e-4 tolerance.  Weights are pre-scaled by 2^6 so their lo
// pieces stay in fp16's normal range; the epilogue undoes the power-of-two scale exactly.
//
// Tiling: one 128 x BN output tile per CTA (BN = 256 for the large per-human GEMMs, BN = 64 for the
// per-environment layers where M is only a few thousand rows and more CTAs matter more than tile
// efficiency), K in blocks of 64 fp16 (= one 128-byte swizzle atom), TMA->smem ring of 2 (BN=256,
// 96 KB per stage) or 4 (BN=64, 48 KB per stage) stages, accumulator 128 lanes x BN columns of TMEM.  Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
// tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld -> registers -> global, each warp owns
// the TMEM lane quadrant warp_id % 4).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define TC_BM 128
#define TC_BK 64
#define TC_A_TILE_BYTES (TC_BM * TC_BK * 2)           // 16 KB
#define TC_EPI_WARPS 8          // two per TMEM lane quadrant, each draining half of the tile's columns
#define TC_THREADS (64 + 32 * TC_EPI_WARPS)

template <int BN>
struct TcCfg {
  static constexpr int kStages = (BN >= 256) ? 2 : 4;
  static constexpr int kBTile = BN * TC_BK * 2;
  static constexpr int kStageBytes = 2 * TC_A_TILE_BYTES + 2 * kBTile;
  // + per-epilogue-warp 4 KB staging slab for the TMA stores (fp32 32x32 box, or fp16 hi + lo 32x32 boxes)
  static constexpr int kStageOutOff = kStages * kStageBytes + 2048;       // barriers (256) + bias (1024), 1024-aligned
  static constexpr int kSmemBytes = kStageOutOff + TC_EPI_WARPS * 4096 + 1024 /*align slack*/;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
};

struct TcEpilogue {
  const float* bias;     // [N] or null
  float inv_scale;       // 1 / scale_b
  int act;               // CN_ACT_* applied to columns [act_lo, act_hi)
  int act_lo, act_hi;
  float* c32;            // fp32 output [M, ldc] or null
  int ldc;
  __half* out_hi;        // split fp16 output [M, ldh] or null (A operand of the next GEMM)
  __half* out_lo;
  int ldh;
  const int* m_ptr;      // optional device-side row count (compacted rows); tiles past it exit
  int dbg_nostore;       // diagnostic: run the epilogue arithmetic but skip the global stores
  const int* m0_ptr;     // optional device-side first row: the launch covers rows [*m0_ptr, *m_ptr) (row chunks)
  // --- PPO update path (cn_update.cuh); all zero / null in the rollout ---
  const float* inv_scale_a;   // optional device scalars: the result is also multiplied by *inv_scale_a * *inv_scale_b
  const float* inv_scale_b;   // (dynamic power-of-two operand scales chosen from the tensors' amax)
  int ksplit;                 // > 1: the K range is cut into `ksplit` slices, every slice is its own tile and ADDS its
                              // partial product into a zero-initialised fp32 C with TMA reduce (bias from slice 0 only,
                              // no activation): wgrad has K = #rows (hundreds of thousands) and a tiny output
};

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// TMA store of a shared-memory box to global memory (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
// TMA reduce (add) of a shared-memory fp32 box into global memory: split-K partial sums
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (=1, unused for swizzled K-major) |
//   [32,46) SBO >> 4 (8 rows * 128 B = 1024 B) | [46,48) version = 1 | [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, K-major
__device__ __forceinline__ uint32_t make_idesc(int bn) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(bn >> 3) << 17) |
         ((uint32_t)(TC_BM >> 4) << 24);
}
// tanh(x) = 1 - 2 / (exp(2x) + 1) with the accurate expf and an IEEE division: ~2e-7 absolute error
// (cheaper than tanhf in the 256-column epilogue, and far inside the 1e-4 policy tolerance)
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = expf(2.0f * x);
  return 1.0f - 2.0f / (e + 1.0f);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace tc

// Persistent kernel: grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x, blockIdx.x + grid, ...
// (n fastest, so CTAs running at the same time share A rows in L2).  The row count may live on the
// device (ep.m_ptr, compacted human rows): no CTA is ever launched for an empty tile.  Two TMEM
// accumulators (2 x BN columns) let the epilogue of tile i overlap the mainloop of tile i + 1.
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
cn_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                  const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                  const __grid_constant__ CUtensorMap map_c32, const __grid_constant__ CUtensorMap map_ohi,
                  const __grid_constant__ CUtensorMap map_olo, int M, int N, int K, TcEpilogue ep) {
  cn_pdl_trigger();                                 // PDL: the successor may be scheduled while this grid runs
  constexpr int TC_STAGES = TcCfg<BN>::kStages;
  constexpr int TC_B_TILE_BYTES = TcCfg<BN>::kBTile;
  constexpr int TC_STAGE_BYTES = TcCfg<BN>::kStageBytes;
  constexpr int TC_BN = BN;
  constexpr uint32_t TMEM_COLS = 2 * TcCfg<BN>::kTmemCols;       // two accumulators
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t raw = tc::smem_u32(tc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* base_ptr = tc_smem_raw + (base - raw);
  const uint32_t bar_base = base + TC_STAGES * TC_STAGE_BYTES;  // barriers after the operand ring
  // full[s] = +8 s ; empty[s] = +32 + 8 s ; tmem_full[b] = +64 + 8 b ; tmem_empty[b] = +80 + 8 b ; tmem ptr at +96
  const uint32_t bar_full = bar_base, bar_empty = bar_base + 32, bar_tfull = bar_base + 64, bar_tempty = bar_base + 80;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(base_ptr + TC_STAGES * TC_STAGE_BYTES + 96);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = K / TC_BK;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      tc::mbar_init(bar_full + 8 * s, 1);         // producer's arrive.expect_tx
      tc::mbar_init(bar_empty + 8 * s, 1);        // one tcgen05.commit
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(bar_tfull + 8 * b, 1);        // accumulator b complete (tcgen05.commit)
      tc::mbar_init(bar_tempty + 8 * b, TC_EPI_WARPS);   // accumulator b drained (one arrive per epilogue warp)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ahi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_alo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_blo) : "memory");
    if (ep.c32) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c32) : "memory");
    if (ep.out_hi) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ohi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_olo) : "memory");
    }
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
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the predecessor's tail;
  // global memory written by it (row counts, operands, bias) is only touched after this wait
  cn_pdl_wait();
  if (ep.m_ptr) { const int mc = *ep.m_ptr; M = mc < M ? mc : M; }
  int m_lo = ep.m0_ptr ? *ep.m0_ptr : 0;
  if (m_lo > M) m_lo = M;
  const int n_ntiles = N / BN;
  const int n_mn = ((M - m_lo + TC_BM - 1) / TC_BM) * n_ntiles;
  const int ksplit = ep.ksplit > 1 ? ep.ksplit : 1;
  const int kb_per = (num_kb + ksplit - 1) / ksplit;                 // k-blocks per slice (the last may be shorter)
  const int n_tiles = n_mn * ksplit;                                 // CTAs beyond it run zero tiles and tear down
  float inv_scale = ep.inv_scale;
  if (ep.inv_scale_a) inv_scale *= __ldg(ep.inv_scale_a);
  if (ep.inv_scale_b) inv_scale *= __ldg(ep.inv_scale_b);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;                                            // running k-block counter across tiles
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mn = tile % n_mn, ks = tile / n_mn;
        const int m0 = m_lo + (mn / n_ntiles) * TC_BM, n0 = (mn % n_ntiles) * TC_BN;
        const int kb0 = ks * kb_per, kb1 = (kb0 + kb_per < num_kb) ? kb0 + kb_per : num_kb;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
          tc::mbar_wait(bar_empty + 8 * s, ph ^ 1u);              // slot free (first pass returns immediately)
          const uint32_t full = bar_full + 8 * s;
          tc::mbar_expect_tx(full, TC_STAGE_BYTES);
          const uint32_t st = base + s * TC_STAGE_BYTES;
          tc::tma_load_2d(st, &map_ahi, full, kb * TC_BK, m0);
          tc::tma_load_2d(st + TC_A_TILE_BYTES, &map_alo, full, kb * TC_BK, m0);
          tc::tma_load_2d(st + 2 * TC_A_TILE_BYTES, &map_bhi, full, kb * TC_BK, n0);
          tc::tma_load_2d(st + 2 * TC_A_TILE_BYTES + TC_B_TILE_BYTES, &map_blo, full, kb * TC_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(BN);
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
        const uint32_t ab = ti & 1u, aph = (ti >> 1) & 1u;
        tc::mbar_wait(bar_tempty + 8 * ab, aph ^ 1u);             // epilogue drained this accumulator
        tc::tcgen05_fence_after();
        const uint32_t tmem_acc = tmem_base + ab * TcCfg<BN>::kTmemCols;
        const int ks = tile / n_mn;
        const int kb0 = ks * kb_per, kb1 = (kb0 + kb_per < num_kb) ? kb0 + kb_per : num_kb;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
          tc::mbar_wait(bar_full + 8 * s, ph);                    // TMA bytes landed
          tc::tcgen05_fence_after();
          const uint32_t st = base + s * TC_STAGE_BYTES;
          const uint32_t a_hi = st, a_lo = st + TC_A_TILE_BYTES, b_hi = st + 2 * TC_A_TILE_BYTES,
                         b_lo = st + 2 * TC_A_TILE_BYTES + TC_B_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint32_t koff = k * 32;                         // 16 fp16 = 32 bytes inside the swizzle atom
            const uint64_t dah = tc::make_desc(a_hi + koff), dal = tc::make_desc(a_lo + koff);
            const uint64_t dbh = tc::make_desc(b_hi + koff), dbl = tc::make_desc(b_lo + koff);
            tc::mma_f16(tmem_acc, dah, dbh, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
            tc::mma_f16(tmem_acc, dah, dbl, idesc, 1u);
            tc::mma_f16(tmem_acc, dal, dbh, idesc, 1u);
          }
          tc::mma_commit(bar_empty + 8 * s);                      // frees the smem slot when the MMAs retire
        }
        tc::mma_commit(bar_tfull + 8 * ab);                       // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                                       // TMEM lane quadrant this warp may access
    const int chalf = (warp - 2) >> 2;                            // which half of the tile's 32-column chunks
    constexpr int kChunks = TC_BN / 32 / (TC_EPI_WARPS / 4);
    const uint32_t stage_out = base + TcCfg<BN>::kStageOutOff + (uint32_t)(warp - 2) * 4096u;   // 1024-byte aligned
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const int mn = tile % n_mn, ks = tile / n_mn;
      const int m0 = m_lo + (mn / n_ntiles) * TC_BM, n0 = (mn % n_ntiles) * TC_BN;
      const uint32_t ab = ti & 1u, aph = (ti >> 1) & 1u;
      tc::mbar_wait(bar_tfull + 8 * ab, aph);
      tc::tcgen05_fence_after();
      const uint32_t tmem_acc = tmem_base + ab * TcCfg<BN>::kTmemCols;
      // stage this tile's bias slice in shared memory once (epilogue warps only: named barrier 1)
      float* bias_s = reinterpret_cast<float*>(base_ptr + TC_STAGES * TC_STAGE_BYTES + 256);
      {
        const int et = threadIdx.x - 64;                          // index within the epilogue warps
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");   // previous tile's readers are done
        for (int c = et; c < TC_BN; c += 32 * TC_EPI_WARPS) bias_s[c] = (ep.bias && ks == 0) ? __ldg(ep.bias + n0 + c) : 0.0f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
      }
      // activation is uniform over the tile unless the [act_lo, act_hi) window cuts through it
      const bool act_full = ep.act_lo <= n0 && ep.act_hi >= n0 + TC_BN;
      const bool act_none = ep.act == 0 || ep.act_hi <= n0 || ep.act_lo >= n0 + TC_BN;
      const int act_mode = act_none ? 0 : (act_full ? ep.act : 3);
#pragma unroll 1
      for (int c = chalf * kChunks; c < (chalf + 1) * kChunks; ++c) {
        uint32_t r[32];
        tc::tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
        const int nb = n0 + c * 32;
        const float* bs = bias_s + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), inv_scale, bs[j]);
        if (act_mode == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        } else if (act_mode == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tc::fast_tanh(v[j]);
        } else if (act_mode == 3) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (nb + j >= ep.act_lo && nb + j < ep.act_hi) v[j] = (ep.act == 1) ? fmaxf(v[j], 0.0f) : tc::fast_tanh(v[j]);
          }
        }
        // Results leave through shared memory and the TMA store engine: the thread = row layout of
        // tcgen05.ld would make every direct STG.128 touch 32 different cache lines (measured: the stores
        // were ~1/3 of the QKV kernel).  Each warp stages its 32 x 32 slab in a private swizzled buffer
        // (conflict-free 16-byte st.shared) and one lane issues the bulk tensor store; rows past the
        // tensor's extent are clipped by the TMA unit.
        if (!ep.dbg_nostore) {
          if (ep.c32) {
            if (lane == 0) tc::tma_store_wait_read();            // previous slab has left the staging buffer
            __syncwarp();
            const uint32_t rowb = stage_out + (uint32_t)lane * 128u;
#pragma unroll
            for (int j = 0; j < 8; ++j)                          // SWIZZLE_128B: 16-byte chunk j of row r sits at j ^ (r & 7)
              tc::st_shared_v4(rowb + (uint32_t)((j ^ (lane & 7)) << 4), __float_as_uint(v[4 * j]),
                               __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (ksplit > 1) tc::tma_reduce_add_2d(&map_c32, stage_out, nb, m0 + q * 32);
              else tc::tma_store_2d(&map_c32, stage_out, nb, m0 + q * 32);
              tc::tma_store_commit();
            }
          }
          if (ep.out_hi) {
            if (lane == 0) tc::tma_store_wait_read();
            __syncwarp();
            const uint32_t rowh = stage_out + (uint32_t)lane * 64u, rowl = rowh + 2048u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t ph[4], pl[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float x0 = fminf(fmaxf(v[8 * j + 2 * t], -65504.0f), 65504.0f);
                const float x1 = fminf(fmaxf(v[8 * j + 2 * t + 1], -65504.0f), 65504.0f);
                const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
                const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
                ph[t] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                pl[t] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
              }
              const uint32_t sw = (uint32_t)((j ^ ((lane >> 1) & 3)) << 4);   // SWIZZLE_64B: chunk j of row r at j ^ ((r >> 1) & 3)
              tc::st_shared_v4(rowh + sw, ph[0], ph[1], ph[2], ph[3]);
              tc::st_shared_v4(rowl + sw, pl[0], pl[1], pl[2], pl[3]);
            }
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tc::tma_store_2d(&map_ohi, stage_out, nb, m0 + q * 32);
              tc::tma_store_2d(&map_olo, stage_out + 2048u, nb, m0 + q * 32);
              tc::tma_store_commit();
            }
          }
        }
      }
      // this warp is done reading the accumulator: hand it back to the MMA issuer
      tc::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tempty + 8 * ab) : "memory");
    }
    if (lane == 0) tc::tma_store_wait_all();                    // bulk stores of this thread are complete before exit
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// fp32 -> (hi, lo) fp16 split with an exact power-of-two pre-scale (weights at finalize time,
// and the generic "split this activation" helper).
__global__ void cn_split_f16_kernel(const float* __restrict__ src, float scale, __half* __restrict__ hi,
                                    __half* __restrict__ lo, size_t count) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float x = fminf(fmaxf(src[i] * scale, -65504.0f), 65504.0f);
  const __half h = __float2half_rn(x);
  hi[i] = h;
  lo[i] = __float2half_rn(x - __half2float(h));
}
