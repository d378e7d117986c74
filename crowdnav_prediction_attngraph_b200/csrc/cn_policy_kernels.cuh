// Device kernels of the attention-graph policy forward (rollout / infer=True path):
//   rl/networks/selfAttn_srnn_temp_node.py:360-449 (selfAttn_merge_SRNN.forward)
//   rl/networks/selfAttn_srnn_temp_node.py:63-91   (SpatialEdgeSelfAttn, nn.MultiheadAttention 8 heads)
//   rl/networks/selfAttn_srnn_temp_node.py:145-223 (EdgeAttention_M)
//   rl/networks/selfAttn_srnn_temp_node.py:262-285 + srnn_model.py:35-47 (EndRNN / GRU step)
//   rl/networks/distributions.py:76-95,36-44       (DiagGaussian / FixedNormal)
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

enum { CN_ACT_NONE = 0, CN_ACT_RELU = 1, CN_ACT_TANH = 2 };

// ------------------------------------------------------------------------------------------
// fp32 CUDA-core GEMM:  C[M,N] = act(A[M,K] * W[N,K]^T + bias[N]),  act on columns [act_lo, act_hi).
// 128x128x16 tiles, 256 threads, 8x8 register tile per thread, register-prefetched double
// buffering through shared memory.  K % 16 == 0 (buffers are zero padded), M and N ragged.
#define CN_GEMM_BM 128
#define CN_GEMM_BN 128
#define CN_GEMM_BK 16
#define CN_GEMM_PAD 4

__device__ __forceinline__ float cn_apply_act(float v, int act) {
  if (act == CN_ACT_RELU) return v > 0.0f ? v : 0.0f;
  if (act == CN_ACT_TANH) return tanhf(v);
  return v;
}

__global__ void __launch_bounds__(256) cn_gemm_f32_kernel(const float* __restrict__ A, int lda,
                                                          const float* __restrict__ W, int ldw,
                                                          const float* __restrict__ bias, float* __restrict__ Cout,
                                                          int ldc, int M, int N, int K, int act, int act_lo,
                                                          int act_hi, const int* __restrict__ m_ptr,
                                                          __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  // m_ptr: optional device-side row count (compacted human rows); tiles past it exit immediately
  if (m_ptr) { const int mc = *m_ptr; M = mc < M ? mc : M; }
  if ((int)(blockIdx.y * CN_GEMM_BM) >= M) return;
  __shared__ __align__(16) float As[2][CN_GEMM_BK][CN_GEMM_BM + CN_GEMM_PAD];
  __shared__ __align__(16) float Bs[2][CN_GEMM_BK][CN_GEMM_BN + CN_GEMM_PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * CN_GEMM_BM, n0 = blockIdx.x * CN_GEMM_BN;
  const int tx = tid & 15, ty = tid >> 4;
  // global -> register staging: each thread moves two float4 of A and two of W per k-tile
  const int lrow = tid >> 2;          // 0..63
  const int lk = (tid & 3) * 4;       // 0,4,8,12
  float4 ra[2], rb[2];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int m = m0 + lrow + i * 64;
      ra[i] = (m < M) ? *reinterpret_cast<const float4*>(A + (size_t)m * lda + k0 + lk) : make_float4(0, 0, 0, 0);
      const int n = n0 + lrow + i * 64;
      rb[i] = (n < N) ? *reinterpret_cast<const float4*>(W + (size_t)n * ldw + k0 + lk) : make_float4(0, 0, 0, 0);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = lrow + i * 64;
      As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y; As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
      Bs[buf][lk + 0][r] = rb[i].x; Bs[buf][lk + 1][r] = rb[i].y; Bs[buf][lk + 2][r] = rb[i].z; Bs[buf][lk + 3][r] = rb[i].w;
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  const int nk = K / CN_GEMM_BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * CN_GEMM_BK);
#pragma unroll
    for (int k = 0; k < CN_GEMM_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int nb = n0 + (jj == 0 ? tx * 4 : 64 + tx * 4);
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = nb + j;
        float x = acc[i][jj * 4 + j];
        if (n < N) {
          if (bias) x += bias[n];
          if (n >= act_lo && n < act_hi) x = cn_apply_act(x, act);
        }
        v[j] = x;
      }
      if (out_hi) {    // (hi, lo) fp16 split for the tensor-core consumer; same leading dimension
        if (nb + 3 < N && ((ldc & 3) == 0)) {
          uint32_t ph[2], pl[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const float c0 = fminf(fmaxf(v[2 * t], -65504.0f), 65504.0f), c1 = fminf(fmaxf(v[2 * t + 1], -65504.0f), 65504.0f);
            const __half h0 = __float2half_rn(c0), h1 = __float2half_rn(c1);
            const __half l0 = __float2half_rn(c0 - __half2float(h0)), l1 = __float2half_rn(c1 - __half2float(h1));
            ph[t] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
            pl[t] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
          }
          *reinterpret_cast<uint2*>(out_hi + (size_t)m * ldc + nb) = make_uint2(ph[0], ph[1]);
          *reinterpret_cast<uint2*>(out_lo + (size_t)m * ldc + nb) = make_uint2(pl[0], pl[1]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (nb + j < N) {
              const float c = fminf(fmaxf(v[j], -65504.0f), 65504.0f);
              const __half hh = __float2half_rn(c);
              out_hi[(size_t)m * ldc + nb + j] = hh;
              out_lo[(size_t)m * ldc + nb + j] = __float2half_rn(c - __half2float(hh));
            }
        }
        if (!Cout) continue;
      }
      float* dst = Cout + (size_t)m * ldc + nb;
      if (nb + 3 < N && ((ldc & 3) == 0)) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (nb + j < N) dst[j] = v[j];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Weight folding at parameter-load time (fp64 accumulate):  C[m,n] = sum_k A[m,k] * B[k,n]
// and  c[m] = sum_k A[m,k] * b[k] + d[m].   Not on the rollout path.
__global__ void cn_fold_mm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Cm,
                                  int M, int N, int K) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N || m >= M) return;
  double acc = 0.0;
  for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * K + k] * (double)B[(size_t)k * N + n];
  Cm[(size_t)m * N + n] = (float)acc;
}
__global__ void cn_fold_mv_kernel(const float* __restrict__ A, const float* __restrict__ b, const float* __restrict__ d,
                                  float* __restrict__ c, int M, int K) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double acc = d ? (double)d[m] : 0.0;
  for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * K + k] * (double)b[k];
  c[m] = (float)acc;
}

// ------------------------------------------------------------------------------------------
// Row compaction.  Humans are sorted by distance and rows j >= n_e = detected_human_num[e] are
// padding: as attention KEYS they are masked (key_padding_mask), and as attention QUERIES their
// outputs only reach the robot-human softmax, where masked_fill(-1e9) gives them weight exactly 0
// (selfAttn_srnn_temp_node.py:49-60,165-170).  They cannot influence any output, so the per-human
// pipeline runs on the compacted valid rows only:  row_start[e] = sum_{e' < e} n_e',  *mc = total.
// Single CTA, 1024 threads, chunked inclusive scan (N <= a few 10^4).
__global__ void __launch_bounds__(1024) cn_row_offsets_kernel(const float* __restrict__ detected, int N, int H,
                                                              int* __restrict__ row_start, int* __restrict__ mc) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < N; base += 1024) {
    const int e = base + tid;
    int n = 0;
    if (e < N) { n = (int)detected[e]; n = n < 1 ? 1 : (n > H ? H : n); }
    int x = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const int incl = x + (warp > 0 ? warp_sums[warp - 1] : 0) + carry;
    if (e < N) row_start[e] = incl - n;
    __syncthreads();
    if (tid == 1023) carry = incl;
    __syncthreads();
  }
  if (tid == 0) { *mc = carry; row_start[N] = carry; }
}

// Input packing: x16[row_start[e] + j, 16] = spatial_edges[e, j] zero-padded to K=16 for j < n_e;
// xr[N,16] = cat(temporal_edges(2), robot_node(7)) zero padded; h0 = h_in * mask.
__global__ void cn_pack_inputs_kernel(const float* __restrict__ spatial, int Win, int H, int N,
                                      const int* __restrict__ row_start, int* __restrict__ row_env,
                                      float* __restrict__ x16,
                                      const float* __restrict__ temporal, const float* __restrict__ robot,
                                      const float* __restrict__ h_in, const float* __restrict__ masks,
                                      float* __restrict__ xr, float* __restrict__ h0, __half* __restrict__ h0_hi,
                                      __half* __restrict__ h0_lo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < N * H * 16) {
    const int r = idx >> 4, c = idx & 15;
    const int e = r / H, j = r - e * H;
    const int rs = row_start[e], n = row_start[e + 1] - rs;
    if (j < n) {
      x16[(size_t)(rs + j) * 16 + c] = c < Win ? spatial[(size_t)r * Win + c] : 0.0f;
      if (c == 0) row_env[rs + j] = e;
    }
  }
  if (idx < N * 16) {
    const int e = idx >> 4, c = idx & 15;
    float v = 0.0f;
    if (c < 2) v = temporal[2 * e + c];
    else if (c < 9) v = robot[7 * e + (c - 2)];
    xr[idx] = v;
  }
  if (idx < N * 128) {
    const float hv = h_in[idx] * masks[idx >> 7];
    h0[idx] = hv;
    if (h0_hi) {
      const float c = fminf(fmaxf(hv, -65504.0f), 65504.0f);
      const __half hh = __float2half_rn(c);
      h0_hi[idx] = hh;
      h0_lo[idx] = __float2half_rn(c - __half2float(hh));
    }
  }
}

// ------------------------------------------------------------------------------------------
// Human-human multi-head self attention for one (environment, head) per CTA.
// qkv: [N*H, 1536] rows = (q | k | v), head hd uses columns hd*64..hd*64+63 of each third.
// Keys j >= n_e are padding (key_padding_mask); query rows >= n_e are never consumed
// downstream (their robot-human attention weight is exactly 0), they are written as zeros.
__global__ void __launch_bounds__(256) cn_hh_attention_kernel(const float* __restrict__ qkv,
                                                              const int* __restrict__ row_start,
                                                              const int* __restrict__ row_env,
                                                              const int* __restrict__ mc_ptr,
                                                              float* __restrict__ out /* [Mc,512] or null */,
                                                              __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
  // One WARP per valid (compacted) human row = one attention query, all 8 heads at once.
  // Every 512-float q / k / v / o row is touched with four fully coalesced LDG.128 per warp: lane l
  // owns float4 #(l + 32 k), k = 0..3, i.e. elements 128 k + 4 l .. + 3, which belong to head
  // 2 k + (l >= 16).  A key's score for head (2k + half) is the 4-FMA partial reduced over the 16 lanes
  // of the half (4 xor-shuffles); soft-max runs online (running max / sum per head), nothing is staged.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mc = *mc_ptr;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const float scale = 0.125f;   // 1/sqrt(head_dim = 64); torch scales q before q k^T
  // grid-stride over the compacted rows: the launch is sized to the machine, not to the worst case
  for (int r = blockIdx.x * (blockDim.x >> 5) + warp; r < mc; r += nwarps) {
    const int e = row_env[r];
    const size_t row0 = (size_t)row_start[e];
    const int n = row_start[e + 1] - row_start[e];
    float4 q[4], acc[4];
    float m[4], l[4];
    {
      const float4* qv = reinterpret_cast<const float4*>(qkv + (size_t)r * 1536) + lane;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 a = __ldg(qv + 32 * k);
        q[k] = make_float4(a.x * scale, a.y * scale, a.z * scale, a.w * scale);
        acc[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        m[k] = -INFINITY; l[k] = 0.0f;
      }
    }
    for (int j = 0; j < n; ++j) {
      const float4* kv = reinterpret_cast<const float4*>(qkv + (row0 + j) * 1536 + 512) + lane;
      const float4* vv = reinterpret_cast<const float4*>(qkv + (row0 + j) * 1536 + 1024) + lane;
      float4 kr[4], vr[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { kr[k] = __ldg(kv + 32 * k); vr[k] = __ldg(vv + 32 * k); }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float s = q[k].x * kr[k].x;
        s = fmaf(q[k].y, kr[k].y, s); s = fmaf(q[k].z, kr[k].z, s); s = fmaf(q[k].w, kr[k].w, s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);        // all 16 lanes of the half hold the head's score
        const float mn = fmaxf(m[k], s);
        const float corr = expf(m[k] - mn);              // exp(-inf) = 0 on the first key
        const float pj = expf(s - mn);
        l[k] = l[k] * corr + pj;
        acc[k].x = fmaf(pj, vr[k].x, acc[k].x * corr); acc[k].y = fmaf(pj, vr[k].y, acc[k].y * corr);
        acc[k].z = fmaf(pj, vr[k].z, acc[k].z * corr); acc[k].w = fmaf(pj, vr[k].w, acc[k].w * corr);
        m[k] = mn;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float inv = 1.0f / l[k];
      acc[k].x *= inv; acc[k].y *= inv; acc[k].z *= inv; acc[k].w *= inv;
    }
    if (out) {
      float4* dst = reinterpret_cast<float4*>(out + (size_t)r * 512) + lane;
#pragma unroll
      for (int k = 0; k < 4; ++k) dst[32 * k] = acc[k];
    }
    if (out_hi) {     // (hi, lo) fp16 split = A operand of the tensor-core out-projection
      uint2* dh = reinterpret_cast<uint2*>(out_hi + (size_t)r * 512) + lane;
      uint2* dl = reinterpret_cast<uint2*>(out_lo + (size_t)r * 512) + lane;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float c[4] = {fminf(fmaxf(acc[k].x, -65504.0f), 65504.0f), fminf(fmaxf(acc[k].y, -65504.0f), 65504.0f),
                            fminf(fmaxf(acc[k].z, -65504.0f), 65504.0f), fminf(fmaxf(acc[k].w, -65504.0f), 65504.0f)};
        uint32_t ph[2], pl[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const __half h0 = __float2half_rn(c[2 * t]), h1 = __float2half_rn(c[2 * t + 1]);
          const __half l0 = __float2half_rn(c[2 * t] - __half2float(h0)), l1 = __float2half_rn(c[2 * t + 1] - __half2float(h1));
          ph[t] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          pl[t] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        }
        dh[32 * k] = make_uint2(ph[0], ph[1]);
        dl[32 * k] = make_uint2(pl[0], pl[1]);
      }
    }
  }   // row loop
}

// ------------------------------------------------------------------------------------------
// Robot-human attention (EdgeAttention_M.att_func) for one environment per warp:
//   score_j = <W_t robot + b_t, W_s s_j + b_s> * (H / sqrt(64))  ==  (u . s_j + cst) * H/8
// with u = W_s^T te (precomputed by a GEMM), cst = <b_s, te>;  masked_fill(-1e9) for j >= n_e;
// softmax over all H; weighted sum of the 256-d human features.
__global__ void __launch_bounds__(128) cn_hr_attention_kernel(const float* __restrict__ s_out /* [N*H,256] */,
                                                              const float* __restrict__ u /* [N,256] */,
                                                              const float* __restrict__ te /* [N, ldte] cols te_off.. */,
                                                              int ldte, int te_off, const float* __restrict__ b_s,
                                                              const int* __restrict__ row_start, int N, int H,
                                                              float* __restrict__ wv /* [N,256] */,
                                                              __half* __restrict__ wv_hi, __half* __restrict__ wv_lo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * 4 + warp;
  if (e >= N) return;
  const size_t row0 = (size_t)row_start[e];
  const int n = row_start[e + 1] - row_start[e];
  float ur[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) ur[t] = u[(size_t)e * 256 + lane + 32 * t];
  float cst = b_s[lane] * te[(size_t)e * ldte + te_off + lane] + b_s[lane + 32] * te[(size_t)e * ldte + te_off + lane + 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cst += __shfl_xor_sync(0xffffffffu, cst, o);
  const float temperature = (float)H / 8.0f;
  // scores for valid humans; lanes cooperate on each 256-d dot product
  float sc[4] = {-1e9f, -1e9f, -1e9f, -1e9f};   // lane holds score of human lane + 32 t
  for (int j = 0; j < n; ++j) {
    const float* sr = s_out + (row0 + j) * 256;
    float d = 0.0f;
#pragma unroll
    for (int t = 0; t < 8; ++t) d = fmaf(ur[t], sr[lane + 32 * t], d);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    const float s = (d + cst) * temperature;
    if ((j & 31) == lane) sc[j >> 5] = s;
  }
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < 4; ++t) if (lane + 32 * t < H) mx = fmaxf(mx, sc[t]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.0f;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    sc[t] = (lane + 32 * t < H) ? expf(sc[t] - mx) : 0.0f;   // masked entries: exp(-1e9 - mx) == 0
    sum += sc[t];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < n; ++j) {
    const float pj = __shfl_sync(0xffffffffu, sc[j >> 5], j & 31) * inv;
    const float* sr = s_out + (row0 + j) * 256;
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = fmaf(pj, sr[lane + 32 * t], acc[t]);
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const size_t o = (size_t)e * 256 + lane + 32 * t;
    wv[o] = acc[t];
    if (wv_hi) {
      const float c = fminf(fmaxf(acc[t], -65504.0f), 65504.0f);
      const __half hh = __float2half_rn(c);
      wv_hi[o] = hh;
      wv_lo[o] = __float2half_rn(c - __half2float(hh));
    }
  }
}

// ------------------------------------------------------------------------------------------
// GRU cell gates (PyTorch order r, z, n; h' = (1 - z) * n + z * h), one thread per (env, unit).
__device__ __forceinline__ float cn_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void cn_gru_gate_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                   const float* __restrict__ h0, int N, float* __restrict__ h1,
                                   __half* __restrict__ h1_hi, __half* __restrict__ h1_lo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * 128) return;
  const int e = idx >> 7, c = idx & 127;
  const float* a = gi + (size_t)e * 384;
  const float* b = gh + (size_t)e * 384;
  const float r = cn_sigmoid(a[c] + b[c]);
  const float z = cn_sigmoid(a[128 + c] + b[128 + c]);
  const float n = tanhf(a[256 + c] + r * b[256 + c]);
  const float hv = (1.0f - z) * n + z * h0[idx];
  h1[idx] = hv;
  if (h1_hi) {       // (hi, lo) fp16 split for the tensor-core output_linear
    const __half hh = __float2half_rn(hv);
    h1_hi[idx] = hh;
    h1_lo[idx] = __float2half_rn(hv - __half2float(hh));
  }
}

// ------------------------------------------------------------------------------------------
// Output heads, one warp per environment: value = critic_linear(hc); mean = fc_mean(ha);
// action = mean + exp(logstd) * noise (torch.normal = randn * std + mean); log-prob summed.
__global__ void __launch_bounds__(128) cn_heads_kernel(const float* __restrict__ ha, int ldha,
                                                       const float* __restrict__ hc, int ldhc,
                                                       const float* __restrict__ w_v, const float* __restrict__ b_v,
                                                       const float* __restrict__ w_m, const float* __restrict__ b_m,
                                                       const float* __restrict__ logstd,
                                                       const float* __restrict__ noise, int N,
                                                       float* __restrict__ value, float* __restrict__ action,
                                                       float* __restrict__ logp, float* __restrict__ mean_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * 4 + warp;
  if (e >= N) return;
  float v = 0.0f, m0 = 0.0f, m1 = 0.0f;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int c = lane + 32 * t;
    const float a = ha[(size_t)e * ldha + c], cc = hc[(size_t)e * ldhc + c];
    v = fmaf(cc, w_v[c], v);
    m0 = fmaf(a, w_m[c], m0);
    m1 = fmaf(a, w_m[256 + c], m1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v += __shfl_xor_sync(0xffffffffu, v, o);
    m0 += __shfl_xor_sync(0xffffffffu, m0, o);
    m1 += __shfl_xor_sync(0xffffffffu, m1, o);
  }
  if (lane == 0) {
    v += b_v[0]; m0 += b_m[0]; m1 += b_m[1];
    value[e] = v;
    if (mean_out) { mean_out[2 * e] = m0; mean_out[2 * e + 1] = m1; }
    const float ls0 = logstd[0], ls1 = logstd[1];
    const float s0 = expf(ls0), s1 = expf(ls1);
    float a0 = m0, a1 = m1;
    if (noise) {
      a0 = __fadd_rn(__fmul_rn(noise[2 * e], s0), m0);
      a1 = __fadd_rn(__fmul_rn(noise[2 * e + 1], s1), m1);
    }
    action[2 * e] = a0; action[2 * e + 1] = a1;
    // Normal.log_prob: -((x - mu)^2) / (2 var) - log(std) - log(sqrt(2 pi))
    const float c = 0.91893853320467274178f;
    const float d0 = a0 - m0, d1 = a1 - m1;
    const float l0 = -(d0 * d0) / (2.0f * (s0 * s0)) - logf(s0) - c;   // log_scale = scale.log()
    const float l1 = -(d1 * d1) / (2.0f * (s1 * s1)) - logf(s1) - c;
    logp[e] = l0 + l1;
  }
}
