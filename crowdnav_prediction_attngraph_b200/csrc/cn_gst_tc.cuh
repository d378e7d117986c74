// Config 3's GST predictor with its dense layers on the tcgen05 3xFP16 GEMM (included at the end of cn_policy.cu so it
// shares that translation unit's TMA / GEMM host helpers).  Same arithmetic as the fused CUDA-core kernel of
// cn_gst.cu (reference lines are cited there); here every layer is a batched [rows, K] GEMM over ALL environments:
//   observation period: rows = N * 5 * H (frames of one environment contiguous: row = (e * 5 + t) * H + n)
//   decoding steps:     rows = N * H
// and small row-wise kernels do embedding + LayerNorm, the H x H attention, residuals, the LSTM cell and the wrapper's
// tail.  Default path of cn_gst_step (CN_GST_MODE=fused selects the single-kernel version); validated by the same
// tests as the fused kernel.
#pragma once

namespace {

#define GT_T 5
#define GT_INVALID (-999.0f)

struct GstTcW {   // fp32 device parameters used by the row-wise kernels
  const float *We_t, *be, *ln0_g, *ln0_b, *ln1_g, *ln1_b, *Wp, *bp;
  const float *bin, *bout, *b1, *b2, *bih, *bhh;
};

__device__ __forceinline__ void gt_split_store(__half* hi, __half* lo, size_t idx, float x) {
  const float c = fminf(fmaxf(x, -65504.0f), 65504.0f);
  const __half h = __float2half_rn(c);
  hi[idx] = h;
  lo[idx] = __float2half_rn(c - __half2float(h));
}

// LayerNorm of one 64-wide row held as two values per lane
__device__ __forceinline__ void gt_ln(float a0, float a1, const float* g, const float* b, int lane, float& o0, float& o1) {
  float s = a0 + a1;
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / 64.0f);
  const float d0 = a0 - mean, d1 = a1 - mean;
  float v = d0 * d0 + d1 * d1;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float inv = rsqrtf(v * (1.0f / 64.0f) + 1e-5f);
  o0 = d0 * inv * g[lane] + b[lane];
  o1 = d1 * inv * g[lane + 32] + b[lane + 32];
}

// ring append + input processing of the 5 observed frames + node embedding + norm_node + mask.  One warp per row.
__global__ void __launch_bounds__(256) gt_prep_kernel(GstTcW w, int N, int H, float* __restrict__ ring_pos,
                                                      uint8_t* __restrict__ ring_mask, int newest,
                                                      const float* __restrict__ robot, const float* __restrict__ sp2,
                                                      const uint8_t* __restrict__ vis, float* __restrict__ X0,
                                                      __half* __restrict__ xh, __half* __restrict__ xl, float* __restrict__ rowm,
                                                      float* __restrict__ fp, float* __restrict__ pos_last,
                                                      float* __restrict__ h32, __half* __restrict__ hh, __half* __restrict__ hl,
                                                      float* __restrict__ c32, float* __restrict__ mu_cum) {
  cn_pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int R = N * GT_T * H;
  if (r >= R) return;
  const int n = r % H, t = (r / H) % GT_T, e = r / (H * GT_T);
  // positions / masks of frame t, t-1 and the newest frame; the newest comes straight from this step's observation
  auto frame_pos = [&](int tt, float& x, float& y, float& m) {
    if (tt == GT_T - 1) {
      x = robot[e * 7] + sp2[((size_t)e * H + n) * 2];
      y = robot[e * 7 + 1] + sp2[((size_t)e * H + n) * 2 + 1];
      m = vis[(size_t)e * H + n] ? 1.0f : 0.0f;
    } else {
      const int slot = (newest + 1 + tt) % GT_T;
      const size_t o = ((size_t)slot * N + e) * H + n;
      x = ring_pos[2 * o]; y = ring_pos[2 * o + 1]; m = (float)ring_mask[o];
    }
  };
  float x, y, m, xp = 0, yp = 0, mp = 0, xl_, yl_, ml_;
  frame_pos(t, x, y, m);
  frame_pos(GT_T - 1, xl_, yl_, ml_);
  if (t > 0) frame_pos(t - 1, xp, yp, mp);
  const float mrel = t == 0 ? m : mp * ml_;                  // interface.forward:77-78 (sic)
  const float dx = t == 0 ? 0.0f : x - xp, dy = t == 0 ? 0.0f : y - yp;
  const float ix = GT_INVALID * (1.0f - mrel) + dx * mrel, iy = GT_INVALID * (1.0f - mrel) + dy * mrel;
  float e0 = fmaf(iy, w.We_t[64 + lane], fmaf(ix, w.We_t[lane], w.be[lane]));
  float e1 = fmaf(iy, w.We_t[96 + lane], fmaf(ix, w.We_t[32 + lane], w.be[32 + lane]));
  float o0, o1;
  gt_ln(e0, e1, w.ln0_g, w.ln0_b, lane, o0, o1);
  o0 *= mrel; o1 *= mrel;
  const size_t b = (size_t)r * 64;
  X0[b + lane] = o0; X0[b + lane + 32] = o1;
  gt_split_store(xh, xl, b + lane, o0); gt_split_store(xh, xl, b + lane + 32, o1);
  if (lane == 0) rowm[r] = mrel;
  if (t == GT_T - 1) {
    const size_t rd = (size_t)e * H + n;
    if (lane == 0) {
      fp[rd] = mrel; pos_last[2 * rd] = x; pos_last[2 * rd + 1] = y; mu_cum[2 * rd] = 0.0f; mu_cum[2 * rd + 1] = 0.0f;
      // traj_buffer.append / mask_buffer.append (readers of this slot in this launch use the observation directly)
      const size_t o = ((size_t)newest * N + e) * H + n;
      ring_pos[2 * o] = x; ring_pos[2 * o + 1] = y; ring_mask[o] = m != 0.0f ? 1 : 0;
    }
    h32[rd * 64 + lane] = 0.0f; h32[rd * 64 + lane + 32] = 0.0f;
    c32[rd * 64 + lane] = 0.0f; c32[rd * 64 + lane + 32] = 0.0f;
    hh[rd * 64 + lane] = __float2half_rn(0.0f); hh[rd * 64 + lane + 32] = __float2half_rn(0.0f);
    hl[rd * 64 + lane] = __float2half_rn(0.0f); hl[rd * 64 + lane + 32] = __float2half_rn(0.0f);
  }
}

// decoding step: node embedding of x_sample + norm_node + mask.  One warp per row of [N*H].
__global__ void __launch_bounds__(256) gt_embed_kernel(GstTcW w, int Rd, const float* __restrict__ xin, const float* __restrict__ fp,
                                                       float* __restrict__ X0, __half* __restrict__ xh, __half* __restrict__ xl) {
  cn_pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= Rd) return;
  const float ix = xin[2 * r], iy = xin[2 * r + 1];
  float e0 = fmaf(iy, w.We_t[64 + lane], fmaf(ix, w.We_t[lane], w.be[lane]));
  float e1 = fmaf(iy, w.We_t[96 + lane], fmaf(ix, w.We_t[32 + lane], w.be[32 + lane]));
  float o0, o1;
  gt_ln(e0, e1, w.ln0_g, w.ln0_b, lane, o0, o1);
  const float m = fp[r];
  o0 *= m; o1 *= m;
  const size_t b = (size_t)r * 64;
  X0[b + lane] = o0; X0[b + lane + 32] = o1;
  gt_split_store(xh, xl, b + lane, o0); gt_split_store(xh, xl, b + lane + 32, o1);
}

// attention within groups of H consecutive rows: one CTA per group (8 H threads = one per (row, head)), the group's
// q | k | v rows staged once in shared memory (every key / value row is read by all 8 H threads of the group: served
// from L1 this kernel was 1/3 of the predictor's time).  rowm: per-row validity.
#define GT_MAXH 32
__global__ void __launch_bounds__(8 * GT_MAXH) gt_attn_kernel(int R, int H, const float* __restrict__ qkv,
                                                              const float* __restrict__ rowm, __half* __restrict__ ah,
                                                              __half* __restrict__ al) {
  cn_pdl_prologue();
  __shared__ __align__(16) float sq[GT_MAXH * 192];
  __shared__ float sm_[GT_MAXH];
  const int g0 = blockIdx.x * H;
  if (g0 >= R) return;
  for (int i = threadIdx.x; i < H * 48; i += blockDim.x)
    reinterpret_cast<float4*>(sq)[i] = __ldg(reinterpret_cast<const float4*>(qkv + (size_t)g0 * 192) + i);
  for (int i = threadIdx.x; i < H; i += blockDim.x) sm_[i] = rowm[g0 + i];
  __syncthreads();
  const int lr = threadIdx.x >> 3, hd = threadIdx.x & 7;
  if (lr >= H) return;
  const float scaling = 0.35355339059327373f;
  float q[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) q[d] = sq[lr * 192 + hd * 8 + d] * scaling;
  float mx = -INFINITY;
  for (int j = 0; j < H; ++j) {
    const float* kj = sq + j * 192 + 64 + hd * 8;
    float s = 0.0f;
#pragma unroll
    for (int d = 0; d < 8; ++d) s = fmaf(q[d], kj[d], s);
    mx = fmaxf(mx, s);
  }
  float den = 0.0f, dm = 0.0f, o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const float mi = sm_[lr];
  for (int j = 0; j < H; ++j) {
    const float* kj = sq + j * 192 + 64 + hd * 8;
    const float* vj = sq + j * 192 + 128 + hd * 8;
    float s = 0.0f;
#pragma unroll
    for (int d = 0; d < 8; ++d) s = fmaf(q[d], kj[d], s);
    const float ex = expf(s - mx);
    den += ex;
    const float em = ex * (mi * sm_[j]);
    dm += em;
#pragma unroll
    for (int d = 0; d < 8; ++d) o[d] = fmaf(em, vj[d], o[d]);
  }
  const float scale = (1.0f / den) / (dm / den + 1e-10f);
  const size_t ob = (size_t)(g0 + lr) * 64 + hd * 8;
#pragma unroll
  for (int d = 0; d < 8; ++d) gt_split_store(ah, al, ob + d, o[d] * scale);
}

// X1 = X0 + O (fp32), Y = norm1(X1) as fp16 hi/lo.  One warp per row.
__global__ void __launch_bounds__(256) gt_res_ln_kernel(GstTcW w, int R, const float* __restrict__ X0, const float* __restrict__ O,
                                                        float* __restrict__ X1, __half* __restrict__ yh, __half* __restrict__ yl) {
  cn_pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= R) return;
  const size_t b = (size_t)r * 64;
  const float a0 = X0[b + lane] + O[b + lane], a1 = X0[b + lane + 32] + O[b + lane + 32];
  X1[b + lane] = a0; X1[b + lane + 32] = a1;
  float o0, o1;
  gt_ln(a0, a1, w.ln1_g, w.ln1_b, lane, o0, o1);
  gt_split_store(yh, yl, b + lane, o0); gt_split_store(yh, yl, b + lane + 32, o1);
}

// XS = (X1 + O2) * rowmask as fp16 hi/lo (input of W_ih)
__global__ void __launch_bounds__(256) gt_res_mask_kernel(size_t count, const float* __restrict__ X1, const float* __restrict__ O2,
                                                          const float* __restrict__ rowm, __half* __restrict__ sh, __half* __restrict__ sl) {
  cn_pdl_prologue();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  gt_split_store(sh, sl, i, (X1[i] + O2[i]) * rowm[i >> 6]);
}

__device__ __forceinline__ float gt_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// LSTM cell.  gx row of (e, n) = gx_row0 + (e * gx_env_stride + n); masked update when fp != null (decoding).
__global__ void __launch_bounds__(256) gt_cell_kernel(int N, int H, const float* __restrict__ GX, int gx_env_stride, int gx_row0,
                                                      const float* __restrict__ GH, const float* __restrict__ fp,
                                                      float* __restrict__ h32, float* __restrict__ c32, __half* __restrict__ hh,
                                                      __half* __restrict__ hl) {
  cn_pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * H * 64) return;
  const int j = i & 63, rd = i >> 6, e = rd / H, n = rd - e * H;
  const float* gx = GX + ((size_t)gx_row0 + (size_t)e * gx_env_stride + n) * 256;
  const float* gh = GH + (size_t)rd * 256;
  const float ig = gt_sigmoid(gx[j] + gh[j]), fg = gt_sigmoid(gx[64 + j] + gh[64 + j]);
  const float gg = tanhf(gx[128 + j] + gh[128 + j]), og = gt_sigmoid(gx[192 + j] + gh[192 + j]);
  float c2 = fg * c32[i] + ig * gg, h2 = og * tanhf(c2);
  if (fp) {
    const float m = fp[rd];
    c2 = c2 * m + c32[i] * (1.0f - m);
    h2 = h2 * m + h32[i] * (1.0f - m);
  }
  c32[i] = c2; h32[i] = h2;
  gt_split_store(hh, hl, i, h2);
}

// after the observation period: h, c *= fp
__global__ void __launch_bounds__(256) gt_mask_state_kernel(int count, const float* __restrict__ fp, float* __restrict__ h32,
                                                            float* __restrict__ c32, __half* __restrict__ hh, __half* __restrict__ hl) {
  cn_pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float m = fp[i >> 6];
  const float h = h32[i] * m;
  h32[i] = h; c32[i] *= m;
  gt_split_store(hh, hl, i, h);
}

// hidden2pos (mean only) -> x_sample, cumulative mean, predicted world position of step tt
__global__ void __launch_bounds__(256) gt_h2p_kernel(GstTcW w, int Rd, int tt, const float* __restrict__ h32, const float* __restrict__ fp,
                                                     const float* __restrict__ pos_last, float* __restrict__ xin,
                                                     float* __restrict__ mu_cum, float* __restrict__ pred) {
  cn_pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Rd * 2) return;
  const int n = i >> 1, d = i & 1;
  float a = w.bp[d];
  for (int k = 0; k < 64; ++k) a = fmaf(h32[(size_t)n * 64 + k], w.Wp[d * 64 + k], a);
  const float m = fp[n];
  xin[i] = a * m;
  const float cum = mu_cum[i] + a;
  mu_cum[i] = cum;
  pred[((size_t)n * GT_T + tt) * 2 + d] = (cum + pos_last[i]) * m + GT_INVALID * (1.0f - m);
}

// process_obs_rew tail: one CTA (32 threads) per environment
__global__ void __launch_bounds__(32) gt_final_kernel(int N, int H, int P, float thr, float collision_penalty,
                                                      const float* __restrict__ robot, const float* __restrict__ sp2,
                                                      const float* __restrict__ fp, const float* __restrict__ pred,
                                                      float* __restrict__ reward, float* __restrict__ penalty_out,
                                                      float* __restrict__ out_sp) {
  cn_pdl_prologue();
  const int e = blockIdx.x;
  const float rx = robot[e * 7], ry = robot[e * 7 + 1];
  float pen = 0.0f;
  for (int i = threadIdx.x; i < H * GT_T; i += 32) {
    const int n = i / GT_T, k = i - n * GT_T;
    if (k < P && fp[(size_t)e * H + n] != 0.0f) {
      const float dx = pred[((size_t)e * H * GT_T + i) * 2] - rx, dy = pred[((size_t)e * H * GT_T + i) * 2 + 1] - ry;
      if (sqrtf(dx * dx + dy * dy) < thr) pen = fminf(pen, collision_penalty / (float)(4 << k));
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) pen = fminf(pen, __shfl_xor_sync(0xffffffffu, pen, o));
  if (threadIdx.x == 0) {
    if (reward) reward[e] += pen;
    if (penalty_out) penalty_out[e] = pen;
  }
  const int W = 2 * (P + 1);
  for (int n = threadIdx.x; n < H; n += 32) {
    const float cx = sp2[((size_t)e * H + n) * 2], cy = sp2[((size_t)e * H + n) * 2 + 1];
    const float key = sqrtf(cx * cx + cy * cy);
    int rank = 0;
    for (int j = 0; j < H; ++j) {
      const float ox = sp2[((size_t)e * H + j) * 2], oy = sp2[((size_t)e * H + j) * 2 + 1];
      const float kj = sqrtf(ox * ox + oy * oy);
      rank += (kj < key || (kj == key && j < n)) ? 1 : 0;
    }
    float* dst = out_sp + ((size_t)e * H + rank) * W;
    dst[0] = cx; dst[1] = cy;
    const bool ok = fp[(size_t)e * H + n] != 0.0f;
    for (int k = 0; k < P; ++k) {
      dst[2 + 2 * k] = ok ? pred[(((size_t)e * H + n) * GT_T + k) * 2] - rx : cx;
      dst[3 + 2 * k] = ok ? pred[(((size_t)e * H + n) * GT_T + k) * 2 + 1] - ry : cy;
    }
  }
}

struct GstTc {
  cn_policy* ctx;
  int N, H, P;
  float thr, pen;
  GstTcW w;
  TcMat tWin, tWout, tW1, tW2, tWih, tWhh;                   // weights (x 2^6, fp16 hi/lo)
  TcMat tX, tA, tY, tF, tXS, tHd;                            // activations (fp16 hi/lo A operands)
  float *X0, *QKV, *O, *X1, *GX, *GH, *rowm, *fp, *pos_last, *h32, *c32, *mu_cum, *xin, *pred;
};

int gt_upload(cn_policy* ctx, const float** dst, const float* src, size_t count) {
  float* q = nullptr;
  int rc = palloc(ctx, &q, count);
  if (rc) return rc;
  if (cudaMemcpy(q, src, count * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return cn_set_error("gst tc: H2D failed");
  *dst = q;
  return 0;
}

}  // namespace

// C++ entry points used by cn_gst.cu (not part of the public C ABI)
void* cn_gst_tc_create(int N, int H, int P, float thr, float pen, int device, const float* const* host /* 20 params, cn_gst.cu order, */,
                       const int* rows, const int* cols) {
  cudaSetDevice(device);
  GstTc* g = new GstTc();
  cn_policy* ctx = new cn_policy();
  ctx->cfg.device = device; ctx->launches = 0; ctx->launch_error = false; ctx->qkv_chunks = 1; ctx->finalized = true;
  ctx->st2 = nullptr; ctx->st3 = nullptr; ctx->profile = false;
  ctx->num_sms = 148;
  cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
  { const char* pd = getenv("CN_PDL"); ctx->pdl = !(pd && pd[0] == '0'); }
  g->ctx = ctx; g->N = N; g->H = H; g->P = P; g->thr = thr; g->pen = pen;
  int rc = tc_set_attrs();
  // indices in cn_gst.cu's kParamNames: 0 We 1 be 2 ln0g 3 ln0b 4 Win 5 bin 6 Wout 7 bout 8 ln1g 9 ln1b 10 W1 11 b1 12 W2 13 b2
  //                                     14 Wih 15 bih 16 Whh 17 bhh 18 Wp 19 bp
  std::vector<float> wet(128);
  for (int c = 0; c < 64; ++c) { wet[c] = host[0][c * 2]; wet[64 + c] = host[0][c * 2 + 1]; }     // [64][2] -> [2][64]
  if (!rc) rc = gt_upload(ctx, &g->w.We_t, wet.data(), 128);
  const float** fdst[] = {&g->w.be, &g->w.ln0_g, &g->w.ln0_b, &g->w.bin, &g->w.bout, &g->w.ln1_g, &g->w.ln1_b, &g->w.b1, &g->w.b2,
                          &g->w.bih, &g->w.bhh, &g->w.Wp, &g->w.bp};
  const int fidx[] = {1, 2, 3, 5, 7, 8, 9, 11, 13, 15, 17, 18, 19};
  for (int i = 0; i < 13 && !rc; ++i) rc = gt_upload(ctx, fdst[i], host[fidx[i]], (size_t)rows[fidx[i]] * cols[fidx[i]]);
  struct { int idx; TcMat* t; } tw[6] = {{4, &g->tWin}, {6, &g->tWout}, {10, &g->tW1}, {12, &g->tW2}, {14, &g->tWih}, {16, &g->tWhh}};
  for (int i = 0; i < 6 && !rc; ++i) {
    const int r = rows[tw[i].idx], k = cols[tw[i].idx];
    const float* d = nullptr;
    rc = gt_upload(ctx, &d, host[tw[i].idx], (size_t)r * k);
    if (!rc) rc = tc_alloc(ctx, *tw[i].t, r, k, r == 256 ? 256 : 64);        // the 256-wide gate GEMMs use BN = 256 tiles
    if (!rc) split16(ctx, 0, d, 64.0f, tw[i].t->hi, tw[i].t->lo, (size_t)r * k);
  }
  const size_t R = (size_t)N * GT_T * H, Rd = (size_t)N * H;
  if (!rc) rc = tc_alloc(ctx, g->tX, (int)R, 64, TC_BM);
  if (!rc) rc = tc_alloc(ctx, g->tA, (int)R, 64, TC_BM);
  if (!rc) rc = tc_alloc(ctx, g->tY, (int)R, 64, TC_BM);
  if (!rc) rc = tc_alloc(ctx, g->tF, (int)R, 128, TC_BM);
  if (!rc) rc = tc_alloc(ctx, g->tXS, (int)R, 64, TC_BM);
  if (!rc) rc = tc_alloc(ctx, g->tHd, (int)Rd, 64, TC_BM);
#define GA(name, count) if (!rc) rc = palloc(ctx, &g->name, (count))
  GA(X0, R * 64); GA(QKV, R * 192); GA(O, R * 64); GA(X1, R * 64); GA(GX, R * 256); GA(GH, Rd * 256); GA(rowm, R); GA(fp, Rd);
  GA(pos_last, Rd * 2); GA(h32, Rd * 64); GA(c32, Rd * 64); GA(mu_cum, Rd * 2); GA(xin, Rd * 2); GA(pred, Rd * GT_T * 2);
#undef GA
  if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = cn_set_error("gst tc: setup failed");
  if (rc) { return nullptr; }
  return g;
}

void cn_gst_tc_destroy(void* handle) {
  GstTc* g = static_cast<GstTc*>(handle);
  if (!g) return;
  for (void* q : g->ctx->allocs) cudaFree(q);
  for (auto* m : g->ctx->omaps) delete m;
  delete g->ctx;
  delete g;
}

int64_t cn_gst_tc_launches(void* handle) { return handle ? static_cast<GstTc*>(handle)->ctx->launches : 0; }

int cn_gst_tc_step(void* handle, float* ring_pos, uint8_t* ring_mask, int newest, const float* robot, const float* sp2,
                   const uint8_t* vis, float* reward, float* penalty, float* out_sp, cudaStream_t st) {
  GstTc* g = static_cast<GstTc*>(handle);
  cn_policy* p = g->ctx;
  const int N = g->N, H = g->H;
  const int R = N * GT_T * H, Rd = N * H;
  auto warps = [](int rows) { return dim3((unsigned)((rows + 7) / 8)); };          // 8 warps (rows) per 256-thread CTA
  auto encoder = [&](int rows, const float* rowm) {
    gemm_tc(p, st, g->tX, g->tWin, rows, 192, 64, 64, g->w.bin, CN_ACT_NONE, out32(g->QKV, 192));
    launch_k(p, gt_attn_kernel, dim3((unsigned)(rows / H)), dim3((unsigned)(8 * H)), 0, st, rows, H, g->QKV, rowm, g->tA.hi, g->tA.lo);
    gemm_tc(p, st, g->tA, g->tWout, rows, 64, 64, 64, g->w.bout, CN_ACT_NONE, out32(g->O, 64));
    launch_k(p, gt_res_ln_kernel, warps(rows), dim3(256), 0, st, g->w, rows, g->X0, g->O, g->X1, g->tY.hi, g->tY.lo);
    gemm_tc(p, st, g->tY, g->tW1, rows, 128, 64, 64, g->w.b1, CN_ACT_RELU, out16(g->tF));
    gemm_tc(p, st, g->tF, g->tW2, rows, 64, 128, 64, g->w.b2, CN_ACT_NONE, out32(g->O, 64));
    launch_k(p, gt_res_mask_kernel, dim3((unsigned)(((size_t)rows * 64 + 255) / 256)), dim3(256), 0, st, (size_t)rows * 64, g->X1, g->O,
             rowm, g->tXS.hi, g->tXS.lo);
    gemm_tc(p, st, g->tXS, g->tWih, rows, 256, 64, 256, g->w.bih, CN_ACT_NONE, out32(g->GX, 256));
  };
  launch_k(p, gt_prep_kernel, warps(R), dim3(256), 0, st, g->w, N, H, ring_pos, ring_mask, newest, robot, sp2, vis, g->X0, g->tX.hi,
           g->tX.lo, g->rowm, g->fp, g->pos_last, g->h32, g->tHd.hi, g->tHd.lo, g->c32, g->mu_cum);
  encoder(R, g->rowm);
  const unsigned cell_grid = (unsigned)((Rd * 64 + 255) / 256);
  for (int t = 0; t < GT_T; ++t) {
    gemm_tc(p, st, g->tHd, g->tWhh, Rd, 256, 64, 256, g->w.bhh, CN_ACT_NONE, out32(g->GH, 256));
    launch_k(p, gt_cell_kernel, dim3(cell_grid), dim3(256), 0, st, N, H, g->GX, GT_T * H, t * H, g->GH, (const float*)nullptr, g->h32,
             g->c32, g->tHd.hi, g->tHd.lo);
  }
  launch_k(p, gt_mask_state_kernel, dim3(cell_grid), dim3(256), 0, st, Rd * 64, g->fp, g->h32, g->c32, g->tHd.hi, g->tHd.lo);
  for (int tt = 0; tt < GT_T; ++tt) {
    if (tt > 0) {
      launch_k(p, gt_embed_kernel, warps(Rd), dim3(256), 0, st, g->w, Rd, g->xin, g->fp, g->X0, g->tX.hi, g->tX.lo);
      encoder(Rd, g->fp);
      gemm_tc(p, st, g->tHd, g->tWhh, Rd, 256, 64, 256, g->w.bhh, CN_ACT_NONE, out32(g->GH, 256));
      launch_k(p, gt_cell_kernel, dim3(cell_grid), dim3(256), 0, st, N, H, g->GX, H, 0, g->GH, g->fp, g->h32, g->c32, g->tHd.hi,
               g->tHd.lo);
    }
    launch_k(p, gt_h2p_kernel, dim3((unsigned)((Rd * 2 + 255) / 256)), dim3(256), 0, st, g->w, Rd, tt, g->h32, g->fp, g->pos_last, g->xin,
             g->mu_cum, g->pred);
  }
  launch_k(p, gt_final_kernel, dim3((unsigned)N), dim3(32), 0, st, N, H, g->P, g->thr, g->pen, robot, sp2, g->fp, g->pred, reward, penalty,
           out_sp);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return cn_set_error("gst tc step: %s", cudaGetErrorString(err));
  if (p->launch_error) { p->launch_error = false; return 1; }
  return 0;
}
