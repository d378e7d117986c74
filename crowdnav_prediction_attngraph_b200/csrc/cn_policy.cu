// Host side of the attention-graph policy forward: parameter upload / folding and the
// per-rollout-step launch sequence behind cn_policy_act (replaces Policy.act,
// rl/networks/model.py:56-74, for base = selfAttn_merge_SRNN with sort_humans = True).
//
// Algebraic folds done once per parameter upload (fp64 accumulate, exact same function):
//   q/k/v_linear  o  MultiheadAttention.in_proj   ->  one 512 -> 1536 projection
//   MultiheadAttention.out_proj  o  spatial_linear -> one 512 -> 256 projection (ReLU after)
//   spatial_edge_layer folded into the robot side of the dot-product attention (u = W_s^T te)
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/crowdnav_b200.h"
#include "cn_host_util.h"
#include "cn_policy_kernels.cuh"
#include "cn_gemm_tc.cuh"

struct cn_policy {
  cn_policy_config cfg;
  int N, H, Win, M;
  int64_t launches;
  int attn_hpc;        // heads per CTA of the HH attention kernel
  bool finalized;
  std::map<std::string, std::vector<float>> host;
  std::vector<void*> allocs;
  size_t ws_allocs;   // allocs[0..ws_allocs) = workspace (kept); the rest = parameters of the last finalize
  // device parameters (kernel layouts)
  float *W1, *b1, *W2, *b2, *Wqkv, *bqkv, *Wos, *bos;
  float *Wr, *br, *Wet, *bet, *WsT, *bs, *Wa, *ba, *Wih, *bih, *Whh, *bhh, *Wo, *bo;
  float *Wac1, *bac1, *Wa2, *ba2, *Wc2, *bc2, *wv_, *bv, *Wm, *bm, *logstd;
  // tcgen05 path (gemm_mode 1): split-fp16 activations / weights and their TMA descriptors
  __half *e1h, *e1l, *e2h, *e2l, *aoh, *aol;
  __half *W2h, *W2l, *Wqkvh, *Wqkvl, *Wosh, *Wosl;
  CUtensorMap m_e1h, m_e1l, m_e2h, m_e2l, m_aoh, m_aol, m_W2h, m_W2l, m_Wqkvh, m_Wqkvl, m_Wosh, m_Wosl;
  // per-environment tail on tensor cores: output_linear, actor.0|critic.0, actor.2, critic.2
  __half *h1h, *h1l, *outh, *outl, *ac1h, *ac1l;
  __half *Woh, *Wol, *Wac1h, *Wac1l, *Wa2h, *Wa2l, *Wc2h, *Wc2l;
  CUtensorMap m_h1h, m_h1l, m_outh, m_outl, m_a1h, m_a1l, m_c1h, m_c1l;
  CUtensorMap m_Woh, m_Wol, m_Wac1h, m_Wac1l, m_Wa2h, m_Wa2l, m_Wc2h, m_Wc2l;
  // optional per-stage profiling
  bool profile;
  std::vector<cudaEvent_t> ev;
  // workspace
  int *row_start, *mc;
  float *x16, *e1, *e2, *qkv, *ao, *sout, *xr, *rs, *t1, *u, *wv, *h0, *gi, *gh, *outb, *ac1, *a2, *c2;
};

namespace {

int palloc(cn_policy* p, float** ptr, size_t count) {
  void* q = nullptr;
  cudaError_t err = cudaMalloc(&q, (count ? count : 4) * sizeof(float));
  if (err != cudaSuccess) return cn_set_error("cudaMalloc(%zu floats): %s", count, cudaGetErrorString(err));
  cudaMemset(q, 0, (count ? count : 4) * sizeof(float));
  p->allocs.push_back(q);
  *ptr = static_cast<float*>(q);
  return 0;
}

int upload(cn_policy* p, float** dst, const std::vector<float>& src) {
  int rc = palloc(p, dst, src.size());
  if (rc) return rc;
  cudaError_t err = cudaMemcpy(*dst, src.data(), src.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (err != cudaSuccess) return cn_set_error("H2D param: %s", cudaGetErrorString(err));
  return 0;
}

const std::vector<float>* get(cn_policy* p, const char* key, size_t count) {
  auto it = p->host.find(key);
  if (it == p->host.end()) { cn_set_error("cn_policy_finalize: parameter '%s' was not set", key); return nullptr; }
  if (it->second.size() != count) {
    cn_set_error("cn_policy_finalize: parameter '%s' has %zu elements, expected %zu", key, it->second.size(), count);
    return nullptr;
  }
  return &it->second;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(sym);
  }
  return fn;
}

// 2-D fp16 row-major [rows, K] tensor, box = 64 (K) x box_rows, 128-byte swizzle
int make_map(CUtensorMap* map, const __half* ptr, int rows, int K, int box_rows, int pitch = 0) {
  EncodeFn enc = get_encode();
  if (!enc) return cn_set_error("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)(pitch ? pitch : K) * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cn_set_error("cuTensorMapEncodeTiled failed (%d) rows=%d K=%d", (int)r, rows, K);
  return 0;
}

int halloc16(cn_policy* p, __half** ptr, size_t count) {
  float* q = nullptr;
  int rc = palloc(p, &q, (count + 1) / 2);
  *ptr = reinterpret_cast<__half*>(q);
  return rc;
}

// tcgen05 GEMM launch: C = act((Ahi+Alo)(Bhi+Blo)^T / scale + bias)
void gemm_tc(cn_policy* p, cudaStream_t st, const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh,
             const CUtensorMap& bl, int M, int N, int K, const float* bias, int act, float* c32, int ldc, __half* oh,
             __half* ol, int ldh, const int* m_ptr = nullptr) {
  TcEpilogue ep;
  ep.m_ptr = m_ptr;
  ep.bias = bias; ep.inv_scale = 1.0f / 64.0f; ep.act = act; ep.c32 = c32; ep.ldc = ldc; ep.out_hi = oh; ep.out_lo = ol;
  ep.ldh = ldh;
  dim3 grid(N / TC_BN, (M + TC_BM - 1) / TC_BM);
  cn_gemm_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(ah, al, bh, bl, M, N, K, ep);
  p->launches += 1;
}

void split16(cn_policy* p, cudaStream_t st, const float* src, float scale, __half* hi, __half* lo, size_t count) {
  cn_split_f16_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(src, scale, hi, lo, count);
  p->launches += 1;
}

void gemm(cn_policy* p, cudaStream_t st, const float* A, int lda, const float* W, int ldw, const float* bias,
          float* C, int ldc, int M, int N, int K, int act, int act_lo = 0, int act_hi = 1 << 30,
          const int* m_ptr = nullptr, __half* oh = nullptr, __half* ol = nullptr) {
  dim3 grid((N + CN_GEMM_BN - 1) / CN_GEMM_BN, (M + CN_GEMM_BM - 1) / CN_GEMM_BM);
  cn_gemm_f32_kernel<<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, C, ldc, M, N, K, act, act_lo, act_hi, m_ptr, oh, ol);
  p->launches += 1;
}

const char* kStageNames[] = {"pack_inputs", "embed1_gemm", "embed2_gemm", "qkv_gemm", "hh_attention",
                             "outproj_spatial_gemm", "robot_branch", "hr_attention", "gru", "actor_critic_heads"};
const int kNumStages = sizeof(kStageNames) / sizeof(kStageNames[0]);

inline void mark(cn_policy* p, cudaStream_t st, int i) {
  if (p->profile) cudaEventRecord(p->ev[i], st);
}

}  // namespace

extern "C" {

int cn_policy_profile(cn_policy* p, int enable) {
  if (!p) return cn_set_error("cn_policy_profile: null argument");
  cudaSetDevice(p->cfg.device);
  if (enable && p->ev.empty()) {
    p->ev.resize(kNumStages + 1);
    for (auto& e : p->ev) cudaEventCreate(&e);
  }
  p->profile = enable != 0;
  return 0;
}
int cn_policy_stage_count(void) { return kNumStages; }
const char* cn_policy_stage_name(int i) { return (i >= 0 && i < kNumStages) ? kStageNames[i] : ""; }
int cn_policy_stage_ms(cn_policy* p, float* out, int n) {
  if (!p || !out) return cn_set_error("cn_policy_stage_ms: null argument");
  if (p->ev.empty()) return cn_set_error("cn_policy_stage_ms: profiling was never enabled");
  cudaSetDevice(p->cfg.device);
  cudaError_t err = cudaEventSynchronize(p->ev[kNumStages]);
  if (err != cudaSuccess) return cn_set_error("cn_policy_stage_ms: %s", cudaGetErrorString(err));
  for (int i = 0; i < n && i < kNumStages; ++i) cudaEventElapsedTime(&out[i], p->ev[i], p->ev[i + 1]);
  return 0;
}

int cn_policy_create(const cn_policy_config* cfg, cn_policy** out) {
  if (!cfg || !out) return cn_set_error("cn_policy_create: null argument");
  *out = nullptr;
  if (cfg->num_envs <= 0 || cfg->human_num <= 0 || cfg->human_num > 128 || cfg->input_size <= 0 || cfg->input_size > 16)
    return cn_set_error("cn_policy_create: unsupported dims N=%d H=%d input=%d", cfg->num_envs, cfg->human_num,
                        cfg->input_size);
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0)
    return cn_set_error("cn_policy_create: no CUDA device (%s); this engine has no CPU fallback",
                        err == cudaSuccess ? "device count 0" : cudaGetErrorString(err));
  if (cfg->device < 0 || cfg->device >= ndev) return cn_set_error("cn_policy_create: bad device %d", cfg->device);
  cudaSetDevice(cfg->device);
  cn_policy* p = new cn_policy();
  memset(static_cast<void*>(&p->cfg), 0, sizeof(p->cfg));
  p->cfg = *cfg;
  p->N = cfg->num_envs; p->H = cfg->human_num; p->Win = cfg->input_size; p->M = p->N * p->H;
  p->launches = 0; p->finalized = false; p->profile = false;
  const size_t M = (size_t)p->M, N = (size_t)p->N;
  int rc = 0;
#define WS(name, count) if (!rc) rc = palloc(p, &p->name, (count))
  {
    float* q = nullptr;
    if (!rc) rc = palloc(p, &q, N + 2); p->row_start = reinterpret_cast<int*>(q);
    if (!rc) rc = palloc(p, &q, 4); p->mc = reinterpret_cast<int*>(q);
  }
  WS(x16, M * 16); WS(e1, M * 128); WS(e2, M * 512); WS(qkv, M * 1536); WS(ao, M * 512); WS(sout, M * 256);
  WS(xr, N * 16); WS(rs, N * 256); WS(t1, N * 128); WS(u, N * 256); WS(wv, N * 256); WS(h0, N * 128);
  WS(gi, N * 384); WS(gh, N * 384); WS(outb, N * 256); WS(ac1, N * 512); WS(a2, N * 256); WS(c2, N * 256);
#undef WS
  if (!rc && cfg->gemm_mode == 1) {
    rc = halloc16(p, &p->e1h, M * 128);
    if (!rc) rc = halloc16(p, &p->e1l, M * 128);
    if (!rc) rc = halloc16(p, &p->e2h, M * 512);
    if (!rc) rc = halloc16(p, &p->e2l, M * 512);
    if (!rc) rc = halloc16(p, &p->aoh, M * 512);
    if (!rc) rc = halloc16(p, &p->aol, M * 512);
    if (!rc) rc = make_map(&p->m_e1h, p->e1h, p->M, 128, TC_BM);
    if (!rc) rc = make_map(&p->m_e1l, p->e1l, p->M, 128, TC_BM);
    if (!rc) rc = make_map(&p->m_e2h, p->e2h, p->M, 512, TC_BM);
    if (!rc) rc = make_map(&p->m_e2l, p->e2l, p->M, 512, TC_BM);
    if (!rc) rc = make_map(&p->m_aoh, p->aoh, p->M, 512, TC_BM);
    if (!rc) rc = make_map(&p->m_aol, p->aol, p->M, 512, TC_BM);
    if (!rc) rc = halloc16(p, &p->h1h, N * 128);
    if (!rc) rc = halloc16(p, &p->h1l, N * 128);
    if (!rc) rc = halloc16(p, &p->outh, N * 256);
    if (!rc) rc = halloc16(p, &p->outl, N * 256);
    if (!rc) rc = halloc16(p, &p->ac1h, N * 512);
    if (!rc) rc = halloc16(p, &p->ac1l, N * 512);
    if (!rc) rc = make_map(&p->m_h1h, p->h1h, p->N, 128, TC_BM);
    if (!rc) rc = make_map(&p->m_h1l, p->h1l, p->N, 128, TC_BM);
    if (!rc) rc = make_map(&p->m_outh, p->outh, p->N, 256, TC_BM);
    if (!rc) rc = make_map(&p->m_outl, p->outl, p->N, 256, TC_BM);
    if (!rc) rc = make_map(&p->m_a1h, p->ac1h, p->N, 256, TC_BM, 512);          // actor half: cols 0..255
    if (!rc) rc = make_map(&p->m_a1l, p->ac1l, p->N, 256, TC_BM, 512);
    if (!rc) rc = make_map(&p->m_c1h, p->ac1h + 256, p->N, 256, TC_BM, 512);    // critic half: cols 256..511
    if (!rc) rc = make_map(&p->m_c1l, p->ac1l + 256, p->N, 256, TC_BM, 512);
    if (!rc) {
      cudaError_t e2 = cudaFuncSetAttribute(cn_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
      if (e2 != cudaSuccess) rc = cn_set_error("cudaFuncSetAttribute(tc): %s", cudaGetErrorString(e2));
    }
  }
  if (rc) { cn_policy_destroy(p); return rc; }
  p->ws_allocs = p->allocs.size();
  p->attn_hpc = 8;
  while (p->attn_hpc > 1 && p->attn_hpc * ((size_t)p->H * 129 + 64) * sizeof(float) > 200 * 1024) p->attn_hpc /= 2;
  const size_t attn_smem = p->attn_hpc * ((size_t)p->H * 129 + 64) * sizeof(float);
  err = cudaFuncSetAttribute(cn_hh_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem);
  if (err != cudaSuccess) { cn_policy_destroy(p); return cn_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(err)); }
  *out = p;
  return 0;
}

int cn_policy_destroy(cn_policy* p) {
  if (!p) return 0;
  cudaSetDevice(p->cfg.device);
  for (void* q : p->allocs) cudaFree(q);
  for (auto& e : p->ev) cudaEventDestroy(e);
  delete p;
  return 0;
}

int cn_policy_set_param(cn_policy* p, const char* key, const float* h_data, size_t count) {
  if (!p || !key || !h_data) return cn_set_error("cn_policy_set_param: null argument");
  p->host[key] = std::vector<float>(h_data, h_data + count);
  p->finalized = false;
  return 0;
}

int cn_policy_finalize(cn_policy* p, void* stream) {
  if (!p) return cn_set_error("cn_policy_finalize: null argument");
  cudaSetDevice(p->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  const int Win = p->Win;
#define GET(var, key, count) const std::vector<float>* var = get(p, key, (count)); if (!var) return 1
  GET(w1, "base.spatial_attn.embedding_layer.0.weight", (size_t)128 * Win);
  GET(b1, "base.spatial_attn.embedding_layer.0.bias", 128);
  GET(w2, "base.spatial_attn.embedding_layer.2.weight", (size_t)512 * 128);
  GET(b2, "base.spatial_attn.embedding_layer.2.bias", 512);
  GET(wq, "base.spatial_attn.q_linear.weight", (size_t)512 * 512);
  GET(bq, "base.spatial_attn.q_linear.bias", 512);
  GET(wk, "base.spatial_attn.k_linear.weight", (size_t)512 * 512);
  GET(bk, "base.spatial_attn.k_linear.bias", 512);
  GET(wvv, "base.spatial_attn.v_linear.weight", (size_t)512 * 512);
  GET(bvv, "base.spatial_attn.v_linear.bias", 512);
  GET(win, "base.spatial_attn.multihead_attn.in_proj_weight", (size_t)1536 * 512);
  GET(bin, "base.spatial_attn.multihead_attn.in_proj_bias", 1536);
  GET(wout, "base.spatial_attn.multihead_attn.out_proj.weight", (size_t)512 * 512);
  GET(bout, "base.spatial_attn.multihead_attn.out_proj.bias", 512);
  GET(wsl, "base.spatial_linear.0.weight", (size_t)256 * 512);
  GET(bsl, "base.spatial_linear.0.bias", 256);
  GET(wr, "base.robot_linear.0.weight", (size_t)256 * 9);
  GET(br, "base.robot_linear.0.bias", 256);
  GET(wt, "base.attn.temporal_edge_layer.0.weight", (size_t)64 * 256);
  GET(bt, "base.attn.temporal_edge_layer.0.bias", 64);
  GET(ws, "base.attn.spatial_edge_layer.0.weight", (size_t)64 * 256);
  GET(bs, "base.attn.spatial_edge_layer.0.bias", 64);
  GET(we, "base.humanNodeRNN.encoder_linear.weight", (size_t)64 * 256);
  GET(be, "base.humanNodeRNN.encoder_linear.bias", 64);
  GET(wa, "base.humanNodeRNN.edge_attention_embed.weight", (size_t)64 * 256);
  GET(ba, "base.humanNodeRNN.edge_attention_embed.bias", 64);
  GET(wih, "base.humanNodeRNN.gru.weight_ih_l0", (size_t)384 * 128);
  GET(whh, "base.humanNodeRNN.gru.weight_hh_l0", (size_t)384 * 128);
  GET(bih, "base.humanNodeRNN.gru.bias_ih_l0", 384);
  GET(bhh, "base.humanNodeRNN.gru.bias_hh_l0", 384);
  GET(wo, "base.humanNodeRNN.output_linear.weight", (size_t)256 * 128);
  GET(bo, "base.humanNodeRNN.output_linear.bias", 256);
  GET(wa0, "base.actor.0.weight", (size_t)256 * 256);
  GET(ba0, "base.actor.0.bias", 256);
  GET(wa2, "base.actor.2.weight", (size_t)256 * 256);
  GET(ba2, "base.actor.2.bias", 256);
  GET(wc0, "base.critic.0.weight", (size_t)256 * 256);
  GET(bc0, "base.critic.0.bias", 256);
  GET(wc2, "base.critic.2.weight", (size_t)256 * 256);
  GET(bc2, "base.critic.2.bias", 256);
  GET(wcl, "base.critic_linear.weight", 256);
  GET(bcl, "base.critic_linear.bias", 1);
  GET(wm, "dist.fc_mean.weight", (size_t)2 * 256);
  GET(bm, "dist.fc_mean.bias", 2);
  GET(ls, "dist.logstd._bias", 2);
#undef GET
  // release device parameters of a previous finalize (workspace allocations come first and stay)
  cudaStreamSynchronize(st);
  for (size_t i = p->ws_allocs; i < p->allocs.size(); ++i) cudaFree(p->allocs[i]);
  p->allocs.resize(p->ws_allocs);
  int rc = 0;
  auto pad_k = [](const std::vector<float>& w, int rows, int k, int kp) {
    std::vector<float> o((size_t)rows * kp, 0.0f);
    for (int r = 0; r < rows; ++r) for (int c = 0; c < k; ++c) o[(size_t)r * kp + c] = w[(size_t)r * k + c];
    return o;
  };
  auto cat = [](const std::vector<float>& a, const std::vector<float>& b) {
    std::vector<float> o(a); o.insert(o.end(), b.begin(), b.end()); return o;
  };
#define UP(dst, vec) if (!rc) rc = upload(p, &p->dst, (vec))
  UP(W1, pad_k(*w1, 128, Win, 16)); UP(b1, *b1); UP(W2, *w2); UP(b2, *b2);
  UP(Wr, pad_k(*wr, 256, 9, 16)); UP(br, *br);
  UP(Wet, cat(*we, *wt)); UP(bet, cat(*be, *bt));           // rows 0..63 encoder_linear, 64..127 temporal_edge_layer
  {
    std::vector<float> wst((size_t)256 * 64);                // W_s^T: [256][64]
    for (int r = 0; r < 64; ++r) for (int c = 0; c < 256; ++c) wst[(size_t)c * 64 + r] = (*ws)[(size_t)r * 256 + c];
    UP(WsT, wst);
  }
  UP(bs, *bs); UP(Wa, *wa); UP(ba, *ba); UP(Wih, *wih); UP(bih, *bih); UP(Whh, *whh); UP(bhh, *bhh);
  UP(Wo, *wo); UP(bo, *bo);
  UP(Wac1, cat(*wa0, *wc0)); UP(bac1, cat(*ba0, *bc0));      // rows 0..255 actor.0, 256..511 critic.0
  UP(Wa2, *wa2); UP(ba2, *ba2); UP(Wc2, *wc2); UP(bc2, *bc2);
  UP(wv_, *wcl); UP(bv, *bcl); UP(Wm, *wm); UP(bm, *bm); UP(logstd, *ls);
  // folded projections
  float *d_win = nullptr, *d_bin = nullptr, *d_wl[3] = {nullptr, nullptr, nullptr}, *d_bl[3] = {nullptr, nullptr, nullptr};
  float *d_wout = nullptr, *d_bout = nullptr, *d_wsl = nullptr, *d_bsl = nullptr;
  if (!rc) rc = upload(p, &d_win, *win);
  if (!rc) rc = upload(p, &d_bin, *bin);
  const std::vector<float>* wl[3] = {wq, wk, wvv};
  const std::vector<float>* bl[3] = {bq, bk, bvv};
  for (int i = 0; i < 3 && !rc; ++i) { rc = upload(p, &d_wl[i], *wl[i]); if (!rc) rc = upload(p, &d_bl[i], *bl[i]); }
  if (!rc) rc = upload(p, &d_wout, *wout);
  if (!rc) rc = upload(p, &d_bout, *bout);
  if (!rc) rc = upload(p, &d_wsl, *wsl);
  if (!rc) rc = upload(p, &d_bsl, *bsl);
  if (!rc) rc = palloc(p, &p->Wqkv, (size_t)1536 * 512);
  if (!rc) rc = palloc(p, &p->bqkv, 1536);
  if (!rc) rc = palloc(p, &p->Wos, (size_t)256 * 512);
  if (!rc) rc = palloc(p, &p->bos, 256);
  if (rc) return rc;
  for (int i = 0; i < 3; ++i) {
    // Wf_i = Win_i (512x512) @ Wl_i (512x512);  bf_i = Win_i @ bl_i + bin_i
    cn_fold_mm_kernel<<<dim3(4, 512), 128, 0, st>>>(d_win + (size_t)i * 512 * 512, d_wl[i],
                                                     p->Wqkv + (size_t)i * 512 * 512, 512, 512, 512);
    cn_fold_mv_kernel<<<4, 128, 0, st>>>(d_win + (size_t)i * 512 * 512, d_bl[i], d_bin + i * 512, p->bqkv + i * 512, 512, 512);
  }
  // Wos = Wsl (256x512) @ Wout (512x512);  bos = Wsl @ bout + bsl
  cn_fold_mm_kernel<<<dim3(4, 256), 128, 0, st>>>(d_wsl, d_wout, p->Wos, 256, 512, 512);
  cn_fold_mv_kernel<<<2, 128, 0, st>>>(d_wsl, d_bout, d_bsl, p->bos, 256, 512);
  if (p->cfg.gemm_mode == 1) {
    // fp16 (hi, lo) split of the tensor-core weights, pre-scaled by 2^6 (exact) so lo stays normal
    if (!rc) rc = halloc16(p, &p->W2h, (size_t)512 * 128);
    if (!rc) rc = halloc16(p, &p->W2l, (size_t)512 * 128);
    if (!rc) rc = halloc16(p, &p->Wqkvh, (size_t)1536 * 512);
    if (!rc) rc = halloc16(p, &p->Wqkvl, (size_t)1536 * 512);
    if (!rc) rc = halloc16(p, &p->Wosh, (size_t)256 * 512);
    if (!rc) rc = halloc16(p, &p->Wosl, (size_t)256 * 512);
    if (rc) return rc;
    split16(p, st, p->W2, 64.0f, p->W2h, p->W2l, (size_t)512 * 128);
    split16(p, st, p->Wqkv, 64.0f, p->Wqkvh, p->Wqkvl, (size_t)1536 * 512);
    split16(p, st, p->Wos, 64.0f, p->Wosh, p->Wosl, (size_t)256 * 512);
    {
      struct { float* src; __half** hi; __half** lo; CUtensorMap* mh; CUtensorMap* ml; int rows, k; } tw[4] = {
          {p->Wo, &p->Woh, &p->Wol, &p->m_Woh, &p->m_Wol, 256, 128},
          {p->Wac1, &p->Wac1h, &p->Wac1l, &p->m_Wac1h, &p->m_Wac1l, 512, 256},
          {p->Wa2, &p->Wa2h, &p->Wa2l, &p->m_Wa2h, &p->m_Wa2l, 256, 256},
          {p->Wc2, &p->Wc2h, &p->Wc2l, &p->m_Wc2h, &p->m_Wc2l, 256, 256}};
      for (auto& t : tw) {
        if (!rc) rc = halloc16(p, t.hi, (size_t)t.rows * t.k);
        if (!rc) rc = halloc16(p, t.lo, (size_t)t.rows * t.k);
        if (rc) return rc;
        split16(p, st, t.src, 64.0f, *t.hi, *t.lo, (size_t)t.rows * t.k);
        rc = make_map(t.mh, *t.hi, t.rows, t.k, TC_BN);
        if (!rc) rc = make_map(t.ml, *t.lo, t.rows, t.k, TC_BN);
        if (rc) return rc;
      }
    }
    rc = make_map(&p->m_W2h, p->W2h, 512, 128, TC_BN);
    if (!rc) rc = make_map(&p->m_W2l, p->W2l, 512, 128, TC_BN);
    if (!rc) rc = make_map(&p->m_Wqkvh, p->Wqkvh, 1536, 512, TC_BN);
    if (!rc) rc = make_map(&p->m_Wqkvl, p->Wqkvl, 1536, 512, TC_BN);
    if (!rc) rc = make_map(&p->m_Wosh, p->Wosh, 256, 512, TC_BN);
    if (!rc) rc = make_map(&p->m_Wosl, p->Wosl, 256, 512, TC_BN);
    if (rc) return rc;
  }
  cudaError_t err = cudaStreamSynchronize(st);
  if (err != cudaSuccess) return cn_set_error("cn_policy_finalize: %s", cudaGetErrorString(err));
  p->finalized = true;
  return 0;
}

int cn_policy_act(cn_policy* p, const cn_act_ptrs* d, void* stream) {
  if (!p || !d) return cn_set_error("cn_policy_act: null argument");
  if (!p->finalized) return cn_set_error("cn_policy_act: call cn_policy_finalize after setting parameters");
  if (!d->robot_node || !d->temporal_edges || !d->spatial_edges || !d->detected_human_num || !d->h_in || !d->masks ||
      !d->value || !d->action || !d->log_prob || !d->h_out)
    return cn_set_error("cn_policy_act: missing input/output pointer");
  cudaSetDevice(p->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  const int N = p->N, H = p->H, M = p->M;
  mark(p, st, 0);
  // 0. compaction offsets, pack / pad inputs, h0 = h * mask
  {
    cn_row_offsets_kernel<<<1, 1024, 0, st>>>(d->detected_human_num, N, H, p->row_start, p->mc);
    const int total = M * 16 > N * 128 ? M * 16 : N * 128;
    cn_pack_inputs_kernel<<<(total + 255) / 256, 256, 0, st>>>(d->spatial_edges, p->Win, H, N, p->row_start, p->x16,
                                                               d->temporal_edges, d->robot_node, d->h_in, d->masks, p->xr,
                                                               p->h0);
    p->launches += 2;
  }
  // 1. human-human branch over the Mc = sum_e n_e valid rows (device-side count p->mc)
  const bool tcm = p->cfg.gemm_mode == 1;
  const int* mc = p->mc;
  mark(p, st, 1);
  if (tcm) gemm(p, st, p->x16, 16, p->W1, 16, p->b1, nullptr, 128, M, 128, 16, CN_ACT_RELU, 0, 1 << 30, mc, p->e1h, p->e1l);
  else gemm(p, st, p->x16, 16, p->W1, 16, p->b1, p->e1, 128, M, 128, 16, CN_ACT_RELU, 0, 1 << 30, mc);
  mark(p, st, 2);
  if (tcm) gemm_tc(p, st, p->m_e1h, p->m_e1l, p->m_W2h, p->m_W2l, M, 512, 128, p->b2, CN_ACT_RELU, nullptr, 0, p->e2h, p->e2l, 512, mc);
  else gemm(p, st, p->e1, 128, p->W2, 128, p->b2, p->e2, 512, M, 512, 128, CN_ACT_RELU, 0, 1 << 30, mc);
  mark(p, st, 3);
  if (tcm) gemm_tc(p, st, p->m_e2h, p->m_e2l, p->m_Wqkvh, p->m_Wqkvl, M, 1536, 512, p->bqkv, CN_ACT_NONE, p->qkv, 1536, nullptr, nullptr, 0, mc);
  else gemm(p, st, p->e2, 512, p->Wqkv, 512, p->bqkv, p->qkv, 1536, M, 1536, 512, CN_ACT_NONE, 0, 1 << 30, mc);
  mark(p, st, 4);
  {
    const size_t smem = p->attn_hpc * ((size_t)H * 129 + 64) * sizeof(float);
    cn_hh_attention_kernel<<<dim3(N, 8 / p->attn_hpc), p->attn_hpc * 32, smem, st>>>(p->qkv, p->row_start, H, tcm ? nullptr : p->ao,
                                                          tcm ? p->aoh : nullptr, tcm ? p->aol : nullptr);
    p->launches += 1;
  }
  mark(p, st, 5);
  if (tcm) gemm_tc(p, st, p->m_aoh, p->m_aol, p->m_Wosh, p->m_Wosl, M, 256, 512, p->bos, CN_ACT_RELU, p->sout, 256, nullptr, nullptr, 0, mc);
  else gemm(p, st, p->ao, 512, p->Wos, 512, p->bos, p->sout, 256, M, 256, 512, CN_ACT_RELU, 0, 1 << 30, mc);
  // 2. robot branch
  mark(p, st, 6);
  gemm(p, st, p->xr, 16, p->Wr, 16, p->br, p->rs, 256, N, 256, 16, CN_ACT_RELU);
  gemm(p, st, p->rs, 256, p->Wet, 256, p->bet, p->t1, 128, N, 128, 256, CN_ACT_RELU, 0, 64);   // [enc | te]
  gemm(p, st, p->t1 + 64, 128, p->WsT, 64, nullptr, p->u, 256, N, 256, 64, CN_ACT_NONE);        // u = W_s^T te
  mark(p, st, 7);
  cn_hr_attention_kernel<<<(N + 3) / 4, 128, 0, st>>>(p->sout, p->u, p->t1, 128, 64, p->bs, p->row_start, N, H,
                                                      p->wv);
  p->launches += 1;
  mark(p, st, 8);
  // emb overwrites the te half of t1 -> t1 = [enc | emb] = GRU input
  gemm(p, st, p->wv, 256, p->Wa, 256, p->ba, p->t1 + 64, 128, N, 64, 256, CN_ACT_RELU);
  gemm(p, st, p->t1, 128, p->Wih, 128, p->bih, p->gi, 384, N, 384, 128, CN_ACT_NONE);
  gemm(p, st, p->h0, 128, p->Whh, 128, p->bhh, p->gh, 384, N, 384, 128, CN_ACT_NONE);
  cn_gru_gate_kernel<<<(N * 128 + 255) / 256, 256, 0, st>>>(p->gi, p->gh, p->h0, N, d->h_out, tcm ? p->h1h : nullptr,
                                                            tcm ? p->h1l : nullptr);
  p->launches += 1;
  mark(p, st, 9);
  if (tcm) {
    gemm_tc(p, st, p->m_h1h, p->m_h1l, p->m_Woh, p->m_Wol, N, 256, 128, p->bo, CN_ACT_NONE, nullptr, 0, p->outh, p->outl, 256);
    gemm_tc(p, st, p->m_outh, p->m_outl, p->m_Wac1h, p->m_Wac1l, N, 512, 256, p->bac1, CN_ACT_TANH, nullptr, 0, p->ac1h, p->ac1l, 512);
    gemm_tc(p, st, p->m_a1h, p->m_a1l, p->m_Wa2h, p->m_Wa2l, N, 256, 256, p->ba2, CN_ACT_TANH, p->a2, 256, nullptr, nullptr, 0);
    gemm_tc(p, st, p->m_c1h, p->m_c1l, p->m_Wc2h, p->m_Wc2l, N, 256, 256, p->bc2, CN_ACT_TANH, p->c2, 256, nullptr, nullptr, 0);
  } else {
    gemm(p, st, d->h_out, 128, p->Wo, 128, p->bo, p->outb, 256, N, 256, 128, CN_ACT_NONE);
    gemm(p, st, p->outb, 256, p->Wac1, 256, p->bac1, p->ac1, 512, N, 512, 256, CN_ACT_TANH);      // [actor.0 | critic.0]
    gemm(p, st, p->ac1, 512, p->Wa2, 256, p->ba2, p->a2, 256, N, 256, 256, CN_ACT_TANH);
    gemm(p, st, p->ac1 + 256, 512, p->Wc2, 256, p->bc2, p->c2, 256, N, 256, 256, CN_ACT_TANH);
  }
  cn_heads_kernel<<<(N + 3) / 4, 128, 0, st>>>(p->a2, 256, p->c2, 256, p->wv_, p->bv, p->Wm, p->bm, p->logstd, d->noise, N,
                                               d->value, d->action, d->log_prob, d->action_mean);
  p->launches += 1;
  mark(p, st, kNumStages);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return cn_set_error("cn_policy_act launch: %s", cudaGetErrorString(err));
  return 0;
}

int64_t cn_policy_launch_count(cn_policy* p) { return p ? p->launches : 0; }

int64_t cn_policy_last_rows(cn_policy* p) {
  if (!p) return -1;
  cudaSetDevice(p->cfg.device);
  int v = 0;
  if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(&v, p->mc, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  return v;
}

// Internal test hook (not part of the public header): C = act(A[M,K] W[N,K]^T + bias) through the
// tcgen05 3xFP16 kernel, fp32 device pointers in/out.  Used by tests/test_gpu_gemm_tc.py.
int cn_internal_gemm_tc(const float* dA, const float* dW, const float* dbias, float* dC, int M, int N, int K, int act) {
  if (N % TC_BN || K % TC_BK) return cn_set_error("cn_internal_gemm_tc: need N %% 256 == 0 and K %% 64 == 0");
  cn_policy tmp;
  tmp.launches = 0;
  __half *ah, *al, *bh, *bl;
  int rc = halloc16(&tmp, &ah, (size_t)M * K);
  if (!rc) rc = halloc16(&tmp, &al, (size_t)M * K);
  if (!rc) rc = halloc16(&tmp, &bh, (size_t)N * K);
  if (!rc) rc = halloc16(&tmp, &bl, (size_t)N * K);
  CUtensorMap mah, mal, mbh, mbl;
  if (!rc) rc = make_map(&mah, ah, M, K, TC_BM);
  if (!rc) rc = make_map(&mal, al, M, K, TC_BM);
  if (!rc) rc = make_map(&mbh, bh, N, K, TC_BN);
  if (!rc) rc = make_map(&mbl, bl, N, K, TC_BN);
  if (!rc) {
    cudaFuncSetAttribute(cn_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    split16(&tmp, 0, dA, 1.0f, ah, al, (size_t)M * K);
    split16(&tmp, 0, dW, 64.0f, bh, bl, (size_t)N * K);
    gemm_tc(&tmp, 0, mah, mal, mbh, mbl, M, N, K, dbias, act, dC, N, nullptr, nullptr, 0);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) rc = cn_set_error("cn_internal_gemm_tc: %s", cudaGetErrorString(err));
  }
  for (void* q : tmp.allocs) cudaFree(q);
  return rc;
}

}  // extern "C"
