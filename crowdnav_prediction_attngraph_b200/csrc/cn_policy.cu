// Host side of the attention-graph policy forward: parameter upload / folding and the
// per-rollout-step launch sequence behind cn_policy_act (replaces Policy.act,
// rl/networks/model.py:56-74, for base = selfAttn_merge_SRNN with sort_humans = True).
//
// Algebraic folds done once per parameter upload (fp64 accumulate, exact same function):
//   q/k/v_linear  o  MultiheadAttention.in_proj   ->  one 512 -> 1536 projection
//   MultiheadAttention.out_proj  o  spatial_linear -> one 512 -> 256 projection (ReLU after)
//   spatial_edge_layer folded into the robot side of the dot-product attention (u = W_s^T te)
//
// gemm_mode 0: every layer on the fp32 CUDA-core GEMM (cn_gemm_f32_kernel).
// gemm_mode 1: every layer with K >= 64 on the tcgen05 3xFP16 GEMM (cn_gemm_tc_kernel): activations
//              travel between layers as (hi, lo) fp16 pairs written by the producing kernel's epilogue.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/crowdnav_b200.h"
#include "cn_host_util.h"
#include "cn_policy_kernels.cuh"
#include "cn_gemm_tc.cuh"
#include "cn_qkv_attn.cuh"

// A split-fp16 matrix [rows, K] (row pitch `pitch` elements) and its TMA descriptors.
struct TcMat {
  __half *hi = nullptr, *lo = nullptr;
  CUtensorMap mh, ml;
  int pitch = 0;
};

struct cn_policy {
  cn_policy_config cfg;
  int N, H, Win, M;
  int64_t launches;
  int num_sms;
  bool pdl;           // programmatic dependent launch along the kernel chain (CN_PDL=0 disables)
  long dbg_launch_idx = 0;   // launch index within the current step (CN_PDL_WINDOW debugging)
  bool launch_error;  // a launch or a GEMM output map failed (cn_last_error has the stage and the reason)
  const char* cur_stage = nullptr;   // stage name of the launches being enqueued (error reports)
  bool fuse_qkv;      // QKV projection + human-human attention in ONE kernel (cn_qkv_attn.cuh; opt-in, CN_FUSE_QKV=1)
  TcMat tWqkvH;       // folded QKV weight, rows head-major: [8][Q 64 | K 64 | V 64][512]
  CUtensorMap qa_ah, qa_al, qa_bh, qa_bl;   // 32-wide (SWIZZLE_64B) boxes of tE2 and tWqkvH for the fused kernel
  float* bqkvH = nullptr;
  int* tile_tab = nullptr;   // row tiles of the fused kernel (cn_qkv_tiles_kernel)
  cudaEvent_t ev_tiles;
  // human-human attention instance: R queries of an environment per warp (CN_ATTN_R = 1 (default), 2 or 4; measured on
  // B200 at 4096 envs: 0.065 / 0.070 / 0.085 ms -- sharing K / V rows between queries does not pay, the kernel is bound by
  // load latency at ~4 keys per query, not by L1 delivery)
  void (*attn_kernel)(const float*, const int*, const int*, const int*, const int*, float*, __half*, __half*);
  int attn_warps;
  int qkv_chunks;     // 1 (default): single pass; 2 (CN_QKV_CHUNKS=2): QKV + attention in two row chunks with overlap
  bool finalized;
  std::map<std::string, std::vector<float>> host;
  std::vector<void*> allocs;
  size_t ws_allocs;   // allocs[0..ws_allocs) = workspace (kept); the rest = parameters of the last finalize
  // device parameters (fp32 kernel layouts)
  float *W1, *b1, *W2, *b2, *Wqkv, *bqkv, *Wos, *bos;
  float *Wr, *br, *Wet, *bet, *WsT, *bs, *Wa, *ba, *Wih, *bih, *Whh, *bhh, *Wo, *bo;
  float *Woac, *boac;      // (actor.0 | critic.0) o output_linear folded: 128 -> 512
  float *Wac1, *bac1, *Wa2, *ba2, *Wc2, *bc2, *wv_, *bv, *Wm, *bm, *logstd;
  // tcgen05 path: split weights (B operands; tile rows = 256 for per-human layers, 64 for per-env layers)
  TcMat tW2, tWqkv, tWos, tWet, tWsT, tWa, tWih, tWhh, tWo, tWac1, tWoac, tWa2, tWc2;
  // tcgen05 path: split activations (A operands)
  TcMat tE1, tE2, tAo;                                  // per human rows
  TcMat tRs, tT1, tTe, tWv, tH0, tH1, tOut, tAc1, tA1, tC1;   // per environment rows (tTe / tA1 / tC1 = column views)
  // second stream: the robot branch / gh / critic.2 are independent of the per-human chain and run
  // concurrently with it (fork / join with events; capturable in a CUDA graph)
  cudaStream_t st2, st3;
  cudaEvent_t ev_fork, ev_join, ev_fork2, ev_join2, ev_fork3, ev_join3;
  // optional per-stage profiling
  bool profile;
  std::vector<cudaEvent_t> ev;
  // cached TMA store maps of GEMM outputs: key = (pointer, element size, columns, rows, leading dimension)
  struct OutMap { const void* ptr; int esize, cols, rows, ld; CUtensorMap map; };
  std::vector<OutMap*> omaps;
  // workspace
  int *row_start, *row_env, *mc;
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

// 2-D fp16 row-major [rows, K] tensor with row pitch `pitch`, box = 64 (K) x box_rows, 128-byte swizzle
// (box_k = 32: 64-byte rows with SWIZZLE_64B, the half-width K blocks of cn_qkv_attn.cuh)
int make_map(CUtensorMap* map, const __half* ptr, int rows, int K, int box_rows, int pitch, int box_k = TC_BK) {
  EncodeFn enc = get_encode();
  if (!enc) return cn_set_error("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_k == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cn_set_error("cuTensorMapEncodeTiled failed (%d) rows=%d K=%d", (int)r, rows, K);
  return 0;
}

// TMA store map of a row-major output [rows, cols] (leading dimension ld elements): box 32 x 32, swizzle
// matching the epilogue's staging layout (fp32: 128-byte rows -> SWIZZLE_128B; fp16: 64-byte rows -> SWIZZLE_64B)
const CUtensorMap* out_map(cn_policy* p, const void* ptr, int esize, int cols, int rows, int ld) {
  for (auto* m : p->omaps)
    if (m->ptr == ptr && m->esize == esize && m->cols == cols && m->rows == rows && m->ld == ld) return &m->map;
  EncodeFn enc = get_encode();
  if (!enc) { cn_set_error("cuTensorMapEncodeTiled entry point not available"); return nullptr; }
  auto* m = new cn_policy::OutMap();
  m->ptr = ptr; m->esize = esize; m->cols = cols; m->rows = rows; m->ld = ld;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * (cuuint64_t)esize};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m->map, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   esize == 4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    cn_set_error("cuTensorMapEncodeTiled(out) failed (%d) cols=%d rows=%d ld=%d esize=%d", (int)r, cols, rows, ld, esize);
    delete m;
    return nullptr;
  }
  p->omaps.push_back(m);
  return &m->map;
}

int halloc16(cn_policy* p, __half** ptr, size_t count) {
  float* q = nullptr;
  int rc = palloc(p, &q, (count + 1) / 2);
  *ptr = reinterpret_cast<__half*>(q);
  return rc;
}

// allocate a split matrix [rows, K] and build its maps (box_rows = 128 for A operands, BN for B operands)
int tc_alloc(cn_policy* p, TcMat& t, int rows, int K, int box_rows) {
  int rc = halloc16(p, &t.hi, (size_t)rows * K);
  if (!rc) rc = halloc16(p, &t.lo, (size_t)rows * K);
  t.pitch = K;
  if (!rc) rc = make_map(&t.mh, t.hi, rows, K, box_rows, K);
  if (!rc) rc = make_map(&t.ml, t.lo, rows, K, box_rows, K);
  return rc;
}
// view of columns [col0, col0 + K) of an existing split matrix
int tc_view(TcMat& v, const TcMat& src, int col0, int rows, int K, int box_rows) {
  v.hi = src.hi + col0; v.lo = src.lo + col0; v.pitch = src.pitch;
  int rc = make_map(&v.mh, v.hi, rows, K, box_rows, src.pitch);
  if (!rc) rc = make_map(&v.ml, v.lo, rows, K, box_rows, src.pitch);
  return rc;
}

// Kernel launch with (optionally) programmatic dependent launch: the kernel may be scheduled before its
// predecessor in the stream has finished; every kernel of the chain calls griddepcontrol.wait before touching
// global memory (cn_pdl_prologue / cn_pdl_wait), so the data dependencies are unchanged.
template <typename... KArgs, typename... Args>
void launch_k(cn_policy* p, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  // debug aid: CN_PDL_WINDOW=lo:hi keeps the attribute only for launches lo <= index < hi of the context
  static int win_lo = -1, win_hi = -1;
  if (win_lo < 0) {
    const char* w = getenv("CN_PDL_WINDOW");
    win_lo = 0; win_hi = 1 << 30;
    if (w) sscanf(w, "%d:%d", &win_lo, &win_hi);
  }
  const long idx = p->dbg_launch_idx++;
  cfg.attrs = at; cfg.numAttrs = (p->pdl && idx >= win_lo && idx < win_hi) ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  if (e != cudaSuccess && !p->launch_error) {     // keep the FIRST failure and the stage it happened in
    p->launch_error = true;
    cn_set_error("kernel launch failed in stage '%s' (launch #%lld of this handle): %s",
                 p->cur_stage ? p->cur_stage : "?", (long long)p->launches, cudaGetErrorString(e));
  }
  p->launches += 1;
}

void split16(cn_policy* p, cudaStream_t st, const float* src, float scale, __half* hi, __half* lo, size_t count) {
  cn_split_f16_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(src, scale, hi, lo, count);
  p->launches += 1;
}

// tcgen05 GEMM launch: C = act((Ahi+Alo)(Bhi+Blo)^T / 64 + bias); bn = B tile rows (256 or 64)
struct TcOut {
  float* c32 = nullptr; int ldc = 0;
  __half *oh = nullptr, *ol = nullptr; int ldh = 0;
};
void gemm_tc(cn_policy* p, cudaStream_t st, const TcMat& A, const TcMat& B, int M, int N, int K, int bn, const float* bias,
             int act, const TcOut& o, const int* m_ptr = nullptr, int act_lo = 0, int act_hi = 1 << 30,
             const int* m0_ptr = nullptr) {
  TcEpilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.bias = bias; ep.inv_scale = 1.0f / 64.0f; ep.act = act; ep.act_lo = act_lo; ep.act_hi = act_hi;
  ep.c32 = o.c32; ep.ldc = o.ldc; ep.out_hi = o.oh; ep.out_lo = o.ol; ep.ldh = o.ldh; ep.m_ptr = m_ptr; ep.m0_ptr = m0_ptr;
  {
    static const int nostore = (getenv("CN_DBG_NOSTORE") && getenv("CN_DBG_NOSTORE")[0] == '1') ? 1 : 0;
    ep.dbg_nostore = nostore;
  }
  // persistent: one CTA per SM at most; tiles beyond the device-side row count are never touched
  const int tiles = (N / bn) * ((M + TC_BM - 1) / TC_BM);
  dim3 grid(tiles < p->num_sms ? tiles : p->num_sms);
  // TMA store maps of the outputs (cached; unused ones alias an operand map and are never dereferenced)
  const CUtensorMap* mc32 = o.c32 ? out_map(p, o.c32, 4, N, M, o.ldc) : &A.mh;
  const CUtensorMap* mhi = o.oh ? out_map(p, o.oh, 2, N, M, o.ldh) : &A.mh;
  const CUtensorMap* mlo = o.ol ? out_map(p, o.ol, 2, N, M, o.ldh) : &A.mh;
  if (!mc32 || !mhi || !mlo) { p->launch_error = true; return; }
  if (bn == 256)
    launch_k(p, cn_gemm_tc_kernel<256>, grid, dim3(TC_THREADS), TcCfg<256>::kSmemBytes, st, A.mh, A.ml, B.mh, B.ml, *mc32, *mhi,
             *mlo, M, N, K, ep);
  else
    launch_k(p, cn_gemm_tc_kernel<64>, grid, dim3(TC_THREADS), TcCfg<64>::kSmemBytes, st, A.mh, A.ml, B.mh, B.ml, *mc32, *mhi,
             *mlo, M, N, K, ep);
}
TcOut out32(float* c, int ldc) { TcOut o; o.c32 = c; o.ldc = ldc; return o; }
TcOut out16(const TcMat& t) { TcOut o; o.oh = t.hi; o.ol = t.lo; o.ldh = t.pitch; return o; }
TcOut out_both(float* c, int ldc, const TcMat& t) { TcOut o = out16(t); o.c32 = c; o.ldc = ldc; return o; }

int tc_set_attrs() {
  cudaError_t e = cudaFuncSetAttribute(cn_gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       TcCfg<256>::kSmemBytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(cn_gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<64>::kSmemBytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(cn_qkv_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, QA_SMEM_BYTES);
  if (e != cudaSuccess) return cn_set_error("cudaFuncSetAttribute(tc): %s", cudaGetErrorString(e));
  return 0;
}

void gemm(cn_policy* p, cudaStream_t st, const float* A, int lda, const float* W, int ldw, const float* bias,
          float* C, int ldc, int M, int N, int K, int act, int act_lo = 0, int act_hi = 1 << 30,
          const int* m_ptr = nullptr, __half* oh = nullptr, __half* ol = nullptr) {
  dim3 grid((N + CN_GEMM_BN - 1) / CN_GEMM_BN, (M + CN_GEMM_BM - 1) / CN_GEMM_BM);
  launch_k(p, cn_gemm_f32_kernel, grid, dim3(256), 0, st, A, lda, W, ldw, bias, C, ldc, M, N, K, act, act_lo, act_hi, m_ptr, oh, ol);
}

const char* kStageNames[] = {"pack_inputs", "embed1_gemm", "embed2_gemm", "qkv_gemm", "hh_attention",
                             "outproj_spatial_gemm", "robot_branch_join", "hr_attention", "gru", "actor_critic_heads"};
const int kNumStages = sizeof(kStageNames) / sizeof(kStageNames[0]);

inline void mark(cn_policy* p, cudaStream_t st, int i) {
  p->cur_stage = i < kNumStages ? kStageNames[i] : "end";
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
  p->cfg = *cfg;
  p->N = cfg->num_envs; p->H = cfg->human_num; p->Win = cfg->input_size; p->M = p->N * p->H;
  p->launches = 0; p->finalized = false; p->profile = false; p->launch_error = false;
  p->num_sms = 148;
  cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, cfg->device);
  {
    const char* pd = getenv("CN_PDL");
    p->pdl = !(pd && pd[0] == '0');
    const char* qc = getenv("CN_QKV_CHUNKS");
    p->qkv_chunks = (qc && qc[0] == '2') ? 2 : 1;
    const char* ar = getenv("CN_ATTN_R");
    const int attn_r = ar ? atoi(ar) : 1;
    p->attn_kernel = attn_r == 1 ? cn_hh_attention_kernel<1, 4, 4> : (attn_r == 4 ? cn_hh_attention_kernel<4, 2, 2> : cn_hh_attention_kernel<2, 4, 4>);
    p->attn_warps = attn_r == 4 ? 2 : 4;
    const char* fq = getenv("CN_FUSE_QKV");
    // opt-in (CN_FUSE_QKV=1): parity green, but the in-epilogue attention is bound by the SM's shared-memory bandwidth,
    // which it shares with the operand fetch of the MMAs -- 0.109-0.129 ms against 0.073 + 0.064 ms of the two-kernel path
    // in stage terms and no gain per rollout step (profiles/r2_qkv_attn_fused.md)
    p->fuse_qkv = cfg->gemm_mode == 1 && p->qkv_chunks == 1 && (fq && fq[0] == '1') && p->N <= QA_MAX_ENVS && p->H <= 128;
  }
  cudaEventCreateWithFlags(&p->ev_tiles, cudaEventDisableTiming);
  cudaStreamCreateWithFlags(&p->st2, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&p->st3, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&p->ev_fork3, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&p->ev_join3, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&p->ev_fork2, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&p->ev_join2, cudaEventDisableTiming);
  const size_t M = (size_t)p->M, N = (size_t)p->N;
  const int Mi = p->M, Ni = p->N;
  int rc = 0;
#define WS(name, count) if (!rc) rc = palloc(p, &p->name, (count))
  {
    float* q = nullptr;
    if (!rc) rc = palloc(p, &q, N + 2);
    p->row_start = reinterpret_cast<int*>(q);
    if (!rc) rc = palloc(p, &q, 4);
    p->mc = reinterpret_cast<int*>(q);
    if (!rc) rc = palloc(p, &q, M + 1);
    p->row_env = reinterpret_cast<int*>(q);
    if (!rc) rc = palloc(p, &q, 2 * N + 4);
    p->tile_tab = reinterpret_cast<int*>(q);
  }
  WS(x16, M * 16); WS(e1, M * 128); WS(e2, M * 512); WS(qkv, M * 1536); WS(ao, M * 512); WS(sout, M * 256);
  WS(xr, N * 16); WS(rs, N * 256); WS(t1, N * 128); WS(u, N * 256); WS(wv, N * 256); WS(h0, N * 128);
  WS(gi, N * 384); WS(gh, N * 384); WS(outb, N * 256); WS(ac1, N * 512); WS(a2, N * 256); WS(c2, N * 256);
#undef WS
  if (!rc && cfg->gemm_mode == 1) {
    // A operands: box rows = 128 (TC_BM)
    rc = tc_alloc(p, p->tE1, Mi, 128, TC_BM);
    if (!rc) rc = tc_alloc(p, p->tE2, Mi, 512, TC_BM);
    if (!rc) rc = tc_alloc(p, p->tAo, Mi, 512, TC_BM);
    if (!rc) rc = tc_alloc(p, p->tRs, Ni, 256, TC_BM);
    if (!rc) rc = tc_alloc(p, p->tT1, Ni, 128, TC_BM);                 // [enc | te] then [enc | emb]
    if (!rc) rc = tc_view(p->tTe, p->tT1, 64, Ni, 64, TC_BM);           // te = columns 64..127
    if (!rc) rc = tc_alloc(p, p->tWv, Ni, 256, TC_BM);
    if (!rc) rc = tc_alloc(p, p->tH0, Ni, 128, TC_BM);
    if (!rc) rc = tc_alloc(p, p->tH1, Ni, 128, TC_BM);
    if (!rc) rc = tc_alloc(p, p->tOut, Ni, 256, TC_BM);
    if (!rc) rc = tc_alloc(p, p->tAc1, Ni, 512, TC_BM);
    if (!rc) rc = tc_view(p->tA1, p->tAc1, 0, Ni, 256, TC_BM);          // actor.0 half
    if (!rc) rc = tc_view(p->tC1, p->tAc1, 256, Ni, 256, TC_BM);        // critic.0 half
    if (!rc) rc = tc_set_attrs();
  }
  if (rc) { cn_policy_destroy(p); return rc; }
  p->ws_allocs = p->allocs.size();
  *out = p;
  return 0;
}

int cn_policy_destroy(cn_policy* p) {
  if (!p) return 0;
  cudaSetDevice(p->cfg.device);
  for (void* q : p->allocs) cudaFree(q);
  for (auto* m : p->omaps) delete m;
  for (auto& e : p->ev) cudaEventDestroy(e);
  if (p->st2) {
    cudaStreamDestroy(p->st2);
    cudaStreamDestroy(p->st3); cudaEventDestroy(p->ev_fork3); cudaEventDestroy(p->ev_join3); cudaEventDestroy(p->ev_tiles);
    cudaEventDestroy(p->ev_fork); cudaEventDestroy(p->ev_join); cudaEventDestroy(p->ev_fork2); cudaEventDestroy(p->ev_join2);
  }
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
#undef UP
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
  if (!rc) rc = palloc(p, &p->Woac, (size_t)512 * 128);
  if (!rc) rc = palloc(p, &p->boac, 512);
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
  // Woac = [actor.0 ; critic.0] (512x256) @ output_linear (256x128);  boac = [actor.0 ; critic.0] @ bo + bac1
  // (output_linear has no activation and feeds only the two MLPs, selfAttn_srnn_temp_node.py:438-447)
  cn_fold_mm_kernel<<<dim3(1, 512), 128, 0, st>>>(p->Wac1, p->Wo, p->Woac, 512, 128, 256);
  cn_fold_mv_kernel<<<4, 128, 0, st>>>(p->Wac1, p->bo, p->bac1, p->boac, 512, 256);
  if (p->cfg.gemm_mode == 1) {
    // fp16 (hi, lo) split of the tensor-core weights, pre-scaled by 2^6 (exact) so lo stays normal.
    // B-tile rows: 256 for the per-human layers (large M), 64 for the per-environment layers.
    struct { float* src; TcMat* t; int rows, k, bn; } tw[13] = {
        {p->Woac, &p->tWoac, 512, 128, 64},
        {p->W2, &p->tW2, 512, 128, 256},    {p->Wqkv, &p->tWqkv, 1536, 512, 256}, {p->Wos, &p->tWos, 256, 512, 256},
        {p->Wet, &p->tWet, 128, 256, 64},   {p->WsT, &p->tWsT, 256, 64, 64},      {p->Wa, &p->tWa, 64, 256, 64},
        {p->Wih, &p->tWih, 384, 128, 64},   {p->Whh, &p->tWhh, 384, 128, 64},     {p->Wo, &p->tWo, 256, 128, 64},
        {p->Wac1, &p->tWac1, 512, 256, 64}, {p->Wa2, &p->tWa2, 256, 256, 64},     {p->Wc2, &p->tWc2, 256, 256, 64}};
    for (auto& t : tw) {
      rc = tc_alloc(p, *t.t, t.rows, t.k, t.bn);
      if (rc) return rc;
      split16(p, st, t.src, 64.0f, t.t->hi, t.t->lo, (size_t)t.rows * t.k);
    }
    if (p->fuse_qkv) {
      // head-major copy of the folded QKV projection for the fused kernel: row h * 192 + s * 64 + d <- row s * 512 + h * 64 + d
      float* wh = nullptr;
      rc = palloc(p, &wh, (size_t)1536 * 512);
      if (!rc) rc = palloc(p, &p->bqkvH, 1536);
      if (!rc) rc = tc_alloc(p, p->tWqkvH, 1536, 512, QA_BN);
      if (!rc) rc = make_map(&p->qa_bh, p->tWqkvH.hi, 1536, 512, QA_BN, 512, QA_BK);
      if (!rc) rc = make_map(&p->qa_bl, p->tWqkvH.lo, 1536, 512, QA_BN, 512, QA_BK);
      if (!rc) rc = make_map(&p->qa_ah, p->tE2.hi, p->M, 512, TC_BM, 512, QA_BK);
      if (!rc) rc = make_map(&p->qa_al, p->tE2.lo, p->M, 512, TC_BM, 512, QA_BK);
      if (rc) return rc;
      cn_head_major_kernel<<<1536, 128, 0, st>>>(p->Wqkv, p->bqkv, wh, p->bqkvH);
      split16(p, st, wh, 64.0f, p->tWqkvH.hi, p->tWqkvH.lo, (size_t)1536 * 512);
    }
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
  CnDeviceGuard guard(p->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  const int N = p->N, H = p->H, M = p->M;
  const bool tcm = p->cfg.gemm_mode == 1;
  const int* mc = p->mc;
  const int ALL = 1 << 30;
  mark(p, st, 0);
  // 0. compaction offsets, pack / pad inputs, h0 = h * mask
  {
    launch_k(p, cn_row_offsets_kernel, dim3(1), dim3(1024), 0, st, d->detected_human_num, N, H, p->row_start, p->mc);
    const int total = (!tcm && M * 16 > N * 128) ? M * 16 : N * 128;
    launch_k(p, cn_pack_inputs_kernel, dim3((total + 255) / 256), dim3(256), 0, st, d->spatial_edges, p->Win, H, N, p->row_start, p->row_env,
                                                               tcm ? nullptr : p->x16,
                                                               d->temporal_edges, d->robot_node, d->h_in, d->masks, p->xr,
                                                               p->h0, tcm ? p->tH0.hi : nullptr, tcm ? p->tH0.lo : nullptr);
  }
  // fork: the robot branch (rs, [enc|te], u) and gh only depend on the packed inputs
  cudaStream_t s2 = p->st2;
  cudaEventRecord(p->ev_fork, st);
  cudaStreamWaitEvent(s2, p->ev_fork, 0);
  if (p->fuse_qkv) {
    cn_qkv_tiles_kernel<<<1, 1024, 0, s2>>>(p->row_start, N, p->tile_tab);
    p->launches += 1;
    cudaEventRecord(p->ev_tiles, s2);
  }
  if (tcm) {
    gemm(p, s2, p->xr, 16, p->Wr, 16, p->br, nullptr, 256, N, 256, 16, CN_ACT_RELU, 0, ALL, nullptr, p->tRs.hi, p->tRs.lo);
    gemm_tc(p, s2, p->tRs, p->tWet, N, 128, 256, 64, p->bet, CN_ACT_RELU, out_both(p->t1, 128, p->tT1), nullptr, 0, 64);
    gemm_tc(p, s2, p->tTe, p->tWsT, N, 256, 64, 64, nullptr, CN_ACT_NONE, out32(p->u, 256));
    gemm_tc(p, s2, p->tH0, p->tWhh, N, 384, 128, 64, p->bhh, CN_ACT_NONE, out32(p->gh, 384));
  } else {
    gemm(p, s2, p->xr, 16, p->Wr, 16, p->br, p->rs, 256, N, 256, 16, CN_ACT_RELU);
    gemm(p, s2, p->rs, 256, p->Wet, 256, p->bet, p->t1, 128, N, 128, 256, CN_ACT_RELU, 0, 64);   // [enc | te]
    gemm(p, s2, p->t1 + 64, 128, p->WsT, 64, nullptr, p->u, 256, N, 256, 64, CN_ACT_NONE);        // u = W_s^T te
    gemm(p, s2, p->h0, 128, p->Whh, 128, p->bhh, p->gh, 384, N, 384, 128, CN_ACT_NONE);
  }
  cudaEventRecord(p->ev_join, s2);
  // 1. human-human branch over the Mc = sum_e n_e valid rows (device-side count p->mc)
  mark(p, st, 1);
  if (tcm) {
    launch_k(p, cn_embed1_kernel, dim3(p->num_sms * 6), dim3(256), 0, st, d->spatial_edges, p->Win, H, p->row_start, p->row_env, p->mc, p->W1, p->b1,
                                                     p->tE1.hi, p->tE1.lo);
  } else gemm(p, st, p->x16, 16, p->W1, 16, p->b1, p->e1, 128, M, 128, 16, CN_ACT_RELU, 0, ALL, mc);
  mark(p, st, 2);
  if (tcm) gemm_tc(p, st, p->tE1, p->tW2, M, 512, 128, 256, p->b2, CN_ACT_RELU, out16(p->tE2), mc);
  else gemm(p, st, p->e1, 128, p->W2, 128, p->b2, p->e2, 512, M, 512, 128, CN_ACT_RELU, 0, ALL, mc);
  mark(p, st, 3);
  if (tcm) {
    // Optional experiment (CN_QKV_CHUNKS=2): QKV projection + attention in two row chunks split at an environment
    // boundary so that chunk 0's attention (side stream) overlaps chunk 1's GEMM.  Measured on B200: no gain
    // (0.516 vs 0.505 ms/step) -- the attention is load-latency bound, not L2-capacity bound -- so it is off.
    const int* mid = p->row_start + N / 2;
    __half* ah = p->tAo.hi;
    __half* al = p->tAo.lo;
    if (p->fuse_qkv) {
      // one kernel: projection tile (128 rows of whole environments x one head's Q | K | V) -> attention -> tAo
      cudaStreamWaitEvent(st, p->ev_tiles, 0);                 // tile table from the side stream
      launch_k(p, cn_qkv_attn_kernel, dim3(p->num_sms), dim3(QA_THREADS), QA_SMEM_BYTES, st, p->qa_ah, p->qa_al, p->qa_bh,
               p->qa_bl, p->bqkvH, 1.0f / 64.0f, p->tile_tab, p->row_start, p->row_env, ah, al,
               getenv("CN_QA_DBG") ? atoi(getenv("CN_QA_DBG")) : 0);
      mark(p, st, 4);
    } else if (p->qkv_chunks == 1) {
      gemm_tc(p, st, p->tE2, p->tWqkv, M, 1536, 512, 256, p->bqkv, CN_ACT_NONE, out32(p->qkv, 1536), mc);
      mark(p, st, 4);
      launch_k(p, p->attn_kernel, dim3(p->num_sms * 64 / p->attn_warps), dim3(p->attn_warps * 32), 0, st, p->qkv, p->row_start, p->row_env, mc, nullptr,
                                                                             nullptr, ah, al);
    } else {
    gemm_tc(p, st, p->tE2, p->tWqkv, M, 1536, 512, 256, p->bqkv, CN_ACT_NONE, out32(p->qkv, 1536), mid);
    cudaEventRecord(p->ev_fork3, st);
    cudaStreamWaitEvent(p->st3, p->ev_fork3, 0);
    launch_k(p, p->attn_kernel, dim3(p->num_sms * 32 / p->attn_warps), dim3(p->attn_warps * 32), 0, p->st3, p->qkv, p->row_start, p->row_env, mid, nullptr,
                                                                              nullptr, ah, al);
    cudaEventRecord(p->ev_join3, p->st3);
    gemm_tc(p, st, p->tE2, p->tWqkv, M, 1536, 512, 256, p->bqkv, CN_ACT_NONE, out32(p->qkv, 1536), mc, 0, 1 << 30, mid);
    mark(p, st, 4);
    launch_k(p, p->attn_kernel, dim3(p->num_sms * 64 / p->attn_warps), dim3(p->attn_warps * 32), 0, st, p->qkv, p->row_start, p->row_env, mc, mid, nullptr,
                                                                           ah, al);
    cudaStreamWaitEvent(st, p->ev_join3, 0);
    }
  } else {
    gemm(p, st, p->e2, 512, p->Wqkv, 512, p->bqkv, p->qkv, 1536, M, 1536, 512, CN_ACT_NONE, 0, ALL, mc);
    mark(p, st, 4);
    launch_k(p, p->attn_kernel, dim3(p->num_sms * 64 / p->attn_warps), dim3(p->attn_warps * 32), 0, st, p->qkv, p->row_start, p->row_env, p->mc, nullptr,
                                                                           p->ao, nullptr, nullptr);
  }
  mark(p, st, 5);
  if (tcm) gemm_tc(p, st, p->tAo, p->tWos, M, 256, 512, 256, p->bos, CN_ACT_RELU, out32(p->sout, 256), mc);
  else gemm(p, st, p->ao, 512, p->Wos, 512, p->bos, p->sout, 256, M, 256, 512, CN_ACT_RELU, 0, ALL, mc);
  // 2. join the robot branch
  mark(p, st, 6);
  cudaStreamWaitEvent(st, p->ev_join, 0);
  mark(p, st, 7);
  launch_k(p, cn_hr_attention_kernel, dim3((N + 3) / 4), dim3(128), 0, st, p->sout, p->u, p->t1, 128, 64, p->bs, p->row_start, N, H, p->wv,
                                                      tcm ? p->tWv.hi : nullptr, tcm ? p->tWv.lo : nullptr);
  // 3. GRU: emb overwrites the te half of t1 -> t1 = [enc | emb] = GRU input (gh came from the side stream)
  mark(p, st, 8);
  if (tcm) {
    TcOut o; o.oh = p->tT1.hi + 64; o.ol = p->tT1.lo + 64; o.ldh = 128;
    gemm_tc(p, st, p->tWv, p->tWa, N, 64, 256, 64, p->ba, CN_ACT_RELU, o);
    gemm_tc(p, st, p->tT1, p->tWih, N, 384, 128, 64, p->bih, CN_ACT_NONE, out32(p->gi, 384));
  } else {
    gemm(p, st, p->wv, 256, p->Wa, 256, p->ba, p->t1 + 64, 128, N, 64, 256, CN_ACT_RELU);
    gemm(p, st, p->t1, 128, p->Wih, 128, p->bih, p->gi, 384, N, 384, 128, CN_ACT_NONE);
  }
  launch_k(p, cn_gru_gate_kernel, dim3((N * 128 + 255) / 256), dim3(256), 0, st, p->gi, p->gh, p->h0, N, d->h_out, tcm ? p->tH1.hi : nullptr,
                                                            tcm ? p->tH1.lo : nullptr);
  // 4. output_linear, actor / critic MLPs (critic.2 on the side stream), heads
  mark(p, st, 9);
  if (tcm) {
    gemm_tc(p, st, p->tH1, p->tWoac, N, 512, 128, 64, p->boac, CN_ACT_TANH, out16(p->tAc1));     // [actor.0 | critic.0]
    cudaEventRecord(p->ev_fork2, st);
    cudaStreamWaitEvent(s2, p->ev_fork2, 0);
    gemm_tc(p, s2, p->tC1, p->tWc2, N, 256, 256, 64, p->bc2, CN_ACT_TANH, out32(p->c2, 256));
    cudaEventRecord(p->ev_join2, s2);
    gemm_tc(p, st, p->tA1, p->tWa2, N, 256, 256, 64, p->ba2, CN_ACT_TANH, out32(p->a2, 256));
  } else {
    gemm(p, st, d->h_out, 128, p->Woac, 128, p->boac, p->ac1, 512, N, 512, 128, CN_ACT_TANH);
    cudaEventRecord(p->ev_fork2, st);
    cudaStreamWaitEvent(s2, p->ev_fork2, 0);
    gemm(p, s2, p->ac1 + 256, 512, p->Wc2, 256, p->bc2, p->c2, 256, N, 256, 256, CN_ACT_TANH);
    cudaEventRecord(p->ev_join2, s2);
    gemm(p, st, p->ac1, 512, p->Wa2, 256, p->ba2, p->a2, 256, N, 256, 256, CN_ACT_TANH);
  }
  cudaStreamWaitEvent(st, p->ev_join2, 0);
  launch_k(p, cn_heads_kernel, dim3((N + 3) / 4), dim3(128), 0, st, p->a2, 256, p->c2, 256, p->wv_, p->bv, p->Wm, p->bm, p->logstd, d->noise, N,
                                               d->value, d->action, d->log_prob, d->action_mean);
  mark(p, st, kNumStages);
  cudaError_t err = cudaGetLastError();
  if (p->launch_error) { p->launch_error = false; return 1; }      // cn_last_error names the stage
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
// tcgen05 3xFP16 kernel with B-tile rows `bn` (256 or 64), fp32 device pointers in/out.
int cn_internal_gemm_tc(const float* dA, const float* dW, const float* dbias, float* dC, int M, int N, int K, int act,
                        int bn) {
  if ((bn != 256 && bn != 64) || N % bn || K % TC_BK)
    return cn_set_error("cn_internal_gemm_tc: need bn in {64,256}, N %% bn == 0 and K %% 64 == 0");
  cn_policy tmp;
  tmp.launches = 0;
  tmp.st2 = nullptr; tmp.st3 = nullptr;
  tmp.num_sms = 148; tmp.qkv_chunks = 1; tmp.launch_error = false; tmp.pdl = false;
  cudaDeviceGetAttribute(&tmp.num_sms, cudaDevAttrMultiProcessorCount, 0);
  TcMat A, B;
  int rc = tc_alloc(&tmp, A, M, K, TC_BM);
  if (!rc) rc = tc_alloc(&tmp, B, N, K, bn);
  if (!rc) rc = tc_set_attrs();
  if (!rc) {
    split16(&tmp, 0, dA, 1.0f, A.hi, A.lo, (size_t)M * K);
    split16(&tmp, 0, dW, 64.0f, B.hi, B.lo, (size_t)N * K);
    gemm_tc(&tmp, 0, A, B, M, N, K, bn, dbias, act, out32(dC, N));
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) rc = cn_set_error("cn_internal_gemm_tc: %s", cudaGetErrorString(err));
    else if (tmp.launch_error) rc = 1;
  }
  for (void* q : tmp.allocs) cudaFree(q);
  for (auto* m : tmp.omaps) delete m;
  return rc;
}

}  // extern "C"

#include "cn_gst_tc.cuh"
#include "cn_update.cuh"
