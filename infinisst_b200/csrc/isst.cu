// infinisst_b200: context, weight ingestion, stream/KV management and the per-chunk step
// (encoder -> adapter -> chunk-prefill -> greedy decode) behind the C-ABI of include/infinisst_b200.h.
//
// Reference call stack being replaced:
//   InfiniSST.policy (agents/infinisst.py:270-394) -> model.generate -> SpeechLlamaModel.forward
//   (model/llm.py:51-126) -> encode_speech (model/speech_encoder.py:219-236) -> uni_w2v2_forward /
//   uni_mha_forward (model/patches/patch_speech_encoder.py) and llama_sdpa_attention_new_forward
//   (model/patches/patch_llm.py:231-336).
#include "../../include/infinisst_b200.h"

#include <cudaTypedefs.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <utility>
#include <vector>

#include "attention.cuh"
#include "common.cuh"
#include "decode_attention.cuh"
#include "decode_attention_group.cuh"
#include "decode_chain.cuh"
#include "gemm_tcgen05.cuh"
#include "gemm_pair.cuh"
#include "prefill_attention_tc.cuh"
#include "beam.cuh"
#include "rowops.cuh"

namespace isst {
thread_local std::string g_last_error;

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;

static int load_driver_api() {
  if (g_encode_tiled) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  ISST_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  ISST_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
  g_encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return 0;
}

template <typename T>
static int dev_alloc(T** p, size_t count) {
  *p = nullptr;
  if (count == 0) return 0;
  ISST_CUDA(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  return 0;
}

struct Weight2D {
  bf16* ptr = nullptr;
  int rows = 0, K = 0;
  CUtensorMap map;          // 128-row boxes
  CUtensorMap map64;        // 64-row boxes (gate / up halves of a CTA-pair tile)
};

static int make_weight_map(Weight2D& w) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(w.K), static_cast<cuuint64_t>(w.rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(w.K) * 2};
  cuuint32_t box[2] = {tc::kBK, tc::kBM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(&w.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w.ptr, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ISST_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight) failed: " + std::to_string(static_cast<int>(r)));
  cuuint32_t box64[2] = {tc::kBK, 64};
  r = g_encode_tiled(&w.map64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w.ptr, dims, strides, box64, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ISST_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight, 64 rows) failed: " + std::to_string(static_cast<int>(r)));
  return 0;
}

// Activation view: element (row, kidx) of batch b lives at
//   ptr + b * batch_stride + ((row * conv_s + kidx / conv_c) * conv_c + kidx % conv_c)
// (plain [rows, K] matrices have conv_c = K, conv_s = 1).  4-D tensor map {conv_c, conv_s, row_groups, batch}.
struct ActView {
  const bf16* ptr;
  int rows;               // logical rows per batch (M_tok)
  int K;
  int conv_c, conv_s;
  int row_groups;         // extent of the row-group dimension (>= rows + (k-1)/s)
  long long batch_stride; // elements
  int batch;
};
static ActView plain_view(const bf16* p, int rows, int K) { return ActView{p, rows, K, K, 1, rows, 0, 1}; }

static int make_act_map(CUtensorMap* map, const ActView& v, int box_rows) {
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(v.conv_c), static_cast<cuuint64_t>(v.conv_s),
                        static_cast<cuuint64_t>(v.row_groups), static_cast<cuuint64_t>(v.batch)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(v.conv_c) * 2,
                           static_cast<cuuint64_t>(v.conv_c) * v.conv_s * 2,
                           static_cast<cuuint64_t>(v.batch > 1 ? v.batch_stride : static_cast<long long>(v.conv_c) * v.conv_s * v.row_groups) * 2};
  cuuint32_t box[4] = {tc::kBK, 1, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(v.ptr), dims, strides,
                              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ISST_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(act) failed: " + std::to_string(static_cast<int>(r)));
  return 0;
}

// Activation tensor maps are pure functions of (pointer, geometry): encode once, reuse on every launch.
struct ActMapKey {
  const void* ptr; int rows, K, conv_c, conv_s, row_groups, batch, box_rows; long long batch_stride;
  bool operator<(const ActMapKey& o) const {
    return std::tie(ptr, rows, K, conv_c, conv_s, row_groups, batch, box_rows, batch_stride) <
           std::tie(o.ptr, o.rows, o.K, o.conv_c, o.conv_s, o.row_groups, o.batch, o.box_rows, o.batch_stride);
  }
};

struct EncLayerW {
  float *ln1_w = nullptr, *ln1_b = nullptr, *bqkv = nullptr, *bo = nullptr, *ln2_w = nullptr, *ln2_b = nullptr,
        *b1 = nullptr, *b2 = nullptr;
  Weight2D wqkv, wo, w1, w2;
};
struct LlmLayerW {
  float *rms1 = nullptr, *rms2 = nullptr;
  Weight2D wqkv, wo, wgu, wd;
};
struct ConvW {
  Weight2D w;                       // conv j >= 1 and adapter: [C_out, k*C_in], K index = tap * C_in + c
  float *w0_t = nullptr;            // conv0 only: [k][C] fp32
  float *bias = nullptr, *ln_w = nullptr, *ln_b = nullptr;
};

// ---- per-kernel-class device timing (bench.py roofline leg): CUDA events around every launch ----
enum ProfCat { P_GEMM_TENSOR = 0, P_GEMM_STREAM, P_ATTN_ENC, P_ATTN_PREFILL, P_ATTN_DECODE, P_NORM, P_CONV0, P_APPEND,
               P_EMBED, P_SELECT, P_NCAT };
static const char* kProfNames[P_NCAT] = {"gemm_tensor", "gemm_stream", "attn_encoder", "attn_prefill", "attn_decode",
                                         "norm", "conv0", "kv_append", "embed_splice", "greedy_select"};
struct ProfRec { int cat; cudaEvent_t e0, e1; double flops, bytes; };
struct Prof {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  std::vector<ProfRec> recs;
  double ms[P_NCAT] = {0}, flops[P_NCAT] = {0}, bytes[P_NCAT] = {0};
  long long n[P_NCAT] = {0};
  cudaEvent_t get() {
    if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[used++];
  }
};

struct StreamHost {
  bool open = false;
  int enc_prefix = 0;               // W2V2RoPECache.n_steps
  int kv_len = 0, sys_len = 0, ring_start = 0;
  long long evicted = 0;            // ring tokens dropped so far: absolute index of a ring token = logical index + evicted
  bool prefilled = false;
  std::vector<int> pages;           // page table entries (sys pages first, then ring pages)
};

}  // namespace isst

using namespace isst;

struct isst_ctx {
  isst_config cfg;
  int device = 0;
  int sm_count = 148;
  bool finalized = false;
  // per-context tuning / test options (isst_debug_option); nothing on the hot path reads the environment
  bool pdl = true;          // "pdl" = 0: plain stream order instead of programmatic dependent launch
  int opt_dec_splits = 0;   // "decode_splits" > 0: fixed key-split count of decode attention (micro-benchmarks)
  bool opt_chain = true;    // "decode_chain" = 0: one kernel per operator instead of the fused decode-layer chain
  bool tap_llm_layers = false;   // "tap_llm_layers" = 1: per-layer residual taps of the LLM prefill (operator path only)
  int opt_pa_l2_ahead = 1;           // "prefill_l2_ahead": K/V tiles the prefill attention asks into L2 ahead of its ring
  bool opt_pair = true;              // "gemm_pair" = 0: one CTA per tile instead of CTA pairs (cta_group::2) above 128 rows (A/B)
  bool opt_defer_as_chain = false;   // "defer_splits_as_chain" (tests): the operator-per-kernel path cuts K like the chain does
  int* chain_ctr = nullptr;                  // folded-norm arrival counters of decode_chain_kernel (two halves, see chain::Params)
  int chain_parity = 0;
  float* chain_sq = nullptr;                 // [2][kMaxTiles][256] squared row sums per feature tile (folded RMSNorm)
  bool opt_fold = true;                      // "chain_fold" = 0: RMSNorm row phases between the decode GEMMs (the reference module's two
                                             //   bf16 roundings; bit-identical to the operator path) instead of norms folded into the GEMMs
  unsigned long long* chain_bar = nullptr;   // grid-barrier counters of decode_chain_kernel, one per phase index (monotonic)
  unsigned long long chain_base[chain::kMaxPhases] = {0};   // their values once every launch issued so far has completed
  std::set<const void*> smem_attr_done;      // kernels whose dynamic shared-memory limit was raised on this device
  std::map<std::string, long long> paths;    // launches per kernel variant (isst_path_count; parity tests assert on them)
  unsigned long long* gemm_dbg = nullptr;   // optional phase stamps of the last stream-K launch
  int64_t launches = 0;
  bool debug = false;
  Prof prof;
  std::map<std::string, std::pair<void*, size_t>> taps;   // name -> (device buffer, bytes)
  std::map<std::string, bool> loaded;
  std::map<ActMapKey, CUtensorMap> act_maps;

  // geometry
  int C = 0, n_tail = 0, rf = 0, total_stride = 0, frames_max = 0, samples_max = 0, enc_cap = 0;
  int speech_per_block = 0;   // LLM tokens per block of frames
  int pages_per_stream = 0;

  // weights
  std::vector<ConvW> conv, adapter;
  float *feat_ln_w = nullptr, *feat_ln_b = nullptr, *post_b = nullptr, *enc_ln_w = nullptr, *enc_ln_b = nullptr,
        *proj_b = nullptr, *final_norm = nullptr;
  Weight2D post_proj, proj, lm_head;
  std::vector<EncLayerW> enc;
  std::vector<LlmLayerW> llm;
  bf16* embed = nullptr;
  float *enc_inv_freq = nullptr, *llm_inv_freq = nullptr;          // RoPE frequencies (fp32, like the reference modules)
  float* enc_xpos_base = nullptr;   // --xpos 1: (2d + 0.4 hd) / (1.4 hd) per rotary pair
  bf16* enc_kx = nullptr;           // --xpos 1: one layer of window keys with the per-call xPos scale applied
  float2 *enc_rope_tab = nullptr, *llm_rope_ring = nullptr, *llm_rope_sys = nullptr;   // (cos, sin) of the new rows
  bf16* lq_sys = nullptr;                                          // q rotated for the pinned prefix keys
  int* d_evicted = nullptr;
  void* staging = nullptr;
  size_t staging_bytes = 0;

  // stream state
  std::vector<StreamHost> streams;
  std::vector<int> free_pages;
  float* tail = nullptr;
  bf16 *enc_k = nullptr, *enc_v = nullptr;
  int *d_enc_prefix = nullptr, *d_page_table = nullptr, *d_kv_len = nullptr, *d_sys_len = nullptr,
      *d_ring_start = nullptr;
  bf16* kv_pool = nullptr;
  CUtensorMap kv_map;        // the whole LLM KV pool as [rows][head_dim]: boxes of one page x 64 dims (prefill attention)
  size_t kv_layer_elems = 0, enc_layer_elems = 0;

  // workspaces
  float* d_pcm = nullptr;
  bf16 *conv_a = nullptr, *conv_b = nullptr;
  bf16 *ex = nullptr, *eh = nullptr, *eqkv = nullptr, *eattn = nullptr, *effn = nullptr, *ead0 = nullptr,
       *ead1 = nullptr, *speech = nullptr;
  int speech_rows_per_stream = 0;
  bf16 *lx = nullptr, *lh = nullptr, *lqkv = nullptr, *lattn = nullptr, *lgu = nullptr, *llast = nullptr;
  float* logits = nullptr;
  float* logits_all = nullptr;      // isst_forward_all: [rows][vocab], grown on demand
  size_t logits_all_rows = 0;
  SelectWs sel_ws{nullptr, nullptr, nullptr};
  float *part_o = nullptr, *part_ml = nullptr;
  int decode_splits = 1;
  float* gemm_ws = nullptr;
  size_t gemm_ws_floats = 0;
  int* gemm_counters = nullptr;
  int n_counters = 0;
  int gemm_parity = 0;          // alternates per stream-K launch (counter halves)
  float* defer_ws = nullptr;    // fp32 split partials of a deferred-reduction GEMM [splits][tokens][features]
  size_t defer_ws_floats = 0;
  int last_defer_splits = 0;    // splits the last gemm() call left in defer_ws (0: output complete)
  // batch metadata (device) + pinned host staging
  int* d_meta = nullptr;
  int* h_meta = nullptr;
  cudaEvent_t ev_active = nullptr;
  int* h_tables = nullptr;   // pinned staging of the per-batch page tables / lengths
  size_t meta_ints = 0;
  int* d_step_logits_dummy = nullptr;
  float* beam_ws = nullptr;     // beam search selection workspace (BeamSel partials / candidates / results)
  int* beam_count = nullptr;    // [max_batch] arrival counters
};

namespace isst {

// Hot-path launcher: cudaLaunchKernelEx with the programmatic-stream-serialization attribute (PDL), so the next
// kernel's prologue overlaps this kernel's tail; every kernel launched this way calls pdl_wait() before it
// touches data of its predecessors.  ISST_PDL=0 falls back to plain stream order (A/B runs).
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(isst_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (ctx->pdl && !ctx->prof.on) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}

#define LAUNCH_CHECK(ctx)                                                                   \
  do {                                                                                      \
    (ctx)->launches++;                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess)                                                                  \
      return set_error(std::string("kernel launch failed: ") + cudaGetErrorString(_e) +     \
                       " at " + __FILE__ + ":" + std::to_string(__LINE__));                 \
  } while (0)

// Scope that brackets the launches inside it with two events when profiling is on.
struct ProfScope {
  isst_ctx* ctx; cudaStream_t st; ProfRec r; bool live;
  ProfScope(isst_ctx* c, cudaStream_t s, int cat, double flops, double bytes) : ctx(c), st(s), live(c->prof.on) {
    if (!live) return;
    r.cat = cat; r.flops = flops; r.bytes = bytes;
    r.e0 = ctx->prof.get(); r.e1 = ctx->prof.get();
    cudaEventRecord(r.e0, st);
  }
  ~ProfScope() {
    if (!live) return;
    cudaEventRecord(r.e1, st);
    ctx->prof.recs.push_back(r);
  }
};
// Folds finished records into the per-class totals (the caller has synchronised the stream).
static void prof_flush(isst_ctx* ctx) {
  Prof& p = ctx->prof;
  for (const ProfRec& r : p.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      p.ms[r.cat] += ms; p.flops[r.cat] += r.flops; p.bytes[r.cat] += r.bytes; p.n[r.cat]++;
    }
  }
  p.recs.clear();
  p.used = 0;
  cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// GEMM dispatch
// ------------------------------------------------------------------------------------------------
struct Epilogue {
  const float* bias = nullptr;
  int act = 0;
  const bf16* resid = nullptr;
  long long ldr = 0;
  long long resid_batch_stride = 0;
  int out_f32 = 0;
  int dual = 0;   // 1: weights hold [gate; up] stacked, rows N_out and dual_off + N_out
  bool defer = false;   // weight-streaming mode only: leave fp32 split partials in ctx->defer_ws for the consumer row kernel
  int dual_off = 0;
};

// k-splits of a weight-streaming GEMM in the fused decode chain: units of TWO 128-row weight tiles (they share one
// activation tile), cut into S k-ranges so that units x S fills the SMs.
static int chain_splits(int sm_count, int n_out, int num_kb, bool dual) {
  const int tiles = ceil_div(n_out, tc::kBM);
  const int units = dual ? tiles : ceil_div(tiles, 2);
  const long long s = std::min<long long>(std::min<long long>(sm_count / std::max(units, 1), 8), num_kb / 8);
  return s >= 2 ? static_cast<int>(s) : 1;
}

// Raises the dynamic shared-memory limit of a kernel once per context (= per device).
template <typename F>
static int ensure_smem(isst_ctx* ctx, F* kern, int bytes) {
  const void* key = reinterpret_cast<const void*>(kern);
  if (!ctx->smem_attr_done.count(key)) {
    ISST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    ctx->smem_attr_done.insert(key);
  }
  return 0;
}

static int get_act_map(isst_ctx* ctx, CUtensorMap* map, const ActView& v, int box_rows) {
  const ActMapKey key{v.ptr, v.rows, v.K, v.conv_c, v.conv_s, v.row_groups, v.batch, box_rows, v.batch_stride};
  auto it = ctx->act_maps.find(key);
  if (it == ctx->act_maps.end()) {
    CUtensorMap m;
    ISST_TRY(make_act_map(&m, v, box_rows));
    if (ctx->act_maps.size() > 4096) ctx->act_maps.clear();      // caller-owned pointers (isst_op_gemm) may churn
    it = ctx->act_maps.emplace(key, m).first;
  }
  *map = it->second;
  return 0;
}

// Work split of one persistent stream-K launch (see tc::SkParams): grid size G, data-parallel / stream-K tile
// counts, and S > 0 when the deferred split reduction is used (every tile cut into S k-ranges, one CTA each).
static void sk_plan(isst_ctx* ctx, int M_tok, int N_out, int K, int batch, int act_rows, int w_rows, bool swap, bool dual,
                    bool want_defer, int force_splits, tc::SkParams* out, long long* G_out, int* S_out) {
  tc::SkParams sk{};
  sk.tiles_tok = ceil_div(M_tok, act_rows);
  sk.tiles_feat = ceil_div(N_out, w_rows);
  sk.num_kb = ceil_div(K, tc::kBK);
  sk.tiles = static_cast<long long>(sk.tiles_tok) * sk.tiles_feat * batch;
  long long G = std::min<long long>(ctx->sm_count, sk.tiles * sk.num_kb);
  if (force_splits > 0) G = std::min<long long>(G, sk.tiles * force_splits);
  // Remainder policy: the tiles beyond the last full round-robin wave are either dealt as one more (partly
  // empty) data-parallel round, or cut stream-K style into equal unit ranges.  Stream-K pays a split reduction
  // (~6 us: park partial, fence, wait, reduce) and wins only when it shortens the critical path by more than
  // that, i.e. when a whole tile is long compared with the reduction: `gain` units saved vs `r_units`.
  const long long rem = sk.tiles % G;
  const long long r_units = (swap && !dual) ? 16 : 24;
  bool use_sk = rem > 0 && (sk.num_kb - ceil_div(static_cast<int>(rem * sk.num_kb), static_cast<int>(G))) > r_units;
  if (force_splits > 1) use_sk = rem > 0;
  sk.tiles_dp = use_sk ? sk.tiles - rem : sk.tiles;
  if (!use_sk && sk.tiles < G) G = sk.tiles;
  sk.units_sk = (sk.tiles - sk.tiles_dp) * sk.num_kb;
  sk.g_sk = static_cast<int>(std::min<long long>(G, sk.units_sk));
  *S_out = 0;
  if (want_defer) {
    // Deferred split reduction: every tile is cut into S equal k-ranges (one CTA each, tiles * S <= #SMs) and the
    // partials are summed by the consumer row kernel; falls back to the in-kernel schemes when S would be 1.
    long long S = std::min<long long>(std::min<long long>(ctx->sm_count / std::max<long long>(sk.tiles, 1), 8), sk.num_kb / 8);
    // test option: the k-ranges of the fused decode chain (its units are tile pairs), so that both paths sum the same
    // partials in the same order; the grid may then exceed the SM count (no CTA of this mode waits for another)
    if (ctx->opt_defer_as_chain && swap && batch == 1) S = chain_splits(ctx->sm_count, N_out, sk.num_kb, false);
    if (swap && batch == 1 && S >= 2 && force_splits == 0 &&
        static_cast<size_t>(S) * M_tok * N_out <= ctx->defer_ws_floats) {
      G = sk.tiles * S;
      sk.tiles_dp = 0;
      sk.units_sk = sk.tiles * sk.num_kb;
      sk.g_sk = static_cast<int>(G);
      *S_out = static_cast<int>(S);
    }
  }
  *out = sk;
  *G_out = G;
}

template <int kBN, bool kDual, bool kSwap>
static int launch_sk(isst_ctx* ctx, cudaStream_t st, const ActView& v, const Weight2D& w, const tc::GemmParams& p,
                     int force_splits) {
  using C = tc::SkCfg<kBN, kDual, kSwap>;
  auto kern = tc::gemm_sk_kernel<kBN, kDual, kSwap>;
  ISST_TRY(ensure_smem(ctx, kern, C::kSmemBytes));
  tc::SkParams sk{};
  long long G = 0;
  int S = 0;
  sk_plan(ctx, p.M_tok, p.N_out, p.K, p.batch, C::kActRows, C::kWRows, kSwap, kDual, p.part_out != nullptr, force_splits, &sk, &G, &S);
  sk.dbg = ctx->gemm_dbg;
  tc::GemmParams pp = p;
  pp.part_out = S ? ctx->defer_ws : nullptr;
  pp.part_splits = S;
  ctx->last_defer_splits = S;
  if (kSwap) ctx->paths[std::string("gemm_sk_swap") + std::to_string(kBN) + (kDual ? "_dual" : "") + (S ? "_deferred" : "")]++;
  else ctx->paths[std::string("gemm_sk_rows") + std::to_string(kBN) + (kDual ? "_dual" : "")]++;
  pp.counter_half = ctx->n_counters / 2;
  pp.counter_parity = ctx->gemm_parity;
  ctx->gemm_parity ^= 1;
  const size_t slot = static_cast<size_t>(C::kAccAll) * tc::kBM;
  ISST_CHECK(2 * static_cast<size_t>(G) * slot <= ctx->gemm_ws_floats && G <= ctx->n_counters / 2,
             "gemm: stream-K workspace too small");
  CUtensorMap amap;
  ISST_TRY(get_act_map(ctx, &amap, v, C::kActRows));
  ISST_CUDA(launch_k(ctx, kern, dim3(static_cast<unsigned>(G)), dim3(C::kThreadsTotal), C::kSmemBytes, st, amap, w.map, pp, sk));
  LAUNCH_CHECK(ctx);
  return 0;
}

// CTA-pair kernel (gemm_pair.cuh): 256 features (gate/up: 128) x tile_tok tokens per pair of SMs.
// tile_tok (any multiple of 16 up to 256) and the remainder policy come from a small cost model fitted to the phase
// stamps of the kernel (tests/gemm_bench.py --stamps): a k-block costs what the SM needs to ingest its 16 KB of weights
// and tile_tok / 2 token rows at ~43 B / clk; whole rounds of tiles (data-parallel) are compared with cutting the last
// partial round stream-K style, which evens the load out but pays the fp32 park + reduce (~12 us).
template <bool kDual>
static int launch_pair(isst_ctx* ctx, cudaStream_t st, const ActView& v, const Weight2D& w, const tc::GemmParams& p) {
  using C = tc::PairCfg<kDual>;
  auto kern = tc::gemm_pair_kernel<kDual>;
  ISST_TRY(ensure_smem(ctx, kern, C::kSmemBytes));
  const long long pairs = ctx->sm_count / 2;
  const int tiles_feat = ceil_div(p.N_out, C::kFeatTile);
  const int num_kb = ceil_div(p.K, tc::kBK);
  int tile_tok = 0;
  bool use_sk = false;
  double best = 0.0;
  for (int tt = 16; tt <= C::kCapN; tt += 16) {
    const double kb_cycles = (16384.0 + 64.0 * tt) / 43.0;
    const long long tiles = static_cast<long long>(ceil_div(p.M_tok, tt)) * tiles_feat;
    const double dp = static_cast<double>(ceil_div(static_cast<int>(tiles), static_cast<int>(pairs))) * num_kb * kb_cycles;
    const double sk = static_cast<double>(tiles) * num_kb / pairs * kb_cycles + 25000.0;
    if (tile_tok == 0 || dp < best) { best = dp; tile_tok = tt; use_sk = false; }
    if (tiles % pairs != 0 && sk < best) { best = sk; tile_tok = tt; use_sk = true; }
  }
  tc::SkParams sk{};
  sk.tiles_tok = ceil_div(p.M_tok, tile_tok);
  sk.tiles_feat = tiles_feat;
  sk.num_kb = num_kb;
  sk.tiles = static_cast<long long>(sk.tiles_tok) * tiles_feat;
  long long G = std::min<long long>(pairs, use_sk ? sk.tiles * num_kb : sk.tiles);
  sk.tiles_dp = use_sk ? sk.tiles - sk.tiles % G : sk.tiles;
  sk.units_sk = (sk.tiles - sk.tiles_dp) * num_kb;
  sk.g_sk = static_cast<int>(std::min<long long>(G, sk.units_sk));
  sk.dbg = ctx->gemm_dbg;
  tc::GemmParams pp = p;
  pp.part_out = nullptr;
  pp.part_splits = 0;
  ctx->paths[std::string("gemm_pair") + (kDual ? "_dual" : "")]++;
  pp.counter_half = ctx->n_counters / 2;
  pp.counter_parity = ctx->gemm_parity;
  ctx->gemm_parity ^= 1;
  const size_t slot = static_cast<size_t>(C::kAccAll) * tc::kBM;
  ISST_CHECK(2 * static_cast<size_t>(2 * G) * slot <= ctx->gemm_ws_floats && 2 * G <= ctx->n_counters / 2,
             "gemm: stream-K workspace too small");
  CUtensorMap amap;
  ISST_TRY(get_act_map(ctx, &amap, v, tile_tok / 2));
  ISST_CUDA(launch_k(ctx, kern, dim3(static_cast<unsigned>(2 * G)), dim3(C::kThreadsTotal), C::kSmemBytes, st, amap,
                     kDual ? w.map64 : w.map, pp, sk, tile_tok));
  LAUNCH_CHECK(ctx);
  return 0;
}

static int gemm(isst_ctx* ctx, cudaStream_t st, const ActView& v, const Weight2D& w, int n_out, void* out,
                long long ldo, long long out_batch_stride, const Epilogue& e, int force_swap = -1,
                int force_splits = 0) {
  tc::GemmParams p{};
  p.M_tok = v.rows;
  p.N_out = n_out;
  p.K = v.K;
  p.batch = v.batch;
  p.splits = 1;
  p.dual_off = e.dual_off;
  p.conv_c = v.conv_c;
  p.conv_s = v.conv_s;
  p.out = out;
  p.ldo = ldo;
  p.out_batch_stride = out_batch_stride;
  p.out_f32 = e.out_f32;
  p.bias = e.bias;
  p.resid = e.resid;
  p.ldr = e.ldr;
  p.resid_batch_stride = e.resid_batch_stride;
  p.act = e.act;
  p.ws = ctx->gemm_ws;
  p.counters = ctx->gemm_counters;
  p.part_out = e.defer ? ctx->defer_ws : nullptr;   // request; launch_sk decides (ctx->last_defer_splits)
  p.part_splits = 0;
  ctx->last_defer_splits = 0;
  ISST_CHECK(v.K == w.K, "gemm: K mismatch");
  // algorithmic work: every operand touched once (weights dominate when few tokens stream them)
  const double nw = e.dual ? 2.0 : 1.0;
  const double g_flops = 2.0 * v.rows * v.batch * static_cast<double>(n_out) * v.K * nw;
  const double act_elems = (v.conv_c == v.K) ? static_cast<double>(v.rows) * v.K
                                             : static_cast<double>(v.rows) * v.conv_s * v.conv_c;   // strided conv: input once
  const double g_bytes = nw * n_out * static_cast<double>(v.K) * 2 + act_elems * v.batch * 2 +
                         static_cast<double>(v.rows) * v.batch * n_out * (e.out_f32 ? 4 : 2) * (e.resid ? 2 : 1);
  // M <= 128 tokens: weight streaming (HBM roofline); more: tensor-pipe roofline
  ProfScope ps(ctx, st, v.rows * v.batch <= 128 ? P_GEMM_STREAM : P_GEMM_TENSOR, g_flops, g_bytes);
  ISST_CHECK(v.conv_c % tc::kBK == 0, "gemm: conv_c must be a multiple of 64");
  // weights on the 128-lane operand up to 64 token rows; measured at 128 / 256 rows (tests/gemm_bench_rows.py):
  // tokens on the 128-lane operand is as fast or faster (133 vs 137 us, 197 vs 210 us per decode layer)
  bool swap = (v.rows <= 64 && v.batch == 1);
  if (force_swap >= 0) swap = force_swap != 0;
  ISST_CHECK(!(swap && v.batch != 1), "gemm: swap mode needs batch == 1");
  // persistent stream-K kernel: tile shape by mode, CTA count = SM count whatever the tile count
  const int sbn = v.rows <= 16 ? 16 : (v.rows <= 32 ? 32 : (v.rows <= 64 ? 64 : 128));
#define ISST_SK(BN, DUAL, SWAP) return launch_sk<BN, DUAL, SWAP>(ctx, st, v, w, p, force_splits)
  if (!swap && ctx->opt_pair && v.rows > 128 && v.batch == 1 && v.conv_c == v.K && v.K >= 256 && force_splits == 0) {
    // tensor-bound regime: one tile per CTA pair (cta_group::2) - a third less operand traffic per SM and flop
    if (e.dual) return launch_pair<true>(ctx, st, v, w, p);
    return launch_pair<false>(ctx, st, v, w, p);
  }
  if (!swap) {
    // one CTA per 128-token tile: strided-conv views (batch > 1), and the A/B arm of the CTA-pair kernel
    if (e.dual) ISST_SK(128, true, false);
    if (n_out >= 256) ISST_SK(256, false, false);
    ISST_SK(128, false, false);
  }
  if (e.dual) {
    switch (sbn) {
      case 16: ISST_SK(16, true, true);
      case 32: ISST_SK(32, true, true);
      case 64: ISST_SK(64, true, true);
      default: ISST_SK(128, true, true);
    }
  }
  switch (sbn) {
    case 16: ISST_SK(16, false, true);
    case 32: ISST_SK(32, false, true);
    case 64: ISST_SK(64, false, true);
    default: ISST_SK(128, false, true);
  }
#undef ISST_SK
  return set_error("gemm: unreachable");
}

// ------------------------------------------------------------------------------------------------
// weight ingestion helpers
// ------------------------------------------------------------------------------------------------
template <typename Src>
__global__ void cvt_rows_kernel(const Src* __restrict__ src, bf16* __restrict__ dst, long long rows, long long cols,
                                float scale) {
  const long long n = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = __float2bfloat16_rn(static_cast<float>(src[i]) * scale);
}
template <typename Src>
__global__ void cvt_vec_kernel(const Src* __restrict__ src, float* __restrict__ dst, long long n, float scale) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = bf16_round(static_cast<float>(src[i])) * scale;   // parameters live in the model dtype (bf16)
}
// conv weight [Co][Ci][k] -> [Co][k][Ci]
template <typename Src>
__global__ void cvt_conv_kernel(const Src* __restrict__ src, bf16* __restrict__ dst, int Co, int Ci, int k) {
  const long long n = static_cast<long long>(Co) * Ci * k;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kk = i % k;
    const int ci = (i / k) % Ci;
    const int co = i / (static_cast<long long>(k) * Ci);
    dst[(static_cast<long long>(co) * k + kk) * Ci + ci] = __float2bfloat16_rn(static_cast<float>(src[i]));
  }
}
// conv0 weight [C][1][k] -> fp32 [k][C]
template <typename Src>
__global__ void cvt_conv0_kernel(const Src* __restrict__ src, float* __restrict__ dst, int C, int k) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * k; i += gridDim.x * blockDim.x) {
    const int kk = i % k, c = i / k;
    dst[kk * C + c] = bf16_round(static_cast<float>(src[i]));
  }
}
__global__ void fill_pattern_kernel(bf16* p, long long n, uint32_t seed) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint32_t h = static_cast<uint32_t>(i) * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    p[i] = __float2bfloat16_rn((static_cast<float>(h & 0xffff) / 65536.f - 0.5f));
  }
}

static int numel(const int64_t* shape, int ndim, long long* out) {
  long long n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  *out = n;
  return 0;
}

// returns a device pointer to the source data (copying host data into the staging buffer)
static int stage_source(isst_ctx* ctx, const void* data, size_t bytes, const void** dev) {
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, data);
  if (e == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged)) {
    *dev = data;
    return 0;
  }
  cudaGetLastError();
  if (bytes > ctx->staging_bytes) {
    if (ctx->staging) cudaFree(ctx->staging);
    ctx->staging = nullptr;
    ctx->staging_bytes = 0;
    ISST_CUDA(cudaMalloc(&ctx->staging, bytes));
    ctx->staging_bytes = bytes;
  }
  ISST_CUDA(cudaMemcpy(ctx->staging, data, bytes, cudaMemcpyHostToDevice));
  *dev = ctx->staging;
  return 0;
}

static int load_matrix(isst_ctx* ctx, bf16* dst, long long rows, long long cols, const void* data, const int64_t* shape,
                       int ndim, int dtype, float scale = 1.f) {
  long long n;
  numel(shape, ndim, &n);
  ISST_CHECK(n == rows * cols, "weight has the wrong number of elements");
  const void* src;
  ISST_TRY(stage_source(ctx, data, static_cast<size_t>(n) * (dtype == ISST_DTYPE_F32 ? 4 : 2), &src));
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 16));
  if (dtype == ISST_DTYPE_F32) cvt_rows_kernel<float><<<blocks, 256>>>(static_cast<const float*>(src), dst, rows, cols, scale);
  else cvt_rows_kernel<bf16><<<blocks, 256>>>(static_cast<const bf16*>(src), dst, rows, cols, scale);
  ISST_CUDA(cudaGetLastError());
  ISST_CUDA(cudaDeviceSynchronize());
  return 0;
}
static int load_vector(isst_ctx* ctx, float* dst, long long n_expect, const void* data, const int64_t* shape, int ndim,
                       int dtype, float scale = 1.f) {
  long long n;
  numel(shape, ndim, &n);
  ISST_CHECK(n == n_expect, "vector weight has the wrong number of elements");
  const void* src;
  ISST_TRY(stage_source(ctx, data, static_cast<size_t>(n) * (dtype == ISST_DTYPE_F32 ? 4 : 2), &src));
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 1024));
  if (dtype == ISST_DTYPE_F32) cvt_vec_kernel<float><<<blocks, 256>>>(static_cast<const float*>(src), dst, n, scale);
  else cvt_vec_kernel<bf16><<<blocks, 256>>>(static_cast<const bf16*>(src), dst, n, scale);
  ISST_CUDA(cudaGetLastError());
  ISST_CUDA(cudaDeviceSynchronize());
  return 0;
}
static int load_conv(isst_ctx* ctx, bf16* dst, int Co, int Ci, int k, const void* data, const int64_t* shape, int ndim,
                     int dtype) {
  long long n;
  numel(shape, ndim, &n);
  ISST_CHECK(n == static_cast<long long>(Co) * Ci * k, "conv weight has the wrong number of elements");
  const void* src;
  ISST_TRY(stage_source(ctx, data, static_cast<size_t>(n) * (dtype == ISST_DTYPE_F32 ? 4 : 2), &src));
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 16));
  if (dtype == ISST_DTYPE_F32) cvt_conv_kernel<float><<<blocks, 256>>>(static_cast<const float*>(src), dst, Co, Ci, k);
  else cvt_conv_kernel<bf16><<<blocks, 256>>>(static_cast<const bf16*>(src), dst, Co, Ci, k);
  ISST_CUDA(cudaGetLastError());
  ISST_CUDA(cudaDeviceSynchronize());
  return 0;
}

static int alloc_w2d(Weight2D& w, int rows, int K) {
  w.rows = rows;
  w.K = K;
  return dev_alloc(&w.ptr, static_cast<size_t>(rows) * K);
}

// ------------------------------------------------------------------------------------------------
// debug taps
// ------------------------------------------------------------------------------------------------
static int tap(isst_ctx* ctx, cudaStream_t st, const std::string& name, const void* src, size_t bytes,
               size_t offset = 0, size_t total = 0) {
  if (!ctx->debug) return 0;
  if (total == 0) total = bytes;
  auto it = ctx->taps.find(name);
  if (it == ctx->taps.end() || it->second.second != total) {
    if (it != ctx->taps.end()) cudaFree(it->second.first);
    void* p = nullptr;
    ISST_CUDA(cudaMalloc(&p, total));
    ISST_CUDA(cudaMemsetAsync(p, 0, total, st));
    ctx->taps[name] = {p, total};
    it = ctx->taps.find(name);
  }
  ISST_CUDA(cudaMemcpyAsync(static_cast<char*>(it->second.first) + offset, src, bytes, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// batch metadata: a handful of int arrays packed into one pinned buffer and copied once per call
// ------------------------------------------------------------------------------------------------
constexpr size_t kEncMetaInts = 2048;   // region reserved for isst_encode_chunk
struct MetaBuilder {
  isst_ctx* ctx;
  size_t used = 0;
  int* host(size_t off) { return ctx->h_meta + off; }
  int* dev(size_t off) { return ctx->d_meta + off; }
  size_t alloc(size_t n) {
    size_t o = used;
    used += (n + 3) & ~size_t(3);
    return o;
  }
};

// ------------------------------------------------------------------------------------------------
// encoder
// ------------------------------------------------------------------------------------------------
static int norm_rows(isst_ctx* ctx, cudaStream_t st, bool rms, bool gelu, const bf16* in, bf16* out, const float* w,
                     const float* b, const int* gather, int rows, int C, float eps, DeferredSum ds = DeferredSum{nullptr, 0, 0, nullptr}) {
  ISST_CHECK(C % 8 == 0 && C <= 4096, "norm_rows: unsupported width");
  if (rows == 0) return 0;
  ProfScope ps(ctx, st, P_NORM, 0.0, static_cast<double>(rows) * C * 4);
  // many narrow rows (conv stack, encoder LayerNorms): one warp per row
  if (rows > 512 && !gather && !ds.part && (C == 256 || C == 512 || C == 1024)) {
    const dim3 grid(ceil_div(rows, 8));
#define ISST_NORMW(RMS, GELU) \
  do { \
    if (C == 256) ISST_CUDA(launch_k(ctx, norm_rows_warp_kernel<RMS, GELU, 1>, grid, dim3(256), 0, st, in, out, w, b, rows, eps)); \
    else if (C == 512) ISST_CUDA(launch_k(ctx, norm_rows_warp_kernel<RMS, GELU, 2>, grid, dim3(256), 0, st, in, out, w, b, rows, eps)); \
    else ISST_CUDA(launch_k(ctx, norm_rows_warp_kernel<RMS, GELU, 4>, grid, dim3(256), 0, st, in, out, w, b, rows, eps)); \
  } while (0)
    if (rms) ISST_NORMW(true, false);
    else if (gelu) ISST_NORMW(false, true);
    else ISST_NORMW(false, false);
#undef ISST_NORMW
    LAUNCH_CHECK(ctx);
    return 0;
  }
  // few rows (decode: one row per stream): one 16-byte column group per thread, every load in flight at once
  // few rows (decode): one 16-byte group per thread, up to 512 threads, every load in flight (latency);
  // many 4096-wide rows (prefill, 1408 rows): 128 threads x 4 groups, 16 CTAs per SM - all rows resident in one wave
  // (norm class 7.08 -> 6.64 ms per step)
  const bool wide = (rows <= 512 && C >= 1024) || (C > 1024 && !(rows > 512 && C <= 4096));
  const int threads = wide ? ((C / 8 + 31) / 32) * 32 : 128;
#define ISST_NORM(RMS, GELU) \
  do { \
    if (wide) ISST_CUDA(launch_k(ctx, norm_rows_kernel<RMS, GELU, 1>, dim3(rows), dim3(threads), 0, st, in, out, w, b, gather, C, eps, ds)); \
    else ISST_CUDA(launch_k(ctx, norm_rows_kernel<RMS, GELU, 4>, dim3(rows), dim3(threads), 0, st, in, out, w, b, gather, C, eps, ds)); \
  } while (0)
  if (rms) ISST_NORM(true, false);
  else if (gelu) ISST_NORM(false, true);
  else ISST_NORM(false, false);
#undef ISST_NORM
  LAUNCH_CHECK(ctx);
  return 0;
}

static int conv_len(int n, int k, int s) { return (n - k) / s + 1; }

static int encode_chunk(isst_ctx* ctx, cudaStream_t st, int n, const int* slots_h, const int* d_slots,
                        const int* d_prefix, int n_new, int multiplier) {
  const isst_config& c = ctx->cfg;
  const int C = ctx->C;
  const int window = ctx->n_tail + n_new;
  // ---- conv feature extractor (E1) ----
  int T = conv_len(window, c.conv_k[0], c.conv_s[0]);
  {
    ProfScope ps(ctx, st, P_CONV0, 2.0 * n * T * C * c.conv_k[0], static_cast<double>(n) * (window * 4 + static_cast<double>(T) * C * 2));
    dim3 grid(ceil_div(T, kConv0FramesPerCta), n);
    const size_t smem = (static_cast<size_t>(c.conv_k[0]) * C + kConv0FramesPerCta * c.conv_s[0] + c.conv_k[0]) * 4;
    ISST_CUDA(launch_k(ctx, conv0_ln_gelu_kernel, grid, dim3(256), smem, st, ctx->d_pcm, ctx->tail, d_slots, n_new, ctx->n_tail,
                       ctx->conv[0].w0_t, ctx->conv[0].bias, ctx->conv[0].ln_w, ctx->conv[0].ln_b, ctx->conv_a, C, c.conv_k[0],
                       c.conv_s[0], T));
    LAUNCH_CHECK(ctx);
    ISST_CUDA(launch_k(ctx, update_tail_kernel, dim3(n), dim3(128), 0, st, ctx->d_pcm, ctx->tail, d_slots, n_new, ctx->n_tail));
    LAUNCH_CHECK(ctx);
  }
  bf16* cur = ctx->conv_a;
  bf16* nxt = ctx->conv_b;
  for (int j = 1; j < c.n_conv; ++j) {
    const int k = c.conv_k[j], s = c.conv_s[j];
    const int To = conv_len(T, k, s);
    ActView v{cur, To, k * C, C, s, ceil_div(T, s) + (k - 1) / s, static_cast<long long>(T) * C, n};
    // keep the row-group extent inside the batch stride (the last group of an odd T spills one row
    // into the next batch / the buffer's pad row; those elements are never multiplied into valid rows)
    v.row_groups = ceil_div(T, s);
    Epilogue e;
    e.bias = ctx->conv[j].bias;
    ISST_TRY(gemm(ctx, st, v, ctx->conv[j].w, C, nxt, C, static_cast<long long>(To) * C, e));
    ISST_TRY(norm_rows(ctx, st, false, true, nxt, nxt, ctx->conv[j].ln_w, ctx->conv[j].ln_b, nullptr, n * To, C, 1e-5f));
    std::swap(cur, nxt);
    T = To;
  }
  const int frames = T;   // new frames per stream
  ISST_CHECK(frames * ctx->total_stride == n_new && frames % c.block_size == 0 && frames <= c.block_size * multiplier,
             "chunk does not produce a whole number of block_size-frame segments");
  const int M = n * frames;
  ISST_TRY(tap(ctx, st, "enc_conv", cur, static_cast<size_t>(M) * C * 2));
  // ---- LayerNorm(C) + post_extract_proj (E3) ----
  ISST_TRY(norm_rows(ctx, st, false, false, cur, nxt, ctx->feat_ln_w, ctx->feat_ln_b, nullptr, M, C, 1e-5f));
  const int D = c.enc_dim, F = c.enc_ffn, H = c.enc_heads, HD = D / H;
  {
    Epilogue e;
    e.bias = ctx->post_b;
    ISST_TRY(gemm(ctx, st, plain_view(nxt, M, C), ctx->post_proj, D, ctx->ex, D, 0, e));
  }
  ISST_TRY(tap(ctx, st, "enc_post_proj", ctx->ex, static_cast<size_t>(M) * D * 2));
  if (c.enc_no_rope) {   // patch_speech_encoder.py:488-493
    ProfScope ps(ctx, st, P_APPEND, 0.0, static_cast<double>(M) * D * 2 * 2);
    ISST_CUDA(launch_k(ctx, enc_sinusoid_add_kernel, dim3(ceil_div(frames * D, 256), n), dim3(256), 0, st, ctx->ex, d_prefix, frames, D));
    LAUNCH_CHECK(ctx);
  }
  // ---- 24 pre-LN layers (E4-E11) ----
  const int blocksize = c.block_size * multiplier;
  {
    // (cos, sin) of the absolute frame indices of this chunk: one table for all layers
    ProfScope ps(ctx, st, P_APPEND, 0.0, static_cast<double>(M) * (D / c.enc_heads / 2) * 8);
    ISST_CUDA(launch_k(ctx, rope_table_kernel<true>, dim3(ceil_div(frames * (D / c.enc_heads / 2), 256), n), dim3(256), 0, st,
                       ctx->enc_rope_tab, nullptr, nullptr, nullptr, d_prefix, nullptr, nullptr, ctx->enc_inv_freq,
                       D / c.enc_heads / 2, frames));
    LAUNCH_CHECK(ctx);
  }
  for (int l = 0; l < c.enc_layers; ++l) {
    EncLayerW& w = ctx->enc[l];
    ISST_TRY(norm_rows(ctx, st, false, false, ctx->ex, ctx->eh, w.ln1_w, w.ln1_b, nullptr, M, D, 1e-5f));
    {
      Epilogue e;
      e.bias = w.bqkv;
      ISST_TRY(gemm(ctx, st, plain_view(ctx->eh, M, D), w.wqkv, 3 * D, ctx->eqkv, 3 * D, 0, e));
    }
    bf16* kr = ctx->enc_k + static_cast<size_t>(l) * ctx->enc_layer_elems;
    bf16* vr = ctx->enc_v + static_cast<size_t>(l) * ctx->enc_layer_elems;
    {
      ProfScope ps(ctx, st, P_APPEND, 0.0, static_cast<double>(n) * frames * D * 2 * 4);
      dim3 grid(ceil_div(frames * D / 8, 256), n);
      const float inv_base = 1.f / 512.f;   // rotary_embedding_torch xpos_scale_base
      ISST_CUDA(launch_k(ctx, enc_rope_append_kernel, grid, dim3(256), 0, st, ctx->eqkv, kr, vr, d_slots, d_prefix, ctx->enc_rope_tab, frames, H, HD, ctx->enc_cap,
                         static_cast<const float*>(ctx->enc_xpos_base), c.max_cache_size, inv_base));
      LAUNCH_CHECK(ctx);
      if (c.enc_xpos) {
        dim3 gx(ceil_div((c.max_cache_size + frames) * D / 8, 256), n);
        ISST_CUDA(launch_k(ctx, enc_xpos_keys_kernel, gx, dim3(256), 0, st, static_cast<const bf16*>(kr), ctx->enc_kx, d_slots, d_prefix, frames, H, HD,
                           ctx->enc_cap, c.max_cache_size, static_cast<const float*>(ctx->enc_xpos_base), inv_base));
        LAUNCH_CHECK(ctx);
      }
    }
    {
      EncAttnParams ep{};
      ep.qkv = ctx->eqkv; ep.out = ctx->eattn; ep.k_ring = c.enc_xpos ? ctx->enc_kx : kr; ep.v_ring = vr; ep.slots = d_slots;
      ep.prefix = d_prefix;
      ep.T = frames; ep.H = H; ep.cap = ctx->enc_cap; ep.max_cache = c.max_cache_size; ep.blocksize = blocksize;
      LlmAttnParams lp{};
      constexpr int NW = 4;
      double keys = 0;
      for (int b = 0; b < n; ++b) keys += std::min(ctx->streams[slots_h[b]].enc_prefix, c.max_cache_size) + frames;
      ProfScope ps(ctx, st, P_ATTN_ENC, 4.0 * frames * keys * D, keys * D * 2 * 2 + static_cast<double>(M) * D * 2 * 2);
      dim3 grid(ceil_div(frames, NW * 16), H, n);
      constexpr int NS = 3;
      constexpr int smem = chunk_attn_smem_bytes<64, NW, NS>();
      ISST_CHECK(HD == 64, "encoder attention kernel is built for head_dim 64");
      ISST_TRY(ensure_smem(ctx, chunk_attention_kernel<64, true, NW, NS>, smem));
      ISST_CUDA(launch_k(ctx, chunk_attention_kernel<64, true, NW, NS>, grid, dim3(NW * 32), smem, st, ep, lp));
      LAUNCH_CHECK(ctx);
    }
    {
      Epilogue e;
      e.bias = w.bo; e.resid = ctx->ex; e.ldr = D;
      ISST_TRY(gemm(ctx, st, plain_view(ctx->eattn, M, D), w.wo, D, ctx->ex, D, 0, e));
    }
    ISST_TRY(norm_rows(ctx, st, false, false, ctx->ex, ctx->eh, w.ln2_w, w.ln2_b, nullptr, M, D, 1e-5f));
    {
      Epilogue e;
      e.bias = w.b1; e.act = 1;
      ISST_TRY(gemm(ctx, st, plain_view(ctx->eh, M, D), w.w1, F, ctx->effn, F, 0, e));
    }
    {
      Epilogue e;
      e.bias = w.b2; e.resid = ctx->ex; e.ldr = D;
      ISST_TRY(gemm(ctx, st, plain_view(ctx->effn, M, F), w.w2, D, ctx->ex, D, 0, e));
    }
    if (ctx->debug) ISST_TRY(tap(ctx, st, "enc_layer_" + std::to_string(l), ctx->ex, static_cast<size_t>(M) * D * 2));
  }
  ISST_TRY(norm_rows(ctx, st, false, false, ctx->ex, ctx->eh, ctx->enc_ln_w, ctx->enc_ln_b, nullptr, M, D, 1e-5f));
  ISST_TRY(tap(ctx, st, "enc_out", ctx->eh, static_cast<size_t>(M) * D * 2));
  // ---- length adapter (E13): k == s strided convs are reshapes + GEMMs ----
  const bf16* a_in = ctx->eh;
  int rows = M, width = D;
  bf16* bufs[2] = {ctx->ead0, ctx->ead1};
  for (int j = 0; j < c.n_adapter; ++j) {
    const int k = c.adapter_k[j], s = c.adapter_s[j];
    ISST_CHECK(k == s && (rows / n) % s == 0, "length adapter needs kernel == stride and divisible frame count");
    rows /= s;
    Epilogue e;
    ISST_TRY(gemm(ctx, st, plain_view(a_in, rows, k * width), ctx->adapter[j].w, c.adapter_dim[j], bufs[j & 1],
                  c.adapter_dim[j], 0, e));
    ISST_TRY(norm_rows(ctx, st, false, true, bufs[j & 1], bufs[j & 1], ctx->adapter[j].ln_w, ctx->adapter[j].ln_b,
                       nullptr, rows, c.adapter_dim[j], 1e-5f));
    a_in = bufs[j & 1];
    width = c.adapter_dim[j];
  }
  {
    Epilogue e;
    e.bias = ctx->proj_b;
    ISST_TRY(gemm(ctx, st, plain_view(a_in, rows, width), ctx->proj, c.hidden, ctx->speech, c.hidden, 0, e));
  }
  ctx->speech_rows_per_stream = rows / n;
  ISST_TRY(tap(ctx, st, "speech_feats", ctx->speech, static_cast<size_t>(rows) * c.hidden * 2));
  // cache.n_steps += T  (patch_speech_encoder.py:533)
  for (int b = 0; b < n; ++b) ctx->streams[slots_h[b]].enc_prefix += frames;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// LLM forward over packed new tokens (prefill: T_b >= 1; decode: T_b == 1)
// ------------------------------------------------------------------------------------------------
struct LlmBatch {
  int n = 0, M = 0, max_T = 0;
  const int* d_slots = nullptr;
  const int* d_tok_base = nullptr;
  const int* d_T = nullptr;
  const int* d_last_row = nullptr;
  const int* d_active = nullptr;   // may be null
  bool decode = false;
  float* all_logits = nullptr;     // isst_forward_all: logits of EVERY position [M][vocab] instead of the last rows
  double kv_tokens = 0;            // sum over streams of the KV length attended to (profiling only)
  double qk_pairs = 0;             // sum over streams of T_b * L_b (profiling only)
  int max_L = 0;                   // longest KV length attended to in this batch (grid sizing)
  // beam search decode: rows come in groups of `group` beams that share their sentence's prompt pages; the shared
  // prefix [0, key_hi[g]) is attended to once per group and the private tails per beam, in one launch
  // (decode_attention_group_kernel)
  int group = 1;
  const int* d_key_hi = nullptr;   // [n / group]
  const int* d_tail_page = nullptr;   // [n / group] page-table index of the first private page
  int max_prefix = 0;
  double prefix_tokens = 0;        // sum over groups of the shared prefix length (profiling only)
};

static PagedKV paged_kv(isst_ctx* ctx, int layer) {
  PagedKV kv;
  kv.pool = ctx->kv_pool + static_cast<size_t>(layer) * ctx->kv_layer_elems;
  kv.page_table = ctx->d_page_table;
  kv.kv_len = ctx->d_kv_len;
  kv.sys_len = ctx->d_sys_len;
  kv.ring_start = ctx->d_ring_start;
  kv.pages_per_stream = ctx->pages_per_stream;
  kv.kv_heads = ctx->cfg.kv_heads;
  kv.head_dim = ctx->cfg.head_dim;
  return kv;
}

// Decode attention launch: split count from the longest stream so that the grid is a few waves of
// 2 CTAs per SM; every split is a whole number of 64-key tiles.
static int decode_splits_for(isst_ctx* ctx, int n, int max_L) {
  if (ctx->opt_dec_splits > 0) return std::max(1, std::min(ctx->decode_splits, ctx->opt_dec_splits));
  const int tiles_total = std::max(1, ceil_div(max_L, kDecTile) + 1);
  // long-lived CTAs stream best (measured: 1 split at 64 streams x 8 kv heads = 512 CTAs beats 2-3 splits by 10%)
  const int target = std::max(1, std::min(ctx->decode_splits, ceil_div(2 * ctx->sm_count, n * ctx->cfg.kv_heads)));
  const int tiles_per = ceil_div(tiles_total, target);
  return ceil_div(tiles_total, tiles_per);
}
struct DecodeFuse {          // fused RoPE + KV append inside the decode attention kernel (see DecodeParams2)
  const float* part = nullptr;
  int n_part = 0;
  long long part_stride = 0;
  const int* active = nullptr;
  bool on = false;
};
static int launch_decode_attention(isst_ctx* ctx, cudaStream_t st, const bf16* qkv, const PagedKV& kv, const int* d_slots,
                                   int n, int splits, float scale_log2, const DecodeFuse& fz = DecodeFuse{}) {
  ISST_TRY(ensure_smem(ctx, decode_attention_mma_kernel<4>, kDecSmemBytes));
  ctx->paths[splits == 1 ? "decode_attention_direct" : "decode_attention_split"]++;
  DecodeParams2 dp{};
  dp.qkv = qkv; dp.q_sys = ctx->lq_sys; dp.kv = kv; dp.slots = d_slots;
  dp.out = ctx->lattn; dp.part_o = ctx->part_o; dp.part_ml = ctx->part_ml; dp.H = ctx->cfg.heads;
  dp.splits = splits; dp.scale_log2 = scale_log2;
  dp.fuse = fz.on ? 1 : 0; dp.part = fz.part; dp.n_part = fz.n_part; dp.part_stride = fz.part_stride;
  dp.tab_ring = ctx->llm_rope_ring; dp.tab_sys = ctx->llm_rope_sys; dp.active = fz.active;
  ISST_CUDA(launch_k(ctx, decode_attention_mma_kernel<4>, dim3(splits, ctx->cfg.kv_heads, n), dim3(kDecThreads), kDecSmemBytes, st, dp));
  LAUNCH_CHECK(ctx);
  return 0;
}


// ------------------------------------------------------------------------------------------------
// fused decode-layer chain (decode_chain.cuh): phase list builder + launcher
// ------------------------------------------------------------------------------------------------
struct ChainBuilder {
  chain::Params p{};
  int bn = 64;
  bool prefill = false;      // path accounting only
  double flops = 0, bytes = 0;
  const void* amap_ptr[chain::kMaxMaps] = {nullptr, nullptr, nullptr, nullptr};
  int amap_K[chain::kMaxMaps] = {0, 0, 0, 0};
};
static void chain_begin(ChainBuilder& cb, int n_tok) {
  cb.p.n_tok = n_tok;
  cb.bn = n_tok <= 16 ? 16 : (n_tok <= 32 ? 32 : (n_tok <= 64 ? 64 : (n_tok <= 128 ? 128 : 256)));
}
// out = act[n_tok, K] . W^T; returns the number of k-splits (EPI_PART: partials [splits][n_tok][n_out] in `out`)
static int chain_gemm(isst_ctx* ctx, ChainBuilder& cb, const bf16* act, const Weight2D& w, int n_out, int dual, int epi,
                      void* out, int* splits_out) {
  ISST_CHECK(cb.p.n_phases < chain::kMaxPhases && cb.p.n_wmaps < chain::kMaxMaps, "decode chain: too many phases");
  ISST_CHECK(w.K % tc::kBK == 0, "decode chain: K must be a multiple of 64");
  chain::Phase& ph = cb.p.ph[cb.p.n_phases++];
  ph = chain::Phase{};
  ph.kind = chain::PH_GEMM;
  ph.wmap = cb.p.n_wmaps;
  cb.p.wmaps[cb.p.n_wmaps++] = w.map;
  int ai = -1;
  for (int i = 0; i < cb.p.n_amaps; ++i)
    if (cb.amap_ptr[i] == act && cb.amap_K[i] == w.K) ai = i;
  if (ai < 0) {
    ISST_CHECK(cb.p.n_amaps < chain::kMaxMaps, "decode chain: too many activation buffers");
    ai = cb.p.n_amaps++;
    ISST_TRY(get_act_map(ctx, &cb.p.amaps[ai], plain_view(act, cb.p.n_tok, w.K), cb.bn));
    cb.amap_ptr[ai] = act; cb.amap_K[ai] = w.K;
  }
  ph.amap = ai;
  // a unit streams TWO weight tiles against one activation tile: gate / up rows of the same 128 features, or two
  // adjacent 128-feature tiles
  const int tiles128 = ceil_div(n_out, tc::kBM);
  ph.tiles = dual ? tiles128 : ceil_div(tiles128, 2);
  ph.tile_rows = dual ? tc::kBM : 2 * tc::kBM;
  ph.sub_off = dual ? n_out : tc::kBM;
  ph.num_kb = w.K / tc::kBK;
  ph.dual = dual;
  ph.epi = epi; ph.n_out = n_out; ph.out = out;
  int S = 1;
  if (epi == chain::EPI_PART) {
    S = chain_splits(ctx->sm_count, n_out, ph.num_kb, false);
    while (S > 1 && static_cast<size_t>(S) * cb.p.n_tok * n_out > ctx->defer_ws_floats) --S;
    ISST_CHECK(static_cast<size_t>(S) * cb.p.n_tok * n_out <= ctx->defer_ws_floats, "decode chain: partial workspace too small");
  }
  ph.splits = S;
  if (splits_out) *splits_out = S;
  const double nw = dual ? 2.0 : 1.0;
  cb.flops += 2.0 * cb.p.n_tok * static_cast<double>(n_out) * w.K * nw;
  cb.bytes += nw * n_out * static_cast<double>(w.K) * 2 + static_cast<double>(cb.p.n_tok) * w.K * 2 +
              static_cast<double>(cb.p.n_tok) * n_out * (epi == chain::EPI_SILU ? 2 : 4) * S;
  return 0;
}
static int chain_rows(ChainBuilder& cb, const bf16* x_in, bf16* x_out, bf16* h_out, const float* w, const float* part,
                      int n_part, long long part_stride, const int* gather, int n_rows, int C, float eps) {
  ISST_CHECK(cb.p.n_phases < chain::kMaxPhases, "decode chain: too many phases");
  ISST_CHECK(C % 8 == 0 && C <= 4096 && n_part <= 8, "decode chain: unsupported row width / split count");
  chain::Phase& ph = cb.p.ph[cb.p.n_phases++];
  ph = chain::Phase{};
  ph.kind = chain::PH_ROWS;
  ph.x_in = x_in; ph.x_out = x_out; ph.h_out = h_out; ph.w = w; ph.part = n_part > 0 ? part : nullptr; ph.n_part = n_part;
  ph.part_stride = part_stride; ph.gather = gather; ph.n_rows = n_rows; ph.C = C; ph.eps = eps;
  cb.bytes += static_cast<double>(n_rows) * C * (4.0 + 4.0 * n_part);
  return 0;
}
// Folded RMSNorm (chain::Phase): the last GEMM phase added sums its split partials + the residual itself and leaves
// x (residual stream), x * w and the per-tile squared sums; the next GEMM phase scales by 1 / rms in its epilogue.
static int chain_fold_reduce(ChainBuilder& cb, const bf16* x_in, bf16* x_out, bf16* h_out, const float* w, float* sq_out) {
  chain::Phase& ph = cb.p.ph[cb.p.n_phases - 1];
  ISST_CHECK(ph.kind == chain::PH_GEMM && ph.epi == chain::EPI_PART && !ph.dual && ph.tiles <= chain::kMaxTiles && ph.splits <= 8 &&
                 ph.n_out % 8 == 0,
             "decode chain: this phase cannot reduce in place");
  ph.fold_reduce = 1; ph.x_in = x_in; ph.x_out = x_out; ph.h_out = h_out; ph.w = w; ph.sq_out = sq_out;
  // the same algorithmic row work as the RMSNorm row phase it replaces (partials + residual in, residual + activation out)
  cb.bytes += static_cast<double>(cb.p.n_tok) * ph.n_out * (4.0 + 4.0 * ph.splits);
  return 0;
}
static void chain_scale(ChainBuilder& cb, const float* sq_in, int n_sq, int C, float eps) {
  chain::Phase& ph = cb.p.ph[cb.p.n_phases - 1];
  ph.sq_in = sq_in; ph.n_sq = n_sq; ph.C = C; ph.eps = eps;
}

template <int kBN, bool kFold>
static int chain_launch_bn(isst_ctx* ctx, cudaStream_t st, ChainBuilder& cb) {
  using C = chain::Cfg<kBN>;
  auto kern = chain::decode_chain_kernel<kBN, kFold>;
  ISST_TRY(ensure_smem(ctx, kern, C::kSmemBytes));
  const int G = ctx->sm_count;
  cb.p.bar = ctx->chain_bar;
  cb.p.tile_ctr = ctx->chain_ctr;
  cb.p.ctr_parity = ctx->chain_parity;
  if (kFold) ctx->chain_parity ^= 1;                  // folded launches alternate between the two counter halves
  for (int i = 0; i + 1 < cb.p.n_phases; ++i) {       // every CTA arrives once at every phase but the last
    cb.p.bar_base[i] = ctx->chain_base[i];
    ctx->chain_base[i] += static_cast<unsigned long long>(G);
  }
  ProfScope ps(ctx, st, P_GEMM_STREAM, cb.flops, cb.bytes);
  ISST_CUDA(launch_k(ctx, kern, dim3(G), dim3(C::kThreads), C::kSmemBytes, st, cb.p));
  LAUNCH_CHECK(ctx);
  ctx->paths[std::string(cb.prefill ? "prefill_chain" : "decode_chain") + std::to_string(kBN)]++;
  if (kFold) ctx->paths["decode_chain_folded"]++;
  return 0;
}
static int chain_launch(isst_ctx* ctx, cudaStream_t st, ChainBuilder& cb) {
  ISST_CHECK(cb.p.n_phases >= 1, "decode chain: empty");
  bool fold = false, rows = false;
  for (int i = 0; i < cb.p.n_phases; ++i) {
    fold |= cb.p.ph[i].fold_reduce || cb.p.ph[i].sq_in;
    rows |= cb.p.ph[i].kind == chain::PH_ROWS;
  }
  ISST_CHECK(!(fold && rows), "decode chain: a folded chain has no row phases");
#define ISST_CHAIN(BN) return fold ? chain_launch_bn<BN, true>(ctx, st, cb) : chain_launch_bn<BN, false>(ctx, st, cb)
  if (cb.bn == 16) ISST_CHAIN(16);
  if (cb.bn == 32) ISST_CHAIN(32);
  if (cb.bn == 64) ISST_CHAIN(64);
  if (cb.bn == 128) ISST_CHAIN(128);
  ISST_CHAIN(256);
#undef ISST_CHAIN
}

// Chunk-prefill attention of layer l over the paged KV (q / q_sys rotated and K / V appended by llm_rope_append_kernel).
static int launch_prefill_attention(isst_ctx* ctx, cudaStream_t st, const LlmBatch& lb, int l, const PagedKV& kv, float scale_log2) {
  const isst_config& c = ctx->cfg;
  const int H = c.heads, Hkv = c.kv_heads, HD = c.head_dim;
  const int M = lb.M;
  {
      LlmAttnParams lp{};
      lp.qkv = ctx->lqkv; lp.q_sys = ctx->lq_sys; lp.out = ctx->lattn; lp.kv = kv; lp.slots = lb.d_slots; lp.tok_base = lb.d_tok_base;
      lp.T = lb.d_T; lp.H = H; lp.scale_log2 = scale_log2;
      ProfScope ps(ctx, st, P_ATTN_PREFILL, 4.0 * lb.qk_pairs * H * HD,
                   lb.kv_tokens * Hkv * HD * 2 * 2 + static_cast<double>(M) * H * HD * 2 * 2);
      ISST_TRY(ensure_smem(ctx, prefill_attention_tc_kernel<4, false>, kPaSmemBytes));
      // few streams: one CTA per (row tile, kv head, stream) leaves most SMs idle and walks the whole KV serially -
      // cut the key range into splits (fp32 partials merged by decode_combine_kernel)
      const int row_tiles = ceil_div(4 * lb.max_T, 128);
      const int ctas = row_tiles * Hkv * lb.n;
      const int key_tiles = ceil_div(lb.max_L, kPaKT) + 1;
      int ks = 1;
      if (2 * ctas <= ctx->sm_count) ks = std::max(1, std::min({ctx->sm_count / ctas, key_tiles / 2, 8}));
      const size_t part_cap = static_cast<size_t>(c.max_batch) * H * ctx->decode_splits;     // (row, head, split) slots of part_o
      while (ks > 1 && static_cast<size_t>(M) * H * ks > part_cap) --ks;
      lp.key_splits = ks; lp.part_o = ctx->part_o; lp.part_ml = ctx->part_ml;
      lp.kv_row0 = static_cast<int>(static_cast<size_t>(l) * (ctx->kv_layer_elems / HD));
      lp.l2_ahead = ctx->opt_pa_l2_ahead;
      ctx->paths[ks > 1 ? "prefill_attention_tc_keysplit" : "prefill_attention_tc_unsplit"]++;
      if (ks > 1) {
        ISST_TRY(ensure_smem(ctx, prefill_attention_tc_kernel<4, true>, kPaSmemBytes));
        ISST_CUDA(launch_k(ctx, prefill_attention_tc_kernel<4, true>, dim3(row_tiles * ks, Hkv, lb.n), dim3(kPaThreads),
                           kPaSmemBytes, st, ctx->kv_map, lp));
      } else {
        ISST_CUDA(launch_k(ctx, prefill_attention_tc_kernel<4, false>, dim3(row_tiles, Hkv, lb.n), dim3(kPaThreads),
                           kPaSmemBytes, st, ctx->kv_map, lp));
      }
      LAUNCH_CHECK(ctx);
      if (ks > 1) {
        ISST_CUDA(launch_k(ctx, decode_combine_kernel, dim3(M * H), dim3(128), 0, st, ctx->part_o, ctx->part_ml, ctx->lattn, H, HD, ks));
        LAUNCH_CHECK(ctx);
      }
  }
  return 0;
}

// One LLM forward on the fused chain: per layer one attention launch (decode: fused RoPE / append; chunk-prefill of up
// to 128 rows: RoPE-append + tensor-core prefill attention) + ONE chain launch instead of five GEMM / norm launches.
static int llm_decode_chain(isst_ctx* ctx, cudaStream_t st, const LlmBatch& lb) {
  const isst_config& c = ctx->cfg;
  const int D = c.hidden, H = c.heads, Hkv = c.kv_heads, HD = c.head_dim, F = c.ffn;
  const int QKV = (H + 2 * Hkv) * HD;
  const int M = lb.M;
  const float scale_log2 = 1.4426950408889634f / std::sqrt(static_cast<float>(HD));
  ISST_CHECK(HD == 128 && H / Hkv == 4, "LLM attention kernels are built for head_dim 128 and 4:1 GQA");
  {
    ProfScope ps(ctx, st, P_APPEND, 0.0, static_cast<double>(M) * (HD / 2) * 16);
    ISST_CUDA(launch_k(ctx, rope_table_kernel<false>, dim3(ceil_div(lb.max_T * (HD / 2), 128), lb.n), dim3(128), 0, st,
                       ctx->llm_rope_ring, ctx->llm_rope_sys, lb.d_tok_base, lb.d_T, ctx->d_kv_len, ctx->d_evicted, lb.d_active,
                       ctx->llm_inv_freq, HD / 2, 0));
    LAUNCH_CHECK(ctx);
  }
  const bool grouped = lb.group == 4 && lb.d_key_hi != nullptr;   // beam search: shared-prefix attention
  // decode ("chain_fold", default on): both RMSNorms of a layer are folded into the GEMMs around them (chain::Phase) -
  // one grid barrier and the row phase less per norm; h = bf16(x * w) reaches the next GEMM un-normalised and 1 / rms is
  // applied to its fp32 accumulators: one bf16 rounding of the normalised activation less than the reference's RMSNorm
  // module (closer to the fp32 oracle); "chain_fold" = 0 keeps the module's two roundings
  const bool fold = ctx->opt_fold && lb.decode;
  float* sq_a = ctx->chain_sq;                                     // o_proj -> gate/up
  float* sq_b = ctx->chain_sq + chain::kMaxTiles * 256;            // down / head -> QKV, lm_head
  const int d_tiles = ceil_div(ceil_div(D, tc::kBM), 2);           // tile pairs of a GEMM with D outputs
  int qkv_splits = 1;
  {
    ChainBuilder cb;
    chain_begin(cb, M);
    cb.prefill = !lb.decode;
    // the head chain (input norm of layer 0 + its QKV; one launch of 33 per forward) keeps the row phase
    ISST_TRY(chain_rows(cb, ctx->lx, nullptr, ctx->lh, ctx->llm[0].rms1, nullptr, 0, 0, nullptr, M, D, c.rms_eps));
    ISST_TRY(chain_gemm(ctx, cb, ctx->lh, ctx->llm[0].wqkv, QKV, 0, chain::EPI_PART, ctx->defer_ws, &qkv_splits));
    ISST_TRY(chain_launch(ctx, st, cb));
  }
  for (int l = 0; l < c.layers; ++l) {
    LlmLayerW& w = ctx->llm[l];
    PagedKV kv = paged_kv(ctx, l);
    if (!lb.decode) {
      // chunk-prefill: RoPE + append of the new rows (summing the QKV split partials), then tensor-core prefill attention
      {
        ProfScope ps(ctx, st, P_APPEND, 0.0, static_cast<double>(M) * (2.0 * H + 4.0 * Hkv) * HD * 2);
        dim3 grid(ceil_div(lb.max_T * (H + 2 * Hkv) * (HD / 16), 128), lb.n);
        ISST_CUDA(launch_k(ctx, llm_rope_append_kernel, grid, dim3(128), 0, st, ctx->lqkv, ctx->lq_sys, kv, lb.d_slots, lb.d_tok_base, lb.d_T,
                           lb.d_active, ctx->llm_rope_ring, ctx->llm_rope_sys, H, static_cast<const float*>(ctx->defer_ws), qkv_splits,
                           static_cast<long long>(M) * QKV));
        LAUNCH_CHECK(ctx);
      }
      ISST_TRY(launch_prefill_attention(ctx, st, lb, l, kv, scale_log2));
    } else if (grouped) {
      // beam search: rows come in groups of 4 beams that share their sentence's prompt pages; the shared prefix is
      // attended to once per group.  The kernel completes (QKV split partials), rotates and appends by itself
      const double tok = lb.prefix_tokens + (lb.kv_tokens - lb.prefix_tokens * lb.group);
      ProfScope ps(ctx, st, P_ATTN_DECODE, 4.0 * lb.kv_tokens * H * HD, tok * Hkv * HD * 2 * 2);
      ISST_TRY(ensure_smem(ctx, decode_attention_group_kernel, kGrpSmemBytes));
      ctx->paths["decode_attention_group"]++;
      const int n_groups = lb.n / lb.group;
      const int splits_p = decode_splits_for(ctx, n_groups, std::max(lb.max_prefix, 1));
      DecodeGroupParams gp{};
      gp.qkv = ctx->lqkv; gp.q_sys = ctx->lq_sys; gp.kv = kv; gp.slots = lb.d_slots; gp.key_hi = lb.d_key_hi;
      gp.tail_page = lb.d_tail_page; gp.out = ctx->lattn;
      gp.part_o = ctx->part_o; gp.part_ml = ctx->part_ml; gp.H = H; gp.splits = splits_p; gp.scale_log2 = scale_log2;
      gp.fuse = 1; gp.part = static_cast<const float*>(ctx->defer_ws); gp.n_part = qkv_splits;
      gp.part_stride = static_cast<long long>(M) * QKV;
      gp.tab_ring = ctx->llm_rope_ring; gp.tab_sys = ctx->llm_rope_sys; gp.active = lb.d_active;
      ISST_CUDA(launch_k(ctx, decode_attention_group_kernel, dim3(splits_p, Hkv, n_groups), dim3(kDecThreads), kGrpSmemBytes, st, gp));
      LAUNCH_CHECK(ctx);
      if (splits_p > 1) {
        ISST_CUDA(launch_k(ctx, decode_combine_kernel, dim3(lb.n * H), dim3(128), 0, st, ctx->part_o, ctx->part_ml, ctx->lattn, H, HD, splits_p));
        LAUNCH_CHECK(ctx);
      }
    } else {
      ProfScope ps(ctx, st, P_ATTN_DECODE, 4.0 * lb.kv_tokens * H * HD, lb.kv_tokens * Hkv * HD * 2 * 2);
      const int splits = decode_splits_for(ctx, lb.n, lb.max_L);
      DecodeFuse fz;
      fz.on = true; fz.active = lb.d_active;
      fz.part = ctx->defer_ws; fz.n_part = qkv_splits; fz.part_stride = static_cast<long long>(M) * QKV;
      ISST_TRY(launch_decode_attention(ctx, st, ctx->lqkv, kv, lb.d_slots, lb.n, splits, scale_log2, fz));
      if (splits > 1) {
        ISST_CUDA(launch_k(ctx, decode_combine_kernel, dim3(lb.n * H), dim3(128), 0, st, ctx->part_o, ctx->part_ml, ctx->lattn, H, HD, splits));
        LAUNCH_CHECK(ctx);
      }
    }
    ChainBuilder cb;
    chain_begin(cb, M);
    cb.prefill = !lb.decode;
    int so = 1, sd = 1;
    ISST_TRY(chain_gemm(ctx, cb, ctx->lattn, w.wo, D, 0, chain::EPI_PART, ctx->defer_ws, &so));
    if (fold) ISST_TRY(chain_fold_reduce(cb, ctx->lx, ctx->lx, ctx->lh, w.rms2, sq_a));
    else ISST_TRY(chain_rows(cb, ctx->lx, ctx->lx, ctx->lh, w.rms2, ctx->defer_ws, so, static_cast<long long>(M) * D, nullptr, M, D, c.rms_eps));
    ISST_TRY(chain_gemm(ctx, cb, ctx->lh, w.wgu, F, 1, chain::EPI_SILU, ctx->lgu, nullptr));
    if (fold) chain_scale(cb, sq_a, d_tiles, D, c.rms_eps);
    ISST_TRY(chain_gemm(ctx, cb, ctx->lgu, w.wd, D, 0, chain::EPI_PART, ctx->defer_ws, &sd));
    if (fold) {
      // down_proj sums itself: residual stream, x * (next norm weight), squared sums; the next QKV / lm_head scales
      const bool last = l + 1 == c.layers;
      ISST_TRY(chain_fold_reduce(cb, ctx->lx, last ? nullptr : ctx->lx, ctx->lh, last ? ctx->final_norm : ctx->llm[l + 1].rms1, sq_b));
      if (!last) ISST_TRY(chain_gemm(ctx, cb, ctx->lh, ctx->llm[l + 1].wqkv, QKV, 0, chain::EPI_PART, ctx->defer_ws, &qkv_splits));
      else ISST_TRY(chain_gemm(ctx, cb, ctx->lh, ctx->lm_head, c.vocab, 0, chain::EPI_F32, ctx->logits, nullptr));   // decode: row b = stream b
      chain_scale(cb, sq_b, d_tiles, D, c.rms_eps);
    } else if (l + 1 < c.layers) {
      ISST_TRY(chain_rows(cb, ctx->lx, ctx->lx, ctx->lh, ctx->llm[l + 1].rms1, ctx->defer_ws, sd, static_cast<long long>(M) * D, nullptr, M, D, c.rms_eps));
      ISST_TRY(chain_gemm(ctx, cb, ctx->lh, ctx->llm[l + 1].wqkv, QKV, 0, chain::EPI_PART, ctx->defer_ws, &qkv_splits));
    } else {
      // final norm on the last row of every stream + lm_head (the reference computes and discards the other rows)
      ISST_TRY(chain_rows(cb, ctx->lx, nullptr, ctx->llast, ctx->final_norm, ctx->defer_ws, sd, static_cast<long long>(M) * D, lb.d_last_row, lb.n, D, c.rms_eps));
      // decode: one row per stream, lm_head is the chain's last phase; chunk-prefill: the chain's token columns are the
      // M prompt rows while lm_head sees the n gathered rows only - it runs as its own GEMM below
      if (lb.decode) ISST_TRY(chain_gemm(ctx, cb, ctx->llast, ctx->lm_head, c.vocab, 0, chain::EPI_F32, ctx->logits, nullptr));
    }
    if (l == 1) cb.p.dbg = ctx->gemm_dbg;             // phase stamps of a middle layer's chain ("gemm_stamps" tap, debug bit 1)
    ISST_TRY(chain_launch(ctx, st, cb));
  }
  if (!lb.decode) {
    Epilogue e;
    e.out_f32 = 1;
    ISST_TRY(gemm(ctx, st, plain_view(ctx->llast, lb.n, D), ctx->lm_head, c.vocab, ctx->logits, c.vocab, 0, e));
  }
  ISST_CUDA(launch_k(ctx, advance_kv_len_kernel, dim3(ceil_div(lb.n, 128)), dim3(128), 0, st, ctx->d_kv_len, lb.d_slots, lb.d_T, lb.d_active, lb.n));
  LAUNCH_CHECK(ctx);
  return 0;
}

static int llm_forward(isst_ctx* ctx, cudaStream_t st, const LlmBatch& lb, bool tap_layers) {
  const isst_config& c = ctx->cfg;
  // the fused chain: decode forwards up to 256 rows, chunk-prefill up to 128 rows (beyond that the GEMMs are tensor-bound
  // and go to the CTA-pair kernel); the per-layer debug taps exist on the operator path only
  if (ctx->opt_chain && !lb.all_logits && !(tap_layers && ctx->tap_llm_layers) &&
      (lb.decode ? (lb.M <= 256 && lb.M == lb.n && (lb.group == 1 || (lb.group == 4 && lb.d_key_hi))) : lb.M <= 128))
    return llm_decode_chain(ctx, st, lb);
  const int D = c.hidden, H = c.heads, Hkv = c.kv_heads, HD = c.head_dim, F = c.ffn;
  const int QKV = (H + 2 * Hkv) * HD;
  const int M = lb.M;
  const float scale_log2 = 1.4426950408889634f / std::sqrt(static_cast<float>(HD));
  ISST_CHECK(HD == 128 && H / Hkv == 4, "LLM attention kernels are built for head_dim 128 and 4:1 GQA");
  {
    // (cos, sin) of the new tokens' absolute / reference positions: one table pair for all layers
    ProfScope ps(ctx, st, P_APPEND, 0.0, static_cast<double>(M) * (HD / 2) * 16);
    ISST_CUDA(launch_k(ctx, rope_table_kernel<false>, dim3(ceil_div(lb.max_T * (HD / 2), 128), lb.n), dim3(128), 0, st,
                       ctx->llm_rope_ring, ctx->llm_rope_sys, lb.d_tok_base, lb.d_T, ctx->d_kv_len, ctx->d_evicted, lb.d_active,
                       ctx->llm_inv_freq, HD / 2, 0));
    LAUNCH_CHECK(ctx);
  }
  // Weight-streaming regime (few tokens): the QKV / o_proj / down_proj GEMMs leave fp32 split partials and the
  // following row kernel (RoPE-append, RMSNorm) sums them while it reads the row anyway, adds the residual and
  // writes the residual stream back - the GEMMs lose their whole cross-CTA reduction phase.
  const bool defer = M <= 128;
  DeferredSum pending{nullptr, 0, 0, nullptr};     // down_proj partials of the previous layer
  for (int l = 0; l < c.layers; ++l) {
    LlmLayerW& w = ctx->llm[l];
    ISST_TRY(norm_rows(ctx, st, true, false, ctx->lx, ctx->lh, w.rms1, nullptr, nullptr, M, D, c.rms_eps, pending));
    pending = DeferredSum{nullptr, 0, 0, nullptr};
    int qkv_splits = 0;
    {
      Epilogue e;
      e.defer = defer;
      ISST_TRY(gemm(ctx, st, plain_view(ctx->lh, M, D), w.wqkv, QKV, ctx->lqkv, QKV, 0, e));
      qkv_splits = ctx->last_defer_splits;
    }
    PagedKV kv = paged_kv(ctx, l);
    const bool grouped = lb.decode && lb.group == 4 && lb.d_key_hi;   // beam search: shared-prefix attention
    const bool fuse_append = lb.decode && !grouped;    // decode: the attention kernel rotates and appends by itself
    if (!fuse_append) {
      ProfScope ps(ctx, st, P_APPEND, 0.0, static_cast<double>(M) * (2.0 * H + 4.0 * Hkv) * HD * 2);
      dim3 grid(ceil_div(lb.max_T * (H + 2 * Hkv) * (HD / 16), 128), lb.n);
      ISST_CUDA(launch_k(ctx, llm_rope_append_kernel, grid, dim3(128), 0, st, ctx->lqkv, ctx->lq_sys, kv, lb.d_slots, lb.d_tok_base, lb.d_T,
                         lb.d_active, ctx->llm_rope_ring, ctx->llm_rope_sys, H, qkv_splits ? ctx->defer_ws : nullptr, qkv_splits,
                         static_cast<long long>(M) * QKV));
      LAUNCH_CHECK(ctx);
    }
    if (!lb.decode) {
      ISST_TRY(launch_prefill_attention(ctx, st, lb, l, kv, scale_log2));
    } else if (grouped) {
      // algorithmic bytes: the shared prefix once per sentence + every beam's private tail
      const double tok = lb.prefix_tokens + (lb.kv_tokens - lb.prefix_tokens * lb.group);
      ProfScope ps(ctx, st, P_ATTN_DECODE, 4.0 * lb.kv_tokens * H * HD, tok * Hkv * HD * 2 * 2);
      ISST_TRY(ensure_smem(ctx, decode_attention_group_kernel, kGrpSmemBytes));
      ctx->paths["decode_attention_group"]++;
      const int n_groups = lb.n / lb.group;
      const int splits_p = decode_splits_for(ctx, n_groups, std::max(lb.max_prefix, 1));
      DecodeGroupParams gp{};
      gp.qkv = ctx->lqkv; gp.q_sys = ctx->lq_sys; gp.kv = kv; gp.slots = lb.d_slots; gp.key_hi = lb.d_key_hi;
      gp.tail_page = lb.d_tail_page; gp.out = ctx->lattn;
      gp.part_o = ctx->part_o; gp.part_ml = ctx->part_ml; gp.H = H; gp.splits = splits_p; gp.scale_log2 = scale_log2;
      ISST_CUDA(launch_k(ctx, decode_attention_group_kernel, dim3(splits_p, Hkv, n_groups), dim3(kDecThreads), kGrpSmemBytes, st, gp));
      LAUNCH_CHECK(ctx);
      if (splits_p > 1) {
        ISST_CUDA(launch_k(ctx, decode_combine_kernel, dim3(lb.n * H), dim3(128), 0, st, ctx->part_o, ctx->part_ml, ctx->lattn, H, HD, splits_p));
        LAUNCH_CHECK(ctx);
      }
    } else {
      // algorithmic bytes: K and V of every attended token once (4096 * L per layer per stream)
      ProfScope ps(ctx, st, P_ATTN_DECODE, 4.0 * lb.kv_tokens * H * HD, lb.kv_tokens * Hkv * HD * 2 * 2);
      const int splits = decode_splits_for(ctx, lb.n, lb.max_L);
      DecodeFuse fz;
      if (fuse_append) {
        fz.on = true; fz.active = lb.d_active;
        if (qkv_splits) { fz.part = ctx->defer_ws; fz.n_part = qkv_splits; fz.part_stride = static_cast<long long>(M) * QKV; }
      }
      ISST_TRY(launch_decode_attention(ctx, st, ctx->lqkv, kv, lb.d_slots, lb.n, splits, scale_log2, fz));
      if (splits > 1) {
        ISST_CUDA(launch_k(ctx, decode_combine_kernel, dim3(lb.n * H), dim3(128), 0, st, ctx->part_o, ctx->part_ml, ctx->lattn, H, HD, splits));
        LAUNCH_CHECK(ctx);
      }
    }
    DeferredSum o_sum{nullptr, 0, 0, nullptr};
    {
      Epilogue e;
      e.defer = defer;
      e.resid = ctx->lx; e.ldr = D;
      ISST_TRY(gemm(ctx, st, plain_view(ctx->lattn, M, H * HD), w.wo, D, ctx->lx, D, 0, e));
      if (ctx->last_defer_splits) o_sum = DeferredSum{ctx->defer_ws, ctx->last_defer_splits, static_cast<long long>(M) * D, ctx->lx};
    }
    ISST_TRY(norm_rows(ctx, st, true, false, ctx->lx, ctx->lh, w.rms2, nullptr, nullptr, M, D, c.rms_eps, o_sum));
    {
      Epilogue e;
      e.dual = 1; e.dual_off = F;
      ISST_TRY(gemm(ctx, st, plain_view(ctx->lh, M, D), w.wgu, F, ctx->lgu, F, 0, e));
    }
    {
      Epilogue e;
      e.defer = defer;
      e.resid = ctx->lx; e.ldr = D;
      ISST_TRY(gemm(ctx, st, plain_view(ctx->lgu, M, F), w.wd, D, ctx->lx, D, 0, e));
      if (ctx->last_defer_splits) pending = DeferredSum{ctx->defer_ws, ctx->last_defer_splits, static_cast<long long>(M) * D, ctx->lx};
    }
    if (tap_layers && ctx->debug && ctx->tap_llm_layers) {
      if (pending.part) {   // materialise the layer output for the tap (debug only)
        ISST_TRY(norm_rows(ctx, st, true, false, ctx->lx, ctx->lh, w.rms1, nullptr, nullptr, M, D, c.rms_eps, pending));
        pending = DeferredSum{nullptr, 0, 0, nullptr};
      }
      ISST_TRY(tap(ctx, st, "llm_layer_" + std::to_string(l), ctx->lx, static_cast<size_t>(M) * D * 2));
    }
  }
  ISST_CUDA(launch_k(ctx, advance_kv_len_kernel, dim3(ceil_div(lb.n, 128)), dim3(128), 0, st, ctx->d_kv_len, lb.d_slots, lb.d_T, lb.d_active, lb.n));
  LAUNCH_CHECK(ctx);
  // final norm + lm_head on the LAST position of each stream only (the reference computes and discards
  // the other T-1 rows, llm.py:236-237 / SURVEY §2.3 L9)
  if (lb.all_logits) {
    // the reference's own shape: final norm + lm_head over all T positions (model/llm.py:236-237)
    ISST_TRY(norm_rows(ctx, st, true, false, ctx->lx, ctx->lh, ctx->final_norm, nullptr, nullptr, M, D, c.rms_eps, pending));
    Epilogue e;
    e.out_f32 = 1;
    ISST_TRY(gemm(ctx, st, plain_view(ctx->lh, M, D), ctx->lm_head, c.vocab, lb.all_logits, c.vocab, 0, e));
    return 0;
  }
  pending.x_out = nullptr;   // only the gathered last rows are completed; the residual stream is not needed any more
  ISST_TRY(norm_rows(ctx, st, true, false, ctx->lx, ctx->llast, ctx->final_norm, nullptr, lb.d_last_row, lb.n, D, c.rms_eps, pending));
  {
    Epilogue e;
    e.out_f32 = 1;
    ISST_TRY(gemm(ctx, st, plain_view(ctx->llast, lb.n, D), ctx->lm_head, c.vocab, ctx->logits, c.vocab, 0, e));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// KV page management (host side; device tables are refreshed before every forward)
// ------------------------------------------------------------------------------------------------
static int ensure_capacity(isst_ctx* ctx, int slot, int extra_tokens, int pin_prefix) {
  StreamHost& s = ctx->streams[slot];
  if (!s.prefilled) {
    s.sys_len = pin_prefix;
    const int sys_pages = ceil_div(pin_prefix, kPageTokens);
    s.ring_start = sys_pages * kPageTokens;
    s.prefilled = true;
  }
  const int need_len = s.kv_len + extra_tokens;
  ISST_CHECK(need_len <= ctx->cfg.max_kv_len, "stream KV length would exceed max_kv_len");
  // highest slot index needed
  const int last_slot = need_len <= s.sys_len ? need_len - 1 : need_len - 1 - s.sys_len + s.ring_start;
  const int need_pages = last_slot / kPageTokens + 1;
  ISST_CHECK(need_pages <= ctx->pages_per_stream, "stream needs more pages than pages_per_stream");
  while (static_cast<int>(s.pages.size()) < need_pages) {
    ISST_CHECK(!ctx->free_pages.empty(), "KV page pool exhausted");
    s.pages.push_back(ctx->free_pages.back());
    ctx->free_pages.pop_back();
  }
  return 0;
}

struct KvView {                 // one batch entry's KV addressing: a stream, or a beam (shared prefix pages + private tail)
  const int* pages0; int n0;    // leading page-table entries
  const int* pages1; int n1;    // entries that follow them (may be empty)
  int kv_len, sys_len, ring_start;
  long long evicted;
};
static int upload_kv_views(isst_ctx* ctx, cudaStream_t st, int n, const KvView* v) {
  const int pps = ctx->pages_per_stream;
  int* h = ctx->h_tables;
  int* h_len = h + static_cast<size_t>(ctx->cfg.max_batch) * pps;
  int* h_sys = h_len + ctx->cfg.max_batch;
  int* h_ring = h_sys + ctx->cfg.max_batch;
  int* h_evi = h_ring + ctx->cfg.max_batch;
  for (int b = 0; b < n; ++b) {
    ISST_CHECK(v[b].n0 + v[b].n1 <= pps, "batch entry needs more pages than pages_per_stream");
    std::copy(v[b].pages0, v[b].pages0 + v[b].n0, h + static_cast<size_t>(b) * pps);
    std::copy(v[b].pages1, v[b].pages1 + v[b].n1, h + static_cast<size_t>(b) * pps + v[b].n0);
    h_len[b] = v[b].kv_len; h_sys[b] = v[b].sys_len; h_ring[b] = v[b].ring_start;
    ISST_CHECK(v[b].evicted + v[b].kv_len + 4096 < 2147483647LL, "stream exceeded 2^31 tokens");
    h_evi[b] = static_cast<int>(v[b].evicted);
  }
  ISST_CUDA(cudaMemcpyAsync(ctx->d_evicted, h_evi, n * sizeof(int), cudaMemcpyHostToDevice, st));
  ISST_CUDA(cudaMemcpyAsync(ctx->d_page_table, h, static_cast<size_t>(n) * pps * sizeof(int), cudaMemcpyHostToDevice, st));
  ISST_CUDA(cudaMemcpyAsync(ctx->d_kv_len, h_len, n * sizeof(int), cudaMemcpyHostToDevice, st));
  ISST_CUDA(cudaMemcpyAsync(ctx->d_sys_len, h_sys, n * sizeof(int), cudaMemcpyHostToDevice, st));
  ISST_CUDA(cudaMemcpyAsync(ctx->d_ring_start, h_ring, n * sizeof(int), cudaMemcpyHostToDevice, st));
  return 0;
}

// The LLM kernels index the per-stream tables by BATCH entry (their `slots` argument is the identity), so one
// call's tables are four contiguous uploads from pinned memory instead of four small copies per stream.
static int upload_stream_tables(isst_ctx* ctx, cudaStream_t st, int n, const int* slots) {
  std::vector<KvView> v(n);
  for (int b = 0; b < n; ++b) {
    const StreamHost& s = ctx->streams[slots[b]];
    v[b] = KvView{s.pages.data(), static_cast<int>(s.pages.size()), nullptr, 0, s.kv_len, s.sys_len, s.ring_start, s.evicted};
  }
  return upload_kv_views(ctx, st, n, v.data());
}

static int check_batch(isst_ctx* ctx, int n, const int* ids) {
  ISST_CHECK(ctx && ctx->finalized, "context not finalized (call isst_finalize_weights)");
  ISST_CHECK(n >= 1 && n <= ctx->cfg.max_batch, "batch size out of range");
  for (int b = 0; b < n; ++b) {
    ISST_CHECK(ids[b] >= 0 && ids[b] < ctx->cfg.max_streams && ctx->streams[ids[b]].open, "bad stream id");
    for (int a = 0; a < b; ++a) ISST_CHECK(ids[a] != ids[b], "duplicate stream id in batch");
  }
  return 0;
}

}  // namespace isst

// =================================================================================================
// C-ABI
// =================================================================================================
extern "C" {

const char* isst_last_error(void) { return g_last_error.c_str(); }

int isst_create(const isst_config* cfg, int device, isst_ctx** out) {
  ISST_CHECK(cfg && out, "null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return set_error("no CUDA device: infinisst_b200 has no CPU fallback");
  ISST_CHECK(device >= 0 && device < ndev, "bad device index");
  ISST_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  ISST_CUDA(cudaGetDeviceProperties(&prop, device));
  ISST_CHECK(prop.major == 10, "infinisst_b200 kernels are built for sm_100a only");
  ISST_TRY(load_driver_api());
  isst_ctx* ctx = new isst_ctx();
  ctx->cfg = *cfg;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  const isst_config& c = ctx->cfg;
  ISST_CHECK(c.n_conv >= 2 && c.n_conv <= ISST_MAX_CONV && c.n_adapter >= 0 && c.n_adapter <= ISST_MAX_CONV, "bad conv config");
  ctx->C = c.conv_dim[0];
  for (int j = 0; j < c.n_conv; ++j) ISST_CHECK(c.conv_dim[j] == ctx->C, "all conv blocks must share one width");
  ISST_CHECK(ctx->C % 64 == 0 && ctx->C <= 512 && c.conv_k[0] <= kConv0MaxK, "unsupported conv width / kernel");
  // receptive field and total stride of the conv stack
  int rf = 1, stride = 1;
  for (int j = 0; j < c.n_conv; ++j) { rf += (c.conv_k[j] - 1) * stride; stride *= c.conv_s[j]; }
  ctx->rf = rf;
  ctx->total_stride = stride;
  ctx->n_tail = rf - 1;   // 79 + 320 for wav2vec2 (agents/infinisst.py:216-218)
  ctx->frames_max = c.block_size * c.max_multiplier;
  ctx->samples_max = ctx->frames_max * stride;
  ctx->enc_cap = c.max_cache_size + ctx->frames_max;
  ctx->pages_per_stream = ceil_div(c.max_kv_len, kPageTokens) + 2;
  const int D = c.enc_dim, F = c.enc_ffn, C = ctx->C, HID = c.hidden;
  ISST_CHECK(D % 64 == 0 && F % 64 == 0 && HID % 64 == 0 && c.ffn % 64 == 0, "model widths must be multiples of 64");

  // ---- weights ----
  ctx->conv.resize(c.n_conv);
  for (int j = 0; j < c.n_conv; ++j) {
    ConvW& w = ctx->conv[j];
    if (j == 0) ISST_TRY(dev_alloc(&w.w0_t, static_cast<size_t>(c.conv_k[0]) * C));
    else ISST_TRY(alloc_w2d(w.w, C, c.conv_k[j] * C));
    ISST_TRY(dev_alloc(&w.bias, C)); ISST_TRY(dev_alloc(&w.ln_w, C)); ISST_TRY(dev_alloc(&w.ln_b, C));
  }
  ISST_TRY(dev_alloc(&ctx->feat_ln_w, C)); ISST_TRY(dev_alloc(&ctx->feat_ln_b, C));
  ISST_TRY(alloc_w2d(ctx->post_proj, D, C)); ISST_TRY(dev_alloc(&ctx->post_b, D));
  ctx->enc.resize(c.enc_layers);
  for (auto& w : ctx->enc) {
    ISST_TRY(dev_alloc(&w.ln1_w, D)); ISST_TRY(dev_alloc(&w.ln1_b, D)); ISST_TRY(dev_alloc(&w.ln2_w, D));
    ISST_TRY(dev_alloc(&w.ln2_b, D)); ISST_TRY(dev_alloc(&w.bqkv, 3 * D)); ISST_TRY(dev_alloc(&w.bo, D));
    ISST_TRY(dev_alloc(&w.b1, F)); ISST_TRY(dev_alloc(&w.b2, D));
    ISST_TRY(alloc_w2d(w.wqkv, 3 * D, D)); ISST_TRY(alloc_w2d(w.wo, D, D));
    ISST_TRY(alloc_w2d(w.w1, F, D)); ISST_TRY(alloc_w2d(w.w2, D, F));
  }
  ISST_TRY(dev_alloc(&ctx->enc_ln_w, D)); ISST_TRY(dev_alloc(&ctx->enc_ln_b, D));
  ctx->adapter.resize(c.n_adapter);
  int width = D;
  for (int j = 0; j < c.n_adapter; ++j) {
    ISST_TRY(alloc_w2d(ctx->adapter[j].w, c.adapter_dim[j], c.adapter_k[j] * width));
    ISST_TRY(dev_alloc(&ctx->adapter[j].ln_w, c.adapter_dim[j])); ISST_TRY(dev_alloc(&ctx->adapter[j].ln_b, c.adapter_dim[j]));
    width = c.adapter_dim[j];
  }
  ISST_TRY(alloc_w2d(ctx->proj, HID, width)); ISST_TRY(dev_alloc(&ctx->proj_b, HID));
  const int QKV = (c.heads + 2 * c.kv_heads) * c.head_dim;
  ctx->llm.resize(c.layers);
  for (auto& w : ctx->llm) {
    ISST_TRY(dev_alloc(&w.rms1, HID)); ISST_TRY(dev_alloc(&w.rms2, HID));
    ISST_TRY(alloc_w2d(w.wqkv, QKV, HID)); ISST_TRY(alloc_w2d(w.wo, HID, c.heads * c.head_dim));
    ISST_TRY(alloc_w2d(w.wgu, 2 * c.ffn, HID)); ISST_TRY(alloc_w2d(w.wd, HID, c.ffn));
  }
  ISST_TRY(dev_alloc(&ctx->final_norm, HID));
  ISST_TRY(alloc_w2d(ctx->lm_head, c.vocab, HID));
  ISST_TRY(dev_alloc(&ctx->embed, static_cast<size_t>(c.vocab) * HID));
  ISST_TRY(dev_alloc(&ctx->enc_inv_freq, D / c.enc_heads / 2));
  ISST_CHECK(!(c.enc_xpos && c.enc_no_rope), "enc_xpos needs the rotary embedding (enc_no_rope = 0)");
  if (c.enc_xpos) {
    const int hd = D / c.enc_heads;
    std::vector<float> base(hd / 2);
    for (int d = 0; d < hd / 2; ++d) base[d] = (2.f * d + 0.4f * hd) / (1.4f * hd);   // rotary_embedding_torch `scale` buffer
    ISST_TRY(dev_alloc(&ctx->enc_xpos_base, hd / 2));
    ISST_CUDA(cudaMemcpy(ctx->enc_xpos_base, base.data(), base.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  ISST_TRY(dev_alloc(&ctx->llm_inv_freq, c.head_dim / 2));

  // ---- stream state ----
  ctx->streams.resize(c.max_streams);
  ISST_TRY(dev_alloc(&ctx->tail, static_cast<size_t>(c.max_streams) * ctx->n_tail));
  ctx->enc_layer_elems = static_cast<size_t>(c.max_streams) * D * ctx->enc_cap;
  ISST_TRY(dev_alloc(&ctx->enc_k, ctx->enc_layer_elems * c.enc_layers));
  ISST_TRY(dev_alloc(&ctx->enc_v, ctx->enc_layer_elems * c.enc_layers));
  if (c.enc_xpos) ISST_TRY(dev_alloc(&ctx->enc_kx, ctx->enc_layer_elems));
  ISST_TRY(dev_alloc(&ctx->d_enc_prefix, c.max_streams));
  // per-batch-entry KV tables (uploaded before every forward): a beam-search batch has streams x beams entries
  const size_t nt = static_cast<size_t>(std::max(c.max_streams, c.max_batch));
  ISST_TRY(dev_alloc(&ctx->d_page_table, nt * ctx->pages_per_stream));
  ISST_TRY(dev_alloc(&ctx->d_kv_len, nt)); ISST_TRY(dev_alloc(&ctx->d_sys_len, nt));
  ISST_TRY(dev_alloc(&ctx->d_ring_start, nt));
  ISST_TRY(dev_alloc(&ctx->d_evicted, nt));
  ISST_CUDA(cudaMemset(ctx->d_evicted, 0, nt * sizeof(int)));
  ISST_CUDA(cudaMemset(ctx->d_page_table, 0, nt * ctx->pages_per_stream * sizeof(int)));
  ISST_CUDA(cudaMemset(ctx->d_kv_len, 0, nt * sizeof(int)));
  ISST_CUDA(cudaMemset(ctx->d_sys_len, 0, nt * sizeof(int)));
  ISST_CUDA(cudaMemset(ctx->d_ring_start, 0, nt * sizeof(int)));
  ISST_CUDA(cudaMemset(ctx->d_enc_prefix, 0, c.max_streams * sizeof(int)));
  ctx->kv_layer_elems = static_cast<size_t>(c.kv_pages) * 2 * c.kv_heads * kPageTokens * c.head_dim;
  ISST_TRY(dev_alloc(&ctx->kv_pool, ctx->kv_layer_elems * c.layers));
  ISST_CUDA(cudaMemset(ctx->kv_pool, 0, ctx->kv_layer_elems * c.layers * sizeof(bf16)));   // masked rows must hold finite values
  for (int p = c.kv_pages - 1; p >= 0; --p) ctx->free_pages.push_back(p);
  {
    // pool layout [layer][page][K|V][kv_head][16 tokens][head_dim]: every (page, K|V, head) is 16 consecutive rows
    const unsigned long long rows = static_cast<unsigned long long>(ctx->kv_layer_elems / c.head_dim) * c.layers;
    ISST_CHECK(c.head_dim == 128 && rows < (1ULL << 31), "KV pool too large for one tensor map");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(c.head_dim), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(c.head_dim) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(kPageTokens)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&ctx->kv_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ctx->kv_pool, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ISST_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(kv pool) failed: " + std::to_string(static_cast<int>(r)));
  }

  // ---- workspaces ----
  const int nb = c.max_batch;
  ISST_TRY(dev_alloc(&ctx->d_pcm, static_cast<size_t>(nb) * (ctx->samples_max + ctx->n_tail)));
  const int T0 = conv_len(ctx->n_tail + ctx->samples_max, c.conv_k[0], c.conv_s[0]);
  ISST_TRY(dev_alloc(&ctx->conv_a, static_cast<size_t>(nb) * T0 * C + 2 * C));
  ISST_TRY(dev_alloc(&ctx->conv_b, static_cast<size_t>(nb) * T0 * C + 2 * C));
  ISST_CUDA(cudaMemset(ctx->conv_a, 0, (static_cast<size_t>(nb) * T0 * C + 2 * C) * 2));
  ISST_CUDA(cudaMemset(ctx->conv_b, 0, (static_cast<size_t>(nb) * T0 * C + 2 * C) * 2));
  const size_t ME = static_cast<size_t>(nb) * ctx->frames_max;
  ISST_TRY(dev_alloc(&ctx->ex, ME * D)); ISST_TRY(dev_alloc(&ctx->eh, ME * D)); ISST_TRY(dev_alloc(&ctx->eqkv, ME * 3 * D));
  ISST_TRY(dev_alloc(&ctx->eattn, ME * D)); ISST_TRY(dev_alloc(&ctx->effn, ME * F));
  int maxw = D;
  for (int j = 0; j < c.n_adapter; ++j) maxw = std::max(maxw, c.adapter_dim[j]);
  ISST_TRY(dev_alloc(&ctx->ead0, ME * maxw)); ISST_TRY(dev_alloc(&ctx->ead1, ME * maxw));
  ISST_TRY(dev_alloc(&ctx->speech, ME * HID));
  const size_t ML = static_cast<size_t>(nb) * std::max(c.max_prompt, 1);
  ISST_TRY(dev_alloc(&ctx->lx, ML * HID)); ISST_TRY(dev_alloc(&ctx->lh, ML * HID)); ISST_TRY(dev_alloc(&ctx->lqkv, ML * QKV));
  ISST_TRY(dev_alloc(&ctx->lattn, ML * c.heads * c.head_dim)); ISST_TRY(dev_alloc(&ctx->lgu, ML * c.ffn));
  ISST_TRY(dev_alloc(&ctx->llast, static_cast<size_t>(nb) * HID));
  ISST_TRY(dev_alloc(&ctx->lq_sys, ML * c.heads * c.head_dim));
  ISST_TRY(dev_alloc(&ctx->llm_rope_ring, ML * (c.head_dim / 2)));
  ISST_TRY(dev_alloc(&ctx->llm_rope_sys, ML * (c.head_dim / 2)));
  ISST_TRY(dev_alloc(&ctx->enc_rope_tab, ME * (D / c.enc_heads / 2)));
  ISST_TRY(dev_alloc(&ctx->logits, static_cast<size_t>(nb) * c.vocab));
  ISST_TRY(dev_alloc(&ctx->sel_ws.best, static_cast<size_t>(nb) * kSelParts));
  ISST_TRY(dev_alloc(&ctx->sel_ws.idx, static_cast<size_t>(nb) * kSelParts));
  ISST_TRY(dev_alloc(&ctx->sel_ws.count, static_cast<size_t>(nb)));
  ISST_CUDA(cudaMemset(ctx->sel_ws.count, 0, static_cast<size_t>(nb) * sizeof(int)));
  ISST_TRY(dev_alloc(&ctx->beam_ws, static_cast<size_t>(nb) * (kSelParts * (2 + 2 * kBeamMaxKeep) + 2 * kBeamMaxKeep)));
  ISST_TRY(dev_alloc(&ctx->beam_count, static_cast<size_t>(nb)));
  ISST_CUDA(cudaMemset(ctx->beam_count, 0, static_cast<size_t>(nb) * sizeof(int)));
  ctx->decode_splits = 32;
  ISST_TRY(dev_alloc(&ctx->part_o, static_cast<size_t>(nb) * c.heads * ctx->decode_splits * c.head_dim));
  ISST_TRY(dev_alloc(&ctx->part_ml, static_cast<size_t>(nb) * c.heads * ctx->decode_splits * 2));
  ctx->gemm_ws_floats = static_cast<size_t>(24) << 20;   // 96 MB: two parked 256 x 256 fp32 partial tiles per CTA
  ISST_TRY(dev_alloc(&ctx->gemm_ws, ctx->gemm_ws_floats));
  ctx->defer_ws_floats = static_cast<size_t>(8) * std::min(std::max(nb, 128), 256) * std::max(QKV, HID);   // <= 8 splits x <= 256 token rows x widest deferred output
  ISST_TRY(dev_alloc(&ctx->defer_ws, ctx->defer_ws_floats));
  ISST_TRY(dev_alloc(&ctx->chain_ctr, 2 * chain::kMaxPhases * chain::kMaxTiles));
  ISST_CUDA(cudaMemset(ctx->chain_ctr, 0, 2 * chain::kMaxPhases * chain::kMaxTiles * sizeof(int)));
  ISST_TRY(dev_alloc(&ctx->chain_sq, 2 * chain::kMaxTiles * 256));
  ISST_TRY(dev_alloc(&ctx->chain_bar, chain::kMaxPhases));
  ISST_CUDA(cudaMemset(ctx->chain_bar, 0, chain::kMaxPhases * sizeof(unsigned long long)));
  ctx->n_counters = 4096;
  ISST_TRY(dev_alloc(&ctx->gemm_counters, ctx->n_counters));
  ISST_CUDA(cudaMemset(ctx->gemm_counters, 0, ctx->n_counters * sizeof(int)));
  ctx->meta_ints = kEncMetaInts + static_cast<size_t>(nb) * (c.max_prompt * 4 + 256 + c.max_new_tokens * 5) + 4096;
  ISST_TRY(dev_alloc(&ctx->d_meta, ctx->meta_ints));
  ISST_CUDA(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_meta), ctx->meta_ints * sizeof(int)));
  ISST_CUDA(cudaEventCreateWithFlags(&ctx->ev_active, cudaEventDisableTiming));
  ISST_CUDA(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_tables), (static_cast<size_t>(nb) * (ctx->pages_per_stream + 4) + 16) * sizeof(int)));
  ISST_CUDA(cudaDeviceSynchronize());
  *out = ctx;
  return 0;
}

void isst_destroy(isst_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  // the context owns every device allocation it made; release the big pools explicitly
  cudaFree(ctx->beam_ws); cudaFree(ctx->beam_count); cudaFree(ctx->chain_bar); cudaFree(ctx->chain_ctr); cudaFree(ctx->chain_sq);
  cudaFree(ctx->kv_pool); cudaFree(ctx->enc_k); cudaFree(ctx->enc_v); cudaFree(ctx->embed);
  cudaFree(ctx->enc_kx); cudaFree(ctx->enc_xpos_base); cudaFree(ctx->logits_all);
  for (auto& w : ctx->llm) { cudaFree(w.wqkv.ptr); cudaFree(w.wo.ptr); cudaFree(w.wgu.ptr); cudaFree(w.wd.ptr); cudaFree(w.rms1); cudaFree(w.rms2); }
  for (auto& w : ctx->enc) { cudaFree(w.wqkv.ptr); cudaFree(w.wo.ptr); cudaFree(w.w1.ptr); cudaFree(w.w2.ptr);
    cudaFree(w.ln1_w); cudaFree(w.ln1_b); cudaFree(w.ln2_w); cudaFree(w.ln2_b); cudaFree(w.bqkv); cudaFree(w.bo); cudaFree(w.b1); cudaFree(w.b2); }
  for (auto& w : ctx->conv) { cudaFree(w.w.ptr); cudaFree(w.w0_t); cudaFree(w.bias); cudaFree(w.ln_w); cudaFree(w.ln_b); }
  for (auto& w : ctx->adapter) { cudaFree(w.w.ptr); cudaFree(w.ln_w); cudaFree(w.ln_b); }
  cudaFree(ctx->lm_head.ptr); cudaFree(ctx->post_proj.ptr); cudaFree(ctx->proj.ptr);
  void* misc[] = {ctx->feat_ln_w, ctx->feat_ln_b, ctx->post_b, ctx->enc_ln_w, ctx->enc_ln_b, ctx->proj_b, ctx->final_norm,
                  ctx->enc_inv_freq, ctx->llm_inv_freq, ctx->enc_rope_tab, ctx->llm_rope_ring, ctx->llm_rope_sys, ctx->lq_sys, ctx->d_evicted, ctx->staging,
                  ctx->tail, ctx->d_enc_prefix, ctx->d_page_table, ctx->d_kv_len, ctx->d_sys_len, ctx->d_ring_start,
                  ctx->d_pcm, ctx->conv_a, ctx->conv_b, ctx->ex, ctx->eh, ctx->eqkv, ctx->eattn, ctx->effn, ctx->ead0,
                  ctx->ead1, ctx->speech, ctx->lx, ctx->lh, ctx->lqkv, ctx->lattn, ctx->lgu, ctx->llast, ctx->logits,
                  ctx->part_o, ctx->part_ml, ctx->gemm_ws, ctx->defer_ws, ctx->gemm_counters, ctx->d_meta,
                  ctx->sel_ws.best, ctx->sel_ws.idx, ctx->sel_ws.count};
  for (void* p : misc) cudaFree(p);
  for (auto& t : ctx->taps) cudaFree(t.second.first);
  cudaFreeHost(ctx->h_meta);
  cudaFreeHost(ctx->h_tables);
  if (ctx->ev_active) cudaEventDestroy(ctx->ev_active);
  delete ctx;
}

static bool starts_with(const std::string& s, const std::string& p) { return s.compare(0, p.size(), p) == 0; }

int isst_load_weight(isst_ctx* ctx, const char* name_c, const void* data, const int64_t* shape, int ndim, int dtype) {
  ISST_CHECK(ctx && name_c && data && shape, "null argument");
  ISST_CHECK(dtype == ISST_DTYPE_F32 || dtype == ISST_DTYPE_BF16, "dtype must be f32 or bf16");
  ISST_CUDA(cudaSetDevice(ctx->device));
  const isst_config& c = ctx->cfg;
  const std::string name(name_c);
  const int D = c.enc_dim, F = c.enc_ffn, C = ctx->C, HID = c.hidden, HD = c.head_dim;
  const std::string ENC = "model.speech_encoder.speech_encoder.";
  const std::string SPE = "model.speech_encoder.";
  ctx->finalized = false;
  auto done = [&]() { ctx->loaded[name] = true; return 0; };
#define LM(dst, rows, cols) do { ISST_TRY(load_matrix(ctx, dst, rows, cols, data, shape, ndim, dtype)); return done(); } while (0)
#define LV(dst, n) do { ISST_TRY(load_vector(ctx, dst, n, data, shape, ndim, dtype)); return done(); } while (0)
  if (name == "rope.enc.inv_freq" || name == "rope.llm.inv_freq") {
    const bool enc = name[5] == 'e';
    const int half = enc ? (D / c.enc_heads / 2) : HD / 2;
    long long ne; numel(shape, ndim, &ne);
    ISST_CHECK(ne == half && dtype == ISST_DTYPE_F32, "inv_freq must be f32 [head_dim / 2]");
    ISST_CUDA(cudaMemcpy(enc ? ctx->enc_inv_freq : ctx->llm_inv_freq, data, half * sizeof(float), cudaMemcpyDefault));
    // --rope 0: zero frequencies make every (cos, sin) = (1, 0), i.e. q and k pass through the append kernel unrotated
    if (enc && c.enc_no_rope) ISST_CUDA(cudaMemset(ctx->enc_inv_freq, 0, half * sizeof(float)));
    return done();
  }
  if (starts_with(name, ENC + "feature_extractor.conv_layers.")) {
    const std::string rest = name.substr((ENC + "feature_extractor.conv_layers.").size());
    const int j = atoi(rest.c_str());
    ISST_CHECK(j >= 0 && j < c.n_conv, "conv layer index out of range");
    const std::string leaf = rest.substr(rest.find('.') + 1);
    ConvW& w = ctx->conv[j];
    if (leaf == "0.weight") {
      if (j == 0) {
        long long n; numel(shape, ndim, &n);
        ISST_CHECK(n == static_cast<long long>(C) * c.conv_k[0], "conv0 weight size");
        const void* src;
        ISST_TRY(stage_source(ctx, data, static_cast<size_t>(n) * (dtype == ISST_DTYPE_F32 ? 4 : 2), &src));
        if (dtype == ISST_DTYPE_F32) cvt_conv0_kernel<float><<<32, 256>>>(static_cast<const float*>(src), w.w0_t, C, c.conv_k[0]);
        else cvt_conv0_kernel<bf16><<<32, 256>>>(static_cast<const bf16*>(src), w.w0_t, C, c.conv_k[0]);
        ISST_CUDA(cudaDeviceSynchronize());
        return done();
      }
      ISST_TRY(load_conv(ctx, w.w.ptr, C, C, c.conv_k[j], data, shape, ndim, dtype));
      return done();
    }
    if (leaf == "0.bias") LV(w.bias, C);
    if (leaf == "2.1.weight") LV(w.ln_w, C);
    if (leaf == "2.1.bias") LV(w.ln_b, C);
    return set_error("unknown conv key: " + name);
  }
  if (name == ENC + "layer_norm.weight") LV(ctx->feat_ln_w, C);
  if (name == ENC + "layer_norm.bias") LV(ctx->feat_ln_b, C);
  if (name == ENC + "post_extract_proj.weight") LM(ctx->post_proj.ptr, D, C);
  if (name == ENC + "post_extract_proj.bias") LV(ctx->post_b, D);
  if (name == ENC + "encoder.layer_norm.weight") LV(ctx->enc_ln_w, D);
  if (name == ENC + "encoder.layer_norm.bias") LV(ctx->enc_ln_b, D);
  if (starts_with(name, ENC + "encoder.layers.")) {
    const std::string rest = name.substr((ENC + "encoder.layers.").size());
    const int i = atoi(rest.c_str());
    ISST_CHECK(i >= 0 && i < c.enc_layers, "encoder layer index out of range");
    const std::string leaf = rest.substr(rest.find('.') + 1);
    EncLayerW& w = ctx->enc[i];
    // q is pre-scaled by head_dim^-0.5 (patch_speech_encoder.py:768); exact for head_dim 64 (power of two)
    const float qs = 1.0f / std::sqrt(static_cast<float>(D / c.enc_heads));
    if (leaf == "self_attn.q_proj.weight") { ISST_TRY(load_matrix(ctx, w.wqkv.ptr, D, D, data, shape, ndim, dtype, qs)); return done(); }
    if (leaf == "self_attn.k_proj.weight") LM(w.wqkv.ptr + static_cast<size_t>(D) * D, D, D);
    if (leaf == "self_attn.v_proj.weight") LM(w.wqkv.ptr + static_cast<size_t>(2) * D * D, D, D);
    if (leaf == "self_attn.q_proj.bias") { ISST_TRY(load_vector(ctx, w.bqkv, D, data, shape, ndim, dtype, qs)); return done(); }
    if (leaf == "self_attn.k_proj.bias") LV(w.bqkv + D, D);
    if (leaf == "self_attn.v_proj.bias") LV(w.bqkv + 2 * D, D);
    if (leaf == "self_attn.out_proj.weight") LM(w.wo.ptr, D, D);
    if (leaf == "self_attn.out_proj.bias") LV(w.bo, D);
    if (leaf == "self_attn_layer_norm.weight") LV(w.ln1_w, D);
    if (leaf == "self_attn_layer_norm.bias") LV(w.ln1_b, D);
    if (leaf == "final_layer_norm.weight") LV(w.ln2_w, D);
    if (leaf == "final_layer_norm.bias") LV(w.ln2_b, D);
    if (leaf == "fc1.weight") LM(w.w1.ptr, F, D);
    if (leaf == "fc1.bias") LV(w.b1, F);
    if (leaf == "fc2.weight") LM(w.w2.ptr, D, F);
    if (leaf == "fc2.bias") LV(w.b2, D);
    // rotary_embedding_torch buffers (freqs: the host passes layer 0's copy as rope.enc.inv_freq; scale / dummy /
    // cached_* are persistent in some versions of the library): recomputed here, ignored like unused modules
    if (starts_with(leaf, "self_attn.rotary_emb.")) return 0;
    return set_error("unknown encoder layer key: " + name);
  }
  if (starts_with(name, SPE + "length_shrink.conv_layers.")) {
    const std::string rest = name.substr((SPE + "length_shrink.conv_layers.").size());
    const int j = atoi(rest.c_str());
    ISST_CHECK(j >= 0 && j < c.n_adapter, "adapter layer index out of range");
    const std::string leaf = rest.substr(rest.find('.') + 1);
    const int cin = j == 0 ? D : c.adapter_dim[j - 1];
    if (leaf == "0.weight") { ISST_TRY(load_conv(ctx, ctx->adapter[j].w.ptr, c.adapter_dim[j], cin, c.adapter_k[j], data, shape, ndim, dtype)); return done(); }
    if (leaf == "2.1.weight") LV(ctx->adapter[j].ln_w, c.adapter_dim[j]);
    if (leaf == "2.1.bias") LV(ctx->adapter[j].ln_b, c.adapter_dim[j]);
    return set_error("unknown adapter key: " + name);
  }
  if (name == SPE + "proj.weight") LM(ctx->proj.ptr, HID, ctx->proj.K);
  if (name == SPE + "proj.bias") LV(ctx->proj_b, HID);
  if (name == "model.embed_tokens.weight") LM(ctx->embed, c.vocab, HID);
  if (name == "model.norm.weight") LV(ctx->final_norm, HID);
  if (name == "lm_head.weight") LM(ctx->lm_head.ptr, c.vocab, HID);
  if (starts_with(name, "model.layers.")) {
    const std::string rest = name.substr(std::string("model.layers.").size());
    const int i = atoi(rest.c_str());
    ISST_CHECK(i >= 0 && i < c.layers, "LLM layer index out of range");
    const std::string leaf = rest.substr(rest.find('.') + 1);
    LlmLayerW& w = ctx->llm[i];
    const size_t qrows = static_cast<size_t>(c.heads) * HD, kvrows = static_cast<size_t>(c.kv_heads) * HD;
    if (leaf == "self_attn.q_proj.weight") LM(w.wqkv.ptr, qrows, HID);
    if (leaf == "self_attn.k_proj.weight") LM(w.wqkv.ptr + qrows * HID, kvrows, HID);
    if (leaf == "self_attn.v_proj.weight") LM(w.wqkv.ptr + (qrows + kvrows) * HID, kvrows, HID);
    if (leaf == "self_attn.o_proj.weight") LM(w.wo.ptr, HID, qrows);
    if (leaf == "mlp.gate_proj.weight") LM(w.wgu.ptr, c.ffn, HID);
    if (leaf == "mlp.up_proj.weight") LM(w.wgu.ptr + static_cast<size_t>(c.ffn) * HID, c.ffn, HID);
    if (leaf == "mlp.down_proj.weight") LM(w.wd.ptr, HID, c.ffn);
    if (leaf == "input_layernorm.weight") LV(w.rms1, HID);
    if (leaf == "post_attention_layernorm.weight") LV(w.rms2, HID);
    if (starts_with(leaf, "self_attn.rotary_emb.")) return 0;   // older HF checkpoints carry per-layer inv_freq buffers
    return set_error("unknown LLM layer key: " + name);
  }
#undef LM
#undef LV
  return 0;   // unused module (encoder.pos_conv.*, mask_emb, ...): ignored like load_state_dict on dead weights
}

int isst_finalize_weights(isst_ctx* ctx) {
  ISST_CHECK(ctx, "null ctx");
  ISST_CUDA(cudaSetDevice(ctx->device));
  const isst_config& c = ctx->cfg;
  // every tensor of the hot path must have been loaded
  std::vector<std::string> need = {"rope.enc.inv_freq", "rope.llm.inv_freq",
                                   "model.embed_tokens.weight", "model.norm.weight", "lm_head.weight",
                                   "model.speech_encoder.proj.weight", "model.speech_encoder.proj.bias"};
  const std::string ENC = "model.speech_encoder.speech_encoder.";
  for (int j = 0; j < c.n_conv; ++j)
    for (const char* leaf : {"0.weight", "0.bias", "2.1.weight", "2.1.bias"})
      need.push_back(ENC + "feature_extractor.conv_layers." + std::to_string(j) + "." + leaf);
  for (const char* leaf : {"layer_norm.weight", "layer_norm.bias", "post_extract_proj.weight", "post_extract_proj.bias",
                           "encoder.layer_norm.weight", "encoder.layer_norm.bias"})
    need.push_back(ENC + leaf);
  for (int i = 0; i < c.enc_layers; ++i)
    for (const char* leaf : {"self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight",
                             "self_attn.out_proj.weight", "self_attn.q_proj.bias", "self_attn.k_proj.bias",
                             "self_attn.v_proj.bias", "self_attn.out_proj.bias", "self_attn_layer_norm.weight",
                             "self_attn_layer_norm.bias", "final_layer_norm.weight", "final_layer_norm.bias",
                             "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"})
      need.push_back(ENC + "encoder.layers." + std::to_string(i) + "." + leaf);
  for (int j = 0; j < c.n_adapter; ++j)
    for (const char* leaf : {"0.weight", "2.1.weight", "2.1.bias"})
      need.push_back("model.speech_encoder.length_shrink.conv_layers." + std::to_string(j) + "." + leaf);
  for (int i = 0; i < c.layers; ++i)
    for (const char* leaf : {"self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight",
                             "self_attn.o_proj.weight", "mlp.gate_proj.weight", "mlp.up_proj.weight",
                             "mlp.down_proj.weight", "input_layernorm.weight", "post_attention_layernorm.weight"})
      need.push_back("model.layers." + std::to_string(i) + "." + leaf);
  for (const auto& k : need)
    if (!ctx->loaded.count(k)) return set_error("missing weight: " + k);
  // tensor maps
  for (int j = 1; j < c.n_conv; ++j) ISST_TRY(make_weight_map(ctx->conv[j].w));
  for (auto& a : ctx->adapter) ISST_TRY(make_weight_map(a.w));
  ISST_TRY(make_weight_map(ctx->post_proj)); ISST_TRY(make_weight_map(ctx->proj)); ISST_TRY(make_weight_map(ctx->lm_head));
  for (auto& w : ctx->enc) { ISST_TRY(make_weight_map(w.wqkv)); ISST_TRY(make_weight_map(w.wo)); ISST_TRY(make_weight_map(w.w1)); ISST_TRY(make_weight_map(w.w2)); }
  for (auto& w : ctx->llm) { ISST_TRY(make_weight_map(w.wqkv)); ISST_TRY(make_weight_map(w.wo)); ISST_TRY(make_weight_map(w.wgu)); ISST_TRY(make_weight_map(w.wd)); }
  ISST_CUDA(cudaDeviceSynchronize());
  if (ctx->staging) { cudaFree(ctx->staging); ctx->staging = nullptr; ctx->staging_bytes = 0; }
  ctx->finalized = true;
  return 0;
}

int isst_stream_open(isst_ctx* ctx, int* stream_id) {
  ISST_CHECK(ctx && stream_id, "null argument");
  ISST_CUDA(cudaSetDevice(ctx->device));
  for (int s = 0; s < ctx->cfg.max_streams; ++s) {
    if (!ctx->streams[s].open) {
      ctx->streams[s] = StreamHost();
      ctx->streams[s].open = true;
      // zero carried samples == the 79+320 zero offset of the first chunk (agents/infinisst.py:216-218)
      ISST_CUDA(cudaMemset(ctx->tail + static_cast<size_t>(s) * ctx->n_tail, 0, ctx->n_tail * sizeof(float)));
      *stream_id = s;
      return 0;
    }
  }
  return set_error("no free stream slot (max_streams reached)");
}

int isst_stream_close(isst_ctx* ctx, int stream_id) {
  ISST_CHECK(ctx && stream_id >= 0 && stream_id < ctx->cfg.max_streams && ctx->streams[stream_id].open, "bad stream id");
  StreamHost& s = ctx->streams[stream_id];
  for (int p : s.pages) ctx->free_pages.push_back(p);
  s = StreamHost();
  return 0;
}

int isst_encode_chunk(isst_ctx* ctx, int n, const int* stream_ids, const float* pcm, int n_samples, int multiplier,
                      void* out_feats, void* cuda_stream) {
  ISST_TRY(check_batch(ctx, n, stream_ids));
  ISST_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  ISST_CHECK(multiplier >= 1 && multiplier <= ctx->cfg.max_multiplier, "multiplier out of range");
  // One call brings j whole segments of block_size frames, 1 <= j <= multiplier: the reference pads a short final
  // chunk to a multiple of ONE segment only (agents/infinisst.py:211-213) while the encoder mask keeps the block size
  // of the multiplier (set_blocksize, speech_encoder.py:143-145); fewer frames then give fewer speech rows.
  const int seg_samples = ctx->cfg.block_size * ctx->total_stride;
  int n_new = ctx->cfg.block_size * multiplier * ctx->total_stride;
  {
    const int body = n_samples >= ctx->n_tail && (n_samples - ctx->n_tail) % seg_samples == 0 && n_samples % seg_samples != 0
                         ? n_samples - ctx->n_tail : n_samples;
    if (body % seg_samples == 0 && body >= seg_samples && body <= n_new) n_new = body;
  }
  // A fresh stream's carried tail is zero (isst_stream_open), which IS the reference's 79+320 zero offset of the first
  // chunk: fresh and running streams may share a batch as long as every row brings exactly n_new samples.  The
  // explicit-offset form (n_new + 79+320 samples, as agents/infinisst.py:216-218 builds it) needs an all-fresh batch.
  bool fresh = true;
  for (int b = 0; b < n; ++b) fresh = fresh && ctx->streams[stream_ids[b]].enc_prefix == 0;
  const bool with_offset = fresh && n_samples == n_new + ctx->n_tail;
  ISST_CHECK(n_samples == n_new || with_offset,
             "n_samples must be j*block_size*320 with 1 <= j <= multiplier (+ 79+320 only when every stream of the batch is "
             "at its first chunk); feed pending audio one policy call at a time");
  ISST_CUDA(cudaStreamSynchronize(st));   // the pinned metadata region below may still be in flight from a previous call
  MetaBuilder mb{ctx};
  const size_t o_slots = mb.alloc(n), o_prefix = mb.alloc(n);
  ISST_CHECK(mb.used <= kEncMetaInts, "max_batch too large for the encoder metadata region");
  for (int b = 0; b < n; ++b) {
    mb.host(o_slots)[b] = stream_ids[b];
    mb.host(o_prefix)[b] = ctx->streams[stream_ids[b]].enc_prefix;
  }
  ISST_CUDA(cudaMemcpyAsync(ctx->d_meta, ctx->h_meta, mb.used * sizeof(int), cudaMemcpyHostToDevice, st));
  if (with_offset) {
    // explicit 79+320 leading samples: they become the carried tail (zeros in the reference)
    ISST_CUDA(cudaMemcpy2DAsync(ctx->d_pcm, static_cast<size_t>(n_samples) * 4, pcm, static_cast<size_t>(n_samples) * 4,
                                static_cast<size_t>(n_samples) * 4, n, cudaMemcpyDefault, st));
    // tail <- first n_tail samples (rounded like the cast at agents/infinisst.py:222), new <- rest, compacted
    // done with two strided copies on the stream
    for (int b = 0; b < n; ++b) {
      update_tail_kernel<<<1, 128, 0, st>>>(ctx->d_pcm + static_cast<size_t>(b) * n_samples, ctx->tail, mb.dev(o_slots) + b,
                                            ctx->n_tail, ctx->n_tail);
      LAUNCH_CHECK(ctx);
    }
    float* tmp = reinterpret_cast<float*>(ctx->conv_b);   // scratch, large enough (n * T0 * C bf16 >> n * n_new floats)
    ISST_CUDA(cudaMemcpy2DAsync(tmp, static_cast<size_t>(n_new) * 4, ctx->d_pcm + ctx->n_tail, static_cast<size_t>(n_samples) * 4,
                                static_cast<size_t>(n_new) * 4, n, cudaMemcpyDeviceToDevice, st));
    ISST_CUDA(cudaMemcpyAsync(ctx->d_pcm, tmp, static_cast<size_t>(n) * n_new * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    ISST_CUDA(cudaMemcpyAsync(ctx->d_pcm, pcm, static_cast<size_t>(n) * n_new * 4, cudaMemcpyDefault, st));
  }
  ISST_TRY(encode_chunk(ctx, st, n, stream_ids, mb.dev(o_slots), mb.dev(o_prefix), n_new, multiplier));
  if (out_feats) {
    ISST_CUDA(cudaMemcpyAsync(out_feats, ctx->speech, static_cast<size_t>(n) * ctx->speech_rows_per_stream * ctx->cfg.hidden * 2,
                              cudaMemcpyDefault, st));
    ISST_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}

// Builds the packed batch description shared by isst_forward / isst_generate.
static int setup_llm_batch(isst_ctx* ctx, cudaStream_t st, MetaBuilder& mb, int n, const int* stream_ids, const int* lens,
                           const int32_t* ids, const int32_t* speech_slot, int extra_decode, int pin_prefix, LlmBatch* lb,
                           size_t* o_ids_out, size_t* o_srow_out) {
  int M = 0, maxT = 0;
  for (int b = 0; b < n; ++b) {
    ISST_CHECK(lens[b] >= 1 && lens[b] <= ctx->cfg.max_prompt, "prompt length out of range");
    M += lens[b];
    maxT = std::max(maxT, lens[b]);
  }
  for (int b = 0; b < n; ++b) ISST_TRY(ensure_capacity(ctx, stream_ids[b], lens[b] + extra_decode, pin_prefix));
  ISST_TRY(upload_stream_tables(ctx, st, n, stream_ids));
  const size_t o_slots = mb.alloc(n), o_base = mb.alloc(n), o_T = mb.alloc(n), o_last = mb.alloc(n);
  const size_t o_ids = mb.alloc(M), o_srow = mb.alloc(M);
  int row = 0;
  for (int b = 0; b < n; ++b) {
    mb.host(o_slots)[b] = b;   // tables are uploaded per batch entry (upload_stream_tables)
    mb.host(o_base)[b] = row;
    mb.host(o_T)[b] = lens[b];
    mb.host(o_last)[b] = row + lens[b] - 1;
    for (int t = 0; t < lens[b]; ++t) {
      mb.host(o_ids)[row + t] = ids ? ids[row + t] : 0;
      const int ss = speech_slot ? speech_slot[row + t] : -1;
      if (ss >= 0) ISST_CHECK(ss < ctx->speech_rows_per_stream, "speech slot index beyond the encoded features");
      mb.host(o_srow)[row + t] = ss >= 0 ? b * ctx->speech_rows_per_stream + ss : -1;
      if (ids) ISST_CHECK(ids[row + t] >= 0 && ids[row + t] < ctx->cfg.vocab, "token id out of range");
    }
    row += lens[b];
  }
  lb->kv_tokens = 0; lb->qk_pairs = 0;
  for (int b = 0; b < n; ++b) {
    const double L = ctx->streams[stream_ids[b]].kv_len + lens[b];
    lb->kv_tokens += L;
    lb->qk_pairs += L * lens[b];
    lb->max_L = std::max(lb->max_L, static_cast<int>(L));
  }
  lb->n = n; lb->M = M; lb->max_T = maxT;
  lb->d_slots = mb.dev(o_slots); lb->d_tok_base = mb.dev(o_base); lb->d_T = mb.dev(o_T); lb->d_last_row = mb.dev(o_last);
  lb->d_active = nullptr; lb->decode = false;
  *o_ids_out = o_ids; *o_srow_out = o_srow;
  return 0;
}

static int forward_impl(isst_ctx* ctx, int n, const int* stream_ids, const int32_t* ids, const int* lens,
                        const int32_t* speech_slot, const void* embeds_override, int pin_prefix, float* out_logits,
                        void* cuda_stream, bool all_positions);

int isst_forward(isst_ctx* ctx, int n, const int* stream_ids, const int32_t* ids, const int* lens,
                 const int32_t* speech_slot, const void* embeds_override, int pin_prefix, float* out_logits,
                 void* cuda_stream) {
  return forward_impl(ctx, n, stream_ids, ids, lens, speech_slot, embeds_override, pin_prefix, out_logits, cuda_stream, false);
}

int isst_forward_all(isst_ctx* ctx, int n, const int* stream_ids, const int32_t* ids, const int* lens,
                     const int32_t* speech_slot, const void* embeds_override, int pin_prefix, float* out_logits,
                     void* cuda_stream) {
  ISST_CHECK(out_logits, "null argument");
  return forward_impl(ctx, n, stream_ids, ids, lens, speech_slot, embeds_override, pin_prefix, out_logits, cuda_stream, true);
}

static int forward_impl(isst_ctx* ctx, int n, const int* stream_ids, const int32_t* ids, const int* lens,
                        const int32_t* speech_slot, const void* embeds_override, int pin_prefix, float* out_logits,
                        void* cuda_stream, bool all_positions) {
  ISST_TRY(check_batch(ctx, n, stream_ids));
  ISST_CHECK(lens && (ids || embeds_override), "null argument");
  ISST_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  MetaBuilder mb{ctx, kEncMetaInts};
  LlmBatch lb;
  size_t o_ids, o_srow;
  ISST_TRY(setup_llm_batch(ctx, st, mb, n, stream_ids, lens, ids, speech_slot, 0, pin_prefix, &lb, &o_ids, &o_srow));
  ISST_CHECK(mb.used <= ctx->meta_ints, "metadata buffer too small");
  ISST_CUDA(cudaMemcpyAsync(ctx->d_meta + kEncMetaInts, ctx->h_meta + kEncMetaInts, (mb.used - kEncMetaInts) * sizeof(int),
                            cudaMemcpyHostToDevice, st));
  if (embeds_override) {
    ISST_CUDA(cudaMemcpyAsync(ctx->lx, embeds_override, static_cast<size_t>(lb.M) * ctx->cfg.hidden * 2, cudaMemcpyDefault, st));
  } else {
    ProfScope ps(ctx, st, P_EMBED, 0.0, static_cast<double>(lb.M) * ctx->cfg.hidden * 4);
    ISST_CUDA(launch_k(ctx, embed_splice_kernel, dim3(lb.M), dim3(128), 0, st, mb.dev(o_ids), mb.dev(o_srow), ctx->embed, ctx->speech, ctx->lx, ctx->cfg.hidden));
    LAUNCH_CHECK(ctx);
  }
  ISST_TRY(tap(ctx, st, "prompt_embeds", ctx->lx, static_cast<size_t>(lb.M) * ctx->cfg.hidden * 2));
  if (all_positions) {
    if (ctx->logits_all_rows < static_cast<size_t>(lb.M)) {       // not a hot path: grown on demand
      ISST_CUDA(cudaStreamSynchronize(st));
      cudaFree(ctx->logits_all);
      ctx->logits_all = nullptr; ctx->logits_all_rows = 0;
      ISST_TRY(dev_alloc(&ctx->logits_all, static_cast<size_t>(lb.M) * ctx->cfg.vocab));
      ctx->logits_all_rows = lb.M;
    }
    lb.all_logits = ctx->logits_all;
  }
  ISST_TRY(llm_forward(ctx, st, lb, true));
  if (all_positions) ISST_CUDA(cudaMemcpyAsync(out_logits, ctx->logits_all, static_cast<size_t>(lb.M) * ctx->cfg.vocab * 4, cudaMemcpyDefault, st));
  else if (out_logits) ISST_CUDA(cudaMemcpyAsync(out_logits, ctx->logits, static_cast<size_t>(n) * ctx->cfg.vocab * 4, cudaMemcpyDefault, st));
  ISST_CUDA(cudaStreamSynchronize(st));
  prof_flush(ctx);
  for (int b = 0; b < n; ++b) ctx->streams[stream_ids[b]].kv_len += lens[b];
  return 0;
}

int isst_generate(isst_ctx* ctx, int n, const int* stream_ids, const int32_t* ids, const int* lens,
                  const int32_t* speech_slot, const int32_t* enc_ids, const int* enc_lens,
                  const isst_gen_params* gen, const int32_t* forced, int32_t* out_tokens, int* out_counts,
                  void* cuda_stream) {
  ISST_TRY(check_batch(ctx, n, stream_ids));
  ISST_CHECK(ids && lens && gen && out_tokens && out_counts, "null argument");
  ISST_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const isst_config& c = ctx->cfg;
  const int max_new = gen->max_new_tokens;
  ISST_CHECK(max_new >= 1 && max_new <= c.max_new_tokens, "max_new_tokens out of range");
  ISST_CHECK(gen->n_eos >= 0 && gen->n_eos <= 8, "too many eos ids");
  MetaBuilder mb{ctx, kEncMetaInts};
  LlmBatch lb;
  size_t o_ids, o_srow;
  ISST_TRY(setup_llm_batch(ctx, st, mb, n, stream_ids, lens, ids, speech_slot, max_new - 1, gen->pin_prefix, &lb, &o_ids, &o_srow));
  // generation state
  const int ctx_cap = c.max_prompt + max_new;
  const int enc_cap = 128;
  const size_t o_ctx = mb.alloc(static_cast<size_t>(n) * ctx_cap), o_ctxlen = mb.alloc(n);
  const size_t o_enc = mb.alloc(static_cast<size_t>(n) * enc_cap), o_enclen = mb.alloc(n);
  const size_t o_active = mb.alloc(n), o_out = mb.alloc(static_cast<size_t>(n) * max_new), o_cnt = mb.alloc(n);
  const size_t o_next = mb.alloc(n), o_forced = mb.alloc(forced ? static_cast<size_t>(n) * max_new : 0);
  const size_t o_sup = mb.alloc(gen->n_suppress), o_eos = mb.alloc(8);
  const size_t o_ones = mb.alloc(n), o_iota = mb.alloc(n);
  const size_t o_picked = mb.alloc(static_cast<size_t>(n) * max_new);
  ISST_CHECK(mb.used <= ctx->meta_ints, "metadata buffer too small");
  int row = 0, eoff = 0;
  for (int b = 0; b < n; ++b) {
    for (int t = 0; t < lens[b]; ++t) mb.host(o_ctx)[static_cast<size_t>(b) * ctx_cap + t] = ids[row + t];
    mb.host(o_ctxlen)[b] = lens[b];
    const int el = enc_lens ? enc_lens[b] : 0;
    ISST_CHECK(el >= 0 && el <= enc_cap, "encoder id history longer than 128");
    for (int t = 0; t < el; ++t) mb.host(o_enc)[static_cast<size_t>(b) * enc_cap + t] = enc_ids[eoff + t];
    mb.host(o_enclen)[b] = el;
    mb.host(o_active)[b] = 1;
    mb.host(o_cnt)[b] = 0;
    mb.host(o_next)[b] = 0;
    mb.host(o_ones)[b] = 1;
    mb.host(o_iota)[b] = b;
    for (int s = 0; s < max_new; ++s) {
      mb.host(o_out)[static_cast<size_t>(b) * max_new + s] = -1;
      mb.host(o_picked)[static_cast<size_t>(b) * max_new + s] = -1;
      if (forced) mb.host(o_forced)[static_cast<size_t>(b) * max_new + s] = forced[static_cast<size_t>(b) * max_new + s];
    }
    row += lens[b];
    eoff += el;
  }
  for (int i = 0; i < gen->n_suppress; ++i) mb.host(o_sup)[i] = gen->suppress_tokens[i];
  for (int i = 0; i < gen->n_eos; ++i) mb.host(o_eos)[i] = gen->eos_token_ids[i];
  ISST_CUDA(cudaMemcpyAsync(ctx->d_meta + kEncMetaInts, ctx->h_meta + kEncMetaInts, (mb.used - kEncMetaInts) * sizeof(int),
                            cudaMemcpyHostToDevice, st));

  GenState g{};
  g.ctx_ids = mb.dev(o_ctx); g.ctx_len = mb.dev(o_ctxlen); g.enc_ids = mb.dev(o_enc); g.enc_len = mb.dev(o_enclen);
  g.active = mb.dev(o_active); g.out_tokens = mb.dev(o_out); g.out_count = mb.dev(o_cnt); g.next_token = mb.dev(o_next);
  g.forced = forced ? mb.dev(o_forced) : nullptr; g.picked = mb.dev(o_picked); g.suppress = mb.dev(o_sup); g.eos = mb.dev(o_eos);
  g.ctx_cap = ctx_cap; g.enc_cap = enc_cap; g.max_new = max_new; g.n_suppress = gen->n_suppress; g.n_eos = gen->n_eos;
  g.ngram = gen->no_repeat_ngram_size; g.penalty = gen->repetition_penalty;

  // ---- step 0: splice + chunk prefill ----
  {
    ProfScope ps(ctx, st, P_EMBED, 0.0, static_cast<double>(lb.M) * c.hidden * 4);
    ISST_CUDA(launch_k(ctx, embed_splice_kernel, dim3(lb.M), dim3(128), 0, st, mb.dev(o_ids), mb.dev(o_srow), ctx->embed, ctx->speech, ctx->lx, c.hidden));
    LAUNCH_CHECK(ctx);
  }
  ISST_TRY(tap(ctx, st, "prompt_embeds", ctx->lx, static_cast<size_t>(lb.M) * c.hidden * 2));
  ISST_TRY(llm_forward(ctx, st, lb, true));
  const size_t lbytes = static_cast<size_t>(n) * c.vocab * 4;
  ISST_TRY(tap(ctx, st, "step_logits", ctx->logits, lbytes, 0, lbytes * max_new));
  g.step = 0;
  {
    ProfScope ps(ctx, st, P_SELECT, 0.0, static_cast<double>(lbytes));
    ISST_CUDA(launch_k(ctx, greedy_select_kernel, dim3(kSelParts, n), dim3(kSelThreads), 0, st, ctx->logits, c.vocab, g, ctx->sel_ws));
    LAUNCH_CHECK(ctx);
  }
  // the select kernel applied the processors in place: `logits` now holds the scores the arg-max saw
  ISST_TRY(tap(ctx, st, "step_scores", ctx->logits, lbytes, 0, lbytes * max_new));
  // ---- decode steps ----
  LlmBatch db = lb;
  db.M = n; db.max_T = 1; db.d_T = mb.dev(o_ones); db.d_tok_base = mb.dev(o_iota); db.d_last_row = mb.dev(o_iota);
  db.d_active = mb.dev(o_active); db.decode = true;
  int* h_active = ctx->h_meta + o_active;
  for (int step = 1; step < max_new; ++step) {
    // Early exit when every stream hit EOS, without stalling the launch queue: the flags of the previous step
    // are copied to pinned memory behind an event; the host only polls it (at most one surplus step is launched,
    // and a finished stream appends nothing: llm_kv_append_kernel / advance_kv_len_kernel check `active`).
    if (step >= 2 && cudaEventQuery(ctx->ev_active) == cudaSuccess) {
      bool any = false;
      for (int b = 0; b < n; ++b) any = any || h_active[b];
      if (!any) break;
    }
    ISST_CUDA(cudaMemcpyAsync(h_active, mb.dev(o_active), n * sizeof(int), cudaMemcpyDeviceToHost, st));
    ISST_CUDA(cudaEventRecord(ctx->ev_active, st));
    {
      ProfScope ps(ctx, st, P_EMBED, 0.0, static_cast<double>(n) * c.hidden * 4);
      ISST_CUDA(launch_k(ctx, embed_splice_kernel, dim3(n), dim3(128), 0, st, mb.dev(o_next), nullptr, ctx->embed, ctx->speech, ctx->lx, c.hidden));
      LAUNCH_CHECK(ctx);
    }
    db.kv_tokens = 0;
    db.max_L = 1;
    for (int b = 0; b < n; ++b) {
      const int Lb = ctx->streams[stream_ids[b]].kv_len + lens[b] + step;
      if (h_active[b]) db.kv_tokens += Lb;
      db.max_L = std::max(db.max_L, Lb);
    }
    ISST_TRY(llm_forward(ctx, st, db, false));
    ISST_TRY(tap(ctx, st, "step_logits", ctx->logits, lbytes, lbytes * step, lbytes * max_new));
    g.step = step;
    {
      ProfScope ps(ctx, st, P_SELECT, 0.0, static_cast<double>(lbytes));
      ISST_CUDA(launch_k(ctx, greedy_select_kernel, dim3(kSelParts, n), dim3(kSelThreads), 0, st, ctx->logits, c.vocab, g, ctx->sel_ws));
      LAUNCH_CHECK(ctx);
    }
    ISST_TRY(tap(ctx, st, "step_scores", ctx->logits, lbytes, lbytes * step, lbytes * max_new));
  }
  ISST_TRY(tap(ctx, st, "step_picked", mb.dev(o_picked), static_cast<size_t>(n) * max_new * sizeof(int)));
  ISST_CUDA(cudaMemcpyAsync(ctx->h_meta + o_out, mb.dev(o_out), static_cast<size_t>(n) * max_new * sizeof(int), cudaMemcpyDeviceToHost, st));
  ISST_CUDA(cudaMemcpyAsync(ctx->h_meta + o_cnt, mb.dev(o_cnt), n * sizeof(int), cudaMemcpyDeviceToHost, st));
  ISST_CUDA(cudaStreamSynchronize(st));
  prof_flush(ctx);
  for (int b = 0; b < n; ++b) {
    const int cnt = ctx->h_meta[o_cnt + b];
    out_counts[b] = cnt;
    for (int s = 0; s < max_new; ++s) out_tokens[static_cast<size_t>(b) * max_new + s] = ctx->h_meta[o_out + static_cast<size_t>(b) * max_new + s];
    // KV holds the prompt and every chosen token except the last (drop-last rule)
    ctx->streams[stream_ids[b]].kv_len += lens[b] + std::max(0, cnt - 1);
  }
  return 0;
}


// ------------------------------------------------------------------------------------------------
// beam search (the reference's shipped decoding: scripts/infer/infinisst.sh:48, patch_hf.py:43-302, 687-967)
// ------------------------------------------------------------------------------------------------
namespace {
struct BeamHypH {               // a finished hypothesis (BeamHypotheses.beams entry, patch_hf.py:296)
  float score;                  // sum_logprobs / generated_len ** length_penalty
  std::vector<int> tokens;      // generated tokens (without the EOS that closed it)
  std::vector<int> pages;       // snapshot of the private tail pages (the KV hand-back, :113-128)
  int kv_extra;                 // generated tokens whose KV the hypothesis holds
  int order;
};
struct BeamRowH {
  std::vector<int> gen;         // tokens generated on this beam
  float score = 0.f;
  std::vector<int> priv[2];     // two sets of private tail pages (ping-pong across reorders)
  int cur = 0;
};
struct BeamGroupH {
  int slot = 0, P = 0, L1 = 0;  // stream slot, prompt length, KV length after the prompt
  int ids_off = 0;              // offset of the prompt in the packed `ids`
  int tail0 = 0;                // index in the stream's page list where the private tail starts
  int n_priv = 0;               // private pages per beam
  bool partial = false;         // the page at tail0 already holds prompt tokens (copied into every beam)
  std::vector<BeamRowH> rows;
  std::vector<BeamHypH> hyps;
  float worst = 1e9f;
  int n_added = 0;
  bool done = false;
  int steps = 0;
};
}  // namespace

static void beam_free_pages(isst_ctx* ctx, std::vector<int>& pages) {
  for (int p : pages) ctx->free_pages.push_back(p);
  pages.clear();
}
static int beam_alloc_pages(isst_ctx* ctx, std::vector<int>& pages, int n) {
  ISST_CHECK(static_cast<int>(ctx->free_pages.size()) >= n, "KV page pool exhausted (beam search tails)");
  for (int i = 0; i < n; ++i) { pages.push_back(ctx->free_pages.back()); ctx->free_pages.pop_back(); }
  return 0;
}
// BeamHypotheses.add as the reference patches it (patch_hf.py:278-302); returns pages of a dropped hypothesis
static void beam_hyp_add(isst_ctx* ctx, BeamGroupH& g, int num_beams, float length_penalty, BeamHypH&& h, float sum_logprobs,
                         int generated_len) {
  const float score = sum_logprobs / std::pow(static_cast<float>(generated_len), length_penalty);
  h.score = score;
  if (static_cast<int>(g.hyps.size()) < num_beams || score > g.worst) {
    h.order = g.n_added++;
    g.hyps.push_back(std::move(h));
    if (static_cast<int>(g.hyps.size()) > num_beams) {
      int lo = 0;
      for (int i = 1; i < static_cast<int>(g.hyps.size()); ++i)
        if (g.hyps[i].score < g.hyps[lo].score) lo = i;       // sorted()[0]: lowest score, earliest on ties
      beam_free_pages(ctx, g.hyps[lo].pages);
      g.hyps.erase(g.hyps.begin() + lo);
      float w = g.hyps[0].score;
      for (const BeamHypH& x : g.hyps) w = std::min(w, x.score);
      g.worst = w;
    } else {
      g.worst = std::min(score, g.worst);
    }
  } else {
    beam_free_pages(ctx, h.pages);
  }
}

int isst_generate_beam(isst_ctx* ctx, int n, const int* stream_ids, const int32_t* ids, const int* lens,
                       const int32_t* speech_slot, const int32_t* enc_ids, const int* enc_lens,
                       const isst_gen_params* gen, int num_beams, float length_penalty, const isst_beam_follow* follow,
                       int32_t* out_tokens, int* out_counts, float* out_scores, isst_beam_trace* trace,
                       void* cuda_stream) {
  ISST_TRY(check_batch(ctx, n, stream_ids));
  ISST_CHECK(ids && lens && gen && out_tokens && out_counts, "null argument");
  ISST_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const isst_config& c = ctx->cfg;
  const int k = num_beams, V = c.vocab;
  const int max_new = gen->max_new_tokens;
  const int R = n * k;
  ISST_CHECK(k >= 2, "beam search needs num_beams >= 2 (isst_generate is the greedy path)");
  ISST_CHECK(R <= c.max_batch, "streams x beams exceeds max_batch");
  ISST_CHECK(max_new >= 1 && max_new <= c.max_new_tokens, "max_new_tokens out of range");
  ISST_CHECK(gen->n_eos >= 0 && gen->n_eos <= 8, "too many eos ids");
  const int n_keep = std::max(2, 1 + gen->n_eos) * k;                       // patch_hf.py:869-870
  ISST_CHECK(n_keep <= kBeamMaxKeep && n_keep <= V, "too many beam candidates per sentence");
  auto is_eos = [&](int t) { for (int e = 0; e < gen->n_eos; ++e) if (gen->eos_token_ids[e] == t) return true; return false; };

  // ---- prompt prefill on one row per stream (the reference runs it on k identical rows, :305-342) ----
  MetaBuilder mb{ctx, kEncMetaInts};
  LlmBatch lb;
  size_t o_ids, o_srow;
  ISST_TRY(setup_llm_batch(ctx, st, mb, n, stream_ids, lens, ids, speech_slot, 0, gen->pin_prefix, &lb, &o_ids, &o_srow));
  const int ctx_cap = c.max_prompt + max_new;
  const int enc_cap = 128;
  const size_t o_ctx = mb.alloc(static_cast<size_t>(R) * ctx_cap), o_ctxlen = mb.alloc(R);
  const size_t o_enc = mb.alloc(static_cast<size_t>(n) * enc_cap), o_enclen = mb.alloc(n);
  const size_t o_active = mb.alloc(R), o_next = mb.alloc(R), o_score = mb.alloc(R);
  const size_t o_sup = mb.alloc(gen->n_suppress), o_ones = mb.alloc(R), o_iota = mb.alloc(R);
  const size_t o_khi = mb.alloc(n), o_tpg = mb.alloc(n);                      // shared-prefix length per sentence, first private page
  const size_t o_pairs = mb.alloc(static_cast<size_t>(R) * 2 * 8 * 2);       // page copy pairs of one step
  const size_t o_res_s = mb.alloc(static_cast<size_t>(n) * n_keep), o_res_i = mb.alloc(static_cast<size_t>(n) * n_keep);
  const size_t meta_end = mb.used;
  ISST_CHECK(meta_end <= ctx->meta_ints, "metadata buffer too small");
  std::vector<BeamGroupH> groups(n);
  int row = 0, eoff = 0;
  for (int b = 0; b < n; ++b) {
    BeamGroupH& g = groups[b];
    g.slot = stream_ids[b]; g.P = lens[b]; g.ids_off = row;
    for (int t = 0; t < lens[b]; ++t) mb.host(o_ctx)[static_cast<size_t>(b) * ctx_cap + t] = ids[row + t];
    mb.host(o_ctxlen)[b] = lens[b];
    const int el = enc_lens ? enc_lens[b] : 0;
    ISST_CHECK(el >= 0 && el <= enc_cap, "encoder id history longer than 128");
    for (int t = 0; t < el; ++t) mb.host(o_enc)[static_cast<size_t>(b) * enc_cap + t] = enc_ids[eoff + t];
    mb.host(o_enclen)[b] = el;
    row += lens[b];
    eoff += el;
  }
  for (int r = 0; r < R; ++r) { mb.host(o_ones)[r] = 1; mb.host(o_iota)[r] = r; mb.host(o_active)[r] = 1; }
  for (int b = 0; b < n; ++b) reinterpret_cast<float*>(mb.host(o_score))[b] = 0.f;
  for (int i = 0; i < gen->n_suppress; ++i) mb.host(o_sup)[i] = gen->suppress_tokens[i];
  ISST_CUDA(cudaMemcpyAsync(ctx->d_meta + kEncMetaInts, ctx->h_meta + kEncMetaInts, (meta_end - kEncMetaInts) * sizeof(int),
                            cudaMemcpyHostToDevice, st));
  {
    ProfScope ps(ctx, st, P_EMBED, 0.0, static_cast<double>(lb.M) * c.hidden * 4);
    ISST_CUDA(launch_k(ctx, embed_splice_kernel, dim3(lb.M), dim3(128), 0, st, mb.dev(o_ids), mb.dev(o_srow), ctx->embed, ctx->speech, ctx->lx, c.hidden));
    LAUNCH_CHECK(ctx);
  }
  ISST_TRY(llm_forward(ctx, st, lb, false));
  for (int b = 0; b < n; ++b) {
    StreamHost& s = ctx->streams[stream_ids[b]];
    s.kv_len += lens[b];
    groups[b].L1 = s.kv_len;
  }

  BeamSel sel{};
  sel.ctx_ids = mb.dev(o_ctx); sel.ctx_len = mb.dev(o_ctxlen); sel.enc_ids = mb.dev(o_enc); sel.enc_len = mb.dev(o_enclen);
  sel.beam_score = reinterpret_cast<const float*>(mb.dev(o_score)); sel.suppress = mb.dev(o_sup);
  sel.n_suppress = gen->n_suppress; sel.ctx_cap = ctx_cap; sel.enc_cap = enc_cap; sel.ngram = gen->no_repeat_ngram_size;
  sel.penalty = gen->repetition_penalty; sel.n_keep = n_keep;
  {
    float* w = ctx->beam_ws;
    const size_t nb = c.max_batch;
    sel.part_max = w; w += nb * kSelParts;
    sel.part_sum = w; w += nb * kSelParts;
    sel.cand_s = w; w += nb * kSelParts * kBeamMaxKeep;
    sel.cand_i = reinterpret_cast<int*>(w); w += nb * kSelParts * kBeamMaxKeep;
    sel.out_s = w; w += nb * kBeamMaxKeep;
    sel.out_i = reinterpret_cast<int*>(w);
    sel.count = ctx->beam_count;
  }
  float* h_res_s = reinterpret_cast<float*>(ctx->h_meta + o_res_s);
  int* h_res_i = ctx->h_meta + o_res_i;
  const size_t page_elems = static_cast<size_t>(2) * c.kv_heads * kPageTokens * c.head_dim;
  std::vector<int> pairs;                               // (src page, dst page) copies queued for this step
  auto flush_pairs = [&]() -> int {
    const size_t cap = static_cast<size_t>(R) * 8 * 2;  // pairs per launch (o_pairs holds 2 ints each)
    for (size_t off = 0; off < pairs.size() / 2; off += cap) {
      const size_t cnt = std::min(cap, pairs.size() / 2 - off);
      // the pinned staging is reused: wait for the previous copy of it to be consumed
      ISST_CUDA(cudaStreamSynchronize(st));
      std::copy(pairs.begin() + 2 * off, pairs.begin() + 2 * (off + cnt), mb.host(o_pairs));
      ISST_CUDA(cudaMemcpyAsync(mb.dev(o_pairs), mb.host(o_pairs), cnt * 2 * sizeof(int), cudaMemcpyHostToDevice, st));
      ProfScope ps(ctx, st, P_APPEND, 0.0, static_cast<double>(cnt) * page_elems * 2 * c.layers * 2);
      ISST_CUDA(launch_k(ctx, kv_page_copy_kernel, dim3(static_cast<unsigned>(cnt), c.layers), dim3(256), 0, st, ctx->kv_pool,
                         ctx->kv_layer_elems, static_cast<int>(page_elems), mb.dev(o_pairs)));
      LAUNCH_CHECK(ctx);
    }
    pairs.clear();
    return 0;
  };
  // Runs the selection kernels on `rows` logits rows (rows_per_group live beams per sentence), reads the 2k
  // candidates per sentence back and applies `beam_search_process` (patch_hf.py:43-157) on the host.
  std::vector<std::vector<std::pair<int, int>>> nxt(n);           // per sentence: (parent beam, token) of the next beams
  std::vector<std::vector<float>> nxt_score(n);
  auto select = [&](int step, int rows_per_group) -> int {
    sel.rows_per_group = rows_per_group;
    const int rows = n * rows_per_group;
    ProfScope ps(ctx, st, P_SELECT, 0.0, static_cast<double>(rows) * V * 4 * 3);
    ISST_CUDA(launch_k(ctx, beam_lse_kernel, dim3(kSelParts, rows), dim3(kSelThreads), 0, st, static_cast<const float*>(ctx->logits), V, sel));
    LAUNCH_CHECK(ctx);
    const int slice_bytes = ceil_div(V, kSelParts) * 4;            // the kernel keeps its vocabulary slice in shared memory
    ISST_TRY(ensure_smem(ctx, beam_topk_kernel, std::max(slice_bytes, 64 * 1024)));
    sel.write_back = follow ? 1 : 0;
    ISST_CUDA(launch_k(ctx, beam_topk_kernel, dim3(kSelParts, rows), dim3(kSelThreads), slice_bytes, st, ctx->logits, V, sel));
    LAUNCH_CHECK(ctx);
    ISST_CUDA(cudaMemcpyAsync(h_res_s, sel.out_s, static_cast<size_t>(n) * n_keep * sizeof(float), cudaMemcpyDeviceToHost, st));
    ISST_CUDA(cudaMemcpyAsync(h_res_i, sel.out_i, static_cast<size_t>(n) * n_keep * sizeof(int), cudaMemcpyDeviceToHost, st));
    ISST_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < n; ++b) {
      BeamGroupH& g = groups[b];
      nxt[b].clear(); nxt_score[b].clear();
      if (g.done) continue;
      const float* cs = h_res_s + static_cast<size_t>(b) * n_keep;
      const int* ci = h_res_i + static_cast<size_t>(b) * n_keep;
      if (trace && trace->cand_scores) {
        for (int j = 0; j < n_keep; ++j) {
          trace->cand_scores[(static_cast<size_t>(b) * max_new + step) * n_keep + j] = cs[j];
          trace->cand_index[(static_cast<size_t>(b) * max_new + step) * n_keep + j] = ci[j];
        }
      }
      const int gen_len = step + 1;                                // cur_len - decoder_prompt_len (:57, :121)
      auto close = [&](int par, float score) -> int {              // EOS candidate: the parent beam becomes a hypothesis
        BeamHypH h;
        if (step > 0) {
          const BeamRowH& pr = g.rows[par];
          h.tokens = pr.gen;
          ISST_TRY(beam_alloc_pages(ctx, h.pages, g.n_priv));
          for (int i = 0; i < g.n_priv; ++i) { pairs.push_back(pr.priv[pr.cur][i]); pairs.push_back(h.pages[i]); }
        }
        h.kv_extra = step;
        beam_hyp_add(ctx, g, k, length_penalty, std::move(h), score, gen_len);
        return 0;
      };
      if (follow) {
        // teacher forcing (parity tests): the caller dictates which beams are closed / continued; their scores are
        // this run's own processed scores, read from the score matrix the selection kernel left in `logits`
        const int32_t* fc = follow->closed + ((static_cast<size_t>(b) * max_new + step) * k) * 2;
        const int32_t* fn = follow->next + ((static_cast<size_t>(b) * max_new + step) * k) * 2;
        auto score_of = [&](int par, int tok, float* out) -> int {
          float lp = 0.f;
          ISST_CUDA(cudaMemcpy(&lp, ctx->logits + (static_cast<size_t>(b) * rows_per_group + par) * V + tok, sizeof(float), cudaMemcpyDeviceToHost));
          *out = lp + (step > 0 ? g.rows[par].score : 0.f);
          return 0;
        };
        for (int j = 0; j < k && fc[2 * j] >= 0; ++j) {
          float sc;
          ISST_TRY(score_of(fc[2 * j], fc[2 * j + 1], &sc));
          ISST_TRY(close(fc[2 * j], sc));
        }
        for (int j = 0; j < k; ++j) {
          float sc;
          ISST_TRY(score_of(fn[2 * j], fn[2 * j + 1], &sc));
          nxt[b].push_back({fn[2 * j], fn[2 * j + 1]});
          nxt_score[b].push_back(sc);
        }
        g.done = step + 1 >= follow->steps[b] && follow->done[b];
      } else {
        for (int rank = 0; rank < n_keep && static_cast<int>(nxt[b].size()) < k; ++rank) {
          ISST_CHECK(ci[rank] >= 0, "beam search ran out of finite candidates");
          const int par = ci[rank] / V, tok = ci[rank] % V;
          if (is_eos(tok)) {
            if (rank >= k) continue;                               // :100-103
            ISST_TRY(close(par, cs[rank]));
          } else {
            nxt[b].push_back({par, tok});
            nxt_score[b].push_back(cs[rank]);
          }
        }
        ISST_CHECK(static_cast<int>(nxt[b].size()) == k, "beam search: fewer than num_beams non-EOS candidates");
        // BeamHypotheses.is_done with early_stopping=False (transformers 4.47): the kept hypotheses cannot be beaten
        if (!g.done && static_cast<int>(g.hyps.size()) >= k) {
          const float highest = cs[0] / std::pow(static_cast<float>(gen_len), length_penalty);
          g.done = g.worst >= highest;
        }
      }
      g.steps = step + 1;
      if (trace && trace->next) {
        for (int j = 0; j < k; ++j) {
          trace->next[((static_cast<size_t>(b) * max_new + step) * k + j) * 2] = nxt[b][j].first;
          trace->next[((static_cast<size_t>(b) * max_new + step) * k + j) * 2 + 1] = nxt[b][j].second;
          if (trace->next_scores) trace->next_scores[(static_cast<size_t>(b) * max_new + step) * k + j] = nxt_score[b][j];
        }
      }
    }
    return 0;
  };

  int db_max_prefix = 0;
  double db_prefix_tokens = 0;
  // ---- step 0: candidates of the single live beam ----
  ISST_TRY(select(0, 1));
  // fork: every beam gets private tail pages; the partially filled last prompt page is copied into each
  for (int b = 0; b < n; ++b) {
    BeamGroupH& g = groups[b];
    StreamHost& s = ctx->streams[g.slot];
    const int slot1 = g.L1 < s.sys_len ? g.L1 : g.L1 - s.sys_len + s.ring_start;
    g.tail0 = slot1 / kPageTokens;
    g.partial = slot1 % kPageTokens != 0;
    g.n_priv = max_new > 1 ? ceil_div(slot1 % kPageTokens + max_new - 1, kPageTokens) : 0;
    ISST_CHECK(g.tail0 + g.n_priv <= ctx->pages_per_stream, "stream needs more pages than pages_per_stream");
    {
      // logical index of the first key that lives in a private page: the shared prefix is [0, lb)
      const int sb = g.tail0 * kPageTokens;
      const int lbnd = sb < s.sys_len ? sb : (sb < s.ring_start ? s.sys_len : sb - s.ring_start + s.sys_len);
      mb.host(o_khi)[b] = lbnd;
      mb.host(o_tpg)[b] = g.tail0;
      db_max_prefix = std::max(db_max_prefix, lbnd);
      db_prefix_tokens += lbnd;
    }
    ISST_CHECK(g.L1 + max_new - 1 <= c.max_kv_len, "stream KV length would exceed max_kv_len");
    g.rows.resize(k);
    for (int j = 0; j < k; ++j) {
      BeamRowH& r = g.rows[j];
      if (!g.done && max_new > 1) {
        ISST_TRY(beam_alloc_pages(ctx, r.priv[0], g.n_priv));
        ISST_TRY(beam_alloc_pages(ctx, r.priv[1], g.n_priv));
        if (g.partial) { pairs.push_back(s.pages[g.tail0]); pairs.push_back(r.priv[0][0]); }
      }
      r.gen.assign(1, nxt[b].empty() ? 0 : nxt[b][j].second);
      r.score = nxt[b].empty() ? 0.f : nxt_score[b][j];
    }
  }
  ISST_TRY(flush_pairs());

  // ---- decode steps on n * k rows ----
  LlmBatch db;
  db.n = R; db.M = R; db.max_T = 1; db.d_slots = mb.dev(o_iota); db.d_tok_base = mb.dev(o_iota); db.d_T = mb.dev(o_ones);
  db.d_last_row = mb.dev(o_iota); db.d_active = mb.dev(o_active); db.decode = true;
  db.group = k; db.d_key_hi = mb.dev(o_khi); db.d_tail_page = mb.dev(o_tpg); db.max_prefix = db_max_prefix;
  db.prefix_tokens = db_prefix_tokens;
  ISST_CUDA(cudaMemcpyAsync(ctx->d_meta + o_khi, ctx->h_meta + o_khi, (o_pairs - o_khi) * sizeof(int), cudaMemcpyHostToDevice, st));
  std::vector<KvView> views(R);
  for (int step = 1; step < max_new; ++step) {
    bool any = false;
    for (const BeamGroupH& g : groups) any = any || !g.done;
    if (!any) break;
    db.kv_tokens = 0; db.max_L = 1;
    for (int b = 0; b < n; ++b) {
      BeamGroupH& g = groups[b];
      const StreamHost& s = ctx->streams[g.slot];
      for (int j = 0; j < k; ++j) {
        const int r = b * k + j;
        BeamRowH& br = g.rows[j];
        // a finished sentence rides along inactive: it attends to its shared prefix only and appends nothing
        views[r] = g.done ? KvView{s.pages.data(), std::min(static_cast<int>(s.pages.size()), g.tail0 + 1), nullptr, 0, g.L1 - 1, s.sys_len, s.ring_start, s.evicted}
                          : KvView{s.pages.data(), g.tail0, br.priv[br.cur].data(), static_cast<int>(br.priv[br.cur].size()),
                                   g.L1 + step - 1, s.sys_len, s.ring_start, s.evicted};
        mb.host(o_active)[r] = g.done ? 0 : 1;
        mb.host(o_next)[r] = br.gen.back();
        reinterpret_cast<float*>(mb.host(o_score))[r] = br.score;
        // context ids of the row: prompt + the beam's own tokens (rebuilt every step: beams were reordered)
        int* cx = mb.host(o_ctx) + static_cast<size_t>(r) * ctx_cap;
        for (int t = 0; t < g.P; ++t) cx[t] = ids[g.ids_off + t];
        for (size_t t = 0; t < br.gen.size(); ++t) cx[g.P + t] = br.gen[t];
        mb.host(o_ctxlen)[r] = g.P + static_cast<int>(br.gen.size());
        if (!g.done) { db.kv_tokens += g.L1 + step; db.max_L = std::max(db.max_L, g.L1 + step); }
      }
    }
    ISST_TRY(upload_kv_views(ctx, st, R, views.data()));
    ISST_CUDA(cudaMemcpyAsync(ctx->d_meta + o_ctx, ctx->h_meta + o_ctx, (o_sup - o_ctx) * sizeof(int), cudaMemcpyHostToDevice, st));
    {
      ProfScope ps(ctx, st, P_EMBED, 0.0, static_cast<double>(R) * c.hidden * 4);
      ISST_CUDA(launch_k(ctx, embed_splice_kernel, dim3(R), dim3(128), 0, st, mb.dev(o_next), nullptr, ctx->embed, ctx->speech, ctx->lx, c.hidden));
      LAUNCH_CHECK(ctx);
    }
    ISST_TRY(llm_forward(ctx, st, db, false));
    ISST_TRY(select(step, k));
    // reorder (`_temporary_reorder_cache`, patch_hf.py:910-913): a beam that continues another beam copies that
    // beam's private tail into its own spare page set; beams that continue themselves keep their pages
    for (int b = 0; b < n; ++b) {
      BeamGroupH& g = groups[b];
      if (nxt[b].empty()) continue;
      std::vector<BeamRowH> old = g.rows;
      for (int j = 0; j < k; ++j) {
        const int par = nxt[b][j].first;
        BeamRowH& r = g.rows[j];                      // the page sets stay with the row index
        r.gen = old[par].gen;
        r.gen.push_back(nxt[b][j].second);
        r.score = nxt_score[b][j];
        if (par != j && !g.done) {                    // (after the last step finalize still hands back the parent's KV)
          for (int i = 0; i < g.n_priv; ++i) { pairs.push_back(old[par].priv[old[par].cur][i]); pairs.push_back(r.priv[1 - r.cur][i]); }
          r.cur = 1 - r.cur;
        }
      }
    }
    ISST_TRY(flush_pairs());
  }
  ISST_TRY(flush_pairs());

  // ---- finalize (patch_hf.py:159-275): open beams become hypotheses WITHOUT the KV of their last token ----
  ISST_CUDA(cudaStreamSynchronize(st));
  prof_flush(ctx);
  for (int b = 0; b < n; ++b) {
    BeamGroupH& g = groups[b];
    StreamHost& s = ctx->streams[g.slot];
    if (!g.done) {
      for (int j = 0; j < k; ++j) {
        BeamRowH& r = g.rows[j];
        BeamHypH h;
        h.tokens = r.gen;
        h.kv_extra = static_cast<int>(r.gen.size()) - 1;
        h.pages = r.priv[r.cur];
        r.priv[r.cur].clear();
        beam_hyp_add(ctx, g, k, length_penalty, std::move(h), r.score, static_cast<int>(r.gen.size()));
      }
    }
    for (BeamRowH& r : g.rows) { beam_free_pages(ctx, r.priv[0]); beam_free_pages(ctx, r.priv[1]); }
    int best = 0;
    for (int i = 1; i < static_cast<int>(g.hyps.size()); ++i)               // sorted(...).pop(): highest score, latest on ties
      if (g.hyps[i].score >= g.hyps[best].score) best = i;
    BeamHypH& w = g.hyps[best];
    int cnt = static_cast<int>(w.tokens.size());
    for (int t = 0; t < cnt; ++t) out_tokens[static_cast<size_t>(b) * (max_new + 1) + t] = w.tokens[t];
    if (cnt < max_new) out_tokens[static_cast<size_t>(b) * (max_new + 1) + cnt++] = gen->eos_token_ids[0];   // :262-264
    for (int t = cnt; t < max_new + 1; ++t) out_tokens[static_cast<size_t>(b) * (max_new + 1) + t] = -1;
    out_counts[b] = cnt;
    if (out_scores) out_scores[b] = w.score;
    if (trace && trace->steps) trace->steps[b] = g.steps;
    // KV hand-back (agents/infinisst.py:334-336): the stream continues from the best hypothesis' cache
    if (w.kv_extra > 0) {
      const int new_len = g.L1 + w.kv_extra;
      const int last_slot = new_len <= s.sys_len ? new_len - 1 : new_len - 1 - s.sys_len + s.ring_start;
      const int need = last_slot / kPageTokens + 1 - g.tail0;               // private pages that hold tokens
      for (size_t i = g.tail0; i < s.pages.size(); ++i) ctx->free_pages.push_back(s.pages[i]);   // replaced by the hypothesis' copy
      s.pages.resize(g.tail0);
      for (int i = 0; i < static_cast<int>(w.pages.size()); ++i) {
        if (i < need) s.pages.push_back(w.pages[i]);
        else ctx->free_pages.push_back(w.pages[i]);
      }
      w.pages.clear();
      s.kv_len = new_len;
    }
    for (BeamHypH& h : g.hyps) beam_free_pages(ctx, h.pages);
  }
  return 0;
}

int isst_kv_len(isst_ctx* ctx, int stream_id, int* len) {
  ISST_CHECK(ctx && len && stream_id >= 0 && stream_id < ctx->cfg.max_streams && ctx->streams[stream_id].open, "bad stream id");
  *len = ctx->streams[stream_id].kv_len;
  return 0;
}

int isst_enc_steps(isst_ctx* ctx, int stream_id, int* n_steps) {
  ISST_CHECK(ctx && n_steps && stream_id >= 0 && stream_id < ctx->cfg.max_streams && ctx->streams[stream_id].open, "bad stream id");
  *n_steps = ctx->streams[stream_id].enc_prefix;
  return 0;
}

int isst_kv_evict(isst_ctx* ctx, int stream_id, int keep_prefix, int drop_upto) {
  ISST_CHECK(ctx && stream_id >= 0 && stream_id < ctx->cfg.max_streams && ctx->streams[stream_id].open, "bad stream id");
  StreamHost& s = ctx->streams[stream_id];
  ISST_CHECK(keep_prefix == s.sys_len, "keep_prefix must equal the stream's pinned prefix (gen.pin_prefix of its first prefill)");
  ISST_CHECK(drop_upto >= keep_prefix && drop_upto <= s.kv_len, "drop range out of bounds");
  const int n_drop = drop_upto - keep_prefix;
  if (n_drop == 0) return 0;
  s.ring_start += n_drop;
  s.kv_len -= n_drop;
  s.evicted += n_drop;
  // release ring pages that fell completely behind ring_start
  const int sys_pages = ceil_div(s.sys_len, kPageTokens);
  int freeable = s.ring_start / kPageTokens - sys_pages;
  freeable = std::min(freeable, static_cast<int>(s.pages.size()) - sys_pages);
  if (freeable > 0) {
    for (int i = 0; i < freeable; ++i) ctx->free_pages.push_back(s.pages[sys_pages + i]);
    s.pages.erase(s.pages.begin() + sys_pages, s.pages.begin() + sys_pages + freeable);
    s.ring_start -= freeable * kPageTokens;
  }
  return 0;
}

int isst_debug_enable(isst_ctx* ctx, int on) {
  ISST_CHECK(ctx, "null ctx");
  ISST_CUDA(cudaSetDevice(ctx->device));
  ctx->debug = (on & 1) != 0;
  if ((on & 2) && !ctx->gemm_dbg) {           // bit 1: per-CTA phase stamps of stream-K GEMM launches ("gemm_stamps" tap)
    ISST_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->gemm_dbg), 4096 * 8 * sizeof(unsigned long long)));
    ISST_CUDA(cudaMemset(ctx->gemm_dbg, 0, 4096 * 8 * sizeof(unsigned long long)));
  }
  if (!(on & 2) && ctx->gemm_dbg) { cudaFree(ctx->gemm_dbg); ctx->gemm_dbg = nullptr; }
  return 0;
}

int isst_debug_read(isst_ctx* ctx, const char* name, void* dst_host, int64_t max_bytes, int64_t* n_bytes) {
  ISST_CHECK(ctx && name && n_bytes, "null argument");
  ISST_CUDA(cudaSetDevice(ctx->device));
  if (std::string(name) == "gemm_stamps" && ctx->gemm_dbg) {
    *n_bytes = 4096 * 8 * sizeof(unsigned long long);
    if (dst_host) {
      ISST_CHECK(max_bytes >= *n_bytes, "destination too small");
      ISST_CUDA(cudaDeviceSynchronize());
      ISST_CUDA(cudaMemcpy(dst_host, ctx->gemm_dbg, *n_bytes, cudaMemcpyDeviceToHost));
    }
    return 0;
  }
  auto it = ctx->taps.find(name);
  if (it == ctx->taps.end()) return set_error(std::string("no such tap: ") + name);
  *n_bytes = static_cast<int64_t>(it->second.second);
  if (dst_host) {
    ISST_CHECK(max_bytes >= *n_bytes, "destination too small");
    ISST_CUDA(cudaDeviceSynchronize());
    ISST_CUDA(cudaMemcpy(dst_host, it->second.first, it->second.second, cudaMemcpyDeviceToHost));
  }
  return 0;
}

int isst_profile_enable(isst_ctx* ctx, int on) {
  ISST_CHECK(ctx, "null ctx");
  ISST_CUDA(cudaSetDevice(ctx->device));
  ISST_CUDA(cudaDeviceSynchronize());
  prof_flush(ctx);
  ctx->prof.on = on != 0;
  return 0;
}

int isst_profile_reset(isst_ctx* ctx) {
  ISST_CHECK(ctx, "null ctx");
  ISST_CUDA(cudaSetDevice(ctx->device));
  ISST_CUDA(cudaDeviceSynchronize());
  prof_flush(ctx);
  for (int i = 0; i < P_NCAT; ++i) { ctx->prof.ms[i] = 0; ctx->prof.flops[i] = 0; ctx->prof.bytes[i] = 0; ctx->prof.n[i] = 0; }
  return 0;
}

int isst_profile_read(isst_ctx* ctx, int index, char* name, int name_cap, int64_t* launches, double* ms, double* flops,
                      double* bytes) {
  ISST_CHECK(ctx && name && launches && ms && flops && bytes, "null argument");
  if (index < 0 || index >= P_NCAT) return 1;   // end of list
  ISST_CUDA(cudaSetDevice(ctx->device));
  ISST_CUDA(cudaDeviceSynchronize());
  prof_flush(ctx);
  snprintf(name, name_cap, "%s", kProfNames[index]);
  *launches = ctx->prof.n[index]; *ms = ctx->prof.ms[index]; *flops = ctx->prof.flops[index]; *bytes = ctx->prof.bytes[index];
  return 0;
}

int isst_debug_shift_positions(isst_ctx* ctx, int stream_id, int64_t delta) {
  ISST_CHECK(ctx && stream_id >= 0 && stream_id < ctx->cfg.max_streams && ctx->streams[stream_id].open, "bad stream id");
  ISST_CHECK(delta >= 0 && ctx->streams[stream_id].evicted + delta < 2000000000LL, "delta out of range");
  ctx->streams[stream_id].evicted += delta;
  return 0;
}

int64_t isst_launch_count(isst_ctx* ctx) { return ctx ? ctx->launches : -1; }

int isst_path_count(isst_ctx* ctx, const char* name, int64_t* count) {
  ISST_CHECK(ctx && name && count, "null argument");
  auto it = ctx->paths.find(name);
  *count = it == ctx->paths.end() ? 0 : it->second;
  return 0;
}

int isst_debug_option(isst_ctx* ctx, const char* key_c, int value) {
  ISST_CHECK(ctx && key_c, "null argument");
  const std::string key(key_c);
  if (key == "pdl") ctx->pdl = value != 0;
  else if (key == "decode_splits") ctx->opt_dec_splits = value;
  else if (key == "decode_chain") ctx->opt_chain = value != 0;
  else if (key == "chain_fold") ctx->opt_fold = value != 0;
  else if (key == "tap_llm_layers") ctx->tap_llm_layers = value != 0;
  else if (key == "defer_splits_as_chain") ctx->opt_defer_as_chain = value != 0;
  else if (key == "gemm_pair") ctx->opt_pair = value != 0;
  else if (key == "prefill_l2_ahead") ctx->opt_pa_l2_ahead = value;
  else return set_error("unknown option: " + key);
  return 0;
}
int isst_pages_free(isst_ctx* ctx) { return ctx ? static_cast<int>(ctx->free_pages.size()) : -1; }

int isst_op_gemm(isst_ctx* ctx, const void* act_bf16, const void* w_bf16, int M, int N, int K, const float* bias,
                 int act_gelu, const void* resid_bf16, int dual, void* out, int out_f32, int force_swap,
                 int force_splits, void* cuda_stream) {
  ISST_CHECK(ctx && act_bf16 && w_bf16 && out, "null argument");
  ISST_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  Weight2D w;
  w.ptr = const_cast<bf16*>(static_cast<const bf16*>(w_bf16));
  w.rows = dual ? 2 * N : N;
  w.K = K;
  ISST_TRY(make_weight_map(w));
  Epilogue e;
  e.bias = bias; e.act = act_gelu; e.resid = static_cast<const bf16*>(resid_bf16); e.ldr = N; e.out_f32 = out_f32;
  e.dual = dual; e.dual_off = dual ? N : 0;
  return gemm(ctx, st, plain_view(static_cast<const bf16*>(act_bf16), M, K), w, N, out, N, 0, e, force_swap, force_splits);
}

int isst_op_decode_attention_bench(isst_ctx* ctx, int n, int L, int iters, float* ms_per_iter, void* cuda_stream) {
  ISST_CHECK(ctx && ms_per_iter && n >= 1 && n <= ctx->cfg.max_batch && iters >= 1, "bad argument");
  ISST_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const isst_config& c = ctx->cfg;
  ISST_CHECK(L + 1 <= c.max_kv_len, "L too large");
  // temporary streams with L tokens of pseudo-random KV in layer 0..layers-1 pools
  std::vector<int> slots(n);
  for (int b = 0; b < n; ++b) {
    ISST_TRY(isst_stream_open(ctx, &slots[b]));
    ISST_TRY(ensure_capacity(ctx, slots[b], L + 1, 0));
    ctx->streams[slots[b]].kv_len = L;
  }
  ISST_TRY(upload_stream_tables(ctx, st, n, slots.data()));
  fill_pattern_kernel<<<1024, 256, 0, st>>>(ctx->kv_pool, static_cast<long long>(ctx->kv_layer_elems) * c.layers, 17u);
  const int QKV = (c.heads + 2 * c.kv_heads) * c.head_dim;
  fill_pattern_kernel<<<64, 256, 0, st>>>(ctx->lqkv, static_cast<long long>(n) * QKV, 3u);
  fill_pattern_kernel<<<64, 256, 0, st>>>(ctx->lq_sys, static_cast<long long>(n) * c.heads * c.head_dim, 5u);
  MetaBuilder mb{ctx};
  const size_t o_slots = mb.alloc(n);
  for (int b = 0; b < n; ++b) mb.host(o_slots)[b] = b;
  ISST_CUDA(cudaMemcpyAsync(ctx->d_meta, ctx->h_meta, mb.used * sizeof(int), cudaMemcpyHostToDevice, st));
  const float scale_log2 = 1.4426950408889634f / std::sqrt(static_cast<float>(c.head_dim));
  cudaEvent_t e0, e1;
  ISST_CUDA(cudaEventCreate(&e0)); ISST_CUDA(cudaEventCreate(&e1));
  const int splits = decode_splits_for(ctx, n, L + 1);
  for (int it = -2; it < iters; ++it) {
    if (it == 0) ISST_CUDA(cudaEventRecord(e0, st));
    // walk the layers so consecutive launches read different (cold) KV: total footprint = layers * n * L * 4 KB
    const int layer = ((it % c.layers) + c.layers) % c.layers;
    ISST_TRY(launch_decode_attention(ctx, st, ctx->lqkv, paged_kv(ctx, layer), mb.dev(o_slots), n, splits, scale_log2));
  }
  ISST_CUDA(cudaEventRecord(e1, st));
  ISST_CUDA(cudaStreamSynchronize(st));
  float ms = 0.f;
  ISST_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *ms_per_iter = ms / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  for (int b = 0; b < n; ++b) ISST_TRY(isst_stream_close(ctx, slots[b]));
  return 0;
}

}  // extern "C"
