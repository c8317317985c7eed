// C ABI of the decode step (include/mtx_b200.h): engine object, TMA descriptors, kernel
// launch plumbing (programmatic dependent launch, CUDA graphs).  Host code only; the kernels
// live in the .cuh files next to this one.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <tuple>

#include "../../include/mtx_b200.h"
#include "attention.cuh"
#include "attention_wide.cuh"
#include "gemm_rows.cuh"
#include "gemm_umma.cuh"
#include "gemma3_kernels.cuh"
#include "paged_kv.cuh"
#include "prefill_attention.cuh"
#include "sampling.cuh"
#include "step_kernels.cuh"
#include "step_persistent.cuh"

using namespace mtx;

#define MTX_STR2(x) #x
#define MTX_STR(x) MTX_STR2(x)

namespace {

thread_local std::string g_error;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

#define MTX_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t err__ = (expr);                                                                 \
    if (err__ != cudaSuccess) return fail(MTX_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(err__)); \
  } while (0)

#define MTX_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != MTX_OK) return rc__; \
  } while (0)

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// ---- TMA descriptors ---------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 matrix [outer, inner] (inner contiguous), box [box_outer, 64] with the 128-byte swizzle.
int make_map(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint32_t box_outer) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail(MTX_ERR_CUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (inner * 2) % 16 != 0)
    return fail(MTX_ERR_ARG, "TMA needs a 16-byte aligned base and row pitch");
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {inner * 2};
  const cuuint32_t box[2] = {64, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return fail(MTX_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(rc));
  return MTX_OK;
}

// uint8 matrix [outer, 64] (one int8 cache row of a kv head per matrix row), box [64, 64] with the 64-byte swizzle.
int make_map_u8(CUtensorMap* m, const void* base, uint64_t outer) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return fail(MTX_ERR_CUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(MTX_ERR_ARG, "TMA needs a 16-byte aligned base");
  const cuuint64_t dims[2] = {64, outer};
  const cuuint64_t strides[1] = {64};
  const cuuint32_t box[2] = {64, 64};
  const cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return fail(MTX_ERR_CUDA, "cuTensorMapEncodeTiled (uint8) failed with CUresult %d", int(rc));
  return MTX_OK;
}

// ---- launch helper -------------------------------------------------------------------------

// Optional per-kernel timing: when a sink is installed every launch is bracketed by CUDA
// events on the launching stream and attributed to the current kernel class.
enum KernelClass { KC_PREPARE = 0, KC_RMSNORM, KC_QKV, KC_ATTENTION, KC_OUTPROJ, KC_MLP_UP, KC_MLP_DOWN, KC_LOGITS, KC_FINALIZE, KC_PERSISTENT, KC_COUNT };
struct ProfileSink {
  struct Rec { int cls; cudaEvent_t a, b; };
  std::vector<Rec> recs;
};
thread_local ProfileSink* g_profile = nullptr;
thread_local int g_class = KC_PREPARE;
long long* g_trace = nullptr;  // debug timeline buffer (device), see mtx_debug_set_trace

bool use_pdl() {
  static int v = env_int("MTX_PDL", 1);
  return v != 0;
}

thread_local int g_cluster_y = 1;  // cluster size along grid.y for the next launch (split-K GEMM)
thread_local bool g_cooperative = false;  // next launch is cooperative: all its CTAs are co-resident or the launch fails

template <typename... KArgs, typename... Args>
int launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_cooperative) {
    attr[na].id = cudaLaunchAttributeCooperative;
    attr[na].val.cooperative = 1;
    ++na;
  } else if (use_pdl()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (g_cluster_y > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = unsigned(g_cluster_y);
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  ProfileSink::Rec rec;
  if (g_profile != nullptr) {
    rec.cls = g_class;
    MTX_CUDA(cudaEventCreate(&rec.a));
    MTX_CUDA(cudaEventCreate(&rec.b));
    MTX_CUDA(cudaEventRecord(rec.a, stream));
  }
  MTX_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  if (g_profile != nullptr) {
    MTX_CUDA(cudaEventRecord(rec.b, stream));
    g_profile->recs.push_back(rec);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MTX_OK;
}

// Attention work items the per-kernel path aims for: a multiple of the SM count, so that the (at most 3 per SM) persistent CTAs
// of decode_attn_kernel each walk several items and finish together (with one item per (row, kv head) pair, batch 256 leaves
// 1024 whole-sequence items for 444 CTAs: two or three each, a 30 % imbalance).
int attn_target_items(int num_sms) { return num_sms * env_int("MTX_ATTN_ITEMS_PER_SM", 1); }

int round_rows(int rows) {
  int t = 16;
  while (t < rows) t *= 2;
  return t;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct GemmPlan {
  int n_tiles, splits, stages;
  size_t smem;
};

// K splits = cluster size: a power of two <= 16 (8 is the portable maximum, 16 needs the
// non-portable opt-in), chosen so that n_tiles * splits is close to one CTA per SM.
GemmPlan plan_gemm(int n, int k, int r_tile, int num_sms, int epi, int forced_splits = 0) {
  GemmPlan g;
  g.n_tiles = (n + kTileN - 1) / kTileN;
  const int kb = k / kBlockK;
  int splits = forced_splits;
  if (splits <= 0) {
    const int target = env_int("MTX_GEMM_TARGET_CTAS", num_sms);
    const int max_split = env_int("MTX_GEMM_MAX_SPLIT", 16);
    const double want = double(target) / g.n_tiles;
    splits = 1;
    while (splits * 2 <= max_split && splits * 2 <= kb && splits * 2 <= r_tile && double(splits) * 1.42 < want) splits *= 2;
  }
  g.splits = splits;
  const int stage_bytes = kWTileBytes + r_tile * kBlockK * 2;
  const int budget = (r_tile <= 128 ? 100 : 200) * 1024;
  int stages = budget / stage_bytes;
  const int max_kb = (kb + splits - 1) / splits;
  if (stages > max_kb) stages = max_kb;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 1) stages = 1;
  g.stages = stages;
  g.smem = gemm_smem_bytes(stages, r_tile, splits, epi);
  return g;
}

template <int EPI>
int launch_gemm(const CUtensorMap& tw, const CUtensorMap& tx, GemmParams p, const EpiArgs& e, const GemmPlan& g, cudaStream_t st) {
  static bool attr_set = false;  // per template instance
  if (!attr_set) {
    MTX_CUDA(cudaFuncSetAttribute(gemm_umma_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    MTX_CUDA(cudaFuncSetAttribute(gemm_umma_kernel<EPI>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_set = true;
  }
  p.splits = g.splits;
  p.stages = g.stages;
  p.trace = g_trace;
  g_cluster_y = g.splits;
  const int rc = launch(gemm_umma_kernel<EPI>, dim3(g.n_tiles, g.splits), dim3(kGemmThreads), g.smem, st, tw, tx, p, e);
  g_cluster_y = 1;
  return rc;
}

// ---- steps of 65..256 rows: gemm_rows.cuh ------------------------------------------------------
struct RowsPlan {
  int n_tiles, splits, stages, grid_x, row_blocks, r_tile_cta;
  size_t smem;
};

RowsPlan plan_rows(int n, int k, int r_tile, int num_sms, int epi, int forced_splits = 0) {
  RowsPlan g;
  g.n_tiles = (n + kTileN - 1) / kTileN;
  const int kb = k / kBlockK;
  int splits = forced_splits;
  if (splits <= 0) {
    splits = 1;
    // few tiles: cut K over a cluster of up to 8 CTAs (the portable maximum; two such clusters fit a GPC) until the
    // grid covers most of the SMs.  Many tiles: one persistent CTA per SM walks them with double-buffered accumulators.
    // (a split unit covers one 128-row block, and two units fit one SM: up to 1.25 units per SM)
    if (epi != EPI_LOGITS && g.n_tiles * 2 <= num_sms) {
      const int max_split = env_int("MTX_ROWS_MAX_SPLIT", 8);
      const int zb = r_tile / 128;
      while (splits * 2 <= max_split && splits * 2 <= kb && g.n_tiles * zb * splits * 2 <= num_sms + num_sms / 4) splits *= 2;
    }
  }
  g.splits = splits;
  g.row_blocks = splits > 1 ? r_tile / 128 : 1;
  g.r_tile_cta = splits > 1 ? 128 : r_tile;
  r_tile = g.r_tile_cta;
  const int stage_bytes = kWTileBytes + r_tile * kBlockK * 2;
  int stages = ((splits > 1 ? 96 : 192) * 1024) / stage_bytes;
  const int max_kb = splits > 1 ? (kb + splits - 1) / splits : kb * 2;
  if (stages > max_kb) stages = max_kb;
  if (stages > kRowsMaxStages) stages = kRowsMaxStages;
  if (stages < 1) stages = 1;
  g.stages = stages;
  g.grid_x = splits > 1 ? g.n_tiles : (g.n_tiles < num_sms ? g.n_tiles : num_sms);
  g.smem = rows_smem_bytes(stages, r_tile, splits);
  return g;
}

template <int EPI>
int launch_rows(const CUtensorMap& tw, const CUtensorMap& tx, GemmParams p, const EpiArgs& e, const RowsPlan& g, cudaStream_t st) {
  static bool attr_set = false;  // per template instance
  if (!attr_set) {
    MTX_CUDA(cudaFuncSetAttribute(gemm_rows_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  p.splits = g.splits;
  p.stages = g.stages;
  p.r_tile = g.r_tile_cta;
  p.trace = g_trace;
  g_cluster_y = g.splits;
  const int rc = launch(gemm_rows_kernel<EPI>, dim3(g.grid_x, g.splits, g.row_blocks), dim3(kRowsThreads), g.smem, st, tw, tx, p, e);
  g_cluster_y = 1;
  return rc;
}

bool use_rows_kernel(int r_tile) {
  static int on = env_int("MTX_ROWS_KERNEL", 1);
  return on != 0 && r_tile >= 128;
}

// (plan_gemm / launch_gemm are declared above)
// out = EPI(x[rows, k] . w[n, k]^T) on whichever tcgen05 GEMM serves the row count (gemm_umma.cuh up to 64 rows, gemm_rows.cuh above)
template <int EPI>
int single_gemm(const void* x, const void* w, int rows, int n, int k, EpiArgs ea, cudaStream_t st) {
  const int r_tile = round_rows(rows);
  CUtensorMap tw, tx;
  MTX_TRY(make_map(&tw, w, k, n, kTileN));
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.n = n;
  p.k = k;
  p.rows = rows;
  p.r_tile = r_tile;
  if (use_rows_kernel(r_tile)) {
    const RowsPlan pl = plan_rows(n, k, r_tile, 148, EPI);
    MTX_TRY(make_map(&tx, x, k, r_tile, pl.splits > 1 ? 128 : r_tile));
    return launch_rows<EPI>(tw, tx, p, ea, pl, st);
  }
  MTX_TRY(make_map(&tx, x, k, r_tile, r_tile));
  return launch_gemm<EPI>(tw, tx, p, ea, plan_gemm(n, k, r_tile, 148, EPI), st);
}

struct XMaps {
  bool built = false;
  CUtensorMap n, attn, act, x, h;
};

int log2_tile(int r_tile) {
  int i = 0;
  while ((16 << i) < r_tile) ++i;
  return i;
}

}  // namespace

// =============================================================================================

struct mtx_engine {
  mtx_model_config cfg;
  int num_sms = 148;
  int max_r_tile = 16;
  int qkv_n = 0;
  int attn_max_chunks = 0;
  bool bound = false;
  mtx_weights w{};
  mtx_decode_state s{};
  // workspace carve-up
  size_t ws_bytes = 0;
  bf16 *x = nullptr, *h = nullptr, *n = nullptr, *q = nullptr, *attn = nullptr, *act = nullptr;
  float *attn_part_o = nullptr, *attn_part_ml = nullptr;
  int* attn_tickets = nullptr;
  RowDesc rd{};
  float* rope_timescale = nullptr;
  float* rope_timescale_w = nullptr;  // gemma3: the local layers' RoPE base
  bf16* qkv_tmp = nullptr;            // gemma3: [max_r_tile, qkv_n] the QKV projection before q/k norm, RoPE and the cache append
  bf16 *kv_tmp_k = nullptr, *kv_tmp_v = nullptr;  // kv_quant 2: [max_r_tile, Hkv, D] the step's keys / values before quantisation
  float *part_score = nullptr, *part_raw = nullptr, *part_max = nullptr, *part_sum = nullptr;
  int* part_idx = nullptr;
  std::vector<float> rope_timescale_host, rope_timescale_w_host;
  // descriptors
  std::vector<CUtensorMap> tm_wqkv, tm_wo, tm_w01, tm_wout;
  CUtensorMap tm_logits, tm_k, tm_v;
  CUtensorMap tm_kq, tm_vq;  // kv_quant: the int8 decode cache (tm_k / tm_v then address the bf16 prefill staging plane)
  CUtensorMap tm_kp, tm_vp;  // attention=paged: the page pools as [L * Hkv * num_pages * tokens_per_page, D], box = min(page, 64) rows
  CUtensorMap tm_all_wqkv, tm_all_wo, tm_all_w01, tm_all_wout;  // all layers stacked: row = layer * N + n
  // persistent step kernel (step_persistent.cuh)
  int pk_ctas = 0;  // CTAs of the persistent grid (0 = unavailable)
  unsigned int* grid_bar = nullptr;
  PkTable* pk_tables = nullptr;
  float *pk_part_ws = nullptr, *pk_ss_x = nullptr, *pk_ss_h = nullptr, *pk_attn_part_o = nullptr;
  int* cand_counters = nullptr;  // [0] nucleus rows truncated to the candidates, [1] commit ticket
  // all-SM top-k / nucleus sampler (sampling.cuh, par_*): per-row state and the per-slice histograms
  float *par_M = nullptr, *par_Z = nullptr;
  unsigned long long *par_target = nullptr, *par_above = nullptr, *par_hist = nullptr;
  uint32_t* par_prefix = nullptr;
  int* par_found = nullptr;
  int* par_cand_count = nullptr;
  int* par_eq = nullptr;
  float *rows_ss_x = nullptr, *rows_ss_h = nullptr;  // gemm_rows.cuh fused RMSNorm statistics: [E/128 rounded up][max_r_tile]
  int *pk_tile_prefix = nullptr, *pk_attn_info = nullptr;
  XMaps xmaps[5];
  // sampling
  int last_prefill_rows = 1;   // rows of the last sampling prefill chunk (the draw's row inside its block)
  unsigned prefill_draws = 0;  // prefill calls that sampled a first token: each draws its noise from its own row namespace
  int strategy = MTX_SAMPLE_GREEDY, top_k = 0;
  float nucleus_p = 0.f, temperature = 1.f;
  // graphs
  cudaStream_t cap_stream = nullptr;
  std::map<int, cudaGraphExec_t> graphs;
  // mtx_decode_step_host: the step's graph with the host copies as nodes, keyed by (rows, tokens, result, log-prob pointers)
  struct HostKey {
    int rows;
    const void *tok, *res, *lp;
    bool operator<(const HostKey& o) const {
      return std::tie(rows, tok, res, lp) < std::tie(o.rows, o.tok, o.res, o.lp);
    }
  };
  std::map<HostKey, cudaGraphExec_t> host_graphs;
};

namespace {

struct WsLayout {
  size_t x, h, n, q, attn, act, attn_part_o, attn_part_ml, attn_tickets;
  size_t token, pos, plane, write_row, len0, ring_first, ring_len, rope_cs, work_items, work_count, rope_timescale;
  size_t skip_w, iota, tmp_row, kv_tmp_k, kv_tmp_v;
  size_t len0_w, ring_first_w, ring_len_w, rope_cs_w, work_items_w, work_count_w, rope_timescale_w, qkv_tmp;
  size_t part_score, part_idx, part_raw, part_max, part_sum, grid_bar;
  size_t pk_tables, pk_part_ws, pk_ss_x, pk_ss_h, pk_attn_part_o, pk_tile_prefix, pk_attn_info;
  size_t rows_ss_x, rows_ss_h, cand_counters;
  size_t par_M, par_Z, par_target, par_above, par_prefix, par_found, par_hist, par_eq, par_cand_count;
  size_t total;
};

// mtx_model_config.kv_quant: 1 / 2 = int8 with kv_quant_axis dkv / heads_and_dkv, 3 / 4 = the same with float8_e4m3fn bytes
int kvq_axis(const mtx_model_config& c) { return c.kv_quant == 0 ? 0 : (c.kv_quant - 1) % 2 + 1; }
bool kvq_fp8(const mtx_model_config& c) { return c.kv_quant >= 3; }
// attention=paged: the decode cache is the page pools; like an int8 engine, prefill then writes ONE bf16 staging plane
bool is_paged(const mtx_model_config& c) { return c.paged_num_pages > 0; }
bool staged_prefill(const mtx_model_config& c) { return c.kv_quant != 0 || is_paged(c); }

WsLayout layout_workspace(const mtx_engine* e) {
  const mtx_model_config& c = e->cfg;
  const size_t rt = e->max_r_tile;
  WsLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t at = off;
    off = align_up(off + bytes, 1024);
    return at;
  };
  const int G = c.num_q_heads / c.num_kv_heads;
  L.x = take(rt * c.emb_dim * 2);
  L.h = take(rt * c.emb_dim * 2);
  L.n = take(rt * c.emb_dim * 2);
  L.q = take(rt * c.num_q_heads * c.head_dim * 2);
  L.attn = take(rt * c.num_q_heads * c.head_dim * 2);
  L.act = take(rt * c.mlp_dim * 2);
  L.attn_part_o = take(size_t(c.max_rows) * c.num_kv_heads * e->attn_max_chunks * G * c.head_dim * 4);
  L.attn_part_ml = take(size_t(c.max_rows) * c.num_kv_heads * e->attn_max_chunks * G * 2 * 4);
  L.attn_tickets = take(size_t(c.max_rows) * c.num_kv_heads * 4);
  L.token = take(rt * 4);
  L.pos = take(rt * 4);
  L.plane = take(rt * 4);
  L.write_row = take(rt * 4);
  L.len0 = take(rt * 4);
  L.ring_first = take(rt * 4);
  L.ring_len = take(rt * 4);
  L.rope_cs = take(rt * (c.head_dim / 2) * 8);
  L.work_items = take(size_t(c.max_rows) * e->attn_max_chunks * 4);
  L.work_count = take(4);
  L.rope_timescale = take((c.head_dim / 2) * 4);
  if (kvq_axis(c) == 2) {
    L.iota = take(rt * 4);
    L.tmp_row = take(rt * 4);
    L.kv_tmp_k = take(rt * c.num_kv_heads * c.head_dim * 2);
    L.kv_tmp_v = take(rt * c.num_kv_heads * c.head_dim * 2);
  }
  if (c.decoder_block == 1) {
    L.len0_w = take(rt * 4);
    L.ring_first_w = take(rt * 4);
    L.ring_len_w = take(rt * 4);
    L.skip_w = take(rt * 4);
    L.rope_cs_w = take(rt * (c.head_dim / 2) * 8);
    L.work_items_w = take(size_t(c.max_rows) * e->attn_max_chunks * 4);
    L.work_count_w = take(4);
    L.rope_timescale_w = take((c.head_dim / 2) * 4);
    L.qkv_tmp = take(rt * size_t(e->qkv_n) * 2);
  }
  size_t vt = (c.vocab_size + kTileN - 1) / kTileN;
  if (vt < size_t(par_slices(c.vocab_size)) * kParWarps) vt = size_t(par_slices(c.vocab_size)) * kParWarps;  // par_scan_kernel's pieces
  L.part_score = take(size_t(c.max_rows) * vt * 4);
  L.part_idx = take(size_t(c.max_rows) * vt * 4);
  L.part_raw = take(size_t(c.max_rows) * vt * 4);
  L.part_max = take(size_t(c.max_rows) * vt * 4);
  L.part_sum = take(size_t(c.max_rows) * vt * 4);
  L.grid_bar = take(64);
  {
    const size_t pk_rows = c.max_rows < kPkMaxRTile ? c.max_rows : kPkMaxRTile;
    const size_t pairs = pk_rows * c.num_kv_heads;
    const size_t ss_tiles = (c.emb_dim + 127) / 128;
    L.pk_tables = take(size_t(2) * e->num_sms * sizeof(PkTable));  // [0] many rows, [1] few rows
    L.pk_part_ws = take(size_t(e->num_sms) * 4 * kPkSlotFloats * 4);
    L.pk_ss_x = take(size_t(kPkMaxRTile) * ss_tiles * 4);
    L.pk_ss_h = take(size_t(kPkMaxRTile) * ss_tiles * 4);
    L.pk_attn_part_o = take(pairs * kPkMaxParts * ((G * c.head_dim + 2 * G + 3) / 4 * 4) * 4);
    L.pk_tile_prefix = take((pk_rows + 1) * 4);
    L.pk_attn_info = take(64);
    L.rows_ss_x = take(rt * ss_tiles * 4);
    L.rows_ss_h = take(rt * ss_tiles * 4);
    L.cand_counters = take(64);
    L.par_M = take(size_t(c.max_rows) * 4);
    L.par_cand_count = take(size_t(c.max_rows) * 4);
    L.par_Z = take(size_t(c.max_rows) * 4);
    L.par_target = take(size_t(c.max_rows) * 8);
    L.par_above = take(size_t(c.max_rows) * 8);
    L.par_prefix = take(size_t(c.max_rows) * 4);
    L.par_found = take(size_t(c.max_rows) * 4);
    L.par_hist = take(size_t(c.max_rows) * kParBins * 8);
    L.par_eq = take(size_t(kParMaxSlices) * kParWarps * c.max_rows * 4);
  }
  L.total = off;
  return L;
}

int get_xmaps(mtx_engine* e, int r_tile, XMaps** out) {
  XMaps& m = e->xmaps[log2_tile(r_tile)];
  if (!m.built) {
    const mtx_model_config& c = e->cfg;
    MTX_TRY(make_map(&m.n, e->n, c.emb_dim, e->max_r_tile, r_tile));
    MTX_TRY(make_map(&m.attn, e->attn, uint64_t(c.num_q_heads) * c.head_dim, e->max_r_tile, r_tile));
    MTX_TRY(make_map(&m.act, e->act, c.mlp_dim, e->max_r_tile, r_tile));
    MTX_TRY(make_map(&m.x, e->x, c.emb_dim, e->max_r_tile, r_tile));
    MTX_TRY(make_map(&m.h, e->h, c.emb_dim, e->max_r_tile, r_tile));
    m.built = true;
  }
  *out = &m;
  return MTX_OK;
}

// gemma3.py:36-48: layers 0..4 of every six use sliding-window attention (and the local RoPE base), the sixth is global.
bool layer_is_local(const mtx_engine* e, int layer) { return e->cfg.decoder_block == 1 && layer % 6 != 5; }

int launch_attention(mtx_engine* e, int layer, int rows, cudaStream_t st, int mode = 0) {
  const mtx_model_config& c = e->cfg;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.q = e->q;
  p.out = e->attn;
  p.plane = e->rd.plane;
  p.len0 = e->rd.len0;
  p.ring_first = e->rd.ring_first;
  p.ring_len = e->rd.ring_len;
  p.work_items = e->rd.work_items;
  p.work_count = e->rd.work_count;
  if (layer_is_local(e, layer)) {  // the window's two cache sub-ranges (attention.cuh, AttnParams.skip0)
    const int R = c.max_target_len - c.max_prefill_len, W = c.sliding_window;
    p.len0 = e->rd.len0_w;
    p.ring_first = e->rd.ring_first_w;
    p.ring_len = e->rd.ring_len_w;
    p.work_items = e->rd.work_items_w;
    p.work_count = e->rd.work_count_w;
    if (mode == 0) {
      p.skip0 = c.max_prefill_len > W ? c.max_prefill_len - W : 0;
      p.ring_off = R > W ? R - W : 0;
      p.ring_size = R - p.ring_off;
    } else {
      p.skip0_rows = e->rd.skip_w;  // prompt positions as decode rows: the window is by position
      p.ring_size = R - (R > W ? R - W : 0);  // (no ring rows in prefill; the tile counts were made for this ring size)
    }
  }
  p.part_o = e->attn_part_o;
  p.part_ml = e->attn_part_ml;
  p.tickets = e->attn_tickets;
  p.rows = rows;
  p.hq = c.num_q_heads;
  p.hkv = c.num_kv_heads;
  p.P = c.max_prefill_len;
  p.T = c.max_target_len;
  p.tiles_per_item = attn_tiles_per_item(rows, c.num_kv_heads, c.max_prefill_len, c.max_target_len, attn_target_items(e->num_sms));
  p.max_chunks = e->attn_max_chunks;
  p.plane_base = layer * c.num_slots;
  p.softcap = c.attn_softcap;
  p.trace = g_trace;
  const size_t smem = attn_smem_bytes(c.head_dim, c.num_q_heads / c.num_kv_heads);
  const int ctas_per_sm = c.head_dim == 64 ? 3 : 1;
  int grid = rows * c.num_kv_heads * attn_max_chunks(c.max_prefill_len, c.max_target_len, p.tiles_per_item);
  const int cap = e->num_sms * ctas_per_sm;
  if (grid > cap) grid = cap;
  if (is_paged(c) && mode == 0) {  // decode rows read the page pools through the group's row of the page map (row r = group r)
    p.page_map = e->s.page_map;
    p.tokens_per_page = c.paged_tokens_per_page;
    p.num_pages = c.paged_num_pages;
    p.max_pages = c.paged_max_pages_per_group;
    p.page_row_base = (long long)layer * c.num_kv_heads * c.paged_num_pages * c.paged_tokens_per_page;
    if (c.head_dim == 64) return launch(decode_attn_kernel<64>, dim3(grid), dim3(kAttnThreads), smem, st, e->tm_kp, e->tm_vp, p);
    return launch(decode_attn_kernel<128>, dim3(grid), dim3(kAttnThreads), smem, st, e->tm_kp, e->tm_vp, p);
  }
  if (c.kv_quant) {
    const size_t smem_q = attn_q8_smem_bytes(c.num_q_heads / c.num_kv_heads);
    int grid_q = rows * c.num_kv_heads * attn_max_chunks(c.max_prefill_len, c.max_target_len, p.tiles_per_item);
    if (grid_q > e->num_sms * 4) grid_q = e->num_sms * 4;
    // groups of up to 8 heads: the transposed tile arithmetic (half the MMAs, fewer registers: four CTAs per SM)
    const bool tr = c.num_q_heads / c.num_kv_heads <= 8 && env_int("MTX_Q8_TRANSPOSED", 1) != 0;
    auto kern = kvq_fp8(c) ? (tr ? decode_attn_q8_kernel<true, true> : decode_attn_q8_kernel<true, false>)
                           : (tr ? decode_attn_q8_kernel<false, true> : decode_attn_q8_kernel<false, false>);
    return launch(kern, dim3(grid_q), dim3(kAttnThreads), smem_q, st, e->tm_kq, e->tm_vq, p, (const float*)e->s.k_scale,
                  (const float*)e->s.v_scale);
  }
  if (c.head_dim == 256) {  // wide heads: the transposed kernel (attention_wide.cuh), one CTA of three warps per SM
    static bool attr_set = false;
    if (!attr_set) {
      MTX_CUDA(cudaFuncSetAttribute(decode_attn_wide_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(attn_wide_smem_bytes(256))));
      attr_set = true;
    }
    int grid_w = rows * c.num_kv_heads * attn_max_chunks(c.max_prefill_len, c.max_target_len, p.tiles_per_item);
    if (grid_w > e->num_sms) grid_w = e->num_sms;
    return launch(decode_attn_wide_kernel<256>, dim3(grid_w), dim3(kWideThreads), attn_wide_smem_bytes(256), st, e->tm_k, e->tm_v, p);
  }
  if (c.head_dim == 64) return launch(decode_attn_kernel<64>, dim3(grid), dim3(kAttnThreads), smem, st, e->tm_k, e->tm_v, p);
  return launch(decode_attn_kernel<128>, dim3(grid), dim3(kAttnThreads), smem, st, e->tm_k, e->tm_v, p);
}

// Causal attention of one prefill chunk (prefill_attention.cuh): the chunk's positions share the K/V tiles.
bool use_prefill_attention(const mtx_engine* e) {
  static int on = env_int("MTX_PREFILL_ATTENTION", 1);
  (void)e;
  return on != 0;
}

int launch_prefill_attention(mtx_engine* e, int layer, int rows, int start_pos, int slot, cudaStream_t st) {
  const mtx_model_config& c = e->cfg;
  const int planes = staged_prefill(c) ? 1 : c.num_slots;  // an int8 / paged engine prefills into its single bf16 staging plane
  PrefillAttnParams p;
  memset(&p, 0, sizeof(p));
  p.q = e->q;
  p.out = e->attn;
  p.rows = rows;
  p.start_pos = start_pos;
  p.hq = c.num_q_heads;
  p.hkv = c.num_kv_heads;
  p.T = c.max_target_len;
  p.plane_row0 = (layer * planes + slot) * c.num_kv_heads * c.max_target_len;
  p.softcap = c.attn_softcap;
  p.window = layer_is_local(e, layer) ? c.sliding_window : 0;
  const dim3 grid(c.num_q_heads, (rows + kPfQRows - 1) / kPfQRows), block(kPfWarps * 32);
  const size_t smem = prefill_attn_smem_bytes(c.head_dim);
  static bool attr_set = false;
  if (!attr_set) {
    MTX_CUDA(cudaFuncSetAttribute(prefill_attn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(prefill_attn_smem_bytes(64))));
    MTX_CUDA(cudaFuncSetAttribute(prefill_attn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(prefill_attn_smem_bytes(128))));
    attr_set = true;
  }
  if (c.head_dim == 64) return launch(prefill_attn_kernel<64>, grid, block, smem, st, e->tm_k, e->tm_v, p);
  return launch(prefill_attn_kernel<128>, grid, block, smem, st, e->tm_k, e->tm_v, p);
}

// ---- persistent step kernel: work tables and launch -------------------------------------------

__global__ void fill_u32_kernel(uint32_t* dst, size_t n, uint32_t value) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) dst[i] = value;
}


// Deals the (weight tile, k-block) units of one GEMM phase to the CTAs in contiguous equal ranges (stream-K).
// A CTA gets at least `q_min` k-blocks so that no tile is shared by more than kPkMaxSplit CTAs.
bool pk_fill_phase(std::vector<PkTable>& tabs, int ph, int n, int k, bool allow_whole_tile) {
  const int n_ctas = int(tabs.size());
  const int n_tiles = (n + kTileN - 1) / kTileN, kbt = k / kBlockK;
  const long long total = (long long)n_tiles * kbt;
  int q_min = (kbt + kPkMaxSplit - 2) / (kPkMaxSplit - 1);
  if (q_min < 2) q_min = 2;
  // short reductions: at most 8 contributors, so a row slice is summed with one round of loads
  if (kbt <= 32 && q_min < (kbt + 6) / 7) q_min = (kbt + 6) / 7;
  if (q_min > kbt) q_min = kbt;
  long long active = total / q_min;
  if (active > n_ctas) active = n_ctas;
  if (active < 1) active = 1;
  std::vector<int> first(n_tiles, -1), count(n_tiles, 0);
  // Small matrices (at least two CTAs per tile): every tile is split evenly over S consecutive CTAs, one unit per
  // CTA, so nobody straddles two tiles and has to sit through two exchanges.
  int uniform_s = 0;
  if (n_tiles * 2 <= n_ctas) {
    uniform_s = n_ctas / n_tiles;
    if (uniform_s > kPkMaxSplit) uniform_s = kPkMaxSplit;
    if (kbt <= 32 && uniform_s > 8) uniform_s = 8;
    if (uniform_s > kbt / 2) uniform_s = kbt / 2 > 0 ? kbt / 2 : 1;
  }
  // Mid-sized matrices (at least half as many tiles as CTAs): one whole tile per CTA.  The CTA streams more bytes
  // than with stream-K, but nothing is exchanged: the accumulator goes straight from TMEM through shared memory to
  // the epilogue, and the exchange of a shared tile costs more (about 12 k-blocks' worth of streaming time) than
  // the extra streaming.  CTAs without a tile run ahead into the next phase's weights.  (The exchange cost grows with
  // the number of rows, so steps of few rows use the second table set, built without this rule.)
  if (allow_whole_tile && uniform_s == 0 && n_tiles <= n_ctas && kbt <= total / n_ctas + env_int("MTX_PK_WHOLE_TILE_SLACK", 12)) uniform_s = 1;
  // MLP up with between one and two CTAs per tile: CTA t < n_tiles OWNS tile t and reduces its first k_own k-blocks;
  // the other CTAs (helpers) share the remaining k-blocks of all tiles in contiguous ranges and hand their partials
  // over through the exchange.  Helpers get less work than owners, so their partials are in L2 by the time an owner
  // has finished its own stream: the owner's accumulator never leaves the chip, it waits for nobody, and it streams
  // k_own instead of all k-blocks (a single SM pulls ~40 GB/s, which bounds the whole-tile variant).
  const int k_own_env = env_int("MTX_PK_UP_OWNER_KB", -1);  // -1 = automatic (an equal share of the k-blocks), 0 = off (one whole tile per CTA)
  if (ph == PK_UP && allow_whole_tile && k_own_env != 0 && n_tiles < n_ctas && 2 * n_tiles > n_ctas && kbt >= 8) {
    int k_own = k_own_env > 0 ? k_own_env : int((total + n_ctas - 1) / n_ctas);
    if (k_own > kbt - 1) k_own = kbt - 1;
    const int helpers = n_ctas - n_tiles, kh = kbt - k_own;
    const long long total_h = (long long)n_tiles * kh;
    for (int c = 0; c < n_ctas; ++c) {
      tabs[c].n_units[ph] = 0;
      tabs[c].kbs[ph] = 0;
    }
    for (int i = 0; i < helpers; ++i) {
      PkTable& t = tabs[n_tiles + i];
      long long lo = i * total_h / helpers, hi = (i + 1) * total_h / helpers;
      while (lo < hi) {
        const int tile = int(lo / kh);
        long long end = (long long)(tile + 1) * kh;
        if (end > hi) end = hi;
        if (t.n_units[ph] >= kPkMaxUnits) return false;
        PkUnit& u = t.u[ph][t.n_units[ph]++];
        u.tile = tile;
        u.kb0 = k_own + int(lo - (long long)tile * kh);
        u.kb1 = k_own + int(end - (long long)tile * kh);
        u.c_first = tile;
        u.S = 0;
        t.kbs[ph] += u.kb1 - u.kb0;
        if (first[tile] < 0) first[tile] = n_tiles + i;
        ++count[tile];
        lo = end;
      }
    }
    for (int tile = 0; tile < n_tiles; ++tile) {
      PkTable& t = tabs[tile];
      PkUnit& u = t.u[ph][t.n_units[ph]++];
      u.tile = tile;
      u.kb0 = 0;
      u.kb1 = k_own;
      u.c_first = first[tile];
      u.S = -count[tile];
      t.kbs[ph] = k_own;
      if (count[tile] < 1 || count[tile] > kPkMaxSplit) return false;
    }
    return true;
  }
  for (int c = 0; c < n_ctas; ++c) {
    PkTable& t = tabs[c];
    t.n_units[ph] = 0;
    t.kbs[ph] = 0;
    if (uniform_s > 0) {
      const int tile = c / uniform_s, si = c % uniform_s;
      if (tile >= n_tiles) continue;
      PkUnit& u = t.u[ph][t.n_units[ph]++];
      u.tile = tile;
      u.kb0 = si * kbt / uniform_s;
      u.kb1 = (si + 1) * kbt / uniform_s;
      t.kbs[ph] = u.kb1 - u.kb0;
      if (first[tile] < 0) first[tile] = c;
      ++count[tile];
      continue;
    }
    if (c >= active) continue;
    long long lo = c * total / active, hi = (c + 1) * total / active;
    while (lo < hi) {
      const int tile = int(lo / kbt);
      long long end = (long long)(tile + 1) * kbt;
      if (end > hi) end = hi;
      if (t.n_units[ph] >= kPkMaxUnits) return false;
      PkUnit& u = t.u[ph][t.n_units[ph]++];
      u.tile = tile;
      u.kb0 = int(lo - (long long)tile * kbt);
      u.kb1 = int(end - (long long)tile * kbt);
      t.kbs[ph] += u.kb1 - u.kb0;
      if (first[tile] < 0) first[tile] = c;
      ++count[tile];
      lo = end;
    }
  }
  for (int c = 0; c < n_ctas; ++c)
    for (int i = 0; i < tabs[c].n_units[ph]; ++i) {
      PkUnit& u = tabs[c].u[ph][i];
      u.c_first = first[u.tile];
      u.S = count[u.tile];
      if (u.S > kPkMaxSplit) return false;
    }
  return true;
}

// Barrier slots of the step trace: 200 (the layout tools/ were written for) or what the layer count needs.
int pk_trace_bars(int layers) { return 2 + 5 * layers + 2 > 200 ? 2 + 5 * layers + 2 : 200; }

bool pk_usable(const mtx_engine* e, int rows) {
  if (e->pk_ctas <= 0 || round_rows(rows) > kPkMaxRTile || e->cfg.decoder_block != 0) return false;
  // (a quantised cache with one scale per token and kv head is quantised by the QKV epilogue; heads_and_dkv needs the maximum over
  //  heads that sit in different CTAs: per-kernel path)
  if (e->cfg.kv_quant && (kvq_axis(e->cfg) != 1 || env_int("MTX_PK_KVQ", 1) == 0)) return false;
  // (paged: a 64-row tile is one slice of a page or two whole pages; smaller pages take the per-kernel path)
  if (is_paged(e->cfg) && (e->cfg.paged_tokens_per_page < 32 || env_int("MTX_PK_PAGED", 1) == 0)) return false;
  if (e->cfg.num_q_heads / e->cfg.num_kv_heads > 8) return false;  // the attention MMA carries the group's heads in its 8 columns
  // an attention warp's tile list holds kPkAttnListMax entries: bound the worst case (every context full)
  const mtx_model_config& c = e->cfg;
  const long long worst = (long long)rows * c.num_kv_heads * attn_max_tiles(c.max_prefill_len, c.max_target_len);
  return worst <= (long long)e->pk_ctas * kPkAttnWarps * kPkAttnListMax;
}

int launch_persistent(mtx_engine* e, int rows, const XMaps& xm, const EpiArgs& logits_epi, cudaStream_t st) {
  const mtx_model_config& c = e->cfg;
  PkParams p;
  memset(&p, 0, sizeof(p));
  p.L = c.num_layers;
  p.E = c.emb_dim;
  p.HD = c.num_q_heads * c.head_dim;
  p.M = c.mlp_dim;
  p.qkv_n = e->qkv_n;
  p.V = c.vocab_size;
  p.hq = c.num_q_heads;
  p.hkv = c.num_kv_heads;
  p.d = c.head_dim;
  p.t_alloc = c.max_target_len;
  p.num_slots = c.num_slots;
  p.rows = rows;
  p.r_tile = round_rows(rows);
  p.P = c.max_prefill_len;
  p.T = c.max_target_len;
  p.eps = c.rms_eps;
  p.softcap = c.attn_softcap;
  p.x = e->x;
  p.h = e->h;
  p.n = e->n;
  p.q = e->q;
  p.attn = e->attn;
  p.act = e->act;
  p.embedding = static_cast<const bf16*>(e->w.embedding);
  p.attn_norm = static_cast<const bf16*>(e->w.attn_norm);
  p.mlp_norm = static_cast<const bf16*>(e->w.mlp_norm);
  p.final_norm = static_cast<const bf16*>(e->w.final_norm);
  p.k_cache = static_cast<bf16*>(e->s.k_cache);
  p.v_cache = static_cast<bf16*>(e->s.v_cache);
  p.kv_layer_elems = (long long)c.num_slots * c.num_kv_heads * c.max_target_len * c.head_dim;
  p.kv_layer_rows = c.num_slots * c.num_kv_heads * c.max_target_len;
  if (is_paged(c)) {  // the page pools: one "plane" of num_pages * tokens_per_page rows per kv head and layer (prepare_rows_kernel)
    p.k_cache = static_cast<bf16*>(e->s.k_pages);
    p.v_cache = static_cast<bf16*>(e->s.v_pages);
    p.t_alloc = c.paged_num_pages * c.paged_tokens_per_page;
    p.kv_layer_rows = c.num_kv_heads * p.t_alloc;
    p.kv_layer_elems = (long long)p.kv_layer_rows * c.head_dim;
    p.page_map = e->s.page_map;
    p.page_tokens = c.paged_tokens_per_page;
    p.num_pages = c.paged_num_pages;
    p.max_pages = c.paged_max_pages_per_group;
  }
  p.token = e->rd.token;
  p.plane = e->rd.plane;
  p.write_row = e->rd.write_row;
  p.len0 = e->rd.len0;
  p.ring_first = e->rd.ring_first;
  p.ring_len = e->rd.ring_len;
  p.rope_cs = e->rd.rope_cs;
  p.tile_prefix = e->pk_tile_prefix;
  p.attn_info = e->pk_attn_info;
  p.attn_part_o = e->pk_attn_part_o;
  p.part_ws = e->pk_part_ws;
  p.ss_x = e->pk_ss_x;
  p.ss_h = e->pk_ss_h;
  p.tables = e->pk_tables + (rows <= env_int("MTX_PK_FEW_ROWS", 0) ? e->pk_ctas : 0);
  p.logits = logits_epi;
  p.grid_bar = e->grid_bar;
  p.trace = g_trace;
  p.trace_bars = pk_trace_bars(c.num_layers);
  p.variant = env_int("MTX_PK_VARIANT", 0);
  p.attn_skew = float(env_int("MTX_PK_ATTN_SKEW", 0)) * 0.01f;  // percent
  p.fold = c.norm_scales_folded ? 1 : 0;
  // The kernel spins on a software grid barrier.  MTX_PK_COOPERATIVE=1 launches it cooperatively, so the driver guarantees that
  // all CTAs are resident together (or fails the launch) whatever else shares the device: the setting for a GPU shared with
  // other streams / processes.  Default off: a dedicated GPU (one process per GPU, one engine per process, as MaxEngine is
  // deployed) passes the occupancy check of mtx_engine_bind once, and the cooperative launch costs programmatic dependent
  // launch against the prepare kernel and 45 us per CUDA-graph replay end to end (profiles/r2k_cooperative_ab.txt).
  g_cooperative = env_int("MTX_PK_COOPERATIVE", 0) != 0;
  int rc;
  if (c.kv_quant) {  // kv_quant_axis dkv only (pk_usable): the byte caches, quantising QKV epilogue, fp16 attention over byte tiles
    p.kq_cache = static_cast<uint8_t*>(e->s.kq_cache);
    p.vq_cache = static_cast<uint8_t*>(e->s.vq_cache);
    p.k_scale = e->s.k_scale;
    p.v_scale = e->s.v_scale;
    auto kern = kvq_fp8(c) ? step_persistent_kernel<2> : step_persistent_kernel<1>;
    rc = launch(kern, dim3(e->pk_ctas), dim3(kPkThreads), pk_smem_bytes(), st, e->tm_all_wqkv, e->tm_all_wo, e->tm_all_w01, e->tm_all_wout,
                e->tm_logits, xm.x, xm.attn, xm.h, xm.act, xm.n, e->tm_kq, e->tm_vq, p);
  } else {
    rc = launch(step_persistent_kernel<0>, dim3(e->pk_ctas), dim3(kPkThreads), pk_smem_bytes(), st, e->tm_all_wqkv, e->tm_all_wo, e->tm_all_w01,
                e->tm_all_wout, e->tm_logits, xm.x, xm.attn, xm.h, xm.act, xm.n, is_paged(c) ? e->tm_kp : e->tm_k,
                is_paged(c) ? e->tm_vp : e->tm_v, p);
  }
  g_cooperative = false;
  return rc;
}

// The kernels of one step, in stream order.  mode 0 = decode, 1 = prefill chunk.
// Noise row of a prefill draw: outside the rows of any decode step (< 256), one block of 256 rows per prefill call.
int prefill_noise_row(const mtx_engine* e) { return 0x40000000 + int(e->prefill_draws % 0x100000u) * 256; }

int enqueue_step(mtx_engine* e, int mode, int rows, const int32_t* chunk_tokens, int start_pos, int slot, int want_logits,
                 int32_t* first_token, float* prefill_logits, cudaStream_t st, float* cand_out = nullptr, float* first_log_prob = nullptr) {
  const mtx_model_config& c = e->cfg;
  // (an int8 cache is appended to by the row-major GEMM's epilogue only: steps of any size run it, padded to 128 rows)
  const bool gemma3 = c.decoder_block == 1;
  const bool mega = mode == 0 && want_logits && pk_usable(e, rows);  // the persistent step kernel (up to 64 rows)
  const int r_tile = (c.kv_quant || gemma3) && !mega && round_rows(rows) < 128 ? 128 : round_rows(rows);
  XMaps* xm;
  MTX_TRY(get_xmaps(e, r_tile, &xm));
  // prefill (mode 1) of an int8 engine writes the ONE bf16 staging plane that k_cache / v_cache then are
  const int planes = (staged_prefill(c) && mode == 1) ? 1 : c.num_slots;
  if (staged_prefill(c) && mode == 1) slot = 0;
  const int E = c.emb_dim, HD = c.num_q_heads * c.head_dim, M = c.mlp_dim, L = c.num_layers;
  const size_t kv_layer = size_t(planes) * c.num_kv_heads * c.max_target_len * c.head_dim;
  const size_t kvq_layer = size_t(c.num_slots) * c.num_kv_heads * c.max_target_len * c.head_dim;  // bytes of one layer of the int8 cache
  const size_t kvs_layer = size_t(c.num_slots) * c.num_kv_heads * c.max_target_len;                // its scales

  PrepareArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.tokens = e->s.tokens;
  pa.next_pos = e->s.next_pos;
  pa.prefill_len = e->s.prefill_len;
  pa.ar_lengths = e->s.ar_lengths;
  pa.ar_index = e->s.ar_index;
  pa.chunk_tokens = chunk_tokens;
  pa.start_pos = start_pos;
  pa.slot = slot;
  pa.mode = mode;
  pa.rows = rows;
  pa.P = c.max_prefill_len;
  pa.T = c.max_target_len;
  pa.D = c.head_dim;
  pa.tiles_per_item = attn_tiles_per_item(rows, c.num_kv_heads, c.max_prefill_len, c.max_target_len, attn_target_items(e->num_sms));
  pa.emb_rows = c.embedding_rows;
  pa.rope_timescale = e->rope_timescale;
  pa.window = gemma3 ? c.sliding_window : 0;
  pa.page_lengths = (is_paged(c) && mode == 0) ? e->s.page_lengths : nullptr;
  pa.active_page = e->s.active_page;
  pa.active_pos = e->s.active_page_pos;
  pa.tokens_per_page = c.paged_tokens_per_page;
  if (is_paged(c) && c.paged_device_state && mode == 0) {  // update_decode_pages at the head of the step (maxengine.py:847-849)
    pa.page_update = 1;
    pa.page_state.page_status = e->s.page_status;
    pa.page_state.page_map = e->s.page_map;
    pa.page_state.num_pages_used = e->s.num_pages_used;
    pa.page_state.sequence_lengths = e->s.page_lengths;
    pa.page_state.active_page = e->s.active_page;
    pa.page_state.has_active_page = e->s.has_active_page;
    pa.page_state.active_page_position = e->s.active_page_pos;
    pa.page_state.num_pages = c.paged_num_pages;
    pa.page_state.groups = c.num_slots;
    pa.page_state.max_pages_per_group = c.paged_max_pages_per_group;
    pa.page_state.tokens_per_page = c.paged_tokens_per_page;
  }
  pa.rope_timescale_w = e->rope_timescale_w;
  if (mega) {
    pa.grid_bar = e->grid_bar;
    pa.tile_prefix = e->pk_tile_prefix;
    pa.attn_info = e->pk_attn_info;
    pa.hkv = c.num_kv_heads;
    pa.pk_ctas = e->pk_ctas;
    pa.pk_max_parts = kPkMaxParts;
    pa.pk_warps = kPkAttnWarps;
    pa.pk_pair_mode_tiles = env_int("MTX_PK_PAIR_MODE_TILES", 8);
  }
  g_class = KC_PREPARE;
  MTX_TRY(launch(prepare_rows_kernel, dim3(1), dim3(256), 0, st, pa, e->rd));
  g_class = KC_RMSNORM;

  GemmParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.rows = rows;
  gp.r_tile = r_tile;
  const GemmPlan plan_logits = plan_gemm(c.vocab_size, c.emb_dim, r_tile, e->num_sms, EPI_LOGITS, 1);
  const bool rows_k = (c.kv_quant || gemma3) ? true : use_rows_kernel(r_tile);  // 65..256 rows: the row-major tensor-core GEMM (gemm_rows.cuh)
  XMaps* xm128 = xm;  // split-K units of that kernel load 128-row activation boxes
  if (rows_k) MTX_TRY(get_xmaps(e, 128, &xm128));
  auto xmap = [&](const RowsPlan& pl, CUtensorMap XMaps::*m) -> const CUtensorMap& { return pl.splits > 1 ? xm128->*m : xm->*m; };
  // 65..256 rows with the norm scales folded into the weights: RMSNorm is fused into the GEMMs around it (the residual
  // epilogues leave the row statistics, the consuming epilogues apply rstd; the per-kernel rmsnorm launches disappear)
  const bool fused_norm = rows_k && c.norm_scales_folded && (gemma3 || env_int("MTX_ROWS_FUSED_NORM", 1) != 0);
  const int ss_tiles = (E + 127) / 128;
  auto ss_args = [&](EpiArgs& ea, const float* in, float* out) {
    ea.ss_in = fused_norm ? in : nullptr;
    ea.ss_out = fused_norm ? out : nullptr;
    ea.ss_tiles = ss_tiles;
    ea.ss_pitch = e->max_r_tile;
    ea.ss_dim = E;
    ea.ss_eps = c.rms_eps;
  };
  if (gemma3) {
    // ---- layers/gemma3.py:62-197 on the row-major GEMMs; the element-wise steps between them in gemma3_kernels.cuh ----
    MTX_TRY(launch(embed_gather_ss_kernel, dim3(rows), dim3(128), 0, st, (const int*)e->rd.token, static_cast<const bf16*>(e->w.embedding), e->x,
                   e->rows_ss_x, ss_tiles, e->max_r_tile, E));
    const RowsPlan pl_qkv = plan_rows(e->qkv_n, E, r_tile, e->num_sms, EPI_STORE_BF16);
    const RowsPlan pl_o = plan_rows(E, HD, r_tile, e->num_sms, EPI_STORE_BF16);
    const RowsPlan pl_up = plan_rows(2 * M, E, r_tile, e->num_sms, EPI_SWIGLU);
    const RowsPlan pl_down = plan_rows(E, M, r_tile, e->num_sms, EPI_STORE_BF16);
    for (int l = 0; l < L; ++l) {
      const bool local = layer_is_local(e, l);
      EpiArgs ea;
      // q, k, v = dense(rms_norm(x))  (the pre-norm's scale is folded into wqkv, its rstd applied to the accumulator)
      memset(&ea, 0, sizeof(ea));
      ea.out = e->qkv_tmp;
      ea.ld_out = e->qkv_n;
      ss_args(ea, e->rows_ss_x, nullptr);
      gp.n = e->qkv_n;
      gp.k = E;
      g_class = KC_QKV;
      MTX_TRY(launch_rows<EPI_STORE_BF16>(e->tm_wqkv[l], xmap(pl_qkv, &XMaps::x), gp, ea, pl_qkv, st));
      QkNormArgs qa;
      memset(&qa, 0, sizeof(qa));
      qa.qkv = e->qkv_tmp;
      qa.q_scale = static_cast<const bf16*>(e->w.q_norm) + size_t(l) * c.head_dim;
      qa.k_scale = static_cast<const bf16*>(e->w.k_norm) + size_t(l) * c.head_dim;
      qa.rope_cs = local ? e->rd.rope_cs_w : e->rd.rope_cs;
      qa.plane = e->rd.plane;
      qa.write_row = e->rd.write_row;
      qa.q_out = e->q;
      qa.k_cache = static_cast<bf16*>(e->s.k_cache) + kv_layer * l;
      qa.v_cache = static_cast<bf16*>(e->s.v_cache) + kv_layer * l;
      qa.hq = c.num_q_heads;
      qa.hkv = c.num_kv_heads;
      qa.d = c.head_dim;
      qa.t_alloc = c.max_target_len;
      qa.eps = c.rms_eps;
      qa.q_scalar = c.query_scalar;
      MTX_TRY(launch(qk_norm_rope_append_kernel, dim3(rows, c.num_q_heads + 2 * c.num_kv_heads), dim3(c.head_dim / 2), 0, st, qa));

      g_class = KC_ATTENTION;
      // A row whose window holds no valid key (a prompt shorter than max_prefill_len - window while the ring's window is still
      // empty) gets no work item: its attention output is defined as zero.  (The reference softmaxes the all-masked scores, i.e.
      // averages every allocated cache row, valid or not: nothing a caller can rely on.)
      if (mode == 0 && local) MTX_CUDA(cudaMemsetAsync(e->attn, 0, size_t(rows) * HD * 2, st));
      if (mode == 1 && c.head_dim != 256) MTX_TRY(launch_prefill_attention(e, l, rows, start_pos, slot, st));
      else MTX_TRY(launch_attention(e, l, rows, st, mode));  // (head_dim 256: prompt positions run as decode rows)

      // h = x + rms_norm(dense(attn))
      memset(&ea, 0, sizeof(ea));
      ea.out = e->n;
      ea.ld_out = E;
      ss_args(ea, nullptr, nullptr);
      gp.n = E;
      gp.k = HD;
      g_class = KC_OUTPROJ;
      MTX_TRY(launch_rows<EPI_STORE_BF16>(e->tm_wo[l], xmap(pl_o, &XMaps::attn), gp, ea, pl_o, st));
      g_class = KC_RMSNORM;
      MTX_TRY(launch(post_norm_residual_kernel, dim3(rows), dim3(128), 0, st, (const bf16*)e->n, (const bf16*)e->x,
                     static_cast<const bf16*>(e->w.post_attn_norm) + size_t(l) * E, e->h, e->rows_ss_h, e->max_r_tile, E, c.rms_eps));

      // x = h + rms_norm(dense(gelu(dense(n2, wi_0)) * dense(n2, wi_1))),  n2 = rms_norm(h) (scale folded into w01)
      memset(&ea, 0, sizeof(ea));
      ea.out = e->act;
      ea.ld_out = M;
      ea.act_gelu = 1;
      ss_args(ea, e->rows_ss_h, nullptr);
      gp.n = 2 * M;
      gp.k = E;
      g_class = KC_MLP_UP;
      MTX_TRY(launch_rows<EPI_SWIGLU>(e->tm_w01[l], xmap(pl_up, &XMaps::h), gp, ea, pl_up, st));
      memset(&ea, 0, sizeof(ea));
      ea.out = e->n;
      ea.ld_out = E;
      ss_args(ea, nullptr, nullptr);
      gp.n = E;
      gp.k = M;
      g_class = KC_MLP_DOWN;
      MTX_TRY(launch_rows<EPI_STORE_BF16>(e->tm_wout[l], xmap(pl_down, &XMaps::act), gp, ea, pl_down, st));
      g_class = KC_RMSNORM;
      MTX_TRY(launch(post_norm_residual_kernel, dim3(rows), dim3(128), 0, st, (const bf16*)e->n, (const bf16*)e->h,
                     static_cast<const bf16*>(e->w.post_ffw_norm) + size_t(l) * E, e->x, e->rows_ss_x, e->max_r_tile, E, c.rms_eps));
    }
    if (want_logits)
      MTX_TRY(launch(rmsnorm_kernel<false>, dim3(rows), dim3(128), 0, st, (const bf16*)e->x, (const int*)nullptr, (const bf16*)nullptr,
                     static_cast<const bf16*>(e->w.final_norm), (bf16*)nullptr, e->n, E, c.rms_eps));
  } else if (!mega) {
  const bf16* attn_norm = static_cast<const bf16*>(e->w.attn_norm);
  const bf16* mlp_norm = static_cast<const bf16*>(e->w.mlp_norm);
  if (fused_norm)
    MTX_TRY(launch(embed_gather_ss_kernel, dim3(rows), dim3(128), 0, st, (const int*)e->rd.token, static_cast<const bf16*>(e->w.embedding), e->x,
                   e->rows_ss_x, ss_tiles, e->max_r_tile, E));
  else
  MTX_TRY(launch(rmsnorm_kernel<true>, dim3(rows), dim3(128), 0, st, (const bf16*)nullptr, (const int*)e->rd.token,
                 static_cast<const bf16*>(e->w.embedding), attn_norm, e->x, e->n, E, c.rms_eps));

  const GemmPlan plan_qkv = plan_gemm(e->qkv_n, E, r_tile, e->num_sms, EPI_QKV_ROPE);
  const GemmPlan plan_o = plan_gemm(E, HD, r_tile, e->num_sms, EPI_RESIDUAL);
  const GemmPlan plan_up = plan_gemm(2 * M, E, r_tile, e->num_sms, EPI_SWIGLU);
  const GemmPlan plan_down = plan_gemm(E, M, r_tile, e->num_sms, EPI_RESIDUAL);

  for (int l = 0; l < L; ++l) {
    EpiArgs ea;
    memset(&ea, 0, sizeof(ea));
    ea.q_out = e->q;
    ea.k_cache = static_cast<bf16*>(e->s.k_cache) + kv_layer * l;
    ea.v_cache = static_cast<bf16*>(e->s.v_cache) + kv_layer * l;
    ea.plane = e->rd.plane;
    ea.write_row = e->rd.write_row;
    ea.rope_cs = e->rd.rope_cs;
    ea.hq = c.num_q_heads;
    ea.hkv = c.num_kv_heads;
    ea.d = c.head_dim;
    ea.t_alloc = c.max_target_len;
    ea.kv_fp8 = kvq_fp8(c) ? 1 : 0;
    if (kvq_axis(c) == 1 && mode == 0) {
      ea.kq_cache = static_cast<uint8_t*>(e->s.kq_cache) + kvq_layer * l;
      ea.vq_cache = static_cast<uint8_t*>(e->s.vq_cache) + kvq_layer * l;
      ea.k_scale = e->s.k_scale + kvs_layer * l;
      ea.v_scale = e->s.v_scale + kvs_layer * l;
    } else if (is_paged(c) && mode == 0) {
      // update_decode_step_pages (paged_attention.py:446-471): the layer's pool is one plane of num_pages * tokens_per_page rows
      // per kv head, the row descriptors address (active page, position) in it (prepare_rows_kernel)
      const size_t pool_layer = size_t(c.num_kv_heads) * c.paged_num_pages * c.paged_tokens_per_page * c.head_dim;
      ea.k_cache = static_cast<bf16*>(e->s.k_pages) + pool_layer * l;
      ea.v_cache = static_cast<bf16*>(e->s.v_pages) + pool_layer * l;
      ea.t_alloc = c.paged_num_pages * c.paged_tokens_per_page;
    } else if (kvq_axis(c) == 2 && mode == 0) {
      // one scale per token over all kv heads (kv_quant_axis heads_and_dkv): the epilogue leaves the rotated keys / values as
      // bf16 in [rows, Hkv, D] scratch matrices (plane = row, write row = 0), kv_quant_rows_kernel quantises and appends
      ea.k_cache = e->kv_tmp_k;
      ea.v_cache = e->kv_tmp_v;
      ea.plane = e->rd.iota;
      ea.write_row = e->rd.tmp_row;
      ea.t_alloc = 1;
    }
    gp.n = e->qkv_n;
    gp.k = E;
    g_class = KC_QKV;
    ss_args(ea, e->rows_ss_x, nullptr);
    if (rows_k) {
      const RowsPlan pl = plan_rows(e->qkv_n, E, r_tile, e->num_sms, EPI_QKV_ROPE);
      MTX_TRY(launch_rows<EPI_QKV_ROPE>(e->tm_wqkv[l], xmap(pl, fused_norm ? &XMaps::x : &XMaps::n), gp, ea, pl, st));
    }
    else MTX_TRY(launch_gemm<EPI_QKV_ROPE>(e->tm_wqkv[l], xm->n, gp, ea, plan_qkv, st));
    if (kvq_axis(c) == 2 && mode == 0) {
      KvQuantRowsArgs ka;
      memset(&ka, 0, sizeof(ka));
      ka.k_tmp = e->kv_tmp_k;
      ka.v_tmp = e->kv_tmp_v;
      ka.plane = e->rd.plane;
      ka.write_row = e->rd.write_row;
      ka.kq_cache = static_cast<uint8_t*>(e->s.kq_cache) + kvq_layer * l;
      ka.vq_cache = static_cast<uint8_t*>(e->s.vq_cache) + kvq_layer * l;
      ka.k_scale = e->s.k_scale + kvs_layer * l;
      ka.v_scale = e->s.v_scale + kvs_layer * l;
      ka.hkv = c.num_kv_heads;
      ka.t_alloc = c.max_target_len;
      ka.fp8 = kvq_fp8(c) ? 1 : 0;
      MTX_TRY(launch(kv_quant_rows_kernel, dim3(rows, 2), dim3(c.num_kv_heads * 32), 0, st, ka));
    }

    g_class = KC_ATTENTION;
    if (mode == 1 && (use_prefill_attention(e) || c.kv_quant)) MTX_TRY(launch_prefill_attention(e, l, rows, start_pos, slot, st));
    else MTX_TRY(launch_attention(e, l, rows, st));

    memset(&ea, 0, sizeof(ea));
    ea.out = e->h;
    ea.resid = e->x;
    ea.ld_out = E;
    gp.n = E;
    gp.k = HD;
    g_class = KC_OUTPROJ;
    ss_args(ea, nullptr, e->rows_ss_h);
    if (rows_k) {
      const RowsPlan pl = plan_rows(E, HD, r_tile, e->num_sms, EPI_RESIDUAL);
      MTX_TRY(launch_rows<EPI_RESIDUAL>(e->tm_wo[l], xmap(pl, &XMaps::attn), gp, ea, pl, st));
    }
    else MTX_TRY(launch_gemm<EPI_RESIDUAL>(e->tm_wo[l], xm->attn, gp, ea, plan_o, st));

    g_class = KC_RMSNORM;
    if (!fused_norm)
    MTX_TRY(launch(rmsnorm_kernel<false>, dim3(rows), dim3(128), 0, st, (const bf16*)e->h, (const int*)nullptr,
                   (const bf16*)nullptr, mlp_norm + size_t(l) * E, (bf16*)nullptr, e->n, E, c.rms_eps));

    memset(&ea, 0, sizeof(ea));
    ea.out = e->act;
    ea.ld_out = M;
    gp.n = 2 * M;
    gp.k = E;
    g_class = KC_MLP_UP;
    ss_args(ea, e->rows_ss_h, nullptr);
    if (rows_k) {
      const RowsPlan pl = plan_rows(2 * M, E, r_tile, e->num_sms, EPI_SWIGLU);
      MTX_TRY(launch_rows<EPI_SWIGLU>(e->tm_w01[l], xmap(pl, fused_norm ? &XMaps::h : &XMaps::n), gp, ea, pl, st));
    }
    else MTX_TRY(launch_gemm<EPI_SWIGLU>(e->tm_w01[l], xm->n, gp, ea, plan_up, st));

    memset(&ea, 0, sizeof(ea));
    ea.out = e->x;
    ea.resid = e->h;
    ea.ld_out = E;
    gp.n = E;
    gp.k = M;
    g_class = KC_MLP_DOWN;
    ss_args(ea, nullptr, e->rows_ss_x);
    if (rows_k) {
      const RowsPlan pl = plan_rows(E, M, r_tile, e->num_sms, EPI_RESIDUAL);
      MTX_TRY(launch_rows<EPI_RESIDUAL>(e->tm_wout[l], xmap(pl, &XMaps::act), gp, ea, pl, st));
    }
    else MTX_TRY(launch_gemm<EPI_RESIDUAL>(e->tm_wout[l], xm->act, gp, ea, plan_down, st));

    g_class = KC_RMSNORM;
    const bf16* next_scale = l + 1 < L ? attn_norm + size_t(l + 1) * E : static_cast<const bf16*>(e->w.final_norm);
    if ((l + 1 < L && !fused_norm) || (l + 1 == L && want_logits))
      MTX_TRY(launch(rmsnorm_kernel<false>, dim3(rows), dim3(128), 0, st, (const bf16*)e->x, (const int*)nullptr,
                     (const bf16*)nullptr, next_scale, (bf16*)nullptr, e->n, E, c.rms_eps));
  }

  }

  if (want_logits) {
    const bool two_pass = e->strategy == MTX_SAMPLE_NUCLEUS || e->strategy == MTX_SAMPLE_TOPK;
    if (two_pass && (mode == 0 ? e->s.logits == nullptr : prefill_logits == nullptr))
      return fail(MTX_ERR_ARG, "nucleus / topk sampling read the materialised logits: decode_state.logits (or logits_out) must be set");
    EpiArgs ea;
    memset(&ea, 0, sizeof(ea));
    ea.logits_out = mode == 0 ? e->s.logits : prefill_logits;
    ea.ld_logits = c.vocab_size;
    ea.logits_only_row = mode == 0 ? -1 : rows - 1;
    ea.part_score = e->part_score;
    ea.part_idx = e->part_idx;
    ea.part_raw = e->part_raw;
    ea.part_max = e->part_max;
    ea.part_sum = e->part_sum;
    ea.n_tiles = plan_logits.n_tiles;
    ea.vocab_offset = c.vocab_offset;
    ea.scale = c.logits_scale;
    ea.softcap = c.final_softcap;
    ea.inv_temp = 1.0f / e->temperature;
    ea.round_bf16 = c.logits_round_bf16;
    ea.gumbel = e->strategy == MTX_SAMPLE_WEIGHTED ? 1 : 0;
    float* lp_out = mode == 0 ? e->s.log_prob : first_log_prob;
    // decode steps with top-k / nucleus run the all-SM sampler, which starts from the per-tile (max, sum exp) partials
    const bool par_sampler = two_pass && mode == 0 && cand_out == nullptr && env_int("MTX_PAR_SAMPLER", 1) != 0;
    // vocab-parallel top-k / nucleus: the shard's 64 candidates come from the same all-SM radix descent (par_collect_kernel)
    const bool par_collect = two_pass && mode == 0 && cand_out != nullptr && env_int("MTX_PAR_COLLECT", 1) != 0;
    ea.want_lse = (((lp_out != nullptr || cand_out != nullptr) && !two_pass) || par_sampler || par_collect) ? 1 : 0;
    ea.rng_state = e->s.rng_state;
    ea.row_offset = mode == 0 ? 0 : prefill_noise_row(e);
    gp.n = c.vocab_size;
    gp.k = E;
    g_class = KC_LOGITS;
    if (mega) {
      g_class = KC_PERSISTENT;
      MTX_TRY(launch_persistent(e, rows, *xm, ea, st));
    } else if (rows_k) {
      MTX_TRY(launch_rows<EPI_LOGITS>(e->tm_logits, xm->n, gp, ea, plan_rows(c.vocab_size, E, r_tile, e->num_sms, EPI_LOGITS), st));
    } else {
      MTX_TRY(launch_gemm<EPI_LOGITS>(e->tm_logits, xm->n, gp, ea, plan_logits, st));
    }
    g_class = KC_FINALIZE;

    FinalizeArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.part_score = e->part_score;
    fa.part_idx = e->part_idx;
    fa.part_raw = e->part_raw;
    fa.part_max = e->part_max;
    fa.part_sum = e->part_sum;
    fa.n_tiles = plan_logits.n_tiles;
    fa.rows = rows;
    fa.stride_r = plan_logits.n_tiles;
    fa.stride_t = 1;
    fa.cand_out = cand_out;
    int finalize_rows = rows;
    auto par_args = [&](ParSampleArgs& ps) {
      memset(&ps, 0, sizeof(ps));
      ps.logits = e->s.logits;
      ps.ld = c.vocab_size;
      ps.vocab = c.vocab_size;
      ps.vocab_offset = c.vocab_offset;
      ps.rows = rows;
      ps.rng_state = e->s.rng_state;
      ps.row_offset = 0;
      ps.part_max = e->part_max;
      ps.part_sum = e->part_sum;
      ps.n_tiles = plan_logits.n_tiles;
      ps.row_M = e->par_M;
      ps.row_Z = e->par_Z;
      ps.target = e->par_target;
      ps.above = e->par_above;
      ps.prefix = e->par_prefix;
      ps.found = e->par_found;
      ps.hist = e->par_hist;
      ps.eq_count = e->par_eq;
      ps.slices = par_slices(c.vocab_size);
      ps.row_group = par_row_group(c.vocab_size, rows);
    };
    if (par_collect) {
      // this shard's kCandK best logits per row + its (max, sum exp): the radix descent with top_k = kCandK, whatever the
      // strategy; the selection happens after the all-gather (mtx_commit_candidates)
      ParSampleArgs ps;
      par_args(ps);
      ps.mode = MTX_SAMPLE_TOPK;
      ps.top_k = kCandK;
      ps.cand = cand_out;
      ps.cand_count = e->par_cand_count;
      const dim3 sweep(ps.slices, (rows + ps.row_group - 1) / ps.row_group);
      MTX_TRY(launch(par_stats_kernel, dim3(rows), dim3(kParThreads), 0, st, ps));
      for (int level = 0; level < kParLevels; ++level) {
        MTX_TRY(launch(par_hist_kernel, sweep, dim3(kParThreads), 0, st, ps, level));
        MTX_TRY(launch(par_select_kernel, dim3(rows), dim3(256), 0, st, ps, level));
      }
      MTX_TRY(launch(par_eqcount_kernel, sweep, dim3(kParThreads), 0, st, ps));
      return launch(par_collect_kernel, sweep, dim3(kParThreads), 0, st, ps);
    }
    if (two_pass && cand_out != nullptr) {
      // (MTX_PAR_COLLECT=0: the one-CTA-per-row extraction of round 1)
      ShardTopkArgs ta;
      memset(&ta, 0, sizeof(ta));
      ta.logits = e->s.logits;
      ta.ld = c.vocab_size;
      ta.vocab = c.vocab_size;
      ta.vocab_offset = c.vocab_offset;
      ta.cand = cand_out;
      return launch(shard_topk_kernel, dim3(rows), dim3(kSampleThreads), 0, st, ta);
    }
    if (par_sampler) {
      // inference_utils.py:87-111 on the logits the GEMM just wrote, cut along the vocabulary over all SMs (sampling.cuh, par_*)
      ParSampleArgs ps;
      memset(&ps, 0, sizeof(ps));
      ps.logits = e->s.logits;
      ps.ld = c.vocab_size;
      ps.vocab = c.vocab_size;
      ps.vocab_offset = c.vocab_offset;
      ps.rows = rows;
      ps.mode = e->strategy;
      ps.top_k = e->top_k;
      ps.nucleus_p = e->nucleus_p;
      ps.inv_temp = 1.0f / e->temperature;
      ps.rng_state = e->s.rng_state;
      ps.row_offset = 0;
      ps.part_max = e->part_max;
      ps.part_sum = e->part_sum;
      ps.n_tiles = plan_logits.n_tiles;
      ps.row_M = e->par_M;
      ps.row_Z = e->par_Z;
      ps.target = e->par_target;
      ps.above = e->par_above;
      ps.prefix = e->par_prefix;
      ps.found = e->par_found;
      ps.hist = e->par_hist;
      ps.eq_count = e->par_eq;
      ps.out_score = e->part_score;
      ps.out_idx = e->part_idx;
      ps.out_raw = e->part_raw;
      ps.out_max = e->part_max;
      ps.out_sum = e->part_sum;
      ps.slices = par_slices(c.vocab_size);
      ps.row_group = par_row_group(c.vocab_size, rows);
      const dim3 sweep(ps.slices, (rows + ps.row_group - 1) / ps.row_group);
      MTX_TRY(launch(par_stats_kernel, dim3(rows), dim3(kParThreads), 0, st, ps));
      for (int level = 0; level < kParLevels; ++level) {
        MTX_TRY(launch(par_hist_kernel, sweep, dim3(kParThreads), 0, st, ps, level));
        MTX_TRY(launch(par_select_kernel, dim3(rows), dim3(256), 0, st, ps, level));
      }
      if (e->strategy == MTX_SAMPLE_TOPK) MTX_TRY(launch(par_eqcount_kernel, sweep, dim3(kParThreads), 0, st, ps));
      MTX_TRY(launch(par_scan_kernel, sweep, dim3(kParThreads), 0, st, ps));
      fa.n_tiles = ps.slices * kParWarps;
      fa.stride_r = ps.slices * kParWarps;
    } else if (two_pass) {
      // inference_utils.py:87-111 on the logits the GEMM just wrote; one candidate per row
      SampleArgs sa;
      memset(&sa, 0, sizeof(sa));
      sa.logits = mode == 0 ? e->s.logits : prefill_logits;
      sa.ld = c.vocab_size;
      sa.vocab = c.vocab_size;
      sa.vocab_offset = c.vocab_offset;
      sa.mode = e->strategy;
      sa.top_k = e->top_k;
      sa.nucleus_p = e->nucleus_p;
      sa.inv_temp = 1.0f / e->temperature;
      sa.rng_state = e->s.rng_state;
      sa.row_offset = mode == 0 ? 0 : prefill_noise_row(e) + rows - 1;
      sa.out_score = e->part_score;
      sa.out_idx = e->part_idx;
      sa.out_raw = e->part_raw;
      sa.out_max = e->part_max;
      sa.out_sum = e->part_sum;
      finalize_rows = mode == 0 ? rows : 1;
      MTX_TRY(launch(sample_rows_kernel, dim3(finalize_rows), dim3(kSampleThreads), 0, st, sa));
      fa.n_tiles = 1;
      fa.stride_r = 1;
      fa.rows = finalize_rows;
    }
    fa.mode = cand_out != nullptr ? 2 : mode;
    fa.have_lse = (two_pass || lp_out != nullptr || cand_out != nullptr) ? 1 : 0;
    fa.tokens = e->s.tokens;
    fa.next_pos = e->s.next_pos;
    fa.generated = e->s.generated;
    fa.ar_lengths = e->s.ar_lengths;
    fa.ar_index = e->s.ar_index;
    fa.result = e->s.result;
    fa.log_prob = lp_out;
    fa.rng_state = e->s.rng_state;
    fa.num_slots = c.num_slots;
    fa.R = c.max_target_len - c.max_prefill_len;
    fa.first_token = first_token;
    MTX_TRY(launch(finalize_kernel, dim3(finalize_rows), dim3(kFinalizeThreads), 0, st, fa));
  }
  return MTX_OK;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================

extern "C" {

const char* mtx_last_error(void) { return g_error.c_str(); }
size_t mtx_step_trace_words(const mtx_engine* e) {
  const size_t ctas = e && e->pk_ctas > 0 ? size_t(e->pk_ctas) : 148, bars = size_t(pk_trace_bars(e ? e->cfg.num_layers : 24));
  return 2 * bars * ctas + ctas * 3 * 64;
}
void mtx_debug_set_trace(void* device_buffer) { g_trace = static_cast<long long*>(device_buffer); }
int mtx_debug_set_timeline(void* device_buffer) {
  unsigned long long* p = static_cast<unsigned long long*>(device_buffer);
  MTX_CUDA(cudaMemcpyToSymbol(d_timeline, &p, sizeof(p)));
  return MTX_OK;
}
const char* mtx_build_info(void) { return "mtx_b200 sm_100a (tcgen05 + TMA + mbarrier), CUDA " MTX_STR(CUDART_VERSION); }
uint64_t mtx_launch_count(void) { return g_launches.load(); }

int mtx_engine_create(const mtx_model_config* cfg, mtx_engine** out) {
  if (cfg == nullptr || out == nullptr) return fail(MTX_ERR_ARG, "null argument");
  const mtx_model_config& c = *cfg;
  if (c.head_dim != 64 && c.head_dim != 128 && !(c.head_dim == 256 && c.decoder_block == 1))
    return fail(MTX_ERR_UNSUPPORTED, "head_dim %d: 64 and 128 (and 256 for the gemma3 block)", c.head_dim);
  if (c.head_dim == 256 && c.num_q_heads / c.num_kv_heads > 8) return fail(MTX_ERR_UNSUPPORTED, "head_dim 256 needs at most 8 query heads per kv head");
  if (c.num_q_heads % c.num_kv_heads != 0 || c.num_q_heads / c.num_kv_heads > 16)
    return fail(MTX_ERR_UNSUPPORTED, "query heads per kv head must divide evenly and be <= 16");
  if (c.emb_dim % 64 != 0 || c.mlp_dim % 64 != 0 || (c.num_q_heads * c.head_dim) % 64 != 0)
    return fail(MTX_ERR_UNSUPPORTED, "emb_dim, mlp_dim and Hq*D must be multiples of 64");
  if (c.mlp_dim % 16 != 0) return fail(MTX_ERR_UNSUPPORTED, "mlp_dim must be a multiple of 16");
  if (c.max_rows < 1 || c.max_rows > 256) return fail(MTX_ERR_ARG, "max_rows must be in [1, 256]");
  if (c.max_target_len <= c.max_prefill_len) return fail(MTX_ERR_ARG, "max_target_len must exceed max_prefill_len");
  if (c.num_layers < 1 || c.vocab_size < 1 || c.num_slots < 1) return fail(MTX_ERR_ARG, "bad layer / vocab / slot count");
  mtx_engine* e = new mtx_engine();
  e->cfg = c;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
    e->num_sms = sms;
  cudaGetLastError();
  if (c.kv_quant < 0 || c.kv_quant > 4)
    return fail(MTX_ERR_ARG, "kv_quant must be 0 (bf16), 1 / 2 (int8, one scale per token and kv head / per token) or 3 / 4 (the same with fp8 bytes)");
  if (kvq_axis(c) == 2 && c.num_kv_heads > 32) return fail(MTX_ERR_UNSUPPORTED, "kv_quant 2: at most 32 kv heads");
  if (c.kv_quant && c.head_dim != 64) return fail(MTX_ERR_UNSUPPORTED, "the int8 KV cache is implemented for head_dim 64");
  if (c.decoder_block != 0 && c.decoder_block != 1) return fail(MTX_ERR_ARG, "decoder_block must be 0 (llama2) or 1 (gemma3)");
  if (c.decoder_block == 1 && (c.kv_quant || !c.norm_scales_folded || c.sliding_window <= 0))
    return fail(MTX_ERR_UNSUPPORTED, "the gemma3 block needs a bf16 KV cache, norm_scales_folded and a sliding_window");
  if (is_paged(c)) {
    const int tpp = c.paged_tokens_per_page;
    if (c.kv_quant || c.decoder_block != 0) return fail(MTX_ERR_UNSUPPORTED, "attention=paged: bf16 cache and the llama2 block only");
    if (c.head_dim != 64 && c.head_dim != 128) return fail(MTX_ERR_UNSUPPORTED, "attention=paged: head_dim 64 or 128");
    if (tpp < 8 || (tpp & (tpp - 1)) != 0) return fail(MTX_ERR_UNSUPPORTED, "pagedattn_tokens_per_page must be a power of two >= 8");
    if (c.paged_num_pages <= 1) return fail(MTX_ERR_ARG, "`pagedattn_num_pages` must be greater than 1.");
    if (c.paged_max_pages_per_group < (c.max_target_len + tpp - 1) / tpp)
      return fail(MTX_ERR_ARG, "`pagedattn_max_pages_per_group` (%d) is insufficient for `max_target_length` (%d). Needs %d.",
                  c.paged_max_pages_per_group, c.max_target_len, (c.max_target_len + tpp - 1) / tpp);
    if (uint64_t(c.num_layers) * c.num_kv_heads * c.paged_num_pages * tpp >= (1ull << 31))
      return fail(MTX_ERR_UNSUPPORTED, "the page pools have too many rows for one tensor map");
  }
  e->max_r_tile = round_rows(c.max_rows);
  if (c.kv_quant && e->max_r_tile < 128) e->max_r_tile = 128;  // every step of an int8 engine runs the 128-row-block GEMM
  if (c.decoder_block == 1 && e->max_r_tile < 128) e->max_r_tile = 128;  // and so does every step of the gemma3 block
  e->qkv_n = (c.num_q_heads + 2 * c.num_kv_heads) * c.head_dim;
  e->attn_max_chunks = attn_max_chunks(c.max_prefill_len, c.max_target_len);
  e->rope_timescale_host.resize(c.head_dim / 2);
  for (int i = 0; i < c.head_dim / 2; ++i) {
    // embeddings.py:270-275, evaluated in fp64 and rounded once
    const double fraction = 2.0 * double(i) / double(c.head_dim);
    e->rope_timescale_host[i] = float(double(c.rope_min_timescale) * pow(double(c.rope_max_timescale) / double(c.rope_min_timescale), fraction));
  }
  e->rope_timescale_w_host.resize(c.head_dim / 2);
  for (int i = 0; i < c.head_dim / 2; ++i) {
    const double fraction = 2.0 * double(i) / double(c.head_dim);
    const double hi = c.local_rope_max_timescale > 0.0f ? double(c.local_rope_max_timescale) : double(c.rope_max_timescale);  // attentions.py:2085-2088
    e->rope_timescale_w_host[i] = float(double(c.rope_min_timescale) * pow(hi / double(c.rope_min_timescale), fraction));
  }
  e->ws_bytes = layout_workspace(e).total;
  *out = e;
  return MTX_OK;
}

int mtx_engine_destroy(mtx_engine* e) {
  if (e == nullptr) return MTX_OK;
  for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
  for (auto& kv : e->host_graphs) cudaGraphExecDestroy(kv.second);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  delete e;
  return MTX_OK;
}

size_t mtx_engine_workspace_bytes(const mtx_engine* e) { return e ? e->ws_bytes : 0; }

int mtx_engine_bind(mtx_engine* e, const mtx_weights* w, const mtx_decode_state* s, void* workspace, size_t workspace_bytes) {
  if (!e || !w || !s || !workspace) return fail(MTX_ERR_ARG, "null argument");
  if (workspace_bytes < e->ws_bytes) return fail(MTX_ERR_ARG, "workspace too small: %zu < %zu", workspace_bytes, e->ws_bytes);
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0) return fail(MTX_ERR_ARG, "workspace must be 1024-byte aligned");
  const mtx_model_config& c = e->cfg;
  if (c.decoder_block == 1 && (!w->q_norm || !w->k_norm || !w->post_attn_norm || !w->post_ffw_norm))
    return fail(MTX_ERR_ARG, "the gemma3 block needs q_norm, k_norm, post_attn_norm and post_ffw_norm");
  e->w = *w;
  e->s = *s;
  for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
  e->graphs.clear();
  for (auto& kv : e->host_graphs) cudaGraphExecDestroy(kv.second);
  e->host_graphs.clear();
  for (auto& m : e->xmaps) m.built = false;
  // (re)binding rewrites the workspace: wait for everything the device still runs, on any stream, before and after
  MTX_CUDA(cudaDeviceSynchronize());
  MTX_CUDA(cudaMemset(workspace, 0, e->ws_bytes));
  const WsLayout L = layout_workspace(e);
  uint8_t* b = static_cast<uint8_t*>(workspace);
  e->x = reinterpret_cast<bf16*>(b + L.x);
  e->h = reinterpret_cast<bf16*>(b + L.h);
  e->n = reinterpret_cast<bf16*>(b + L.n);
  e->q = reinterpret_cast<bf16*>(b + L.q);
  e->attn = reinterpret_cast<bf16*>(b + L.attn);
  e->act = reinterpret_cast<bf16*>(b + L.act);
  e->attn_part_o = reinterpret_cast<float*>(b + L.attn_part_o);
  e->attn_part_ml = reinterpret_cast<float*>(b + L.attn_part_ml);
  e->attn_tickets = reinterpret_cast<int*>(b + L.attn_tickets);
  e->rd.token = reinterpret_cast<int*>(b + L.token);
  e->rd.pos = reinterpret_cast<int*>(b + L.pos);
  e->rd.plane = reinterpret_cast<int*>(b + L.plane);
  e->rd.write_row = reinterpret_cast<int*>(b + L.write_row);
  e->rd.len0 = reinterpret_cast<int*>(b + L.len0);
  e->rd.ring_first = reinterpret_cast<int*>(b + L.ring_first);
  e->rd.ring_len = reinterpret_cast<int*>(b + L.ring_len);
  e->rd.rope_cs = reinterpret_cast<float2*>(b + L.rope_cs);
  e->rd.work_items = reinterpret_cast<int*>(b + L.work_items);
  e->rd.work_count = reinterpret_cast<int*>(b + L.work_count);
  e->rope_timescale = reinterpret_cast<float*>(b + L.rope_timescale);
  e->rd.iota = e->rd.tmp_row = nullptr;
  if (kvq_axis(c) == 2) {
    e->rd.iota = reinterpret_cast<int*>(b + L.iota);
    e->rd.tmp_row = reinterpret_cast<int*>(b + L.tmp_row);
    e->kv_tmp_k = reinterpret_cast<bf16*>(b + L.kv_tmp_k);
    e->kv_tmp_v = reinterpret_cast<bf16*>(b + L.kv_tmp_v);
  }
  if (c.decoder_block == 1) {
    e->rd.len0_w = reinterpret_cast<int*>(b + L.len0_w);
    e->rd.ring_first_w = reinterpret_cast<int*>(b + L.ring_first_w);
    e->rd.ring_len_w = reinterpret_cast<int*>(b + L.ring_len_w);
    e->rd.skip_w = reinterpret_cast<int*>(b + L.skip_w);
    e->rd.rope_cs_w = reinterpret_cast<float2*>(b + L.rope_cs_w);
    e->rd.work_items_w = reinterpret_cast<int*>(b + L.work_items_w);
    e->rd.work_count_w = reinterpret_cast<int*>(b + L.work_count_w);
    e->rope_timescale_w = reinterpret_cast<float*>(b + L.rope_timescale_w);
    e->qkv_tmp = reinterpret_cast<bf16*>(b + L.qkv_tmp);
  }
  e->part_score = reinterpret_cast<float*>(b + L.part_score);
  e->part_idx = reinterpret_cast<int*>(b + L.part_idx);
  e->part_raw = reinterpret_cast<float*>(b + L.part_raw);
  e->part_max = reinterpret_cast<float*>(b + L.part_max);
  e->part_sum = reinterpret_cast<float*>(b + L.part_sum);
  e->grid_bar = reinterpret_cast<unsigned int*>(b + L.grid_bar);
  e->pk_tables = reinterpret_cast<PkTable*>(b + L.pk_tables);
  e->pk_part_ws = reinterpret_cast<float*>(b + L.pk_part_ws);
  e->pk_ss_x = reinterpret_cast<float*>(b + L.pk_ss_x);
  e->pk_ss_h = reinterpret_cast<float*>(b + L.pk_ss_h);
  e->pk_attn_part_o = reinterpret_cast<float*>(b + L.pk_attn_part_o);
  e->pk_tile_prefix = reinterpret_cast<int*>(b + L.pk_tile_prefix);
  e->pk_attn_info = reinterpret_cast<int*>(b + L.pk_attn_info);
  e->rows_ss_x = reinterpret_cast<float*>(b + L.rows_ss_x);
  e->rows_ss_h = reinterpret_cast<float*>(b + L.rows_ss_h);
  e->cand_counters = reinterpret_cast<int*>(b + L.cand_counters);
  e->par_M = reinterpret_cast<float*>(b + L.par_M);
  e->par_cand_count = reinterpret_cast<int*>(b + L.par_cand_count);
  e->par_Z = reinterpret_cast<float*>(b + L.par_Z);
  e->par_target = reinterpret_cast<unsigned long long*>(b + L.par_target);
  e->par_above = reinterpret_cast<unsigned long long*>(b + L.par_above);
  e->par_prefix = reinterpret_cast<uint32_t*>(b + L.par_prefix);
  e->par_found = reinterpret_cast<int*>(b + L.par_found);
  e->par_hist = reinterpret_cast<unsigned long long*>(b + L.par_hist);
  e->par_eq = reinterpret_cast<int*>(b + L.par_eq);
  MTX_CUDA(cudaMemcpy(e->rope_timescale, e->rope_timescale_host.data(), e->rope_timescale_host.size() * 4, cudaMemcpyHostToDevice));
  if (c.decoder_block == 1)
    MTX_CUDA(cudaMemcpy(e->rope_timescale_w, e->rope_timescale_w_host.data(), e->rope_timescale_w_host.size() * 4, cudaMemcpyHostToDevice));

  const int E = c.emb_dim, HD = c.num_q_heads * c.head_dim, M = c.mlp_dim, L_ = c.num_layers;
  e->tm_wqkv.resize(L_);
  e->tm_wo.resize(L_);
  e->tm_w01.resize(L_);
  e->tm_wout.resize(L_);
  for (int l = 0; l < L_; ++l) {
    MTX_TRY(make_map(&e->tm_wqkv[l], static_cast<const bf16*>(w->wqkv) + size_t(l) * e->qkv_n * E, E, e->qkv_n, kTileN));
    MTX_TRY(make_map(&e->tm_wo[l], static_cast<const bf16*>(w->wo) + size_t(l) * E * HD, HD, E, kTileN));
    MTX_TRY(make_map(&e->tm_w01[l], static_cast<const bf16*>(w->w01) + size_t(l) * 2 * M * E, E, 2 * M, kTileN));
    MTX_TRY(make_map(&e->tm_wout[l], static_cast<const bf16*>(w->wout) + size_t(l) * E * M, M, E, kTileN));
  }
  MTX_TRY(make_map(&e->tm_logits, w->logits, E, c.vocab_size, kTileN));
  MTX_TRY(make_map(&e->tm_all_wqkv, w->wqkv, E, uint64_t(L_) * e->qkv_n, kTileN));
  MTX_TRY(make_map(&e->tm_all_wo, w->wo, HD, uint64_t(L_) * E, kTileN));
  MTX_TRY(make_map(&e->tm_all_w01, w->w01, E, uint64_t(L_) * 2 * M, kTileN));
  MTX_TRY(make_map(&e->tm_all_wout, w->wout, M, uint64_t(L_) * E, kTileN));
  e->pk_ctas = 0;
  if (c.head_dim == 64 && env_int("MTX_PERSISTENT", 1) != 0) {
    // one CTA per SM, all co-resident (the grid barrier spins): needs the full shared-memory carve-out
    int blocks = 0;
    auto pk_kernel = c.kv_quant == 0 ? step_persistent_kernel<0> : kvq_fp8(c) ? step_persistent_kernel<2> : step_persistent_kernel<1>;
    if (cudaFuncSetAttribute(pk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pk_smem_bytes())) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, pk_kernel, kPkThreads, pk_smem_bytes()) == cudaSuccess && blocks >= 1) {
      int ctas = e->num_sms;
      const int cap = env_int("MTX_PERSISTENT_CTAS", 0);
      if (cap > 0 && cap < ctas) ctas = cap;
      std::vector<PkTable> tabs(ctas), tabs_few(ctas);
      memset(tabs.data(), 0, tabs.size() * sizeof(PkTable));
      memset(tabs_few.data(), 0, tabs_few.size() * sizeof(PkTable));
      bool ok = true;
      for (int set = 0; set < 2 && ok; ++set) {
        std::vector<PkTable>& t = set == 0 ? tabs : tabs_few;
        ok = pk_fill_phase(t, PK_QKV, e->qkv_n, E, set == 0) && pk_fill_phase(t, PK_OPROJ, E, HD, set == 0) &&
             pk_fill_phase(t, PK_UP, 2 * M, E, set == 0) && pk_fill_phase(t, PK_DOWN, E, M, set == 0);
      }
      if (ok) {
        MTX_CUDA(cudaMemcpy(e->pk_tables, tabs.data(), tabs.size() * sizeof(PkTable), cudaMemcpyHostToDevice));
        MTX_CUDA(cudaMemcpy(e->pk_tables + ctas, tabs_few.data(), tabs_few.size() * sizeof(PkTable), cudaMemcpyHostToDevice));
        // the split-K exchange workspace starts out (and is left by every reader) as "nothing written"
        fill_u32_kernel<<<256, 256>>>(reinterpret_cast<uint32_t*>(e->pk_part_ws), size_t(e->num_sms) * 4 * kPkSlotFloats, kPkSentinel);
        {
          const mtx_model_config& cc = e->cfg;
          const size_t pk_rows = cc.max_rows < kPkMaxRTile ? cc.max_rows : kPkMaxRTile;
          const size_t Gq = cc.num_q_heads / cc.num_kv_heads;
          const size_t words = pk_rows * cc.num_kv_heads * kPkMaxParts * ((Gq * cc.head_dim + 2 * Gq + 3) / 4 * 4);
          fill_u32_kernel<<<256, 256>>>(reinterpret_cast<uint32_t*>(e->pk_attn_part_o), words, kPkSentinel);
        }
        MTX_CUDA(cudaGetLastError());
        MTX_CUDA(cudaDeviceSynchronize());
        e->pk_ctas = ctas;
      }
    }
    cudaGetLastError();
  }
  const uint64_t kv_rows = uint64_t(L_) * c.num_slots * c.num_kv_heads * c.max_target_len;
  if (kv_rows >= (1ull << 31)) return fail(MTX_ERR_UNSUPPORTED, "KV cache has too many rows for one tensor map");
  if (c.kv_quant) {
    if (!s->kq_cache || !s->vq_cache || !s->k_scale || !s->v_scale) return fail(MTX_ERR_ARG, "kv_quant: kq_cache / vq_cache / k_scale / v_scale must be set");
    const uint64_t stage_rows = uint64_t(L_) * c.num_kv_heads * c.max_target_len;  // one bf16 plane per layer
    MTX_TRY(make_map(&e->tm_k, s->k_cache, c.head_dim, stage_rows, kAttnTileRows));
    MTX_TRY(make_map(&e->tm_v, s->v_cache, c.head_dim, stage_rows, kAttnTileRows));
    MTX_TRY(make_map_u8(&e->tm_kq, s->kq_cache, kv_rows));
    MTX_TRY(make_map_u8(&e->tm_vq, s->vq_cache, kv_rows));
    MTX_CUDA(cudaFuncSetAttribute(decode_attn_q8_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(attn_q8_smem_bytes(16))));
    MTX_CUDA(cudaFuncSetAttribute(decode_attn_q8_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(attn_q8_smem_bytes(16))));
    MTX_CUDA(cudaFuncSetAttribute(decode_attn_q8_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(attn_q8_smem_bytes(16))));
    MTX_CUDA(cudaFuncSetAttribute(decode_attn_q8_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(attn_q8_smem_bytes(16))));
    // four CTAs of 46 KB per SM need the large shared-memory carve-out
    MTX_CUDA(cudaFuncSetAttribute(decode_attn_q8_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    MTX_CUDA(cudaFuncSetAttribute(decode_attn_q8_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  } else if (is_paged(c)) {
    if (!s->k_pages || !s->v_pages || !s->page_map || !s->page_lengths || !s->active_page || !s->active_page_pos)
      return fail(MTX_ERR_ARG, "attention=paged: k_pages / v_pages / page_map / page_lengths / active_page / active_page_pos must be set");
    if (c.paged_device_state && (!s->page_status || !s->num_pages_used || !s->has_active_page))
      return fail(MTX_ERR_ARG, "paged_device_state: page_status / num_pages_used / has_active_page must be set");
    const uint64_t stage_rows = uint64_t(L_) * c.num_kv_heads * c.max_target_len;  // one bf16 plane per layer
    MTX_TRY(make_map(&e->tm_k, s->k_cache, c.head_dim, stage_rows, kAttnTileRows));
    MTX_TRY(make_map(&e->tm_v, s->v_cache, c.head_dim, stage_rows, kAttnTileRows));
    const uint64_t pool_rows = uint64_t(L_) * c.num_kv_heads * c.paged_num_pages * c.paged_tokens_per_page;
    const uint32_t box = c.paged_tokens_per_page < kAttnTileRows ? c.paged_tokens_per_page : kAttnTileRows;
    MTX_TRY(make_map(&e->tm_kp, s->k_pages, c.head_dim, pool_rows, box));
    MTX_TRY(make_map(&e->tm_vp, s->v_pages, c.head_dim, pool_rows, box));
  } else {
  MTX_TRY(make_map(&e->tm_k, s->k_cache, c.head_dim, kv_rows, kAttnTileRows));
  MTX_TRY(make_map(&e->tm_v, s->v_cache, c.head_dim, kv_rows, kAttnTileRows));
  }
  if (c.head_dim == 64)
    MTX_CUDA(cudaFuncSetAttribute(decode_attn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(attn_smem_bytes(64, 16))));
  else
    MTX_CUDA(cudaFuncSetAttribute(decode_attn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(attn_smem_bytes(128, 16))));
  e->bound = true;
  return MTX_OK;
}

int mtx_engine_set_sampling(mtx_engine* e, int strategy, int top_k, float nucleus_p, float temperature) {
  if (!e) return fail(MTX_ERR_ARG, "null engine");
  if (strategy < MTX_SAMPLE_GREEDY || strategy > MTX_SAMPLE_TOPK) return fail(MTX_ERR_ARG, "Sampling algorithm=%d not supported!", strategy);
  if (strategy == MTX_SAMPLE_TOPK && top_k <= 0) return fail(MTX_ERR_ARG, "Can't apply algorithm topk with parameter topk=%d less than or equal to zero", top_k);
  if (strategy == MTX_SAMPLE_NUCLEUS && nucleus_p < 0) return fail(MTX_ERR_ARG, "Can't apply nucleus with parameter nucleus_topp=%f less zero", nucleus_p);
  if (!(temperature > 0.f)) return fail(MTX_ERR_ARG, "temperature must be positive");
  e->strategy = strategy;
  e->top_k = top_k;
  e->nucleus_p = nucleus_p;
  e->temperature = temperature;
  for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
  e->graphs.clear();
  for (auto& kv : e->host_graphs) cudaGraphExecDestroy(kv.second);
  e->host_graphs.clear();
  return MTX_OK;
}

int mtx_decode_step(mtx_engine* e, int rows, mtx_stream stream) {
  if (!e || !e->bound) return fail(MTX_ERR_ARG, "engine is not bound");
  if (rows < 1 || rows > e->cfg.max_rows) return fail(MTX_ERR_ARG, "rows %d outside [1, %d]", rows, e->cfg.max_rows);
  return enqueue_step(e, 0, rows, nullptr, 0, 0, 1, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

int mtx_decode_step_graph(mtx_engine* e, int rows, mtx_stream stream) {
  if (!e || !e->bound) return fail(MTX_ERR_ARG, "engine is not bound");
  if (rows < 1 || rows > e->cfg.max_rows) return fail(MTX_ERR_ARG, "rows %d outside [1, %d]", rows, e->cfg.max_rows);
  auto it = e->graphs.find(rows);
  if (it == e->graphs.end()) {
    if (!e->cap_stream) MTX_CUDA(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    MTX_CUDA(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = enqueue_step(e, 0, rows, nullptr, 0, 0, 1, nullptr, nullptr, e->cap_stream);
    const cudaError_t end = cudaStreamEndCapture(e->cap_stream, &graph);
    if (rc != MTX_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (end != cudaSuccess) return fail(MTX_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(end));
    cudaGraphExec_t exec = nullptr;
    const cudaError_t inst = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (inst != cudaSuccess) return fail(MTX_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(inst));
    it = e->graphs.emplace(rows, exec).first;
  }
  MTX_CUDA(cudaGraphLaunch(it->second, static_cast<cudaStream_t>(stream)));
  return MTX_OK;
}

int mtx_decode_step_host(mtx_engine* e, int rows, const int32_t* tokens_host, int32_t* result_host, float* log_prob_host, mtx_stream stream) {
  if (!e || !e->bound || !result_host) return fail(MTX_ERR_ARG, "engine is not bound / null result buffer");
  if (rows < 1 || rows > e->cfg.max_rows) return fail(MTX_ERR_ARG, "rows %d outside [1, %d]", rows, e->cfg.max_rows);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool want_lp = log_prob_host != nullptr && e->s.log_prob != nullptr;
  auto copies_in = [&](cudaStream_t s2) -> int {
    if (tokens_host != nullptr) MTX_CUDA(cudaMemcpyAsync(e->s.tokens, tokens_host, size_t(rows) * 4, cudaMemcpyHostToDevice, s2));
    return MTX_OK;
  };
  auto copies_out = [&](cudaStream_t s2) -> int {
    MTX_CUDA(cudaMemcpyAsync(result_host, e->s.result, size_t(rows) * 12, cudaMemcpyDeviceToHost, s2));
    if (want_lp) MTX_CUDA(cudaMemcpyAsync(log_prob_host, e->s.log_prob, size_t(rows) * 4, cudaMemcpyDeviceToHost, s2));
    return MTX_OK;
  };
  // Pinned buffers (what a serving loop reuses every step): ONE graph holds the copy in, the step's kernels and the copies
  // out, so a step is one driver call.  Pageable buffers cannot be captured: three calls.
  auto pinned = [](const void* ptr) {
    if (ptr == nullptr) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return at.type == cudaMemoryTypeHost;
  };
  const mtx_engine::HostKey key{rows, tokens_host, result_host, want_lp ? log_prob_host : nullptr};
  auto it = e->host_graphs.find(key);
  if (it == e->host_graphs.end()) {
    if (!(pinned(tokens_host) && pinned(result_host) && (!want_lp || pinned(log_prob_host))) || env_int("MTX_HOST_GRAPH", 1) == 0) {
      MTX_TRY(copies_in(st));
      MTX_TRY(mtx_decode_step_graph(e, rows, stream));
      return copies_out(st);
    }
    if (!e->cap_stream) MTX_CUDA(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    MTX_CUDA(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    int rc = copies_in(e->cap_stream);
    if (rc == MTX_OK) rc = enqueue_step(e, 0, rows, nullptr, 0, 0, 1, nullptr, nullptr, e->cap_stream);
    if (rc == MTX_OK) rc = copies_out(e->cap_stream);
    const cudaError_t end = cudaStreamEndCapture(e->cap_stream, &graph);
    if (rc != MTX_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (end != cudaSuccess) return fail(MTX_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(end));
    cudaGraphExec_t exec = nullptr;
    const cudaError_t inst = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (inst != cudaSuccess) return fail(MTX_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(inst));
    if (e->host_graphs.size() >= 16) {  // a caller cycling through buffers: do not grow without bound
      for (auto& kv : e->host_graphs) cudaGraphExecDestroy(kv.second);
      e->host_graphs.clear();
    }
    it = e->host_graphs.emplace(key, exec).first;
  }
  MTX_CUDA(cudaGraphLaunch(it->second, st));
  return MTX_OK;
}

int mtx_decode_step_host_sync(mtx_engine* e, int rows, const int32_t* tokens_host, int32_t* result_host, float* log_prob_host, mtx_stream stream) {
  MTX_TRY(mtx_decode_step_host(e, rows, tokens_host, result_host, log_prob_host, stream));
  MTX_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return MTX_OK;
}

int mtx_decode_step_candidates(mtx_engine* e, int rows, float* candidates, mtx_stream stream) {
  if (!e || !e->bound || !candidates) return fail(MTX_ERR_ARG, "engine is not bound / null candidates");
  if (rows < 1 || rows > e->cfg.max_rows) return fail(MTX_ERR_ARG, "rows %d outside [1, %d]", rows, e->cfg.max_rows);
  if (e->strategy == MTX_SAMPLE_TOPK && e->top_k > kCandK)
    return fail(MTX_ERR_UNSUPPORTED, "vocab-parallel top-k carries %d candidates per shard: decode_sampling_top_k must be <= %d", kCandK, kCandK);
  return enqueue_step(e, 0, rows, nullptr, 0, 0, 1, nullptr, nullptr, static_cast<cudaStream_t>(stream), candidates);
}

int mtx_sample_logits(mtx_engine* e, const float* logits, int rows, long long ld, int vocab, int row_offset, int32_t* token_out,
                      float* log_prob_out, mtx_stream stream) {
  if (!e || !e->bound || !logits || !token_out) return fail(MTX_ERR_ARG, "engine is not bound / null argument");
  if (rows < 1 || rows > e->cfg.max_rows || vocab < 1 || ld < vocab) return fail(MTX_ERR_ARG, "bad rows / vocab / ld");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SampleArgs sa;
  memset(&sa, 0, sizeof(sa));
  sa.logits = logits;
  sa.ld = ld;
  sa.vocab = vocab;
  sa.vocab_offset = 0;
  sa.mode = e->strategy;
  sa.top_k = e->top_k;
  sa.nucleus_p = e->nucleus_p;
  sa.inv_temp = 1.0f / e->temperature;
  sa.rng_state = e->s.rng_state;
  sa.row_offset = row_offset < 0 ? prefill_noise_row(e) + e->last_prefill_rows - 1 : row_offset;
  sa.out_score = e->part_score;
  sa.out_idx = e->part_idx;
  sa.out_raw = e->part_raw;
  sa.out_max = e->part_max;
  sa.out_sum = e->part_sum;
  MTX_TRY(launch(sample_rows_kernel, dim3(rows), dim3(kSampleThreads), 0, st, sa));
  FinalizeArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.part_score = e->part_score;
  fa.part_idx = e->part_idx;
  fa.part_raw = e->part_raw;
  fa.part_max = e->part_max;
  fa.part_sum = e->part_sum;
  fa.n_tiles = 1;
  fa.rows = rows;
  fa.stride_r = 1;
  fa.stride_t = 1;
  fa.mode = 3;
  fa.have_lse = 1;
  fa.first_token = token_out;
  fa.log_prob = log_prob_out;
  fa.rng_state = e->s.rng_state;
  return launch(finalize_kernel, dim3(rows), dim3(kFinalizeThreads), 0, st, fa);
}

size_t mtx_candidate_floats(const mtx_engine* e) {
  if (!e) return 0;
  return (e->strategy == MTX_SAMPLE_NUCLEUS || e->strategy == MTX_SAMPLE_TOPK) ? size_t(kCandFloats) : size_t(5);
}

int mtx_engine_counter(mtx_engine* e, int which, long long* value) {
  if (!e || !e->bound || !value || which != 0) return fail(MTX_ERR_ARG, "bad counter request");
  int v = 0;
  MTX_CUDA(cudaMemcpy(&v, e->cand_counters + which, sizeof(int), cudaMemcpyDeviceToHost));
  *value = v;
  return MTX_OK;
}

int mtx_commit_candidates(mtx_engine* e, int rows, const float* gathered, int n_shards, mtx_stream stream) {
  if (!e || !e->bound || !gathered || n_shards < 1) return fail(MTX_ERR_ARG, "bad commit arguments");
  if (rows < 1 || rows > e->cfg.max_rows) return fail(MTX_ERR_ARG, "rows %d outside [1, %d]", rows, e->cfg.max_rows);
  const mtx_model_config& c = e->cfg;
  if (e->strategy == MTX_SAMPLE_NUCLEUS || e->strategy == MTX_SAMPLE_TOPK) {
    if (n_shards > kCandMaxShards) return fail(MTX_ERR_UNSUPPORTED, "at most %d vocabulary shards", kCandMaxShards);
    CommitTopkArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.gathered = gathered;
    ca.n_shards = n_shards;
    ca.rows = rows;
    ca.shard_vocab = c.vocab_size;
    ca.mode = e->strategy;
    ca.top_k = e->top_k;
    ca.nucleus_p = e->nucleus_p;
    ca.inv_temp = 1.0f / e->temperature;
    ca.tokens = e->s.tokens;
    ca.next_pos = e->s.next_pos;
    ca.generated = e->s.generated;
    ca.ar_lengths = e->s.ar_lengths;
    ca.ar_index = e->s.ar_index;
    ca.result = e->s.result;
    ca.log_prob = e->s.log_prob;
    ca.rng_state = e->s.rng_state;
    ca.num_slots = c.num_slots;
    ca.R = c.max_target_len - c.max_prefill_len;
    ca.truncated = e->cand_counters;
    ca.ticket = e->cand_counters + 1;
    g_class = KC_FINALIZE;
    return launch(commit_topk_kernel, dim3(rows), dim3(kSampleThreads), 0, static_cast<cudaStream_t>(stream), ca);
  }
  FinalizeArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.part_score = gathered + 0 * rows;
  fa.part_idx = reinterpret_cast<const int*>(gathered + 1 * rows);
  fa.part_raw = gathered + 2 * rows;
  fa.part_max = gathered + 3 * rows;
  fa.part_sum = gathered + 4 * rows;
  fa.n_tiles = n_shards;
  fa.rows = rows;
  fa.stride_r = 1;
  fa.stride_t = 5LL * rows;
  fa.mode = 0;
  fa.have_lse = 1;
  fa.tokens = e->s.tokens;
  fa.next_pos = e->s.next_pos;
  fa.generated = e->s.generated;
  fa.ar_lengths = e->s.ar_lengths;
  fa.ar_index = e->s.ar_index;
  fa.result = e->s.result;
  fa.log_prob = e->s.log_prob;
  fa.rng_state = e->s.rng_state;
  fa.num_slots = c.num_slots;
  fa.R = c.max_target_len - c.max_prefill_len;
  g_class = KC_FINALIZE;
  return launch(finalize_kernel, dim3(rows), dim3(kFinalizeThreads), 0, static_cast<cudaStream_t>(stream), fa);
}

int mtx_profile_decode_step(mtx_engine* e, int rows, mtx_stream stream, float* class_ms, int32_t* class_launches) {
  if (!e || !e->bound || !class_ms || !class_launches) return fail(MTX_ERR_ARG, "bad profile arguments");
  if (rows < 1 || rows > e->cfg.max_rows) return fail(MTX_ERR_ARG, "rows %d outside [1, %d]", rows, e->cfg.max_rows);
  ProfileSink sink;
  g_profile = &sink;
  const int rc = enqueue_step(e, 0, rows, nullptr, 0, 0, 1, nullptr, nullptr, static_cast<cudaStream_t>(stream));
  g_profile = nullptr;
  cudaError_t sync = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  for (int i = 0; i < KC_COUNT; ++i) { class_ms[i] = 0.f; class_launches[i] = 0; }
  for (auto& r : sink.recs) {
    float ms = 0.f;
    if (sync == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      class_ms[r.cls] += ms;
      class_launches[r.cls] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  if (rc != MTX_OK) return rc;
  if (sync != cudaSuccess) return fail(MTX_ERR_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(sync));
  return MTX_OK;
}

int mtx_prefill_chunk(mtx_engine* e, const int32_t* tokens, int count, int start_pos, int slot, int sample_last,
                      int32_t* first_token, float* logits_out, float* first_log_prob, mtx_stream stream) {
  if (!e || !e->bound) return fail(MTX_ERR_ARG, "engine is not bound");
  if (count < 1 || count > e->cfg.max_rows) return fail(MTX_ERR_ARG, "count %d outside [1, %d]", count, e->cfg.max_rows);
  if (start_pos < 0 || start_pos + count > e->cfg.max_prefill_len) return fail(MTX_ERR_ARG, "prompt positions exceed max_prefill_len");
  if (slot < 0 || slot >= e->cfg.num_slots) return fail(MTX_ERR_ARG, "slot out of range");
  if (sample_last && first_token == nullptr) return fail(MTX_ERR_ARG, "first_token is null");
  if (sample_last) {
    ++e->prefill_draws;
    e->last_prefill_rows = count;
  }
  return enqueue_step(e, 1, count, tokens, start_pos, slot, sample_last, first_token, logits_out, static_cast<cudaStream_t>(stream), nullptr,
                      first_log_prob);
}

int mtx_insert_prefix(mtx_engine* e, const void* k_src, const void* v_src, int n_rows, int n_src_rows, int slot, int next_pos, int generated,
                      int token, mtx_stream stream) {
  if (!e || !e->bound || !k_src || !v_src) return fail(MTX_ERR_ARG, "engine is not bound / null prefix");
  const mtx_model_config& c = e->cfg;
  if (slot < 0 || slot >= c.num_slots) return fail(MTX_ERR_ARG, "slot out of range");
  if (n_rows < 1 || n_rows > n_src_rows || n_rows > c.max_prefill_len) return fail(MTX_ERR_ARG, "prefix rows %d outside [1, min(%d, %d)]", n_rows, n_src_rows, c.max_prefill_len);
  InsertArgs a;
  memset(&a, 0, sizeof(a));
  a.k_src = static_cast<const bf16*>(k_src);
  a.v_src = static_cast<const bf16*>(v_src);
  a.k_cache = static_cast<bf16*>(e->s.k_cache);
  a.v_cache = static_cast<bf16*>(e->s.v_cache);
  a.L = c.num_layers;
  a.hkv = c.num_kv_heads;
  a.D = c.head_dim;
  a.T = c.max_target_len;
  a.planes = c.num_slots;
  a.n = n_rows;
  a.n_src = n_src_rows;
  a.slot = slot;
  a.next_pos = next_pos;
  a.generated = generated;
  a.token = token;
  a.prefill_len = e->s.prefill_len;
  a.ar_lengths = e->s.ar_lengths;
  a.next_pos_out = e->s.next_pos;
  a.generated_out = e->s.generated;
  a.tokens_out = e->s.tokens;
  if (is_paged(c)) {
    // _copy_paged (maxengine.py:1104-1131): the prefix rows into the pages of group `slot`; then the slot's bookkeeping
    PagedInsertArgs g;
    memset(&g, 0, sizeof(g));
    g.k_src = a.k_src;
    g.v_src = a.v_src;
    g.k_pages = static_cast<bf16*>(e->s.k_pages);
    g.v_pages = static_cast<bf16*>(e->s.v_pages);
    g.page_map_row = e->s.page_map + size_t(slot) * c.paged_max_pages_per_group;
    g.layers = c.num_layers;
    g.hkv = c.num_kv_heads;
    g.d = c.head_dim;
    g.src_rows = n_src_rows;
    g.n_tokens = n_rows;
    g.num_pages = c.paged_num_pages;
    g.tokens_per_page = c.paged_tokens_per_page;
    const long long vecs_p = (long long)c.num_layers * c.num_kv_heads * n_rows * (c.head_dim / 8);
    int grid_p = int((vecs_p + 255) / 256);
    if (grid_p > e->num_sms * 8) grid_p = e->num_sms * 8;
    MTX_TRY(launch(paged_insert_kernel, dim3(grid_p), dim3(256), 0, static_cast<cudaStream_t>(stream), g));
    a.n = 0;  // (no rows: only the bookkeeping of insert_prefix_kernel)
    return launch(insert_prefix_kernel, dim3(1), dim3(32), 0, static_cast<cudaStream_t>(stream), a);
  }
  if (c.kv_quant) {
    InsertQ8Args q;
    q.base = a;
    q.kq_cache = static_cast<uint8_t*>(e->s.kq_cache);
    q.vq_cache = static_cast<uint8_t*>(e->s.vq_cache);
    q.k_scale = e->s.k_scale;
    q.v_scale = e->s.v_scale;
    q.shared_scale = kvq_axis(c) == 2 ? 1 : 0;
    q.fp8 = kvq_fp8(c) ? 1 : 0;
    const long long warps = (long long)c.num_layers * (kvq_axis(c) == 2 ? 1 : c.num_kv_heads) * n_rows * 2;
    int grid_q = int((warps + 7) / 8);
    if (grid_q > e->num_sms * 8) grid_q = e->num_sms * 8;
    if (kvq_axis(c) == 2) return launch(insert_prefix_q8_shared_kernel, dim3(grid_q), dim3(256), 0, static_cast<cudaStream_t>(stream), q);
    return launch(insert_prefix_q8_kernel, dim3(grid_q), dim3(256), 0, static_cast<cudaStream_t>(stream), q);
  }
  const long long vecs = (long long)c.num_layers * c.num_kv_heads * n_rows * (c.head_dim / 8);
  int grid = int((vecs + 255) / 256);
  if (grid > e->num_sms * 8) grid = e->num_sms * 8;
  return launch(insert_prefix_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), a);
}

// ---- single ops ------------------------------------------------------------------------------

int mtx_rmsnorm(const void* x, const void* scale, void* out, int rows, int emb_dim, float eps, mtx_stream stream) {
  if (!x || !scale || !out || rows < 1 || emb_dim % 8 != 0) return fail(MTX_ERR_ARG, "bad rmsnorm arguments");
  return launch(rmsnorm_kernel<false>, dim3(rows), dim3(128), 0, static_cast<cudaStream_t>(stream), static_cast<const bf16*>(x),
                (const int*)nullptr, (const bf16*)nullptr, static_cast<const bf16*>(scale), (bf16*)nullptr, static_cast<bf16*>(out),
                emb_dim, eps);
}

int mtx_linear(const void* x, const void* w, void* out, int rows, int n, int k, int splits, mtx_stream stream) {
  if (!x || !w || !out || rows < 1 || rows > 256 || n < 1 || k < 64 || k % 64 != 0) return fail(MTX_ERR_ARG, "bad linear arguments");
  const int r_tile = round_rows(rows);
  if (splits < 1) splits = 1;
  if ((splits & (splits - 1)) != 0 || splits > 16 || splits > k / kBlockK || splits > r_tile)
    return fail(MTX_ERR_ARG, "splits must be a power of two <= min(16, k/64, padded rows)");
  CUtensorMap tw, tx;
  MTX_TRY(make_map(&tw, w, k, n, kTileN));
  MTX_TRY(make_map(&tx, x, k, r_tile, r_tile));
  if (use_rows_kernel(r_tile) && splits <= 8) {
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.n = n;
    p.k = k;
    p.rows = rows;
    p.r_tile = r_tile;
    EpiArgs e;
    memset(&e, 0, sizeof(e));
    e.out = static_cast<bf16*>(out);
    e.ld_out = n;
    const RowsPlan pl = plan_rows(n, k, r_tile, 148, EPI_STORE_BF16, splits);
    if (pl.splits > 1) MTX_TRY(make_map(&tx, x, k, r_tile, 128));
    return launch_rows<EPI_STORE_BF16>(tw, tx, p, e, pl, static_cast<cudaStream_t>(stream));
  }
  const GemmPlan g = plan_gemm(n, k, r_tile, 148, EPI_STORE_BF16, splits);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.n = n;
  p.k = k;
  p.rows = rows;
  p.r_tile = r_tile;
  EpiArgs e;
  memset(&e, 0, sizeof(e));
  e.out = static_cast<bf16*>(out);
  e.ld_out = n;
  return launch_gemm<EPI_STORE_BF16>(tw, tx, p, e, g, static_cast<cudaStream_t>(stream));
}

int mtx_outproj_residual(const void* attn, const void* wo, const void* x, void* out, int rows, int emb_dim, int q_dim, mtx_stream stream) {
  if (!attn || !wo || !x || !out || rows < 1 || rows > 256 || emb_dim < 8 || emb_dim % 8 != 0 || q_dim < 64 || q_dim % 64 != 0)
    return fail(MTX_ERR_ARG, "bad out-projection arguments");
  EpiArgs ea;
  memset(&ea, 0, sizeof(ea));
  ea.out = static_cast<bf16*>(out);
  ea.resid = static_cast<const bf16*>(x);
  ea.ld_out = emb_dim;
  return single_gemm<EPI_RESIDUAL>(attn, wo, rows, emb_dim, q_dim, ea, static_cast<cudaStream_t>(stream));
}

size_t mtx_mlp_scratch_bytes(int rows, int emb_dim, int mlp_dim) {
  const size_t rt = size_t(round_rows(rows < 1 ? 1 : rows));
  return align_up(rt * emb_dim * 2, 1024) + align_up(rt * mlp_dim * 2, 1024);
}

int mtx_mlp(const void* h, const void* norm_scale, const void* w01, const void* wout, void* out, int rows, int emb_dim, int mlp_dim, float eps,
            void* scratch, mtx_stream stream) {
  if (!h || !norm_scale || !w01 || !wout || !out || !scratch || rows < 1 || rows > 256) return fail(MTX_ERR_ARG, "bad MLP arguments");
  if (emb_dim < 64 || emb_dim % 64 != 0 || mlp_dim < 64 || mlp_dim % 64 != 0) return fail(MTX_ERR_UNSUPPORTED, "emb_dim and mlp_dim must be multiples of 64");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t rt = size_t(round_rows(rows));
  bf16* n = static_cast<bf16*>(scratch);
  bf16* act = reinterpret_cast<bf16*>(static_cast<uint8_t*>(scratch) + align_up(rt * emb_dim * 2, 1024));
  // the rows past `rows` of the two scratch matrices feed MMA rows whose results are never stored; keep them finite
  MTX_CUDA(cudaMemsetAsync(scratch, 0, mtx_mlp_scratch_bytes(rows, emb_dim, mlp_dim), st));
  MTX_TRY(launch(rmsnorm_kernel<false>, dim3(rows), dim3(128), 0, st, static_cast<const bf16*>(h), (const int*)nullptr, (const bf16*)nullptr,
                 static_cast<const bf16*>(norm_scale), (bf16*)nullptr, n, emb_dim, eps));
  EpiArgs ea;
  memset(&ea, 0, sizeof(ea));
  ea.out = act;
  ea.ld_out = mlp_dim;
  MTX_TRY(single_gemm<EPI_SWIGLU>(n, w01, rows, 2 * mlp_dim, emb_dim, ea, st));
  memset(&ea, 0, sizeof(ea));
  ea.out = static_cast<bf16*>(out);
  ea.resid = static_cast<const bf16*>(h);
  ea.ld_out = emb_dim;
  return single_gemm<EPI_RESIDUAL>(act, wout, rows, emb_dim, mlp_dim, ea, st);
}

size_t mtx_attention_scratch_bytes(int rows, int num_kv_heads, int num_q_heads, int head_dim, int max_prefill_len, int max_target_len) {
  const size_t mc = attn_max_chunks(max_prefill_len, max_target_len);
  const size_t G = num_q_heads / num_kv_heads;
  size_t b = 0;
  b += align_up(size_t(rows) * mc * 4, 1024);                           // work items
  b += 1024;                                                            // work count
  b += align_up(size_t(rows) * num_kv_heads * 4, 1024);                 // tickets
  b += align_up(size_t(rows) * num_kv_heads * mc * G * 2 * 4, 1024);    // (max, sum)
  b += align_up(size_t(rows) * num_kv_heads * mc * G * head_dim * 4, 1024);
  return b;
}

int mtx_decode_attention(const void* q, const void* k_cache, const void* v_cache, const int32_t* plane, const int32_t* len0,
                         const int32_t* ring_first, const int32_t* ring_len, void* out, int rows, int num_slots, int num_q_heads,
                         int num_kv_heads, int head_dim, int max_prefill_len, int max_target_len, float softcap, void* scratch,
                         mtx_stream stream) {
  if (!q || !k_cache || !v_cache || !plane || !len0 || !ring_first || !ring_len || !out || !scratch) return fail(MTX_ERR_ARG, "null argument");
  if (head_dim != 64 && head_dim != 128) return fail(MTX_ERR_UNSUPPORTED, "head_dim %d: only 64 and 128", head_dim);
  if (rows < 1 || rows > 256) return fail(MTX_ERR_ARG, "rows must be in [1, 256]");
  if (num_q_heads % num_kv_heads != 0 || num_q_heads / num_kv_heads > 16) return fail(MTX_ERR_UNSUPPORTED, "bad head grouping");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t mc = attn_max_chunks(max_prefill_len, max_target_len);
  const size_t G = num_q_heads / num_kv_heads;
  uint8_t* b = static_cast<uint8_t*>(scratch);
  AttnParams p;
  memset(&p, 0, sizeof(p));
  int* work_items = reinterpret_cast<int*>(b);
  b += align_up(size_t(rows) * mc * 4, 1024);
  int* work_count = reinterpret_cast<int*>(b);
  b += 1024;
  p.tickets = reinterpret_cast<int*>(b);
  const size_t ticket_bytes = align_up(size_t(rows) * num_kv_heads * 4, 1024);
  b += ticket_bytes;
  p.part_ml = reinterpret_cast<float*>(b);
  b += align_up(size_t(rows) * num_kv_heads * mc * G * 2 * 4, 1024);
  p.part_o = reinterpret_cast<float*>(b);
  MTX_CUDA(cudaMemsetAsync(p.tickets, 0, ticket_bytes, st));
  MTX_TRY(launch(attn_build_worklist_kernel, dim3(1), dim3(256), 0, st, (const int*)len0, (const int*)ring_first, (const int*)ring_len,
                 rows, max_prefill_len, max_target_len, attn_tiles_per_item(rows, num_kv_heads, max_prefill_len, max_target_len, 148), work_items, work_count));
  CUtensorMap tk, tv;
  const uint64_t kv_rows = uint64_t(num_slots) * num_kv_heads * max_target_len;
  MTX_TRY(make_map(&tk, k_cache, head_dim, kv_rows, kAttnTileRows));
  MTX_TRY(make_map(&tv, v_cache, head_dim, kv_rows, kAttnTileRows));
  p.q = static_cast<const bf16*>(q);
  p.out = static_cast<bf16*>(out);
  p.plane = plane;
  p.len0 = len0;
  p.ring_first = ring_first;
  p.ring_len = ring_len;
  p.work_items = work_items;
  p.work_count = work_count;
  p.rows = rows;
  p.hq = num_q_heads;
  p.hkv = num_kv_heads;
  p.P = max_prefill_len;
  p.T = max_target_len;
  p.tiles_per_item = attn_tiles_per_item(rows, num_kv_heads, max_prefill_len, max_target_len, 148);
  p.max_chunks = int(mc);
  p.plane_base = 0;
  p.softcap = softcap;
  p.trace = g_trace;
  const size_t smem = attn_smem_bytes(head_dim, int(G));
  int grid = rows * num_kv_heads * attn_max_chunks(max_prefill_len, max_target_len, p.tiles_per_item);
  const int cap = 148 * (head_dim == 64 ? 3 : 1);
  if (grid > cap) grid = cap;
  if (head_dim == 64) {
    MTX_CUDA(cudaFuncSetAttribute(decode_attn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    return launch(decode_attn_kernel<64>, dim3(grid), dim3(kAttnThreads), smem, st, tk, tv, p);
  }
  MTX_CUDA(cudaFuncSetAttribute(decode_attn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  return launch(decode_attn_kernel<128>, dim3(grid), dim3(kAttnThreads), smem, st, tk, tv, p);
}

// ---- gpu_ragged_attention contract ----------------------------------------------------------------

namespace {
__global__ void ragged_rows_kernel(int* plane, int* zeros_a, int* zeros_b, int rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) {
    plane[i] = i;
    zeros_a[i] = 0;
    zeros_b[i] = 0;
  }
}
// RoPE table of mtx_qkv_rope_append: (cos, sin) of pos / timescale, rounded to bf16 (embeddings.py:270-307)
__global__ void rope_table_kernel(const int* pos, float2* cs, int rows, int half, float min_ts, float max_ts) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * half) return;
  const int r = idx / half, i = idx - r * half;
  const double fraction = 2.0 * double(i) / double(2 * half);
  const float ts = float(double(min_ts) * pow(double(max_ts) / double(min_ts), fraction));
  const float ang = float(pos[r]) / ts;
  cs[idx] = make_float2(bf16r(cosf(ang)), bf16r(sinf(ang)));
}
}  // namespace

size_t mtx_ragged_attention_scratch_bytes(int rows, int num_kv_heads, int num_q_heads, int head_dim, int seq_len) {
  return mtx_attention_scratch_bytes(rows, num_kv_heads, num_q_heads, head_dim, seq_len, seq_len + 64) + align_up(size_t(rows) * 12, 1024);
}

int mtx_ragged_attention(const void* q, const void* k, const void* v, const int32_t* lengths, void* out, float* out_max, float* out_sum,
                         int rows, int seq_len, int num_q_heads, int num_kv_heads, int head_dim, int seq_major, float softcap,
                         void* scratch, mtx_stream stream) {
  if (!q || !k || !v || !lengths || !out || !out_max || !out_sum || !scratch) return fail(MTX_ERR_ARG, "null argument");
  if (head_dim != 64 && head_dim != 128) return fail(MTX_ERR_UNSUPPORTED, "head_dim %d: only 64 and 128", head_dim);
  if (rows < 1 || rows > 256 || seq_len < 1) return fail(MTX_ERR_ARG, "rows must be in [1, 256], seq_len >= 1");
  if (num_q_heads % num_kv_heads != 0 || num_q_heads / num_kv_heads > 16) return fail(MTX_ERR_UNSUPPORTED, "bad head grouping");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the segment is addressed as a "prefill segment" of seq_len rows with an empty ring behind it
  const int P = seq_len, T = seq_len + 64;
  const size_t mc = attn_max_chunks(P, T);
  const size_t G = num_q_heads / num_kv_heads;
  uint8_t* b = static_cast<uint8_t*>(scratch);
  AttnParams p;
  memset(&p, 0, sizeof(p));
  int* work_items = reinterpret_cast<int*>(b);
  b += align_up(size_t(rows) * mc * 4, 1024);
  int* work_count = reinterpret_cast<int*>(b);
  b += 1024;
  p.tickets = reinterpret_cast<int*>(b);
  const size_t ticket_bytes = align_up(size_t(rows) * num_kv_heads * 4, 1024);
  b += ticket_bytes;
  p.part_ml = reinterpret_cast<float*>(b);
  b += align_up(size_t(rows) * num_kv_heads * mc * G * 2 * 4, 1024);
  p.part_o = reinterpret_cast<float*>(b);
  b += align_up(size_t(rows) * num_kv_heads * mc * G * head_dim * 4, 1024);
  int* plane = reinterpret_cast<int*>(b);
  int* zeros_a = plane + rows;
  int* zeros_b = zeros_a + rows;
  MTX_CUDA(cudaMemsetAsync(p.tickets, 0, ticket_bytes, st));
  ragged_rows_kernel<<<(rows + 127) / 128, 128, 0, st>>>(plane, zeros_a, zeros_b, rows);
  const int tpi = attn_tiles_per_item(rows, num_kv_heads, P, T, 148);
  MTX_TRY(launch(attn_build_worklist_kernel, dim3(1), dim3(256), 0, st, (const int*)lengths, (const int*)zeros_a, (const int*)zeros_b, rows, P, T, tpi,
                 work_items, work_count));
  CUtensorMap tk, tv;
  if (seq_major) {  // [rows, S, Hkv, D]: a tile is 64 rows of one head's D columns out of Hkv * D
    MTX_TRY(make_map(&tk, k, uint64_t(num_kv_heads) * head_dim, uint64_t(rows) * seq_len, kAttnTileRows));
    MTX_TRY(make_map(&tv, v, uint64_t(num_kv_heads) * head_dim, uint64_t(rows) * seq_len, kAttnTileRows));
  } else {          // [rows, Hkv, S, D]
    MTX_TRY(make_map(&tk, k, head_dim, uint64_t(rows) * num_kv_heads * seq_len, kAttnTileRows));
    MTX_TRY(make_map(&tv, v, head_dim, uint64_t(rows) * num_kv_heads * seq_len, kAttnTileRows));
  }
  p.q = static_cast<const bf16*>(q);
  p.out = static_cast<bf16*>(out);
  p.out_max = out_max;
  p.out_sum = out_sum;
  p.seq_major = seq_major ? 1 : 0;
  p.plane = plane;
  p.len0 = lengths;
  p.ring_first = zeros_a;
  p.ring_len = zeros_b;
  p.work_items = work_items;
  p.work_count = work_count;
  p.rows = rows;
  p.hq = num_q_heads;
  p.hkv = num_kv_heads;
  p.P = P;
  p.T = seq_len;  // rows allocated per (plane, head): the stride of a plane in the tensor map
  p.tiles_per_item = tpi;
  p.max_chunks = int(mc);
  p.softcap = softcap;
  // (attn_tile only ever sees prefill tiles here: ring_len = 0, so R = T - P is never used as a divisor)
  const size_t smem = attn_smem_bytes(head_dim, int(G));
  int grid = rows * num_kv_heads * attn_max_chunks(P, T, tpi);
  const int cap = 148 * (head_dim == 64 ? 3 : 1);
  if (grid > cap) grid = cap;
  if (head_dim == 64) {
    MTX_CUDA(cudaFuncSetAttribute(decode_attn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    return launch(decode_attn_kernel<64>, dim3(grid), dim3(kAttnThreads), smem, st, tk, tv, p);
  }
  MTX_CUDA(cudaFuncSetAttribute(decode_attn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  return launch(decode_attn_kernel<128>, dim3(grid), dim3(kAttnThreads), smem, st, tk, tv, p);
}

// ---- paged KV cache as single ops (inference/paged_attention.py) ---------------------------------------

int mtx_paged_append(void* k_pages, void* v_pages, const void* k_new, const void* v_new, const int32_t* active_page,
                     const int32_t* active_pos, int rows, int num_kv_heads, int head_dim, int num_pages, int tokens_per_page,
                     mtx_stream stream) {
  if (!k_pages || !v_pages || !k_new || !v_new || !active_page || !active_pos) return fail(MTX_ERR_ARG, "null argument");
  if (rows < 1 || num_kv_heads < 1 || head_dim % 8 != 0 || num_pages < 1 || tokens_per_page < 1) return fail(MTX_ERR_ARG, "bad paged append shape");
  PagedAppendArgs a;
  memset(&a, 0, sizeof(a));
  a.k_new = static_cast<const bf16*>(k_new);
  a.v_new = static_cast<const bf16*>(v_new);
  a.k_pages = static_cast<bf16*>(k_pages);
  a.v_pages = static_cast<bf16*>(v_pages);
  a.active_page = active_page;
  a.active_pos = active_pos;
  a.hkv = num_kv_heads;
  a.d = head_dim;
  a.num_pages = num_pages;
  a.tokens_per_page = tokens_per_page;
  return launch(paged_append_kernel, dim3(rows), dim3(128), 0, static_cast<cudaStream_t>(stream), a);
}

size_t mtx_paged_attention_scratch_bytes(int rows, int num_kv_heads, int num_q_heads, int head_dim, int max_tokens) {
  return mtx_attention_scratch_bytes(rows, num_kv_heads, num_q_heads, head_dim, max_tokens, max_tokens + 64) + align_up(size_t(rows) * 12, 1024);
}

int mtx_paged_attention(const void* q, const void* k_pages, const void* v_pages, const int32_t* lengths, const int32_t* page_map,
                        void* out, int rows, int num_q_heads, int num_kv_heads, int head_dim, int num_pages, int tokens_per_page,
                        int max_pages_per_group, float softcap, void* scratch, mtx_stream stream) {
  if (!q || !k_pages || !v_pages || !lengths || !page_map || !out || !scratch) return fail(MTX_ERR_ARG, "null argument");
  if (head_dim != 64 && head_dim != 128) return fail(MTX_ERR_UNSUPPORTED, "head_dim %d: only 64 and 128", head_dim);
  if (rows < 1 || rows > 256) return fail(MTX_ERR_ARG, "rows must be in [1, 256]");
  if (num_q_heads % num_kv_heads != 0 || num_q_heads / num_kv_heads > 16) return fail(MTX_ERR_UNSUPPORTED, "bad head grouping");
  if (tokens_per_page < 8 || (tokens_per_page & (tokens_per_page - 1)) != 0)
    return fail(MTX_ERR_UNSUPPORTED, "tokens_per_page must be a power of two >= 8");
  if (num_pages < 1 || max_pages_per_group < 1) return fail(MTX_ERR_ARG, "bad page counts");
  const uint64_t pool_rows = uint64_t(num_kv_heads) * num_pages * tokens_per_page;
  if (pool_rows >= (1ull << 31)) return fail(MTX_ERR_UNSUPPORTED, "the page pool has too many rows for one tensor map");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // a group's tokens are addressed as a "prefill segment" of up to max_tokens rows with an empty ring behind it
  const int P = max_pages_per_group * tokens_per_page, T = P + 64;
  const size_t mc = attn_max_chunks(P, T);
  const size_t G = num_q_heads / num_kv_heads;
  uint8_t* b = static_cast<uint8_t*>(scratch);
  AttnParams p;
  memset(&p, 0, sizeof(p));
  int* work_items = reinterpret_cast<int*>(b);
  b += align_up(size_t(rows) * mc * 4, 1024);
  int* work_count = reinterpret_cast<int*>(b);
  b += 1024;
  p.tickets = reinterpret_cast<int*>(b);
  const size_t ticket_bytes = align_up(size_t(rows) * num_kv_heads * 4, 1024);
  b += ticket_bytes;
  p.part_ml = reinterpret_cast<float*>(b);
  b += align_up(size_t(rows) * num_kv_heads * mc * G * 2 * 4, 1024);
  p.part_o = reinterpret_cast<float*>(b);
  b += align_up(size_t(rows) * num_kv_heads * mc * G * head_dim * 4, 1024);
  int* plane = reinterpret_cast<int*>(b);
  int* zeros_a = plane + rows;
  int* zeros_b = zeros_a + rows;
  MTX_CUDA(cudaMemsetAsync(p.tickets, 0, ticket_bytes, st));
  ragged_rows_kernel<<<(rows + 127) / 128, 128, 0, st>>>(plane, zeros_a, zeros_b, rows);
  const int tpi = attn_tiles_per_item(rows, num_kv_heads, P, T, 148);
  MTX_TRY(launch(attn_build_worklist_kernel, dim3(1), dim3(256), 0, st, (const int*)lengths, (const int*)zeros_a, (const int*)zeros_b, rows, P, T, tpi,
                 work_items, work_count));
  CUtensorMap tk, tv;
  const uint32_t box = tokens_per_page < kAttnTileRows ? tokens_per_page : kAttnTileRows;
  MTX_TRY(make_map(&tk, k_pages, head_dim, pool_rows, box));
  MTX_TRY(make_map(&tv, v_pages, head_dim, pool_rows, box));
  p.q = static_cast<const bf16*>(q);
  p.out = static_cast<bf16*>(out);
  p.plane = plane;
  p.len0 = lengths;
  p.ring_first = zeros_a;
  p.ring_len = zeros_b;
  p.work_items = work_items;
  p.work_count = work_count;
  p.rows = rows;
  p.hq = num_q_heads;
  p.hkv = num_kv_heads;
  p.P = P;
  p.T = T;
  p.tiles_per_item = tpi;
  p.max_chunks = int(mc);
  p.softcap = softcap;
  p.page_map = page_map;
  p.tokens_per_page = tokens_per_page;
  p.num_pages = num_pages;
  p.max_pages = max_pages_per_group;
  p.page_row_base = 0;
  const size_t smem = attn_smem_bytes(head_dim, int(G));
  int grid = rows * num_kv_heads * attn_max_chunks(P, T, tpi);
  const int cap = 148 * (head_dim == 64 ? 3 : 1);
  if (grid > cap) grid = cap;
  if (head_dim == 64) {
    MTX_CUDA(cudaFuncSetAttribute(decode_attn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    return launch(decode_attn_kernel<64>, dim3(grid), dim3(kAttnThreads), smem, st, tk, tv, p);
  }
  MTX_CUDA(cudaFuncSetAttribute(decode_attn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  return launch(decode_attn_kernel<128>, dim3(grid), dim3(kAttnThreads), smem, st, tk, tv, p);
}

int mtx_page_update_decode(int32_t* page_status, int32_t* page_map, int32_t* num_pages_used, int32_t* sequence_lengths, int32_t* active_page,
                           const int32_t* has_active_page, int32_t* active_page_position, int num_pages, int groups, int max_pages_per_group,
                           int tokens_per_page, mtx_stream stream) {
  if (!page_status || !page_map || !num_pages_used || !sequence_lengths || !active_page || !has_active_page || !active_page_position)
    return fail(MTX_ERR_ARG, "null argument");
  if (groups < 1 || groups > 256 || num_pages < 2 || max_pages_per_group < 1 || tokens_per_page < 1) return fail(MTX_ERR_ARG, "bad page state shape");
  PageStateDev a;
  a.page_status = page_status;
  a.page_map = page_map;
  a.num_pages_used = num_pages_used;
  a.sequence_lengths = sequence_lengths;
  a.active_page = active_page;
  a.has_active_page = has_active_page;
  a.active_page_position = active_page_position;
  a.num_pages = num_pages;
  a.groups = groups;
  a.max_pages_per_group = max_pages_per_group;
  a.tokens_per_page = tokens_per_page;
  return launch(page_update_decode_kernel, dim3(1), dim3(256), 0, static_cast<cudaStream_t>(stream), a);
}

int mtx_paged_insert(void* k_pages, void* v_pages, const void* k_src, const void* v_src, const int32_t* page_map_row, int layers,
                     int num_kv_heads, int head_dim, int n_src_rows, int n_tokens, int num_pages, int tokens_per_page, mtx_stream stream) {
  if (!k_pages || !v_pages || !k_src || !v_src || !page_map_row) return fail(MTX_ERR_ARG, "null argument");
  if (layers < 1 || num_kv_heads < 1 || head_dim % 8 != 0 || n_tokens < 1 || n_tokens > n_src_rows || num_pages < 1 || tokens_per_page < 1)
    return fail(MTX_ERR_ARG, "bad paged insert shape");
  PagedInsertArgs g;
  memset(&g, 0, sizeof(g));
  g.k_src = static_cast<const bf16*>(k_src);
  g.v_src = static_cast<const bf16*>(v_src);
  g.k_pages = static_cast<bf16*>(k_pages);
  g.v_pages = static_cast<bf16*>(v_pages);
  g.page_map_row = page_map_row;
  g.layers = layers;
  g.hkv = num_kv_heads;
  g.d = head_dim;
  g.src_rows = n_src_rows;
  g.n_tokens = n_tokens;
  g.num_pages = num_pages;
  g.tokens_per_page = tokens_per_page;
  const long long vecs = (long long)layers * num_kv_heads * n_tokens * (head_dim / 8);
  int grid = int((vecs + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  return launch(paged_insert_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), g);
}

// ---- fused QKV projection + RoPE + KV append ------------------------------------------------------

size_t mtx_qkv_rope_append_scratch_bytes(int rows, int head_dim) { return align_up(size_t(round_rows(rows)) * (head_dim / 2) * 8, 1024); }

int mtx_qkv_rope_append(const void* n, const void* wqkv, const int32_t* pos, const int32_t* plane, const int32_t* write_row, void* q_out,
                        void* k_cache, void* v_cache, int rows, int emb_dim, int num_q_heads, int num_kv_heads, int head_dim,
                        int rows_per_plane, float rope_min_timescale, float rope_max_timescale, void* scratch, mtx_stream stream) {
  if (!n || !wqkv || !pos || !plane || !write_row || !q_out || !k_cache || !v_cache || !scratch) return fail(MTX_ERR_ARG, "null argument");
  if (head_dim != 64 && head_dim != 128) return fail(MTX_ERR_UNSUPPORTED, "head_dim %d: only 64 and 128", head_dim);
  if (rows < 1 || rows > 256 || emb_dim % 64 != 0) return fail(MTX_ERR_ARG, "rows must be in [1, 256], emb_dim a multiple of 64");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int r_tile = round_rows(rows);
  const int qkv_n = (num_q_heads + 2 * num_kv_heads) * head_dim;
  float2* cs = static_cast<float2*>(scratch);
  const int half = head_dim / 2;
  rope_table_kernel<<<(rows * half + 127) / 128, 128, 0, st>>>(pos, cs, rows, half, rope_min_timescale, rope_max_timescale);
  CUtensorMap tw, tx;
  MTX_TRY(make_map(&tw, wqkv, emb_dim, qkv_n, kTileN));
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.n = qkv_n;
  p.k = emb_dim;
  p.rows = rows;
  p.r_tile = r_tile;
  EpiArgs ea;
  memset(&ea, 0, sizeof(ea));
  ea.q_out = static_cast<bf16*>(q_out);
  ea.k_cache = static_cast<bf16*>(k_cache);
  ea.v_cache = static_cast<bf16*>(v_cache);
  ea.plane = plane;
  ea.write_row = write_row;
  ea.rope_cs = cs;
  ea.hq = num_q_heads;
  ea.hkv = num_kv_heads;
  ea.d = head_dim;
  ea.t_alloc = rows_per_plane;
  if (use_rows_kernel(r_tile)) {
    const RowsPlan pl = plan_rows(qkv_n, emb_dim, r_tile, 148, EPI_QKV_ROPE);
    MTX_TRY(make_map(&tx, n, emb_dim, r_tile, pl.splits > 1 ? 128 : r_tile));
    return launch_rows<EPI_QKV_ROPE>(tw, tx, p, ea, pl, st);
  }
  MTX_TRY(make_map(&tx, n, emb_dim, r_tile, r_tile));
  return launch_gemm<EPI_QKV_ROPE>(tw, tx, p, ea, plan_gemm(qkv_n, emb_dim, r_tile, 148, EPI_QKV_ROPE), st);
}

// Re-attach the decode-state buffers (an XLA FFI caller receives them anew at every call): only what depends on the
// state pointers is rebuilt (the K/V tensor maps; captured graphs are dropped).
int mtx_engine_rebind_state(mtx_engine* e, const mtx_decode_state* s) {
  if (!e || !e->bound || !s) return fail(MTX_ERR_ARG, "engine is not bound / null state");
  if (memcmp(&e->s, s, sizeof(*s)) == 0) return MTX_OK;
  const mtx_model_config& c = e->cfg;
  // (checked before anything is changed: a quantised or paged engine has more tensor maps than the two rebuilt below)
  if (c.kv_quant) return fail(MTX_ERR_UNSUPPORTED, "mtx_engine_rebind_state with an int8 KV cache: bind again instead");
  if (is_paged(c)) return fail(MTX_ERR_UNSUPPORTED, "mtx_engine_rebind_state with a paged KV cache: bind again instead");
  const bool kv_moved = e->s.k_cache != s->k_cache || e->s.v_cache != s->v_cache;
  e->s = *s;
  for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
  e->graphs.clear();
  for (auto& kv : e->host_graphs) cudaGraphExecDestroy(kv.second);
  e->host_graphs.clear();
  if (kv_moved) {
    const uint64_t kv_rows = uint64_t(c.num_layers) * c.num_slots * c.num_kv_heads * c.max_target_len;
    MTX_TRY(make_map(&e->tm_k, s->k_cache, c.head_dim, kv_rows, kAttnTileRows));
    MTX_TRY(make_map(&e->tm_v, s->v_cache, c.head_dim, kv_rows, kAttnTileRows));
  }
  return MTX_OK;
}

}  // extern "C"
