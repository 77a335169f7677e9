// mlp_engine.cu -- the csb_mlp_* C ABI: dense-stack column emulator (MLP_v1 / ED / HSR-style nets) on one B200.
//
// Data layout in HBM (all row-major, feature dimension contiguous, every feature dimension padded to a multiple of
// 64 so that one 128-byte swizzle row == 64 bf16 and TMA boxes never straddle a row pitch):
//   params / grads / adam m,v : one flat fp32 buffer, per layer  W_l [Kp_l x Np_l] (Keras kernel layout) then b_l [Np_l]
//                               (then gamma_l, beta_l [Np_l] when LayerNorm) -- padding entries are zero and stay zero.
//   bf16 mode weight copies   : W16_l [Kp x Np] (operand of the data-gradient GEMM) and Wt16_l [Np x Kp] (operand of
//                               the forward GEMM), refreshed by the optimizer step.
//   activations               : xn [cap x Kp_0], act_l [cap x Np_l] (bf16 in CSB_BF16 mode, fp32 in CSB_F32 mode),
//                               kept for the backward pass; dZ ping-pong [cap x max Np]; pred fp32 [cap x Np_last].
//   split-K workspace         : per layer  splits_l x (Kp_l*Np_l) fp32 partials of dW_l and S x Np_l partials of db_l,
//                               reduced in a fixed order (deterministic) into `grads`.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "common.cuh"
#include "simt_kernels.cuh"
#include "fit_kernels.cuh"
#include "dp_kernels.cuh"
#include "tc_gemm.cuh"
#include "tail_kernel.cuh"

namespace csb {

static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------------------------------------------
// TMA descriptor creation through the driver entry point (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// bf16 matrix [rows, ld] with `cols` valid columns; box = box_cols x box_rows; SWIZZLE_128B (box_cols == 64)
static int make_tmap_bf16(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_cols,
                          uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  CSB_REQUIRE(fn != nullptr, CSB_ECUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CSB_REQUIRE(r == CUDA_SUCCESS, CSB_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (cols %llu rows %llu ld %llu box %ux%u)",
              (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld, box_cols, box_rows);
  return CSB_OK;
}

// fp32 matrix as a kind::tf32 operand: the 128-byte swizzle row holds 32 elements
static int make_tmap_f32(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  CSB_REQUIRE(fn != nullptr, CSB_ECUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CSB_REQUIRE(r == CUDA_SUCCESS, CSB_ECUDA, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d (cols %llu rows %llu ld %llu box 32x%u)",
              (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)ld, box_rows);
  return CSB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// tensor-core launch helpers
// ---------------------------------------------------------------------------------------------------------------
static bool g_use_pairs = true;     // CSB_NO_PAIRS=1 falls back to the single-CTA kernels (debugging aid)
static bool g_use_pdl = true;       // CSB_NO_PDL=1 launches every kernel fully serialised (debugging aid)
static bool g_use_tail = true;      // CSB_NO_TAIL_FUSION=1: output layer, its data gradient and its weight gradient as three launches (A/B aid)
static bool g_use_balanced = false; // CSB_BALANCED=1 (opt-in): balanced contiguous tile ranges instead of round-robin -- measured SLOWER

// launch with (optionally) programmatic stream serialisation: see pdl_wait() in common.cuh
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <int BN, int STAGES, int EPI, int CG, int VAR = 0>
static int launch_tn(const CUtensorMap& ta, const CUtensorMap& tb, const tc::GemmParams& p, int sm_count, cudaStream_t st,
                     const CUtensorMap* ta2 = nullptr) {
  constexpr bool BF16_OUT = (EPI == tc::EPI_BIAS_ACT || EPI == tc::EPI_HEAD_LOSS || EPI == tc::EPI_DGRAD || EPI == tc::EPI_DGRAD_MASK ||
                             EPI == tc::EPI_BIAS_ADD);
  using L = tc::TnSmem<BN, STAGES, CG, (VAR & tc::VAR_STAGED) != 0 && BF16_OUT>;
  auto kern = tc::gemm_tn_kernel<BN, STAGES, EPI, CG, VAR>;
  CSB_REQUIRE((EPI == tc::EPI_HEAD_LOSS || (VAR & tc::VAR_ACC2) ? 2 : 1) * (int)round_up(p.N, BN) <= tc::TN_BIAS_SMEM, CSB_EUNSUPPORTED,
              "layer width %d too large for the %d-float bias / loss-weight area in shared memory", p.N, tc::TN_BIAS_SMEM);
  static bool attr_set = false;
  if (!attr_set) {
    CSB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  int tiles = (int)(ceil_div(ceil_div(p.M, tc::BM), CG) * ceil_div(p.N, BN));   // CG m-blocks per tile
  if ((VAR & tc::VAR_KSPLIT) != 0) tiles *= std::max(1, p.k_splits);              // (split, m-group, n-block) grid
  if (tiles == 0) return CSB_OK;
  int grid = std::min(tiles, sm_count / CG) * CG;
  if (p.dbg >> 16) grid = std::min(grid, (p.dbg >> 16) * CG);      // micro-benchmark: run on a few SMs only (no power capping)
  tc::GemmParams q = p;
  q.b_box_rows = std::min(p.N, BN) / CG;     // must equal the box the B tensor map was encoded with
  q.balanced = (g_use_balanced && EPI != tc::EPI_HEAD_LOSS && EPI != tc::EPI_HEAD_OUT && EPI != tc::EPI_F32) ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(tc::TN_THREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CG; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  CSB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, ta, tb, q, ta2 ? *ta2 : ta));
  return CSB_OK;
}

// Tile-shape policy.  N > 128: 256-wide tiles; CTA pairs (cta_group::2, the B tensor map must then have been encoded
// with a box of min(N,256)/2 rows) whenever the layer is wide enough.
static inline bool tn_use_pairs(int N) { return g_use_pairs && N > 128; }
static inline int tn_b_box_rows(int N) { return N > 128 ? (tn_use_pairs(N) ? std::min(N, 256) / 2 : std::min(N, 256)) : std::min(N, 128); }

static bool g_use_staged = true;    // CSB_NO_STAGED_EPI=1: register -> global stores in every launch (debugging aid)
// Small batches: when the 256 x 256 pair tiles of a wide layer would occupy at most half of the SMs (e.g. the reference's own batch
// size 3072: 36 pairs), the layer runs on 128 x 128 single-CTA tiles instead -- four times as many CTAs, each a quarter of the
// mainloop.  `tb_small` is the B tensor map encoded with the 128-row box those tiles need (nullptr: policy off).
static inline bool tn_small_tiles(int M, int N, int sm_count, const CUtensorMap* tb_small) {
  static const bool off = getenv("CSB_NO_SMALL_TILES") != nullptr;          // debugging aid: always the 256-wide pair tiles
  if (off || tb_small == nullptr || N <= 128 || !tn_use_pairs(N)) return false;
  const int64_t pairs = ceil_div(ceil_div(M, 128), 2) * ceil_div(N, 256);
  return 4 * pairs <= sm_count;
}
template <int EPI, int VAR>
static int launch_tn_shape(const CUtensorMap& ta, const CUtensorMap& tb, const tc::GemmParams& p, int sm_count, cudaStream_t st,
                           const CUtensorMap* tb_small = nullptr) {
  constexpr bool BF16_OUT = (EPI == tc::EPI_BIAS_ACT || EPI == tc::EPI_HEAD_LOSS || EPI == tc::EPI_DGRAD || EPI == tc::EPI_DGRAD_MASK ||
                             EPI == tc::EPI_BIAS_ADD);
  if (tn_small_tiles(p.M, p.N, sm_count, tb_small)) {
    if constexpr (BF16_OUT) {
      if (g_use_staged && p.K <= 256 && p.kb_per_tap == 0) return launch_tn<128, 4, EPI, 1, VAR | tc::VAR_STAGED>(ta, *tb_small, p, sm_count, st);
    }
    return launch_tn<128, 6, EPI, 1, VAR>(ta, *tb_small, p, sm_count, st);
  }
  if constexpr (BF16_OUT) {
    // Short contractions (K <= 256: at most four k-blocks per tile) are bound by the epilogue's stores, not by the mainloop:
    // they trade operand-ring stages for an output staging tile and coalesced stores.
    if (g_use_staged && p.K <= 256 && p.kb_per_tap == 0) {
      constexpr int V = VAR | tc::VAR_STAGED;
      if (p.N > 128) {
        if (tn_use_pairs(p.N)) return launch_tn<256, 4, EPI, 2, V>(ta, tb, p, sm_count, st);
        return launch_tn<256, 2, EPI, 1, V>(ta, tb, p, sm_count, st);
      }
      return launch_tn<128, 4, EPI, 1, V>(ta, tb, p, sm_count, st);
    }
  }
  if (p.N > 128) {
    if (tn_use_pairs(p.N)) return launch_tn<256, 6, EPI, 2, VAR>(ta, tb, p, sm_count, st);
    return launch_tn<256, 4, EPI, 1, VAR>(ta, tb, p, sm_count, st);
  }
  return launch_tn<128, 6, EPI, 1, VAR>(ta, tb, p, sm_count, st);
}
template <int EPI>
static int launch_tn_auto(const CUtensorMap& ta, const CUtensorMap& tb, const tc::GemmParams& p, int sm_count, cudaStream_t st,
                          const CUtensorMap* tb_small = nullptr) {
  // ELU (expm1f in the epilogue) is a separate instantiation of the epilogues that evaluate an activation
  constexpr bool HAS_ACT = (EPI == tc::EPI_BIAS_ACT || EPI == tc::EPI_HEAD_LOSS || EPI == tc::EPI_HEAD_OUT || EPI == tc::EPI_DGRAD);
  if constexpr (EPI == tc::EPI_HEAD_LOSS) {
    // the lean loss loop covers MSE without an output mask on a non-ELU head; everything else takes the general variant
    if (p.act == CSB_ACT_ELU || p.loss_kind != CSB_LOSS_MSE || p.out_mask != nullptr)
      return p.act == CSB_ACT_ELU ? launch_tn_shape<EPI, tc::VAR_ELU | tc::VAR_GENERAL_LOSS>(ta, tb, p, sm_count, st, tb_small)
                                  : launch_tn_shape<EPI, tc::VAR_GENERAL_LOSS>(ta, tb, p, sm_count, st, tb_small);
  } else if constexpr (HAS_ACT) {
    if (p.act == CSB_ACT_ELU) return launch_tn_shape<EPI, tc::VAR_ELU>(ta, tb, p, sm_count, st, tb_small);
  }
  return launch_tn_shape<EPI, 0>(ta, tb, p, sm_count, st, tb_small);
}
static inline int tn_block_n(int N) { return N > 128 ? 256 : 128; }

// kind::tf32 launches (CSB_TF32): 256-wide pair tiles for layers wider than 128 columns, else 128-wide single-CTA tiles; the B tensor
// map must have been encoded with tf32_b_box(N) rows
template <int EPI, int VARX = 0>
static int launch_tn_tf32(const CUtensorMap& ta, const CUtensorMap& tb, const tc::GemmParams& p, int sm_count, cudaStream_t st) {
  constexpr int V = tc::VAR_TF32 | VARX;
  if (g_use_pairs && p.N > 128) return launch_tn<256, 6, EPI, 2, V>(ta, tb, p, sm_count, st);
  return launch_tn<128, 6, EPI, 1, V>(ta, tb, p, sm_count, st);
}

template <int BN, int STAGES, int CG>
static int launch_nt(const CUtensorMap& ta, const CUtensorMap& tb, const tc::NtParams& p, int splits, cudaStream_t st) {
  using L = tc::NtSmem<BN, STAGES, CG>;
  auto kern = tc::gemm_nt_kernel<BN, STAGES, CG>;
  static bool attr_set = false;
  if (!attr_set) {
    CSB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  // pairs: every n-block's width splits into two halves of whole 8-column pieces that are valid cta_group::2 MMA widths (N % 32 == 0);
  // the halves need not be 64-column-chunk aligned in global memory (the TMA box of the second CTA simply starts mid-chunk)
  CSB_REQUIRE(CG == 1 || p.N % 32 == 0, CSB_EUNSUPPORTED, "CTA-pair weight-gradient tiles need N %% 32 == 0 (N = %d)", p.N);
  const unsigned m_tiles = (unsigned)ceil_div(ceil_div(p.M, tc::BM), CG), n_blocks = (unsigned)ceil_div(p.N, BN);
  const unsigned tiles = m_tiles * n_blocks;                                                     // CG m-blocks per tile
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles * CG, (unsigned)splits);
  if (p.splits_narrow > 0) {      // uneven split counts: one CTA group per (split, full-width tile) and per (narrow split, narrow tile)
    CSB_REQUIRE(n_blocks >= 2 && p.splits == splits && p.splits_narrow <= splits && 2 * p.splits_narrow >= splits, CSB_EINVAL,
                "inconsistent uneven split geometry (%d, %d, %d)", splits, p.splits, p.splits_narrow);
    cfg.gridDim = dim3((m_tiles * (n_blocks - 1) * (unsigned)splits + m_tiles * (unsigned)p.splits_narrow) * CG, 1);
  }
  cfg.blockDim = dim3(tc::NUM_THREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CG; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  CSB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  return CSB_OK;
}
// Weight-gradient tile policy: CTA pairs (256 x BN tiles, cg == 2) whenever the layer has at least two 128-row m-blocks and
// its width splits into two 64-column-chunk-aligned halves; an odd m-block count leaves half of the last pair's rows empty
// (zero-filled by TMA, never stored), which still beats single CTAs that cannot take their operands in fast enough.
static bool g_use_nt_pairs = true;  // CSB_NO_NT_PAIRS=1: single-CTA weight-gradient tiles (debugging aid)
// CSB_NT_NARROW=1 (opt-in): uneven split counts for a half-width last n-block (NtParams.splits_narrow).  Correct (tests/test_gemm_gpu.py)
// but measured slower on the MLP_v1 step on one B200 -- 0.913 against 0.890 ms with the 2/3 rule, 0.902 against 0.860 ms with 1/2 --
// so the default keeps one split count per layer.  Likely cause: with one split count all tiles of a split walk the same row range
// together, so an activation / dZ row block is fetched from HBM once and hit in L2 by the sibling tiles; narrow tiles with their own
// (longer) row ranges read their operands at other times and re-fetch them from HBM (+~100 MB on a launch that already moves 130-180 MB).
static bool g_use_nt_narrow = false;
static inline int nt_cta_group(int M, int N) { return (g_use_nt_pairs && N % 128 == 0 && N > 128 && M > 128) ? 2 : 1; }   // 128-wide layers: measured no gain
// Splits of a half-width tile for S splits of the full-width ones: a row block of the narrow tile costs half the MMA cycles but
// 3/4 of the operand bytes (the H^T tile is not amortised over fewer columns), and at 96 B/clk it is bound by what an SM can take
// in (~69 B/clk): about 2/3 of a full-width row block.  
static inline int nt_narrow_splits(int S) { return std::max((S + 1) / 2, (2 * S + 1) / 3); }
static inline int nt_m_tiles(int M, int cg) { return (int)ceil_div(ceil_div(M, 128), cg); }
static int launch_nt_auto(const CUtensorMap& ta, const CUtensorMap& tb, const tc::NtParams& p, int splits, cudaStream_t st, int cg = 1) {
  if (cg == 2) {
    if (p.N > 128) return launch_nt<256, 6, 2>(ta, tb, p, splits, st);
    return launch_nt<128, 8, 2>(ta, tb, p, splits, st);
  }
  if (p.N > 128) return launch_nt<256, 4, 1>(ta, tb, p, splits, st);
  return launch_nt<128, 6, 1>(ta, tb, p, splits, st);
}

static int grid_for(int64_t work_items, int threads, int sm_count, int per_thread = 1) {
  int64_t blocks = ceil_div(work_items, (int64_t)threads * per_thread);
  return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)sm_count * 8));
}

}  // namespace csb

using namespace csb;

// ---------------------------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------------------------
struct LayerInfo {
  int K, N, Kp, Np;
  int act;
  float alpha;
  int ln;
  size_t w_off, b_off, g_off;  // offsets into the padded flat buffers (g_off: gamma [Np] then beta [Np], LayerNorm layers)
  size_t w_off_user, b_off_user, g_off_user;
  size_t ws_g_off;             // LayerNorm parameter-gradient partials [LN_SPLITS][2][Np]
  // split-K workspace
  size_t ws_w_off, ws_b_off;
  int max_w_splits, b_splits;
  int split_cap = 64;      // most weight-gradient splits a step uses (the fused optimizer walks a tile's partials serially)
  int nt_block_n;
  int nt_cg;                   // CTAs per weight-gradient tile (2: cta_group::2 pairs)
  bool nt_narrow;              // half-width last n-block: uneven split counts (NtParams.rb_per_split_narrow)
};

struct ActMaps {                 // TMA descriptors that depend on the batch size
  CUtensorMap a_k128;            // buffer as K-major A operand (box 64 x 128 rows)
  CUtensorMap mn64;              // buffer as MN-major operand of the weight-gradient GEMM (box 64 x 64 rows)
};

enum ProfKind { K_BEGIN = 0, K_NORMALIZE, K_GEMM_FWD, K_GEMM_HEAD, K_LOSS, K_GEMM_WGRAD, K_COLSUM, K_GEMM_DGRAD, K_REDUCE, K_OPT,
                K_REPACK, K_MISC, K_COUNT };
static const char* kProfNames[K_COUNT] = {"begin", "normalize", "gemm_tn_fwd", "gemm_tn_head", "loss", "gemm_nt_wgrad", "colsum_bias_grad",
                                          "gemm_tn_dgrad", "reduce_partials", "optimizer", "repack_bf16", "misc"};

struct csb_mlp {
  csb_mlp_cfg cfg;
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_kind;
  int prof_used = 0;
  double prof_ms[K_COUNT] = {};
  int64_t prof_n[K_COUNT] = {};
  int L = 0;
  LayerInfo layer[CSB_MAX_LAYERS];
  int in_dim = 0, in_p = 0, out_dim = 0, out_p = 0, max_np = 0;
  int64_t cap = 0;
  size_t P_pad = 0, P_user = 0, ws_elems = 0;
  int sm_count = 0;
  bool bf16 = false;

  float *params = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr, *ws = nullptr;
  __nv_bfloat16* w16[CSB_MAX_LAYERS] = {};
  __nv_bfloat16* wt16[CSB_MAX_LAYERS] = {};
  void* xn = nullptr;                       // [cap x in_p]   bf16 or fp32
  void* act[CSB_MAX_LAYERS] = {};           // [cap x Np_l]   (l < L-1)
  void* dz[2] = {nullptr, nullptr};         // [cap x max_np]
  uint32_t* amask[CSB_MAX_LAYERS] = {};     // bf16 mode, ReLU / LeakyReLU layers without LayerNorm: sign bits of act_l [cap x Np_l/32]
  void* zbuf[CSB_MAX_LAYERS] = {};          // LayerNorm layers: pre-norm z [cap x Np_l]
  float* ln_stats[CSB_MAX_LAYERS] = {};     // LayerNorm layers: (mean, rstd) per row [cap x 2]
  float* pred = nullptr;                    // [cap x out_p]
  float* dx_tmp = nullptr;                  // [cap x in_p] (lazily allocated)
  float *d_sub = nullptr, *d_div = nullptr, *d_out_scale = nullptr, *d_inv_out_scale = nullptr, *d_loss_w = nullptr, *d_out_mask = nullptr;
  bool has_mask = false;
  float* d_xform = nullptr;                 // csb_mlp_set_input_transform: [4][in_p] lambda, keep, clip_lo, clip_hi (NULL: plain normalisation)
  float *loss_partials = nullptr, *d_loss = nullptr;
  int n_loss_partials = 0;
  // device staging for the *_host entry points: two slots, filled on an internal copy stream so that the H2D copy of
  // batch i+1 overlaps the compute of batch i (csb_mlp_stage_host_batch / csb_mlp_train_step_host_async)
  float *x_stage[2] = {nullptr, nullptr}, *y_stage[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_released[2] = {nullptr, nullptr};
  bool released_valid[2] = {false, false};
  int stage_next = 0, staged_cur = -1;

  CUtensorMap tm_wt[CSB_MAX_LAYERS];        // Wt16_l as K-major B operand of the forward GEMM
  CUtensorMap tm_w[CSB_MAX_LAYERS];         // W16_l  as K-major B operand of the data-gradient GEMM
  CUtensorMap tm_wt_s[CSB_MAX_LAYERS], tm_w_s[CSB_MAX_LAYERS];     // the same two with a 128-row box (small-batch tile policy)
  int64_t maps_B = -1;
  // CSB_TF32: fp32 buffers as kind::tf32 operands.  wt32_l = W_l^T [Np, Kp] (B operand of the forward GEMM; W_l itself, [Kp, Np], is the
  // K-major B operand of the data gradient); tr_a / tr_b = transposed copies of a layer's input and dZ ([features, batch]: the weight
  // gradient contracts over the batch)
  bool tf32 = false;
  // CSB_TF32X3 (x3): nothing is rounded; every GEMM contracts over the tripled operands [hi | lo | hi] x [hi | hi | lo] written by
  // split3_f32_kernel (sp_a: the A-side copy of the current layer input / dZ; the weight copies and the transposed weight-gradient
  // operands carry their three blocks side by side), i.e. three kind::tf32 products per fp32 product: fp32-class results
  bool x3 = false;
  float* sp_a = nullptr;
  // CSB_TF32 (not x3): persistent [features, batch] copies written by the GEMM epilogues themselves (GemmParams.out_t): xn^T, act_l^T,
  // dZ^T (ping-pong like dz).  A copy the epilogue could not write (LayerNorm layers: the activation comes from ln_fwd and dZ from
  // ln_bwd; dropout: the activation changes after the GEMM; the output layer's dZ) is produced by transpose_f32_kernel into the SAME buffer.
  float* xn_t = nullptr;
  float* act_t[CSB_MAX_LAYERS] = {};
  float* dz_t[2] = {nullptr, nullptr};
  bool act_t_valid[CSB_MAX_LAYERS] = {};
  bool dz_t_valid = false;                  // the transposed copy of the dZ the backward chain currently holds is valid
  float* wt32[CSB_MAX_LAYERS] = {};
  float* w32r[CSB_MAX_LAYERS] = {};        // W_l rounded to the TF32 grid, [Kp, Np] (the fp32 master weights stay exact)
  float *tr_a = nullptr, *tr_b = nullptr;
  int64_t tr_ld = 0;
  CUtensorMap tm32_wt[CSB_MAX_LAYERS], tm32_w[CSB_MAX_LAYERS];
  CUtensorMap tm32_in[CSB_MAX_LAYERS], tm32_dz[CSB_MAX_LAYERS], tm32_tra[CSB_MAX_LAYERS], tm32_trb[CSB_MAX_LAYERS];
  ActMaps tm_in[CSB_MAX_LAYERS];            // input of layer l (xn or act[l-1]) with `maps_B` rows
  ActMaps tm_dz[CSB_MAX_LAYERS];            // dZ_l (ping-pong buffer (L-1-l)&1, ld = Np_l) with `maps_B` rows
  CUtensorMap tm_z[CSB_MAX_LAYERS];         // LayerNorm layers: pre-norm z buffer (TMA-store target of the forward GEMM)

  int64_t step = 0, launches = 0;
  // Dropout behind every hidden layer's activation (hsr.py:20-25, online mlp.py:41-45: Linear -> [LayerNorm] -> Dropout -> ReLU, and
  // relu(dropout(u)) == dropout(relu(u))): training forwards only, counter-based keep decisions keyed by (seed, training forward, layer,
  // element) -- no mask is stored, the backward pass recognises a dropped element by its saved output being exactly 0
  float dropout = 0.f;
  uint32_t drop_seed = 0;
  int64_t drop_fwd = 0;                     // training forwards so far (advances the masks)
  bool drop_live = false;                   // the activations in the handle come from a forward that dropped
  struct GraphEntry { const float* x; const float* y; int64_t B; float gs; uint32_t flags; float* loss_out; int64_t maps_B;
                      cudaGraphExec_t exec; int64_t n_launches; uint64_t use; };
  std::vector<GraphEntry> graphs;           // cached CUDA graphs of the training step (LRU, 8 entries)
  uint64_t graph_clock = 0;
  bool graphs_on = true;
  cudaStream_t cap_stream = nullptr;
  // optional: weight-gradient GEMMs on a side stream, concurrently with the data-gradient GEMM of the same layer (both only read dZ_l):
  // the CTAs of one fill the SMs the other leaves idle in its last wave.  ev_dz[l]: dZ_l complete (main stream);
  // ev_w[l]: dW_l partials complete (side stream)
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_dz[CSB_MAX_LAYERS] = {}, ev_w[CSB_MAX_LAYERS] = {};
  bool side_on = false;
  // CSB_TRAIN_FUSED_OPT: the split partials of the last training step are still unreduced; csb_mlp_apply_opt consumes them
  bool pending = false;
  bool pending_tail = false;                // the pending partials of the output layer were written by tail_kernel (one per CTA)
  int64_t pending_B = 0;
  int pending_n_loss = 0;
  float* pending_loss_out = nullptr;
  // data parallelism over NVLink peer memory (dp_kernels.cuh): `grads` then lives at the head of an IPC-exported slab
  bool dp_on = false;
  int dp_rank = 0, dp_world = 1;
  int64_t dp_n = 0;                         // floats in grads / gsum (P_pad + 4)
  void* dp_peer_base[simt::DP_MAX_RANKS] = {};
  unsigned long long dp_epoch = 0;
  int64_t acts_B = -1;                      // batch of the last forward that kept activations
  bool acts_normalized = false;
};

// ---- per-launch device timing (optional): an event is recorded after every kernel launch; the time attributed to a
// launch is the interval since the previous mark on the same stream (kernel duration + launch gap).
static void prof_mark(csb_mlp* h, int kind, cudaStream_t st) {
  if (kind != K_BEGIN) h->launches++;
  if (!h->prof_on) return;
  if (h->prof_used == (int)h->prof_ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return; }
    h->prof_ev.push_back(e);
    h->prof_kind.push_back(kind);
  }
  h->prof_kind[h->prof_used] = kind;
  cudaEventRecord(h->prof_ev[h->prof_used], st);
  h->prof_used++;
}

static inline void* layer_in(csb_mlp* h, int l) { return l == 0 ? h->xn : h->act[l - 1]; }
static inline size_t esize(const csb_mlp* h) { return h->bf16 ? 2 : 4; }

static void free_all(csb_mlp* h) {
  auto F = [](void* p) { if (p) cudaFree(p); };
  F(h->params); F(h->grads); F(h->m); F(h->v); F(h->ws);
  for (int l = 0; l < CSB_MAX_LAYERS; ++l) { F(h->w16[l]); F(h->wt16[l]); F(h->wt32[l]); F(h->w32r[l]); F(h->act[l]); F(h->zbuf[l]); F(h->ln_stats[l]); F(h->amask[l]); }
  F(h->tr_a); F(h->tr_b); F(h->sp_a); F(h->xn_t); F(h->dz_t[0]); F(h->dz_t[1]);
  for (int l = 0; l < CSB_MAX_LAYERS; ++l) F(h->act_t[l]);
  F(h->xn); F(h->dz[0]); F(h->dz[1]); F(h->pred); F(h->dx_tmp);
  F(h->d_sub); F(h->d_div); F(h->d_out_scale); F(h->d_inv_out_scale); F(h->d_loss_w); F(h->d_out_mask);
  F(h->loss_partials); F(h->d_loss); F(h->d_xform);
  for (int i = 0; i < 2; ++i) {
    F(h->x_stage[i]); F(h->y_stage[i]);
    if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
    if (h->ev_released[i]) cudaEventDestroy(h->ev_released[i]);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (int l = 0; l < CSB_MAX_LAYERS; ++l) {
    if (h->ev_dz[l]) cudaEventDestroy(h->ev_dz[l]);
    if (h->ev_w[l]) cudaEventDestroy(h->ev_w[l]);
  }
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  for (int r = 0; r < simt::DP_MAX_RANKS; ++r)
    if (h->dp_peer_base[r] && r != h->dp_rank) cudaIpcCloseMemHandle(h->dp_peer_base[r]);
}

#define CSB_ALLOC(ptr, bytes)                                                                          \
  do {                                                                                                 \
    if (cudaMalloc(reinterpret_cast<void**>(&(ptr)), (bytes)) != cudaSuccess) {                        \
      set_last_error("cudaMalloc of %zu bytes failed (%s)", (size_t)(bytes), #ptr);                    \
      cudaGetLastError();                                                                              \
      return CSB_ENOMEM;                                                                               \
    }                                                                                                  \
    if (cudaMemset((ptr), 0, (bytes)) != cudaSuccess) { set_last_error("cudaMemset failed"); return CSB_ECUDA; } \
  } while (0)

static inline int transpose_grid(const csb_mlp* h, int64_t rows, int cols) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(rows, 32) * ceil_div(cols, 32), (int64_t)h->sm_count * 16));
}
static int repack_weights(csb_mlp* h, cudaStream_t st) {
  if (h->tf32) {                       // W_l [Kp, Np] -> wt32_l [Np, Kp] and w32r_l [Kp, Np], both rounded to the TF32 grid
    for (int l = 0; l < h->L && h->x3; ++l) {      // [hi | hi | lo] copies: wt32_l [Np, 3 Kp] (transposed), w32r_l [Kp, 3 Np]
      const LayerInfo& li = h->layer[l];
      simt::split3_f32_kernel<<<transpose_grid(h, li.Kp, li.Np), 256, 0, st>>>(h->params + li.w_off, li.Np, li.Kp, li.Np, h->wt32[l], 3 * (int64_t)li.Kp, li.Kp, 1, 1, li.Kp);
      simt::split3_f32_kernel<<<transpose_grid(h, li.Kp, li.Np), 256, 0, st>>>(h->params + li.w_off, li.Np, li.Kp, li.Np, h->w32r[l], 3 * (int64_t)li.Np, li.Np, 0, 1, li.Kp);
      CSB_CUDA_CHECK(cudaGetLastError());
    }
    for (int l = 0; l < h->L && !h->x3; ++l) {
      const LayerInfo& li = h->layer[l];
      simt::transpose_f32_kernel<<<transpose_grid(h, li.Kp, li.Np), 256, 0, st>>>(h->params + li.w_off, li.Np, li.Kp, li.Np, h->wt32[l], li.Kp, 1);
      simt::transpose_f32_kernel<<<transpose_grid(h, li.Kp, li.Np), 256, 0, st>>>(h->params + li.w_off, li.Np, li.Kp, li.Np, h->w32r[l], 0, 1);
      CSB_CUDA_CHECK(cudaGetLastError());
    }
    prof_mark(h, K_REPACK, st);
    return CSB_OK;
  }
  if (!h->bf16) return CSB_OK;
  simt::RepackTable tab;
  tab.n = h->L;
  int max_tiles = 1;
  for (int l = 0; l < h->L; ++l) {
    const LayerInfo& li = h->layer[l];
    tab.l[l] = {h->params + li.w_off, h->w16[l], h->wt16[l], li.Kp, li.Np};
    max_tiles = std::max(max_tiles, (li.Kp / 32) * (li.Np / 32));
  }
  dim3 grid((unsigned)std::min(max_tiles, 4 * h->sm_count), (unsigned)h->L);
  CSB_CUDA_CHECK(launch_pdl(simt::repack_kernel, grid, dim3(256), 0, st, tab));
  prof_mark(h, K_REPACK, st);
  return CSB_OK;
}

static int build_weight_maps(csb_mlp* h) {
  for (int l = 0; l < h->L; ++l) {
    const LayerInfo& li = h->layer[l];
    // forward:  D[B, Np] = in[B, Kp] . Wt16[Np, Kp]^T          B-operand rows = Np, contraction = Kp
    int rc = make_tmap_bf16(&h->tm_wt[l], h->wt16[l], li.Kp, li.Np, li.Kp, 64, (uint32_t)tn_b_box_rows(li.Np));
    if (rc) return rc;
    // dgrad:    D[B, Kp] = dZ[B, Np] . W16[Kp, Np]^T           B-operand rows = Kp, contraction = Np
    rc = make_tmap_bf16(&h->tm_w[l], h->w16[l], li.Np, li.Kp, li.Np, 64, (uint32_t)tn_b_box_rows(li.Kp));
    if (rc) return rc;
    rc = make_tmap_bf16(&h->tm_wt_s[l], h->wt16[l], li.Kp, li.Np, li.Kp, 64, (uint32_t)std::min(li.Np, 128));
    if (rc) return rc;
    rc = make_tmap_bf16(&h->tm_w_s[l], h->w16[l], li.Np, li.Kp, li.Np, 64, (uint32_t)std::min(li.Kp, 128));
    if (rc) return rc;
  }
  return CSB_OK;
}

static inline __nv_bfloat16* dz16(csb_mlp* h, int l) { return reinterpret_cast<__nv_bfloat16*>(h->dz[(h->L - 1 - l) & 1]); }
static inline float* dz32(csb_mlp* h, int l) { return reinterpret_cast<float*>(h->dz[(h->L - 1 - l) & 1]); }

// B-operand box rows of the kind::tf32 launches: 256-wide pair tiles when the layer is wider than 128 columns, else 128-wide single-CTA tiles
static inline bool tf32_pairs(int N) { return g_use_pairs && N > 128; }
static inline uint32_t tf32_b_box(int N) { return (uint32_t)(tf32_pairs(N) ? std::min(N, 256) / 2 : std::min(N, 128)); }

static int build_act_maps(csb_mlp* h, int64_t B) {
  if (h->tf32 && h->maps_B != B) {
    const int64_t bb = round_up(B, 32);               // x3: the three [features, batch] blocks are bb columns apart
    for (int l = 0; l < h->L; ++l) {
      const LayerInfo& li = h->layer[l];
      int rc;
      if (h->x3) {
        rc = make_tmap_f32(&h->tm32_in[l], h->sp_a, 3 * (uint64_t)li.Kp, B, 3 * (uint64_t)li.Kp, 128);
        if (!rc) rc = make_tmap_f32(&h->tm32_dz[l], h->sp_a, 3 * (uint64_t)li.Np, B, 3 * (uint64_t)li.Np, 128);
        if (!rc) rc = make_tmap_f32(&h->tm32_tra[l], h->tr_a, 3 * (uint64_t)bb, li.Kp, 3 * (uint64_t)bb, 128);
        if (!rc) rc = make_tmap_f32(&h->tm32_trb[l], h->tr_b, 3 * (uint64_t)bb, li.Np, 3 * (uint64_t)bb, tf32_b_box(li.Np));
      } else {
        rc = make_tmap_f32(&h->tm32_in[l], layer_in(h, l), li.Kp, B, li.Kp, 128);
        if (!rc) rc = make_tmap_f32(&h->tm32_dz[l], dz32(h, l), li.Np, B, li.Np, 128);
        // transposed copies [features, batch]: A operand rows = Kp (box 128), B operand rows = Np
        if (!rc) rc = make_tmap_f32(&h->tm32_tra[l], l == 0 ? h->xn_t : h->act_t[l - 1], B, li.Kp, h->tr_ld, 128);
        if (!rc) rc = make_tmap_f32(&h->tm32_trb[l], h->dz_t[(h->L - 1 - l) & 1], B, li.Np, h->tr_ld, tf32_b_box(li.Np));
      }
      if (rc) return rc;
    }
    h->maps_B = B;
    return CSB_OK;
  }
  if (!h->bf16 || h->maps_B == B) return CSB_OK;
  for (int l = 0; l < h->L; ++l) {
    const LayerInfo& li = h->layer[l];
    int rc = make_tmap_bf16(&h->tm_in[l].a_k128, layer_in(h, l), li.Kp, B, li.Kp, 64, 128);
    if (rc) return rc;
    rc = make_tmap_bf16(&h->tm_in[l].mn64, layer_in(h, l), li.Kp, B, li.Kp, 64, 64);
    if (rc) return rc;
    rc = make_tmap_bf16(&h->tm_dz[l].a_k128, dz16(h, l), li.Np, B, li.Np, 64, 128);
    if (rc) return rc;
    rc = make_tmap_bf16(&h->tm_dz[l].mn64, dz16(h, l), li.Np, B, li.Np, 64, 64);
    if (rc) return rc;
    if (li.ln) {
      rc = make_tmap_bf16(&h->tm_z[l], h->zbuf[l], li.Np, B, li.Np, 64, 128);
      if (rc) return rc;
    }
  }
  h->maps_B = B;
  return CSB_OK;
}

extern "C" {

static int flush_pending(csb_mlp* h, cudaStream_t st);   // reduces gradient partials left behind by CSB_TRAIN_FUSED_OPT

int csb_version(void) { return CSB_VERSION; }

const char* csb_strerror(int code) {
  switch (code) {
    case CSB_OK: return "ok";
    case CSB_EINVAL: return "invalid argument";
    case CSB_ENODEV: return "no sm_100 CUDA device";
    case CSB_ENOMEM: return "device allocation failed";
    case CSB_ECUDA: return "CUDA call or kernel launch failed";
    case CSB_ESTATE: return "call order / state violation";
    case CSB_EUNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}
const char* csb_last_error(void) { return g_err; }

int csb_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* hbm_bytes) {
  int dev = 0, n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); set_last_error("no CUDA device"); return CSB_ENODEV; }
  CSB_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CSB_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (hbm_bytes) *hbm_bytes = prop.totalGlobalMem;
  return CSB_OK;
}

int csb_mlp_create(const csb_mlp_cfg* cfg, csb_mlp** out) {
  CSB_REQUIRE(cfg && out, CSB_EINVAL, "null argument");
  *out = nullptr;
  CSB_REQUIRE(cfg->n_layers >= 1 && cfg->n_layers <= CSB_MAX_LAYERS, CSB_EINVAL, "n_layers %d out of range", cfg->n_layers);
  CSB_REQUIRE(cfg->in_dim >= 1 && cfg->max_batch >= 1, CSB_EINVAL, "in_dim / max_batch must be positive");
  CSB_REQUIRE(cfg->dtype == CSB_F32 || cfg->dtype == CSB_BF16 || cfg->dtype == CSB_TF32 || cfg->dtype == CSB_TF32X3, CSB_EINVAL, "unknown dtype %d",
              cfg->dtype);
  CSB_REQUIRE(cfg->loss >= CSB_LOSS_MSE && cfg->loss <= CSB_LOSS_HUBER, CSB_EINVAL, "unknown loss %d", cfg->loss);
  int sm = 0, maj = 0, min = 0;
  int rc = csb_device_info(&sm, &maj, &min, nullptr);
  if (rc) return rc;
  CSB_REQUIRE(maj == 10, CSB_ENODEV, "device has compute capability %d.%d; this library is sm_100a only", maj, min);
  for (int l = 0; l < cfg->n_layers; ++l) {
    CSB_REQUIRE(cfg->units[l] >= 1, CSB_EINVAL, "units[%d] must be positive", l);
    CSB_REQUIRE(cfg->act[l] >= CSB_ACT_NONE && cfg->act[l] <= CSB_ACT_LEAKYRELU, CSB_EINVAL, "act[%d] unknown", l);
    CSB_REQUIRE(cfg->layernorm[l] == 0 || l + 1 < cfg->n_layers, CSB_EUNSUPPORTED, "layernorm on the output layer is not supported");
  }

  g_use_pairs = getenv("CSB_NO_PAIRS") == nullptr;
  g_use_nt_pairs = getenv("CSB_NO_NT_PAIRS") == nullptr;
  g_use_nt_narrow = getenv("CSB_NT_NARROW") != nullptr;
  g_use_staged = getenv("CSB_NO_STAGED_EPI") == nullptr;
  g_use_balanced = getenv("CSB_BALANCED") != nullptr;
  g_use_tail = getenv("CSB_NO_TAIL_FUSION") == nullptr;
  g_use_pdl = getenv("CSB_NO_PDL") == nullptr;
  csb_mlp* h = new (std::nothrow) csb_mlp();
  CSB_REQUIRE(h != nullptr, CSB_ENOMEM, "host allocation failed");
  h->graphs_on = getenv("CSB_NO_GRAPHS") == nullptr;
  h->cfg = *cfg;
  h->L = cfg->n_layers;
  h->sm_count = sm;
  h->bf16 = cfg->dtype == CSB_BF16;
  h->tf32 = cfg->dtype == CSB_TF32 || cfg->dtype == CSB_TF32X3;      // fp32 storage (every `!bf16` path below), GEMMs on the tensor cores
  h->x3 = cfg->dtype == CSB_TF32X3;
  // opt-in (CSB_CONCURRENT_WGRAD=1): measured 0.896 against 0.875-0.89 ms/step for the plain in-order chain on one B200 -- the
  // cross-stream dependencies cost the programmatic-launch overlap that the single stream has, and both kernels want every SM
  if (h->bf16 && getenv("CSB_CONCURRENT_WGRAD") != nullptr) {         // side stream for the weight-gradient GEMMs (see csb_mlp)
    bool ok = cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int l = 0; ok && l < h->L; ++l)
      ok = cudaEventCreateWithFlags(&h->ev_dz[l], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h->ev_w[l], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { cudaGetLastError(); }
    h->side_on = ok;
  }
  h->in_dim = cfg->in_dim;
  h->in_p = (int)round_up(cfg->in_dim, 64);
  h->cap = round_up(cfg->max_batch, 128);
  size_t off = 0, off_user = 0, ws_off = 0;
  int k = cfg->in_dim;
  for (int l = 0; l < h->L; ++l) {
    LayerInfo& li = h->layer[l];
    li.K = k; li.N = cfg->units[l];
    li.Kp = (int)round_up(li.K, 64); li.Np = (int)round_up(li.N, 64);
    li.act = cfg->act[l]; li.alpha = cfg->alpha[l]; li.ln = cfg->layernorm[l];
    li.w_off = off; off += (size_t)li.Kp * li.Np;
    li.b_off = off; off += (size_t)li.Np;
    li.g_off = off; if (li.ln) off += 2 * (size_t)li.Np;
    li.w_off_user = off_user; off_user += (size_t)li.K * li.N;
    li.b_off_user = off_user; off_user += (size_t)li.N;
    li.g_off_user = off_user; if (li.ln) off_user += 2 * (size_t)li.N;
    li.nt_block_n = tn_block_n(li.Np);
    li.nt_cg = h->bf16 ? nt_cta_group(li.Kp, li.Np) : 1;
    const int tiles = nt_m_tiles(li.Kp, li.nt_cg) * (int)ceil_div(li.Np, li.nt_block_n);
    // one wave of CTAs, but at most 64 partials: the reduction walks a tile's partials serially
    li.max_w_splits = h->bf16 ? std::max(1, std::min(64, (sm / li.nt_cg) / tiles)) : 1;
    // a layer that is ONE 128 x 128 tile (ED's 115 -> 57 -> 28 -> 5 -> ... chain) gets a split per SM: its launch is bound by how many SMs
    // pull the two [B, <= 128] operands out of HBM, not by the partials (64 KB each; the fused optimizer walks them in 8-row items)
    li.split_cap = (h->bf16 && tiles == 1 && li.Kp <= 128 && li.Np <= 128 && getenv("CSB_NO_WIDE_SPLITS") == nullptr) ? sm : 64;
    if (li.split_cap > 64) li.max_w_splits = sm;
    if (h->tf32) {                    // split contraction of the TF32 weight gradient: one wave of 256 x 256 pair tiles (or 128 x 128 tiles)
      const bool pr = tf32_pairs(li.Np);
      const int t32 = pr ? (int)(ceil_div(ceil_div(li.Kp, 128), 2) * ceil_div(li.Np, 256)) : (int)(ceil_div(li.Kp, 128) * ceil_div(li.Np, 128));
      li.max_w_splits = std::max(1, std::min(64, (pr ? sm / 2 : sm) / t32));
    }
    // a last n-block of exactly half width (640 = 256 + 256 + 128) takes fewer, longer splits (NtParams.splits_narrow): nt_narrow_splits(S)
    // of them; the largest S with  m_tiles * ((n_blocks - 1) * S + nt_narrow_splits(S))  CTA groups in one wave
    li.nt_narrow = h->bf16 && g_use_nt_narrow && li.Np > li.nt_block_n && li.Np % li.nt_block_n == li.nt_block_n / 2;
    if (li.nt_narrow) {
      const int mt = nt_m_tiles(li.Kp, li.nt_cg), nb = (int)ceil_div(li.Np, li.nt_block_n), groups = sm / li.nt_cg;
      int S = li.max_w_splits;
      while (S < 64 && mt * ((nb - 1) * (S + 1) + nt_narrow_splits(S + 1)) <= groups) ++S;
      li.max_w_splits = S;
    }
    if (h->bf16 && l == h->L - 1 && li.Kp == 128 && li.Np == 128) li.max_w_splits = std::max(li.max_w_splits, sm);   // tail_kernel: one partial per CTA
    li.b_splits = h->bf16 ? li.max_w_splits * nt_m_tiles(li.Kp, li.nt_cg) : (h->tf32 ? 128 : 32);
    li.ws_w_off = ws_off; ws_off += (size_t)li.max_w_splits * li.Kp * li.Np;
    li.ws_b_off = ws_off; ws_off += (size_t)li.b_splits * li.Np;
    li.ws_g_off = ws_off; if (li.ln) ws_off += (size_t)std::max(32, 2 * sm) * 2 * li.Np;      // one partial pair per block of ln_bwd_bf16_kernel
    h->max_np = std::max(h->max_np, li.Np);
    k = li.N;
  }
  h->out_dim = k;
  h->out_p = h->layer[h->L - 1].Np;
  h->P_pad = off; h->P_user = off_user; h->ws_elems = ws_off;
  if (h->bf16) {
    if (h->out_dim % 4 != 0) { delete h; set_last_error("CSB_BF16 mode needs out_dim %% 4 == 0"); return CSB_EUNSUPPORTED; }
  }

#define CK(x) do { int _rc = (x); if (_rc) { free_all(h); delete h; return _rc; } } while (0)
#define CKA(ptr, bytes) do { int _rc = [&]() -> int { CSB_ALLOC(ptr, bytes); return CSB_OK; }(); if (_rc) { free_all(h); delete h; return _rc; } } while (0)
  CKA(h->params, h->P_pad * 4); CKA(h->grads, h->P_pad * 4); CKA(h->m, h->P_pad * 4); CKA(h->v, h->P_pad * 4);
  CKA(h->ws, h->ws_elems * 4);
  const size_t es = esize(h);
  CKA(h->xn, (size_t)h->cap * h->in_p * es);
  for (int l = 0; l + 1 < h->L; ++l) {
    CKA(h->act[l], (size_t)h->cap * h->layer[l].Np * es);
    if ((h->bf16 || (h->tf32 && !h->x3)) && !h->layer[l].ln && (h->layer[l].act == CSB_ACT_RELU || h->layer[l].act == CSB_ACT_LEAKYRELU) &&
        getenv("CSB_NO_MASK") == nullptr)
      CKA(h->amask[l], (size_t)h->cap * (h->layer[l].Np / 32) * 4);
    if (h->layer[l].ln) {
      CKA(h->zbuf[l], (size_t)h->cap * h->layer[l].Np * es);
      CKA(h->ln_stats[l], (size_t)h->cap * 2 * 4);
    }
  }
  CKA(h->dz[0], (size_t)h->cap * h->max_np * es);
  CKA(h->dz[1], (size_t)h->cap * h->max_np * es);
  CKA(h->pred, (size_t)h->cap * h->out_p * 4);
  CKA(h->d_sub, (size_t)h->in_p * 4); CKA(h->d_div, (size_t)h->in_p * 4);
  CKA(h->d_out_scale, (size_t)h->out_p * 4); CKA(h->d_inv_out_scale, (size_t)h->out_p * 4); CKA(h->d_loss_w, (size_t)h->out_p * 4);
  CKA(h->d_out_mask, (size_t)h->out_p * 4);
  h->n_loss_partials = (int)std::max<int64_t>(h->cap / 128 * ceil_div(h->out_p, 128) * tc::TN_EPI_WARPS, 8 * sm);   // 128-wide n-blocks: the most partials any tile policy writes
  CKA(h->loss_partials, (size_t)h->n_loss_partials * 4);
  CKA(h->d_loss, 4);
  if (h->bf16) {
    for (int l = 0; l < h->L; ++l) {
      CKA(h->w16[l], (size_t)h->layer[l].Kp * h->layer[l].Np * 2);
      CKA(h->wt16[l], (size_t)h->layer[l].Kp * h->layer[l].Np * 2);
    }
    CK(build_weight_maps(h));
  }
  if (h->tf32) {
    int max_dim = h->in_p;
    for (int l = 0; l < h->L; ++l) {
      const LayerInfo& li = h->layer[l];
      if (l + 1 < h->L && li.act == CSB_ACT_ELU) { free_all(h); delete h; set_last_error("CSB_TF32: ELU hidden layers are not instantiated"); return CSB_EUNSUPPORTED; }
      max_dim = std::max(max_dim, std::max(li.Kp, li.Np));
      const size_t kx = h->x3 ? 3 : 1;
      CKA(h->wt32[l], kx * li.Kp * li.Np * 4);
      CKA(h->w32r[l], kx * li.Kp * li.Np * 4);
      CK(make_tmap_f32(&h->tm32_wt[l], h->wt32[l], kx * li.Kp, li.Np, kx * li.Kp, tf32_b_box(li.Np)));
      CK(make_tmap_f32(&h->tm32_w[l], h->w32r[l], kx * li.Np, li.Kp, kx * li.Np, tf32_b_box(li.Kp)));
    }
    h->tr_ld = h->cap;                  // a multiple of 128 floats
    if (h->x3) {
      CKA(h->tr_a, (size_t)3 * max_dim * h->tr_ld * 4);
      CKA(h->tr_b, (size_t)3 * max_dim * h->tr_ld * 4);
      CKA(h->sp_a, (size_t)3 * max_dim * h->cap * 4);
    } else {
      CKA(h->xn_t, (size_t)h->in_p * h->tr_ld * 4);
      for (int l = 0; l + 1 < h->L; ++l) CKA(h->act_t[l], (size_t)h->layer[l].Np * h->tr_ld * 4);
      CKA(h->dz_t[0], (size_t)h->max_np * h->tr_ld * 4);
      CKA(h->dz_t[1], (size_t)h->max_np * h->tr_ld * 4);
    }
  }
  CK(csb_mlp_set_norm(h, nullptr, nullptr, nullptr, nullptr));
#undef CK
#undef CKA
  *out = h;
  return CSB_OK;
}

int csb_mlp_destroy(csb_mlp* h) {
  if (!h) return CSB_OK;
  cudaDeviceSynchronize();
  free_all(h);
  for (auto& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  delete h;
  return CSB_OK;
}

size_t csb_mlp_param_count(const csb_mlp* h) { return h ? h->P_user : 0; }

int csb_mlp_profile(csb_mlp* h, int enable) {
  CSB_REQUIRE(h, CSB_EINVAL, "null handle");
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  h->prof_on = enable != 0;
  h->prof_used = 0;
  for (int k = 0; k < K_COUNT; ++k) { h->prof_ms[k] = 0.0; h->prof_n[k] = 0; }
  return CSB_OK;
}
int csb_mlp_profile_read(csb_mlp* h, double* ms_by_kind, int64_t* launches_by_kind, int n_kinds) {
  CSB_REQUIRE(h && ms_by_kind && launches_by_kind && n_kinds >= K_COUNT, CSB_EINVAL, "need room for %d kinds", (int)K_COUNT);
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  for (int i = 1; i < h->prof_used; ++i) {
    const int kind = h->prof_kind[i];
    if (kind == K_BEGIN) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->prof_ev[i - 1], h->prof_ev[i]) != cudaSuccess) { cudaGetLastError(); continue; }
    h->prof_ms[kind] += ms;
    h->prof_n[kind] += 1;
  }
  h->prof_used = 0;
  for (int k = 0; k < K_COUNT; ++k) { ms_by_kind[k] = h->prof_ms[k]; launches_by_kind[k] = h->prof_n[k]; }
  return CSB_OK;
}
int csb_profile_kind_count(void) { return K_COUNT; }
const char* csb_profile_kind_name(int kind) { return (kind >= 0 && kind < K_COUNT) ? kProfNames[kind] : "?"; }
int64_t csb_mlp_launch_count(const csb_mlp* h) { return h ? h->launches : 0; }

// user (unpadded) blob <-> padded device buffer
static int upload_padded(csb_mlp* h, const float* user, float* dev) {
  std::vector<float> pad(h->P_pad, 0.f);
  for (int l = 0; l < h->L; ++l) {
    const LayerInfo& li = h->layer[l];
    for (int r = 0; r < li.K; ++r) memcpy(&pad[li.w_off + (size_t)r * li.Np], user + li.w_off_user + (size_t)r * li.N, (size_t)li.N * 4);
    memcpy(&pad[li.b_off], user + li.b_off_user, (size_t)li.N * 4);
    if (li.ln) {
      memcpy(&pad[li.g_off], user + li.g_off_user, (size_t)li.N * 4);
      memcpy(&pad[li.g_off + li.Np], user + li.g_off_user + li.N, (size_t)li.N * 4);
    }
  }
  CSB_CUDA_CHECK(cudaMemcpy(dev, pad.data(), h->P_pad * 4, cudaMemcpyHostToDevice));
  return CSB_OK;
}
static int download_padded(csb_mlp* h, const float* dev, float* user) {
  std::vector<float> pad(h->P_pad);
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  CSB_CUDA_CHECK(cudaMemcpy(pad.data(), dev, h->P_pad * 4, cudaMemcpyDeviceToHost));
  for (int l = 0; l < h->L; ++l) {
    const LayerInfo& li = h->layer[l];
    for (int r = 0; r < li.K; ++r) memcpy(user + li.w_off_user + (size_t)r * li.N, &pad[li.w_off + (size_t)r * li.Np], (size_t)li.N * 4);
    memcpy(user + li.b_off_user, &pad[li.b_off], (size_t)li.N * 4);
    if (li.ln) {
      memcpy(user + li.g_off_user, &pad[li.g_off], (size_t)li.N * 4);
      memcpy(user + li.g_off_user + li.N, &pad[li.g_off + li.Np], (size_t)li.N * 4);
    }
  }
  return CSB_OK;
}

int csb_mlp_set_params(csb_mlp* h, const float* params_host) {
  CSB_REQUIRE(h && params_host, CSB_EINVAL, "null argument");
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  int rc = upload_padded(h, params_host, h->params);
  if (rc) return rc;
  rc = repack_weights(h, 0);
  if (rc) return rc;
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  return CSB_OK;
}
int csb_mlp_get_params(csb_mlp* h, float* params_host) {
  CSB_REQUIRE(h && params_host, CSB_EINVAL, "null argument");
  return download_padded(h, h->params, params_host);
}
int csb_mlp_get_grads(csb_mlp* h, float* grads_host) {
  CSB_REQUIRE(h && grads_host, CSB_EINVAL, "null argument");
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  int rc = flush_pending(h, 0);
  if (rc) return rc;
  return download_padded(h, h->grads, grads_host);
}
static int pad_copy(csb_mlp* h, float* padded, float* user, int dir, cudaStream_t st) {
  simt::PadTable tab;
  tab.n = h->L;
  int64_t mx = 1;
  for (int l = 0; l < h->L; ++l) {
    const LayerInfo& li = h->layer[l];
    tab.l[l] = {li.K, li.N, li.Np, li.ln, li.w_off, li.b_off, li.g_off, li.w_off_user, li.b_off_user, li.g_off_user};
    mx = std::max<int64_t>(mx, (int64_t)(li.K + 3) * li.N);
  }
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(mx, 256), 4 * h->sm_count), (unsigned)h->L);
  simt::pad_copy_kernel<<<grid, 256, 0, st>>>(padded, user, dir, tab);
  CSB_CUDA_CHECK(cudaGetLastError());
  prof_mark(h, K_MISC, st);
  return CSB_OK;
}
int csb_mlp_set_params_device(csb_mlp* h, const float* params_dev, void* stream) {
  CSB_REQUIRE(h && params_dev, CSB_EINVAL, "null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = pad_copy(h, h->params, const_cast<float*>(params_dev), 0, st);
  if (rc) return rc;
  return repack_weights(h, st);
}
int csb_mlp_get_params_device(csb_mlp* h, float* params_dev, void* stream) {
  CSB_REQUIRE(h && params_dev, CSB_EINVAL, "null argument");
  return pad_copy(h, h->params, params_dev, 1, reinterpret_cast<cudaStream_t>(stream));
}
int csb_mlp_get_grads_device(csb_mlp* h, float* grads_dev, void* stream) {
  CSB_REQUIRE(h && grads_dev, CSB_EINVAL, "null argument");
  int rc = flush_pending(h, reinterpret_cast<cudaStream_t>(stream));
  if (rc) return rc;
  return pad_copy(h, h->grads, grads_dev, 1, reinterpret_cast<cudaStream_t>(stream));
}

int csb_mlp_get_opt_state(csb_mlp* h, float* m_host, float* v_host, int64_t* step) {
  CSB_REQUIRE(h, CSB_EINVAL, "null handle");
  int rc = CSB_OK;
  if (m_host) rc = download_padded(h, h->m, m_host);
  if (!rc && v_host) rc = download_padded(h, h->v, v_host);
  if (step) *step = h->step;
  return rc;
}
int csb_mlp_set_opt_state(csb_mlp* h, const float* m_host, const float* v_host, int64_t step) {
  CSB_REQUIRE(h, CSB_EINVAL, "null handle");
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  int rc = CSB_OK;
  if (m_host) rc = upload_padded(h, m_host, h->m);
  if (!rc && v_host) rc = upload_padded(h, v_host, h->v);
  h->step = step;
  return rc;
}

int csb_mlp_set_norm(csb_mlp* h, const float* inp_sub, const float* inp_div, const float* out_scale, const float* loss_w) {
  CSB_REQUIRE(h, CSB_EINVAL, "null handle");
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  auto up = [&](float* dev, const float* src, int n, int np, float fill, float padfill) -> int {
    std::vector<float> t(np, padfill);
    for (int i = 0; i < n; ++i) t[i] = src ? src[i] : fill;
    CSB_CUDA_CHECK(cudaMemcpy(dev, t.data(), (size_t)np * 4, cudaMemcpyHostToDevice));
    return CSB_OK;
  };
  int rc;
  // NULL keeps the previous value, except on the very first call from create (all NULL -> defaults)
  const bool init = !inp_sub && !inp_div && !out_scale && !loss_w;
  if (inp_sub || init) { rc = up(h->d_sub, inp_sub, h->in_dim, h->in_p, 0.f, 0.f); if (rc) return rc; }
  if (inp_div || init) { rc = up(h->d_div, inp_div, h->in_dim, h->in_p, 1.f, 1.f); if (rc) return rc; }
  if (out_scale || init) {
    rc = up(h->d_out_scale, out_scale, h->out_dim, h->out_p, 1.f, 1.f); if (rc) return rc;
    std::vector<float> inv(h->out_p, 1.f);
    for (int i = 0; i < h->out_dim; ++i) inv[i] = out_scale ? 1.f / out_scale[i] : 1.f;
    CSB_CUDA_CHECK(cudaMemcpy(h->d_inv_out_scale, inv.data(), (size_t)h->out_p * 4, cudaMemcpyHostToDevice));
  }
  if (loss_w || init) { rc = up(h->d_loss_w, loss_w, h->out_dim, h->out_p, 1.f, 0.f); if (rc) return rc; }
  return CSB_OK;
}

int csb_mlp_set_input_transform(csb_mlp* h, const float* exp_lambda, const float* keep, const float* clip_lo, const float* clip_hi) {
  CSB_REQUIRE(h, CSB_EINVAL, "null handle");
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  for (auto& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);    // cached graphs captured the old prologue
  h->graphs.clear();
  if (!exp_lambda && !keep && !clip_lo && !clip_hi) {
    if (h->d_xform) { cudaFree(h->d_xform); h->d_xform = nullptr; }
    return CSB_OK;
  }
  const int np = h->in_p;
  std::vector<float> t((size_t)4 * np);
  for (int c = 0; c < np; ++c) {
    const bool v = c < h->in_dim;
    t[c] = (v && exp_lambda) ? exp_lambda[c] : 0.f;
    t[np + c] = v ? ((keep == nullptr || keep[c] != 0.f) ? 1.f : 0.f) : 0.f;
    t[2 * np + c] = (v && clip_lo) ? clip_lo[c] : -INFINITY;
    t[3 * np + c] = (v && clip_hi) ? clip_hi[c] : INFINITY;
    CSB_REQUIRE(!(t[2 * np + c] > t[3 * np + c]), CSB_EINVAL, "clip_lo[%d] > clip_hi[%d]", c, c);
  }
  if (!h->d_xform) { CSB_ALLOC(h->d_xform, (size_t)4 * np * 4); }
  CSB_CUDA_CHECK(cudaMemcpy(h->d_xform, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
  return CSB_OK;
}

static inline uint32_t mlp_drop_seed(const csb_mlp* h, int layer) {
  return h->drop_seed ^ (uint32_t)((uint64_t)h->drop_fwd * 0x9E3779B97F4A7C15ull >> 32) ^ (uint32_t)(layer + 1) * 0x85EBCA77u;
}
static inline uint32_t mlp_drop_threshold(const csb_mlp* h) { return (uint32_t)lrintf(h->dropout * 16777216.f); }   // keep iff 24 random bits >= p * 2^24

int csb_mlp_set_dropout(csb_mlp* h, float rate, uint32_t seed) {
  CSB_REQUIRE(h, CSB_EINVAL, "null handle");
  CSB_REQUIRE(rate >= 0.f && rate < 1.f, CSB_EINVAL, "dropout rate %g outside [0, 1)", (double)rate);
  for (int l = 0; l + 1 < h->L && rate > 0.f; ++l)
    CSB_REQUIRE(h->layer[l].act == CSB_ACT_RELU || h->layer[l].act == CSB_ACT_LEAKYRELU || h->layer[l].act == CSB_ACT_NONE, CSB_EUNSUPPORTED,
                "dropout needs hidden activations whose derivative follows from the sign of the stored output (ReLU, LeakyReLU, linear)");
  h->dropout = rate; h->drop_seed = seed; h->drop_fwd = 0; h->drop_live = false;
  return CSB_OK;
}

int csb_mlp_debug_dropout_mask(csb_mlp* h, int layer, float* dst_dev, int64_t B, void* stream) {
  CSB_REQUIRE(h && dst_dev, CSB_EINVAL, "null argument");
  CSB_REQUIRE(layer >= 0 && layer + 1 < h->L && B >= 1 && B <= h->cfg.max_batch, CSB_EINVAL, "bad layer / batch");
  CSB_REQUIRE(h->dropout > 0.f && h->drop_fwd > 0, CSB_ESTATE, "no training forward with dropout has run on this handle");
  const LayerInfo& li = h->layer[layer];
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  simt::dropout_mask_kernel<<<grid_for(B * li.Np / 8, 256, h->sm_count), 256, 0, st>>>(dst_dev, B, li.N, li.Np, mlp_drop_seed(h, layer), mlp_drop_threshold(h),
                                                                                     1.f / (1.f - h->dropout));
  CSB_CUDA_CHECK(cudaGetLastError());
  return CSB_OK;
}

int csb_mlp_set_output_mask(csb_mlp* h, const float* mask_host) {
  CSB_REQUIRE(h, CSB_EINVAL, "null handle");
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  h->has_mask = mask_host != nullptr;
  if (mask_host) {
    std::vector<float> m(h->out_p, 0.f);
    for (int i = 0; i < h->out_dim; ++i) m[i] = mask_host[i] != 0.f ? 1.f : 0.f;
    CSB_CUDA_CHECK(cudaMemcpy(h->d_out_mask, m.data(), (size_t)h->out_p * 4, cudaMemcpyHostToDevice));
  }
  for (auto& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);    // cached graphs captured the old epilogue parameters
  h->graphs.clear();
  return CSB_OK;
}

int csb_mlp_grad_buffer(csb_mlp* h, float** ptr, size_t* n) {
  CSB_REQUIRE(h && ptr && n, CSB_EINVAL, "null argument");
  *ptr = h->grads;
  *n = h->P_pad;
  return CSB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// forward pieces
// ---------------------------------------------------------------------------------------------------------------
static int run_normalize_raw(csb_mlp* h, const float* x, int64_t B, int apply, cudaStream_t st) {
  if (apply && h->d_xform != nullptr) {       // generalised prologue (exp transforms, pruned columns, clipping) of the online models
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(B, 2), (int64_t)h->sm_count * 8), (unsigned)ceil_div(h->in_p, 128));
    simt::prepare_input_kernel<<<grid, 256, 0, st>>>(x, h->in_dim, h->d_sub, h->d_div, h->d_xform,
                                                     h->bf16 ? nullptr : reinterpret_cast<float*>(h->xn),
                                                     h->bf16 ? reinterpret_cast<__nv_bfloat16*>(h->xn) : nullptr, h->in_p, B, h->in_dim, h->in_p);
    CSB_CUDA_CHECK(cudaGetLastError());
    prof_mark(h, K_NORMALIZE, st);
    return CSB_OK;
  }
  if (h->bf16 && h->in_dim % 4 == 0 && 256 % (h->in_p / 4) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int grid = grid_for(B * (h->in_p / 4), 256, h->sm_count, 4);
    CSB_CUDA_CHECK(launch_pdl(simt::normalize_bf16_vec4_kernel, dim3(grid), dim3(256), 0, st, x, h->d_sub, h->d_div, apply,
                              reinterpret_cast<__nv_bfloat16*>(h->xn), B, h->in_dim, h->in_p));
  } else {
    const int grid = grid_for(B * h->in_p, 256, h->sm_count);
    simt::normalize_kernel<<<grid, 256, 0, st>>>(x, h->in_dim, h->d_sub, h->d_div, apply,
                                                 h->bf16 ? nullptr : reinterpret_cast<float*>(h->xn), h->in_p,
                                                 h->bf16 ? reinterpret_cast<__nv_bfloat16*>(h->xn) : nullptr, h->in_p, B,
                                                 h->in_dim, h->in_p);
  }
  CSB_CUDA_CHECK(cudaGetLastError());
  prof_mark(h, K_NORMALIZE, st);
  return CSB_OK;
}

static int run_normalize(csb_mlp* h, const float* x, int64_t B, int apply, cudaStream_t st) {
  int rc = run_normalize_raw(h, x, B, apply, st);
  if (rc || !h->tf32 || h->x3) return rc;
  // CSB_TF32: the first GEMM's A operand rounded (to nearest) onto the TF32 grid, like every other stored tensor of this mode
  simt::transpose_f32_kernel<<<transpose_grid(h, B, h->in_p), 256, 0, st>>>(reinterpret_cast<const float*>(h->xn), h->in_p, B, h->in_p,
                                                                          reinterpret_cast<float*>(h->xn), 0, 1);
  simt::transpose_f32_kernel<<<transpose_grid(h, B, h->in_p), 256, 0, st>>>(reinterpret_cast<const float*>(h->xn), h->in_p, B, h->in_p, h->xn_t, h->tr_ld, 1);
  CSB_CUDA_CHECK(cudaGetLastError());
  return CSB_OK;
}

static int run_hidden_forward(csb_mlp* h, int64_t B, cudaStream_t st, bool training = false) {
  const bool drop = training && h->dropout > 0.f;
  if (drop) h->drop_fwd++;
  h->drop_live = drop;
  for (int l = 0; l + 1 < h->L; ++l) {
    const LayerInfo& li = h->layer[l];
    // LayerNorm layers: the GEMM writes z = hW + b (no activation) to zbuf; ln_fwd_kernel then produces act(LN(z))
    const int gemm_act = li.ln ? CSB_ACT_NONE : li.act;
    if (h->bf16) {
      tc::GemmParams p = {};
      p.M = (int)B; p.N = li.Np; p.K = li.Kp; p.act = gemm_act; p.alpha = li.alpha; p.head_relu_from = -1;
      p.bias = h->params + li.b_off; p.out = li.ln ? h->zbuf[l] : h->act[l]; p.ld_out = li.Np;
      p.mask_out = h->amask[l]; p.ld_mask = (int)h->cap;
      int rc = launch_tn_auto<tc::EPI_BIAS_ACT>(h->tm_in[l].a_k128, h->tm_wt[l], p, h->sm_count, st, &h->tm_wt_s[l]);
      if (rc) return rc;
    } else if (h->tf32) {
      tc::GemmParams p = {};
      p.M = (int)B; p.N = li.Np; p.K = li.Kp; p.act = gemm_act; p.alpha = li.alpha; p.head_relu_from = -1;
      if (h->x3) {
        simt::split3_f32_kernel<<<transpose_grid(h, B, li.Kp), 256, 0, st>>>(reinterpret_cast<const float*>(layer_in(h, l)), li.Kp, B, li.Kp, h->sp_a, 3 * (int64_t)li.Kp, li.Kp, 0, 0, B);
        CSB_CUDA_CHECK(cudaGetLastError());
        p.K = 3 * li.Kp; p.tf32_exact_store = 1;
      }
      p.bias = h->params + li.b_off; p.out = li.ln ? h->zbuf[l] : h->act[l]; p.ld_out = li.Np;
      h->act_t_valid[l] = false;
      if (!h->x3) {
        p.mask_out = h->amask[l]; p.ld_mask = (int)h->cap;
        if (!li.ln && !drop) { p.out_t = h->act_t[l]; p.ld_out_t = h->tr_ld; h->act_t_valid[l] = true; }
      }
      int rc = launch_tn_tf32<tc::EPI_BIAS_ACT>(h->tm32_in[l], h->tm32_wt[l], p, h->sm_count, st);
      if (rc) return rc;
    } else {
      simt::SgemmParams p = {};
      p.M = (int)B; p.N = li.Np; p.K = li.Kp;
      p.A = reinterpret_cast<const float*>(layer_in(h, l)); p.lda = li.Kp;
      p.B = h->params + li.w_off; p.ldb = li.Np;
      p.C = reinterpret_cast<float*>(li.ln ? h->zbuf[l] : h->act[l]); p.ldc = li.Np;
      p.bias = h->params + li.b_off; p.act = gemm_act; p.alpha = li.alpha; p.head_relu_from = -1;
      dim3 grid((unsigned)(li.Np / 64), (unsigned)ceil_div(B, 64));
      simt::sgemm_kernel<false, false, simt::SEPI_BIAS_ACT><<<grid, 256, 0, st>>>(p);
      CSB_CUDA_CHECK(cudaGetLastError());
    }
    prof_mark(h, K_GEMM_FWD, st);
    if (li.ln) {
      const int grid = (int)std::min<int64_t>(ceil_div(B, 8), (int64_t)h->sm_count * 8);
      const float* gamma = h->params + li.g_off;
      if (h->bf16 && li.Np <= 256 * simt::LN_MAXP)
        simt::ln_fwd_bf16_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(h->zbuf[l]), reinterpret_cast<__nv_bfloat16*>(h->act[l]),
                                                       li.Np, gamma, gamma + li.Np, h->ln_stats[l], B, li.N, li.Np, li.act, li.alpha, 1e-5f);
      else if (h->bf16)
        simt::ln_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(h->zbuf[l]),
                                                                 reinterpret_cast<__nv_bfloat16*>(h->act[l]), li.Np, gamma, gamma + li.Np,
                                                                 h->ln_stats[l], B, li.N, li.Np, li.act, li.alpha, 1e-5f);
      else
        simt::ln_fwd_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(h->zbuf[l]), reinterpret_cast<float*>(h->act[l]),
                                                         li.Np, gamma, gamma + li.Np, h->ln_stats[l], B, li.N, li.Np, li.act, li.alpha, 1e-5f);
      CSB_CUDA_CHECK(cudaGetLastError());
      prof_mark(h, K_MISC, st);
    }
    if (drop) {
      // inverted dropout in place on the stored activation (padding columns are zero and stay zero)
      const int64_t n8 = B * li.Np / 8;
      const float scale = 1.f / (1.f - h->dropout);
      if (h->bf16) simt::dropout_bf16_kernel<<<grid_for(n8, 256, h->sm_count), 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(h->act[l]), n8, mlp_drop_seed(h, l), mlp_drop_threshold(h), scale);
      else simt::dropout_f32_kernel<<<grid_for(n8, 256, h->sm_count), 256, 0, st>>>(reinterpret_cast<float*>(h->act[l]), n8, mlp_drop_seed(h, l), mlp_drop_threshold(h), scale);
      CSB_CUDA_CHECK(cudaGetLastError());
      prof_mark(h, K_MISC, st);
    }
  }
  return CSB_OK;
}

// output layer.  mode 0: predictions only (pred buffer);  mode 1 (bf16 only): fused loss + dZ
static int run_head(csb_mlp* h, int64_t B, int fused_loss, const float* y, float grad_scale, cudaStream_t st) {
  const int l = h->L - 1;
  const LayerInfo& li = h->layer[l];
  if (h->bf16) {
    tc::GemmParams p = {};
    p.M = (int)B; p.N = li.Np; p.K = li.Kp; p.act = li.act; p.alpha = li.alpha; p.head_relu_from = h->cfg.head_relu_from;
    p.bias = h->params + li.b_off; p.out_dim = h->out_dim;
    p.pred = h->pred; p.ld_pred = h->out_p;
    p.out_mask = h->has_mask ? h->d_out_mask : nullptr;
    int rc;
    if (fused_loss) {
      p.out = dz16(h, l); p.ld_out = li.Np;
      p.y = y; p.ld_y = h->out_dim; p.loss_w = h->d_loss_w; p.grad_scale = grad_scale; p.loss_kind = h->cfg.loss;
      p.loss_partials = h->loss_partials;
      p.pred = nullptr;
      rc = launch_tn_auto<tc::EPI_HEAD_LOSS>(h->tm_in[l].a_k128, h->tm_wt[l], p, h->sm_count, st, &h->tm_wt_s[l]);
    } else {
      rc = launch_tn_auto<tc::EPI_HEAD_OUT>(h->tm_in[l].a_k128, h->tm_wt[l], p, h->sm_count, st, &h->tm_wt_s[l]);
    }
    if (rc) return rc;
  } else if (h->tf32) {
    // predictions (fp32) straight from the head epilogue; loss and dL/dz follow in head_grad_kernel as in the fp32 mode
    tc::GemmParams p = {};
    p.M = (int)B; p.N = li.Np; p.K = li.Kp; p.act = li.act; p.alpha = li.alpha; p.head_relu_from = h->cfg.head_relu_from;
    if (h->x3) {
      simt::split3_f32_kernel<<<transpose_grid(h, B, li.Kp), 256, 0, st>>>(reinterpret_cast<const float*>(layer_in(h, l)), li.Kp, B, li.Kp, h->sp_a, 3 * (int64_t)li.Kp, li.Kp, 0, 0, B);
      CSB_CUDA_CHECK(cudaGetLastError());
      p.K = 3 * li.Kp;
    }
    p.bias = h->params + li.b_off; p.out_dim = h->out_dim;
    p.pred = h->pred; p.ld_pred = h->out_p;
    p.out_mask = h->has_mask ? h->d_out_mask : nullptr;
    int rc = li.act == CSB_ACT_ELU ? launch_tn_tf32<tc::EPI_HEAD_OUT, tc::VAR_ELU>(h->tm32_in[l], h->tm32_wt[l], p, h->sm_count, st)
                                   : launch_tn_tf32<tc::EPI_HEAD_OUT>(h->tm32_in[l], h->tm32_wt[l], p, h->sm_count, st);
    if (rc) return rc;
  } else {
    simt::SgemmParams p = {};
    p.M = (int)B; p.N = li.Np; p.K = li.Kp;
    p.A = reinterpret_cast<const float*>(layer_in(h, l)); p.lda = li.Kp;
    p.B = h->params + li.w_off; p.ldb = li.Np;
    p.C = h->pred; p.ldc = h->out_p;
    p.bias = h->params + li.b_off; p.act = li.act; p.alpha = li.alpha; p.head_relu_from = h->cfg.head_relu_from;
    dim3 grid((unsigned)(li.Np / 64), (unsigned)ceil_div(B, 64));
    simt::sgemm_kernel<false, false, simt::SEPI_BIAS_ACT><<<grid, 256, 0, st>>>(p);
    CSB_CUDA_CHECK(cudaGetLastError());
    if (h->has_mask) {
      simt::colmask_kernel<<<grid_for(B * h->out_dim, 256, h->sm_count), 256, 0, st>>>(h->pred, h->out_p, h->d_out_mask, B, h->out_dim);
      CSB_CUDA_CHECK(cudaGetLastError());
    }
  }
  prof_mark(h, K_GEMM_HEAD, st);
  return CSB_OK;
}

// The fused tail (tail_kernel.cuh) covers an output layer of exactly 128 x 128 padded columns behind a ReLU / LeakyReLU layer whose
// sign mask exists, MSE without an output mask on a non-ELU head: MLP_v1.  Everything else keeps the three separate launches.
static inline bool tail_fusable(const csb_mlp* h) {
  if (!h->bf16 || !g_use_tail || h->L < 2 || h->dropout > 0.f) return false;      // (the tail's data gradient reads the sign mask, which knows nothing of dropped elements)
  const LayerInfo& li = h->layer[h->L - 1];
  return li.Kp == 128 && li.Np == 128 && !li.ln && h->amask[h->L - 2] != nullptr && h->cfg.loss == CSB_LOSS_MSE && !h->has_mask &&
         li.act != CSB_ACT_ELU && h->out_dim % 4 == 0;
}
static inline int tail_grid(const csb_mlp* h, int64_t B) { return (int)std::min<int64_t>(ceil_div(B, 128), h->sm_count); }

// output layer + loss + its data gradient + its weight / bias gradient partials in one launch (tail_kernel.cuh)
static int run_tail(csb_mlp* h, int64_t B, const float* y, float grad_scale, cudaStream_t st) {
  const int l = h->L - 1;
  const LayerInfo &li = h->layer[l], &lp = h->layer[l - 1];
  tc::TailParams p = {};
  p.M = (int)B; p.act = li.act; p.alpha = li.alpha; p.head_relu_from = h->cfg.head_relu_from;
  p.bias = h->params + li.b_off; p.loss_w = h->d_loss_w; p.y = y; p.ld_y = h->out_dim; p.out_dim = h->out_dim; p.grad_scale = grad_scale;
  p.loss_partials = h->loss_partials;
  p.prev_act = lp.act; p.prev_alpha = lp.alpha; p.mask_in = h->amask[l - 1]; p.ld_mask = (int)h->cap;
  p.dz_prev = dz16(h, l - 1); p.ld_dz = lp.Np;
  p.dw_out = h->ws + li.ws_w_off; p.db_out = h->ws + li.ws_b_off;
  static bool attr_set = false;
  if (!attr_set) {
    CSB_CUDA_CHECK(cudaFuncSetAttribute(tc::tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TailSmem::TOTAL));
    attr_set = true;
  }
  CSB_CUDA_CHECK(launch_pdl(tc::tail_kernel, dim3((unsigned)tail_grid(h, B)), dim3(tc::TN_THREADS), (size_t)tc::TailSmem::TOTAL, st,
                            h->tm_in[l].a_k128, h->tm_wt_s[l], h->tm_w_s[l], p));
  prof_mark(h, K_GEMM_HEAD, st);
  return CSB_OK;
}

static int check_batch(csb_mlp* h, int64_t B) {
  CSB_REQUIRE(B >= 0 && B <= h->cfg.max_batch, CSB_ESTATE, "batch %lld exceeds max_batch %lld", (long long)B, (long long)h->cfg.max_batch);
  return CSB_OK;
}

int csb_mlp_forward(csb_mlp* h, const float* x, float* y_pred, int64_t B, uint32_t flags, void* stream) {
  CSB_REQUIRE(h && x && y_pred, CSB_EINVAL, "null argument");
  int rc = check_batch(h, B);
  if (rc) return rc;
  if (B == 0) return CSB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((rc = build_act_maps(h, B))) return rc;
  prof_mark(h, K_BEGIN, st);
  if ((rc = run_normalize(h, x, B, (flags & CSB_FWD_NORMALIZE_IN) ? 1 : 0, st))) return rc;
  if ((rc = run_hidden_forward(h, B, st, (flags & CSB_FWD_TRAINING) != 0))) return rc;
  if ((rc = run_head(h, B, 0, nullptr, 0.f, st))) return rc;
  const int grid = grid_for(B * h->out_dim, 256, h->sm_count);
  simt::scale_copy_kernel<<<grid, 256, 0, st>>>(h->pred, h->out_p, (flags & CSB_FWD_DENORM_OUT) ? h->d_inv_out_scale : nullptr,
                                                y_pred, h->out_dim, B, h->out_dim);
  CSB_CUDA_CHECK(cudaGetLastError());
  prof_mark(h, K_MISC, st);
  h->acts_B = (flags & CSB_FWD_KEEP_ACTIVATIONS) ? B : -1;
  h->acts_normalized = (flags & CSB_FWD_NORMALIZE_IN) != 0;
  return CSB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// backward chain: given dZ_{L-1} in dz(L-1), produce all parameter gradients (and optionally dx)
// ---------------------------------------------------------------------------------------------------------------
// split-K geometry of the weight-gradient GEMM of layer l at batch B (CSB_BF16)
static inline int wgrad_splits(const csb_mlp* h, int l, int64_t B, int* rb_per_split, int* rb_per_split_narrow = nullptr, bool tail = false) {
  const LayerInfo& li = h->layer[l];
  if (tail && l == h->L - 1) return tail_grid(h, B);               // one partial per CTA of tail_kernel
  const int num_rb = (int)ceil_div(B, 64);
  int splits = std::max(1, std::min(std::min(li.max_w_splits, li.split_cap), num_rb));      // (slots beyond the cap exist for tail_kernel only)
  const int rps = (int)ceil_div(num_rb, splits);
  if (rb_per_split) *rb_per_split = rps;
  splits = (int)ceil_div(num_rb, rps);
  if (rb_per_split_narrow) *rb_per_split_narrow = (li.nt_narrow && splits >= 2) ? (int)ceil_div(num_rb, nt_narrow_splits(splits)) : 0;
  return splits;
}
static inline void set_wgrad_geometry(const csb_mlp* h, int l, int64_t B, tc::NtParams& p) {
  p.splits = wgrad_splits(h, l, B, &p.rb_per_split, &p.rb_per_split_narrow);
  p.splits_narrow = p.rb_per_split_narrow > 0 ? nt_narrow_splits(p.splits) : 0;
}
static inline bool fused_opt_supported(const csb_mlp* h) {
  static const bool off = getenv("CSB_NO_FUSED_OPT") != nullptr;     // debugging aid: always take the three-launch path
  return h->bf16 && !off;          // LayerNorm layers included: their (gamma, beta) partials ride in the same launch
}
// number of (dgamma, dbeta) partial pairs the LayerNorm backward of layer l writes at batch B (one per block of the kernel that runs)
static inline int ln_grad_splits(const csb_mlp* h, int l, int64_t B) {
  const LayerInfo& li = h->layer[l];
  if (h->bf16 && li.Np <= 256 * simt::LN_MAXP) return (int)std::max<int64_t>(1, std::min<int64_t>(2 * h->sm_count, ceil_div(B, 8)));
  return (int)std::max<int64_t>(1, std::min<int64_t>(32, ceil_div(B, 256)));
}

// reduce the pending split partials into the gradient buffer (and finish the pending loss) with the plain reduction kernel
static int flush_pending(csb_mlp* h, cudaStream_t st) {
  if (!h->pending) return CSB_OK;
  simt::SegmentTable tab;
  tab.n = 0;
  tab.loss_partials = h->loss_partials; tab.n_loss = h->pending_n_loss; tab.loss_out = h->pending_loss_out;
  int64_t max_len = 4;
  for (int l = h->L - 1; l >= 0; --l) {
    const LayerInfo& li = h->layer[l];
    const int splits = wgrad_splits(h, l, h->pending_B, nullptr, nullptr, h->pending_tail);
    tab.seg[tab.n++] = {h->ws + li.ws_w_off, (size_t)li.Kp * li.Np, h->grads + li.w_off, (int64_t)li.Kp * li.Np, splits};
    tab.seg[tab.n++] = {h->ws + li.ws_b_off, (size_t)li.Np, h->grads + li.b_off, (int64_t)li.Np, splits * nt_m_tiles(li.Kp, li.nt_cg)};
    if (li.ln) {
      const int S = ln_grad_splits(h, l, h->pending_B);
      tab.seg[tab.n++] = {h->ws + li.ws_g_off, (size_t)2 * li.Np, h->grads + li.g_off, (int64_t)2 * li.Np, S};
    }
    max_len = std::max<int64_t>(max_len, (int64_t)li.Kp * li.Np);
  }
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(max_len / 4, 256), 2 * h->sm_count), (unsigned)(tab.n + (tab.loss_out ? 1 : 0)));
  CSB_CUDA_CHECK(launch_pdl(simt::reduce_partials_kernel, grid, dim3(256), 0, st, tab));
  prof_mark(h, K_REDUCE, st);
  h->pending = false;
  return CSB_OK;
}

// loss partials the output-layer kernel writes for a batch of B rows: one per (m-block, n-block, epilogue warp) of the tile shape
// the launch policy picks (launch_tn_shape)
static inline int head_loss_partials(const csb_mlp* h, int64_t B) {
  const int l = h->L - 1;
  const int bn = tn_small_tiles((int)B, h->out_p, h->sm_count, &h->tm_wt_s[l]) ? 128 : tn_block_n(h->out_p);
  return (int)(ceil_div(B, 128) * ceil_div(h->out_p, bn)) * tc::TN_EPI_WARPS;
}

static int run_backward_chain(csb_mlp* h, int64_t B, float* dx, cudaStream_t st, int n_loss_partials = 0, float* loss_out = nullptr,
                              bool defer_reduce = false, bool tail_done = false) {
  simt::SegmentTable tab;
  tab.n = 0;
  tab.loss_partials = h->loss_partials; tab.n_loss = n_loss_partials; tab.loss_out = loss_out;
  int64_t max_len = 4;
  const bool conc = h->bf16 && h->side_on && !h->prof_on;      // per-kind profiling wants the launches back to back on one stream
  h->pending_tail = false;
  if (h->tf32 && !h->x3) {
    // the output layer's dL/dz (written in full fp32 by the loss kernels) onto the TF32 grid, like every dZ this mode stores: all of its
    // consumers (bias column sums, weight gradient, data gradient) then see one and the same value
    const LayerInfo& lo = h->layer[h->L - 1];
    simt::transpose_f32_kernel<<<transpose_grid(h, B, lo.Np), 256, 0, st>>>(dz32(h, h->L - 1), lo.Np, B, lo.Np, dz32(h, h->L - 1), 0, 1);
    CSB_CUDA_CHECK(cudaGetLastError());
    h->dz_t_valid = false;
  }
  for (int l = h->L - 1; l >= 0; --l) {
    const LayerInfo& li = h->layer[l];
    if (tail_done && l == h->L - 1) {
      // tail_kernel already produced dZ of the layer below and this layer's gradient partials (one per CTA)
      const int splits = tail_grid(h, B);
      tab.seg[tab.n++] = {h->ws + li.ws_w_off, (size_t)li.Kp * li.Np, h->grads + li.w_off, (int64_t)li.Kp * li.Np, splits};
      tab.seg[tab.n++] = {h->ws + li.ws_b_off, (size_t)li.Np, h->grads + li.b_off, (int64_t)li.Np, splits};
      max_len = std::max<int64_t>(max_len, (int64_t)li.Kp * li.Np);
      h->pending_tail = true;
      continue;
    }
    if (li.ln) {
      // the buffer holds du_l = dA_l * act'(a_l): LayerNorm parameter gradients, then du -> dz in place
      const int S = ln_grad_splits(h, l, B);
      dim3 gridp((unsigned)(li.Np / 64), (unsigned)S);
      const int gridb = (int)std::min<int64_t>(ceil_div(B, 8), (int64_t)h->sm_count * 8);
      const float* gamma = h->params + li.g_off;
      if (h->bf16 && li.Np <= 256 * simt::LN_MAXP) {
        // one fused kernel: du -> dz and the (dgamma, dbeta) partials of each block
        // S = two resident blocks per SM
        const size_t smem = (size_t)(1 + 2 * 8) * 256 * simt::LN_MAXP * 4;        // gamma + (dgamma, dbeta) per warp
        static bool attr_set = false;
        if (!attr_set) {
          CSB_CUDA_CHECK(cudaFuncSetAttribute(simt::ln_bwd_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          attr_set = true;
        }
        simt::ln_bwd_bf16_kernel<<<S, 256, smem, st>>>(dz16(h, l), reinterpret_cast<const __nv_bfloat16*>(h->zbuf[l]), li.Np, gamma, h->ln_stats[l],
                                                       B, li.N, li.Np, h->ws + li.ws_g_off);
      } else if (h->bf16) {
        simt::ln_param_grad_kernel<__nv_bfloat16><<<gridp, 256, 0, st>>>(dz16(h, l), reinterpret_cast<const __nv_bfloat16*>(h->zbuf[l]), li.Np,
                                                                        h->ln_stats[l], B, h->ws + li.ws_g_off, (size_t)2 * li.Np, li.Np);
        simt::ln_bwd_kernel<__nv_bfloat16><<<gridb, 256, 0, st>>>(dz16(h, l), reinterpret_cast<const __nv_bfloat16*>(h->zbuf[l]), li.Np, gamma,
                                                                 h->ln_stats[l], B, li.N);
      } else {
        simt::ln_param_grad_kernel<float><<<gridp, 256, 0, st>>>(dz32(h, l), reinterpret_cast<const float*>(h->zbuf[l]), li.Np, h->ln_stats[l], B,
                                                                h->ws + li.ws_g_off, (size_t)2 * li.Np, li.Np);
        simt::ln_bwd_kernel<float><<<gridb, 256, 0, st>>>(dz32(h, l), reinterpret_cast<const float*>(h->zbuf[l]), li.Np, gamma, h->ln_stats[l], B, li.N);
      }
      CSB_CUDA_CHECK(cudaGetLastError());
      prof_mark(h, K_MISC, st);
      prof_mark(h, K_MISC, st);
      tab.seg[tab.n++] = {h->ws + li.ws_g_off, (size_t)2 * li.Np, h->grads + li.g_off, (int64_t)li.Np, S};
      tab.seg[tab.n++] = {h->ws + li.ws_g_off + li.Np, (size_t)2 * li.Np, h->grads + li.g_off + li.Np, (int64_t)li.Np, S};
    }
    // ---- weight gradient dW_l = in_l^T . dZ_l  (contraction over the B rows) and bias gradient, on stream `ws`
    auto weight_grads = [&](cudaStream_t ws) -> int {
      int splits = 1;
      if (h->bf16) {
        tc::NtParams p = {};
        p.M = li.Kp; p.N = li.Np; p.R = (int)B;
        set_wgrad_geometry(h, l, B, p);
        splits = p.splits;
        p.out = h->ws + li.ws_w_off; p.ld_out = li.Np; p.split_stride = (size_t)li.Kp * li.Np;
        p.colsum_out = h->ws + li.ws_b_off; p.colsum_stride = (size_t)li.Np;      // bias gradient fused into this kernel
        int rc = launch_nt_auto(h->tm_in[l].mn64, h->tm_dz[l].mn64, p, splits, ws, li.nt_cg);
        if (rc) return rc;
      } else if (h->tf32) {
        // dW_l = in_l^T . dZ_l contracts over the batch: both operands transposed to [features, batch] (K-major for tcgen05), the
        // contraction cut into `splits` ranges whose fp32 partial products are summed in a fixed order with every other gradient
        const int64_t bb = round_up(B, 32);
        if (h->x3) {
          simt::split3_f32_kernel<<<transpose_grid(h, bb, li.Kp), 256, 0, ws>>>(reinterpret_cast<const float*>(layer_in(h, l)), li.Kp, B, li.Kp, h->tr_a, 3 * bb, bb, 1, 0, bb);
          simt::split3_f32_kernel<<<transpose_grid(h, bb, li.Np), 256, 0, ws>>>(dz32(h, l), li.Np, B, li.Np, h->tr_b, 3 * bb, bb, 1, 1, bb);
        } else {
          // the [features, batch] copies the epilogues did not write themselves
          if (l > 0 && !h->act_t_valid[l - 1])
            simt::transpose_f32_kernel<<<transpose_grid(h, B, li.Kp), 256, 0, ws>>>(reinterpret_cast<const float*>(layer_in(h, l)), li.Kp, B, li.Kp, h->act_t[l - 1], h->tr_ld, 1);
          if (!h->dz_t_valid)
            simt::transpose_f32_kernel<<<transpose_grid(h, B, li.Np), 256, 0, ws>>>(dz32(h, l), li.Np, B, li.Np, h->dz_t[(h->L - 1 - l) & 1], h->tr_ld, 1);
        }
        CSB_CUDA_CHECK(cudaGetLastError());
        tc::GemmParams p = {};
        p.M = li.Kp; p.N = li.Np; p.K = (int)((h->x3 ? 3 : 1) * bb);          // columns past B are zero (TMA fill / split3's padding)
        const int num_kb = p.K / 32;
        const int want = std::max(1, std::min(li.max_w_splits, num_kb));
        const int kps = (int)ceil_div(num_kb, want);
        splits = (int)ceil_div(num_kb, kps);                                  // every range holds at least one k-block
        p.k_splits = splits; p.split_stride = (size_t)li.Kp * li.Np;
        p.out = h->ws + li.ws_w_off; p.ld_out = li.Np;
        int rc = launch_tn_tf32<tc::EPI_F32, tc::VAR_KSPLIT>(h->tm32_tra[l], h->tm32_trb[l], p, h->sm_count, ws);
        if (rc) return rc;
      } else {
        simt::SgemmParams p = {};
        p.M = li.Kp; p.N = li.Np; p.K = (int)B;
        p.A = reinterpret_cast<const float*>(layer_in(h, l)); p.lda = li.Kp;
        p.B = dz32(h, l); p.ldb = li.Np;
        p.C = h->ws + li.ws_w_off; p.ldc = li.Np;
        dim3 grid((unsigned)(li.Np / 64), (unsigned)(li.Kp / 64));
        simt::sgemm_kernel<true, false, simt::SEPI_STORE><<<grid, 256, 0, ws>>>(p);
        CSB_CUDA_CHECK(cudaGetLastError());
      }
      prof_mark(h, K_GEMM_WGRAD, ws);
      tab.seg[tab.n++] = {h->ws + li.ws_w_off, (size_t)li.Kp * li.Np, h->grads + li.w_off, (int64_t)li.Kp * li.Np, splits};
      max_len = std::max<int64_t>(max_len, (int64_t)li.Kp * li.Np);
      // ---- bias gradient db_l = column sums of dZ_l (CSB_BF16: computed inside the weight-gradient kernel above)
      if (h->bf16) {
        tab.seg[tab.n++] = {h->ws + li.ws_b_off, (size_t)li.Np, h->grads + li.b_off, (int64_t)li.Np, splits * nt_m_tiles(li.Kp, li.nt_cg)};
      } else {
        const int S = (int)std::max<int64_t>(1, std::min<int64_t>(li.b_splits, ceil_div(B, 256)));
        dim3 grid((unsigned)(li.Np / 64), (unsigned)S);
        simt::colsum_kernel<float><<<grid, 256, 0, ws>>>(dz32(h, l), li.Np, B, h->ws + li.ws_b_off, (size_t)li.Np);
        CSB_CUDA_CHECK(cudaGetLastError());
        prof_mark(h, K_COLSUM, ws);
        tab.seg[tab.n++] = {h->ws + li.ws_b_off, (size_t)li.Np, h->grads + li.b_off, (int64_t)li.Np, S};
      }
      return CSB_OK;
    };
    if (!conc) {
      int rcw = weight_grads(st);
      if (rcw) return rcw;
    } else {
      CSB_CUDA_CHECK(cudaEventRecord(h->ev_dz[l], st));
      // dZ_{l-1} goes into the buffer dZ_{l+1} lived in: its last reader, the weight-gradient GEMM of layer l+1, must be done
      if (l + 1 < h->L) CSB_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_w[l + 1], 0));
    }
    // ---- data gradient dZ_{l-1} = (dZ_l . W_l^T) * act'_{l-1}(act_{l-1})
    if (l > 0) {
      const LayerInfo& lp = h->layer[l - 1];
      if (h->bf16) {
        tc::GemmParams p = {};
        p.M = (int)B; p.N = li.Kp; p.K = li.Np; p.act = lp.act; p.alpha = lp.alpha; p.head_relu_from = -1;
        p.out = dz16(h, l - 1); p.ld_out = lp.Np;
        p.saved = reinterpret_cast<const __nv_bfloat16*>(h->act[l - 1]); p.ld_saved = lp.Np;
        p.dgrad_scale = h->drop_live ? 1.f / (1.f - h->dropout) : 0.f;
        int rc;
        if (h->amask[l - 1] != nullptr && !h->drop_live) {        // ReLU-family layer: act' from the forward pass's sign bits (no activation re-read)
          p.mask_in = h->amask[l - 1]; p.ld_mask = (int)h->cap;
          rc = launch_tn_auto<tc::EPI_DGRAD_MASK>(h->tm_dz[l].a_k128, h->tm_w[l], p, h->sm_count, st, &h->tm_w_s[l]);
        } else {
          rc = launch_tn_auto<tc::EPI_DGRAD>(h->tm_dz[l].a_k128, h->tm_w[l], p, h->sm_count, st, &h->tm_w_s[l]);
        }
        if (rc) return rc;
      } else if (h->tf32) {
        tc::GemmParams p = {};
        p.M = (int)B; p.N = li.Kp; p.K = li.Np; p.act = lp.act; p.alpha = lp.alpha; p.head_relu_from = -1;
        p.out = dz32(h, l - 1); p.ld_out = lp.Np;
        p.saved = reinterpret_cast<const __nv_bfloat16*>(h->act[l - 1]); p.ld_saved = lp.Np;      // fp32 in this mode (VAR_TF32 epilogue)
        p.dgrad_scale = h->drop_live ? 1.f / (1.f - h->dropout) : 0.f;
        if (h->x3) {
          simt::split3_f32_kernel<<<transpose_grid(h, B, li.Np), 256, 0, st>>>(dz32(h, l), li.Np, B, li.Np, h->sp_a, 3 * (int64_t)li.Np, li.Np, 0, 0, B);
          CSB_CUDA_CHECK(cudaGetLastError());
          p.K = 3 * li.Np; p.tf32_exact_store = 1;
        }
        // dZ_{l-1} also lands transposed, ready for its weight gradient -- unless a LayerNorm backward still has to turn du into dz
        h->dz_t_valid = !h->x3 && !lp.ln;
        if (h->dz_t_valid) { p.out_t = h->dz_t[(h->L - l) & 1]; p.ld_out_t = h->tr_ld; }
        int rc;
        if (!h->x3 && h->amask[l - 1] != nullptr && !h->drop_live) {      // act' from the sign bits: no fp32 activation re-read
          p.mask_in = h->amask[l - 1]; p.ld_mask = (int)h->cap;
          rc = launch_tn_tf32<tc::EPI_DGRAD_MASK>(h->tm32_dz[l], h->tm32_w[l], p, h->sm_count, st);
        } else {
          rc = launch_tn_tf32<tc::EPI_DGRAD>(h->tm32_dz[l], h->tm32_w[l], p, h->sm_count, st);
        }
        if (rc) return rc;
      } else {
        simt::SgemmParams p = {};
        p.M = (int)B; p.N = li.Kp; p.K = li.Np;
        p.A = dz32(h, l); p.lda = li.Np;
        p.B = h->params + li.w_off; p.ldb = li.Np;          // stored [Kp, Np] = [N_out, K_contract] -> TB
        p.C = dz32(h, l - 1); p.ldc = lp.Np;
        p.act = lp.act; p.alpha = lp.alpha; p.head_relu_from = -1;
        p.saved = reinterpret_cast<const float*>(h->act[l - 1]); p.ld_saved = lp.Np;
        p.dgrad_scale = h->drop_live ? 1.f / (1.f - h->dropout) : 0.f;
        dim3 grid((unsigned)(li.Kp / 64), (unsigned)ceil_div(B, 64));
        simt::sgemm_kernel<false, true, simt::SEPI_DGRAD><<<grid, 256, 0, st>>>(p);
        CSB_CUDA_CHECK(cudaGetLastError());
      }
      prof_mark(h, K_GEMM_DGRAD, st);
    } else if (dx != nullptr) {
      // dL/dxn = dZ_0 . W_0^T  (fp32 result), then / div if the forward normalised
      if (!h->dx_tmp) { CSB_ALLOC(h->dx_tmp, (size_t)h->cap * h->in_p * 4); }
      if (h->bf16) {
        tc::GemmParams p = {};
        p.M = (int)B; p.N = li.Kp; p.K = li.Np; p.out = h->dx_tmp; p.ld_out = h->in_p;
        int rc = launch_tn_auto<tc::EPI_F32>(h->tm_dz[0].a_k128, h->tm_w[0], p, h->sm_count, st, &h->tm_w_s[0]);
        if (rc) return rc;
      } else {
        simt::SgemmParams p = {};
        p.M = (int)B; p.N = li.Kp; p.K = li.Np;
        p.A = dz32(h, 0); p.lda = li.Np; p.B = h->params + li.w_off; p.ldb = li.Np; p.C = h->dx_tmp; p.ldc = h->in_p;
        dim3 grid((unsigned)(li.Kp / 64), (unsigned)ceil_div(B, 64));
        simt::sgemm_kernel<false, true, simt::SEPI_STORE><<<grid, 256, 0, st>>>(p);
        CSB_CUDA_CHECK(cudaGetLastError());
      }
      prof_mark(h, K_GEMM_DGRAD, st);
    }
    if (conc) {          // launched after the data-gradient GEMM so that the critical path takes the SMs first
      CSB_CUDA_CHECK(cudaStreamWaitEvent(h->side_stream, h->ev_dz[l], 0));
      int rcw = weight_grads(h->side_stream);
      if (rcw) return rcw;
      CSB_CUDA_CHECK(cudaEventRecord(h->ev_w[l], h->side_stream));
    }
  }
  if (conc) CSB_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_w[0], 0));      // the side stream is in order: layer 0 finishes last
  // ---- deterministic reduction of all split partials into the flat gradient buffer (one launch, blockIdx.y = segment);
  // with CSB_TRAIN_FUSED_OPT it is left to csb_mlp_apply_opt, which fuses it with the update
  if (defer_reduce) {
    h->pending = true; h->pending_B = B; h->pending_n_loss = n_loss_partials; h->pending_loss_out = loss_out;
  } else {
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(max_len / 4, 256), 2 * h->sm_count), (unsigned)(tab.n + (loss_out ? 1 : 0)));
    CSB_CUDA_CHECK(launch_pdl(simt::reduce_partials_kernel, grid, dim3(256), 0, st, tab));
    prof_mark(h, K_REDUCE, st);
  }
  return CSB_OK;
}

static int train_step_body(csb_mlp* h, const float* x, const float* y, int64_t B, float grad_scale, uint32_t flags, float* loss_out,
                           cudaStream_t st) {
  int rc;
  prof_mark(h, K_BEGIN, st);
  if ((rc = run_normalize(h, x, B, (flags & CSB_FWD_NORMALIZE_IN) ? 1 : 0, st))) return rc;
  if ((rc = run_hidden_forward(h, B, st, true))) return rc;
  int n_partials;
  const int l = h->L - 1;
  const bool tail = tail_fusable(h);
  if (tail) {
    if ((rc = run_tail(h, B, y, grad_scale, st))) return rc;
    n_partials = (int)ceil_div(B, 128) * tc::TN_EPI_WARPS;
  } else if (h->bf16) {
    if ((rc = run_head(h, B, 1, y, grad_scale, st))) return rc;
    n_partials = head_loss_partials(h, B);
  } else {
    if ((rc = run_head(h, B, 0, nullptr, 0.f, st))) return rc;
    const int grid = std::min(h->n_loss_partials, grid_for(B * h->out_p, 256, h->sm_count));
    simt::head_grad_kernel<float><<<grid, 256, 0, st>>>(h->pred, h->out_p, y, h->out_dim, h->d_loss_w, grad_scale, h->cfg.loss, 0,
                                                        h->layer[l].act, h->layer[l].alpha, h->cfg.head_relu_from, h->has_mask ? h->d_out_mask : nullptr,
                                                        dz32(h, l), h->out_p, B, h->out_dim, h->out_p, h->loss_partials);
    CSB_CUDA_CHECK(cudaGetLastError());
    prof_mark(h, K_LOSS, st);
    n_partials = grid;
  }
  // the scalar loss is summed inside the gradient-reduction launch at the end of the backward pass
  const bool defer = (flags & CSB_TRAIN_FUSED_OPT) != 0 && fused_opt_supported(h);
  if ((rc = run_backward_chain(h, B, nullptr, st, n_partials, loss_out, defer, tail))) return rc;
  h->acts_B = -1;
  return CSB_OK;
}


// The ~24 launches of a training step are replayed as one CUDA graph once the same (buffers, batch, scale) key has been
// seen twice: first sight runs eagerly (also warms the one-time kernel attribute setup), second sight captures on an
// internal stream (the caller's stream may be the legacy default stream, which cannot be captured) and instantiates.
int csb_mlp_train_step(csb_mlp* h, const float* x, const float* y, int64_t B, float grad_scale, uint32_t flags,
                       float* loss_out, void* stream) {
  CSB_REQUIRE(h && x && y, CSB_EINVAL, "null argument");
  int rc = check_batch(h, B);
  if (rc) return rc;
  CSB_REQUIRE(B > 0, CSB_EINVAL, "empty batch");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (grad_scale <= 0.f) grad_scale = 1.f / ((float)B * (float)h->out_dim);
  if ((rc = build_act_maps(h, B))) return rc;
  if ((rc = flush_pending(h, st))) return rc;          // a deferred reduction nobody consumed (no csb_mlp_apply_opt in between)
  float* lo = loss_out ? loss_out : h->d_loss;
  // (dropout: the per-step seeds are launch arguments, so those steps are not replayed from a captured graph)
  if (!h->graphs_on || h->prof_on || h->dropout > 0.f) return train_step_body(h, x, y, B, grad_scale, flags, lo, st);

  h->graph_clock++;
  for (auto& g : h->graphs) {
    if (g.x == x && g.y == y && g.B == B && g.gs == grad_scale && g.flags == flags && g.loss_out == lo && g.maps_B == h->maps_B) {
      g.use = h->graph_clock;
      if (g.exec == nullptr) {            // second sight: capture
        if (h->cap_stream == nullptr) CSB_CUDA_CHECK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
        const int64_t l0 = h->launches;
        CSB_CUDA_CHECK(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
        rc = train_step_body(h, x, y, B, grad_scale, flags, lo, h->cap_stream);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
        g.n_launches = h->launches - l0;
        h->launches = l0;
        if (rc || e != cudaSuccess || graph == nullptr) {
          cudaGetLastError();
          if (graph) cudaGraphDestroy(graph);
          h->graphs_on = false;           // capture not possible here: stay on the eager path for good
          return train_step_body(h, x, y, B, grad_scale, flags, lo, st);
        }
        e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { cudaGetLastError(); g.exec = nullptr; h->graphs_on = false; return train_step_body(h, x, y, B, grad_scale, flags, lo, st); }
      }
      CSB_CUDA_CHECK(cudaGraphLaunch(g.exec, st));
      h->launches += g.n_launches;
      h->acts_B = -1;
      if ((flags & CSB_TRAIN_FUSED_OPT) != 0 && fused_opt_supported(h)) {      // what run_backward_chain records when it runs eagerly
        h->pending = true; h->pending_B = B; h->pending_tail = tail_fusable(h);
        h->pending_n_loss = h->pending_tail ? (int)ceil_div(B, 128) * tc::TN_EPI_WARPS : head_loss_partials(h, B);
        h->pending_loss_out = lo;
      }
      return CSB_OK;
    }
  }
  // first sight: remember the key, run eagerly
  if (h->graphs.size() >= 8) {
    size_t lru = 0;
    for (size_t i = 1; i < h->graphs.size(); ++i) if (h->graphs[i].use < h->graphs[lru].use) lru = i;
    if (h->graphs[lru].exec) cudaGraphExecDestroy(h->graphs[lru].exec);
    h->graphs.erase(h->graphs.begin() + lru);
  }
  h->graphs.push_back({x, y, B, grad_scale, flags, lo, h->maps_B, nullptr, 0, h->graph_clock});
  return train_step_body(h, x, y, B, grad_scale, flags, lo, st);
}

int csb_mlp_backward(csb_mlp* h, const float* dy, float* dx, int64_t B, void* stream) {
  CSB_REQUIRE(h && dy, CSB_EINVAL, "null argument");
  CSB_REQUIRE(h->acts_B == B && B > 0, CSB_ESTATE, "csb_mlp_backward needs a preceding forward with CSB_FWD_KEEP_ACTIVATIONS on the same batch");
  CSB_REQUIRE(!(dx != nullptr && h->acts_normalized && h->d_xform != nullptr), CSB_EUNSUPPORTED,
              "dL/dx through the input transform of csb_mlp_set_input_transform is not available");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  { int rc0 = flush_pending(h, st); if (rc0) return rc0; }
  const int l = h->L - 1;
  const int grid = grid_for(B * h->out_p, 256, h->sm_count);
  prof_mark(h, K_BEGIN, st);
  if (h->bf16)
    simt::head_grad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(h->pred, h->out_p, dy, h->out_dim, h->d_loss_w, 1.f, 0, 1, h->layer[l].act,
                                                                 h->layer[l].alpha, h->cfg.head_relu_from, h->has_mask ? h->d_out_mask : nullptr, dz16(h, l), h->out_p,
                                                                 B, h->out_dim, h->out_p, nullptr);
  else
    simt::head_grad_kernel<float><<<grid, 256, 0, st>>>(h->pred, h->out_p, dy, h->out_dim, h->d_loss_w, 1.f, 0, 1, h->layer[l].act,
                                                         h->layer[l].alpha, h->cfg.head_relu_from, h->has_mask ? h->d_out_mask : nullptr, dz32(h, l), h->out_p, B,
                                                         h->out_dim, h->out_p, nullptr);
  CSB_CUDA_CHECK(cudaGetLastError());
  prof_mark(h, K_LOSS, st);
  int rc = run_backward_chain(h, B, dx, st);
  if (rc) return rc;
  if (dx) {
    // dx_tmp [B, in_p] -> dx [B, in_dim], divided by inp_div when the forward normalised: reuse scale_copy with 1/div
    // (inp_div entries that are 0 produced xn = 0 in the forward; their gradient is defined as 0)
    const int g2 = grid_for(B * h->in_dim, 256, h->sm_count);
    simt::scale_copy_kernel<<<g2, 256, 0, st>>>(h->dx_tmp, h->in_p, nullptr, dx, h->in_dim, B, h->in_dim);
    CSB_CUDA_CHECK(cudaGetLastError());
    prof_mark(h, K_MISC, st);
  }
  return CSB_OK;
}

int csb_mlp_apply_opt(csb_mlp* h, int rule, float lr, float beta1, float beta2, float eps, float wd, void* stream) {
  CSB_REQUIRE(h, CSB_EINVAL, "null handle");
  CSB_REQUIRE(rule >= CSB_OPT_ADAM_KERAS && rule <= CSB_OPT_RMSPROP, CSB_EINVAL, "unknown optimizer rule %d", rule);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  prof_mark(h, K_BEGIN, st);
  h->step++;
  simt::OptParams o;
  o.rule = rule; o.lr = lr; o.beta1 = beta1; o.beta2 = beta2; o.eps = eps; o.wd = wd;
  o.bc1 = (float)(1.0 - pow((double)beta1, (double)h->step));
  o.bc2 = (float)(1.0 - pow((double)beta2, (double)h->step));
  o.radam_r = -1.f;
  if (rule == CSB_OPT_RADAM) {
    const double t = (double)h->step, b2t = pow((double)beta2, t);
    const double sma_inf = 2.0 / (1.0 - (double)beta2) - 1.0;
    const double sma_t = sma_inf - 2.0 * t * b2t / (1.0 - b2t);
    if (sma_t >= 5.0) o.radam_r = (float)sqrt((sma_t - 4.0) / (sma_inf - 4.0) * (sma_t - 2.0) / (sma_inf - 2.0) * sma_inf / sma_t);
  }
  if (h->pending) {
    // split partials -> gradient -> update -> bf16 copies in one launch (bit-identical to the three-launch path)
    simt::FusedOptTable tab;
    tab.n = h->L;
    tab.params = h->params; tab.grads = h->grads; tab.m = h->m; tab.v = h->v;
    tab.loss_partials = h->loss_partials; tab.n_loss = h->pending_n_loss; tab.loss_out = h->pending_loss_out;
    int max_items = 1;
    for (int l = 0; l < h->L; ++l) {
      const LayerInfo& li = h->layer[l];
      const int splits = wgrad_splits(h, l, h->pending_B, nullptr, nullptr, h->pending_tail);
      tab.l[l] = {li.Kp, li.Np, li.w_off, li.b_off, h->w16[l], h->wt16[l], h->ws + li.ws_w_off, splits,
                  h->ws + li.ws_b_off, splits * nt_m_tiles(li.Kp, li.nt_cg),
                  li.g_off, h->ws + li.ws_g_off, li.ln ? ln_grad_splits(h, l, h->pending_B) : 0, simt::fused_opt_tk(splits)};
      max_items = std::max(max_items, (li.Kp / simt::fused_opt_tk(splits)) * (li.Np / 64) + (int)ceil_div(li.Np / 4, 256) + (li.ln ? (int)ceil_div(li.Np / 2, 256) : 0));
    }
    dim3 grid((unsigned)max_items, (unsigned)(h->L + 1));
    CSB_CUDA_CHECK(launch_pdl(simt::opt_fused_kernel, grid, dim3(256), 0, st, tab, o));
    prof_mark(h, K_OPT, st);
    h->pending = false;
    return CSB_OK;
  }
  if (fused_opt_supported(h)) {
    // gradients already reduced (data-parallel path: partial reduction -> all-reduce -> here): the same fused kernel with the
    // gradient buffer itself as the single "partial" updates the weights and writes the bf16 copies in one launch
    simt::FusedOptTable tab;
    tab.n = h->L;
    tab.params = h->params; tab.grads = h->grads; tab.m = h->m; tab.v = h->v;
    tab.loss_partials = nullptr; tab.n_loss = 0; tab.loss_out = nullptr;
    int max_items = 1;
    for (int l = 0; l < h->L; ++l) {
      const LayerInfo& li = h->layer[l];
      tab.l[l] = {li.Kp, li.Np, li.w_off, li.b_off, h->w16[l], h->wt16[l], h->grads + li.w_off, 1, h->grads + li.b_off, 1,
                  li.g_off, h->grads + li.g_off, li.ln ? 1 : 0, 32};
      max_items = std::max(max_items, (li.Kp / 32) * (li.Np / 64) + (int)ceil_div(li.Np / 4, 256) + (li.ln ? (int)ceil_div(li.Np / 2, 256) : 0));
    }
    dim3 grid((unsigned)max_items, (unsigned)(h->L + 1));
    CSB_CUDA_CHECK(launch_pdl(simt::opt_fused_kernel, grid, dim3(256), 0, st, tab, o));
    prof_mark(h, K_OPT, st);
    return CSB_OK;
  }
  const int grid = grid_for((int64_t)h->P_pad / 4, 256, h->sm_count);
  CSB_CUDA_CHECK(launch_pdl(simt::opt_kernel, dim3(grid), dim3(256), 0, st, h->params, h->grads, h->m, h->v, (int64_t)h->P_pad, o));
  prof_mark(h, K_OPT, st);
  return repack_weights(h, st);
}

static int ensure_stage(csb_mlp* h) {
  for (int i = 0; i < 2; ++i) {
    if (!h->x_stage[i]) { CSB_ALLOC(h->x_stage[i], (size_t)h->cap * h->in_dim * 4); }
    if (!h->y_stage[i]) { CSB_ALLOC(h->y_stage[i], (size_t)h->cap * h->out_dim * 4); }
    if (!h->ev_copied[i]) CSB_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
    if (!h->ev_released[i]) CSB_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_released[i], cudaEventDisableTiming));
  }
  // non-blocking: must not synchronise implicitly with the legacy default stream the caller may be computing on
  if (!h->copy_stream) CSB_CUDA_CHECK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  return CSB_OK;
}

int csb_mlp_stage_host_batch(csb_mlp* h, const float* x_host, const float* y_host, int64_t B, void* stream, float** x_dev, float** y_dev) {
  CSB_REQUIRE(h && x_host && x_dev, CSB_EINVAL, "null argument");
  CSB_REQUIRE((y_host == nullptr) == (y_dev == nullptr), CSB_EINVAL, "y_host and y_dev go together");
  int rc = check_batch(h, B);
  if (rc) return rc;
  if ((rc = ensure_stage(h))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int slot = h->stage_next;
  h->stage_next ^= 1;
  // the slot was last read by the step staged two calls ago: wait (on the host) until that step has released it.  This also
  // bounds how far the host runs ahead of the device (one step) and therefore how long a caller must keep its host buffers.
  if (h->released_valid[slot]) CSB_CUDA_CHECK(cudaEventSynchronize(h->ev_released[slot]));
  CSB_CUDA_CHECK(cudaMemcpyAsync(h->x_stage[slot], x_host, (size_t)B * h->in_dim * 4, cudaMemcpyHostToDevice, h->copy_stream));
  if (y_host) CSB_CUDA_CHECK(cudaMemcpyAsync(h->y_stage[slot], y_host, (size_t)B * h->out_dim * 4, cudaMemcpyHostToDevice, h->copy_stream));
  CSB_CUDA_CHECK(cudaEventRecord(h->ev_copied[slot], h->copy_stream));
  CSB_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_copied[slot], 0));
  h->staged_cur = slot;
  h->released_valid[slot] = false;
  *x_dev = h->x_stage[slot];
  if (y_dev) *y_dev = h->y_stage[slot];
  return CSB_OK;
}

int csb_mlp_release_staged(csb_mlp* h, void* stream) {
  CSB_REQUIRE(h, CSB_EINVAL, "null handle");
  CSB_REQUIRE(h->staged_cur >= 0, CSB_ESTATE, "no staged batch");
  CSB_CUDA_CHECK(cudaEventRecord(h->ev_released[h->staged_cur], reinterpret_cast<cudaStream_t>(stream)));
  h->released_valid[h->staged_cur] = true;
  return CSB_OK;
}

int csb_mlp_forward_host(csb_mlp* h, const float* x_host, float* y_pred_host, int64_t B, uint32_t flags, void* stream) {
  CSB_REQUIRE(h && x_host && y_pred_host, CSB_EINVAL, "null argument");
  int rc = check_batch(h, B);
  if (rc) return rc;
  if (B == 0) return CSB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* xd = nullptr;
  if ((rc = csb_mlp_stage_host_batch(h, x_host, nullptr, B, stream, &xd, nullptr))) return rc;
  float* yd = h->y_stage[h->staged_cur];               // the slot's target buffer doubles as the prediction staging
  if ((rc = csb_mlp_forward(h, xd, yd, B, flags, stream))) return rc;
  CSB_CUDA_CHECK(cudaMemcpyAsync(y_pred_host, yd, (size_t)B * h->out_dim * 4, cudaMemcpyDeviceToHost, st));
  if ((rc = csb_mlp_release_staged(h, stream))) return rc;
  CSB_CUDA_CHECK(cudaStreamSynchronize(st));
  return CSB_OK;
}

int csb_mlp_train_step_host_async(csb_mlp* h, const float* x_host, const float* y_host, int64_t B, float grad_scale, uint32_t flags,
                                  int rule, float lr, float beta1, float beta2, float eps, float wd, float* loss_host, void* stream) {
  CSB_REQUIRE(h && x_host && y_host, CSB_EINVAL, "null argument");
  int rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float *xd = nullptr, *yd = nullptr;
  if ((rc = csb_mlp_stage_host_batch(h, x_host, y_host, B, stream, &xd, &yd))) return rc;
  if ((rc = csb_mlp_train_step(h, xd, yd, B, grad_scale, flags | CSB_TRAIN_FUSED_OPT, nullptr, stream))) return rc;
  if ((rc = csb_mlp_release_staged(h, stream))) return rc;
  if ((rc = csb_mlp_apply_opt(h, rule, lr, beta1, beta2, eps, wd, stream))) return rc;
  if (loss_host) CSB_CUDA_CHECK(cudaMemcpyAsync(loss_host, h->d_loss, 4, cudaMemcpyDeviceToHost, st));
  return CSB_OK;
}

int csb_mlp_train_step_host(csb_mlp* h, const float* x_host, const float* y_host, int64_t B, float grad_scale, uint32_t flags,
                            int rule, float lr, float beta1, float beta2, float eps, float wd, float* loss_host, void* stream) {
  float loss = 0.f;      // pageable: the 4-byte D2H is staged by the driver and complete after the synchronise below
  int rc = csb_mlp_train_step_host_async(h, x_host, y_host, B, grad_scale, flags, rule, lr, beta1, beta2, eps, wd, &loss, stream);
  if (rc) return rc;
  CSB_CUDA_CHECK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  if (loss_host) *loss_host = loss;
  return CSB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// data_utils helpers
// ---------------------------------------------------------------------------------------------------------------
int csb_normalize(const float* x_raw, const float* sub, const float* div, float* x_out, int64_t N, int32_t F, void* stream) {
  CSB_REQUIRE(x_raw && sub && div && x_out && N >= 0 && F > 0, CSB_EINVAL, "bad argument");
  if (N == 0) return CSB_OK;
  int sm = 148;
  csb_device_info(&sm, nullptr, nullptr, nullptr);
  simt::normalize_kernel<<<grid_for(N * F, 256, sm), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x_raw, F, sub, div, 1, x_out, F, nullptr, 0, N, F, F);
  CSB_CUDA_CHECK(cudaGetLastError());
  return CSB_OK;
}
int csb_reshape_input_for_cnn(const float* x, float* out, int64_t N, void* stream) {
  CSB_REQUIRE(x && out && N >= 0, CSB_EINVAL, "bad argument");
  if (N == 0) return CSB_OK;
  simt::cnn_reshape_in_kernel<<<(unsigned)std::min<int64_t>(ceil_div(N * 360, 256), 148 * 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, N, 2, 4, 124);
  CSB_CUDA_CHECK(cudaGetLastError());
  return CSB_OK;
}
int csb_reshape_target_for_cnn(const float* y, float* out, int64_t N, void* stream) {
  CSB_REQUIRE(y && out && N >= 0, CSB_EINVAL, "bad argument");
  if (N == 0) return CSB_OK;
  simt::cnn_reshape_in_kernel<<<(unsigned)std::min<int64_t>(ceil_div(N * 600, 256), 148 * 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(y, out, N, 2, 8, 128);
  CSB_CUDA_CHECK(cudaGetLastError());
  return CSB_OK;
}
int csb_reshape_target_from_cnn(const float* p, float* out, int64_t N, void* stream) {
  CSB_REQUIRE(p && out && N >= 0, CSB_EINVAL, "bad argument");
  if (N == 0) return CSB_OK;
  simt::cnn_reshape_out_kernel<<<(unsigned)std::min<int64_t>(ceil_div(N * 128, 256), 148 * 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, out, N);
  CSB_CUDA_CHECK(cudaGetLastError());
  return CSB_OK;
}

int csb_eval_metrics(const float* pred, const float* target, const float* x_norm, int64_t N, int32_t ncol, const double* hyai,
                     const double* hybi, double p0, const double* area_wgt, const double* out_scale, double ps_mean, double ps_max,
                     double ps_min, int normalize, double* out, double* scratch, void* stream) {
  CSB_REQUIRE(pred && target && x_norm && hyai && hybi && area_wgt && out_scale && out && scratch, CSB_EINVAL, "null argument");
  CSB_REQUIRE(ncol >= 1 && N >= ncol && N % ncol == 0, CSB_EINVAL, "N (%lld) must be a positive multiple of ncol (%d)", (long long)N, ncol);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  simt::EvalConsts c;
  for (int i = 0; i < 61; ++i) { c.hyai[i] = hyai[i]; c.hybi[i] = hybi[i]; }
  c.p0 = p0; c.ps_mean = ps_mean; c.ps_span = ps_max - ps_min; c.grav = 9.80616; c.normalize = normalize;
  // energy-unit conversion (data_utils.py:480-494) and the small constant vectors live at the tail of the scratch buffer
  std::vector<double> host((size_t)ncol + 128 + 128);
  for (int i = 0; i < ncol; ++i) host[i] = area_wgt[i];
  for (int j = 0; j < 128; ++j) host[ncol + j] = out_scale[j];
  const double cp = 1.00464e3, lv = 2.501e6, rho_h2o = 1.0e3;
  for (int j = 0; j < 128; ++j) host[ncol + 128 + j] = j < 60 ? cp : (j < 120 ? lv : ((j == 122 || j == 123) ? lv * rho_h2o : 1.0));
  double* dconst = scratch + (size_t)4 * 128 * ncol;
  CSB_CUDA_CHECK(cudaMemcpyAsync(dconst, host.data(), host.size() * 8, cudaMemcpyHostToDevice, st));
  CSB_CUDA_CHECK(cudaStreamSynchronize(st));       // `host` goes out of scope
  simt::eval_metrics_kernel<<<ncol, 128, 0, st>>>(pred, target, x_norm, N / ncol, ncol, dconst, dconst + ncol, dconst + ncol + 128, c, scratch);
  CSB_CUDA_CHECK(cudaGetLastError());
  simt::eval_gridmean_kernel<<<4, 128, 0, st>>>(scratch, ncol, out);
  CSB_CUDA_CHECK(cudaGetLastError());
  return CSB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// kernel self-test hooks
// ---------------------------------------------------------------------------------------------------------------
int csb_test_gemm_tn(const uint16_t* A, const uint16_t* Bt, float* C, int M, int N, int K, int block_n, void* stream) {
  CSB_REQUIRE(A && Bt && C, CSB_EINVAL, "null argument");
  CSB_REQUIRE(N % 64 == 0 && K % 64 == 0 && M > 0, CSB_EINVAL, "N and K must be multiples of 64");
  CSB_REQUIRE(block_n == 128 || block_n == 256 || block_n == 512, CSB_EINVAL, "block_n must be 128, 256 (single CTA) or 512 (= 256 on CTA pairs)");
  const bool pairs = block_n == 512;
  if (pairs) block_n = 256;
  int sm = 0;
  int rc = csb_device_info(&sm, nullptr, nullptr, nullptr);
  if (rc) return rc;
  CUtensorMap ta, tb;
  if ((rc = make_tmap_bf16(&ta, A, K, M, K, 64, 128))) return rc;
  if ((rc = make_tmap_bf16(&tb, Bt, K, N, K, 64, (uint32_t)(std::min(N, block_n) / (pairs ? 2 : 1))))) return rc;
  tc::GemmParams p = {};
  p.M = M; p.N = N; p.K = K; p.out = C; p.ld_out = N;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pairs) return launch_tn<256, 6, tc::EPI_F32, 2>(ta, tb, p, sm, st);
  if (block_n == 256) return launch_tn<256, 4, tc::EPI_F32, 1>(ta, tb, p, sm, st);
  return launch_tn<128, 6, tc::EPI_F32, 1>(ta, tb, p, sm, st);
}

// C[M,N] (fp32) = A[M,K] . Bt[N,K]^T with fp32 operands through tcgen05 kind::tf32 (block_n as above)
int csb_test_gemm_tn_tf32(const float* A, const float* Bt, float* C, int M, int N, int K, int block_n, void* stream) {
  CSB_REQUIRE(A && Bt && C, CSB_EINVAL, "null argument");
  CSB_REQUIRE(N % 64 == 0 && K % 32 == 0 && M > 0, CSB_EINVAL, "N must be a multiple of 64 and K of 32");
  CSB_REQUIRE(block_n == 128 || block_n == 512, CSB_EINVAL, "block_n must be 128 (single CTA) or 512 (= 256 on CTA pairs)");
  const bool pairs = block_n == 512;
  int sm = 0;
  int rc = csb_device_info(&sm, nullptr, nullptr, nullptr);
  if (rc) return rc;
  CUtensorMap ta, tb;
  if ((rc = make_tmap_f32(&ta, A, K, M, K, 128))) return rc;
  if ((rc = make_tmap_f32(&tb, Bt, K, N, K, (uint32_t)(pairs ? std::min(N, 256) / 2 : std::min(N, 128))))) return rc;
  tc::GemmParams p = {};
  p.M = M; p.N = N; p.K = K; p.out = C; p.ld_out = N;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pairs) return launch_tn<256, 6, tc::EPI_F32, 2, tc::VAR_TF32>(ta, tb, p, sm, st);
  return launch_tn<128, 6, tc::EPI_F32, 1, tc::VAR_TF32>(ta, tb, p, sm, st);
}

static int g_test_dbg = 0;
static unsigned long long* g_test_stats = nullptr;
void csb_test_set_debug(int flags) { g_test_dbg = flags; }
void csb_test_set_stats(void* dev_u64x4_per_cta) { g_test_stats = reinterpret_cast<unsigned long long*>(dev_u64x4_per_cta); }

int csb_eval_crps(const void* samples, const void* target, int is_f64, int64_t n_tc, int L, int S, double* out, double* scratch, void* stream) {
  CSB_REQUIRE(samples && target && out && scratch, CSB_EINVAL, "null argument");
  CSB_REQUIRE(n_tc > 0 && L > 0 && S >= 2 && S <= 32, CSB_EINVAL, "need n_tc > 0, L > 0 and 2 <= S <= 32 ensemble members (got %lld, %d, %d)",
              (long long)n_tc, L, S);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int NB = (int)std::max<int64_t>(1, std::min<int64_t>(CSB_CRPS_BLOCKS, ceil_div(n_tc, 8)));
  dim3 grid((unsigned)L, (unsigned)NB);
  if (is_f64) simt::crps_kernel<double><<<grid, 256, 0, st>>>(reinterpret_cast<const double*>(samples), reinterpret_cast<const double*>(target), n_tc, L, S, scratch);
  else simt::crps_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(samples), reinterpret_cast<const float*>(target), n_tc, L, S, scratch);
  CSB_CUDA_CHECK(cudaGetLastError());
  simt::crps_finalize_kernel<<<(unsigned)ceil_div(L, 64), 64, 0, st>>>(scratch, NB, n_tc, L, out);
  CSB_CUDA_CHECK(cudaGetLastError());
  return CSB_OK;
}

static int* g_gather_bad = nullptr;                       // device flag: a gather saw an index outside [0, src_rows)

int csb_gather_rows(const float* src, const int64_t* idx, float* dst, int64_t n_rows, int row_len, int64_t src_rows, void* stream) {
  CSB_REQUIRE(src && idx && dst, CSB_EINVAL, "null argument");
  CSB_REQUIRE(n_rows >= 0 && row_len > 0 && src_rows > 0, CSB_EINVAL, "bad shape (n_rows %lld, row_len %d, src_rows %lld)", (long long)n_rows,
              row_len, (long long)src_rows);
  if (n_rows == 0) return CSB_OK;
  static int sm = 0;
  int rc;
  if (sm == 0 && (rc = csb_device_info(&sm, nullptr, nullptr, nullptr))) return rc;
  if (g_gather_bad == nullptr) {
    CSB_ALLOC(g_gather_bad, sizeof(int));
    CSB_CUDA_CHECK(cudaMemset(g_gather_bad, 0, sizeof(int)));
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = row_len % 4 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  const dim3 block(32, 8), grid((unsigned)std::min<int64_t>(ceil_div(n_rows, 8), (int64_t)sm * 16));
  if (vec) simt::gather_rows_kernel<true><<<grid, block, 0, st>>>(src, idx, dst, n_rows, row_len, src_rows, g_gather_bad);
  else simt::gather_rows_kernel<false><<<grid, block, 0, st>>>(src, idx, dst, n_rows, row_len, src_rows, g_gather_bad);
  CSB_CUDA_CHECK(cudaGetLastError());
  return CSB_OK;
}

int csb_gather_rows_check(void* stream) {
  if (g_gather_bad == nullptr) return CSB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int bad = 0;
  CSB_CUDA_CHECK(cudaMemcpyAsync(&bad, g_gather_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  CSB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (bad) {
    CSB_CUDA_CHECK(cudaMemsetAsync(g_gather_bad, 0, sizeof(int), st));
    set_last_error("csb_gather_rows: an index outside [0, src_rows) was skipped (its destination row is unwritten)");
    return CSB_EINVAL;
  }
  return CSB_OK;
}

int csb_test_linear_fwd(const uint16_t* A, const uint16_t* Wt, const float* bias, uint16_t* out, int M, int N, int K, int act,
                        float alpha, int pairs, void* stream) {
  CSB_REQUIRE(A && Wt && bias && out, CSB_EINVAL, "null argument");
  CSB_REQUIRE(N % 64 == 0 && K % 64 == 0 && M > 0, CSB_EINVAL, "N and K must be multiples of 64");
  static int sm = 0;
  int rc;
  if (sm == 0 && (rc = csb_device_info(&sm, nullptr, nullptr, nullptr))) return rc;
  const bool wide = N > 128, use_pairs = wide && pairs != 0;
  const int bn = wide ? 256 : 128;
  CUtensorMap ta, tb;
  if ((rc = make_tmap_bf16(&ta, A, K, M, K, 64, 128))) return rc;
  if ((rc = make_tmap_bf16(&tb, Wt, K, N, K, 64, (uint32_t)(std::min(N, bn) / (use_pairs ? 2 : 1))))) return rc;
  tc::GemmParams p = {};
  p.M = M; p.N = N; p.K = K; p.act = act; p.alpha = alpha; p.head_relu_from = -1; p.bias = bias; p.out = out; p.ld_out = N;
  p.dbg = g_test_dbg;
  p.stats = g_test_stats;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pairs == 2) {                       // the engine's own policy (tile shape, CTA pairs, staged epilogue for short contractions)
    CUtensorMap tb2;
    if ((rc = make_tmap_bf16(&tb2, Wt, K, N, K, 64, (uint32_t)tn_b_box_rows(N)))) return rc;
    return act == CSB_ACT_ELU ? launch_tn_shape<tc::EPI_BIAS_ACT, tc::VAR_ELU>(ta, tb2, p, sm, st) : launch_tn_shape<tc::EPI_BIAS_ACT, 0>(ta, tb2, p, sm, st);
  }
  if (use_pairs && p.stats != nullptr) return launch_tn<256, 6, tc::EPI_BIAS_ACT, 2, tc::VAR_STATS>(ta, tb, p, sm, st);   // probe_mainloop.py
  if (use_pairs) return launch_tn<256, 6, tc::EPI_BIAS_ACT, 2>(ta, tb, p, sm, st);
  if (wide) return launch_tn<256, 4, tc::EPI_BIAS_ACT, 1>(ta, tb, p, sm, st);
  return launch_tn<128, 6, tc::EPI_BIAS_ACT, 1>(ta, tb, p, sm, st);
}

int csb_test_gemm_nt(const uint16_t* A, const uint16_t* B, float* C, float* colsum, int M, int N, int Kr, int splits, void* stream) {
  CSB_REQUIRE(A && B && C, CSB_EINVAL, "null argument");
  CSB_REQUIRE(M % 64 == 0 && N % 64 == 0 && Kr > 0 && splits >= 1, CSB_EINVAL, "M and N must be multiples of 64");
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap_bf16(&ta, A, M, Kr, M, 64, 64))) return rc;
  if ((rc = make_tmap_bf16(&tb, B, N, Kr, N, 64, 64))) return rc;
  const int num_rb = (int)ceil_div(Kr, 64);
  tc::NtParams p = {};
  p.M = M; p.N = N; p.R = Kr;
  p.rb_per_split = (int)ceil_div(num_rb, splits);
  p.out = C; p.ld_out = N; p.split_stride = (size_t)M * N;    // caller provides splits * M * N floats
  p.colsum_out = colsum; p.colsum_stride = (size_t)N;         // optional: splits * N floats
  return launch_nt_auto(ta, tb, p, splits, reinterpret_cast<cudaStream_t>(stream));
}

// the same with an explicit tile policy: cg = 1 single CTAs, 2 CTA pairs, 0 what the engine would pick.  *m_tiles_out receives the
// number of bias-gradient partial rows per split (colsum must hold splits * m_tiles * N floats; ceil(M / 128) rows always suffice).
int csb_test_gemm_nt_cg(const uint16_t* A, const uint16_t* B, float* C, float* colsum, int M, int N, int Kr, int splits, int cg,
                        int* m_tiles_out, void* stream) {
  CSB_REQUIRE(A && B && C, CSB_EINVAL, "null argument");
  const bool uneven = splits < 0;           // negative: |splits| slots with the uneven split geometry for a half-width last n-block
  if (uneven) splits = -splits;
  CSB_REQUIRE(M % 64 == 0 && N % 64 == 0 && Kr > 0 && splits >= 1, CSB_EINVAL, "M and N must be multiples of 64");
  CSB_REQUIRE(cg >= 0 && cg <= 2, CSB_EINVAL, "cg must be 0, 1 or 2");
  if (cg == 0) cg = nt_cta_group(M, N);
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_tmap_bf16(&ta, A, M, Kr, M, 64, 64))) return rc;
  if ((rc = make_tmap_bf16(&tb, B, N, Kr, N, 64, 64))) return rc;
  const int num_rb = (int)ceil_div(Kr, 64);
  tc::NtParams p = {};
  p.M = M; p.N = N; p.R = Kr;
  p.rb_per_split = (int)ceil_div(num_rb, splits);
  p.out = C; p.ld_out = N; p.split_stride = (size_t)M * N;
  p.colsum_out = colsum; p.colsum_stride = (size_t)N;
  if (uneven) {
    const int bn = tn_block_n(N);
    CSB_REQUIRE(N > bn && N % bn == bn / 2 && splits >= 2, CSB_EINVAL, "uneven splits need a half-width last n-block and >= 2 splits");
    p.splits = splits; p.splits_narrow = (splits + 1) / 2;
    p.rb_per_split_narrow = (int)ceil_div(num_rb, p.splits_narrow);
  }
  if (m_tiles_out) *m_tiles_out = nt_m_tiles(M, cg);
  return launch_nt_auto(ta, tb, p, splits, reinterpret_cast<cudaStream_t>(stream), cg);
}


// ---------------------------------------------------------------------------------------------------------------
// Keras metrics of a batch (Trainer.fit / evaluate)
// ---------------------------------------------------------------------------------------------------------------
int csb_batch_metrics(const float* pred, const float* y, int64_t B, int32_t F, double* out5, double* scratch, void* stream) {
  CSB_REQUIRE(pred && y && out5 && scratch, CSB_EINVAL, "null argument");
  CSB_REQUIRE(B > 0 && F > 0, CSB_EINVAL, "bad shape (B %lld, F %d)", (long long)B, F);
  static int sm = 0;
  int rc;
  if (sm == 0 && (rc = csb_device_info(&sm, nullptr, nullptr, nullptr))) return rc;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(B, 8), std::min(4 * sm, (CSB_BATCH_METRICS_SCRATCH - 1) / 3)));
  simt::batch_metrics_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, y, B, F, out5, scratch);
  CSB_CUDA_CHECK(cudaGetLastError());
  return CSB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Heteroskedastic regression: one training step of BOTH networks inside the engine (hsr.py:122-140)
// ---------------------------------------------------------------------------------------------------------------
static int mlp_forward_keep(csb_mlp* h, const float* x, int64_t B, uint32_t flags, cudaStream_t st) {
  int rc;
  if ((rc = build_act_maps(h, B))) return rc;
  if ((rc = flush_pending(h, st))) return rc;
  prof_mark(h, K_BEGIN, st);
  if ((rc = run_normalize(h, x, B, (flags & CSB_FWD_NORMALIZE_IN) ? 1 : 0, st))) return rc;
  if ((rc = run_hidden_forward(h, B, st, true))) return rc;
  if ((rc = run_head(h, B, 0, nullptr, 0.f, st))) return rc;
  h->acts_B = B;
  h->acts_normalized = (flags & CSB_FWD_NORMALIZE_IN) != 0;
  return CSB_OK;
}

int csb_hsr_train_step(csb_mlp* mean, csb_mlp* logprec, const float* x, const float* y, int64_t B, int mle, uint32_t flags, int rule,
                       float lr, float beta1, float beta2, float eps, float wd_mean, float wd_logprec, float* loss_out, double* scratch,
                       void* stream) {
  CSB_REQUIRE(mean && logprec && x && y && loss_out && scratch, CSB_EINVAL, "null argument");
  CSB_REQUIRE(mean != logprec, CSB_EINVAL, "the two networks need separate handles");
  CSB_REQUIRE(mean->in_dim == logprec->in_dim && mean->out_dim == logprec->out_dim && mean->out_p == logprec->out_p && mean->bf16 == logprec->bf16,
              CSB_EINVAL, "the mean and log-precision networks must agree in input / output width and dtype");
  for (csb_mlp* h : {mean, logprec}) {
    CSB_REQUIRE(h->layer[h->L - 1].act == CSB_ACT_NONE && h->cfg.head_relu_from < 0 && !h->has_mask, CSB_EUNSUPPORTED,
                "csb_hsr_train_step expects linear output layers without a mask (hsr.py:27-35)");
    int rc = check_batch(h, B);
    if (rc) return rc;
  }
  CSB_REQUIRE(B > 0 && mean->out_dim % 4 == 0, CSB_EINVAL, "empty batch or output width not a multiple of 4");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc;
  if ((rc = mlp_forward_keep(mean, x, B, flags, st))) return rc;
  if (mle && (rc = mlp_forward_keep(logprec, x, B, flags, st))) return rc;      // the MSE phase never looks at the log-precision
  const int F = mean->out_dim, ldp = mean->out_p;
  const int64_t n = B * F;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n / 4, 256), std::min(4 * mean->sm_count, CSB_BATCH_METRICS_SCRATCH - 1)));
  simt::hsr_loss_kernel<<<grid, 256, 0, st>>>(mean->pred, logprec->pred, y, B, F, ldp, mle, loss_out, scratch);
  CSB_CUDA_CHECK(cudaGetLastError());
  prof_mark(mean, K_LOSS, st);
  const int lm = mean->L - 1, ll = logprec->L - 1;
  if (mean->bf16)
    simt::hsr_grad_kernel<__nv_bfloat16><<<grid_for(B * (ldp / 4), 256, mean->sm_count), 256, 0, st>>>(
        mean->pred, logprec->pred, y, B, F, ldp, mle, loss_out, dz16(mean, lm), mle ? dz16(logprec, ll) : nullptr);
  else
    simt::hsr_grad_kernel<float><<<grid_for(B * (ldp / 4), 256, mean->sm_count), 256, 0, st>>>(
        mean->pred, logprec->pred, y, B, F, ldp, mle, loss_out, dz32(mean, lm), mle ? dz32(logprec, ll) : nullptr);
  CSB_CUDA_CHECK(cudaGetLastError());
  prof_mark(mean, K_LOSS, st);
  // torch.optim.Adam skips parameters without a gradient: in the MSE phase the log-precision network is not touched at all
  // (no decay, no moment update, no step count) -- hsr.py:109-112,128-131
  csb_mlp* nets[2] = {mean, mle ? logprec : nullptr};
  const float wds[2] = {wd_mean, wd_logprec};
  for (int i = 0; i < 2; ++i) {
    csb_mlp* h = nets[i];
    if (!h) continue;
    const bool no_opt = (flags & CSB_HSR_NO_OPT) != 0;      // data parallelism: the caller all-reduces csb_mlp_grad_buffer, then csb_mlp_apply_opt
    if ((rc = run_backward_chain(h, B, nullptr, st, 0, nullptr, !no_opt && fused_opt_supported(h)))) return rc;
    h->acts_B = -1;
    if (!no_opt && (rc = csb_mlp_apply_opt(h, rule, lr, beta1, beta2, eps, wds[i], stream))) return rc;
  }
  return CSB_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// data parallelism over NVLink peer memory (dp_kernels.cuh)
// ---------------------------------------------------------------------------------------------------------------
static inline size_t dp_slab_bytes(int64_t n) { return (size_t)2 * n * 4 + 64 * 8; }

int csb_mlp_dp_export(csb_mlp* h, void* ipc_handle_out) {
  CSB_REQUIRE(h && ipc_handle_out, CSB_EINVAL, "null argument");
  CSB_REQUIRE(h->bf16, CSB_EUNSUPPORTED, "the peer-memory exchange is fused with the bf16 engine's optimizer launch");
  static_assert(sizeof(cudaIpcMemHandle_t) == CSB_IPC_HANDLE_BYTES, "CSB_IPC_HANDLE_BYTES");
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  if (h->dp_n == 0) {
    const int64_t n = (int64_t)h->P_pad + 4;
    float* slab = nullptr;
    CSB_ALLOC(slab, dp_slab_bytes(n));                    // zero-filled: flags and the grid-barrier counter start at 0
    CSB_CUDA_CHECK(cudaMemcpy(slab, h->grads, h->P_pad * 4, cudaMemcpyDeviceToDevice));
    cudaFree(h->grads);
    h->grads = slab;
    h->dp_n = n;
    for (auto& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);     // cached graphs hold the old gradient pointer
    h->graphs.clear();
  }
  cudaIpcMemHandle_t hd;
  CSB_CUDA_CHECK(cudaIpcGetMemHandle(&hd, h->grads));
  memcpy(ipc_handle_out, &hd, sizeof(hd));
  return CSB_OK;
}

int csb_mlp_dp_attach(csb_mlp* h, int rank, int world, const void* ipc_handles) {
  CSB_REQUIRE(h && ipc_handles, CSB_EINVAL, "null argument");
  CSB_REQUIRE(h->dp_n > 0, CSB_ESTATE, "csb_mlp_dp_export first");
  CSB_REQUIRE(world >= 2 && world <= simt::DP_MAX_RANKS && rank >= 0 && rank < world, CSB_EINVAL, "rank %d / world %d out of range", rank, world);
  CSB_REQUIRE(!h->dp_on, CSB_ESTATE, "already attached");
  for (int r = 0; r < world; ++r) {
    if (r == rank) { h->dp_peer_base[r] = h->grads; continue; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, reinterpret_cast<const char*>(ipc_handles) + (size_t)r * sizeof(hd), sizeof(hd));
    void* base = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      for (int q = 0; q < r; ++q) if (q != rank && h->dp_peer_base[q]) { cudaIpcCloseMemHandle(h->dp_peer_base[q]); h->dp_peer_base[q] = nullptr; }
      set_last_error("cudaIpcOpenMemHandle of rank %d's slab failed: %s (ranks must be GPUs of one node with peer access)", r, cudaGetErrorString(e));
      return CSB_EUNSUPPORTED;
    }
    h->dp_peer_base[r] = base;
  }
  h->dp_rank = rank; h->dp_world = world; h->dp_on = true; h->dp_epoch = 0;
  return CSB_OK;
}

int csb_mlp_dp_debug(csb_mlp* h, unsigned long long* stamps6_host) {
  CSB_REQUIRE(h && stamps6_host && h->dp_n > 0, CSB_EINVAL, "bad argument");
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  CSB_CUDA_CHECK(cudaMemcpy(stamps6_host, reinterpret_cast<unsigned long long*>(h->grads + 2 * h->dp_n) + 40, 6 * 8, cudaMemcpyDeviceToHost));
  return CSB_OK;
}

int csb_mlp_dp_step(csb_mlp* h, int rule, float lr, float beta1, float beta2, float eps, float wd, float* loss_out, void* stream) {
  CSB_REQUIRE(h, CSB_EINVAL, "null handle");
  CSB_REQUIRE(h->dp_on, CSB_ESTATE, "csb_mlp_dp_attach first");
  CSB_REQUIRE(rule >= CSB_OPT_ADAM_KERAS && rule <= CSB_OPT_RMSPROP, CSB_EINVAL, "unknown optimizer rule %d", rule);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  prof_mark(h, K_BEGIN, st);
  h->step++;
  simt::OptParams o;
  o.rule = rule; o.lr = lr; o.beta1 = beta1; o.beta2 = beta2; o.eps = eps; o.wd = wd;
  o.bc1 = (float)(1.0 - pow((double)beta1, (double)h->step));
  o.bc2 = (float)(1.0 - pow((double)beta2, (double)h->step));
  o.radam_r = -1.f;
  if (rule == CSB_OPT_RADAM) {
    const double t = (double)h->step, b2t = pow((double)beta2, t);
    const double sma_inf = 2.0 / (1.0 - (double)beta2) - 1.0, sma_t = sma_inf - 2.0 * t * b2t / (1.0 - b2t);
    if (sma_t >= 5.0) o.radam_r = (float)sqrt((sma_t - 4.0) / (sma_inf - 4.0) * (sma_t - 2.0) / (sma_inf - 2.0) * sma_inf / sma_t);
  }
  // phase 0: the split partials a CSB_TRAIN_FUSED_OPT step left behind (none: the gradient buffer is taken as it is)
  simt::SegmentTable seg;
  seg.n = 0; seg.loss_partials = h->loss_partials; seg.n_loss = 0; seg.loss_out = nullptr;
  if (h->pending) {
    seg.n_loss = h->pending_n_loss; seg.loss_out = h->grads + h->P_pad;      // (the kernel writes the loss share behind the gradient)
    for (int l = h->L - 1; l >= 0; --l) {
      const LayerInfo& li = h->layer[l];
      const int splits = wgrad_splits(h, l, h->pending_B, nullptr, nullptr, h->pending_tail);
      seg.seg[seg.n++] = {h->ws + li.ws_w_off, (size_t)li.Kp * li.Np, h->grads + li.w_off, (int64_t)li.Kp * li.Np, splits};
      seg.seg[seg.n++] = {h->ws + li.ws_b_off, (size_t)li.Np, h->grads + li.b_off, (int64_t)li.Np, splits * nt_m_tiles(li.Kp, li.nt_cg)};
      if (li.ln) seg.seg[seg.n++] = {h->ws + li.ws_g_off, (size_t)2 * li.Np, h->grads + li.g_off, (int64_t)2 * li.Np, ln_grad_splits(h, l, h->pending_B)};
    }
    h->pending = false;
  }
  simt::DpTable dp;
  dp.rank = h->dp_rank; dp.world = h->dp_world; dp.epoch = ++h->dp_epoch; dp.n = h->dp_n;
  dp.slice = (int64_t)ceil_div(ceil_div(h->dp_n, 4), h->dp_world) * 4;
  for (int r = 0; r < simt::DP_MAX_RANKS; ++r) {
    float* base = reinterpret_cast<float*>(h->dp_peer_base[r < h->dp_world ? r : h->dp_rank]);
    dp.grads_peer[r] = base; dp.gsum_peer[r] = base + h->dp_n;
    dp.flags_peer[r] = reinterpret_cast<unsigned long long*>(base + 2 * h->dp_n);
  }
  float* gsum = h->grads + h->dp_n;
  simt::FusedOptTable tab;
  tab.n = h->L;
  tab.params = h->params; tab.grads = h->grads; tab.m = h->m; tab.v = h->v;
  tab.loss_partials = nullptr; tab.n_loss = 0; tab.loss_out = loss_out ? loss_out : h->d_loss;
  int64_t items = 0;
  for (int l = 0; l < h->L; ++l) {
    const LayerInfo& li = h->layer[l];
    tab.l[l] = {li.Kp, li.Np, li.w_off, li.b_off, h->w16[l], h->wt16[l], gsum + li.w_off, 1, gsum + li.b_off, 1,
                li.g_off, gsum + li.g_off, li.ln ? 1 : 0, 32};
    dp.opt_item_base[l] = (int)items;
    items += (li.Kp / 32) * (li.Np / 64) + (int)ceil_div(li.Np / 4, 256) + (li.ln ? (int)ceil_div(li.Np / 2, 256) : 0);
  }
  dp.opt_item_base[h->L] = (int)items;
  dp.opt_items_total = items;
  dp.seg_base[0] = 0;
  for (int i = 0; i < seg.n; ++i) dp.seg_base[i + 1] = dp.seg_base[i] + seg.seg[i].len / 4;
  // as many blocks as are co-resident (the kernel spins on flags and on its own grid barrier); the SAME grid every step (the barrier
  // counter advances by 2 x grid per step); no programmatic overlap with neighbours
  static int occ = 0;
  if (occ == 0) {
    CSB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, simt::dp_reduce_opt_kernel, 256, 0));
    occ = std::max(1, std::min(occ, 4));
  }
  simt::dp_reduce_opt_kernel<<<h->sm_count * occ, 256, 0, st>>>(seg, dp, tab, o);
  CSB_CUDA_CHECK(cudaGetLastError());
  prof_mark(h, K_OPT, st);
  return CSB_OK;
}

}  // extern "C"

#include "cnn_engine.cuh"
