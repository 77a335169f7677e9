// tc_gemm.cuh -- hand-written Blackwell (sm_100a) tensor-core GEMMs for the dense layers of the column emulators.
//
// Two kernels, both: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory -> tcgen05.mma (kind::f16, bf16
// operands, fp32 accumulators in TMEM) -> tcgen05.ld -> fused epilogue.  Warp-specialised: warps 0-3 epilogue (one
// TMEM lane quadrant each), warp 4 = TMA producer, warp 5 = TMEM owner + single-thread MMA issuer.
//
//   gemm_tn_kernel   D[M,N] = A[M,K] . Bt[N,K]^T      both operands K-major.  Forward layers (A = activations,
//                    Bt = W^T) and data-gradient layers (A = dZ, Bt = W).  Persistent over 128 x BN tiles,
//                    STAGES-deep smem ring, two TMEM accumulator buffers so the epilogue of tile i overlaps the
//                    MMAs of tile i+1.  Epilogues: bias+activation->bf16, head+loss (+dZ), act'-masked dgrad, fp32.
//   gemm_nt_kernel   D[M,N] = A[R,M]^T . B[R,N]        both operands MN-major (row index R = batch is the
//                    contraction).  Weight gradients dW = H^T dZ.  One 128 x BN tile per CTA, split over R
//                    (blockIdx.y) into fp32 partials that a later kernel reduces deterministically.
//
// Shared-memory operand layouts are the canonical UMMA SWIZZLE_128B layouts (what TMA writes with
// CU_TENSOR_MAP_SWIZZLE_128B): K-major: rows of 64 bf16 (128 B), 8-row groups 1024 B apart (SBO);
// MN-major: 64 contiguous MN elements (128 B) per contraction row, 8-row groups 1024 B apart (SBO), 64-element MN
// chunks LBO apart.
#pragma once
#include "common.cuh"

namespace csb {
namespace tc {

constexpr int BM = 128;          // tile rows == UMMA M == TMEM lanes
constexpr int BK = 64;           // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;       // K per tcgen05.mma for 16-bit operands
constexpr int NUM_THREADS = 192; // 4 epilogue warps + producer + mma

enum Epi : int {
  EPI_BIAS_ACT = 0,   // out(bf16) = act(acc + bias)                       forward hidden layer
  EPI_HEAD_LOSS = 1,  // p = head(acc + bias); loss, dZ(bf16), optional p  forward output layer fused with the loss
  EPI_DGRAD = 2,      // out(bf16) = acc * act'(saved activation)          backward data gradient
  EPI_F32 = 3,        // out(fp32) = acc                                   (self-test, fp32 consumers)
  EPI_HEAD_OUT = 4    // p = head(acc + bias) -> fp32 (optionally / out_scale)   inference output layer
};

struct GemmParams {
  int M, N, K;                 // problem (N, K multiples of 64; M arbitrary)
  int b_box_rows;              // rows of the B-operand TMA box (== min(N, BN)): sets the expected transaction bytes
  int act;                     // CSB_ACT_* for EPI_BIAS_ACT / EPI_DGRAD (activation of the layer whose output is stored / was saved)
  float alpha;
  int head_relu_from;          // EPI_HEAD_*: columns >= this get ReLU (-1: none); otherwise `act` applies
  const float* bias;           // [N]
  void* out;                   // bf16 or fp32 [M, ld_out]
  int ld_out;
  const __nv_bfloat16* saved;  // EPI_DGRAD: saved activation [M, ld_saved]
  int ld_saved;
  // EPI_HEAD_LOSS / EPI_HEAD_OUT
  const float* y;              // targets [M, ld_y]
  int ld_y;
  int out_dim;                 // valid output columns (<= N)
  const float* loss_w;         // [N] (zero in padding)
  float grad_scale;            // multiplies loss and dL/dp
  int loss_kind;               // CSB_LOSS_*
  float* pred;                 // optional fp32 predictions [M, ld_pred]
  int ld_pred;
  const float* inv_out_scale;  // optional [N]: pred *= inv_out_scale
  float* loss_partials;        // [num_m_blocks * 4]
};

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Wait with a watchdog: a protocol bug becomes a trap (launch error) after ~2 s instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (globaltimer_ns() - t0 > 2000000000ull) __trap();
  }
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base_lane + i), columns [c, c+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// descriptors (bit layouts: cute/arch/mma_sm100_desc.hpp SmemDescriptor / InstrDescriptor)
// ---------------------------------------------------------------------------------------------------------------
// K-major SWIZZLE_128B operand tile: start address, LBO (unused for swizzled K-major, encoded 1), SBO = 1024 B,
// descriptor version 1 (bit 46), layout SWIZZLE_128B = 2 (bits 61-63).
__device__ __forceinline__ uint64_t make_desc_kmajor_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major SWIZZLE_128B operand tile: LBO = byte distance between 64-element MN chunks, SBO = 1024 B between 8-row
// contraction groups.
__device__ __forceinline__ uint64_t make_desc_mnmajor_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D fp32 (bits 4-5 = 1), A,B bf16 (bits 7-9 / 10-12 = 1), majors (bit 15/16: 1 = MN),
// N>>3 at bits 17-22, M>>4 at bits 24-28.
__device__ __forceinline__ uint32_t make_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------------
// epilogue helpers: each thread owns one output row (TMEM lane) and 32 consecutive columns
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32]) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
    u.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
    u.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
    u.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
    d[q] = u;
  }
}
__device__ __forceinline__ void store_f32x32(float* dst, const float (&v)[32]) {
  float4* d = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int q = 0; q < 8; ++q) d[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
__device__ __forceinline__ void load_f32x32(const float* src, float (&v)[32]) {
  const float4* s = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4 t = __ldg(s + q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}
__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* src, float (&v)[32]) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 t = __ldg(s + q);
    v[8 * q + 0] = bf16_lo(t.x); v[8 * q + 1] = bf16_hi(t.x);
    v[8 * q + 2] = bf16_lo(t.y); v[8 * q + 3] = bf16_hi(t.y);
    v[8 * q + 4] = bf16_lo(t.z); v[8 * q + 5] = bf16_hi(t.z);
    v[8 * q + 6] = bf16_lo(t.w); v[8 * q + 7] = bf16_hi(t.w);
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int row, int col, const uint32_t (&raw)[32],
                                               float& loss_acc) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);

  if constexpr (EPI == EPI_F32) {
    store_f32x32(reinterpret_cast<float*>(p.out) + (size_t)row * p.ld_out + col, v);
  } else if constexpr (EPI == EPI_BIAS_ACT) {
    float b[32];
    load_f32x32(p.bias + col, b);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = act_fwd(p.act, p.alpha, v[j] + b[j]);
    store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ld_out + col, v);
  } else if constexpr (EPI == EPI_DGRAD) {
    float a[32];
    load_bf16x32(p.saved + (size_t)row * p.ld_saved + col, a);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= act_bwd_from_out(p.act, p.alpha, a[j]);
    store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ld_out + col, v);
  } else if constexpr (EPI == EPI_HEAD_OUT) {
    float b[32];
    load_f32x32(p.bias + col, b);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float z = v[j] + b[j];
      const bool relu_col = p.head_relu_from >= 0 && (col + j) >= p.head_relu_from;
      v[j] = relu_col ? fmaxf(z, 0.f) : act_fwd(p.act, p.alpha, z);
    }
    if (p.inv_out_scale != nullptr) {
      load_f32x32(p.inv_out_scale + col, b);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= b[j];
    }
    if (col + 32 <= p.out_dim) {
      store_f32x32(p.pred + (size_t)row * p.ld_pred + col, v);
    } else {
      for (int j = 0; j < 32; ++j)
        if (col + j < p.out_dim) p.pred[(size_t)row * p.ld_pred + col + j] = v[j];
    }
  } else {  // EPI_HEAD_LOSS
    float b[32], dz[32];
    load_f32x32(p.bias + col, b);
    // p = head activation; g = d p / d z expressed through p
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float z = v[j] + b[j];
      const bool relu_col = p.head_relu_from >= 0 && (col + j) >= p.head_relu_from;
      float pv = relu_col ? fmaxf(z, 0.f) : act_fwd(p.act, p.alpha, z);
      dz[j] = relu_col ? (pv > 0.f ? 1.f : 0.f) : act_bwd_from_out(p.act, p.alpha, pv);
      v[j] = pv;
    }
    if (p.pred != nullptr && col + 32 <= p.out_dim) store_f32x32(p.pred + (size_t)row * p.ld_pred + col, v);
    float w[32], yv[32];
    load_f32x32(p.loss_w + col, w);
    if (col + 32 <= p.out_dim) {
      load_f32x32(p.y + (size_t)row * p.ld_y + col, yv);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) yv[j] = (col + j < p.out_dim) ? __ldg(p.y + (size_t)row * p.ld_y + col + j) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float d = v[j] - yv[j];
      if (p.loss_kind == CSB_LOSS_MSE) {
        loss_acc += w[j] * d * d;
        dz[j] *= 2.f * w[j] * d * p.grad_scale;
      } else {
        loss_acc += w[j] * fabsf(d);
        dz[j] *= w[j] * p.grad_scale * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
      }
    }
    store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ld_out + col, dz);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// gemm_tn_kernel: persistent, K-major x K-major
// ---------------------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
struct TnSmem {
  static constexpr int A_BYTES = BM * BK * 2;   // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // barriers + slack for manual 1024 B alignment
};

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using L = TnSmem<BN, STAGES>;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static_assert(2 * BN <= 512, "two accumulator buffers must fit the 512 TMEM columns");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B atoms are 1024 B aligned
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + L::BAR_OFFSET + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m_blocks = (p.M + BM - 1) / BM;
  const int num_n_blocks = (p.N + BN - 1) / BN;
  const int num_tiles = num_m_blocks * num_n_blocks;
  const int num_kb = p.K / BK;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n_blocks) * BM, n0 = (tile % num_n_blocks) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), (uint32_t)(L::A_BYTES + p.b_box_rows * BK * 2));
          const uint32_t sa = smem_base + s * L::STAGE_BYTES;
          tma_load_2d(sa, &tmap_a, full_bar(s), kb * BK, m0);
          tma_load_2d(sa + L::A_BYTES, &tmap_b, full_bar(s), kb * BK, n0);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0; int t = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const int n0 = (tile % num_n_blocks) * BN;
        const int n_valid = min(BN, p.N - n0);
        const uint32_t idesc = make_idesc_bf16(BM, n_valid, 0, 0);
        const int acc = t & 1;
        mbar_wait(tempty_bar(acc), ((uint32_t)(t >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * L::STAGE_BYTES;
          const uint64_t da = make_desc_kmajor_sw128(sa);
          const uint64_t db = make_desc_kmajor_sw128(sa + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(empty_bar(s));            // smem slot reusable once these MMAs have read it
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        umma_commit(tfull_bar(acc));            // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue (warps 0-3 <-> TMEM lanes 32w .. 32w+31) =====================
    int t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int mb = tile / num_n_blocks;
      const int m0 = mb * BM, n0 = (tile % num_n_blocks) * BN;
      const int n_valid = min(BN, p.N - n0);
      const int acc = t & 1;
      const int row = m0 + warp * 32 + lane;
      mbar_wait(tfull_bar(acc), (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
      float loss_acc = 0.f;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN);
      for (int c = 0; c < n_valid; c += 32) {
        uint32_t raw[32];
        tmem_ld_32x32(taddr + (uint32_t)c, raw);
        tmem_ld_wait();
        if (row < p.M) epilogue_chunk<EPI>(p, row, n0 + c, raw, loss_acc);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if constexpr (EPI == EPI_HEAD_LOSS) {
        // deterministic: one partial per (m-block, warp); requires num_n_blocks == 1 (host asserts)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
        if (lane == 0) p.loss_partials[mb * 4 + warp] = loss_acc * p.grad_scale;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// gemm_nt_kernel: D[M,N] = sum_r A[r, m] * B[r, n]; both operands MN-major; split over r (blockIdx.y)
// ---------------------------------------------------------------------------------------------------------------
struct NtParams {
  int M, N, R;             // M, N feature dims (multiples of 64), R rows to contract
  int rb_per_split;        // 64-row blocks per split
  float* out;              // partials: out + split * split_stride + m * ld_out + n
  int ld_out;
  size_t split_stride;
};

template <int BN, int STAGES>
struct NtSmem {
  static constexpr int A_BYTES = BK * BM * 2;   // 64 contraction rows x 128 m (2 chunks of [64 rows x 128 B])
  static constexpr int B_BYTES = BK * BN * 2;   // BN/64 chunks
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const NtParams p) {
  using L = NtSmem<BN, STAGES>;
  constexpr uint32_t TMEM_COLS = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : (BN <= 256) ? 256 : 512;
  constexpr int CHUNK_BYTES = BK * 128;   // one [64 rows x 64 elements] box = 8 KB

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + L::BAR_OFFSET + 8 * (2 * STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_n_blocks = (p.N + BN - 1) / BN;
  const int m0 = (blockIdx.x / num_n_blocks) * BM, n0 = (blockIdx.x % num_n_blocks) * BN;
  const int m_valid = min(BM, p.M - m0), n_valid = min(BN, p.N - n0);
  const int num_rb = (p.R + BK - 1) / BK;
  const int rb_begin = blockIdx.y * p.rb_per_split;
  const int rb_end = min(num_rb, rb_begin + p.rb_per_split);
  const int nrb = max(0, rb_end - rb_begin);

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      const int a_chunks = (m_valid + 63) / 64, b_chunks = (n_valid + 63) / 64;
      for (int rb = rb_begin; rb < rb_end; ++rb) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), (uint32_t)((a_chunks + b_chunks) * CHUNK_BYTES));
        const uint32_t sa = smem_base + s * L::STAGE_BYTES;
        for (int c = 0; c < a_chunks; ++c) tma_load_2d(sa + c * CHUNK_BYTES, &tmap_a, full_bar(s), m0 + 64 * c, rb * BK);
        for (int c = 0; c < b_chunks; ++c) tma_load_2d(sa + L::A_BYTES + c * CHUNK_BYTES, &tmap_b, full_bar(s), n0 + 64 * c, rb * BK);
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && nrb > 0) {
      int s = 0; uint32_t ph = 0;
      // M is always issued as 128: rows >= m_valid read stale smem (never stored by the epilogue)
      const uint32_t idesc = make_idesc_bf16(BM, n_valid, 1, 1);
      for (int i = 0; i < nrb; ++i) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * L::STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // 16 contraction rows = 2 groups of 8 rows x 128 B = 2048 B further into every chunk
          const uint64_t da = make_desc_mnmajor_sw128(sa + k * 2048, CHUNK_BYTES);
          const uint64_t db = make_desc_mnmajor_sw128(sa + L::A_BYTES + k * 2048, CHUNK_BYTES);
          umma_f16(tmem_base, da, db, idesc, (uint32_t)((i | k) != 0));
        }
        umma_commit(empty_bar(s));
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      umma_commit(tfull_bar);
    }
  } else {
    const int row = m0 + warp * 32 + lane;
    float* out = p.out + (size_t)blockIdx.y * p.split_stride + (size_t)row * p.ld_out + n0;
    if (nrb > 0) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int c = 0; c < n_valid; c += 32) {
        uint32_t raw[32];
        tmem_ld_32x32(taddr + (uint32_t)c, raw);
        tmem_ld_wait();
        if (row < p.M) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          store_f32x32(out + c, v);
        }
      }
    } else if (row < p.M) {
      for (int c = 0; c < n_valid; c += 4) *reinterpret_cast<float4*>(out + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace csb
