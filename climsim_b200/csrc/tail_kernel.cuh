// tail_kernel.cuh -- the 128-wide tail of the training step as ONE persistent kernel (sm_100a, tcgen05 / TMEM / TMA):
//
//     output layer  z = a W + b  ->  head activation  ->  weighted MSE + dL/dz   (was gemm_tn_kernel<EPI_HEAD_LOSS>)
//     data gradient dZ_prev = (dZ W^T) * act'_prev                               (was gemm_tn_kernel<EPI_DGRAD_MASK>)
//     weight / bias gradient  dW = a^T dZ,  db = colsum(dZ)                      (was gemm_nt_kernel)
//
// for an output layer of 128 -> 128 (padded) columns, i.e. MLP_v1's Dense(128) -> [Dense(120) | relu(Dense(8))] pair
// (hpo_baseline_v1.py:89-101) with Keras 'mse' (:127-129).  A CTA walks 128-row blocks of the batch; per block the activation tile
// a [128 x 128] arrives ONCE by TMA and serves as the K-major A operand of the output GEMM and, read the other way round (the same
// bytes are a valid MN-major SWIZZLE_128B tile), as the A operand of the weight gradient; dZ never goes to HBM: the loss epilogue
// writes it as bf16 into shared memory in the UMMA operand layout, where it is the A operand of the data gradient and the B operand of
// the weight gradient.  Both weight copies (W^T for the forward, W for the data gradient: 32 KB each) stay resident in shared memory;
// dW accumulates in TMEM over all blocks of the CTA and leaves as one fp32 partial per CTA (reduced in a fixed order with every other
// gradient), db as column sums of the rounded dZ (warp butterfly, fixed order).
//
// Against the three separate launches this removes two kernel launches, the write + two reads of dZ (3 x 16.8 MB at B = 65 536),
// the second and third read of a (2 x 16.8 MB), and the latency-bound one-tile-at-a-time structure of the 128 x 128 x 128 head GEMM.
//
// TMEM (512 columns): [0,128) and [128,256) output accumulators (double-buffered), [256,384) data-gradient accumulator,
// [384,512) weight-gradient accumulator (persistent).  Shared memory: 3 x 32 KB activation tiles, 2 x 32 KB dZ tiles, 2 x 32 KB weights.
// Warps: 0 = TMEM owner + MMA issuer, 1 = TMA producer, 2..17 = epilogue (lane quadrant = warp % 4, 32-column slice = (warp - 2) / 4).
#pragma once
#include "tc_gemm.cuh"

namespace csb {
namespace tc {

struct TailParams {
  int M;                        // rows (batch)
  int act;                      // head activation of columns < head_relu_from (CSB_ACT_NONE / RELU / LEAKYRELU)
  float alpha;
  int head_relu_from;           // columns >= this get ReLU (-1: none)
  const float* bias;            // [128] (zero in padding)
  const float* loss_w;          // [128] (zero in padding)
  const float* y;               // targets [M, ld_y]
  int ld_y, out_dim;
  float grad_scale;
  float* loss_partials;         // [ceil(M/128) * TN_EPI_WARPS]
  int prev_act;                 // activation of the layer below (its sign mask gates the data gradient)
  float prev_alpha;
  const uint32_t* mask_in;      // sign bits of the previous activation, chunk-major [4][ld_mask]
  int ld_mask;
  __nv_bfloat16* dz_prev;       // out: dZ of the layer below [M, ld_dz]
  int ld_dz;
  float* dw_out;                // out: [gridDim.x][128 x 128] fp32 partials of dW (row = input feature, column = output feature)
  float* db_out;                // out: [gridDim.x][128] fp32 partials of db
};

struct TailSmem {
  static constexpr int TILE = 128 * 128 * 2;            // one [128 rows x 128 columns] bf16 tile = two 64-column SWIZZLE_128B chunks
  static constexpr int CHUNK = 128 * 64 * 2;
  static constexpr int A_STAGES = 3, Z_STAGES = 2;
  static constexpr int A_OFF = 0;
  static constexpr int Z_OFF = A_OFF + A_STAGES * TILE;
  static constexpr int WT_OFF = Z_OFF + Z_STAGES * TILE;   // W^T [n][k]: B operand of the output GEMM
  static constexpr int W_OFF = WT_OFF + TILE;              // W [k][n]: B operand of the data gradient
  static constexpr int VEC_OFF = W_OFF + TILE;             // bias [128], loss weights [128]
  static constexpr int BAR_OFF = VEC_OFF + 2 * 128 * 4;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
  static_assert(TOTAL <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
};

// column sums over the 32 lanes of a warp of x[32] (lane = row, index = column): lane l ends up with the sum of column l.
// Reduce-scatter butterfly: 31 shuffles instead of 160.
__device__ __forceinline__ float warp_colsum32(float (&x)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float send = up ? x[j] : x[j + s];
      const float keep = up ? x[j + s] : x[j];
      x[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return x[0];
}

__global__ void __launch_bounds__(TN_THREADS, 1)
tail_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_wt, const __grid_constant__ CUtensorMap tmap_w,
            const TailParams p) {
  using L = TailSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  float* sbias = reinterpret_cast<float*>(smem_gen + L::VEC_OFF);
  float* sloss_w = sbias + 128;
  float* sdb = reinterpret_cast<float*>(smem_gen + L::Z_OFF);      // [4 lane quadrants][128]: aliases a dZ tile, used after the last MMA
  const uint32_t bar_base = smem_base + L::BAR_OFF;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (3 + s); };
  auto acc_full = [&](int a) { return bar_base + 8u * (6 + a); };
  auto acc_empty = [&](int a) { return bar_base + 8u * (8 + a); };
  auto z_full = [&](int a) { return bar_base + 8u * (10 + a); };
  auto z_empty = [&](int a) { return bar_base + 8u * (12 + a); };
  const uint32_t dz_full = bar_base + 8u * 14, dz_empty = bar_base + 8u * 15, w_full = bar_base + 8u * 16, dw_full = bar_base + 8u * 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + L::BAR_OFF + 8 * 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_blocks = (p.M + BM - 1) / BM;
  const int my_blocks = ((int)blockIdx.x < num_blocks) ? (num_blocks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == TN_PRODUCER_WARP && lane == 0) {
    tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_wt); tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < 3; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1); mbar_init(acc_empty(a), TN_EPI_WARPS);
      mbar_init(z_full(a), TN_EPI_WARPS); mbar_init(z_empty(a), 1);
    }
    mbar_init(dz_full, 1); mbar_init(dz_empty, TN_EPI_WARPS); mbar_init(w_full, 1); mbar_init(dw_full, 1);
    fence_mbar_init();
  }
  if (warp == TN_MMA_WARP) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < 128; i += TN_THREADS) { sbias[i] = __ldg(p.bias + i); sloss_w[i] = __ldg(p.loss_w + i); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t COL_ACC = 0, COL_DZ = 256, COL_DW = 384;

  if (warp == TN_PRODUCER_WARP) {
    // ===================== TMA producer: the two weight tiles once, then one activation tile per block =====================
    if (lane == 0) {
      mbar_expect_tx(w_full, 2u * L::TILE);
      for (int c = 0; c < 2; ++c) {
        tma_load_2d(smem_base + L::WT_OFF + c * L::CHUNK, &tmap_wt, w_full, 64 * c, 0);
        tma_load_2d(smem_base + L::W_OFF + c * L::CHUNK, &tmap_w, w_full, 64 * c, 0);
      }
      for (int i = 0; i < my_blocks; ++i) {
        const int s = i % 3;
        const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * BM;
        mbar_wait(a_empty(s), ((uint32_t)(i / 3) & 1u) ^ 1u);
        mbar_expect_tx(a_full(s), (uint32_t)L::TILE);
        for (int c = 0; c < 2; ++c) tma_load_2d(smem_base + L::A_OFF + s * L::TILE + c * L::CHUNK, &tmap_a, a_full(s), 64 * c, m0);
      }
    }
  } else if (warp == TN_MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0 && my_blocks > 0) {
      const uint32_t idesc_kk = make_idesc_bf16(BM, 128, 0, 0);        // K-major x K-major: output GEMM, data gradient
      const uint32_t idesc_mn = make_idesc_bf16(BM, 128, 1, 1);        // MN-major x MN-major: weight gradient (contraction = rows)
      auto backward = [&](int j) {
        const uint32_t sa = smem_base + L::A_OFF + (j % 3) * L::TILE, sz = smem_base + L::Z_OFF + (j & 1) * L::TILE;
        mbar_wait(z_full(j & 1), (uint32_t)(j >> 1) & 1u);
        mbar_wait(dz_empty, ((uint32_t)j & 1u) ^ 1u);
        tc_fence_after();
        // dZ_prev accumulator = dZ [128 x 128] . W [k][n]^T   (contraction over the output features)
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_f16(tmem_base + COL_DZ, make_desc_kmajor_sw128(sz + kb * L::CHUNK) + (uint64_t)(2 * k),
                     make_desc_kmajor_sw128(smem_base + L::W_OFF + kb * L::CHUNK) + (uint64_t)(2 * k), idesc_kk, (uint32_t)((kb | k) != 0));
        umma_commit(dz_full);
        // dW += a^T . dZ: both tiles read MN-major (64-column chunks L::CHUNK apart, 16 rows = 2048 B per instruction)
#pragma unroll
        for (int k = 0; k < BM / UMMA_K; ++k)
          umma_f16(tmem_base + COL_DW, make_desc_mnmajor_sw128(sa + k * 2048, L::CHUNK), make_desc_mnmajor_sw128(sz + k * 2048, L::CHUNK), idesc_mn,
                   (uint32_t)((j | k) != 0));
        umma_commit(a_empty(j % 3));           // every MMA that reads this activation tile / this dZ tile has completed
        umma_commit(z_empty(j & 1));
      };
      mbar_wait(w_full, 0);
      for (int i = 0; i < my_blocks; ++i) {
        const uint32_t sa = smem_base + L::A_OFF + (i % 3) * L::TILE;
        mbar_wait(a_full(i % 3), (uint32_t)(i / 3) & 1u);
        mbar_wait(acc_empty(i & 1), ((uint32_t)(i >> 1) & 1u) ^ 1u);
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_f16(tmem_base + COL_ACC + (uint32_t)(128 * (i & 1)), make_desc_kmajor_sw128(sa + kb * L::CHUNK) + (uint64_t)(2 * k),
                     make_desc_kmajor_sw128(smem_base + L::WT_OFF + kb * L::CHUNK) + (uint64_t)(2 * k), idesc_kk, (uint32_t)((kb | k) != 0));
        umma_commit(acc_full(i & 1));
        if (i > 0) backward(i - 1);            // one block behind: its dZ tile was written while this block's output GEMM ran
      }
      backward(my_blocks - 1);
      umma_commit(dw_full);
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - TN_FIRST_EPI_WARP, q = warp & 3, cq = ew >> 2;
    const int tile_row = q * 32 + lane, c0 = 32 * cq;
    const float slope = p.act == CSB_ACT_RELU ? 0.f : (p.act == CSB_ACT_LEAKYRELU ? p.alpha : 1.f);
    const bool relu_cols = p.head_relu_from >= 0 && c0 + 32 > p.head_relu_from;
    const float two_gs = 2.f * p.grad_scale;
    const float neg_prev = p.prev_act == CSB_ACT_LEAKYRELU ? p.prev_alpha : 0.f;
    // this thread's row inside a [128 x 64] SWIZZLE_128B chunk, and the 16-byte pieces its 32 columns occupy
    const uint32_t row_off = (uint32_t)((tile_row >> 3) * 1024 + (tile_row & 7) * 128) + (uint32_t)((c0 >> 6) * L::CHUNK);
    const int piece0 = (c0 & 63) >> 3;
    float db_acc = 0.f;                          // column c0 + lane of db, summed over this warp's rows of every block

    // the targets (fp32, 128 B per thread and block) come straight from HBM: their latency is hidden by asking L2 for the NEXT
    // block's lines one block ahead (no registers, no shared memory -- both are spoken for)
    auto prefetch_block = [&](int i) {
      const int grow = ((int)blockIdx.x + i * (int)gridDim.x) * BM + tile_row;
      if (i < my_blocks && grow < p.M) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.y + (size_t)grow * p.ld_y + c0));
        if (cq == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.mask_in + grow));
      }
    };
    auto loss_phase = [&](int i) {
      const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * BM, grow = m0 + tile_row;
      const bool row_ok = grow < p.M;
      prefetch_block(i + 1);
      float yv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) yv[j] = 0.f;
      if (row_ok) {
        const float* yptr = p.y + (size_t)grow * p.ld_y + c0;
        if (c0 + 32 <= p.out_dim && (p.ld_y & 7) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 31) == 0) load_f32x32_global_v8(yptr, yv);
        else
          for (int j = 0; j < 32; ++j) if (c0 + j < p.out_dim) yv[j] = __ldg(yptr + j);
      }
      mbar_wait_long(acc_full(i & 1), (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      uint32_t raw[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + COL_ACC + (uint32_t)(128 * (i & 1) + c0), raw);
      tmem_ld_wait();
      float v[32], loss_acc = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 b4 = *reinterpret_cast<const float4*>(sbias + c0 + 4 * g), w4 = *reinterpret_cast<const float4*>(sloss_w + c0 + 4 * g);
        const float b[4] = {b4.x, b4.y, b4.z, b4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 4 * g + j;
          const float z = __uint_as_float(raw[c]) + b[j];
          const float sc = (relu_cols && c0 + c >= p.head_relu_from) ? 0.f : slope;
          const float sl = z > 0.f ? 1.f : sc;
          const float d = sl * z - yv[c];
          const float wd = w[j] * d;
          loss_acc = fmaf(wd, d, loss_acc);
          v[c] = row_ok ? sl * (two_gs * wd) : 0.f;
        }
      }
      if (!row_ok) loss_acc = 0.f;
      // dZ (bf16) into the operand tile: wait until the MMAs of two blocks ago have finished reading this buffer
      mbar_wait(z_empty(i & 1), ((uint32_t)(i >> 1) & 1u) ^ 1u);
      const uint32_t zrow = smem_base + L::Z_OFF + (i & 1) * L::TILE + row_off;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const uint32_t w0 = pack_bf16x2(v[8 * h], v[8 * h + 1]), w1 = pack_bf16x2(v[8 * h + 2], v[8 * h + 3]);
        const uint32_t w2 = pack_bf16x2(v[8 * h + 4], v[8 * h + 5]), w3 = pack_bf16x2(v[8 * h + 6], v[8 * h + 7]);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(zrow + (uint32_t)(((piece0 + h) ^ (tile_row & 7)) << 4)), "r"(w0), "r"(w1), "r"(w2),
                     "r"(w3) : "memory");
        // the bias gradient sums what the weight gradient sees: the ROUNDED values
        v[8 * h] = bf16_lo(w0); v[8 * h + 1] = bf16_hi(w0); v[8 * h + 2] = bf16_lo(w1); v[8 * h + 3] = bf16_hi(w1);
        v[8 * h + 4] = bf16_lo(w2); v[8 * h + 5] = bf16_hi(w2); v[8 * h + 6] = bf16_lo(w3); v[8 * h + 7] = bf16_hi(w3);
      }
      fence_proxy_async_smem();                  // generic-proxy stores -> visible to the tensor core's (async proxy) operand reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(z_full(i & 1)); mbar_arrive(acc_empty(i & 1)); }
      db_acc += warp_colsum32(v, lane);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
      if (lane == 0) p.loss_partials[(size_t)(m0 / BM) * TN_EPI_WARPS + ew] = loss_acc * p.grad_scale;
    };
    auto dgrad_phase = [&](int j) {
      const int m0 = ((int)blockIdx.x + j * (int)gridDim.x) * BM, grow = m0 + tile_row;
      const bool in_range = grow < p.M;
      const uint32_t mask_word = in_range ? __ldg(p.mask_in + (size_t)cq * p.ld_mask + grow) : 0u;
      mbar_wait_long(dz_full, (uint32_t)j & 1u);
      tc_fence_after();
      uint32_t raw[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + COL_DZ + (uint32_t)c0, raw);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dz_empty);
      float v[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        v[c] = __uint_as_float(raw[c]);
        if (!(mask_word & (1u << c))) v[c] *= neg_prev;
      }
      if (in_range) store_bf16x32_global(p.dz_prev + (size_t)grow * p.ld_dz + c0, v);
    };

    for (int i = 0; i < my_blocks; ++i) {
      loss_phase(i);
      if (i > 0) dgrad_phase(i - 1);
    }
    if (my_blocks > 0) dgrad_phase(my_blocks - 1);

    // ---- the CTA's partial of dW (TMEM -> global, one row of 32 columns per thread) and of db (four lane quadrants folded in order)
    float* dw = p.dw_out + (size_t)blockIdx.x * 128 * 128 + (size_t)tile_row * 128 + c0;
    if (my_blocks > 0) {
      mbar_wait_long(dw_full, 0);                     // every MMA has completed: the dZ tiles are free (sdb aliases one of them)
      tc_fence_after();
      uint32_t raw[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + COL_DW + (uint32_t)c0, raw);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<float4*>(dw + 4 * g) = make_float4(__uint_as_float(raw[4 * g]), __uint_as_float(raw[4 * g + 1]), __uint_as_float(raw[4 * g + 2]),
                                                             __uint_as_float(raw[4 * g + 3]));
    } else {
#pragma unroll
      for (int g = 0; g < 8; ++g) *reinterpret_cast<float4*>(dw + 4 * g) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    sdb[q * 128 + c0 + lane] = db_acc;
    asm volatile("bar.sync 1, %0;" ::"n"(TN_EPI_THREADS) : "memory");
    if (q == 0) p.db_out[(size_t)blockIdx.x * 128 + c0 + lane] = ((sdb[c0 + lane] + sdb[128 + c0 + lane]) + sdb[256 + c0 + lane]) + sdb[384 + c0 + lane];
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TN_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tc
}  // namespace csb
