// simt_kernels.cuh -- fp32 CUDA-core kernels: the CSB_F32 parity mode (every GEMM in FFMA) and the HBM-bound
// element-wise / reduction kernels shared by both modes (normalise, loss, column sums, partial reduce, optimizer,
// weight repacking).  All are written for coalesced 128-bit access along the contiguous (feature) dimension.
#pragma once
#include "common.cuh"

namespace csb {
namespace simt {

// ---------------------------------------------------------------------------------------------------------------
// normalise: xn = (x - sub)/div, inf/nan -> 0   (climsim_utils/data_utils.py:806-809, 894-897)
// Writes fp32 [N, ld_f32] and/or bf16 [N, ld_bf16]; padding columns [F, ld) are zero-filled.
// ---------------------------------------------------------------------------------------------------------------
__global__ void normalize_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ sub,
                                 const float* __restrict__ div, int apply, float* __restrict__ out_f32, int ld_f32,
                                 __nv_bfloat16* __restrict__ out_bf16, int ld_bf16, int64_t N, int F, int Fp) {
  const int64_t total = N * Fp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Fp;
    const int c = (int)(i - r * Fp);
    float v = 0.f;
    if (c < F) {
      v = x[r * ld_x + c];
      if (apply) {
        v = (v - sub[c]) / div[c];
        if (isinf(v) || isnan(v)) v = 0.f;
      }
    }
    if (out_f32) out_f32[r * ld_f32 + c] = v;
    if (out_bf16) out_bf16[r * ld_bf16 + c] = __float2bfloat16_rn(v);
  }
}

// CRPS by the sorted-sample identity (climsim_utils/data_utils.py:1499-1524), averaged over time and grid:
//   out[l] = mean over (t, c) of [ mean_s |p_s - y|  -  sum_i (p_(i+1) - p_(i)) * (i + 1)(S - 1 - i) / (S (S - 1)) ]
// samples [n_tc, L, S] (S <= 32 ensemble members, contiguous), target [n_tc, L].  A warp owns one (t, c, l) element: lane s holds
// member s, the 32 lanes are sorted with a bitonic network of shuffles, the spread is one neighbour difference per lane.
// grid = (L, NB): block b of level l takes elements b*8 + warp, stepping NB*8; fp64 accumulation, fixed-order block reduction into
// partials[l * NB + b]; crps_finalize_kernel sums the NB partials of a level in order.
template <typename T>
__global__ void __launch_bounds__(256)
crps_kernel(const T* __restrict__ samples, const T* __restrict__ target, int64_t n_tc, int L, int S, double* __restrict__ partials) {
  __shared__ double red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l = blockIdx.x, NB = gridDim.y;
  const double norm = 1.0 / ((double)S * (double)(S - 1));
  double acc = 0.0;
  for (int64_t tc = (int64_t)blockIdx.y * 8 + warp; tc < n_tc; tc += (int64_t)NB * 8) {
    const bool has = lane < S;
    double v = has ? (double)samples[(tc * L + l) * S + lane] : INFINITY;      // +inf pads sort to the end
    const double y = (double)target[tc * L + l];
    double a = has ? fabs(v - y) : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, v, j);
        const bool up = (lane & k) == 0, lower = (lane & j) == 0;
        v = (lower == up) ? fmin(v, other) : fmax(v, other);
      }
    }
    const double next = __shfl_down_sync(0xffffffffu, v, 1);
    double d = (lane < S - 1) ? (next - v) * (double)((lane + 1) * (S - 1 - lane)) : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    acc += a / (double)S - d * norm;
  }
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partials[(size_t)l * NB + blockIdx.y] = t;
  }
}
__global__ void crps_finalize_kernel(const double* __restrict__ partials, int NB, int64_t n_tc, int L, double* __restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  double t = 0.0;
  for (int b = 0; b < NB; ++b) t += partials[(size_t)l * NB + b];
  out[l] = t / (double)n_tc;
}

// Row gather dst[i, :] = src[idx[i], :]: the shuffle stage of the input pipeline on the device (the reference shuffles samples on the
// host: tf.data `unbatch().shuffle(11520).batch(B)`, hpo_baseline_v1.py:140-143; DistributedSampler(shuffle=True),
// train_mlp_h5loader.py:126-134).  blockDim = (32, 8): a warp per row; VEC: rows are multiples of four floats and 16-byte aligned.
template <bool VEC>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, float* __restrict__ dst, int64_t n_rows, int row_len,
                   int64_t src_rows, int* __restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; i < n_rows; i += (int64_t)gridDim.x * blockDim.y) {
    const int64_t r = idx[i];
    if (r < 0 || r >= src_rows) {                      // never dereference an index outside the source; reported to the caller
      if (threadIdx.x == 0) atomicExch(bad, 1);
      continue;
    }
    if constexpr (VEC) {
      const float4* s = reinterpret_cast<const float4*>(src + r * row_len);
      float4* d = reinterpret_cast<float4*>(dst + i * row_len);
      for (int c = threadIdx.x; c < row_len / 4; c += 32) d[c] = __ldcs(s + c);
    } else {
      for (int c = threadIdx.x; c < row_len; c += 32) dst[i * row_len + c] = __ldcs(src + r * row_len + c);
    }
  }
}

// Generalised input prologue of the online models (online_testing/model_postprocessing/v2_nn_wrapper.ipynb cell 5 `preprocessing`,
// online_testing/baseline_models/MLP_v2rh/training/climsim_datapip_h5.py:132-168), in the reference's order:
//   x' = 1 - exp(-lambda_c x) where lambda_c != 0  ->  (x' - sub_c) / div_c  ->  nan, inf -> 0  ->  0 where keep_c == 0  ->  clamp to [lo_c, hi_c]
// xform = [4][Fp] floats: lambda, keep, lo, hi.  One thread per column (its constants live in registers), blockIdx.y = 128-column
// chunk, two rows per block pass; reads and writes are coalesced along the row.
__global__ void __launch_bounds__(256)
prepare_input_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ sub, const float* __restrict__ div,
                     const float* __restrict__ xform, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, int ld_out,
                     int64_t N, int F, int Fp) {
  const int c = blockIdx.y * 128 + (threadIdx.x & 127);
  if (c >= Fp) return;
  const bool valid = c < F;
  const float lam = valid ? xform[c] : 0.f, keep = valid ? xform[Fp + c] : 0.f, lo = xform[2 * Fp + c], hi = xform[3 * Fp + c];
  const float s = valid ? sub[c] : 0.f, d = valid ? div[c] : 1.f;
  for (int64_t r = (int64_t)blockIdx.x * 2 + (threadIdx.x >> 7); r < N; r += (int64_t)gridDim.x * 2) {
    float v = 0.f;
    if (valid && keep != 0.f) {
      v = __ldcs(x + r * ld_x + c);
      if (lam != 0.f) v = __fsub_rn(1.f, expf(-__fmul_rn(v, lam)));
      v = __fdiv_rn(__fsub_rn(v, s), d);
      if (isinf(v) || isnan(v)) v = 0.f;
      v = fminf(fmaxf(v, lo), hi);
    }
    if (out_f32) out_f32[r * ld_out + c] = v;
    if (out_bf16) out_bf16[r * ld_out + c] = __float2bfloat16_rn(v);
  }
}

// Vectorised variant for F % 4 == 0 and bf16 output: one thread per 4 columns (128-bit load, 64-bit store).  A thread keeps its
// column slot for the whole launch (the block is q = Fp / 4 slots wide, 256 / q rows tall; host guarantees 256 % q == 0), so the
// normalisation constants are loaded once and no index division runs per element; four rows are in flight per thread.
__device__ __forceinline__ float norm_elem(float x, float s, float d) {
  const float v = __fdiv_rn(__fsub_rn(x, s), d);
  return (isinf(v) || isnan(v)) ? 0.f : v;                 // (max == min) columns: inf / nan -> 0 as the reference does
}
__global__ void __launch_bounds__(256)
normalize_bf16_vec4_kernel(const float* __restrict__ x, const float* __restrict__ sub, const float* __restrict__ div, int apply,
                           __nv_bfloat16* __restrict__ out, int64_t N, int F, int Fp) {
  const int q = Fp >> 2, qv = F >> 2;                    // float4 slots per padded row / valid slots
  const int c4 = (int)threadIdx.x % q, rows_per_block = 256 / q;
  const bool valid = c4 < qv;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), d = make_float4(1.f, 1.f, 1.f, 1.f);
  if (valid && apply) { s = __ldg(reinterpret_cast<const float4*>(sub) + c4); d = __ldg(reinterpret_cast<const float4*>(div) + c4); }
  pdl_launch_dependents();
  pdl_wait();
  const int64_t step = (int64_t)gridDim.x * rows_per_block;
  for (int64_t r0 = (int64_t)blockIdx.x * rows_per_block + (int)threadIdx.x / q; r0 < N; r0 += 4 * step) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = r0 + u * step;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid && r < N) v[u] = __ldcs(reinterpret_cast<const float4*>(x + r * F) + c4);       // streamed once
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = r0 + u * step;
      if (r >= N) break;
      float4 w = v[u];
      if (apply && valid) { w.x = norm_elem(w.x, s.x, d.x); w.y = norm_elem(w.y, s.y, d.y); w.z = norm_elem(w.z, s.z, d.z); w.w = norm_elem(w.w, s.w, d.w); }
      uint2 o;
      o.x = pack_bf16x2(w.x, w.y);
      o.y = pack_bf16x2(w.z, w.w);
      *reinterpret_cast<uint2*>(out + r * Fp + 4 * c4) = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 GEMM  C[M,N] = epilogue( opA(A) . opB(B) )  -- 64x64 tile, BK 16, 256 threads, 4x4 micro-tile.
//   TA = false: A stored [M, K] (lda)        TA = true : A stored [K, M] (lda)     (weight gradient: H^T)
//   TB = false: B stored [K, N] (ldb)        TB = true : B stored [N, K] (ldb)     (data gradient:  W^T)
// N and the contiguous feature dimensions are multiples of 64 (padded layout); M (and K when it is the batch) may be
// ragged and is bounds-checked.
// ---------------------------------------------------------------------------------------------------------------
enum SEpi : int { SEPI_STORE = 0, SEPI_BIAS_ACT = 1, SEPI_DGRAD = 2, SEPI_BIAS_ADD = 3 };

struct SgemmParams {
  int M, N, K;
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;
  const float* bias;          // SEPI_BIAS_ACT
  int act; float alpha; int head_relu_from;
  const float* saved; int ld_saved;   // SEPI_DGRAD (act' of the saved activation) / SEPI_BIAS_ADD (tile to add)
  float dgrad_scale;                  // SEPI_DGRAD, != 0: a dropout layer sits behind the activation -- zero where the saved output is
                                      // exactly 0 (dropped), the kept elements' gradient times 1 / (1 - p)
  // Conv1D('same') over the halo-padded channels-last layout (see cnn_engine.cuh):
  int a_tap_k;       // !TA: contraction index k = t*a_tap_k + c reads A row (m + t - tap_center), column c.   0 = plain GEMM
  int tap_center;
  int b_tap_k;       // TB: contraction index k = t*b_tap_k + c reads B[((taps-1-t)*b_tap_rows + n) * ldb + c]  (flipped, transposed taps)
  int b_tap_rows, taps;
  int a_row_off;     // TA: A stored [K rows, M]: row k + a_row_off (rows outside [0, K) read as zero)
  int halo_period;   // rows r with r % halo_period in {0, halo_period-1} are written as zeros
};

template <bool TA, bool TB, int EPI>
__global__ void __launch_bounds__(256) sgemm_kernel(const SgemmParams p) {
  constexpr int T = 64, BKS = 16;
  __shared__ float As[BKS][T + 4];
  __shared__ float Bs[BKS][T + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * T, n0 = blockIdx.x * T;
  const int tx = tid & 15, ty = tid >> 4;   // micro-tile: rows ty*4.., cols tx*4..
  float acc[4][4] = {};

  for (int k0 = 0; k0 < p.K; k0 += BKS) {
    // ---- load A tile -> As[k][m]
    if constexpr (!TA) {
      const int r = tid >> 2, kq = (tid & 3) * 4;          // 64 rows x 4 float4
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      int arow = m0 + r, acol = k0 + kq;
      if (p.a_tap_k > 0) { const int t = acol / p.a_tap_k; acol -= t * p.a_tap_k; arow += t - p.tap_center; }
      if (m0 + r < p.M && arow >= 0 && arow < p.M) {
        const float* src = p.A + (size_t)arow * p.lda + acol;
        if (k0 + kq + 3 < p.K) v = *reinterpret_cast<const float4*>(src);
        else { float t[4] = {0, 0, 0, 0}; for (int i = 0; i < 4; ++i) if (k0 + kq + i < p.K) t[i] = src[i]; v = make_float4(t[0], t[1], t[2], t[3]); }
      }
      As[kq + 0][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
    } else {
      const int k = tid >> 4, mq = (tid & 15) * 4;         // 16 k-rows x 16 float4
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int krow = k0 + k + p.a_row_off;
      if (k0 + k < p.K && krow >= 0 && krow < p.K) {
        const float* src = p.A + (size_t)krow * p.lda + m0 + mq;
        if (m0 + mq + 3 < p.M) v = *reinterpret_cast<const float4*>(src);
        else { float t[4] = {0, 0, 0, 0}; for (int i = 0; i < 4; ++i) if (m0 + mq + i < p.M) t[i] = src[i]; v = make_float4(t[0], t[1], t[2], t[3]); }
      }
      *reinterpret_cast<float4*>(&As[k][mq]) = v;
    }
    // ---- load B tile -> Bs[k][n]
    if constexpr (!TB) {
      const int k = tid >> 4, nq = (tid & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + k < p.K) v = *reinterpret_cast<const float4*>(p.B + (size_t)(k0 + k) * p.ldb + n0 + nq);
      *reinterpret_cast<float4*>(&Bs[k][nq]) = v;
    } else {
      const int r = tid >> 2, kq = (tid & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* src = p.B + (size_t)(n0 + r) * p.ldb + k0 + kq;
      if (p.b_tap_k > 0) {
        const int t = (k0 + kq) / p.b_tap_k;
        src = p.B + ((size_t)(p.taps - 1 - t) * p.b_tap_rows + n0 + r) * p.ldb + (k0 + kq - t * p.b_tap_k);
      }
      if (k0 + kq + 3 < p.K) v = *reinterpret_cast<const float4*>(src);
      else { float t[4] = {0, 0, 0, 0}; for (int i = 0; i < 4; ++i) if (k0 + kq + i < p.K) t[i] = src[i]; v = make_float4(t[0], t[1], t[2], t[3]); }
      Bs[kq + 0][r] = v.x; Bs[kq + 1][r] = v.y; Bs[kq + 2][r] = v.z; Bs[kq + 3][r] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BKS; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= p.M) continue;
    const int c = n0 + tx * 4;
    float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
    if constexpr (EPI == SEPI_BIAS_ACT) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float z = v[j] + p.bias[c + j];
        const bool relu_col = p.head_relu_from >= 0 && (c + j) >= p.head_relu_from;
        v[j] = relu_col ? fmaxf(z, 0.f) : act_fwd(p.act, p.alpha, z);
      }
    } else if constexpr (EPI == SEPI_DGRAD) {
      const float4 a = *reinterpret_cast<const float4*>(p.saved + (size_t)r * p.ld_saved + c);
      v[0] *= act_bwd_from_out(p.act, p.alpha, a.x);
      v[1] *= act_bwd_from_out(p.act, p.alpha, a.y);
      v[2] *= act_bwd_from_out(p.act, p.alpha, a.z);
      v[3] *= act_bwd_from_out(p.act, p.alpha, a.w);
      if (p.dgrad_scale != 0.f) {
        v[0] = a.x == 0.f ? 0.f : v[0] * p.dgrad_scale; v[1] = a.y == 0.f ? 0.f : v[1] * p.dgrad_scale;
        v[2] = a.z == 0.f ? 0.f : v[2] * p.dgrad_scale; v[3] = a.w == 0.f ? 0.f : v[3] * p.dgrad_scale;
      }
    } else if constexpr (EPI == SEPI_BIAS_ADD) {
      const float4 a = *reinterpret_cast<const float4*>(p.saved + (size_t)r * p.ld_saved + c);
      v[0] += p.bias[c] + a.x; v[1] += p.bias[c + 1] + a.y; v[2] += p.bias[c + 2] + a.z; v[3] += p.bias[c + 3] + a.w;
    }
    if (p.halo_period > 0) {
      const int rr = r % p.halo_period;
      if (rr == 0 || rr == p.halo_period - 1) { v[0] = v[1] = v[2] = v[3] = 0.f; }
    }
    *reinterpret_cast<float4*>(p.C + (size_t)r * p.ldc + c) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// loss + dL/dz of the head (fp32 mode), and the dy -> dz conversion used by csb_mlp_backward.
//   mode 0: dz = loss gradient (MSE / MAE, weights w, scale) * head'(p);  accumulates the loss partial
//   mode 1: dz = dy * head'(p)                                            (external upstream gradient)
// p is the saved head output [M, ldp]; out dz [M, ldz] (fp32 or bf16); padding columns written as zero.
// ---------------------------------------------------------------------------------------------------------------
template <typename TZ>
__device__ __forceinline__ void store_dz(TZ* dst, float v);
template <> __device__ __forceinline__ void store_dz<float>(float* dst, float v) { *dst = v; }
template <> __device__ __forceinline__ void store_dz<__nv_bfloat16>(__nv_bfloat16* dst, float v) { *dst = __float2bfloat16_rn(v); }

template <typename TZ>
__global__ void __launch_bounds__(256)
head_grad_kernel(const float* __restrict__ pred, int ldp, const float* __restrict__ y_or_dy, int ldy,
                 const float* __restrict__ w, float scale, int loss_kind, int mode, int act, float alpha,
                 int head_relu_from, const float* __restrict__ out_mask, TZ* __restrict__ dz, int ldz,
                 int64_t M, int out_dim, int Np, float* __restrict__ loss_partials, int halo_period = 0) {
  float lacc = 0.f;
  const int64_t total = M * Np;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Np;
    const int c = (int)(i - r * Np);
    float g = 0.f;
    bool interior = true;
    int64_t yr = r;                                         // row of the (halo-free) target / upstream-gradient array
    if (halo_period > 0) {
      const int rr = (int)(r % halo_period);
      interior = rr != 0 && rr != halo_period - 1;
      yr = (r / halo_period) * (halo_period - 2) + rr - 1;
    }
    if (c < out_dim && interior) {
      const float pv = pred[r * ldp + c];                 // already masked by the forward pass
      const bool relu_col = head_relu_from >= 0 && c >= head_relu_from;
      float dact = relu_col ? (pv > 0.f ? 1.f : 0.f) : act_bwd_from_out(act, alpha, pv);
      if (out_mask != nullptr) dact *= out_mask[c];
      if (mode == 0) {
        const float d = pv - y_or_dy[yr * ldy + c];
        if (loss_kind == CSB_LOSS_MSE) { lacc += w[c] * d * d; g = 2.f * w[c] * d * scale * dact; }
        else if (loss_kind == CSB_LOSS_HUBER) {
          const float ad = fabsf(d);
          lacc += w[c] * (ad <= 1.f ? 0.5f * d * d : ad - 0.5f);
          g = w[c] * scale * (ad <= 1.f ? d : (d > 0.f ? 1.f : -1.f)) * dact;
        }
        else { lacc += w[c] * fabsf(d); g = w[c] * scale * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * dact; }
      } else {
        g = y_or_dy[yr * ldy + c] * dact;
      }
    }
    store_dz<TZ>(dz + r * ldz + c, g);
  }
  if (loss_partials != nullptr) {
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lacc += __shfl_xor_sync(0xffffffffu, lacc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lacc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
      loss_partials[blockIdx.x] = s * scale;
    }
  }
}

// final deterministic reduction of the loss partials (single block, fp64 accumulation)
__device__ __forceinline__ void block_sum_loss(const float* __restrict__ partials, int n, float* __restrict__ out);
__global__ void __launch_bounds__(256) loss_finalize_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
  block_sum_loss(partials, n, out);
}

// ---------------------------------------------------------------------------------------------------------------
// bias gradient: column sums of dZ [M, ld] over rows, two-stage deterministic.
// grid = (Np/64, S); block = 256 = 64 columns x 4 row lanes; partial [S][Np] written to `out + s*stride`.
// ---------------------------------------------------------------------------------------------------------------
template <typename TZ>
__device__ __forceinline__ float load_as_float(const TZ* p);
template <> __device__ __forceinline__ float load_as_float<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename TZ>
__global__ void __launch_bounds__(256)
colsum_kernel(const TZ* __restrict__ dz, int ld, int64_t M, float* __restrict__ out, size_t split_stride) {
  __shared__ float red[4][64];
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int rl = threadIdx.x >> 6;
  const int S = gridDim.y;
  const int64_t rows_per = (M + S - 1) / S;
  const int64_t r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float s = 0.f;
  for (int64_t r = r0 + rl; r < r1; r += 4) s += load_as_float<TZ>(dz + r * ld + c);
  red[rl][threadIdx.x & 63] = s;
  __syncthreads();
  if (rl == 0) out[(size_t)blockIdx.y * split_stride + c] = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm (eps 1e-5, affine, biased variance -- torch.nn.LayerNorm, baseline_models/HSR/training/hsr.py:23) between
// a Linear and its activation:  u = (z - mean)/sqrt(var + eps) * gamma + beta,  a = act(u).   One warp per row.
// ---------------------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const T* __restrict__ z, T* __restrict__ a, int ld, const float* __restrict__ gamma, const float* __restrict__ beta,
              float* __restrict__ stats, int64_t M, int N, int Np, int act, float alpha, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < M; r += warps) {
    const T* zr = z + r * ld;
    float s = 0.f;
    for (int c = lane; c < N; c += 32) s += to_f32<T>(zr[c]);
    const float mean = warp_sum(s) / (float)N;
    float q = 0.f;
    for (int c = lane; c < N; c += 32) { const float d = to_f32<T>(zr[c]) - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)N + eps);
    if (lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
    T* ar = a + r * ld;
    for (int c = lane; c < Np; c += 32) {
      float v = 0.f;
      if (c < N) v = act_fwd(act, alpha, (to_f32<T>(zr[c]) - mean) * rstd * gamma[c] + beta[c]);
      ar[c] = from_f32<T>(v);
    }
  }
}

// du -> dz in place:  g = du*gamma;  dz = rstd * (g - mean(g) - xhat * mean(g*xhat))
template <typename T>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(T* __restrict__ du_dz, const T* __restrict__ z, int ld, const float* __restrict__ gamma, const float* __restrict__ stats,
              int64_t M, int N) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < M; r += warps) {
    const float mean = stats[2 * r], rstd = stats[2 * r + 1];
    T* dr = du_dz + r * ld;
    const T* zr = z + r * ld;
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < N; c += 32) {
      const float g = to_f32<T>(dr[c]) * gamma[c];
      s1 += g;
      s2 += g * (to_f32<T>(zr[c]) - mean) * rstd;
    }
    s1 = warp_sum(s1) / (float)N;
    s2 = warp_sum(s2) / (float)N;
    for (int c = lane; c < N; c += 32) {
      const float xh = (to_f32<T>(zr[c]) - mean) * rstd;
      dr[c] = from_f32<T>(rstd * (to_f32<T>(dr[c]) * gamma[c] - s1 - xh * s2));
    }
  }
}

// ---- bf16 LayerNorm, rows of at most 256 * MAXP columns: a warp owns a row and keeps it in registers (lane -> eight consecutive
// columns per 256-column pass, 128-bit loads and stores), so z is read once; the backward kernel fuses the parameter gradients
// (per-lane column sums carried across the warp's rows, reduced over the block's warps through shared memory in a fixed order)
// with du -> dz.  Statistics run over the N valid columns; padding columns [N, Np) are written as zeros.
constexpr int LN_MAXP = 4;       // up to 1024 columns (the HSR networks: 1024)

__device__ __forceinline__ void ln_load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 w = *reinterpret_cast<const uint4*>(p);
  v[0] = bf16_lo(w.x); v[1] = bf16_hi(w.x); v[2] = bf16_lo(w.y); v[3] = bf16_hi(w.y);
  v[4] = bf16_lo(w.z); v[5] = bf16_hi(w.z); v[6] = bf16_lo(w.w); v[7] = bf16_hi(w.w);
}
__device__ __forceinline__ void ln_store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 w;
  w.x = pack_bf16x2(v[0], v[1]); w.y = pack_bf16x2(v[2], v[3]); w.z = pack_bf16x2(v[4], v[5]); w.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = w;
}
__device__ __forceinline__ void ln_load8f(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__global__ void __launch_bounds__(256, 3)
ln_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ a, int ld, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ stats, int64_t M, int N, int Np, int act, float alpha, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float inv_n = 1.f / (float)N;
  for (int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < M; r += warps) {
    float v[LN_MAXP][8];
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < LN_MAXP; ++p) {
      const int c = p * 256 + lane * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[p][j] = 0.f;
      if (c < Np) ln_load8(z + r * ld + c, v[p]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += (c + j < N) ? v[p][j] : 0.f;
    }
    const float mean = warp_sum(s) * inv_n;
    float q = 0.f;
#pragma unroll
    for (int p = 0; p < LN_MAXP; ++p) {
      const int c = p * 256 + lane * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[p][j] - mean; q += (c + j < N) ? d * d : 0.f; }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_n + eps);
    if (lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
#pragma unroll
    for (int p = 0; p < LN_MAXP; ++p) {
      const int c = p * 256 + lane * 8;
      if (c < Np) {
        float g[8], b[8], o[8];
        ln_load8f(gamma + c, g);
        ln_load8f(beta + c, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (c + j < N) ? act_fwd(act, alpha, (v[p][j] - mean) * rstd * g[j] + b[j]) : 0.f;
        ln_store8(a + r * ld + c, o);
      }
    }
  }
}

// du -> dz in place (g = du * gamma; dz = rstd * (g - mean(g) - xhat * mean(g * xhat))) fused with the parameter gradients
// dgamma = sum_rows du * xhat, dbeta = sum_rows du: one partial pair per block at partials + blockIdx.x * 2 * Np.
// The running column sums live in shared memory, one private slice per warp in a lane-major layout (element j of pass p of lane l at
// (p*8 + j)*32 + l: conflict-free scalar accesses, no atomics), as does gamma: that keeps the kernel at two blocks per SM, which a
// streaming kernel with one row in flight per warp needs to cover the HBM latency (the first version carried the sums in registers:
// 200 registers, one block per SM, 2.2 TB/s).
__global__ void __launch_bounds__(256, 2)
ln_bwd_bf16_kernel(__nv_bfloat16* __restrict__ du_dz, const __nv_bfloat16* __restrict__ z, int ld, const float* __restrict__ gamma,
                   const float* __restrict__ stats, int64_t M, int N, int Np, float* __restrict__ partials) {
  extern __shared__ float sm[];                      // gamma [LN_MAXP*256] | per warp: dgamma [LN_MAXP*256], dbeta [LN_MAXP*256]
  constexpr int W = LN_MAXP * 256;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* gm = sm;
  float* ag = sm + W + warp * 2 * W;
  float* ab = ag + W;
  for (int i = threadIdx.x; i < W; i += 256) {       // lane-major copy of gamma: slot (p*8 + j)*32 + l <- column p*256 + l*8 + j
    const int l = i & 31, pj = i >> 5, c = (pj >> 3) * 256 + l * 8 + (pj & 7);
    gm[i] = c < N ? gamma[c] : 0.f;
  }
  for (int i = lane; i < 2 * W; i += 32) ag[i] = 0.f;
  __syncthreads();
  const int64_t warps = (int64_t)gridDim.x * 8;
  const float inv_n = 1.f / (float)N;
  for (int64_t r = blockIdx.x * (int64_t)8 + warp; r < M; r += warps) {
    const float mean = stats[2 * r], rstd = stats[2 * r + 1];
    float du[LN_MAXP][8], xh[LN_MAXP][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int p = 0; p < LN_MAXP; ++p) {
      const int c = p * 256 + lane * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) { du[p][j] = 0.f; xh[p][j] = 0.f; }
      if (c < Np) { ln_load8(du_dz + r * ld + c, du[p]); ln_load8(z + r * ld + c, xh[p]); }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = c + j < N;
        const int slot = (p * 8 + j) * 32 + lane;
        xh[p][j] = ok ? (xh[p][j] - mean) * rstd : 0.f;
        du[p][j] = ok ? du[p][j] : 0.f;
        const float g = du[p][j] * gm[slot];
        s1 += g;
        s2 += g * xh[p][j];
        ag[slot] += du[p][j] * xh[p][j];
        ab[slot] += du[p][j];
        du[p][j] = g;                                // from here on: du holds g = du * gamma
      }
    }
    s1 = warp_sum(s1) * inv_n;
    s2 = warp_sum(s2) * inv_n;
#pragma unroll
    for (int p = 0; p < LN_MAXP; ++p) {
      const int c = p * 256 + lane * 8;
      if (c < Np) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (c + j < N) ? rstd * (du[p][j] - s1 - xh[p][j] * s2) : 0.f;
        ln_store8(du_dz + r * ld + c, o);
      }
    }
  }
  // block reduction over the eight warps, warp 0 first (fixed order); slot -> column as above
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * W; i += 256) {
    const int which = i / W, s_ = i - which * W;
    const int l = s_ & 31, pj = s_ >> 5, c = (pj >> 3) * 256 + l * 8 + (pj & 7);
    if (c < Np) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += sm[W + w * 2 * W + i];
      partials[(size_t)blockIdx.x * 2 * Np + which * Np + c] = t;
    }
  }
}

// dgamma = sum_rows du*xhat, dbeta = sum_rows du: two-stage column sums; partials [S][2][Np] at out + s*stride
template <typename T>
__global__ void __launch_bounds__(256)
ln_param_grad_kernel(const T* __restrict__ du, const T* __restrict__ z, int ld, const float* __restrict__ stats, int64_t M,
                     float* __restrict__ out, size_t split_stride, int Np) {
  __shared__ float red[2][4][64];
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int rl = threadIdx.x >> 6;
  const int S = gridDim.y;
  const int64_t rows_per = (M + S - 1) / S;
  const int64_t r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float sg = 0.f, sb = 0.f;
  for (int64_t r = r0 + rl; r < r1; r += 4) {
    const float d = to_f32<T>(du[r * ld + c]);
    sg += d * (to_f32<T>(z[r * ld + c]) - stats[2 * r]) * stats[2 * r + 1];
    sb += d;
  }
  red[0][rl][threadIdx.x & 63] = sg;
  red[1][rl][threadIdx.x & 63] = sb;
  __syncthreads();
  if (rl == 0) {
    const int t = threadIdx.x;
    float* o = out + (size_t)blockIdx.y * split_stride;
    o[c] = red[0][0][t] + red[0][1][t] + red[0][2][t] + red[0][3][t];
    o[Np + c] = red[1][0][t] + red[1][1][t] + red[1][2][t] + red[1][3][t];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// reduce split partials in a fixed order: grad[i] = sum_s ws[s*stride + i]
// ---------------------------------------------------------------------------------------------------------------
struct Segment { const float* ws; size_t stride; float* grad; int64_t len; int splits; };

// sum_s p[s * stride] over four consecutive floats, splits taken in order 0, 1, 2, ... (the fixed order every gradient in the
// engine is reduced in); loads are issued eight (then four) splits at a time so that several are in flight per thread
__device__ __forceinline__ float4 sum_partials4(const float* __restrict__ p, size_t stride, int splits) {
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  int s = 0;
  // (the additions stay in split order: only the loads are batched.  Sixteen in flight first: the output layer of the fused tail has
  // one partial per CTA -- 148 -- and this walk is a pure latency chain, ~1.2 us per batch whatever its size)
  for (; s + 16 <= splits; s += 16) {
    float4 v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(p + (size_t)(s + u) * stride));
#pragma unroll
    for (int u = 0; u < 16; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
  }
  for (; s + 8 <= splits; s += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(p + (size_t)(s + u) * stride));
#pragma unroll
    for (int u = 0; u < 8; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
  }
  for (; s + 4 <= splits; s += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(p + (size_t)(s + u) * stride));
#pragma unroll
    for (int u = 0; u < 4; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
  }
  for (; s < splits; ++s) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(p + (size_t)s * stride));
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  return a;
}
struct SegmentTable {
  int n; Segment seg[4 * CSB_MAX_LAYERS];
  // optional: the scalar loss rides in the same launch (row blockIdx.y == n, one block), off the backward pass's critical path
  const float* loss_partials; int n_loss; float* loss_out;
};

// fixed-order fp64 sum of the loss partials by one block (deterministic)
__device__ __forceinline__ void block_sum_loss(const float* __restrict__ partials, int n, float* __restrict__ out) {
  __shared__ double red[256];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int i = threadIdx.x;
  for (; i + 3 * 256 < n; i += 4 * 256) {          // four independent loads in flight per thread
    s0 += (double)partials[i]; s1 += (double)partials[i + 256]; s2 += (double)partials[i + 512]; s3 += (double)partials[i + 768];
  }
  for (; i < n; i += 256) s0 += (double)partials[i];
  red[threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)red[0];
}

// one launch for every gradient segment: blockIdx.y = segment
__global__ void __launch_bounds__(256) reduce_partials_kernel(const SegmentTable tab) {
  pdl_launch_dependents();
  pdl_wait();
  if ((int)blockIdx.y == tab.n) {
    if (blockIdx.x == 0) block_sum_loss(tab.loss_partials, tab.n_loss, tab.loss_out);
    return;
  }
  const Segment sg = tab.seg[blockIdx.y];
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4; i < sg.len; i += (int64_t)gridDim.x * blockDim.x * 4) {
    *reinterpret_cast<float4*>(sg.grad + i) = sum_partials4(sg.ws + i, sg.stride, sg.splits);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// optimizer: flat element-wise update over the padded parameter buffer (padding has zero gradient -> stays zero)
// ---------------------------------------------------------------------------------------------------------------
struct OptParams {
  int rule;
  float lr, beta1, beta2, eps, wd;
  float bc1, bc2;          // 1 - beta^t
  float radam_r;           // RADAM: rectification factor r_t, or < 0 while sma_t < 5 (un-rectified momentum step)
};

// one parameter element; the arithmetic of each rule is written exactly once and shared by the flat and the fused kernels.
// Every operation is an explicit round-to-nearest intrinsic so that the compiler cannot contract multiply-adds differently in
// the two kernels (their results are compared bit for bit).
__device__ __forceinline__ void opt_update(const OptParams& o, float& w, const float g, float& m, float& v) {
  const float one_m_b1 = __fsub_rn(1.f, o.beta1), one_m_b2 = __fsub_rn(1.f, o.beta2);
  if (o.rule == CSB_OPT_SGD) {
    w = __fsub_rn(w, __fmul_rn(o.lr, __fmaf_rn(o.wd, w, g)));
  } else if (o.rule == CSB_OPT_ADAM_KERAS) {
    // keras Adam.update_step: alpha = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g^2-v)(1-b2); w -= alpha*m/(sqrt(v)+eps)
    const float alpha = __fdiv_rn(__fmul_rn(o.lr, __fsqrt_rn(o.bc2)), o.bc1);
    m = __fmaf_rn(__fsub_rn(g, m), one_m_b1, m);
    v = __fmaf_rn(__fmaf_rn(g, g, -v), one_m_b2, v);
    w = __fsub_rn(w, __fdiv_rn(__fmul_rn(alpha, m), __fadd_rn(__fsqrt_rn(v), o.eps)));
  } else if (o.rule == CSB_OPT_RADAM) {
    // tfa RectifiedAdam._resource_apply_dense (no warm-up, no amsgrad)
    m = __fmaf_rn(o.beta1, m, __fmul_rn(one_m_b1, g));
    v = __fmaf_rn(o.beta2, v, __fmul_rn(__fmul_rn(one_m_b2, g), g));
    const float mhat = __fdiv_rn(m, o.bc1);
    const float step = (o.radam_r >= 0.f) ? __fdiv_rn(__fmul_rn(o.radam_r, mhat), __fadd_rn(__fsqrt_rn(__fdiv_rn(v, o.bc2)), o.eps)) : mhat;
    w = __fsub_rn(w, __fmul_rn(o.lr, __fmaf_rn(o.wd, w, step)));
  } else if (o.rule == CSB_OPT_RMSPROP) {
    // keras RMSprop.update_step (rho in beta2; epsilon inside the square root)
    v = __fmaf_rn(o.beta2, v, __fmul_rn(__fmul_rn(one_m_b2, g), g));
    w = __fsub_rn(w, __fdiv_rn(__fmul_rn(o.lr, g), __fsqrt_rn(__fadd_rn(v, o.eps))));
  } else {
    // torch.optim.Adam (L2 decay folded into g)
    const float ge = __fmaf_rn(o.wd, w, g);
    m = __fmaf_rn(o.beta1, m, __fmul_rn(one_m_b1, ge));
    v = __fmaf_rn(o.beta2, v, __fmul_rn(__fmul_rn(one_m_b2, ge), ge));
    w = __fsub_rn(w, __fdiv_rn(__fmul_rn(__fdiv_rn(o.lr, o.bc1), m), __fadd_rn(__fdiv_rn(__fsqrt_rn(v), __fsqrt_rn(o.bc2)), o.eps)));
  }
}

__global__ void __launch_bounds__(256)
opt_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
           const OptParams o) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4; i < n; i += (int64_t)gridDim.x * blockDim.x * 4) {
    const float4 w4 = *reinterpret_cast<float4*>(w + i), g4 = *reinterpret_cast<const float4*>(g + i);
    const float4 m4 = *reinterpret_cast<float4*>(m + i), v4 = *reinterpret_cast<float4*>(v + i);
    float wv[4] = {w4.x, w4.y, w4.z, w4.w}, mv[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
    const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) opt_update(o, wv[j], gv[j], mv[j], vv[j]);
    if (o.rule != CSB_OPT_SGD) {
      *reinterpret_cast<float4*>(m + i) = make_float4(mv[0], mv[1], mv[2], mv[3]);
      *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
    *reinterpret_cast<float4*>(w + i) = make_float4(wv[0], wv[1], wv[2], wv[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fused "reduce split partials -> optimizer rule -> bf16 weight copies" (CSB_TRAIN_FUSED_OPT): one launch replaces
// reduce_partials_kernel + opt_kernel + repack_kernel.  blockIdx.y = layer (row n: the scalar loss), blockIdx.x walks the
// 32 x 32 tiles of W_l [Kp, Np] and then 256-wide pieces of b_l.  The partials are summed in the same fixed order as
// reduce_partials_kernel does, so gradients (also written to the gradient buffer) and weights are bit-identical.
// ---------------------------------------------------------------------------------------------------------------
struct FusedOptLayer {
  int Kp, Np;
  size_t w_off, b_off;                  // offsets into params / grads / m / v
  __nv_bfloat16 *w16, *wt16;
  const float* ws_w; int w_splits;      // partials of dW: ws_w + s * Kp * Np
  const float* ws_b; int b_splits;      // partials of db: ws_b + s * Np
  // LayerNorm layers: gamma [Np] then beta [Np] at g_off, partials ws_g + s * 2 * Np (g_splits == 0: no LayerNorm)
  size_t g_off; const float* ws_g; int g_splits;
  // rows of W per work item: 32, or 8 for a layer with many partials (the fused tail writes one per CTA): a block that walks 148
  // partials of a 32 x 64 tile pulls 1.2 MB through one SM, which alone takes longer than the rest of the launch
  int tk;
};
static inline int fused_opt_tk(int w_splits) { return w_splits > 32 ? 8 : 32; }
struct FusedOptTable {
  int n; FusedOptLayer l[CSB_MAX_LAYERS];
  float *params, *grads, *m, *v;
  const float* loss_partials; int n_loss; float* loss_out;
};

// work items of one layer in the fused optimizer: tiles of W, then 256-wide pieces of the vector segments
__device__ __forceinline__ int fused_opt_items(const FusedOptLayer& L) {
  return (L.Kp / L.tk) * (L.Np / 64) + (L.Np / 4 + 255) / 256 + (L.g_splits > 0 ? (2 * L.Np / 4 + 255) / 256 : 0);
}
// one work item (called by all 256 threads of a block; `t` is the block's transpose staging tile)
__device__ __forceinline__ void fused_opt_item(const FusedOptTable& tab, const OptParams& o, const FusedOptLayer& L, int item, float (*t)[65]) {
  // W_l in tiles of TK (k) x 64 (n): thread -> 4 consecutive n (one float4) of rows ty, ty + 16
  const int TK = L.tk;
  const int tiles_n = L.Np / 64, tiles = (L.Kp / TK) * tiles_n;
  const int vec_b = (L.Np / 4 + 255) / 256;
  const size_t wsz = (size_t)L.Kp * L.Np;
  if (item < tiles) {
    const int k0 = (item / tiles_n) * TK, n0 = (item % tiles_n) * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int k = ty; k < TK; k += 16) {
      const size_t idx = (size_t)(k0 + k) * L.Np + n0 + 4 * tx, e = L.w_off + idx;
      const float4 g4 = sum_partials4(L.ws_w + idx, wsz, L.w_splits);
      const float4 w4 = *reinterpret_cast<const float4*>(tab.params + e), m4 = *reinterpret_cast<const float4*>(tab.m + e),
                   v4 = *reinterpret_cast<const float4*>(tab.v + e);
      float w[4] = {w4.x, w4.y, w4.z, w4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
      const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { opt_update(o, w[j], g[j], m[j], v[j]); t[k][4 * tx + j] = w[j]; }
      *reinterpret_cast<float4*>(tab.grads + e) = g4;
      *reinterpret_cast<float4*>(tab.params + e) = make_float4(w[0], w[1], w[2], w[3]);
      if (o.rule != CSB_OPT_SGD) {
        *reinterpret_cast<float4*>(tab.m + e) = make_float4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<float4*>(tab.v + e) = make_float4(v[0], v[1], v[2], v[3]);
      }
      *reinterpret_cast<uint2*>(L.w16 + idx) = make_uint2(pack_bf16x2(w[0], w[1]), pack_bf16x2(w[2], w[3]));
    }
    __syncthreads();
    {  // transposed copy: 64 n-rows x 32 k; thread -> 8 consecutive k of one n (one 16-byte store)
      const int n = threadIdx.x >> 2, kq = (threadIdx.x & 3) * 8;
      if (kq < TK) {
        uint4 u;
        u.x = pack_bf16x2(t[kq + 0][n], t[kq + 1][n]); u.y = pack_bf16x2(t[kq + 2][n], t[kq + 3][n]);
        u.z = pack_bf16x2(t[kq + 4][n], t[kq + 5][n]); u.w = pack_bf16x2(t[kq + 6][n], t[kq + 7][n]);
        *reinterpret_cast<uint4*>(L.wt16 + (size_t)(n0 + n) * L.Kp + k0 + kq) = u;
      }
    }
    __syncthreads();
  } else {
    // vector segments: the bias [Np] and, behind it, a LayerNorm layer's (gamma, beta) [2 Np]
    int c = ((item - tiles) * 256 + threadIdx.x) * 4;
    size_t e; const float* wsp; size_t stride; int splits; bool ok;
    if (c < L.Np) { e = L.b_off + c; wsp = L.ws_b + c; stride = (size_t)L.Np; splits = L.b_splits; ok = true; }
    else { c -= vec_b * 1024; e = L.g_off + c; wsp = L.ws_g + c; stride = (size_t)2 * L.Np; splits = L.g_splits; ok = c >= 0 && c < 2 * L.Np && L.g_splits > 0; }
    if (ok) {
      const float4 g4 = sum_partials4(wsp, stride, splits);
      const float4 w4 = *reinterpret_cast<const float4*>(tab.params + e), m4 = *reinterpret_cast<const float4*>(tab.m + e),
                   v4 = *reinterpret_cast<const float4*>(tab.v + e);
      float w[4] = {w4.x, w4.y, w4.z, w4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
      const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) opt_update(o, w[j], g[j], m[j], v[j]);
      *reinterpret_cast<float4*>(tab.grads + e) = g4;
      *reinterpret_cast<float4*>(tab.params + e) = make_float4(w[0], w[1], w[2], w[3]);
      if (o.rule != CSB_OPT_SGD) {
        *reinterpret_cast<float4*>(tab.m + e) = make_float4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<float4*>(tab.v + e) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
  }
}

__global__ void __launch_bounds__(256) opt_fused_kernel(const FusedOptTable tab, const OptParams o) {
  pdl_launch_dependents();
  pdl_wait();
  if ((int)blockIdx.y == tab.n) {
    if (blockIdx.x == 0 && tab.loss_out != nullptr) block_sum_loss(tab.loss_partials, tab.n_loss, tab.loss_out);
    return;
  }
  const FusedOptLayer L = tab.l[blockIdx.y];
  __shared__ float t[32][65];
  const int items = fused_opt_items(L);
  for (int item = blockIdx.x; item < items; item += gridDim.x) fused_opt_item(tab, o, L, item, t);
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 weight repack: from the fp32 master W [Kp, Np] write W16 [Kp, Np] and Wt16 [Np, Kp] (32x32 smem transpose)
// one launch for all layers: blockIdx.y = layer, blockIdx.x = 32x32 tile index
// ---------------------------------------------------------------------------------------------------------------
struct RepackLayer { const float* w; __nv_bfloat16* w16; __nv_bfloat16* wt16; int Kp, Np; };
struct RepackTable { int n; RepackLayer l[CSB_MAX_LAYERS]; };

__global__ void __launch_bounds__(256) repack_kernel(const RepackTable tab) {
  const RepackLayer L = tab.l[blockIdx.y];
  const int tiles_n = L.Np / 32, tiles = (L.Kp / 32) * tiles_n;
  __shared__ float t[32][33];
  pdl_launch_dependents();
  pdl_wait();
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int k0 = (tile / tiles_n) * 32, n0 = (tile % tiles_n) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + ty + 8 * i;
      const float v = L.w[(size_t)k * L.Np + n0 + tx];
      t[ty + 8 * i][tx] = v;
      L.w16[(size_t)k * L.Np + n0 + tx] = __float2bfloat16_rn(v);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + ty + 8 * i;
      L.wt16[(size_t)n * L.Kp + k0 + tx] = __float2bfloat16_rn(t[tx][ty + 8 * i]);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// flat user blob (unpadded, W_l [K x N], b_l [N]) <-> padded internal buffer (W_l [Kp x Np], b_l [Np])
// dir 0: user -> padded (padding entries untouched = zero);  dir 1: padded -> user.   blockIdx.y = layer
// ---------------------------------------------------------------------------------------------------------------
struct PadLayer { int K, N, Np, ln; size_t w_off, b_off, g_off, w_off_user, b_off_user, g_off_user; };   // gamma at g_off, beta at g_off + Np
struct PadTable { int n; PadLayer l[CSB_MAX_LAYERS]; };

__global__ void __launch_bounds__(256) pad_copy_kernel(float* __restrict__ padded, float* __restrict__ user, int dir, const PadTable tab) {
  const PadLayer L = tab.l[blockIdx.y];
  const int64_t total = (int64_t)(L.K + 1 + 2 * L.ln) * L.N;  // K weight rows + the bias row (+ gamma, beta rows)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / L.N), c = (int)(i - (int64_t)r * L.N);
    float *pp, *pu;
    if (r < L.K) { pp = padded + L.w_off + (size_t)r * L.Np + c; pu = user + L.w_off_user + (size_t)r * L.N + c; }
    else if (r == L.K) { pp = padded + L.b_off + c; pu = user + L.b_off_user + c; }
    else { pp = padded + L.g_off + (size_t)(r - L.K - 1) * L.Np + c; pu = user + L.g_off_user + (size_t)(r - L.K - 1) * L.N + c; }
    if (dir == 0) *pp = *pu; else *pu = *pp;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
// in-place column mask: a[r, c] *= mask[c]
__global__ void colmask_kernel(float* __restrict__ a, int ld, const float* __restrict__ mask, int64_t M, int F) {
  const int64_t total = M * F;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / F;
    const int c = (int)(i - r * F);
    a[r * ld + c] *= mask[c];
  }
}

// out[r, c] = in[r, c] * scale[c] for c < F  (denormalise predictions), in ld_in -> out ld_out
__global__ void scale_copy_kernel(const float* __restrict__ in, int ld_in, const float* __restrict__ scale,
                                  float* __restrict__ out, int ld_out, int64_t M, int F) {
  const int64_t total = M * F;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / F;
    const int c = (int)(i - r * F);
    const float v = in[r * ld_in + c];
    out[r * ld_out + c] = scale ? v * scale[c] : v;
  }
}

// CNN layout helpers (climsim_utils/data_utils.py:1693-1760)
__global__ void cnn_reshape_in_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t N, int nprof, int nscal, int ld) {
  // (N, 60*nprof + nscal) -> (N, 60, nprof + nscal)
  const int C = nprof + nscal;
  const int64_t total = N * 60 * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int l = (int)((i / C) % 60);
    const int64_t n = i / (60 * C);
    out[i] = (c < nprof) ? x[n * ld + c * 60 + l] : x[n * ld + nprof * 60 + (c - nprof)];
  }
}
__global__ void cnn_reshape_out_kernel(const float* __restrict__ p, float* __restrict__ out, int64_t N) {
  // (N, 60, 10) -> (N, 128): channels 0,1 as profiles; channels 2..9 = mean over the 60 levels
  const int64_t total = N * 128;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i & 127);
    const int64_t n = i >> 7;
    const float* src = p + n * 600;
    if (c < 120) {
      out[i] = src[(c % 60) * 10 + (c / 60)];
    } else {
      float s = 0.f;
      for (int l = 0; l < 60; ++l) s += src[l * 10 + (c - 120 + 2)];
      out[i] = s / 60.f;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CNN helpers (channels-last, halo-padded rows: sample b occupies rows b*(L+2) .. b*(L+2)+L+1, first and last are zero)
// ---------------------------------------------------------------------------------------------------------------
// x (B, L, C) fp32 -> bf16 [B*(L+2), Cp]
__global__ void __launch_bounds__(256)
cnn_pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t B, int L, int C, int Cp) {
  const int64_t total = B * (L + 2) * Cp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const int64_t r = i / Cp;
    const int l = (int)(r % (L + 2)) - 1;
    const int64_t b = r / (L + 2);
    float v = 0.f;
    if (c < C && l >= 0 && l < L) v = x[(b * L + l) * C + c];
    out[i] = __float2bfloat16_rn(v);
  }
}
// fp32 variants for the CSB_F32 parity mode
__global__ void __launch_bounds__(256)
cnn_pack_input_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t B, int L, int C, int Cp) {
  const int64_t total = B * (L + 2) * Cp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cp);
    const int64_t r = i / Cp;
    const int l = (int)(r % (L + 2)) - 1;
    const int64_t b = r / (L + 2);
    out[i] = (c < C && l >= 0 && l < L) ? x[(b * L + l) * C + c] : 0.f;
  }
}
__global__ void __launch_bounds__(256)
act_mask_f32_kernel(const float* __restrict__ g, const float* __restrict__ a, float* __restrict__ dz, int64_t n, int act, float alpha) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dz[i] = g[i] * act_bwd_from_out(act, alpha, a[i]);
}
// pred [B*(L+2), ld] (halo layout) -> y (B, L, C)
__global__ void __launch_bounds__(256)
cnn_unpack_output_kernel(const float* __restrict__ pred, int ld, float* __restrict__ y, int64_t B, int L, int C) {
  const int64_t total = B * L * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int l = (int)((i / C) % L);
    const int64_t b = i / ((int64_t)C * L);
    y[i] = pred[(b * (L + 2) + l + 1) * ld + c];
  }
}

// Inverted dropout in place (Keras Dropout(rate) in training mode, CNN/training/hpo_train.py:170,177): an element is kept with
// probability 1 - rate and scaled by 1 / (1 - rate), else zeroed.  The keep decisions come from a counter-based generator: one
// 32-bit mix of (seed, index of the thread's eight elements) starts an LCG whose high 24 bits decide the eight elements, so a
// (seed, position) pair always gives the same mask (reproducible steps; TensorFlow's own random stream cannot be matched).
// The backward pass needs no stored mask: a dropped element is zero, and relu'(0) = 0 already blocks its gradient.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__global__ void __launch_bounds__(256)
dropout_bf16_kernel(__nv_bfloat16* __restrict__ a, int64_t n8, uint32_t seed, uint32_t keep_threshold, float scale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 av = reinterpret_cast<const uint4*>(a)[i];
    const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
    uint32_t st = mix32(seed ^ mix32((uint32_t)i) ^ (uint32_t)(i >> 32) * 0x9E3779B1u);
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      st = st * 747796405u + 2891336453u;
      const float lo = (st >> 8) >= keep_threshold ? bf16_lo(aw[j]) * scale : 0.f;
      st = st * 747796405u + 2891336453u;
      const float hi = (st >> 8) >= keep_threshold ? bf16_hi(aw[j]) * scale : 0.f;
      o[j] = pack_bf16x2(lo, hi);
    }
    reinterpret_cast<uint4*>(a)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// dst[c][r] = src[r][c] for r < rows, c < cols (fp32; 32 x 32 tiles through shared memory, both sides coalesced): the TF32 mode's weight
// gradient contracts over the batch, which the K-major tcgen05 operands want contiguous
// round_tf32 != 0: values rounded to nearest onto the TF32 grid on the way (tc::round_tf32); ld_dst == 0: no transposition (a rounded copy).
__global__ void __launch_bounds__(256) transpose_f32_kernel(const float* __restrict__ src, int64_t ld_src, int64_t rows, int cols,
                                                            float* __restrict__ dst, int64_t ld_dst, int round_tf32) {
  __shared__ float t[32][33];
  const int64_t tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
  for (int64_t tile = blockIdx.x; tile < tiles_c * tiles_r; tile += gridDim.x) {
    const int64_t r0 = (tile / tiles_c) * 32;
    const int c0 = (int)(tile % tiles_c) * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t r = r0 + ty + 8 * j;
      float v = (r < rows && c0 + tx < cols) ? src[r * ld_src + c0 + tx] : 0.f;
      if (round_tf32) v = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
      t[ty + 8 * j][tx] = v;
    }
    __syncthreads();
    if (ld_dst == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t r = r0 + ty + 8 * j;
        if (r < rows && c0 + tx < cols) dst[r * ld_src + c0 + tx] = t[ty + 8 * j][tx];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + ty + 8 * j;
        if (c < cols && r0 + tx < rows) dst[(int64_t)c * ld_dst + r0 + tx] = t[tx][ty + 8 * j];
      }
    }
    __syncthreads();
  }
}

// CSB_TF32X3 operands: dst holds THREE copies of src (transposed if `transpose`), `blk` elements apart along dst's column (contraction)
// axis:  A-side [hi | lo | hi],  B-side (mode_b) [hi | hi | lo],  hi = x rounded to the TF32 grid, lo = (x - hi) rounded to it again.
// One GEMM over the tripled contraction then sums  hi.hi + lo.hi + hi.lo  -- the fp32 product up to the dropped lo.lo term and the
// rounding of lo, ~2^-22 relative each and unbiased.  transpose: dst columns rows..rows_pad-1 of every
// copy are written as zeros (the contraction is cut into 32-element blocks).
__global__ void __launch_bounds__(256) split3_f32_kernel(const float* __restrict__ src, int64_t ld_src, int64_t rows, int cols,
                                                         float* __restrict__ dst, int64_t ld_dst, int64_t blk, int transpose, int mode_b,
                                                         int64_t rows_pad) {
  __shared__ float t[32][33];
  const int64_t rows_all = transpose ? rows_pad : rows;
  const int64_t tiles_c = (cols + 31) / 32, tiles_r = (rows_all + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int64_t tile = blockIdx.x; tile < tiles_c * tiles_r; tile += gridDim.x) {
    const int64_t r0 = (tile / tiles_c) * 32;
    const int c0 = (int)(tile % tiles_c) * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t r = r0 + ty + 8 * j;
      t[ty + 8 * j][tx] = (r < rows && c0 + tx < cols) ? src[r * ld_src + c0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x;
      int64_t o;
      bool ok;
      if (transpose) {
        const int c = c0 + ty + 8 * j;
        x = t[tx][ty + 8 * j];
        ok = c < cols && r0 + tx < rows_all;
        o = (int64_t)c * ld_dst + r0 + tx;
      } else {
        const int64_t r = r0 + ty + 8 * j;
        x = t[ty + 8 * j][tx];
        ok = r < rows && c0 + tx < cols;
        o = r * ld_dst + c0 + tx;
      }
      if (ok) {
        // both parts ROUNDED to nearest onto the TF32 grid (the tensor core then reads them exactly): truncating instead leaves a
        // one-sided 2^-20 error per product that seven chained layers add up to 1e-5
        const float hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
        const float lo = __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xFFFFE000u);
        dst[o] = hi;
        dst[o + blk] = mode_b ? hi : lo;
        dst[o + 2 * blk] = mode_b ? lo : hi;
      }
    }
    __syncthreads();
  }
}

// the same keep decisions on an fp32 buffer (element e belongs to group e / 8, LCG step e % 8), so that the CSB_F32 and CSB_BF16 engines
// drop the same elements for the same (seed, step, layer)
__global__ void __launch_bounds__(256)
dropout_f32_kernel(float* __restrict__ a, int64_t n8, uint32_t seed, uint32_t keep_threshold, float scale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v0 = reinterpret_cast<const float4*>(a)[2 * i], v1 = reinterpret_cast<const float4*>(a)[2 * i + 1];
    float e[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint32_t st = mix32(seed ^ mix32((uint32_t)i) ^ (uint32_t)(i >> 32) * 0x9E3779B1u);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      st = st * 747796405u + 2891336453u;
      e[j] = (st >> 8) >= keep_threshold ? e[j] * scale : 0.f;
    }
    reinterpret_cast<float4*>(a)[2 * i] = make_float4(e[0], e[1], e[2], e[3]);
    reinterpret_cast<float4*>(a)[2 * i + 1] = make_float4(e[4], e[5], e[6], e[7]);
  }
}
// test hook: the multipliers (0 or scale) those decisions amount to, for the first `n` of `ld` columns of every row -> dst [rows, n]
__global__ void __launch_bounds__(256)
dropout_mask_kernel(float* __restrict__ dst, int64_t rows, int n, int ld, uint32_t seed, uint32_t keep_threshold, float scale) {
  const int64_t n8 = rows * ld / 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t st = mix32(seed ^ mix32((uint32_t)i) ^ (uint32_t)(i >> 32) * 0x9E3779B1u);
    const int64_t e0 = 8 * i, r = e0 / ld;
    const int c0 = (int)(e0 - r * ld);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      st = st * 747796405u + 2891336453u;
      if (c0 + j < n) dst[r * n + c0 + j] = (st >> 8) >= keep_threshold ? scale : 0.f;
    }
  }
}

// test hook: hidden activation buffer (bf16, halo layout [B*(L+2), ld]) -> fp32 (B, L, C)
__global__ void cnn_unpack_hidden_kernel(const __nv_bfloat16* __restrict__ hbuf, int ld, float* __restrict__ y, int64_t B, int L, int C) {
  const int64_t total = B * L * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int l = (int)((i / C) % L);
    const int64_t b = i / ((int64_t)C * L);
    y[i] = __bfloat162float(hbuf[(b * (L + 2) + l + 1) * ld + c]);
  }
}

// dz = g * act'(a) * scale element-wise (bf16, 8 elements per thread)
__global__ void __launch_bounds__(256)
act_mask_bf16_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ a, __nv_bfloat16* __restrict__ dz,
                     int64_t n8, int act, float alpha, float scale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 gv = reinterpret_cast<const uint4*>(g)[i], av = reinterpret_cast<const uint4*>(a)[i];
    const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w}, aw[4] = {av.x, av.y, av.z, av.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = pack_bf16x2(bf16_lo(gw[j]) * act_bwd_from_out(act, alpha, bf16_lo(aw[j])) * scale,
                         bf16_hi(gw[j]) * act_bwd_from_out(act, alpha, bf16_hi(aw[j])) * scale);
    reinterpret_cast<uint4*>(dz)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
// conv weight repack from the fp32 master W [taps][Cinp][Coutp]:
//   wt16 [Coutp][taps*Cinp]            (B operand of the forward GEMM:  k = t*Cinp + ci)
//   wd16 [Cinp][taps*Coutp], flipped   (B operand of the data-gradient GEMM: k = t'*Coutp + co with t' = taps-1-t)
// The copies may sit inside a wider matrix shared with another layer (row pitch *_ld, first column *_col0): the residual 1x1 kernel
// rides behind the k = 3 kernel of the same block as a fourth "tap" of one merged GEMM (cnn_engine.cuh).
struct ConvRepack { const float* w; __nv_bfloat16* wt16; __nv_bfloat16* wd16; int taps, Cinp, Coutp; int wt_ld, wt_col0, wd_ld, wd_col0; };
struct ConvRepackTable { int n; ConvRepack l[48]; };
__global__ void __launch_bounds__(256) conv_repack_kernel(const ConvRepackTable tab) {
  const ConvRepack L = tab.l[blockIdx.y];
  const int64_t total = (int64_t)L.taps * L.Cinp * L.Coutp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % L.Coutp);
    const int ci = (int)((i / L.Coutp) % L.Cinp);
    const int t = (int)(i / ((int64_t)L.Coutp * L.Cinp));
    const __nv_bfloat16 v = __float2bfloat16_rn(L.w[i]);
    L.wt16[(size_t)co * L.wt_ld + L.wt_col0 + (size_t)t * L.Cinp + ci] = v;
    L.wd16[(size_t)ci * L.wd_ld + L.wd_col0 + (size_t)(L.taps - 1 - t) * L.Coutp + co] = v;
  }
}
// flat user blob <-> padded conv parameters: W [taps][Cin][Cout] <-> [taps][Cinp][Coutp], b [Cout] <-> [Coutp]
struct ConvPad { int taps, Cin, Cout, Cinp, Coutp; size_t w_off, b_off, w_off_user, b_off_user; };
struct ConvPadTable { int n; ConvPad l[48]; };
__global__ void __launch_bounds__(256) conv_pad_copy_kernel(float* __restrict__ padded, float* __restrict__ user, int dir, const ConvPadTable tab) {
  const ConvPad L = tab.l[blockIdx.y];
  const int64_t nw = (int64_t)L.taps * L.Cin * L.Cout, total = nw + L.Cout;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float *pp, *pu;
    if (i < nw) {
      const int co = (int)(i % L.Cout);
      const int ci = (int)((i / L.Cout) % L.Cin);
      const int t = (int)(i / ((int64_t)L.Cout * L.Cin));
      pp = padded + L.w_off + ((size_t)t * L.Cinp + ci) * L.Coutp + co;
      pu = user + L.w_off_user + i;
    } else {
      pp = padded + L.b_off + (i - nw);
      pu = user + L.b_off_user + (i - nw);
    }
    if (dir == 0) *pp = *pu; else *pu = *pp;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fused evaluation: output weighting + per-(column, output) time reductions, then the grid mean (data_utils.py:1112-1362,
// 1432-1497).  Thread = one (grid column, output index); loops over the T time samples (coalesced over the output index).
// ---------------------------------------------------------------------------------------------------------------
struct EvalConsts {
  double hyai[61], hybi[61];
  double p0, ps_mean, ps_span, grav;
  int normalize;
};

__global__ void __launch_bounds__(128)
eval_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ x_norm, int64_t T, int ncol,
                    const double* __restrict__ area_wgt, const double* __restrict__ out_scale, const double* __restrict__ conv,
                    const EvalConsts c, double* __restrict__ scratch) {
  const int j = threadIdx.x;                         // output index 0..127
  const int col = blockIdx.x;
  const double inv_scale = c.normalize ? 1.0 / out_scale[j] : 1.0;      // the reference divides; x * (1/s) differs by <= 1 ulp
  const double aw = area_wgt[col], cv = conv[j];
  const int lev = j < 60 ? j : j - 60;
  const bool prof = j < 120;
  double s_abs = 0, s_sq = 0, s_p = 0, s_t = 0, s_tt = 0;
  for (int64_t t = 0; t < T; ++t) {
    const int64_t r = t * ncol + col;
    double wgt = aw * cv;
    if (prof) {
      double ps = (double)x_norm[r * 124 + 120];
      if (c.normalize) ps = ps * c.ps_span + c.ps_mean;
      const double dp = (c.p0 * c.hyai[lev + 1] + c.hybi[lev + 1] * ps) - (c.p0 * c.hyai[lev] + c.hybi[lev] * ps);
      wgt *= dp / c.grav;
    }
    const double pw = (double)pred[r * 128 + j] * inv_scale * wgt;
    const double tw = (double)target[r * 128 + j] * inv_scale * wgt;
    const double d = pw - tw;
    s_abs += fabs(d); s_sq += d * d; s_p += pw; s_t += tw; s_tt += tw * tw;
  }
  const double n = (double)T;
  double* o = scratch + ((size_t)col * 4) * 128 + j;
  o[0 * 128] = s_abs / n;
  o[1 * 128] = sqrt(s_sq / n);
  o[2 * 128] = 1.0 - s_sq / (s_tt - s_t * s_t / n);
  o[3 * 128] = s_p / n - s_t / n;
}
// grid mean in a fixed order: out[m][j] = mean_col scratch[col][m][j]
__global__ void __launch_bounds__(128) eval_gridmean_kernel(const double* __restrict__ scratch, int ncol, double* __restrict__ out) {
  const int j = threadIdx.x, m = blockIdx.x;
  double s = 0;
  for (int col = 0; col < ncol; ++col) s += scratch[((size_t)col * 4 + m) * 128 + j];
  out[m * 128 + j] = s / (double)ncol;
}

}  // namespace simt
}  // namespace csb
