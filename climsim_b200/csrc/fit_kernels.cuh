// fit_kernels.cuh -- small HBM-bound kernels around the training step that the reference expresses as Keras metrics / torch losses:
//   batch_metrics_kernel     sufficient statistics of Keras' metrics=['mse','mae','accuracy'] for one batch (hpo_baseline_v1.py:127-129)
//   hsr_loss_kernel          the heteroskedastic-regression losses of hsr.py:126-138 (MSE on the mean / Gaussian NLL) and their
//                            gradients w.r.t. both networks' outputs, two passes (loss, then clip-aware gradient)
// All of them stream their inputs once with 128-bit loads, accumulate in fp64, and reduce in a fixed order (deterministic).
#pragma once
#include "common.cuh"

namespace csb {
namespace simt {

// block-level fixed-order sum of `n` doubles per thread into thread 0 (256 threads, 8 warps)
template <int NV>
__device__ __forceinline__ void block_sum_f64(double (&v)[NV], double (*red)[NV]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) red[warp][k] = v[k];
  __syncthreads();
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][k];
      v[k] = t;
    }
}

// The last block to arrive (ticket counter behind the partials) sums the per-block partials in block order.
// partials: [gridDim.x][NV] doubles followed by one 8-byte slot used as the ticket counter (zero before the first launch; the last
// block re-arms it).  Returns true in thread 0 of the last block with the totals in v.
template <int NV>
__device__ __forceinline__ bool last_block_total(double (&v)[NV], double* partials) {
  __shared__ bool is_last;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(partials + (size_t)gridDim.x * NV);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) partials[(size_t)blockIdx.x * NV + k] = v[k];
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return false;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = 0.0;
  for (unsigned b = 0; b < gridDim.x; ++b)
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] += __ldcg(partials + (size_t)b * NV + k);
  *ticket = 0u;
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Keras metrics of one batch.  pred, y: [B, F] fp32 row-major.  out5 = {sum (p-y)^2, sum |p-y|, #rows with argmax(p) == argmax(y),
// B*F, B}.  'accuracy' on a (B, 128) regression output is what Keras resolves to categorical accuracy: argmax over the feature axis,
// first maximum wins (tf.argmax).  A warp per row.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
batch_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ y, int64_t B, int F, double* __restrict__ out5,
                     double* __restrict__ partials) {
  __shared__ double red[8][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[3] = {0.0, 0.0, 0.0};
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < B; r += (int64_t)gridDim.x * 8) {
    const float* pr = pred + r * F;
    const float* yr = y + r * F;
    float se = 0.f, ae = 0.f, bp = -INFINITY, by = -INFINITY;
    int ip = 0x7fffffff, iy = 0x7fffffff;
    for (int c = lane; c < F; c += 32) {
      const float p = __ldcs(pr + c), t = __ldcs(yr + c), d = p - t;
      se = fmaf(d, d, se); ae += fabsf(d);
      if (p > bp) { bp = p; ip = c; }              // ascending c per lane: strict > keeps the first maximum
      if (t > by) { by = t; iy = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      se += __shfl_xor_sync(0xffffffffu, se, o); ae += __shfl_xor_sync(0xffffffffu, ae, o);
      const float op = __shfl_xor_sync(0xffffffffu, bp, o), oy = __shfl_xor_sync(0xffffffffu, by, o);
      const int oip = __shfl_xor_sync(0xffffffffu, ip, o), oiy = __shfl_xor_sync(0xffffffffu, iy, o);
      if (op > bp || (op == bp && oip < ip)) { bp = op; ip = oip; }
      if (oy > by || (oy == by && oiy < iy)) { by = oy; iy = oiy; }
    }
    if (lane == 0) { acc[0] += (double)se; acc[1] += (double)ae; acc[2] += (ip == iy) ? 1.0 : 0.0; }
  }
  block_sum_f64<3>(acc, red);
  if (last_block_total<3>(acc, partials)) {
    out5[0] = acc[0]; out5[1] = acc[1]; out5[2] = acc[2]; out5[3] = (double)B * (double)F; out5[4] = (double)B;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Heteroskedastic regression (baseline_models/HSR/training/hsr.py:126-138).  mu, logprec, y: [B, F] fp32.
//   mle == 0:  loss = mean((y - mu)^2)                                 d/dmu = 2 (mu - y) / n,  d/dlogprec = 0
//   mle == 1:  loss = mean(exp(lp) (y - mu)^2 - lp)                    d/dmu = 2 exp(lp) (mu - y) / n,  d/dlp = (exp(lp) (y - mu)^2 - 1) / n
// followed by torch.clip(loss, -1e5, 1e5): the gradient of a clipped value is zero, so pass 2 reads the scalar loss of pass 1.
// Pass 1 (hsr_loss_kernel): fixed-order fp64 sum -> loss_out (fp32, unclipped mean, as `losses += [loss.item()]` records it).
// Pass 2 (hsr_grad_kernel): dmu / dlp in fp32 (the `dy` the two networks' backward passes take).
// ---------------------------------------------------------------------------------------------------------------
// pred rows have pitch ldp >= F (the engine's padded output width); y is dense [B, F]; F % 4 == 0.
__global__ void __launch_bounds__(256)
hsr_loss_kernel(const float* __restrict__ mu, const float* __restrict__ lp, const float* __restrict__ y, int64_t B, int F, int ldp, int mle,
                float* __restrict__ loss_out, double* __restrict__ partials) {
  __shared__ double red[8][1];
  double acc[1] = {0.0};
  const int q = F / 4;
  const int64_t n4 = B * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / q;
    const int c = (int)(i - r * q) * 4;
    const float4 m4 = *reinterpret_cast<const float4*>(mu + r * ldp + c), y4 = __ldcs(reinterpret_cast<const float4*>(y + r * F + c));
    const float mv[4] = {m4.x, m4.y, m4.z, m4.w}, yv[4] = {y4.x, y4.y, y4.z, y4.w};
    float lv[4] = {0.f, 0.f, 0.f, 0.f};
    if (mle) { const float4 l4 = *reinterpret_cast<const float4*>(lp + r * ldp + c); lv[0] = l4.x; lv[1] = l4.y; lv[2] = l4.z; lv[3] = l4.w; }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d = yv[j] - mv[j];
      s += mle ? (expf(lv[j]) * d * d - lv[j]) : d * d;
    }
    acc[0] += (double)s;
  }
  block_sum_f64<1>(acc, red);
  if (last_block_total<1>(acc, partials)) *loss_out = (float)(acc[0] / ((double)B * (double)F));
}

template <typename TZ> __device__ __forceinline__ void store4_dz(TZ* dst, const float (&g)[4]);
template <> __device__ __forceinline__ void store4_dz<float>(float* dst, const float (&g)[4]) {
  *reinterpret_cast<float4*>(dst) = make_float4(g[0], g[1], g[2], g[3]);
}
template <> __device__ __forceinline__ void store4_dz<__nv_bfloat16>(__nv_bfloat16* dst, const float (&g)[4]) {
  *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]));
}

// dL/dz of the two (linear) output layers, written straight into the engines' dZ buffers [B, ldp] (padding columns zero)
template <typename TZ>
__global__ void __launch_bounds__(256)
hsr_grad_kernel(const float* __restrict__ mu, const float* __restrict__ lp, const float* __restrict__ y, int64_t B, int F, int ldp, int mle,
                const float* __restrict__ loss, TZ* __restrict__ dmu, TZ* __restrict__ dlp) {
  const float l = *loss;
  const float inv_n = (l > 1e5f || l < -1e5f) ? 0.f : (float)(1.0 / ((double)B * (double)F));   // clip(loss): zero gradient outside [-1e5, 1e5]
  const int q = ldp / 4;
  const int64_t n4 = B * q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / q;
    const int c = (int)(i - r * q) * 4;
    float gm[4] = {0.f, 0.f, 0.f, 0.f}, gl[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < F) {
      const float4 m4 = *reinterpret_cast<const float4*>(mu + r * ldp + c), y4 = __ldcs(reinterpret_cast<const float4*>(y + r * F + c));
      const float mv[4] = {m4.x, m4.y, m4.z, m4.w}, yv[4] = {y4.x, y4.y, y4.z, y4.w};
      float lv[4] = {0.f, 0.f, 0.f, 0.f};
      if (mle) { const float4 l4 = *reinterpret_cast<const float4*>(lp + r * ldp + c); lv[0] = l4.x; lv[1] = l4.y; lv[2] = l4.z; lv[3] = l4.w; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = mv[j] - yv[j];
        const float prec = mle ? expf(lv[j]) : 1.f;
        gm[j] = 2.f * prec * d * inv_n;
        gl[j] = mle ? (prec * d * d - 1.f) * inv_n : 0.f;
      }
    }
    store4_dz<TZ>(dmu + r * ldp + c, gm);
    if (dlp) store4_dz<TZ>(dlp + r * ldp + c, gl);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CNN autograd entry: upstream gradient dy (B, L, C) w.r.t. the network output -> dL/dz of the fused Dense heads in the halo layout
// [B*(L+2), ld]: channels < out_lin are linear, the others ReLU (their derivative from the saved prediction: p > 0); halo rows and
// padding channels are zero.
// ---------------------------------------------------------------------------------------------------------------
template <typename TZ>
__global__ void __launch_bounds__(256)
cnn_head_grad_kernel(const float* __restrict__ y_pred, const float* __restrict__ dy, TZ* __restrict__ dz, int ld, int64_t B, int L, int C, int out_lin) {
  const int64_t total = B * (L + 2) * ld;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld);
    const int64_t r = i / ld;
    const int l = (int)(r % (L + 2)) - 1;
    const int64_t b = r / (L + 2);
    float g = 0.f;
    if (c < C && l >= 0 && l < L) {
      const int64_t e = (b * L + l) * C + c;
      g = dy[e] * ((c < out_lin || y_pred[e] > 0.f) ? 1.f : 0.f);
    }
    if constexpr (sizeof(TZ) == 2) dz[i] = __float2bfloat16_rn(g); else dz[i] = g;
  }
}

}  // namespace simt
}  // namespace csb
