// dp_kernels.cuh -- the data-parallel exchange of the training step as ONE kernel over NVLink peer memory (no NCCL on the step's path):
//
//     split partials -> this rank's gradient        (what reduce_partials_kernel does)
//     reduce-scatter over the ranks                 (peer-to-peer LOADS of the other ranks' gradient slices, summed in rank order)
//     all-gather of the reduced slices              (peer-to-peer STORES into every rank's copy of the summed gradient)
//     optimizer rule + bf16 weight copies           (what opt_fused_kernel does)
//
// The reference's only collective is DDP's gradient all-reduce (online_testing/baseline_models/MLP_v2rh/training/
// train_mlp_h5loader.py:195-207); round 1 issued it as one ncclAllReduce between two kernels, fully exposed (87 us per step at 8 GPUs).
// Here every rank maps its peers' gradient slabs (CUDA IPC; NVSwitch gives every pair full bandwidth), and one persistent kernel per
// rank -- one block per SM, all co-resident, so it may spin on flags -- does the four phases back to back.  Cross-GPU ordering: data
// stores, __threadfence_system(), then a monotonically increasing epoch written into the consumer's flag array; consumers poll their
// LOCAL flags (volatile) and then read with L1-bypassing loads.  Every rank sums the ranks' slices in the same order 0..N-1, so the
// summed gradient -- and therefore every replica's weights -- are bit-identical on all ranks, and independent of timing.
//
// Slab layout (one cudaMalloc per rank, exported with cudaIpcGetMemHandle):
//     grads [n]   this rank's gradient (n = P_pad + 4: the last four floats carry {loss share, 0, 0, 0})
//     gsum  [n]   the sum over ranks, written by the owners of the slices
//     flags [64]  uint64: [0..15] "rank r's gradient is complete for epoch e", [16..31] "rank r has written its slice of gsum",
//                 [32] grid-barrier counter (local)
#pragma once
#include "simt_kernels.cuh"

namespace csb {
namespace simt {

constexpr int DP_MAX_RANKS = 8;           // one NVSwitch node

struct DpTable {
  int rank, world;
  unsigned long long epoch;          // 1, 2, 3, ... one per step
  float* grads_peer[DP_MAX_RANKS];   // peers' `grads` (index = rank; [rank] is the local one)
  float* gsum_peer[DP_MAX_RANKS];
  unsigned long long* flags_peer[DP_MAX_RANKS];
  int64_t n;                         // floats in grads / gsum (multiple of 4)
  int64_t slice;                     // floats per rank slice (multiple of 4)
  int64_t opt_items_total;           // sum of fused_opt_items over the layers
  int opt_item_base[CSB_MAX_LAYERS + 1];
  int64_t seg_base[4 * CSB_MAX_LAYERS + 1];   // phase 0: first float4 index of every segment in one flat index space
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// flag store: every thread has executed __threadfence_system() after its data stores and before the grid barrier that precedes this
// store (fence + relaxed store = release pattern), so the flag itself is a plain system-scope relaxed store -- a st.release.sys here
// made the peers wait ~10 us for it
__device__ __forceinline__ void st_flag_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns64() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *p >= want; a peer that never arrives becomes a trap (launch error) after 20 s instead of a hung GPU
__device__ __forceinline__ void spin_until_ge(const unsigned long long* p, unsigned long long want) {
  if (ld_volatile_u64(p) >= want) return;
  const unsigned long long t0 = globaltimer_ns64();
  unsigned it = 0;
  while (ld_volatile_u64(p) < want) {
    if ((++it & 1023u) == 0 && globaltimer_ns64() - t0 > 20000000000ull) __trap();
  }
}
// all blocks of the (co-resident) grid; `counter` only ever grows, `target` = arrivals expected in total after this barrier.
// The block barrier orders every thread's stores before thread 0's system-scope fence (causality is transitive through bar.sync --
// the cooperative-groups grid.sync pattern), so ONE fence per block publishes the block's global and peer stores: a MEMBAR.SYS in
// every thread cost this kernel ~10 us per barrier.
__device__ __forceinline__ void grid_barrier(unsigned long long* counter, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    atomicAdd(counter, 1ull);
    spin_until_ge(counter, target);
    __threadfence_system();
  }
  __syncthreads();
}

// grid = (resident blocks per SM) x SMs, all co-resident (the host sizes it with the occupancy API), 256 threads.  seg: the split partials of this rank (phase 0); opt: the fused optimizer
// table whose "partials" are gsum with one split (phase 3).
__global__ void __launch_bounds__(256, 2) dp_reduce_opt_kernel(const SegmentTable seg, const DpTable dp, const FusedOptTable opt, const OptParams o) {
  __shared__ float t[32][65];
  const int G = gridDim.x;
  unsigned long long* flags = dp.flags_peer[dp.rank];
  unsigned long long* counter = flags + 32;
  float* grads = dp.grads_peer[dp.rank];
  float* gsum = dp.gsum_peer[dp.rank];
  const unsigned long long arrivals0 = (dp.epoch - 1) * 2ull * (unsigned long long)G;
  // phase timestamps of the last step (block 0; flags[40..45], globaltimer ns): read back by csb_mlp_dp_debug
  auto stamp = [&](int k) { if (blockIdx.x == 0 && threadIdx.x == 0) flags[40 + k] = globaltimer_ns64(); };
  stamp(0);

  // ---- phase 0: this rank's gradient = fixed-order sum of its split partials (+ its share of the loss behind the gradient)
  // (one flat float4 index space over all segments: a bias vector of 128 floats does not cost the grid a pass of its own)
  {
    int sidx = 0;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < dp.seg_base[seg.n]; q += (int64_t)G * blockDim.x) {
      while (q >= dp.seg_base[sidx + 1]) ++sidx;
      const Segment& sg = seg.seg[sidx];
      const int64_t i = (q - dp.seg_base[sidx]) * 4;
      *reinterpret_cast<float4*>(sg.grad + i) = sum_partials4(sg.ws + i, sg.stride, sg.splits);
    }
  }
  if (blockIdx.x == G - 1) {
    if (seg.loss_out != nullptr) block_sum_loss(seg.loss_partials, seg.n_loss, grads + dp.n - 4);
    else if (threadIdx.x == 0) grads[dp.n - 4] = 0.f;
  }
  grid_barrier(counter, arrivals0 + (unsigned long long)G);
  stamp(1);

  // ---- tell every rank (self included) that this gradient is complete, then wait for everybody's
  if (blockIdx.x == 0 && (int)threadIdx.x < dp.world) st_flag_sys_u64(dp.flags_peer[threadIdx.x] + dp.rank, dp.epoch);
  if ((int)threadIdx.x < dp.world) spin_until_ge(flags + threadIdx.x, dp.epoch);
  __syncthreads();
  stamp(2);

  // ---- phase 1 + 2: my slice summed over the ranks in rank order (peer loads), written into every rank's gsum (peer stores).
  // A remote load takes ~3 us, so every thread keeps eight in flight whatever the world size: with W ranks it works on 8 / W
  // positions of the slice per iteration (slot s of the unrolled loops = position s / W, rank s % W).
  {
    const int W = dp.world;
    const int P = W <= 2 ? 4 : (W <= 4 ? 2 : 1);
    const int64_t lo = min(dp.n, (int64_t)dp.rank * dp.slice), hi = min(dp.n, lo + dp.slice);
    const int64_t step = (int64_t)G * blockDim.x * 4;
    for (int64_t base = lo + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; base < hi; base += step * P) {
      float4 v[8];
#pragma unroll
      for (int sl = 0; sl < 8; ++sl) {
        const int ps = sl / W, r = sl - ps * W;
        const int64_t i = base + ps * step;
        v[sl] = (ps < P && i < hi) ? __ldcv(reinterpret_cast<const float4*>(dp.grads_peer[r] + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float4 a[4];
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) a[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int sl = 0; sl < 8; ++sl) {             // slots in ascending order = ranks in ascending order within a position
        const int ps_of = sl / W;
#pragma unroll
        for (int ps = 0; ps < 4; ++ps)
          if (ps_of == ps) { a[ps].x += v[sl].x; a[ps].y += v[sl].y; a[ps].z += v[sl].z; a[ps].w += v[sl].w; }
      }
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const int64_t i = base + ps * step;
        if (ps < P && i < hi) {
#pragma unroll
          for (int r = 0; r < DP_MAX_RANKS; ++r)
            if (r < W) *reinterpret_cast<float4*>(dp.gsum_peer[r] + i) = a[ps];
        }
      }
    }
  }
  grid_barrier(counter, arrivals0 + 2ull * (unsigned long long)G);
  stamp(3);
  if (blockIdx.x == 0 && (int)threadIdx.x < dp.world) st_flag_sys_u64(dp.flags_peer[threadIdx.x] + 16 + dp.rank, dp.epoch);
  if ((int)threadIdx.x < dp.world) spin_until_ge(flags + 16 + threadIdx.x, dp.epoch);
  __syncthreads();
  __threadfence_system();                    // acquire side: the slices other ranks stored into gsum are read below
  stamp(4);

  // ---- phase 3: optimizer + bf16 copies from the summed gradient (the same work items as opt_fused_kernel, flattened over layers)
  for (int64_t w = blockIdx.x; w < dp.opt_items_total; w += G) {
    int l = 0;
    while (l + 1 < opt.n && w >= dp.opt_item_base[l + 1]) ++l;
    fused_opt_item(opt, o, opt.l[l], (int)(w - dp.opt_item_base[l]), t);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && opt.loss_out != nullptr) *opt.loss_out = __ldcv(gsum + dp.n - 4);   // the GLOBAL loss
  stamp(5);                                  // (block 0's own end: the other blocks finish within one work item of it)
}

}  // namespace simt
}  // namespace csb
