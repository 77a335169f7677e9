// tc_gemm.cuh -- hand-written Blackwell (sm_100a) tensor-core GEMMs for the dense layers of the column emulators.
//
// Two kernels, both: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory -> tcgen05.mma (kind::f16, bf16
// operands, fp32 accumulators in TMEM) -> tcgen05.ld -> fused epilogue; warp-specialised (one TMA producer thread, one
// MMA-issuing thread, the remaining warps run the epilogue, one TMEM lane quadrant each); optionally on CTA pairs
// (tcgen05 cta_group::2: a 256-row tile across two SMs, each loading half of the B operand).
//
//   gemm_tn_kernel   D[M,N] = A[M,K] . Bt[N,K]^T      both operands K-major.  Forward layers (A = activations,
//                    Bt = W^T) and data-gradient layers (A = dZ, Bt = W).  Persistent over 128 x BN tiles (256 x BN per
//                    pair), STAGES-deep smem ring, two TMEM accumulator buffers so the epilogue of tile i overlaps the
//                    MMAs of tile i+1; 18 warps (issuer, producer, 16 epilogue).  Epilogues: bias+activation->bf16 (+ sign
//                    mask), head+loss (+dZ), act'-masked dgrad, residual add, fp32; results go registers -> global, or through a
//                    per-warp staging tile and coalesced stores for the short-contraction launches (VAR_STAGED).
//                    The same kernel with fp32 operands and tcgen05 kind::tf32 (VAR_TF32) is the CSB_TF32 / CSB_TF32X3 mode.
//   gemm_nt_kernel   D[M,N] = A[R,M]^T . B[R,N]        both operands MN-major (row index R = batch is the
//                    contraction).  Weight gradients dW = H^T dZ, bias gradients as column sums of the dZ tiles in flight.
//                    One 128 x BN (256 x BN per pair) tile per CTA, split over R (blockIdx.y) into fp32 partials that a
//                    later kernel reduces deterministically; 6 warps; output through shared memory + bulk-copy stores.
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
constexpr int NUM_THREADS = 192; // gemm_nt_kernel: 4 epilogue warps + producer + mma issuer

enum Epi : int {
  EPI_BIAS_ACT = 0,   // out(bf16) = act(acc + bias)                       forward hidden layer
  EPI_HEAD_LOSS = 1,  // p = head(acc + bias); loss, dZ(bf16), optional p  forward output layer fused with the loss
  EPI_DGRAD = 2,      // out(bf16) = acc * act'(saved activation)          backward data gradient
  EPI_F32 = 3,        // out(fp32) = acc                                   (self-test, fp32 consumers)
  EPI_HEAD_OUT = 4,   // p = head(acc + bias) -> fp32 (optionally / out_scale)   inference output layer
  EPI_BIAS_ADD = 5,   // out(bf16) = acc + bias + saved tile                 residual add / gradient accumulation
  EPI_DGRAD_MASK = 6  // out(bf16) = acc * act'(sign bit)                    backward data gradient of ReLU / LeakyReLU layers:
                      //   act' needs only "activation > 0", stored by the forward epilogue as one bit per element
};

struct GemmParams {
  int M, N, K;                 // problem (N, K multiples of 64; M arbitrary)
  int b_box_rows;              // rows of the B-operand TMA box (== min(N, BN)): sets the expected transaction bytes
  // Conv1D('same') as a row-shifted GEMM over a halo-padded channels-last layout [B*(L+2), C]: contraction block kb
  // belongs to tap t = kb / kb_per_tap and reads the A rows shifted by (t - tap_center); rows whose index modulo
  // halo_period is 0 or halo_period-1 are the zero halo rows of each sample and are written as zeros.
  int kb_per_tap;              // 0: plain GEMM
  int tap_center;
  int halo_period;             // 0: no halo rows
  uint32_t* mask_out;          // EPI_BIAS_ACT: optional sign-bit mask (bit j of word w <-> column 32 w + j, set iff out > 0)
  const uint32_t* mask_in;     // EPI_DGRAD_MASK
  int ld_mask;                 // chunk-major layout [N/32][ld_mask]: word of (row r, columns 32w..32w+31) at w * ld_mask + r
  int balanced;                // tile schedule: 0 round-robin over the (m, n-block) grid, 1 balanced contiguous unit ranges (TileIter)
  int dbg;                     // micro-benchmark knobs (scripts/microbench_gemm.py): 1 skip bias staging, 2 skip epilogue math + smem
                               // stores, 4 skip TMA store, 8 skip TMEM loads, 16 skip all TMA loads, 32 skip the MMAs, 64 skip the
                               // B loads, 128 skip the A loads.  0 in production.
  unsigned long long* stats;   // micro-benchmark: per CTA {clock64 at start, at end, globaltimer ns at start, at end}; NULL in production
  int act;                     // CSB_ACT_* for EPI_BIAS_ACT / EPI_DGRAD (activation of the layer whose output is stored / was saved)
  float alpha;
  int head_relu_from;          // EPI_HEAD_*: columns >= this get ReLU (-1: none); otherwise `act` applies
  float dgrad_scale;           // EPI_DGRAD: extra factor on the result (1 / (1 - p) of a dropout layer behind the activation); 0 = none
  // VAR_DROPOUT (EPI_BIAS_ACT): inverted dropout behind the activation, folded into the epilogue -- the same counter-based keep
  // decisions as simt::dropout_bf16_kernel (one 32-bit mix of (seed, index of the 8-element group in the output buffer) starts an LCG
  // whose high 24 bits decide the eight elements), so the separate element-wise pass over the activation disappears
  uint32_t drop_seed, drop_threshold;
  float drop_scale;
  const float* bias;           // [N]
  void* out;                   // bf16 or fp32 [M, ld_out]
  int ld_out;
  const __nv_bfloat16* saved;  // EPI_DGRAD: saved activation [M, ld_saved]
  int ld_saved;
  // VAR_A2: contraction blocks kb >= a2_from_kb read their A tiles from the SECOND tensor map (columns (kb - a2_from_kb) * 64, rows
  // unshifted): d(x_in) = conv^T(dz1; W1) + conv1x1^T(d_out; Wr) as one GEMM over [dz1 taps | d_out] (cnn_engine.cuh)
  int a2_from_kb;
  // VAR_DUAL (EPI_DGRAD): the accumulator itself, rounded to bf16, is a second output (`out2`), and the masked result is computed from
  // that ROUNDED value -- exactly what a separate act' pass over the stored tensor gives
  void* out2;
  int ld_out2;
  // > 0: the last contraction block of every tap holds only tap_tail_k * 16 non-zero channels (406 = 6 * 64 + 22 -> 2): the issuer
  // skips the MMAs over the all-zero remainder of that block
  int tap_tail_k;
  int a2_tail_k;               // the same for the last contraction block of the second A source (0: all four)
  // VAR_ACC2 (EPI_BIAS_ACT + VAR_A2, BN = 256): the contraction blocks of the second A source accumulate into a SECOND accumulator
  // (TMEM columns 256..511; one tile in flight instead of two).  out = act(acc1 + bias) [dropout] as usual, and
  // out2 = (acc2 + bias2) + bf16(out): a residual block's  relu(conv2(h1)) + conv1x1(x_in)  in one launch (cnn_engine.cuh)
  const float* bias2;
  // VAR_KSPLIT (EPI_F32): the contraction is cut into k_splits ranges of whole k-blocks, range s writes its partial product to
  // out + s * split_stride (floats); the caller sums the partials in a fixed order (the weight gradient of the TF32 mode, whose
  // contraction runs over the batch)
  int k_splits;
  size_t split_stride;
  // VAR_TF32: optional second, TRANSPOSED copy of the stored tile, out_t[col * ld_out_t + row] (the weight gradient of this mode
  // contracts over the batch and wants [features, batch] operands; a warp's 32 rows make every such store one 128-byte line)
  float* out_t;
  int64_t ld_out_t;
  int tf32_exact_store;        // VAR_TF32: 1 = store the fp32 results as they are (CSB_TF32X3: operands are split into hi / lo copies
                               // elsewhere); 0 = round them onto the TF32 grid (CSB_TF32)
  // EPI_HEAD_LOSS / EPI_HEAD_OUT
  const float* y;              // targets [M, ld_y]
  int ld_y;
  int out_dim;                 // valid output columns (<= N)
  const float* loss_w;         // [N] (zero in padding)
  float grad_scale;            // multiplies loss and dL/dp
  int loss_kind;               // CSB_LOSS_*
  float* pred;                 // EPI_HEAD_OUT: fp32 predictions [M, ld_pred]
  int ld_pred;
  const float* inv_out_scale;  // optional [N]: pred *= inv_out_scale
  const float* out_mask;       // optional [N] of 0/1: predictions (and their gradients) of masked columns are zero
  float* loss_partials;        // [num_m_blocks * num_n_blocks * TN_EPI_WARPS]
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
// try_wait with a suspend-time hint (ns): the thread may sleep in hardware until the phase completes or the hint expires, instead of
// coming back after the short default time-out -- far fewer wake-ups (and issued instructions) over a long wait
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok;
}
// Wait with a watchdog: a protocol bug becomes a trap (launch error) instead of a hung GPU.  The watchdog counts wake-ups; it does
// not read a timer (ncu: in the wide forward kernel the try_wait / globaltimer / compare / branch loops of the 18 mostly-waiting
// warps were 60 % of all issued instructions, competing for issue slots with the one thread that feeds the tensor pipe and with the
// epilogue warps that do have work).
//   mbar_wait       latency-critical waits (MMA issuer, TMA producer): plain try_wait retries
//   mbar_wait_long  waits that are expected to take a while and tolerate ~a hundred cycles of wake-up latency (an epilogue warp
//                   waiting for its accumulator): hardware sleep of up to 2 us per try
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++n > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void mbar_wait_long(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t n = 0;
  while (!mbar_try_wait_hint(bar, parity, 2000u)) {
    if (++n > (1u << 22)) __trap();
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

// bulk (TMA, non-tensor) copy shared -> global of `bytes` (multiple of 16; both addresses 16-byte aligned), tracked by bulk groups
__device__ __forceinline__ void bulk_store_s2g(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory (st.shared) -> visible to the async proxy (TMA) of this thread's later bulk copies
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
// kind::tf32: the operands are fp32 words in shared memory, of which the tensor core uses sign, exponent and the top 10 mantissa bits
// (what TF32 on the reference's A100 runs does); 8 elements = 32 B of contraction per instruction, fp32 accumulation
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- cta_group::2 (CTA pair) variants --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in the pair's leader (even) CTA
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  // default semantics on purpose: `.release.cluster` makes ptxas put MEMBAR.ALL.GPU in front of every arrive (and an
  // `.acquire.cluster` wait adds CCTL.IVALL), which halves the kernel.  What the arrive orders here is shared memory written by
  // TMA / read by tcgen05 (async proxy, observed through the mbarrier chain) and TMEM, neither of which those fences serve.
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
// TMA load issued by either CTA of the pair; the transaction bytes are credited to the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256 (128 rows per CTA), B split along N between the two CTAs; leader only
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit: arrive on the mbarrier at this offset in BOTH CTAs of the pair once the issued MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
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

// kind::tf32 instruction descriptor: the same fields with A, B format 2 (TF32)
__device__ __forceinline__ uint32_t make_idesc_tf32(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------------
// epilogue helpers: each thread owns one output row (TMEM lane) and 32 consecutive columns per step
// ---------------------------------------------------------------------------------------------------------------
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
// ---- register <-> global row pieces.  Each epilogue thread owns one output row; a 32-column step is 64 contiguous bytes of
// bf16 in that row = two full 32-byte sectors, moved with 256-bit accesses.  (The first version staged the tile in shared
// memory and wrote it with a TMA store; on B200 the TMA store queues in front of the mainloop's TMA loads and stalls them
// for the ~2 k cycles the SM needs to push 64 KB to L2 -- see profiles/r01_probe_mainloop.txt.)
__device__ __forceinline__ void store_bf16x32_global(__nv_bfloat16* dst, const float (&v)[32]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(v[16 * h + 2 * i], v[16 * h + 2 * i + 1]);
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 16 * h), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
  }
}
// 32 bf16 (64 B) into shared memory at a 16-byte aligned address
__device__ __forceinline__ void stage_bf16x32(uint32_t saddr, const float (&v)[32]) {
#pragma unroll
  for (int h = 0; h < 4; ++h)
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr + 16u * h), "r"(pack_bf16x2(v[8 * h], v[8 * h + 1])),
                 "r"(pack_bf16x2(v[8 * h + 2], v[8 * h + 3])), "r"(pack_bf16x2(v[8 * h + 4], v[8 * h + 5])),
                 "r"(pack_bf16x2(v[8 * h + 6], v[8 * h + 7])) : "memory");
}
__device__ __forceinline__ void load_bf16x32_global_raw(const __nv_bfloat16* src, uint32_t (&w)[16]) {
#pragma unroll
  for (int h = 0; h < 2; ++h)
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[8 * h + 0]), "=r"(w[8 * h + 1]), "=r"(w[8 * h + 2]), "=r"(w[8 * h + 3]), "=r"(w[8 * h + 4]), "=r"(w[8 * h + 5]),
                   "=r"(w[8 * h + 6]), "=r"(w[8 * h + 7])
                 : "l"(src + 16 * h));
}
// 32 consecutive fp32 (128 B, 32 B aligned) of one row: four 256-bit loads (full sectors, no reliance on L1)
__device__ __forceinline__ void load_f32x32_global_v8(const float* src, float (&v)[32]) {
#pragma unroll
  for (int h = 0; h < 4; ++h)
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[8 * h + 0]), "=f"(v[8 * h + 1]), "=f"(v[8 * h + 2]), "=f"(v[8 * h + 3]), "=f"(v[8 * h + 4]), "=f"(v[8 * h + 5]),
                   "=f"(v[8 * h + 6]), "=f"(v[8 * h + 7])
                 : "l"(src + 8 * h));
}
__device__ __forceinline__ void unpack_bf16x32(const uint32_t (&w)[16], float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) { v[2 * i] = bf16_lo(w[i]); v[2 * i + 1] = bf16_hi(w[i]); }
}
__device__ __forceinline__ void load_smem_f32x32(const float* s, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(s + 4 * q);   // same address across the warp: broadcast
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}

constexpr int TN_BIAS_SMEM = 4096;     // widest layer (columns) the kernel accepts: its bias vector lives in shared memory

// cheap activations (ReLU / LeakyReLU / none) over a 32-vector, branch-free inside the element loop; ELU is a separate kernel
// instantiation (template flag) so that expm1f's code never sits in the instruction stream of the common kernels
template <bool ELU>
__device__ __forceinline__ void act_fwd32(int act, float alpha, float (&v)[32]) {
  if constexpr (ELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : expm1f(v[j]);
  } else {
    if (act == CSB_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == CSB_ACT_LEAKYRELU) {
      if (alpha >= 0.f && alpha <= 1.f) {               // warp-uniform: leaky(v) == max(v, alpha v) for a slope in [0, 1]
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], alpha * v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : alpha * v[j];
      }
    }
  }
}
// v *= act'(a) with a = saved activation output
template <bool ELU>
__device__ __forceinline__ void act_bwd32(int act, float alpha, float (&v)[32], const float (&a)[32]) {
  if constexpr (ELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = a[j] > 0.f ? v[j] : v[j] * (a[j] + 1.f);
  } else {
    if (act == CSB_ACT_RELU || act == CSB_ACT_LEAKYRELU) {
      const float neg = act == CSB_ACT_LEAKYRELU ? alpha : 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = a[j] > 0.f ? v[j] : neg * v[j];
    }
  }
}

// per-tile facts about the output row a thread owns
struct RowInfo {
  int grow;          // global row of the GEMM
  bool in_range;     // grow < M: the row exists (stores allowed)
  bool row_ok;       // a real sample row (not a conv halo row, not past M): contributes to the loss
  bool zero_row;     // conv halo row: stored as zeros
  int64_t yrow;      // row of the user's target / prediction arrays (no halo rows there)
};
__device__ __forceinline__ RowInfo make_row_info(const GemmParams& p, int grow) {
  RowInfo r;
  r.grow = grow; r.in_range = grow < p.M; r.row_ok = r.in_range; r.yrow = grow;
  if (p.halo_period > 0) {
    const int rr = grow % p.halo_period;
    if (rr == 0 || rr == p.halo_period - 1) r.row_ok = false;
    r.yrow = (int64_t)(grow / p.halo_period) * (p.halo_period - 2) + rr - 1;
  }
  r.zero_row = p.halo_period > 0 && !r.row_ok;
  return r;
}

// One 32-column step of the epilogue.  gcol = global column; sbias = the layer's bias vector in shared memory;
// `sv` holds the 32 saved bf16 values of EPI_DGRAD / EPI_BIAS_ADD.
// STAGED: the 32 bf16 results go to this thread's row of the warp's staging tile in shared memory (`stage_addr`); the warp writes the
// tile to global memory afterwards with fully coalesced stores (see the kernel).
__device__ __forceinline__ uint32_t drop_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// round-to-nearest onto the TF32 grid (10 mantissa bits): kind::tf32 itself TRUNCATES the low 13 bits of what it reads, which biases
// every product downwards by ~2^-11 and adds up over a chain of layers; tensors that will be tensor-core operands again are therefore
// stored already rounded (the truncation is then exact)
__device__ __forceinline__ float round_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
// F32IO store of one 32-column step: rounded onto the TF32 grid unless the caller keeps exact values, plus the optional transposed copy
__device__ __forceinline__ void store_tf32_step(const GemmParams& p, const RowInfo& ri, int gcol, float (&v)[32], bool st_ok) {
  if (!p.tf32_exact_store) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
  }
  if (!st_ok) return;
  store_f32x32(reinterpret_cast<float*>(p.out) + (size_t)ri.grow * p.ld_out + gcol, v);
  if (p.out_t != nullptr) {
    float* t = p.out_t + (size_t)gcol * (size_t)p.ld_out_t + ri.grow;
#pragma unroll
    for (int j = 0; j < 32; ++j) t[(size_t)j * (size_t)p.ld_out_t] = v[j];
  }
}
// F32IO (VAR_TF32): the stored tensors (outputs, saved activations) are fp32 instead of bf16.
template <int EPI, bool ELU, bool GENERAL_LOSS, bool STAGED, bool DROPOUT = false, bool DUAL = false, bool ACC2 = false, bool F32IO = false>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const float* sbias, const float* sloss_w, const RowInfo& ri, int gcol,
                                               const uint32_t (&raw)[32], const uint32_t (&sv)[16], const float (&yv)[32], float& loss_acc,
                                               uint32_t mask_word, uint32_t stage_addr, uint32_t taddr2 = 0, size_t out_off = 0) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
  const bool st_ok = ri.in_range && !(p.dbg & 4);
  __nv_bfloat16* out16 = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)ri.grow * p.ld_out + gcol;
  [[maybe_unused]] float* out32 = reinterpret_cast<float*>(p.out) + (size_t)ri.grow * p.ld_out + gcol;      // F32IO
  constexpr bool USE_BIAS = (EPI == EPI_BIAS_ACT || EPI == EPI_BIAS_ADD || EPI == EPI_HEAD_LOSS || EPI == EPI_HEAD_OUT);
  if constexpr (USE_BIAS) {
    if (!(p.dbg & 1)) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(sbias + gcol + 4 * q);   // same address across the warp: broadcast
        v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
      }
    }
  }

  if constexpr (EPI == EPI_F32) {
    if (ri.row_ok) store_f32x32(reinterpret_cast<float*>(p.out) + out_off + (size_t)ri.grow * p.ld_out + gcol, v);
  } else if constexpr (EPI == EPI_BIAS_ACT) {
    if (p.mask_out != nullptr) {
      // sign bits of the pre-activation (the activations here preserve the sign; act' of ReLU / LeakyReLU needs nothing else):
      // one funnel shift per element collects "z < 0" from column 31 down to column 0, so that bit j <-> column j.
      // Chunk-major layout: the 32 rows of a warp write 32 consecutive words.  (z == +0 counts as positive; TF's ReluGrad
      // gives 0 there -- a measure-zero difference.)
      uint32_t m = 0;
#pragma unroll
      for (int j = 31; j >= 0; --j) m = __funnelshift_l(__float_as_uint(v[j]), m, 1);
      if (st_ok) p.mask_out[(size_t)(gcol >> 5) * p.ld_mask + ri.grow] = ~m;
    }
    act_fwd32<ELU>(p.act, p.alpha, v);
    if constexpr (DROPOUT) {
      const unsigned long long g0 = ((unsigned long long)ri.grow * (unsigned long long)p.ld_out + (unsigned long long)gcol) >> 3;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const unsigned long long i = g0 + g;
        uint32_t st = drop_mix32(p.drop_seed ^ drop_mix32((uint32_t)i) ^ (uint32_t)(i >> 32) * 0x9E3779B1u);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          st = st * 747796405u + 2891336453u;
          v[8 * g + j] = (st >> 8) >= p.drop_threshold ? v[8 * g + j] * p.drop_scale : 0.f;
        }
      }
    }
    if (ri.zero_row) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    if constexpr (F32IO) store_tf32_step(p, ri, gcol, v, st_ok);
    else if constexpr (STAGED) stage_bf16x32(stage_addr, v); else if (st_ok) store_bf16x32_global(out16, v);
    if constexpr (ACC2) {
      // second accumulator (loaded only now: the first one's registers are free again), its bias behind the first bias vector
      uint32_t raw2[32];
      tmem_ld_32x32(taddr2, raw2);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));      // the stored value is what gets added
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(sloss_w + gcol + 4 * q);
        v[4 * q] += __uint_as_float(raw2[4 * q]) + t.x; v[4 * q + 1] += __uint_as_float(raw2[4 * q + 1]) + t.y;
        v[4 * q + 2] += __uint_as_float(raw2[4 * q + 2]) + t.z; v[4 * q + 3] += __uint_as_float(raw2[4 * q + 3]) + t.w;
      }
      if (ri.zero_row) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (st_ok) store_bf16x32_global(reinterpret_cast<__nv_bfloat16*>(p.out2) + (size_t)ri.grow * p.ld_out2 + gcol, v);
    }
  } else if constexpr (EPI == EPI_DGRAD_MASK) {
    // act'(a) through the sign bit: relu -> {1, 0}, leaky relu -> {1, alpha}
    const float neg = p.act == CSB_ACT_LEAKYRELU ? p.alpha : 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (!(mask_word & (1u << j))) v[j] *= neg;               // one predicate-setting LOP3 + one predicated FMUL per element
    if (ri.zero_row) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    if constexpr (F32IO) store_tf32_step(p, ri, gcol, v, st_ok);
    else if constexpr (STAGED) stage_bf16x32(stage_addr, v); else if (st_ok) store_bf16x32_global(out16, v);
  } else if constexpr (EPI == EPI_DGRAD) {
    if constexpr (DUAL) {
      if (ri.zero_row) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (st_ok) store_bf16x32_global(reinterpret_cast<__nv_bfloat16*>(p.out2) + (size_t)ri.grow * p.ld_out2 + gcol, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
    }
    float a[32];
    if constexpr (F32IO) {
      if (ri.in_range) load_f32x32(reinterpret_cast<const float*>(p.saved) + (size_t)ri.grow * p.ld_saved + gcol, a);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = 0.f;
      }
    } else {
      unpack_bf16x32(sv, a);                        // saved activation (same rows / columns as the output)
    }
    act_bwd32<ELU>(p.act, p.alpha, v, a);
    if (p.dgrad_scale != 0.f) {
      // kernel-uniform: a dropout layer sits behind the activation -- a dropped element's saved output is exactly 0 (also where
      // relu'(0) = 0 already did it); the kept ones carry 1 / (1 - p)
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = a[j] == 0.f ? 0.f : v[j] * p.dgrad_scale;
    }
    if (ri.zero_row) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    if constexpr (F32IO) store_tf32_step(p, ri, gcol, v, st_ok);
    else if constexpr (STAGED) stage_bf16x32(stage_addr, v); else if (st_ok) store_bf16x32_global(out16, v);
  } else if constexpr (EPI == EPI_BIAS_ADD) {
    float a[32];
    unpack_bf16x32(sv, a);                          // tile to add (residual branch / partial gradient)
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = ri.zero_row ? 0.f : v[j] + a[j];
    if constexpr (STAGED) stage_bf16x32(stage_addr, v); else if (st_ok) store_bf16x32_global(out16, v);
  } else {
    // head: p = (col >= head_relu_from) ? relu(z) : act(z); everything is computed element by element in place (v: z -> p -> dL/dz)
    // so that the live set stays at the accumulator chunk + the prefetched targets (a spill here exposes the target loads' latency)
    const float slope = p.act == CSB_ACT_RELU ? 0.f : (p.act == CSB_ACT_LEAKYRELU ? p.alpha : 1.f);   // relu / leaky / identity as one formula
    const bool relu_cols = p.head_relu_from >= 0 && gcol + 32 > p.head_relu_from;                      // warp-uniform
    auto head_elem = [&](int j, float z, float& pj, float& da) {
      if constexpr (ELU) {
        pj = z > 0.f ? z : expm1f(z);
        da = z > 0.f ? 1.f : pj + 1.f;
      } else {
        pj = z > 0.f ? z : slope * z;
        da = z > 0.f ? 1.f : slope;
      }
      if (relu_cols && gcol + j >= p.head_relu_from) { pj = fmaxf(z, 0.f); da = z > 0.f ? 1.f : 0.f; }
    };
    if constexpr (EPI == EPI_HEAD_OUT) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float pj, da;
        head_elem(j, v[j], pj, da);
        v[j] = pj;
      }
      if (p.out_mask != nullptr) {
        float mk[32];
        load_f32x32(p.out_mask + gcol, mk);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= mk[j];
      }
      if (p.inv_out_scale != nullptr) {
        float sc[32];
        load_f32x32(p.inv_out_scale + gcol, sc);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= sc[j];
      }
      if (ri.row_ok) {
        if (gcol + 32 <= p.out_dim) {
          store_f32x32(p.pred + (size_t)ri.yrow * p.ld_pred + gcol, v);
        } else {
          for (int j = 0; j < 32; ++j)
            if (gcol + j < p.out_dim) p.pred[(size_t)ri.yrow * p.ld_pred + gcol + j] = v[j];
        }
      }
    } else {  // EPI_HEAD_LOSS: targets were requested before the accumulator wait (yv), loss weights sit in shared memory
      // the common configuration -- identity (or ReLU / LeakyReLU) head, MSE, no output mask -- gets a loop of ~6 instructions
      // per element; the general loop below costs ~30 and made this kernel ALU-bound (128 x 128 outputs per 2 k-blocks of MMA)
      if constexpr (!ELU && !GENERAL_LOSS) {      // host guarantees: MSE, no output mask
        const float two_gs = 2.f * p.grad_scale;
        if (!ri.row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        } else {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 w4 = *reinterpret_cast<const float4*>(sloss_w + gcol + 4 * g);
            const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = 4 * g + j;
              const float z = v[c];
              const float sc = (relu_cols && gcol + c >= p.head_relu_from) ? 0.f : slope;      // warp-uniform per column
              const float sl = z > 0.f ? 1.f : sc;
              const float wd = w[j] * (sl * z - yv[c]);        // w (p - y)
              loss_acc = fmaf(wd, sl * z - yv[c], loss_acc);
              v[c] = sl * (two_gs * wd);
            }
          }
        }
      } else {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 w4 = *reinterpret_cast<const float4*>(sloss_w + gcol + 4 * g);      // broadcast; zero in padding columns (general path)
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
        float mk[4] = {1.f, 1.f, 1.f, 1.f};
        if (p.out_mask != nullptr) {
          const float4 m4 = __ldg(reinterpret_cast<const float4*>(p.out_mask + gcol + 4 * g));
          mk[0] = m4.x; mk[1] = m4.y; mk[2] = m4.z; mk[3] = m4.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 4 * g + j;
          float pj, da;
          head_elem(c, v[c], pj, da);
          pj *= mk[j]; da *= mk[j];
          const float d = ri.row_ok ? pj - yv[c] : 0.f;
          float dl;                                              // dL/dp
          if (p.loss_kind == CSB_LOSS_MSE) {
            loss_acc += w[j] * d * d;
            dl = 2.f * w[j] * d * p.grad_scale;
          } else if (p.loss_kind == CSB_LOSS_HUBER) {
            const float ad = fabsf(d);
            loss_acc += w[j] * (ad <= 1.f ? 0.5f * d * d : ad - 0.5f);
            dl = w[j] * p.grad_scale * (ad <= 1.f ? d : (d > 0.f ? 1.f : -1.f));
          } else {
            loss_acc += w[j] * fabsf(d);
            dl = w[j] * p.grad_scale * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
          }
          v[c] = ri.row_ok ? da * dl : 0.f;
        }
      }
      }
      if constexpr (STAGED) stage_bf16x32(stage_addr, v); else if (st_ok) store_bf16x32_global(out16, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// gemm_tn_kernel: persistent, K-major x K-major.   18 warps: 0 = TMEM owner + MMA issuer, 1 = TMA producer, 2-17 epilogue
// (four per TMEM lane quadrant = four per SM sub-partition, each taking a quarter of the tile's columns).
// The epilogue warps are independent of each other: each waits for the accumulator, walks its 32-column steps
// (tcgen05.ld -> registers -> fused math -> 256-bit global stores) and releases the TMEM buffer; no staging tile, no
// named barriers, so all of the shared memory beyond the bias vector belongs to the operand ring.  Four warps per
// sub-partition (instead of two with twice the columns) is what hides the tcgen05.ld / store latencies.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TN_EPI_WARPS = 16;
constexpr int TN_EPI_THREADS = TN_EPI_WARPS * 32;
constexpr int TN_THREADS = TN_EPI_THREADS + 64;
// The MMA issuer and the TMA producer are the two LOWEST-numbered warps: the warp scheduler favours older (lower-numbered) warps
// among the ready ones, and the single thread that feeds the tensor pipe must never queue behind the epilogue warps' long
// ALU bursts (as warp 17 it did: the epilogue's issue cycles showed up one-for-one in the tile time).
constexpr int TN_MMA_WARP = 0, TN_PRODUCER_WARP = 1, TN_FIRST_EPI_WARP = 2;

template <int BN, int STAGES, int CG = 1, bool STAGED = false>
struct TnSmem {
  static constexpr int A_BYTES = BM * BK * 2;   // 16 KB
  static constexpr int B_BYTES = (BN / CG) * BK * 2;   // cta_group::2: each CTA of the pair holds half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BIAS_OFFSET = STAGES * STAGE_BYTES;
  // STAGED: one output staging tile per epilogue warp, 32 rows x (BN / 4 bf16 + 16 B): the odd multiple of 16 B keeps the
  // row-per-lane 16-byte writes and the row-contiguous reads both free of bank conflicts
  static constexpr int OUT_PITCH = (BN / 4) * 2 + 16;
  static constexpr int OUT_WARP_BYTES = 32 * OUT_PITCH;
  static constexpr int OUT_OFFSET = BIAS_OFFSET + TN_BIAS_SMEM * 4;
  static constexpr int BAR_OFFSET = OUT_OFFSET + (STAGED ? TN_EPI_WARPS * OUT_WARP_BYTES : 0);
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;   // barriers + slack for manual 1024 B alignment
  static_assert(TOTAL <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
};

// CG = 1: one CTA per 128 x BN tile.  CG = 2 (launched as clusters of two): a CTA pair owns a 256 x BN tile with
// tcgen05 cta_group::2 -- each CTA loads its 128 rows of A and HALF of the B tile, the leader CTA issues the MMAs for
// both, each CTA runs the epilogue of its own 128 rows.  Halves the B traffic and the B footprint per stage.
// VAR: bit 0 = ELU activation (expm1f in the epilogue), bit 1 = general loss path of EPI_HEAD_LOSS (MAE / Huber / output mask);
// both keep rarely used code out of the instruction stream (and the register budget) of the common kernels
// bit 2 = results staged in shared memory and written with coalesced stores (the launches whose time is the epilogue: K <= 256)
// bit 3 = in-kernel cycle counters for the micro-benchmark (p.stats); production instantiations carry none of that code
// bit 4 = inverted dropout behind the activation inside the epilogue (CNN blocks)
// bit 5 = the last contraction blocks come from a SECOND A tensor map (GemmParams.a2_from_kb): two GEMMs with one accumulator
// bit 6 = EPI_DGRAD with two outputs: the rounded accumulator itself and its act'-masked version (CNN: d(block input) + dz2)
// bit 7 = the second A source accumulates into a second TMEM accumulator (conv2 + residual 1x1 forward; opt-in, no gain)
// bit 8 = skip the MMAs over the all-zero channel tail of a tap's last block (GemmParams.tap_tail_k)
// bit 9 = fp32 operands through tcgen05 kind::tf32, fp32 stored tensors (CSB_TF32 / CSB_TF32X3), optional transposed second store
// bit 10 = split contraction: the tile grid is (split, m-group, n-block), fp32 partial products (the TF32 weight gradient)
constexpr int VAR_ELU = 1, VAR_GENERAL_LOSS = 2, VAR_STAGED = 4, VAR_STATS = 8, VAR_DROPOUT = 16, VAR_A2 = 32, VAR_DUAL = 64, VAR_ACC2 = 128, VAR_KTRIM = 256, VAR_TF32 = 512, VAR_KSPLIT = 1024;

// Tile schedule of the persistent kernel; all three warp roles walk the same sequence.
//   round-robin (the first design): tile t = (m-group t / n_blocks, n-block t % n_blocks), CTA group g takes t = g, g + G, ...
//     With 768 tiles on 74 CTA pairs (N = 640 or 768 at B = 65 536) that is 10.4 waves -- the last one 38 % full -- and N = 640
//     leaves every third tile half as wide (256 + 256 + 128), which is bound by operand ingest instead of the tensor pipe.
//   balanced (GemmParams.balanced): the output is cut into 64-column units, m-group-major; every CTA group owns one CONTIGUOUS range of
//     total / G units (+-1), and walks it in tiles of up to BN / 64 units that never cross an m-group: what is left of the row inside
//     the range is split evenly (10 units -> 4 + 3 + 3 rather than 4 + 4 + 2).  All groups finish together (no partial wave), narrow
//     tiles appear only where a range boundary cuts a row, and consecutive tiles of a group re-read the same activation rows from L2.
template <int BN>
struct TileIter {
  static constexpr int MAXU = BN / 64;
  bool BALANCED;
  int num_n_blocks, upm, N;
  long long u, u_end;          // balanced: current / last unit of this group's range
  int tile, num_tiles, stride; // round-robin
  int mg, n0, n_valid, slot;   // current tile: m-group (pairs when CG == 2), first column, width; slot = unique index of the tile in its row
  __device__ __forceinline__ TileIter(bool balanced, int num_m_groups, int N_, int group, int groups) : BALANCED(balanced), N(N_) {
    num_n_blocks = (N_ + BN - 1) / BN;
    upm = (N_ + 63) / 64;
    if (BALANCED) {
      const long long total = (long long)num_m_groups * upm;
      u = total * group / groups;
      u_end = total * (group + 1) / groups;
      tile = num_tiles = stride = 0;
    } else {
      u = u_end = 0;
      tile = group; num_tiles = num_m_groups * num_n_blocks; stride = groups;
    }
    mg = n0 = n_valid = slot = 0;
  }
  __device__ __forceinline__ bool next() {
    if (BALANCED) {
      if (u >= u_end) return false;
      mg = (int)(u / upm);
      const int c = (int)(u - (long long)mg * upm);
      const long long row_end = min(u_end, (long long)(mg + 1) * upm);
      const int r = (int)(row_end - u);
      const int pieces = (r + MAXU - 1) / MAXU;           // tiles still needed for the rest of this row inside the range
      const int w = (r + pieces - 1) / pieces;            // ... of even width
      n0 = c * 64; n_valid = min(w * 64, N - n0); slot = c;
      u += w;
      return true;
    } else {
      if (tile >= num_tiles) return false;
      mg = tile / num_n_blocks;
      slot = tile - mg * num_n_blocks;
      n0 = slot * BN; n_valid = min(BN, N - n0);
      tile += stride;
      return true;
    }
  }
};
template <int BN, int STAGES, int EPI, int CG, int VAR>
__global__ void __launch_bounds__(TN_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p,
               const __grid_constant__ CUtensorMap tmap_a2) {
  constexpr bool ELU = (VAR & VAR_ELU) != 0, GENERAL_LOSS = (VAR & VAR_GENERAL_LOSS) != 0;
  // VAR_TF32: fp32 operands through kind::tf32 (a 128-byte k-block holds 32 elements, an instruction contracts 8), fp32 stored tensors
  constexpr bool TF32 = (VAR & VAR_TF32) != 0;
  static_assert(!TF32 || EPI == EPI_BIAS_ACT || EPI == EPI_DGRAD || EPI == EPI_DGRAD_MASK || EPI == EPI_HEAD_OUT || EPI == EPI_F32,
                "no fp32-storage variant of this epilogue");
  constexpr int BKE = TF32 ? 32 : BK;                  // elements per k-block (TMA coordinates are in elements)
  constexpr bool BF16_OUT = !TF32 && (EPI == EPI_BIAS_ACT || EPI == EPI_HEAD_LOSS || EPI == EPI_DGRAD || EPI == EPI_DGRAD_MASK || EPI == EPI_BIAS_ADD);
  constexpr bool STAGED = (VAR & VAR_STAGED) != 0 && BF16_OUT;
  constexpr bool STATS = (VAR & VAR_STATS) != 0;
  constexpr bool DROPOUT = (VAR & VAR_DROPOUT) != 0 && EPI == EPI_BIAS_ACT;
  constexpr bool A2 = (VAR & VAR_A2) != 0, DUAL = (VAR & VAR_DUAL) != 0 && EPI == EPI_DGRAD;
  constexpr bool ACC2 = (VAR & VAR_ACC2) != 0 && EPI == EPI_BIAS_ACT && A2;
  constexpr bool KTRIM = (VAR & VAR_KTRIM) != 0;       // GemmParams.tap_tail_k / a2_tail_k honoured (the issuer loop of every other kernel stays as it was)
  static_assert(!ACC2 || BN == 256, "two accumulators per tile: 2 x 256 TMEM columns");
  static_assert(!TF32 || (VAR & (VAR_ACC2 | VAR_KTRIM | VAR_A2 | VAR_DUAL | VAR_DROPOUT | VAR_STAGED)) == 0, "kind::tf32 exists for the plain variants");
  constexpr bool KSPLIT = (VAR & VAR_KSPLIT) != 0;
  static_assert(!KSPLIT || (EPI == EPI_F32 && (VAR & (VAR_ACC2 | VAR_KTRIM | VAR_A2)) == 0), "split contraction: fp32 partial products only");
  const bool BALANCED = p.balanced != 0 && EPI != EPI_HEAD_LOSS;      // the loss partials are indexed by the round-robin tile grid
  using L = TnSmem<BN, STAGES, CG, STAGED>;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static_assert(2 * BN <= 512, "two accumulator buffers must fit the 512 TMEM columns");
  constexpr bool SAVED_IN = (EPI == EPI_DGRAD || EPI == EPI_BIAS_ADD) && !TF32;      // (fp32 storage: loaded inside the epilogue step)
  constexpr bool USE_BIAS = (EPI == EPI_BIAS_ACT || EPI == EPI_HEAD_LOSS || EPI == EPI_HEAD_OUT || EPI == EPI_BIAS_ADD);
  constexpr int QCOLS = BN / 4, NCH = QCOLS / 32;      // columns / 32-column steps per epilogue warp
  static_assert(NCH >= 1, "BN must be at least 128");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B atoms are 1024 B aligned
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  float* sbias = reinterpret_cast<float*>(smem_gen + L::BIAS_OFFSET);
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + L::BAR_OFFSET + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m_blocks = ((p.M + BM - 1) / BM + CG - 1) / CG;       // in units of CG m-blocks (pairs when CG == 2)
  const int num_n_blocks = (p.N + BN - 1) / BN;
  const int num_kb = p.K / BKE;
  // KSPLIT: the tile grid is (split, m-group, n-block); split s covers the k-blocks [s * kps, (s + 1) * kps)
  const int m_groups_real = num_m_blocks;
  const int kps = KSPLIT ? (num_kb + p.k_splits - 1) / p.k_splits : num_kb;
  const int tile_m_groups = KSPLIT ? num_m_blocks * p.k_splits : num_m_blocks;
  const int my_group = (int)(blockIdx.x / CG), num_groups = (int)(gridDim.x / CG);      // both CTAs of a pair walk the same tiles
  using Tiles = TileIter<BN>;

  if (warp == TN_PRODUCER_WARP && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if constexpr (A2) tma_prefetch_desc(&tmap_a2);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    // the leader's "accumulator drained" barrier collects the epilogue warps of BOTH CTAs
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), TN_EPI_WARPS * CG); }
    fence_mbar_init();
  }
  if (warp == TN_MMA_WARP) {
    if constexpr (CG == 2) { tmem_alloc_2sm(smem_u32(tmem_slot), TMEM_COLS); tmem_relinquish_2sm(); }
    else { tmem_alloc(smem_u32(tmem_slot), TMEM_COLS); tmem_relinquish(); }
  }
  pdl_launch_dependents();                           // the next kernel may begin its own prologue as SMs free up
  pdl_wait();                                        // everything below reads memory the previous kernel may have written
  const float* sloss_w = sbias + num_n_blocks * BN;   // EPI_HEAD_LOSS: loss weights behind the bias vector (host checks 2 N <= TN_BIAS_SMEM)
  if constexpr (USE_BIAS) {                          // whole bias vector, zero-extended to the tile grid (host checks N <= TN_BIAS_SMEM)
    for (int i = threadIdx.x; i < num_n_blocks * BN; i += TN_THREADS) {
      sbias[i] = (i < p.N) ? __ldg(p.bias + i) : 0.f;
      if constexpr (EPI == EPI_HEAD_LOSS) sbias[num_n_blocks * BN + i] = (i < p.N) ? __ldg(p.loss_w + i) : 0.f;
      if constexpr (ACC2) sbias[num_n_blocks * BN + i] = (i < p.N) ? __ldg(p.bias2 + i) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();       // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (STATS && p.stats != nullptr && threadIdx.x == 32) {
    p.stats[8 * blockIdx.x + 0] = (unsigned long long)clock64();
    p.stats[8 * blockIdx.x + 2] = globaltimer_ns();
  }

  if (warp == TN_PRODUCER_WARP) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (Tiles it(BALANCED, tile_m_groups, p.N, my_group, num_groups); it.next();) {
        const int split = KSPLIT ? it.mg / m_groups_real : 0;
        const int mg = KSPLIT ? it.mg - split * m_groups_real : it.mg;
        const int kb0 = KSPLIT ? split * kps : 0, kb1 = KSPLIT ? min(num_kb, kb0 + kps) : num_kb;
        const int m0 = (mg * CG + (int)cta_rank) * BM, n0 = it.n0;
        const int nb0 = n0 + (int)cta_rank * (it.n_valid / CG);              // this CTA's share of the B rows (the box may over-fetch)
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t sa = smem_base + s * L::STAGE_BYTES;
          const int tap = p.kb_per_tap ? kb / p.kb_per_tap : 0;
          int ka = (kb - tap * p.kb_per_tap) * BKE;               // column block inside the (un-replicated) A matrix
          int tap_shift = p.kb_per_tap ? tap - p.tap_center : 0;  // may be -1 at the top: TMA zero-fills out-of-range rows
          const CUtensorMap* ta = &tmap_a;
          if constexpr (A2) {
            if (kb >= p.a2_from_kb) { ta = &tmap_a2; ka = (kb - p.a2_from_kb) * BKE; tap_shift = 0; }
          }
          const bool ld_a = !(p.dbg & (16 | 128)), ld_b = !(p.dbg & (16 | 64));      // both true in production
          const uint32_t tx = (ld_a ? (uint32_t)L::A_BYTES : 0u) + (ld_b ? (uint32_t)(p.b_box_rows * 128) : 0u);
          if constexpr (CG == 2) {
            // one expect_tx on the leader's barrier covers the four loads of the pair
            if (is_leader) mbar_expect_tx(full_bar(s), 2 * tx);
            if (ld_a) tma_load_2d_2sm(sa, ta, full_bar(s), ka, m0 + tap_shift);
            if (ld_b) tma_load_2d_2sm(sa + L::A_BYTES, &tmap_b, full_bar(s), kb * BKE, nb0);
          } else {
            mbar_expect_tx(full_bar(s), tx);
            if (ld_a) tma_load_2d(sa, ta, full_bar(s), ka, m0 + tap_shift);
            if (ld_b) tma_load_2d(sa + L::A_BYTES, &tmap_b, full_bar(s), kb * BKE, nb0);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == TN_MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0 && is_leader) {
      int s = 0; uint32_t ph = 0; int t = 0;
      long long st_tempty = 0, st_full = 0;          // micro-benchmark: cycles the issuer waited for a free accumulator / for operands
      for (Tiles it(BALANCED, tile_m_groups, p.N, my_group, num_groups); it.next(); ++t) {
        const int n_valid = it.n_valid;
        const int kb0 = KSPLIT ? (it.mg / m_groups_real) * kps : 0, kb1 = KSPLIT ? min(num_kb, kb0 + kps) : num_kb;
        const uint32_t idesc = TF32 ? make_idesc_tf32(BM * CG, n_valid, 0, 0) : make_idesc_bf16(BM * CG, n_valid, 0, 0);
        const int acc = ACC2 ? 0 : (t & 1);          // ACC2: one tile in flight (both accumulators belong to it)
        const uint32_t acc_par = ACC2 ? ((uint32_t)t & 1u) : ((uint32_t)(t >> 1) & 1u);
        long long c0 = (STATS && p.stats) ? clock64() : 0;
        mbar_wait(tempty_bar(acc), acc_par ^ 1u);
        if (STATS && p.stats) st_tempty += clock64() - c0;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        [[maybe_unused]] int kin = 0;                // position inside the tap (tap_tail_k)
        for (int kb = kb0; kb < kb1; ++kb) {
          [[maybe_unused]] int nk = BK / UMMA_K;
          if constexpr (KTRIM) {
            if (A2 && kb >= p.a2_from_kb) { if (kb == num_kb - 1 && p.a2_tail_k > 0) nk = p.a2_tail_k; }
            else if (p.tap_tail_k > 0 && ++kin == p.kb_per_tap) { kin = 0; nk = p.tap_tail_k; }
          }
          c0 = (STATS && p.stats) ? clock64() : 0;
          mbar_wait(full_bar(s), ph);
          if (STATS && p.stats) st_full += clock64() - c0;
          tc_fence_after();
          const uint32_t sa = smem_base + s * L::STAGE_BYTES;
          const uint64_t da = make_desc_kmajor_sw128(sa);
          const uint64_t db = make_desc_kmajor_sw128(sa + L::A_BYTES);
          if constexpr (ACC2 || KTRIM) {
            uint32_t d_cols = tmem_d;
            int kfirst = kb;                         // 0 in the block that starts an accumulation
            if constexpr (ACC2) {
              if (kb >= p.a2_from_kb) { d_cols = tmem_d + (uint32_t)BN; kfirst = kb - p.a2_from_kb; }
            }
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              if (k >= nk) break;
              if constexpr (CG == 2) umma_f16_2sm(d_cols, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kfirst | k) != 0));
              else umma_f16(d_cols, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)((kfirst | k) != 0));
            }
          } else {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            if (p.dbg & 32) break;
            // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in the (addr >> 4) field  (tf32: 8 elements = 32 B)
            if constexpr (TF32) {
              if constexpr (CG == 2) umma_tf32_2sm(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)(((kb - kb0) | k) != 0));
              else umma_tf32(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)(((kb - kb0) | k) != 0));
            } else {
            if constexpr (CG == 2) umma_f16_2sm(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)(((kb - kb0) | k) != 0));
            else umma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)(((kb - kb0) | k) != 0));
            }
          }
          }
          // smem slot reusable (in both CTAs of a pair) once these MMAs have read it
          if constexpr (CG == 2) umma_commit_2sm(empty_bar(s)); else umma_commit(empty_bar(s));
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        // accumulator complete -> epilogue (of both CTAs of a pair)
        if constexpr (CG == 2) umma_commit_2sm(tfull_bar(acc)); else umma_commit(tfull_bar(acc));
      }
      if (STATS && p.stats) { p.stats[8 * blockIdx.x + 4] = (unsigned long long)st_tempty; p.stats[8 * blockIdx.x + 5] = (unsigned long long)st_full; }
    }
  } else {
    // ===================== epilogue: warp w -> TMEM lanes 32*(w&3).., columns [QCOLS*(w>>2), QCOLS*(w>>2)+QCOLS) ==========
    const int ew = warp - TN_FIRST_EPI_WARP;                 // 0..15
    const int q = warp & 3, cq = ew >> 2;                    // TMEM lane quadrant = hardware warp id % 4; column quarter
    const int tile_row = q * 32 + lane;
    const uint32_t stage_warp = smem_base + (uint32_t)(L::OUT_OFFSET + ew * L::OUT_WARP_BYTES);     // STAGED only
    int t = 0;
    long long st_epi_wait = 0, st_epi_busy = 0;      // micro-benchmark (epilogue warp 0): cycles waiting for an accumulator / working on it
    for (Tiles it(BALANCED, tile_m_groups, p.N, my_group, num_groups); it.next(); ++t) {
      const int split = KSPLIT ? it.mg / m_groups_real : 0;
      const int mb = (KSPLIT ? it.mg - split * m_groups_real : it.mg) * CG + (int)cta_rank;
      const int m0 = mb * BM, n0 = it.n0;
      const int n_valid = it.n_valid;
      const int acc = ACC2 ? 0 : (t & 1);
      const uint32_t acc_par = ACC2 ? ((uint32_t)t & 1u) : ((uint32_t)(t >> 1) & 1u);
      const RowInfo ri = make_row_info(p, m0 + tile_row);
      const int c0 = cq * QCOLS;                              // tile-relative first column of this warp (warp-uniform)
      // everything that does not depend on the accumulator is requested before waiting for it
      // (arrays an epilogue kind does not use are left uninitialised on purpose: zero-filling them cost 32 register-pair moves per
      // tile and warp in every kind -- 5 % of the forward epilogue's instructions)
      uint32_t mask_words[NCH];
      if constexpr (EPI == EPI_DGRAD_MASK) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) mask_words[i] = 0u;
        if (ri.in_range) {
#pragma unroll
          for (int i = 0; i < NCH; ++i)
            if (c0 + 32 * i < n_valid) mask_words[i] = __ldg(p.mask_in + (size_t)((n0 + c0 + 32 * i) >> 5) * p.ld_mask + ri.grow);
        }
      }
      uint32_t sv[NCH][16];
      if constexpr (SAVED_IN) {
#pragma unroll
        for (int i = 0; i < NCH; ++i)
#pragma unroll
          for (int j = 0; j < 16; ++j) sv[i][j] = 0u;
        if (ri.in_range) {
#pragma unroll
          for (int i = 0; i < NCH; ++i)
            if (c0 + 32 * i < n_valid) load_bf16x32_global_raw(p.saved + (size_t)ri.grow * p.ld_saved + n0 + c0 + 32 * i, sv[i]);
        }
      }
      // EPI_HEAD_LOSS: the 32 targets of the first step, requested before the accumulator wait (later steps: before their tcgen05.ld)
      float yv[32];
      auto load_targets = [&](int c) {
        if constexpr (EPI == EPI_HEAD_LOSS) {
#pragma unroll
          for (int j = 0; j < 32; ++j) yv[j] = 0.f;
          if (ri.row_ok && c < n_valid) {
            const int gc = n0 + c;
            const float* yptr = p.y + (size_t)ri.yrow * p.ld_y + gc;
            if (gc + 32 <= p.out_dim && (p.ld_y & 7) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 31) == 0) {
              load_f32x32_global_v8(yptr, yv);
            } else {
              for (int j = 0; j < 32; ++j)
                if (gc + j < p.out_dim) yv[j] = __ldg(yptr + j);
            }
          }
        }
      };
      load_targets(c0);
      long long ck0 = (STATS && p.stats) ? clock64() : 0;
      mbar_wait_long(tfull_bar(acc), acc_par);
      long long ck1 = (STATS && p.stats) ? clock64() : 0;
      tc_fence_after();
      float loss_acc = 0.f;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0);
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int c = c0 + 32 * i;                            // tile-relative column of this step (warp-uniform)
        if (c < n_valid) {
          if (i > 0) load_targets(c);
          uint32_t raw[32];
          if (!(p.dbg & 8)) {
            tmem_ld_32x32(taddr + (uint32_t)(32 * i), raw);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) raw[j] = 0u;
          }
          if (!(p.dbg & 2))
            epilogue_chunk<EPI, ELU, GENERAL_LOSS, STAGED, DROPOUT, DUAL, ACC2, TF32>(p, sbias, sloss_w, ri, n0 + c, raw, sv[i], yv, loss_acc, mask_words[i],
                                                           stage_warp + (uint32_t)(lane * L::OUT_PITCH + 64 * i),
                                                           taddr + (uint32_t)(BN + 32 * i), KSPLIT ? (size_t)split * p.split_stride : (size_t)0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                                        // TMEM buffer free: the MMA of tile t+2 may start
        if constexpr (CG == 2) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc));
      }
      if constexpr (STAGED) {
        // the warp's 32 x QCOLS bf16 tile, shared -> global: each instruction moves 16 B per lane, a row's lanes adjacent, so
        // a store instruction covers 4 (QCOLS = 64) or 8 (QCOLS = 32) whole row segments instead of 32 separate 32-byte pieces
        if (c0 < n_valid && !(p.dbg & (2 | 4))) {
          constexpr int CPR = QCOLS * 2 / 16, RPI = 32 / CPR;     // 16-byte chunks per row, rows per instruction
          const int rr = lane / CPR, ch = lane % CPR;
          __nv_bfloat16* gbase = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)(m0 + q * 32) * p.ld_out + n0 + c0 + 8 * ch;
#pragma unroll
          for (int j = 0; j < 32 / RPI; ++j) {
            const int r = j * RPI + rr;
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                         : "r"(stage_warp + (uint32_t)(r * L::OUT_PITCH + 16 * ch)) : "memory");
            if (m0 + q * 32 + r < p.M)
              asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(gbase + (size_t)r * p.ld_out), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
          }
        }
        __syncwarp();                                         // the tile is read before the next tile's rows overwrite it
      }
      if (STATS && p.stats && ew == 0 && lane == 0) { st_epi_wait += ck1 - ck0; st_epi_busy += clock64() - ck1; }
      if constexpr (EPI == EPI_HEAD_LOSS) {
        // deterministic: one partial per (m-block, n-block, epilogue warp)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
        if (lane == 0 && m0 < p.M) p.loss_partials[(mb * num_n_blocks + it.slot) * TN_EPI_WARPS + ew] = loss_acc * p.grad_scale;
      }
    }
    if (STATS && p.stats && ew == 0 && lane == 0) { p.stats[8 * blockIdx.x + 6] = (unsigned long long)st_epi_wait; p.stats[8 * blockIdx.x + 7] = (unsigned long long)st_epi_busy; }
  }

  tc_fence_before();
  __syncthreads();
  if (STATS && p.stats != nullptr && threadIdx.x == 32) {
    p.stats[8 * blockIdx.x + 1] = (unsigned long long)clock64();
    p.stats[8 * blockIdx.x + 3] = globaltimer_ns();
  }
  if constexpr (CG == 2) cluster_sync_all();       // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == TN_MMA_WARP) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_2sm(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// gemm_nt_kernel: D[M,N] = sum_r A[r, m] * B[r, n]; both operands MN-major; split over r (blockIdx.y, or NtParams.splits_narrow).
// Weight gradient dW = H^T dZ.  The epilogue warps, idle during the mainloop, reduce the dZ tiles that pass through shared
// memory over their rows (every m-tile takes its share of the row blocks): that is the bias gradient, at no extra HBM traffic.
// ---------------------------------------------------------------------------------------------------------------
struct NtParams {
  int M, N, R;             // M, N feature dims (multiples of 64), R rows to contract
  int rb_per_split;        // 64-row blocks per split
  // Uneven split counts (splits_narrow > 0; the grid is then one-dimensional): the tiles of a half-width last n-block do half
  // the work per row block, so they take `splits_narrow` = ceil(splits / 2) splits of `rb_per_split_narrow` row blocks -- as long
  // as a full-width tile's split -- and the CTAs saved go into a larger `splits` for everybody.  A narrow tile's CTA also writes
  // the zeros of the split slots it leaves unused, so that the reduction keeps one slot count per layer.
  int splits, splits_narrow, rb_per_split_narrow;
  float* out;              // partials: out + split * split_stride + m * ld_out + n
  int ld_out;
  size_t split_stride;
  float* colsum_out;       // optional: column sums of B (bias gradient) partials: colsum_out + split * colsum_stride + n
  size_t colsum_stride;
  int a_row_offset;        // Conv1D weight gradient of tap t: A rows shifted by (t - center); out-of-range rows read as zero
  // All taps of a Conv1D in ONE launch: the output rows are stacked tap-major (M = taps * a_tap_m, exactly the layout of the
  // [taps][Cin][Cout] kernel), and 64-column chunk c of the stacked A^T belongs to tap t = c / (a_tap_m / 64): it is loaded from
  // columns (c mod a_tap_m / 64) * 64 of the activation matrix with its rows shifted by (t - a_tap_center).  0 = off.
  int a_tap_m, a_tap_center;
};

template <int BN, int STAGES, int CG = 1>
struct NtSmem {
  static constexpr int A_BYTES = BK * BM * 2;          // 64 contraction rows x 128 m (2 chunks of [64 rows x 128 B])
  static constexpr int B_BYTES = BK * (BN / CG) * 2;   // BN/64 chunks; a CTA pair holds half of the columns each
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
  static_assert(TOTAL <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
};

// CG = 1: one 128 x BN tile per CTA.
// CG = 2 (clusters of two): a CTA pair owns a 256 x BN tile with tcgen05 cta_group::2 -- CTA r loads its 128 rows of the
// m-range and HALF of the dZ columns, the leader issues the MMAs for both.  Per SM that is 32 KB of operands per 64-row block
// and 512 MMA cycles (64 B/cycle) instead of 48 KB (96 B/cycle), which is more than an SM can take in from L2 (~69 B/cycle
// measured, profiles/r01_probe_mainloop.txt).  Each CTA fills its own shared memory through its own `full` barrier (its
// bias-gradient warps read the dZ chunks from there); a relay lane in the peer forwards "my stage is full" to the leader.
template <int BN, int STAGES, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const NtParams p) {
  using L = NtSmem<BN, STAGES, CG>;
  constexpr uint32_t TMEM_COLS = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : (BN <= 256) ? 256 : 512;
  constexpr int CHUNK_BYTES = BK * 128;   // one [64 rows x 64 elements] box = 8 KB
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto peer_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };     // leader only: the peer's stage s is full
  const uint32_t tfull_bar = bar_base + 8u * (3 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + L::BAR_OFFSET + 8 * (3 * STAGES + 1));
  static_assert(8 * (3 * STAGES + 2) <= 256, "barrier area");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_n_blocks = (p.N + BN - 1) / BN;
  // bias gradient: every m-tile takes the row blocks i with i % num_m_tiles == its index, so no CTA is a straggler;
  // partial index = split * num_m_tiles + m-tile
  const int num_m_tiles = ((p.M + BM - 1) / BM + CG - 1) / CG;
  // (m-tile, n-block, split) of this CTA (both CTAs of a pair share them).  Uniform mode: blockIdx.x = tile, blockIdx.y = split.
  // Uneven mode: all (split, full-width tile) pairs first, split-major, then the (split, narrow tile) pairs.
  int m_tile, n_blk, split;
  bool narrow_tile = false;
  if (p.splits_narrow > 0) {
    const int lin = (int)(blockIdx.x / CG), tiles_wide = num_m_tiles * (num_n_blocks - 1), wide_ctas = tiles_wide * p.splits;
    if (lin < wide_ctas) {
      split = lin / tiles_wide;
      const int tw = lin - split * tiles_wide;
      m_tile = tw / (num_n_blocks - 1); n_blk = tw - m_tile * (num_n_blocks - 1);
    } else {
      const int ln = lin - wide_ctas;
      split = ln / num_m_tiles; m_tile = ln - split * num_m_tiles; n_blk = num_n_blocks - 1;
      narrow_tile = true;
    }
  } else {
    const int tile = (int)(blockIdx.x / CG);
    m_tile = tile / num_n_blocks; n_blk = tile - m_tile * num_n_blocks; split = (int)blockIdx.y;
  }
  const int m0 = (m_tile * CG + (int)cta_rank) * BM, n0 = n_blk * BN;
  const int n_valid = min(BN, p.N - n0);
  const int nb0 = n0 + (int)cta_rank * (n_valid / CG);                   // first dZ column this CTA loads
  // CG == 2 always loads both 64-wide m chunks (columns past M are zero-filled by TMA) so that the transaction size is uniform
  const int a_chunks = (CG == 2) ? 2 : (min(BM, p.M - m0) + 63) / 64, b_chunks = (n_valid / CG + 63) / 64;
  const int num_rb = (p.R + BK - 1) / BK;
  const int rps = narrow_tile ? p.rb_per_split_narrow : p.rb_per_split;
  const int rb_begin = min(num_rb, split * rps);
  const int rb_end = min(num_rb, rb_begin + rps);
  const int nrb = max(0, rb_end - rb_begin);
  const bool do_colsum = p.colsum_out != nullptr;                       // kernel-uniform
  const int colsum_warps = do_colsum ? min(4, b_chunks) : 0;            // warp w sums the columns of this CTA's chunk w

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1 + colsum_warps);
      mbar_init(peer_full_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    if constexpr (CG == 2) { tmem_alloc_2sm(smem_u32(tmem_slot), TMEM_COLS); tmem_relinquish_2sm(); }
    else { tmem_alloc(smem_u32(tmem_slot), TMEM_COLS); tmem_relinquish(); }
  }
  pdl_launch_dependents();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int rb = rb_begin; rb < rb_end; ++rb) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), (uint32_t)((a_chunks + b_chunks) * CHUNK_BYTES));
        const uint32_t sa = smem_base + s * L::STAGE_BYTES;
        for (int c = 0; c < a_chunks; ++c) {
          int mcol = m0 + 64 * c, roff = p.a_row_offset;
          if (p.a_tap_m > 0) {                       // stacked taps: chunk -> (tap, column inside the activation matrix)
            const int t = mcol / p.a_tap_m;
            roff = t - p.a_tap_center;
            mcol = (mcol < p.M) ? mcol - t * p.a_tap_m : p.a_tap_m;      // past the last tap: an out-of-range column, zero-filled by TMA
          }
          tma_load_2d(sa + c * CHUNK_BYTES, &tmap_a, full_bar(s), mcol, rb * BK + roff);
        }
        for (int c = 0; c < b_chunks; ++c) tma_load_2d(sa + L::A_BYTES + c * CHUNK_BYTES, &tmap_b, full_bar(s), nb0 + 64 * c, rb * BK);
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && nrb > 0) {
      int s = 0; uint32_t ph = 0;
      if (is_leader) {
        // M is always issued in full: rows >= M read zero-filled / stale smem and are never stored by the epilogue
        const uint32_t idesc = make_idesc_bf16(BM * CG, n_valid, 1, 1);
        for (int i = 0; i < nrb; ++i) {
          mbar_wait(full_bar(s), ph);
          if constexpr (CG == 2) mbar_wait(peer_full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * L::STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // 16 contraction rows = 2 groups of 8 rows x 128 B = 2048 B further into every chunk
            const uint64_t da = make_desc_mnmajor_sw128(sa + k * 2048, CHUNK_BYTES);
            const uint64_t db = make_desc_mnmajor_sw128(sa + L::A_BYTES + k * 2048, CHUNK_BYTES);
            if constexpr (CG == 2) umma_f16_2sm(tmem_base, da, db, idesc, (uint32_t)((i | k) != 0));
            else umma_f16(tmem_base, da, db, idesc, (uint32_t)((i | k) != 0));
          }
          if constexpr (CG == 2) umma_commit_2sm(empty_bar(s)); else umma_commit(empty_bar(s));
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        if constexpr (CG == 2) umma_commit_2sm(tfull_bar); else umma_commit(tfull_bar);
      } else {
        // peer of a pair: relay "stage s of my shared memory is full" to the leader's MMA issuer
        for (int i = 0; i < nrb; ++i) {
          mbar_wait(full_bar(s), ph);
          mbar_arrive_leader(peer_full_bar(s));
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // ---- bias gradient: column sums of the dZ tiles streaming through the ring (warps 0..colsum_warps-1)
    if (warp < colsum_warps) {
      // lane -> logical 16-byte piece (8 columns) lp = lane & 7 of row 4*i + (lane >> 3); a quarter-warp reads one full
      // 128-byte row (conflict-free); the physical piece position is XOR-swizzled with (row & 7)
      float acc8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const int lp = lane & 7, rs = lane >> 3;
      int s = 0; uint32_t ph = 0;
      for (int i = 0; i < nrb; ++i) {
        mbar_wait(full_bar(s), ph);
        const uint8_t* chunk = smem_gen + s * L::STAGE_BYTES + L::A_BYTES + warp * CHUNK_BYTES;
        if (i % num_m_tiles == m_tile)
#pragma unroll
        for (int it = 0; it < BK / 4; ++it) {         // rows past R were zero-filled by TMA
          const int r = 4 * it + rs;
          const uint4 w = *reinterpret_cast<const uint4*>(chunk + r * 128 + ((lp ^ (r & 7)) << 4));
          acc8[0] += bf16_lo(w.x); acc8[1] += bf16_hi(w.x); acc8[2] += bf16_lo(w.y); acc8[3] += bf16_hi(w.y);
          acc8[4] += bf16_lo(w.z); acc8[5] += bf16_hi(w.z); acc8[6] += bf16_lo(w.w); acc8[7] += bf16_hi(w.w);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));     // this warp is done reading the slot
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {                   // fold the four row-lanes that share a piece
        acc8[j] += __shfl_xor_sync(0xffffffffu, acc8[j], 8);
        acc8[j] += __shfl_xor_sync(0xffffffffu, acc8[j], 16);
      }
      // (a CTA's share of the columns need not be a whole number of 64-column chunks -- n_valid = 160 on a pair gives 80 -- so the
      // last chunk may hold over-fetched columns of the neighbour: only this CTA's own 8-column pieces are written)
      if (rs == 0 && 64 * warp + 8 * lp < n_valid / CG) {
        float* dst = p.colsum_out + ((size_t)split * num_m_tiles + m_tile) * p.colsum_stride + nb0 + 64 * warp + 8 * lp;
        *reinterpret_cast<float4*>(dst) = make_float4(acc8[0], acc8[1], acc8[2], acc8[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(acc8[4], acc8[5], acc8[6], acc8[7]);
        if (narrow_tile) {                            // the split slots this narrow tile leaves unused hold zeros
          for (int zs = split + p.splits_narrow; zs < p.splits; zs += p.splits_narrow) {
            float* z = p.colsum_out + ((size_t)zs * num_m_tiles + m_tile) * p.colsum_stride + nb0 + 64 * warp + 8 * lp;
            *reinterpret_cast<float4*>(z) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(z + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
    }
    // ---- weight-gradient tile
    const int row = m0 + warp * 32 + lane;
    float* out = p.out + (size_t)split * p.split_stride + (size_t)row * p.ld_out + n0;
    if (nrb > 0) {
      mbar_wait_long(tfull_bar, 0);
      tc_fence_after();
      // Every MMA has completed, so the operand ring is free: each thread parks its output row (n_valid fp32) there and hands
      // it to the bulk-copy engine as ONE contiguous shared -> global copy.  (Storing straight from registers makes every
      // store instruction touch 32 different lines, ~16 B/clk per SM: 4 us for the 128 KB tile of a CTA, with nothing to
      // overlap it in this one-tile-per-CTA kernel.)  Row pitch n_valid * 4 + 16 B keeps the 16-byte shared stores conflict-free.
      // (the other epilogue warps may still be column-summing the last ring slot: all four pass this barrier before any writes)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
      const uint32_t pitch = (uint32_t)n_valid * 4u + 16u;
      const uint32_t srow = smem_base + (uint32_t)(warp * 32 + lane) * pitch;
      static_assert(BM * (BN * 4 + 16) <= L::BAR_OFFSET, "output staging tile must fit the operand ring");
      for (int c = 0; c < n_valid; c += 32) {
        uint32_t raw[32];
        tmem_ld_32x32(taddr + (uint32_t)c, raw);
        tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4)
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(srow + (uint32_t)(c + 4 * q4) * 4u), "r"(raw[4 * q4]), "r"(raw[4 * q4 + 1]),
                       "r"(raw[4 * q4 + 2]), "r"(raw[4 * q4 + 3]) : "memory");
      }
      fence_proxy_async_smem();
      if (row < p.M) bulk_store_s2g(out, srow, (uint32_t)n_valid * 4u);
      bulk_commit();
      bulk_wait_read_all();                         // shared memory must stay intact until the copy engine has read it
    } else if (row < p.M) {
      for (int c = 0; c < n_valid; c += 4) *reinterpret_cast<float4*>(out + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (narrow_tile) {                               // zero partials for the unused split slots of this narrow tile, a row per instruction
      for (int zs = split + p.splits_narrow; zs < p.splits; zs += p.splits_narrow) {
        float* z = p.out + (size_t)zs * p.split_stride + (size_t)(m0 + warp * 32) * p.ld_out + n0;
        for (int r = 0; r < 32; ++r) {
          if (m0 + warp * 32 + r >= p.M) break;
          for (int c = 4 * lane; c < n_valid; c += 128) *reinterpret_cast<float4*>(z + (size_t)r * p.ld_out + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();       // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == 5) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_2sm(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace csb
