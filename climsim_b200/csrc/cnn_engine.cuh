// cnn_engine.cuh -- the csb_cnn_* C ABI: ResNet-1D column emulator of baseline_models/CNN/training/hpo_train.py:131-200
// (included at the end of mlp_engine.cu: shares its TMA / launch / profiling helpers).
//
// Layout: channels-last with one zero halo row above and below every column: sample b owns rows b*62 .. b*62+61 of a
// [B*62, Cp] bf16 matrix (Cp = channels padded to 64).  Conv1D(k=3,'same') is then ONE GEMM whose contraction runs over
// (tap, channel): contraction block kb reads the activation rows shifted by (tap - 1) -- the TMA producer of
// gemm_tn_kernel does the shift, out-of-range rows are zero-filled by TMA, and the epilogue writes zeros into the halo
// rows so that they keep acting as 'same' padding for the next layer.  1x1 convolutions and the per-level Dense heads
// are plain GEMMs over the same rows.  Residual `x = relu(conv2) + conv1x1(block input)` = EPI_BIAS_ADD (the relu(conv2)
// tile arrives by TMA into the staging tile and is added in place).
//
// Backward: data gradients are convolutions with the tap-flipped, transposed kernels (wd16 copies); weight gradients
// are one gemm_nt launch per tap with the activation rows shifted by (tap - 1); bias gradients ride in the tap-0 launch.
// Per residual block, d(block input) = conv1^T(dz1) + conv1x1^T(d_out) is ONE launch over [dz1 taps | d_out] (second A tensor
// map) whose second output is the previous block's dz2 (cnn_dgrad_dual); the dropout of the forward pass sits in the conv
// epilogues, and the MMAs over the all-zero tail of a 406 -> 448 channel block are skipped (cnn_tail_k).
// CSB_F32 parity mode: the same flow on fp32 buffers with sgemm_kernel (tap-aware loaders), no split-K workspace.
#pragma once

struct ConvLayerInfo {
  int taps, Cin, Cout, Cinp, Coutp, act;
  size_t w_off, b_off, w_off_user, b_off_user;
  size_t ws_w_off, ws_b_off;
  int max_splits;
  __nv_bfloat16 *wt16 = nullptr, *wd16 = nullptr;
  // pitch / first column of the copies: the residual 1x1 layer's copies live behind the k = 3 layer they are merged with
  // (wd16 behind conv1's, wt16 behind conv2's), as one more tap of the same B operand
  int wt_ld = 0, wt_col0 = 0, wd_ld = 0, wd_col0 = 0;
  bool wt_alias = false, wd_alias = false;     // the storage belongs to another layer
  CUtensorMap tm_wt, tm_wd;
  CUtensorMap tm_wtc, tm_wdc;                  // the merged operands (on the k = 3 layers that host a 1x1 layer's copy)
  bool has_wtc = false, has_wdc = false;
};

struct BufMaps { CUtensorMap k128, mn64; };
struct CnnBuf { void* ptr = nullptr; int Cp = 0; BufMaps maps; };      // bf16 (CSB_BF16) or fp32 (CSB_F32) [cap x Cp]

struct csb_cnn {
  csb_cnn_cfg cfg;
  int depth = 0, L = 60, P = 62;                 // levels, rows per sample incl. halo
  int n_layers = 0;                              // 3*depth + 2
  ConvLayerInfo layer[3 * 16 + 2];
  int in_ch = 6, in_p = 64, width = 0, width_p = 0, out_ch = 10, out_p = 64, out_lin = 2;
  int64_t cap = 0;                               // rows capacity = max_batch * P rounded to 128
  size_t P_pad = 0, P_user = 0, ws_elems = 0;
  int sm_count = 0;
  float *params = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr, *ws = nullptr, *zero_bias = nullptr;
  // activation buffers: x0, per block (h1, h2, out), e;  gradient buffers G0, G1, Z1, Z2, T, dzh (head), dze
  bool bf16 = true;
  CnnBuf x0, h1[16], h2[16], ob[16], e, G[2], Z1, Z2, T, dzh, dze;
  float* pred = nullptr;                         // fp32 mode: head output in the halo layout [cap x out_p]
  float *d_loss_w = nullptr, *loss_partials = nullptr, *d_loss = nullptr;
  int n_loss_partials = 0;
  int64_t maps_B = -1;
  int64_t step = 0, launches = 0;
  float dropout = 0.f;                           // Dropout rate behind the two ReLUs of every block (training steps only)
  uint32_t drop_seed = 0;
  int64_t train_steps = 0;                       // advances the dropout masks from step to step
  int64_t fwd_B = -1;                            // batch of the last csb_cnn_forward (its activations are what csb_cnn_backward uses)
  // launch merging (read from the environment when the handle is created; A/B aids):
  bool merge_bwd = true;                         // CSB_CNN_NO_MERGE=1 off: d(block input) + the previous block's dz2 as one GEMM launch
  bool merge_fwd = false;                        // CSB_CNN_MERGE_FWD=1 on: conv2 + residual 1x1 as one launch (measured: no gain)
};
static bool g_cnn_ktrim = true;                  // CSB_CNN_NO_KTRIM=1 off

static void cnn_free(csb_cnn* h) {
  auto F = [](void* p) { if (p) cudaFree(p); };
  F(h->params); F(h->grads); F(h->m); F(h->v); F(h->ws); F(h->zero_bias);
  for (int l = 0; l < h->n_layers; ++l) {
    if (!h->layer[l].wt_alias) F(h->layer[l].wt16);
    if (!h->layer[l].wd_alias) F(h->layer[l].wd16);
  }
  F(h->x0.ptr); F(h->e.ptr); F(h->G[0].ptr); F(h->G[1].ptr); F(h->Z1.ptr); F(h->Z2.ptr); F(h->T.ptr); F(h->dzh.ptr); F(h->dze.ptr);
  for (int i = 0; i < 16; ++i) { F(h->h1[i].ptr); F(h->h2[i].ptr); F(h->ob[i].ptr); }
  F(h->pred); F(h->d_loss_w); F(h->loss_partials); F(h->d_loss);
}

static int cnn_repack(csb_cnn* h, cudaStream_t st) {
  if (!h->bf16) return CSB_OK;
  simt::ConvRepackTable tab;
  tab.n = h->n_layers;
  int64_t mx = 1;
  for (int l = 0; l < h->n_layers; ++l) {
    ConvLayerInfo& li = h->layer[l];
    tab.l[l] = {h->params + li.w_off, li.wt16, li.wd16, li.taps, li.Cinp, li.Coutp, li.wt_ld, li.wt_col0, li.wd_ld, li.wd_col0};
    mx = std::max<int64_t>(mx, (int64_t)li.taps * li.Cinp * li.Coutp);
  }
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(mx, 256), 4 * h->sm_count), (unsigned)h->n_layers);
  simt::conv_repack_kernel<<<grid, 256, 0, st>>>(tab);
  CSB_CUDA_CHECK(cudaGetLastError());
  h->launches++;
  return CSB_OK;
}

static int cnn_pad_copy(csb_cnn* h, float* padded, float* user, int dir, cudaStream_t st) {
  simt::ConvPadTable tab;
  tab.n = h->n_layers;
  int64_t mx = 1;
  for (int l = 0; l < h->n_layers; ++l) {
    const ConvLayerInfo& li = h->layer[l];
    tab.l[l] = {li.taps, li.Cin, li.Cout, li.Cinp, li.Coutp, li.w_off, li.b_off, li.w_off_user, li.b_off_user};
    mx = std::max<int64_t>(mx, (int64_t)li.taps * li.Cin * li.Cout + li.Cout);
  }
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(mx, 256), 4 * h->sm_count), (unsigned)h->n_layers);
  simt::conv_pad_copy_kernel<<<grid, 256, 0, st>>>(padded, user, dir, tab);
  CSB_CUDA_CHECK(cudaGetLastError());
  h->launches++;
  return CSB_OK;
}

static int cnn_buf_maps(CnnBuf* b, int64_t rows) {
  int rc = make_tmap_bf16(&b->maps.k128, b->ptr, b->Cp, rows, b->Cp, 64, 128);
  if (rc) return rc;
  return make_tmap_bf16(&b->maps.mn64, b->ptr, b->Cp, rows, b->Cp, 64, 64);
}

static int cnn_build_maps(csb_cnn* h, int64_t B) {
  if (!h->bf16 || h->maps_B == B) return CSB_OK;
  const int64_t R = B * h->P;
  int rc;
  CnnBuf* all[] = {&h->x0, &h->e, &h->G[0], &h->G[1], &h->Z1, &h->Z2, &h->T, &h->dzh, &h->dze};
  for (CnnBuf* b : all) if ((rc = cnn_buf_maps(b, R))) return rc;
  for (int i = 0; i < h->depth; ++i) {
    if ((rc = cnn_buf_maps(&h->h1[i], R))) return rc;
    if ((rc = cnn_buf_maps(&h->h2[i], R))) return rc;
    if ((rc = cnn_buf_maps(&h->ob[i], R))) return rc;
  }
  h->maps_B = B;
  return CSB_OK;
}

// Output width the tensor cores compute for a layer of `c` real channels stored with pitch `cp` (a multiple of 64): the last n-block
// stops at the next multiple of 32 (406 -> 416 instead of 448: 7 % fewer MMA columns).  The columns in between are never written
// and stay what the allocation made them, zero, which is what the next layer's contraction over the full pitch needs.  Only when the
// trimmed width still fills a 256-wide first block, so that the B tensor-map box (encoded for the padded width) is unchanged.
static inline int cnn_mma_width(int c, int cp) {
  const int w = (int)round_up(c, 32);
  return (cp > 256 && w >= 256 && w < cp) ? w : cp;
}

// MMAs needed for the last 64-channel contraction block of a tensor with `c` real channels stored with pitch `cp`: the channels behind
// c are zero, so the 16-channel instructions that would only see zeros are skipped (406 -> 2 of 4).  0 = all four.
static inline int cnn_tail_k(int c, int cp) {
  const int r = c - (cp - 64);
  return (!g_cnn_ktrim || r <= 0 || r > 48) ? 0 : (r + 15) / 16;
}

// one convolution-as-GEMM launch.  `dgrad` selects the flipped/transposed weights (output width = Cinp).
//   kind 0: out = act(conv + bias)   kind 1: out = conv + bias + saved   kind 2: out = conv * act'(saved)
static inline uint32_t cnn_drop_seed(const csb_cnn* h, int layer_id) {
  return h->drop_seed ^ (uint32_t)(h->train_steps * 0x9E3779B97F4A7C15ull >> 32) ^ (uint32_t)(layer_id + 1) * 0x85EBCA77u;
}
static inline uint32_t cnn_drop_threshold(const csb_cnn* h) { return (uint32_t)lrintf(h->dropout * 16777216.f); }   // keep iff 24 random bits >= rate * 2^24

// drop_layer >= 0 (kind 0, bf16): inverted dropout behind the activation inside the epilogue (VAR_DROPOUT), keyed like cnn_dropout
static int cnn_conv(csb_cnn* h, const ConvLayerInfo& li, bool dgrad, const CnnBuf& in, CnnBuf& out, int kind, const CnnBuf* saved, int act,
                    const float* bias, int64_t B, cudaStream_t st, float dgrad_scale = 0.f, int drop_layer = -1) {
  const int M = (int)(B * h->P), N = dgrad ? li.Cinp : li.Coutp, Kt = dgrad ? li.Coutp : li.Cinp;
  int rc = CSB_OK;
  if (h->bf16) {
    tc::GemmParams p = {};
    p.M = M; p.N = cnn_mma_width(dgrad ? li.Cin : li.Cout, N); p.K = li.taps * Kt; p.kb_per_tap = Kt / 64; p.tap_center = (li.taps - 1) / 2;
    p.halo_period = h->P;
    p.act = act; p.head_relu_from = -1; p.bias = bias; p.dgrad_scale = dgrad_scale;
    p.tap_tail_k = cnn_tail_k(dgrad ? li.Cout : li.Cin, Kt);
    const CUtensorMap& w = dgrad ? li.tm_wd : li.tm_wt;
    p.out = out.ptr; p.ld_out = out.Cp;
    if (saved) { p.saved = reinterpret_cast<const __nv_bfloat16*>(saved->ptr); p.ld_saved = saved->Cp; }
    // the instantiations that skip the all-zero MMAs of a tap's last block (VAR_KTRIM) exist for the wide pair tiles only
    const bool trim = p.tap_tail_k > 0 && g_use_pairs && p.N > 128 && act != CSB_ACT_ELU;
    if (!trim) p.tap_tail_k = 0;
    constexpr int KT = tc::VAR_KTRIM;
    if (kind == 0 && drop_layer >= 0 && act != CSB_ACT_ELU) {
      p.drop_seed = cnn_drop_seed(h, drop_layer); p.drop_threshold = cnn_drop_threshold(h); p.drop_scale = 1.f / (1.f - h->dropout);
      if (trim) rc = launch_tn<256, 6, tc::EPI_BIAS_ACT, 2, tc::VAR_DROPOUT | KT>(in.maps.k128, w, p, h->sm_count, st);
      else rc = launch_tn_shape<tc::EPI_BIAS_ACT, tc::VAR_DROPOUT>(in.maps.k128, w, p, h->sm_count, st);
    } else if (trim) {
      if (kind == 0) rc = launch_tn<256, 6, tc::EPI_BIAS_ACT, 2, KT>(in.maps.k128, w, p, h->sm_count, st);
      else if (kind == 1) rc = launch_tn<256, 6, tc::EPI_BIAS_ADD, 2, KT>(in.maps.k128, w, p, h->sm_count, st);
      else rc = launch_tn<256, 6, tc::EPI_DGRAD, 2, KT>(in.maps.k128, w, p, h->sm_count, st);
    } else if (kind == 0) rc = launch_tn_auto<tc::EPI_BIAS_ACT>(in.maps.k128, w, p, h->sm_count, st);
    else if (kind == 1) rc = launch_tn_auto<tc::EPI_BIAS_ADD>(in.maps.k128, w, p, h->sm_count, st);
    else rc = launch_tn_auto<tc::EPI_DGRAD>(in.maps.k128, w, p, h->sm_count, st);
  } else {
    simt::SgemmParams p = {};
    p.M = M; p.N = N; p.K = li.taps * Kt;
    p.A = reinterpret_cast<const float*>(in.ptr); p.lda = in.Cp;
    p.C = reinterpret_cast<float*>(out.ptr); p.ldc = out.Cp;
    p.bias = bias; p.act = act; p.head_relu_from = -1; p.halo_period = h->P;
    p.a_tap_k = Kt; p.tap_center = (li.taps - 1) / 2;
    if (saved) { p.saved = reinterpret_cast<const float*>(saved->ptr); p.ld_saved = saved->Cp; }
    dim3 grid((unsigned)(N / 64), (unsigned)ceil_div(M, 64));
    if (!dgrad) {
      p.B = h->params + li.w_off; p.ldb = li.Coutp;               // W [taps*Cinp, Coutp]
      if (kind == 0) simt::sgemm_kernel<false, false, simt::SEPI_BIAS_ACT><<<grid, 256, 0, st>>>(p);
      else if (kind == 1) simt::sgemm_kernel<false, false, simt::SEPI_BIAS_ADD><<<grid, 256, 0, st>>>(p);
      else simt::sgemm_kernel<false, false, simt::SEPI_DGRAD><<<grid, 256, 0, st>>>(p);
    } else {
      p.B = h->params + li.w_off; p.ldb = li.Coutp;               // W[t][ci][co] read as B'[ci][(taps-1-t', co)]
      p.b_tap_k = li.Coutp; p.b_tap_rows = li.Cinp; p.taps = li.taps;
      if (kind == 0) simt::sgemm_kernel<false, true, simt::SEPI_BIAS_ACT><<<grid, 256, 0, st>>>(p);
      else if (kind == 1) simt::sgemm_kernel<false, true, simt::SEPI_BIAS_ADD><<<grid, 256, 0, st>>>(p);
      else simt::sgemm_kernel<false, true, simt::SEPI_DGRAD><<<grid, 256, 0, st>>>(p);
    }
    if (cudaGetLastError() != cudaSuccess) { set_last_error("sgemm launch failed"); rc = CSB_ECUDA; }
  }
  if (rc) return rc;
  h->launches++;
  return CSB_OK;
}

// Second half of a residual block in ONE launch (bf16, CTA pairs): the residual 1x1 convolution is a fourth "tap" of conv2's GEMM,
// read from the block input (VAR_A2) and accumulated separately (VAR_ACC2) because the ReLU [+ dropout] sits on conv2's sum alone:
//   h2  = dropout(relu(conv2(h1) + b2))                 -> out_h2   (kept for the backward pass)
//   out = (conv1x1(x_in) + br) + bf16(h2)               -> out_block
// Replaces conv2 -> h2 and conv1x1(x_in) + h2 -> out (EPI_BIAS_ADD): one launch and one pass over h2 fewer, bit-identical outputs.
// OPT-IN (CSB_CNN_MERGE_FWD=1): two 256-column accumulators fill the 512 TMEM columns, so the epilogue of a tile no longer overlaps
// the next tile's MMAs, and that costs what the saved launch gains (28.5 against 28.4 ms per step on the same box).
static int cnn_conv2_residual(csb_cnn* h, const ConvLayerInfo& c2, const ConvLayerInfo& cr, const CnnBuf& h1, const CnnBuf& xin, CnnBuf& out_h2,
                              CnnBuf& out_block, int64_t B, cudaStream_t st, int drop_layer) {
  tc::GemmParams p = {};
  p.M = (int)(B * h->P); p.N = cnn_mma_width(c2.Cout, c2.Coutp);
  p.kb_per_tap = c2.Cinp / 64; p.tap_center = (c2.taps - 1) / 2;
  p.a2_from_kb = c2.taps * p.kb_per_tap;
  p.K = c2.taps * c2.Cinp + cr.Cinp;
  p.halo_period = h->P;
  p.act = c2.act; p.head_relu_from = -1;
  p.bias = h->params + c2.b_off; p.bias2 = h->params + cr.b_off;
  p.tap_tail_k = cnn_tail_k(c2.Cin, c2.Cinp); p.a2_tail_k = cnn_tail_k(cr.Cin, cr.Cinp);
  p.out = out_h2.ptr; p.ld_out = out_h2.Cp;
  p.out2 = out_block.ptr; p.ld_out2 = out_block.Cp;
  int rc;
  if (drop_layer >= 0) {
    p.drop_seed = cnn_drop_seed(h, drop_layer); p.drop_threshold = cnn_drop_threshold(h); p.drop_scale = 1.f / (1.f - h->dropout);
    rc = launch_tn<256, 6, tc::EPI_BIAS_ACT, 2, tc::VAR_A2 | tc::VAR_ACC2 | tc::VAR_KTRIM | tc::VAR_DROPOUT>(h1.maps.k128, c2.tm_wtc, p, h->sm_count, st, &xin.maps.k128);
  } else {
    rc = launch_tn<256, 6, tc::EPI_BIAS_ACT, 2, tc::VAR_A2 | tc::VAR_ACC2 | tc::VAR_KTRIM>(h1.maps.k128, c2.tm_wtc, p, h->sm_count, st, &xin.maps.k128);
  }
  if (rc) return rc;
  h->launches++;
  return CSB_OK;
}

// Data gradient with two outputs in ONE launch (bf16, CTA pairs):
//   G      = [in taps | in2] . Wd                       (in2 = a second A source behind the taps, VAR_A2; nullptr: none)   -> out_plain
//   masked = bf16(G) * act'(saved) [* ds]                                                                                  -> out_masked
// For a residual block: in = dz1, in2 = d(block output), Wd = [W1^T flipped | Wr^T] -> G = d(block input) = d(previous block's output),
// masked = dz2 of the previous block.  Replaces conv^T(dz1) -> T, conv1x1^T(d_out) + T -> G, act_mask(G, h2) -> dz2.
static int cnn_dgrad_dual(csb_cnn* h, const ConvLayerInfo& li, const CUtensorMap& w, const CnnBuf& in, const CnnBuf* in2, const CnnBuf& saved,
                          int act, CnnBuf& out_masked, CnnBuf& out_plain, int64_t B, cudaStream_t st, float ds) {
  tc::GemmParams p = {};
  p.M = (int)(B * h->P); p.N = cnn_mma_width(li.Cin, li.Cinp);
  p.kb_per_tap = li.Coutp / 64; p.tap_center = (li.taps - 1) / 2;
  p.a2_from_kb = li.taps * p.kb_per_tap;
  p.K = (li.taps + (in2 ? 1 : 0)) * li.Coutp;
  p.halo_period = h->P;
  p.act = act; p.head_relu_from = -1; p.dgrad_scale = ds;
  p.tap_tail_k = cnn_tail_k(li.Cout, li.Coutp); p.a2_tail_k = p.tap_tail_k;
  p.out = out_masked.ptr; p.ld_out = out_masked.Cp;
  p.out2 = out_plain.ptr; p.ld_out2 = out_plain.Cp;
  p.saved = reinterpret_cast<const __nv_bfloat16*>(saved.ptr); p.ld_saved = saved.Cp;
  int rc;
  if (in2) rc = launch_tn<256, 6, tc::EPI_DGRAD, 2, tc::VAR_DUAL | tc::VAR_A2 | tc::VAR_KTRIM>(in.maps.k128, w, p, h->sm_count, st, &in2->maps.k128);
  else rc = launch_tn<256, 6, tc::EPI_DGRAD, 2, tc::VAR_DUAL | tc::VAR_KTRIM>(in.maps.k128, w, p, h->sm_count, st);
  if (rc) return rc;
  h->launches++;
  return CSB_OK;
}

// weight (+ bias) gradient of one conv layer: in [R, Cinp], dz [R, Coutp]
static int cnn_wgrad(csb_cnn* h, const ConvLayerInfo& li, const CnnBuf& in, const CnnBuf& dz, int64_t B, simt::SegmentTable& tab,
                     int64_t& max_len, cudaStream_t st) {
  const int64_t R = B * h->P;
  const size_t tap_elems = (size_t)li.Cinp * li.Coutp;
  if (!h->bf16) {
    for (int t = 0; t < li.taps; ++t) {                       // straight into the gradient buffer: no split, no partials
      simt::SgemmParams p = {};
      p.M = li.Cinp; p.N = li.Coutp; p.K = (int)R;
      p.A = reinterpret_cast<const float*>(in.ptr); p.lda = in.Cp; p.a_row_off = t - (li.taps - 1) / 2;
      p.B = reinterpret_cast<const float*>(dz.ptr); p.ldb = dz.Cp;
      p.C = h->grads + li.w_off + (size_t)t * tap_elems; p.ldc = li.Coutp;
      dim3 grid((unsigned)(li.Coutp / 64), (unsigned)(li.Cinp / 64));
      simt::sgemm_kernel<true, false, simt::SEPI_STORE><<<grid, 256, 0, st>>>(p);
      CSB_CUDA_CHECK(cudaGetLastError());
      h->launches++;
    }
    const int S = (int)std::max<int64_t>(1, std::min<int64_t>(li.max_splits, ceil_div(R, 256)));
    dim3 grid((unsigned)(li.Coutp / 64), (unsigned)S);
    simt::colsum_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(dz.ptr), dz.Cp, R, h->ws + li.ws_b_off, (size_t)li.Coutp);
    CSB_CUDA_CHECK(cudaGetLastError());
    h->launches++;
    tab.seg[tab.n++] = {h->ws + li.ws_b_off, (size_t)li.Coutp, h->grads + li.b_off, (int64_t)li.Coutp, S};
    return CSB_OK;
  }
  const int num_rb = (int)ceil_div(R, 64);
  int splits = std::max(1, std::min(li.max_splits, num_rb));
  // CTA pairs (256-row tiles, each CTA loading half of the dZ columns) whenever the layer has at least two 128-row blocks of input
  // channels and more than 128 output columns: a single CTA needs 96 B/clk of operands, more than an SM takes in
  static const bool no_pairs = getenv("CSB_CNN_NT_SINGLE") != nullptr;            // A/B aid
  const int n_mma = cnn_mma_width(li.Cout, li.Coutp);
  const int cg = (!no_pairs && li.Cinp > 128 && n_mma > 128 && n_mma % 32 == 0) ? 2 : 1;
  // One launch per tap is the default.  CSB_CNN_NT_STACKED=1 (opt-in) runs all taps in one launch -- 24 launches fewer per step and dZ
  // read once per split, correct (same tests), but measured 4 % SLOWER on the step (28.7 against 27.6 ms): fewer, longer splits.
  static const bool per_tap = getenv("CSB_CNN_NT_STACKED") == nullptr;
  int m_tiles_colsum = nt_m_tiles(li.Cinp, cg);
  if (li.taps > 1 && !per_tap) {
    // all taps in one launch: M = taps * Cinp stacked tap-major = the [taps][Cin][Cout] kernel's own layout; dZ is read once per split
    tc::NtParams p = {};
    p.M = li.taps * li.Cinp; p.N = n_mma; p.R = (int)R;
    p.a_tap_m = li.Cinp; p.a_tap_center = (li.taps - 1) / 2;
    const int tiles = nt_m_tiles(p.M, cg) * (int)ceil_div(p.N, tn_block_n(p.N));
    splits = std::max(1, std::min(std::min(li.max_splits, num_rb), std::max(1, (h->sm_count / cg) / tiles)));
    p.rb_per_split = (int)ceil_div(num_rb, splits);
    splits = (int)ceil_div(num_rb, p.rb_per_split);
    p.out = h->ws + li.ws_w_off; p.ld_out = li.Coutp; p.split_stride = (size_t)li.taps * tap_elems;
    p.colsum_out = h->ws + li.ws_b_off; p.colsum_stride = (size_t)li.Coutp;
    m_tiles_colsum = nt_m_tiles(p.M, cg);
    int rc = launch_nt_auto(in.maps.mn64, dz.maps.mn64, p, splits, st, cg);
    if (rc) return rc;
    h->launches++;
  } else {
  for (int t = 0; t < li.taps; ++t) {
    tc::NtParams p = {};
    p.M = li.Cinp; p.N = n_mma; p.R = (int)R;
    p.rb_per_split = (int)ceil_div(num_rb, splits);
    const int eff = (int)ceil_div(num_rb, p.rb_per_split);
    p.out = h->ws + li.ws_w_off + (size_t)t * tap_elems; p.ld_out = li.Coutp; p.split_stride = (size_t)li.taps * tap_elems;
    p.a_row_offset = t - (li.taps - 1) / 2;
    if (t == 0) { p.colsum_out = h->ws + li.ws_b_off; p.colsum_stride = (size_t)li.Coutp; }
    int rc = launch_nt_auto(in.maps.mn64, dz.maps.mn64, p, eff, st, cg);
    if (rc) return rc;
    h->launches++;
    splits = eff;
  }
  }
  tab.seg[tab.n++] = {h->ws + li.ws_w_off, (size_t)li.taps * tap_elems, h->grads + li.w_off, (int64_t)(li.taps * tap_elems), splits};
  tab.seg[tab.n++] = {h->ws + li.ws_b_off, (size_t)li.Coutp, h->grads + li.b_off, (int64_t)li.Coutp, splits * m_tiles_colsum};
  max_len = std::max<int64_t>(max_len, (int64_t)(li.taps * tap_elems));
  return CSB_OK;
}

// inverted dropout on a hidden activation buffer (training forward); the mask depends on (seed, training step, layer)
static int cnn_dropout(csb_cnn* h, CnnBuf& a, int layer_id, int64_t B, cudaStream_t st) {
  const int64_t n8 = B * h->P * a.Cp / 8;
  const uint32_t seed = cnn_drop_seed(h, layer_id);
  const uint32_t thr = cnn_drop_threshold(h);
  simt::dropout_bf16_kernel<<<grid_for(n8, 256, h->sm_count), 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(a.ptr), n8, seed, thr,
                                                                           1.f / (1.f - h->dropout));
  CSB_CUDA_CHECK(cudaGetLastError());
  h->launches++;
  return CSB_OK;
}

static int cnn_forward_body(csb_cnn* h, const float* x, int64_t B, cudaStream_t st, bool training = false) {
  const bool drop = training && h->dropout > 0.f;
  static const bool drop_kernel = getenv("CSB_CNN_DROPOUT_KERNEL") != nullptr;     // A/B aid: the separate element-wise dropout pass
  const bool fold = drop && h->bf16 && !drop_kernel;                              // dropout inside the conv epilogue
  const int64_t R = B * h->P;
  if (h->bf16) simt::cnn_pack_input_kernel<<<grid_for(R * h->in_p, 256, h->sm_count), 256, 0, st>>>(x, reinterpret_cast<__nv_bfloat16*>(h->x0.ptr), B, h->L, h->in_ch, h->in_p);
  else simt::cnn_pack_input_f32_kernel<<<grid_for(R * h->in_p, 256, h->sm_count), 256, 0, st>>>(x, reinterpret_cast<float*>(h->x0.ptr), B, h->L, h->in_ch, h->in_p);
  CSB_CUDA_CHECK(cudaGetLastError());
  h->launches++;
  const CnnBuf* xin = &h->x0;
  int rc;
  const bool merged = h->bf16 && h->merge_fwd && g_use_pairs && h->width_p > 128 && h->layer[1].act != CSB_ACT_ELU && (!drop || fold);
  for (int i = 0; i < h->depth; ++i) {
    const ConvLayerInfo &c1 = h->layer[3 * i], &c2 = h->layer[3 * i + 1], &cr = h->layer[3 * i + 2];
    if ((rc = cnn_conv(h, c1, false, *xin, h->h1[i], 0, nullptr, c1.act, h->params + c1.b_off, B, st, 0.f, fold ? 2 * i : -1))) return rc;
    if (drop && !fold && (rc = cnn_dropout(h, h->h1[i], 2 * i, B, st))) return rc;
    if (merged && c2.has_wtc) {
      if ((rc = cnn_conv2_residual(h, c2, cr, h->h1[i], *xin, h->h2[i], h->ob[i], B, st, fold ? 2 * i + 1 : -1))) return rc;
      xin = &h->ob[i];
      continue;
    }
    if ((rc = cnn_conv(h, c2, false, h->h1[i], h->h2[i], 0, nullptr, c2.act, h->params + c2.b_off, B, st, 0.f, fold ? 2 * i + 1 : -1))) return rc;
    if (drop && !fold && (rc = cnn_dropout(h, h->h2[i], 2 * i + 1, B, st))) return rc;
    // out = conv1x1(x_in) + b + relu(conv2)
    if ((rc = cnn_conv(h, cr, false, *xin, h->ob[i], 1, &h->h2[i], CSB_ACT_NONE, h->params + cr.b_off, B, st))) return rc;
    xin = &h->ob[i];
  }
  const ConvLayerInfo& co = h->layer[3 * h->depth];
  return cnn_conv(h, co, false, *xin, h->e, 0, nullptr, co.act, h->params + co.b_off, B, st);
}

// fp32 mode: head GEMM -> pred (halo layout, fp32)
static int cnn_head_f32(csb_cnn* h, int64_t B, cudaStream_t st) {
  const ConvLayerInfo& cd = h->layer[3 * h->depth + 1];
  simt::SgemmParams p = {};
  p.M = (int)(B * h->P); p.N = cd.Coutp; p.K = cd.Cinp;
  p.A = reinterpret_cast<const float*>(h->e.ptr); p.lda = h->e.Cp;
  p.B = h->params + cd.w_off; p.ldb = cd.Coutp;
  p.C = h->pred; p.ldc = h->out_p;
  p.bias = h->params + cd.b_off; p.act = CSB_ACT_NONE; p.head_relu_from = h->out_lin; p.halo_period = h->P;
  dim3 grid((unsigned)(cd.Coutp / 64), (unsigned)ceil_div(p.M, 64));
  simt::sgemm_kernel<false, false, simt::SEPI_BIAS_ACT><<<grid, 256, 0, st>>>(p);
  CSB_CUDA_CHECK(cudaGetLastError());
  h->launches++;
  return CSB_OK;
}

// everything behind dL/dz of the Dense heads (already in h->dzh): weight / bias gradients of every layer into h->grads
static int cnn_backward_chain(csb_cnn* h, int64_t B, cudaStream_t st) {
  const float ds = h->dropout > 0.f ? 1.f / (1.f - h->dropout) : 1.f;        // gradient through a kept element of a dropout layer
  const int D = h->depth;
  const ConvLayerInfo &co = h->layer[3 * D], &cd = h->layer[3 * D + 1];
  const int64_t R = B * h->P;
  int rc;
  simt::SegmentTable tab;
  tab.n = 0;
  int64_t max_len = 4;
  auto flush_reduce = [&]() -> int {
    if (tab.n == 0) return CSB_OK;
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(max_len / 4, 256), 2 * h->sm_count), (unsigned)tab.n);
    simt::reduce_partials_kernel<<<grid, 256, 0, st>>>(tab);
    CSB_CUDA_CHECK(cudaGetLastError());
    h->launches++;
    tab.n = 0; max_len = 4;
    return CSB_OK;
  };
  auto act_mask = [&](const CnnBuf& g, const CnnBuf& a, CnnBuf& dz, int act) -> int {
    const int64_t n = R * g.Cp;
    if (h->bf16) simt::act_mask_bf16_kernel<<<grid_for(n / 8, 256, h->sm_count), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(g.ptr), reinterpret_cast<const __nv_bfloat16*>(a.ptr), reinterpret_cast<__nv_bfloat16*>(dz.ptr), n / 8, act, 0.f, ds);
    else simt::act_mask_f32_kernel<<<grid_for(n, 256, h->sm_count), 256, 0, st>>>(reinterpret_cast<const float*>(g.ptr), reinterpret_cast<const float*>(a.ptr), reinterpret_cast<float*>(dz.ptr), n, act, 0.f);
    CSB_CUDA_CHECK(cudaGetLastError());
    h->launches++;
    return CSB_OK;
  };
  // ---- Dense heads and the 1x1 output convolution
  if ((rc = cnn_wgrad(h, cd, h->e, h->dzh, B, tab, max_len, st))) return rc;
  if ((rc = cnn_conv(h, cd, true, h->dzh, h->dze, 2, &h->e, co.act, nullptr, B, st))) return rc;            // dze = (dzh . Wd^T) * elu'(e)
  const CnnBuf& last_out = D > 0 ? h->ob[D - 1] : h->x0;
  if ((rc = cnn_wgrad(h, co, last_out, h->dze, B, tab, max_len, st))) return rc;
  int g = 0;
  // Merged launches (bf16, CTA pairs, ReLU-family blocks): every kernel that produces d(block output) also writes dz2 = d_out * relu'(h2)
  // of that block (two outputs), and d(block input) = conv^T(dz1; W1) + conv1x1^T(d_out; Wr) is ONE GEMM over [dz1 taps | d_out]:
  // three launches and four passes over a [R, Cp] tensor fewer per block.  CSB_CNN_NO_MERGE=1: the separate launches (A/B aid).
  const bool merged = h->bf16 && h->merge_bwd && g_use_pairs && D > 0 && h->width_p > 128 && h->layer[1].act != CSB_ACT_ELU;
  // d(block output): no activation after the residual add
  if (merged) {
    if ((rc = cnn_dgrad_dual(h, co, co.tm_wd, h->dze, nullptr, h->h2[D - 1], h->layer[3 * (D - 1) + 1].act, h->Z2, h->G[g], B, st, h->dropout > 0.f ? ds : 0.f))) return rc;
  } else if ((rc = cnn_conv(h, co, true, h->dze, h->G[g], 0, nullptr, CSB_ACT_NONE, h->zero_bias, B, st))) return rc;
  // ---- residual blocks, last to first.  G[g] holds dL/d(block output).
  for (int i = D - 1; i >= 0; --i) {
    const ConvLayerInfo &c1 = h->layer[3 * i], &c2 = h->layer[3 * i + 1], &cr = h->layer[3 * i + 2];
    const CnnBuf& xin = i > 0 ? h->ob[i - 1] : h->x0;
    if (!merged && (rc = act_mask(h->G[g], h->h2[i], h->Z2, c2.act))) return rc;                            // dz2 = d_out * relu'(h2)
    if ((rc = cnn_wgrad(h, cr, xin, h->G[g], B, tab, max_len, st))) return rc;                               // residual 1x1: dW = xin^T d_out
    if ((rc = cnn_wgrad(h, c2, h->h1[i], h->Z2, B, tab, max_len, st))) return rc;
    if ((rc = cnn_conv(h, c2, true, h->Z2, h->Z1, 2, &h->h1[i], c1.act, nullptr, B, st, h->dropout > 0.f ? ds : 0.f))) return rc;   // dz1 = conv^T(dz2; W2) * relu'(h1) [* 1/(1-p)]
    if ((rc = cnn_wgrad(h, c1, xin, h->Z1, B, tab, max_len, st))) return rc;
    if (i > 0 && merged && c1.has_wdc) {
      // d(x_in) = conv^T(dz1; W1) + conv1x1^T(d_out; Wr), and the previous block's dz2 from it
      if ((rc = cnn_dgrad_dual(h, c1, c1.tm_wdc, h->Z1, &h->G[g], h->h2[i - 1], h->layer[3 * (i - 1) + 1].act, h->Z2, h->G[g ^ 1], B, st,
                               h->dropout > 0.f ? ds : 0.f))) return rc;
      g ^= 1;
    } else if (i > 0) {
      // d(x_in) = conv^T(dz1; W1) + conv1x1^T(d_out; Wr)
      if ((rc = cnn_conv(h, c1, true, h->Z1, h->T, 0, nullptr, CSB_ACT_NONE, h->zero_bias, B, st))) return rc;
      if ((rc = cnn_conv(h, cr, true, h->G[g], h->G[g ^ 1], 1, &h->T, CSB_ACT_NONE, h->zero_bias, B, st))) return rc;
      g ^= 1;
    }
    if (tab.n + 8 > (int)(sizeof(tab.seg) / sizeof(tab.seg[0]))) { if ((rc = flush_reduce())) return rc; }
  }
  return flush_reduce();
}

extern "C" {

int csb_cnn_create(const csb_cnn_cfg* cfg, csb_cnn** out) {
  CSB_REQUIRE(cfg && out, CSB_EINVAL, "null argument");
  *out = nullptr;
  CSB_REQUIRE(cfg->depth >= 1 && cfg->depth <= 16, CSB_EINVAL, "depth %d out of range 1..16", cfg->depth);
  CSB_REQUIRE(cfg->kernel == 3 || cfg->kernel == 1, CSB_EINVAL, "kernel width must be 1 or 3");
  CSB_REQUIRE(cfg->width >= 1 && cfg->in_ch >= 1 && cfg->out_ch >= 1 && cfg->out_lin >= 0 && cfg->out_lin <= cfg->out_ch, CSB_EINVAL, "bad channel configuration");
  CSB_REQUIRE(cfg->levels >= 1 && cfg->max_batch >= 1, CSB_EINVAL, "levels / max_batch must be positive");
  CSB_REQUIRE(cfg->dtype == CSB_BF16 || cfg->dtype == CSB_F32, CSB_EINVAL, "unknown dtype %d", cfg->dtype);
  CSB_REQUIRE(cfg->loss == CSB_LOSS_MSE || cfg->loss == CSB_LOSS_MAE, CSB_EINVAL, "unknown loss %d", cfg->loss);
  int sm = 0, maj = 0, mnr = 0;
  int rc = csb_device_info(&sm, &maj, &mnr, nullptr);
  if (rc) return rc;
  CSB_REQUIRE(maj == 10, CSB_ENODEV, "device has compute capability %d.%d; this library is sm_100a only", maj, mnr);
  g_use_pairs = getenv("CSB_NO_PAIRS") == nullptr;
  g_cnn_ktrim = getenv("CSB_CNN_NO_KTRIM") == nullptr;

  csb_cnn* h = new (std::nothrow) csb_cnn();
  CSB_REQUIRE(h != nullptr, CSB_ENOMEM, "host allocation failed");
  h->merge_bwd = getenv("CSB_CNN_NO_MERGE") == nullptr;
  h->merge_fwd = getenv("CSB_CNN_MERGE_FWD") != nullptr;
  h->cfg = *cfg;
  h->depth = cfg->depth; h->L = cfg->levels; h->P = cfg->levels + 2;
  h->in_ch = cfg->in_ch; h->in_p = (int)round_up(cfg->in_ch, 64);
  h->width = cfg->width; h->width_p = (int)round_up(cfg->width, 64);
  h->out_ch = cfg->out_ch; h->out_p = (int)round_up(cfg->out_ch, 64); h->out_lin = cfg->out_lin;
  h->sm_count = sm;
  h->bf16 = cfg->dtype == CSB_BF16;
  h->cap = round_up(cfg->max_batch * h->P, 128);
  h->n_layers = 3 * h->depth + 2;
  size_t off = 0, off_user = 0, ws_off = 0;
  auto add = [&](int idx, int taps, int cin, int cout, int act) {
    ConvLayerInfo& li = h->layer[idx];
    li.taps = taps; li.Cin = cin; li.Cout = cout; li.Cinp = (int)round_up(cin, 64); li.Coutp = (int)round_up(cout, 64); li.act = act;
    li.w_off = off; off += (size_t)taps * li.Cinp * li.Coutp;
    li.b_off = off; off += (size_t)li.Coutp;
    li.w_off_user = off_user; off_user += (size_t)taps * cin * cout;
    li.b_off_user = off_user; off_user += (size_t)cout;
    const int tiles = (int)(ceil_div(li.Cinp, 128) * ceil_div(li.Coutp, tn_block_n(li.Coutp)));
    li.max_splits = h->bf16 ? std::max(1, std::min(64, sm / tiles)) : 32;
    li.ws_w_off = ws_off; if (h->bf16) ws_off += (size_t)li.max_splits * taps * li.Cinp * li.Coutp;   // fp32 mode writes dW directly
    li.ws_b_off = ws_off; ws_off += (size_t)li.max_splits * ceil_div((int64_t)taps * li.Cinp, 128) * li.Coutp;
  };
  int c = cfg->in_ch;
  for (int i = 0; i < h->depth; ++i) {
    add(3 * i, cfg->kernel, c, cfg->width, cfg->act);
    add(3 * i + 1, cfg->kernel, cfg->width, cfg->width, cfg->act);
    add(3 * i + 2, 1, c, cfg->width, CSB_ACT_NONE);
    c = cfg->width;
  }
  add(3 * h->depth, 1, c, cfg->out_ch, cfg->pre_out_act);
  add(3 * h->depth + 1, 1, cfg->out_ch, cfg->out_ch, CSB_ACT_NONE);      // Dense(out_lin, linear) || Dense(out_ch - out_lin, relu), fused column-wise
  h->P_pad = off; h->P_user = off_user; h->ws_elems = ws_off;

#define CKA(ptr, bytes) do { int _rc = [&]() -> int { CSB_ALLOC(ptr, bytes); return CSB_OK; }(); if (_rc) { cnn_free(h); delete h; return _rc; } } while (0)
  CKA(h->params, h->P_pad * 4); CKA(h->grads, h->P_pad * 4); CKA(h->m, h->P_pad * 4); CKA(h->v, h->P_pad * 4);
  CKA(h->ws, h->ws_elems * 4);
  CKA(h->zero_bias, (size_t)std::max(h->width_p, h->out_p) * 4);
  const size_t es = h->bf16 ? 2 : 4;
  auto mk = [&](CnnBuf& b, int Cp) -> int { b.Cp = Cp; CSB_ALLOC(b.ptr, (size_t)h->cap * Cp * es); return CSB_OK; };
#define CKB(buf, Cp) do { int _rc = mk(buf, Cp); if (_rc) { cnn_free(h); delete h; return _rc; } } while (0)
  CKB(h->x0, h->in_p);
  for (int i = 0; i < h->depth; ++i) { CKB(h->h1[i], h->width_p); CKB(h->h2[i], h->width_p); CKB(h->ob[i], h->width_p); }
  CKB(h->e, h->out_p); CKB(h->dzh, h->out_p); CKB(h->dze, h->out_p);
  CKB(h->G[0], h->width_p); CKB(h->G[1], h->width_p); CKB(h->Z1, h->width_p); CKB(h->Z2, h->width_p); CKB(h->T, h->width_p);
#undef CKB
  if (!h->bf16) CKA(h->pred, (size_t)h->cap * h->out_p * 4);
  CKA(h->d_loss_w, (size_t)h->out_p * 4);
  h->n_loss_partials = (int)std::max<int64_t>(h->cap / 128 * tc::TN_EPI_WARPS, 8 * sm);
  CKA(h->loss_partials, (size_t)h->n_loss_partials * 4);
  CKA(h->d_loss, 4);
  for (int l = 0; l < h->n_layers && h->bf16; ++l) {
    ConvLayerInfo& li = h->layer[l];
    li.wt_ld = li.taps * li.Cinp; li.wd_ld = li.taps * li.Coutp;
  }
  // residual blocks: the 1x1 layer's transposed copy behind conv1's (both map d(block output) / dz1 -> d(block input)), its forward
  // copy behind conv2's (both produce the block output)
  for (int i = 0; i < h->depth && h->bf16; ++i) {
    ConvLayerInfo &c1 = h->layer[3 * i], &c2 = h->layer[3 * i + 1], &cr = h->layer[3 * i + 2];
    if (c1.Cinp == cr.Cinp && c1.Coutp == cr.Coutp) {
      c1.wd_ld = cr.wd_ld = (c1.taps + 1) * c1.Coutp; cr.wd_col0 = c1.taps * c1.Coutp; cr.wd_alias = true; c1.has_wdc = true;
    }
    if (c2.Coutp == cr.Coutp) {
      c2.wt_ld = cr.wt_ld = c2.taps * c2.Cinp + cr.Cinp; cr.wt_col0 = c2.taps * c2.Cinp; cr.wt_alias = true; c2.has_wtc = true;
    }
  }
  for (int l = 0; l < h->n_layers && h->bf16; ++l) {
    ConvLayerInfo& li = h->layer[l];
    if (!li.wt_alias) CKA(li.wt16, (size_t)li.Coutp * li.wt_ld * 2);
    if (!li.wd_alias) CKA(li.wd16, (size_t)li.Cinp * li.wd_ld * 2);
  }
#undef CKA
  for (int i = 0; i < h->depth && h->bf16; ++i) {
    ConvLayerInfo &c1 = h->layer[3 * i], &c2 = h->layer[3 * i + 1], &cr = h->layer[3 * i + 2];
    if (cr.wd_alias) cr.wd16 = c1.wd16;
    if (cr.wt_alias) cr.wt16 = c2.wt16;
  }
  for (int l = 0; l < h->n_layers && h->bf16; ++l) {
    ConvLayerInfo& li = h->layer[l];
    rc = make_tmap_bf16(&li.tm_wt, li.wt16 + li.wt_col0, (uint64_t)li.taps * li.Cinp, li.Coutp, (uint64_t)li.wt_ld, 64, (uint32_t)tn_b_box_rows(li.Coutp));
    if (!rc) rc = make_tmap_bf16(&li.tm_wd, li.wd16 + li.wd_col0, (uint64_t)li.taps * li.Coutp, li.Cinp, (uint64_t)li.wd_ld, 64, (uint32_t)tn_b_box_rows(li.Cinp));
    if (!rc && li.has_wtc) rc = make_tmap_bf16(&li.tm_wtc, li.wt16, (uint64_t)li.wt_ld, li.Coutp, (uint64_t)li.wt_ld, 64, (uint32_t)tn_b_box_rows(li.Coutp));
    if (!rc && li.has_wdc) rc = make_tmap_bf16(&li.tm_wdc, li.wd16, (uint64_t)li.wd_ld, li.Cinp, (uint64_t)li.wd_ld, 64, (uint32_t)tn_b_box_rows(li.Cinp));
    if (rc) { cnn_free(h); delete h; return rc; }
  }
  // default loss weights: the reference's *_adjusted losses (hpo_train.py:114-121): per-sample sum over (level, channel) of
  // w_c * e with w = (120/128)/(levels*out_lin) on the profile channels and (8/128)/(levels*(out_ch-out_lin)) on the scalars
  {
    std::vector<float> w(h->out_p, 0.f);
    for (int ch = 0; ch < h->out_ch; ++ch)
      w[ch] = ch < h->out_lin ? (120.f / 128.f) / (float)(h->L * h->out_lin) : (8.f / 128.f) / (float)(h->L * (h->out_ch - h->out_lin));
    if (cudaMemcpy(h->d_loss_w, w.data(), (size_t)h->out_p * 4, cudaMemcpyHostToDevice) != cudaSuccess) { cnn_free(h); delete h; set_last_error("cudaMemcpy failed"); return CSB_ECUDA; }
  }
  *out = h;
  return CSB_OK;
}

int csb_cnn_destroy(csb_cnn* h) {
  if (!h) return CSB_OK;
  cudaDeviceSynchronize();
  cnn_free(h);
  delete h;
  return CSB_OK;
}

size_t csb_cnn_param_count(const csb_cnn* h) { return h ? h->P_user : 0; }
int64_t csb_cnn_launch_count(const csb_cnn* h) { return h ? h->launches : 0; }

int csb_cnn_set_params(csb_cnn* h, const float* params_host) {
  CSB_REQUIRE(h && params_host, CSB_EINVAL, "null argument");
  float* tmp = nullptr;
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  CSB_CUDA_CHECK(cudaMalloc(&tmp, h->P_user * 4));
  cudaError_t e = cudaMemcpy(tmp, params_host, h->P_user * 4, cudaMemcpyHostToDevice);
  int rc = (e == cudaSuccess) ? cnn_pad_copy(h, h->params, tmp, 0, 0) : CSB_ECUDA;
  if (!rc) rc = cnn_repack(h, 0);
  cudaDeviceSynchronize();
  cudaFree(tmp);
  return rc;
}
static int cnn_download(csb_cnn* h, float* dev_padded, float* host) {
  float* tmp = nullptr;
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  CSB_CUDA_CHECK(cudaMalloc(&tmp, h->P_user * 4));
  int rc = cnn_pad_copy(h, dev_padded, tmp, 1, 0);
  cudaError_t e = cudaMemcpy(host, tmp, h->P_user * 4, cudaMemcpyDeviceToHost);
  cudaFree(tmp);
  if (e != cudaSuccess) { set_last_error("cudaMemcpy failed"); return CSB_ECUDA; }
  return rc;
}
int csb_cnn_get_params(csb_cnn* h, float* params_host) {
  CSB_REQUIRE(h && params_host, CSB_EINVAL, "null argument");
  return cnn_download(h, h->params, params_host);
}
int csb_cnn_get_grads(csb_cnn* h, float* grads_host) {
  CSB_REQUIRE(h && grads_host, CSB_EINVAL, "null argument");
  return cnn_download(h, h->grads, grads_host);
}
int csb_cnn_grad_buffer(csb_cnn* h, float** ptr, size_t* n) {
  CSB_REQUIRE(h && ptr && n, CSB_EINVAL, "null argument");
  *ptr = h->grads; *n = h->P_pad;
  return CSB_OK;
}
int csb_cnn_set_loss_weights(csb_cnn* h, const float* w_host) {
  CSB_REQUIRE(h && w_host, CSB_EINVAL, "null argument");
  std::vector<float> w(h->out_p, 0.f);
  for (int c = 0; c < h->out_ch; ++c) w[c] = w_host[c];
  CSB_CUDA_CHECK(cudaDeviceSynchronize());
  CSB_CUDA_CHECK(cudaMemcpy(h->d_loss_w, w.data(), (size_t)h->out_p * 4, cudaMemcpyHostToDevice));
  return CSB_OK;
}

int csb_cnn_set_dropout(csb_cnn* h, float rate, uint32_t seed) {
  CSB_REQUIRE(h, CSB_EINVAL, "null handle");
  CSB_REQUIRE(rate >= 0.f && rate < 1.f, CSB_EINVAL, "dropout rate %g outside [0, 1)", (double)rate);
  CSB_REQUIRE(rate == 0.f || h->bf16, CSB_EUNSUPPORTED, "dropout is implemented for the CSB_BF16 engine (the fp32 engine is the parity mode)");
  for (int i = 0; i < h->depth && rate > 0.f; ++i)
    CSB_REQUIRE(h->layer[3 * i].act == CSB_ACT_RELU && h->layer[3 * i + 1].act == CSB_ACT_RELU, CSB_EUNSUPPORTED,
                "the mask-free dropout backward relies on ReLU (relu'(0) = 0)");
  h->dropout = rate; h->drop_seed = seed; h->train_steps = 0;
  return CSB_OK;
}

int csb_cnn_debug_read_hidden(csb_cnn* h, int which, int block, float* dst, int64_t B, void* stream) {
  CSB_REQUIRE(h && dst, CSB_EINVAL, "null argument");
  CSB_REQUIRE(h->bf16, CSB_EUNSUPPORTED, "bf16 engine only");
  CSB_REQUIRE((which == 1 || which == 2) && block >= 0 && block < h->depth && B >= 1 && B <= h->cfg.max_batch, CSB_EINVAL, "bad selector");
  const CnnBuf& b = which == 1 ? h->h1[block] : h->h2[block];
  const int C = h->layer[3 * block].Cout;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  simt::cnn_unpack_hidden_kernel<<<grid_for(B * h->L * C, 256, h->sm_count), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(b.ptr), b.Cp, dst, B, h->L, C);
  CSB_CUDA_CHECK(cudaGetLastError());
  return CSB_OK;
}

int csb_cnn_forward(csb_cnn* h, const float* x, float* y_pred, int64_t B, void* stream) {
  CSB_REQUIRE(h && x && y_pred, CSB_EINVAL, "null argument");
  CSB_REQUIRE(B >= 1 && B <= h->cfg.max_batch, CSB_ESTATE, "batch %lld outside 1..max_batch %lld", (long long)B, (long long)h->cfg.max_batch);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc;
  if ((rc = cnn_build_maps(h, B))) return rc;
  if ((rc = cnn_forward_body(h, x, B, st))) return rc;
  h->fwd_B = B;
  const ConvLayerInfo& cd = h->layer[3 * h->depth + 1];
  if (!h->bf16) {
    if ((rc = cnn_head_f32(h, B, st))) return rc;
    simt::cnn_unpack_output_kernel<<<grid_for(B * h->L * h->out_ch, 256, h->sm_count), 256, 0, st>>>(h->pred, h->out_p, y_pred, B, h->L, h->out_ch);
    CSB_CUDA_CHECK(cudaGetLastError());
    h->launches++;
    return CSB_OK;
  }
  tc::GemmParams p = {};
  p.M = (int)(B * h->P); p.N = cd.Coutp; p.K = cd.Cinp; p.halo_period = h->P;
  p.act = CSB_ACT_NONE; p.head_relu_from = h->out_lin; p.bias = h->params + cd.b_off;
  p.out_dim = h->out_ch; p.pred = y_pred; p.ld_pred = h->out_ch;
  rc = launch_tn_auto<tc::EPI_HEAD_OUT>(h->e.maps.k128, cd.tm_wt, p, h->sm_count, st);
  if (rc) return rc;
  h->launches++;
  return CSB_OK;
}

int csb_cnn_train_step(csb_cnn* h, const float* x, const float* y, int64_t B, float grad_scale, float* loss_out, void* stream) {
  CSB_REQUIRE(h && x && y, CSB_EINVAL, "null argument");
  CSB_REQUIRE(B >= 1 && B <= h->cfg.max_batch, CSB_ESTATE, "batch %lld outside 1..max_batch %lld", (long long)B, (long long)h->cfg.max_batch);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (grad_scale <= 0.f) grad_scale = 1.f / (float)B;       // the *_adjusted losses average over the batch (weights carry the rest)
  int rc;
  if ((rc = cnn_build_maps(h, B))) return rc;
  if ((rc = cnn_forward_body(h, x, B, st, true))) return rc;
  h->train_steps++;
  h->fwd_B = -1;
  const ConvLayerInfo& cd = h->layer[3 * h->depth + 1];
  const int64_t R = B * h->P;
  // ---- head + loss: dzh = dL/dz of the fused Dense heads
  if (h->bf16) {
    tc::GemmParams p = {};
    p.M = (int)R; p.N = cd.Coutp; p.K = cd.Cinp; p.halo_period = h->P;
    p.act = CSB_ACT_NONE; p.head_relu_from = h->out_lin; p.bias = h->params + cd.b_off; p.out_dim = h->out_ch;
    p.y = y; p.ld_y = h->out_ch; p.loss_w = h->d_loss_w; p.grad_scale = grad_scale; p.loss_kind = h->cfg.loss;
    p.loss_partials = h->loss_partials;
    p.out = h->dzh.ptr; p.ld_out = h->dzh.Cp;
    if ((rc = launch_tn_auto<tc::EPI_HEAD_LOSS>(h->e.maps.k128, cd.tm_wt, p, h->sm_count, st))) return rc;
    h->launches++;
    simt::loss_finalize_kernel<<<1, 256, 0, st>>>(h->loss_partials, (int)ceil_div(R, 128) * tc::TN_EPI_WARPS, loss_out ? loss_out : h->d_loss);
  } else {
    if ((rc = cnn_head_f32(h, B, st))) return rc;
    const int grid = std::min(h->n_loss_partials, grid_for(R * h->out_p, 256, h->sm_count));
    simt::head_grad_kernel<float><<<grid, 256, 0, st>>>(h->pred, h->out_p, y, h->out_ch, h->d_loss_w, grad_scale, h->cfg.loss, 0, CSB_ACT_NONE, 0.f,
                                                        h->out_lin, nullptr, reinterpret_cast<float*>(h->dzh.ptr), h->out_p, R, h->out_ch, h->out_p,
                                                        h->loss_partials, h->P);
    CSB_CUDA_CHECK(cudaGetLastError());
    h->launches++;
    simt::loss_finalize_kernel<<<1, 256, 0, st>>>(h->loss_partials, grid, loss_out ? loss_out : h->d_loss);
  }
  CSB_CUDA_CHECK(cudaGetLastError());
  h->launches++;
  return cnn_backward_chain(h, B, st);
}

int csb_cnn_backward(csb_cnn* h, const float* y_pred, const float* dy, int64_t B, void* stream) {
  CSB_REQUIRE(h && y_pred && dy, CSB_EINVAL, "null argument");
  CSB_REQUIRE(B >= 1 && B <= h->cfg.max_batch && h->fwd_B == B, CSB_ESTATE,
              "csb_cnn_backward follows csb_cnn_forward on the same batch (the activations live in the handle)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t n = B * h->P * h->out_p;
  if (h->bf16) simt::cnn_head_grad_kernel<__nv_bfloat16><<<grid_for(n, 256, h->sm_count), 256, 0, st>>>(y_pred, dy, reinterpret_cast<__nv_bfloat16*>(h->dzh.ptr), h->out_p, B, h->L, h->out_ch, h->out_lin);
  else simt::cnn_head_grad_kernel<float><<<grid_for(n, 256, h->sm_count), 256, 0, st>>>(y_pred, dy, reinterpret_cast<float*>(h->dzh.ptr), h->out_p, B, h->L, h->out_ch, h->out_lin);
  CSB_CUDA_CHECK(cudaGetLastError());
  h->launches++;
  const float keep = h->dropout;
  h->dropout = 0.f;                          // csb_cnn_forward never drops: no 1 / (1 - p) factors in its backward pass
  const int rc = cnn_backward_chain(h, B, st);
  h->dropout = keep;
  return rc;
}
int csb_cnn_set_params_device(csb_cnn* h, const float* params_dev, void* stream) {
  CSB_REQUIRE(h && params_dev, CSB_EINVAL, "null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = cnn_pad_copy(h, h->params, const_cast<float*>(params_dev), 0, st);
  return rc ? rc : cnn_repack(h, st);
}
int csb_cnn_get_grads_device(csb_cnn* h, float* grads_dev, void* stream) {
  CSB_REQUIRE(h && grads_dev, CSB_EINVAL, "null argument");
  return cnn_pad_copy(h, h->grads, grads_dev, 1, reinterpret_cast<cudaStream_t>(stream));
}

int csb_cnn_apply_opt(csb_cnn* h, int rule, float lr, float beta1, float beta2, float eps, float wd, void* stream) {
  CSB_REQUIRE(h, CSB_EINVAL, "null handle");
  CSB_REQUIRE(rule >= CSB_OPT_ADAM_KERAS && rule <= CSB_OPT_RMSPROP, CSB_EINVAL, "unknown optimizer rule %d", rule);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
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
  simt::opt_kernel<<<grid_for((int64_t)h->P_pad / 4, 256, h->sm_count), 256, 0, st>>>(h->params, h->grads, h->m, h->v, (int64_t)h->P_pad, o);
  CSB_CUDA_CHECK(cudaGetLastError());
  h->launches++;
  return cnn_repack(h, st);
}

}  // extern "C"
