/* climsim_b200.h -- C ABI of libclimsim_b200.so: the B200-native column-emulator engine.
 *
 * The reference (leap-stc/ClimSim) has no FFI / plugin interface: its hot path is whatever TensorFlow or PyTorch
 * executes for the model graphs built inline in its training scripts.  This header therefore DEFINES the boundary a
 * maintainer would bind; every entry point names the reference statement(s) it replaces.
 * Paths are relative to the reference checkout.
 *
 * Conventions
 *   - plain C: opaque handle, POD config, raw pointers and sizes; no C++/torch types.
 *   - unless a name ends in `_host`, data pointers are DEVICE pointers; the caller owns every buffer.
 *   - every call returns 0 (CSB_OK) or a negative CSB_E* code, never throws; csb_strerror() describes it and
 *     csb_last_error() returns the detail string of the most recent failure on the calling thread.
 *   - all device work is enqueued on the caller's stream (`void* stream` is a cudaStream_t; NULL = default stream);
 *     calls do not synchronise unless documented (the `_host` variants synchronise before returning).
 *   - a handle is not thread-safe; distinct handles are independent.
 *   - arrays are row-major fp32 with the reference's column order (climsim_utils/data_utils.py:172-188,815-820):
 *     inputs (B,124) = state_t[60] | state_q0001[60] | ps | SOLIN | LHFLX | SHFLX,
 *     targets (B,128) = ptend_t[60] | ptend_q0001[60] | NETSW FLWDS PRECSC PRECC SOLS SOLL SOLSD SOLLD.
 *   - there is NO CPU fallback: on a machine without an sm_100 device csb_*_create returns CSB_ENODEV.
 */
#ifndef CLIMSIM_B200_H
#define CLIMSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSB_VERSION 100          /* 0.1.0 */
#define CSB_MAX_LAYERS 24

/* ---- error codes ---------------------------------------------------------------------------------------- */
enum {
  CSB_OK = 0,
  CSB_EINVAL = -1,    /* bad argument / configuration */
  CSB_ENODEV = -2,    /* no CUDA device of compute capability 10.x */
  CSB_ENOMEM = -3,    /* device allocation failed */
  CSB_ECUDA = -4,     /* a CUDA runtime / driver call or kernel launch failed */
  CSB_ESTATE = -5,    /* call order violated (e.g. backward before forward, batch larger than max_batch) */
  CSB_EUNSUPPORTED = -6
};

/* ---- enums ---------------------------------------------------------------------------------------------- */
/* Activations of the reference's Keras hyper-parameter space (baseline_v1/hpo_baseline_v1.py:67,82-87) */
enum { CSB_ACT_NONE = 0, CSB_ACT_RELU = 1, CSB_ACT_ELU = 2 /* alpha = 1 */, CSB_ACT_LEAKYRELU = 3 };

/* Arithmetic mode.  CSB_F32: fp32 FFMA everywhere (parity mode, <= 1e-5 relative to the fp32 CPU reference).
 * CSB_BF16: bf16 operands on the tcgen05 tensor cores with fp32 accumulation in TMEM, fp32 master weights,
 * fp32 loss / optimizer (throughput mode; tolerance stated in tests/test_mlp_gpu.py).
 * CSB_TF32 (MLP family): fp32 storage everywhere, every GEMM of the step on the SAME tcgen05 kernels with kind::tf32 products (the
 * tensor core reads sign, exponent and 10 mantissa bits of each fp32 operand, accumulates in fp32) -- the arithmetic the reference's
 * own A100 runs used (TF32 is TensorFlow's default there; step3_prediction/step3_inference.ipynb cell 2).  Agreement with the fp32
 * oracle ~1e-3; with operands representable in TF32 the products are exact up to summation order (tests/test_gemm_gpu.py).
 * CSB_TF32X3 (MLP family): the same kernels with every operand split into hi + lo TF32 parts and three products per fp32 product
 * (hi.hi + lo.hi + hi.lo over a tripled contraction), no stored tensor rounded: fp32-class results (<= 3e-5 against the fp32 oracle;
 * the rest is the tensor core's truncating fp32 accumulation) from the tensor-core pipeline -- the parity mode that shares the
 * benchmarked kernels.  CSB_F32 stays the <= 1e-5 reference point. */
enum { CSB_F32 = 0, CSB_BF16 = 1, CSB_TF32 = 2, CSB_TF32X3 = 3 };

/* Loss.  MSE: mean_ij w_j (p_ij - y_ij)^2  (w = 1 is Keras 'mse', hpo_baseline_v1.py:127-129).
 * MAE: mean_ij w_j |p_ij - y_ij|           (CNN mae_adjusted through w, CNN/training/hpo_train.py:114-121). */
/* HUBER: torch.nn.HuberLoss(delta=1) (online_testing/.../train_mlp_h5loader.py:226-232): mean_ij w_j h(p-y), h(d) = d^2/2 if |d| <= 1 else |d| - 1/2 */
enum { CSB_LOSS_MSE = 0, CSB_LOSS_MAE = 1, CSB_LOSS_HUBER = 2 };

/* Optimizer update rule.
 * ADAM_KERAS: keras.optimizers.Adam update_step (hpo_baseline_v1.py:116-117): w -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)
 * ADAM_TORCH: torch.optim.Adam with L2 weight decay (HSR/training/hsr.py:109-112): w -= lr/(1-b1^t) * m/(sqrt(v/(1-b2^t))+eps)
 * SGD:        w -= lr * (g + wd*w)
 * RADAM:      tensorflow_addons.optimizers.RectifiedAdam, no warm-up, sma_threshold 5 (the best MLP_v1 trial's optimizer,
 *             hpo_baseline_v1.py:118-119):  w -= lr * (sma_t >= 5 ? r_t * mhat/(sqrt(vhat)+eps) : mhat)
 * RMSPROP:    keras.optimizers.RMSprop (rho = beta2 argument, hpo_baseline_v1.py:120-121): v = rho v + (1-rho) g^2; w -= lr * g * rsqrt(v + eps)
 */
enum { CSB_OPT_ADAM_KERAS = 0, CSB_OPT_ADAM_TORCH = 1, CSB_OPT_SGD = 2, CSB_OPT_RADAM = 3, CSB_OPT_RMSPROP = 4 };

/* flags for csb_mlp_forward */
enum {
  CSB_FWD_NORMALIZE_IN = 1,   /* x is raw: apply (x - inp_sub)/inp_div with inf/nan -> 0 (data_utils.py:806-809,894-897) */
  CSB_FWD_DENORM_OUT = 2,     /* divide predictions by out_scale (data_utils.py:1187-1197, step [0] of output_weighting) */
  CSB_FWD_KEEP_ACTIVATIONS = 4, /* keep per-layer activations so that csb_mlp_backward may follow */
  /* csb_mlp_train_step only: the caller will call csb_mlp_apply_opt next and does not need the gradient buffer before that.
   * The reduction of the split-K gradient partials (and the loss sum) is then fused into the optimizer launch, which also
   * refreshes the bf16 weight copies and still fills the gradient buffer.  Results are bit-identical to the unfused path.
   * Not for data-parallel callers (they all-reduce the gradient buffer between the two calls).  If something else reads the
   * gradients first (csb_mlp_get_grads*, another train step) the engine reduces them on demand. */
  CSB_TRAIN_FUSED_OPT = 8,
  /* csb_mlp_forward: this is the forward of a training step (module.train() in the reference's PyTorch models): the Dropout layers
   * set with csb_mlp_set_dropout are active.  Combine with CSB_FWD_KEEP_ACTIVATIONS when csb_mlp_backward follows. */
  CSB_FWD_TRAINING = 32
};

/* ---- MLP family ----------------------------------------------------------------------------------------- */
/* A dense stack  h_{l+1} = act_l( LN_l?( h_l W_l + b_l ) ),  l = 0..n_layers-1.
 *  - MLP_v1 (hpo_baseline_v1.py:75-103): units {768,640,512,640,640,128,128}, act LEAKYRELU(0.15) on the first six
 *    layers, last layer = the two Keras output Dense layers fused column-wise ([120 linear | 8 relu]):
 *    act[last] = NONE, head_relu_from = 120.
 *  - ED (ED/training/ClimSIM_ED_1_3_train.py:56-92): 14 layers, RELU, last ELU.
 *  - HSR mean / log-precision nets (HSR/training/hsr.py:14-35): layernorm = 1 and RELU on the hidden layers. */
typedef struct csb_mlp_cfg {
  int32_t in_dim;                      /* 124 */
  int32_t n_layers;                    /* 1..CSB_MAX_LAYERS, the output layer included */
  int32_t units[CSB_MAX_LAYERS];       /* output width of each layer */
  int32_t act[CSB_MAX_LAYERS];         /* CSB_ACT_* per layer */
  float   alpha[CSB_MAX_LAYERS];       /* LeakyReLU slope (Keras alpha) */
  int32_t layernorm[CSB_MAX_LAYERS];   /* 1: LayerNorm(eps 1e-5, affine) between the Linear and the activation */
  int32_t head_relu_from;              /* last layer only: columns >= this index get ReLU; -1 = none */
  int32_t dtype;                       /* CSB_F32 | CSB_BF16 */
  int32_t loss;                        /* CSB_LOSS_* */
  int64_t max_batch;                   /* capacity (columns per call) of the internal activation buffers */
} csb_mlp_cfg;

typedef struct csb_mlp csb_mlp;

int  csb_mlp_create(const csb_mlp_cfg* cfg, csb_mlp** out);   /* replaces keras.Model(...)/model.compile (hpo_baseline_v1.py:103,127) */
int  csb_mlp_destroy(csb_mlp* h);

/* Flat parameter blob, fp32, layer by layer: W_l (in_l x out_l, row-major == Keras kernel layout), b_l (out_l),
 * then gamma_l, beta_l (out_l each) if layernorm[l].  csb_mlp_param_count() is its length (1 753 472 for MLP_v1). */
size_t csb_mlp_param_count(const csb_mlp* h);
int  csb_mlp_set_params(csb_mlp* h, const float* params_host);        /* model.set_weights / load_state_dict */
int  csb_mlp_get_params(csb_mlp* h, float* params_host);              /* model.get_weights / state_dict (synchronises) */
int  csb_mlp_get_grads(csb_mlp* h, float* grads_host);                /* same order as the parameter blob (synchronises) */
/* the same three with DEVICE blobs, stream-ordered (no synchronisation): how an nn.Module / optimizer that keeps its
 * own flat parameter tensor exchanges weights and gradients with the engine */
int  csb_mlp_set_params_device(csb_mlp* h, const float* params_dev, void* stream);
int  csb_mlp_get_params_device(csb_mlp* h, float* params_dev, void* stream);
int  csb_mlp_get_grads_device(csb_mlp* h, float* grads_dev, void* stream);
/* optimizer state (m, v, each csb_mlp_param_count() long) and step counter: ModelCheckpoint .h5 incl. optimizer
 * state (step2_retrain.py:253-261) / torch.save(state_dict) (hsr.py:120-121) */
int  csb_mlp_get_opt_state(csb_mlp* h, float* m_host, float* v_host, int64_t* step);
int  csb_mlp_set_opt_state(csb_mlp* h, const float* m_host, const float* v_host, int64_t step);

/* Normalisation / loss vectors (HOST pointers; NULL keeps the default): inp_sub[in_dim] (0), inp_div[in_dim] (1),
 * out_scale[out_dim] (1), loss_w[out_dim] (1).  data_utils.save_norm (data_utils.py:954-988) produces the first three. */
int  csb_mlp_set_norm(csb_mlp* h, const float* inp_sub, const float* inp_div, const float* out_scale, const float* loss_w);

/* Generalised input prologue of the online models, applied by every call that passes CSB_FWD_NORMALIZE_IN, in the reference's
 * order (online_testing/model_postprocessing/v2_nn_wrapper.ipynb cell 5 `preprocessing`;
 * online_testing/baseline_models/MLP_v2rh/training/climsim_datapip_h5.py:132-168):
 *   x' = 1 - exp(-exp_lambda[j] * x) where exp_lambda[j] != 0 (cloud liquid / ice);  (x' - inp_sub[j]) / inp_div[j];  nan, inf -> 0;
 *   column j zeroed where keep[j] == 0 (pruned stratospheric inputs);  clamped to [clip_lo[j], clip_hi[j]] (relative humidity).
 * Host arrays of in_dim floats; NULL = no transform / keep everything / no bound.  All NULL restores the plain normalisation.
 * csb_mlp_backward cannot return dL/dx through a transform (CSB_EUNSUPPORTED). */
int  csb_mlp_set_input_transform(csb_mlp* h, const float* exp_lambda, const float* keep, const float* clip_lo, const float* clip_hi);

/* Per-output-column 0/1 mask applied to the predictions (and therefore to their gradients): the online MLP's `output_prune`
 * zeroing of the stratospheric levels (online_testing/baseline_models/MLP_v2rh/training/mlp.py:56-61).  NULL removes the mask. */
int  csb_mlp_set_output_mask(csb_mlp* h, const float* mask_host);

/* torch.nn.Dropout(p) behind every hidden layer (baseline_models/HSR/training/hsr.py:20-25: Linear -> LayerNorm -> Dropout -> ReLU;
 * online_testing/baseline_models/MLP_v2rh/training/mlp.py:41-45: Linear -> Dropout -> ReLU; relu(dropout(u)) == dropout(relu(u))).
 * Active in training forwards only: csb_mlp_train_step*, csb_hsr_train_step, csb_mlp_forward with CSB_FWD_TRAINING.  Inverted dropout
 * with counter-based keep decisions keyed by (seed, training forward, layer, element): no mask is stored, and both arithmetic modes drop
 * the same elements.  PyTorch's random stream is not reproduced (parity is checked with the engine's own masks replayed in the oracle:
 * csb_mlp_debug_dropout_mask writes the multipliers, 0 or 1/(1-p), of the LAST training forward as fp32 [B, units[layer]] on the device).
 * Hidden activations must be ReLU / LeakyReLU / none.  rate 0 switches it off.  Training steps with dropout are not graph-replayed. */
int  csb_mlp_set_dropout(csb_mlp* h, float rate, uint32_t seed);
int  csb_mlp_debug_dropout_mask(csb_mlp* h, int layer, float* dst_dev, int64_t B, void* stream);

/* model.predict / module.forward: x (B,in_dim) -> y_pred (B,out_dim).  step3_inference.ipynb cell 2; hsr.py:28-35 */
int  csb_mlp_forward(csb_mlp* h, const float* x, float* y_pred, int64_t B, uint32_t flags, void* stream);
int  csb_mlp_forward_host(csb_mlp* h, const float* x_host, float* y_pred_host, int64_t B, uint32_t flags, void* stream);

/* autograd entry: given dL/dy_pred (B,out_dim) for the batch of the last KEEP_ACTIVATIONS forward, accumulate the
 * parameter gradients into the internal gradient buffer (overwriting it) and, if dx != NULL, write dL/dx (B,in_dim). */
int  csb_mlp_backward(csb_mlp* h, const float* dy, float* dx, int64_t B, void* stream);

/* One model.fit inner step minus the optimizer (TF train_function, step2_retrain.py:280-285; hsr.py:122-140):
 * forward + loss + backward on normalised x (B,in_dim), scaled targets y (B,out_dim).
 *   loss   = grad_scale * sum_ij w_j (p_ij - y_ij)^2        (or |.| for MAE)
 *   grads  = d loss / d params, written to the internal flat gradient buffer
 * grad_scale <= 0 selects 1/(B*out_dim) (Keras global mean); a data-parallel caller passes 1/(global_B*out_dim)
 * and all-reduces (sums) the gradient buffer.  loss_out (device float, may be NULL) receives the scalar. */
int  csb_mlp_train_step(csb_mlp* h, const float* x, const float* y, int64_t B, float grad_scale, uint32_t flags,
                        float* loss_out, void* stream);

/* The internal gradient buffer (device, fp32, padded layout, *n elements) -- the message of the data-parallel
 * ncclAllReduce (the reference's only collective: DDP gradient allreduce, online_testing/.../train_mlp_h5loader.py:195-207). */
int  csb_mlp_grad_buffer(csb_mlp* h, float** ptr, size_t* n);

/* optimizer.apply_gradients / opt.step(): consumes the gradient buffer, advances the internal step counter.
 * rule = CSB_OPT_*;  wd = L2 weight decay (ADAM_TORCH/SGD).  In CSB_BF16 mode also refreshes the bf16 weight copies. */
int  csb_mlp_apply_opt(csb_mlp* h, int rule, float lr, float beta1, float beta2, float eps, float wd, void* stream);

/* End-to-end step from HOST buffers (pinned recommended): H2D copies of x,y + train_step + apply_opt + D2H of the
 * loss; synchronises; *loss_host receives the scalar.  This is what a model.fit batch costs a host-side caller. */
int  csb_mlp_train_step_host(csb_mlp* h, const float* x_host, const float* y_host, int64_t B, float grad_scale,
                             uint32_t flags, int rule, float lr, float beta1, float beta2, float eps, float wd,
                             float* loss_host, void* stream);

/* Pipelined host feed -- the engine-side counterpart of the reference's input prefetch (tf.data .prefetch at
 * baseline_models/MLP/.../step2_retrain.py:266-276; DataLoader(pin_memory) + non-blocking copies in
 * online_testing/.../train_mlp_h5loader.py:126-150).
 * csb_mlp_stage_host_batch: enqueue the H2D copies of (x_host, y_host) into one of two internal staging slots on an internal
 *   copy stream and make `stream` wait for them; *x_dev / *y_dev receive the slot's device pointers (y_host and y_dev may both
 *   be NULL).  Blocks the host only while the slot is still being read by the step staged two calls earlier, so the copy of
 *   batch i+1 overlaps the compute of batch i and a caller may reuse its host buffers after two further calls.
 * csb_mlp_release_staged: call once the work reading the staged batch has been enqueued on `stream`.
 * csb_mlp_train_step_host_async: csb_mlp_train_step_host without the final synchronisation; *loss_host (pinned memory) is
 *   written by a D2H copy enqueued on `stream` and is valid once the stream has passed it. */
int  csb_mlp_stage_host_batch(csb_mlp* h, const float* x_host, const float* y_host, int64_t B, void* stream, float** x_dev, float** y_dev);
int  csb_mlp_release_staged(csb_mlp* h, void* stream);
int  csb_mlp_train_step_host_async(csb_mlp* h, const float* x_host, const float* y_host, int64_t B, float grad_scale,
                                   uint32_t flags, int rule, float lr, float beta1, float beta2, float eps, float wd,
                                   float* loss_host, void* stream);

/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
int64_t csb_mlp_launch_count(const csb_mlp* h);

/* Optional per-launch device timing (CUDA events on the caller's stream after every kernel launch).
 * csb_mlp_profile(h, 1) resets and enables, csb_mlp_profile_read() synchronises and returns, per kernel kind, the
 * accumulated milliseconds and launch count since the last enable (arrays of csb_profile_kind_count() entries). */
int  csb_mlp_profile(csb_mlp* h, int enable);
int  csb_mlp_profile_read(csb_mlp* h, double* ms_by_kind, int64_t* launches_by_kind, int n_kinds);
int  csb_profile_kind_count(void);
const char* csb_profile_kind_name(int kind);

/* ---- CNN family ----------------------------------------------------------------------------------------- */
/* ResNet-1D of baseline_models/CNN/training/hpo_train.py:131-200: depth x { Conv1D(width,k,'same') -> act -> Conv1D(width,k,
 * 'same') -> act -> + Conv1D(width,1)(block input) } -> Conv1D(out_ch,1,pre_out_act) -> per-level Dense(out_lin, linear) ||
 * Dense(out_ch-out_lin, relu).  Tensors are channels-last fp32: x (B,levels,in_ch), y / predictions (B,levels,out_ch)
 * (what data_utils.reshape_input_for_cnn / reshape_target_for_cnn produce, data_utils.py:1693-1738).  Dropout is inference
 * mode (p = 0).  Flat parameter blob = Keras get_weights() order: per block Wc1 (k,Cin,Cout), bc1, Wc2, bc2, Wres (1,Cin,Cout),
 * bres; then Wout (1,width,out_ch), bout, then the two Dense heads fused column-wise: W (out_ch,out_ch) = [W_lin | W_relu], b.
 * Loss = batch mean of sum_{level,channel} w_c * e (e = squared or absolute error); the default w reproduces the reference's
 * mse_adjusted / mae_adjusted (hpo_train.py:114-121). */
typedef struct csb_cnn_cfg {
  int32_t depth;        /* 12 */
  int32_t width;        /* 406 */
  int32_t kernel;       /* 3 */
  int32_t in_ch;        /* 6 */
  int32_t out_ch;       /* 10 */
  int32_t out_lin;      /* 2 */
  int32_t levels;       /* 60 */
  int32_t act;          /* CSB_ACT_RELU */
  int32_t pre_out_act;  /* CSB_ACT_ELU */
  int32_t dtype;        /* CSB_BF16 (the only mode implemented for the CNN) */
  int32_t loss;         /* CSB_LOSS_MAE (final reference configuration) | CSB_LOSS_MSE */
  int64_t max_batch;
} csb_cnn_cfg;
typedef struct csb_cnn csb_cnn;

int  csb_cnn_create(const csb_cnn_cfg* cfg, csb_cnn** out);       /* CNNHyperModel.build, hpo_train.py:124-236 */
int  csb_cnn_destroy(csb_cnn* h);
size_t csb_cnn_param_count(const csb_cnn* h);                     /* 13 215 420 for the reference configuration */
int  csb_cnn_set_params(csb_cnn* h, const float* params_host);
int  csb_cnn_get_params(csb_cnn* h, float* params_host);
int  csb_cnn_get_grads(csb_cnn* h, float* grads_host);
int  csb_cnn_set_loss_weights(csb_cnn* h, const float* w_host);   /* out_ch per-channel weights */
/* keras.layers.Dropout(rate) behind the two ReLUs of every residual block (hpo_train.py:143,170,177: rate 0.175), active in
 * csb_cnn_train_step only (csb_cnn_forward is model.predict: no dropout).  Inverted dropout with a counter-based generator keyed by
 * (seed, training step, layer, position): reproducible, but not TensorFlow's random stream.  CSB_BF16 engines; 0 switches it off. */
int  csb_cnn_set_dropout(csb_cnn* h, float rate, uint32_t seed);
/* test hook: the hidden activation after conv1 (which = 1) / conv2 (which = 2) of residual block `block`, as left by the last
 * csb_cnn_train_step / csb_cnn_forward, unpacked to fp32 (B, levels, width) on the device */
int  csb_cnn_debug_read_hidden(csb_cnn* h, int which, int block, float* dst, int64_t B, void* stream);
int  csb_cnn_forward(csb_cnn* h, const float* x, float* y_pred, int64_t B, void* stream);                  /* model.predict */
int  csb_cnn_train_step(csb_cnn* h, const float* x, const float* y, int64_t B, float grad_scale, float* loss_out, void* stream);
int  csb_cnn_grad_buffer(csb_cnn* h, float** ptr, size_t* n);
/* autograd entry (the reference's CNN is a Keras model driven by model.fit; a torch.nn.Module face needs backward for an arbitrary
 * upstream gradient): after csb_cnn_forward(h, x, y_pred, B) on the same batch, csb_cnn_backward takes dL/d(y_pred) (B, levels, out_ch)
 * and leaves every parameter gradient in the gradient buffer (no dropout, as in csb_cnn_forward; dL/dx is not produced).
 * csb_cnn_set_params_device / csb_cnn_get_grads_device move the flat unpadded blobs without leaving the device. */
int  csb_cnn_backward(csb_cnn* h, const float* y_pred, const float* dy, int64_t B, void* stream);
int  csb_cnn_set_params_device(csb_cnn* h, const float* params_dev, void* stream);
int  csb_cnn_get_grads_device(csb_cnn* h, float* grads_dev, void* stream);
int  csb_cnn_apply_opt(csb_cnn* h, int rule, float lr, float beta1, float beta2, float eps, float wd, void* stream);
int64_t csb_cnn_launch_count(const csb_cnn* h);

/* ---- data_utils device helpers -------------------------------------------------------------------------- */
/* (x - sub)/div, inf/nan -> 0 on (N,F) fp32; data_utils.py:806-809,894-897. */
int  csb_normalize(const float* x_raw, const float* sub, const float* div, float* x_out, int64_t N, int32_t F, void* stream);
/* CNN tensor layout helpers, data_utils.py:1693-1760: (N,124)->(N,60,6), (N,128)->(N,60,10), (N,60,10)->(N,128). */
int  csb_reshape_input_for_cnn(const float* x, float* out, int64_t N, void* stream);
int  csb_reshape_target_for_cnn(const float* y, float* out, int64_t N, void* stream);
int  csb_reshape_target_from_cnn(const float* p, float* out, int64_t N, void* stream);

/* Fused evaluation (V1 variable set): denormalise -> dp/g vertical weighting -> area weighting -> energy units
 * (data_utils.output_weighting, data_utils.py:1112-1362) and the time-axis metrics MAE / RMSE / R2 / bias followed by the grid
 * mean (data_utils.calc_*, :1432-1497), in one pass over pred / target (N,128) fp32 with fp64 accumulation.
 *   x_norm (N,124) is the (normalised) input array the pressure grid is built from (set_pressure_grid, :1037-1086);
 *   N must be a multiple of ncol (rows are ordered time-major, column fastest, as the reference reshapes them);
 *   consts (host, fp64): hyai[61], hybi[61], area_wgt[ncol], out_scale[128];  ps_* de-normalise state_ps.
 * out (device, fp64) [4][128]: rows MAE, RMSE, R2, bias per output index.  scratch (device) needs 4*128*ncol + ncol + 256 doubles. */
int  csb_eval_metrics(const float* pred, const float* target, const float* x_norm, int64_t N, int32_t ncol, const double* hyai,
                      const double* hybi, double p0, const double* area_wgt, const double* out_scale, double ps_mean, double ps_max,
                      double ps_min, int normalize, double* out, double* scratch, void* stream);

/* CRPS of an ensemble by the sorted-sample identity, averaged over time and grid (data_utils.calc_CRPS, data_utils.py:1499-1524):
 * samples [n_tc, L, S] and target [n_tc, L] on the device (fp32, or fp64 with is_f64 != 0; n_tc = time x grid, L = 60 levels or 1,
 * 2 <= S <= 32 members contiguous); out (device, fp64) [L]; scratch (device) needs L * CSB_CRPS_BLOCKS doubles.
 * fp64 arithmetic, fixed summation order (deterministic). */
#define CSB_CRPS_BLOCKS 64
int  csb_eval_crps(const void* samples, const void* target, int is_f64, int64_t n_tc, int L, int S, double* out, double* scratch, void* stream);

/* ---- fit-loop helpers ------------------------------------------------------------------------------------------------------------- */
/* Sufficient statistics of Keras' `metrics=['mse','mae','accuracy']` (hpo_baseline_v1.py:127-129; step2_retrain.py:262 logs them and
 * their val_ twins through CSVLogger) for one batch: pred, y device fp32 [B, F] dense.
 * out5 (device, fp64) = {sum (p-y)^2, sum |p-y|, rows with argmax(p) == argmax(y) [first maximum, as tf.argmax], B*F, B}.
 * scratch (device) needs CSB_BATCH_METRICS_SCRATCH doubles, zero before the first use (the kernel re-arms it).  Deterministic. */
#define CSB_BATCH_METRICS_SCRATCH 2048
int  csb_batch_metrics(const float* pred, const float* y, int64_t B, int32_t F, double* out5, double* scratch, void* stream);

/* One training step of the heteroskedastic-regression model -- BOTH networks, loss, backward, optimizer -- as
 * HeteroskedasticRegression.trainer runs it per batch (baseline_models/HSR/training/hsr.py:122-140):
 *   mle == 0: loss = mean((y - mu)^2) (the first third of the epochs; the log-precision network gets no gradient and, like
 *             torch.optim.Adam with grad None, no update);   mle != 0: loss = mean(exp(lp) (y - mu)^2 - lp);
 *   torch.clip(loss, -1e5, 1e5).backward(); per-group L2 weight decay (wd_mean = alpha, wd_logprec = beta, hsr.py:100-107) with
 *   `rule` (CSB_OPT_ADAM_TORCH or CSB_OPT_SGD in the reference).
 * loss_out (device, fp32): the unclipped mean, what the reference appends to `losses`.  scratch: CSB_BATCH_METRICS_SCRATCH doubles,
 * zero before the first use.  `flags`: CSB_FWD_NORMALIZE_IN; CSB_HSR_NO_OPT stops after the backward passes with the gradients in the
 * handles' gradient buffers (data parallelism: all-reduce csb_mlp_grad_buffer of each network, then csb_mlp_apply_opt with its own
 * decay -- in the MSE phase only for `mean`).  Both handles: linear output layer, same widths / dtype. */
#define CSB_HSR_NO_OPT 16u
int  csb_hsr_train_step(csb_mlp* mean, csb_mlp* logprec, const float* x, const float* y, int64_t B, int mle, uint32_t flags, int rule,
                        float lr, float beta1, float beta2, float eps, float wd_mean, float wd_logprec, float* loss_out, double* scratch,
                        void* stream);

/* ---- data parallelism over NVLink peer memory -------------------------------------------------------------------------------------- */
/* The reference's one collective is DDP's gradient all-reduce (online_testing/baseline_models/MLP_v2rh/training/
 * train_mlp_h5loader.py:195-207).  Here it is fused with the optimizer into ONE kernel per rank that reads and writes its peers'
 * gradient slabs directly (reduce-scatter by peer loads, all-gather by peer stores, fixed rank order => bit-identical replicas):
 *   csb_mlp_dp_export   moves the gradient buffer into an IPC-exportable slab and returns its handle (CSB_IPC_HANDLE_BYTES bytes);
 *   csb_mlp_dp_attach   maps the peers' slabs: `ipc_handles` = world x CSB_IPC_HANDLE_BYTES bytes in rank order (gathered by the
 *                       caller, e.g. torch.distributed.all_gather_object); ranks must be GPUs of one node with peer access
 *                       (CSB_EUNSUPPORTED otherwise: fall back to csb_mlp_grad_buffer + ncclAllReduce + csb_mlp_apply_opt);
 *   csb_mlp_dp_step     after csb_mlp_train_step(..., CSB_TRAIN_FUSED_OPT) called with grad_scale = 1 / (global batch x out_dim) on
 *                       EVERY rank: partial reduction, cross-rank sum, optimizer rule and bf16 weight copies in one launch; loss_out
 *                       (device, optional) receives the loss of the GLOBAL batch.  All ranks must call it the same number of times. */
#define CSB_IPC_HANDLE_BYTES 64
int  csb_mlp_dp_export(csb_mlp* h, void* ipc_handle_out);
int  csb_mlp_dp_attach(csb_mlp* h, int rank, int world, const void* ipc_handles);
int  csb_mlp_dp_step(csb_mlp* h, int rule, float lr, float beta1, float beta2, float eps, float wd, float* loss_out, void* stream);
/* measurement hook: globaltimer (ns) of the last csb_mlp_dp_step at {start, gradient summed, all ranks ready, slice exchanged, all
 * ranks done, end} on this rank; synchronises the device. */
int  csb_mlp_dp_debug(csb_mlp* h, unsigned long long* stamps6_host);

/* ---- input pipeline ------------------------------------------------------------------------------------------------------------- */
/* dst[i, :] = src[idx[i], :] for i < n_rows (fp32 rows of row_len floats, device pointers, idx int64 on the device): the sample
 * shuffle of the reference's input pipelines -- tf.data `unbatch().shuffle(384*30).batch(B)` (hpo_baseline_v1.py:140-143,
 * step2_retrain.py:266-271) and DistributedSampler(shuffle=True) (online_testing/.../train_mlp_h5loader.py:126-134) -- done on the
 * device over a window of columns already resident in HBM.  Never synchronises; an index outside [0, src_rows) is skipped and
 * reported by the next csb_gather_rows_check (CSB_EINVAL), which synchronises the stream. */
int  csb_gather_rows(const float* src, const int64_t* idx, float* dst, int64_t n_rows, int row_len, int64_t src_rows, void* stream);
int  csb_gather_rows_check(void* stream);

/* ---- kernel self-test hooks (used by tests/test_gemm_gpu.py; device pointers) ----------------------------- */
/* C[M,N] (fp32) = A[M,K] * Bt[N,K]^T on the tcgen05 path (both operands K-major bf16, raw uint16 payloads). */
int  csb_test_gemm_tn(const uint16_t* A, const uint16_t* Bt, float* C, int M, int N, int K, int block_n, void* stream);
/* the same kernel on fp32 operands through tcgen05 kind::tf32 (block_n 128 or 512; K a multiple of 32) */
int  csb_test_gemm_tn_tf32(const float* A, const float* Bt, float* C, int M, int N, int K, int block_n, void* stream);
/* out[M,N] (bf16) = act(A[M,K] * Wt[N,K]^T + bias): one forward layer as the engine runs it.  pairs = 0: single-CTA tiles,
 * 1: cta_group::2 pairs (layers wider than 128), 2: the engine's own launch policy (pairs + the staged coalesced-store epilogue
 * for K <= 256).  For kernel micro-benchmarks (scripts/microbench_gemm.py) and tests/test_gemm_gpu.py. */
int  csb_test_linear_fwd(const uint16_t* A, const uint16_t* Wt, const float* bias, uint16_t* out, int M, int N, int K, int act,
                         float alpha, int pairs, void* stream);
void csb_test_set_debug(int flags);   /* epilogue ablation knobs for the micro-benchmark (0 = normal operation) */
void csb_test_set_stats(void* dev_u64x8_per_cta);   /* micro-benchmark: per-CTA {clock64 start, end, globaltimer start, end, MMA issuer cycles waiting for a
                                                      * free accumulator, ... for operands, epilogue warp 0 cycles waiting for an accumulator, ... working} of
                                                      * csb_test_linear_fwd; NULL = off */
/* C[s][M,N] (fp32 partials, s < splits) = A[Kr,M]^T * B[Kr,N]  (both operands MN-major bf16): the weight-gradient
 * contraction over rows; colsum (optional, [splits * ceil(M/128)][N]) receives partial column sums of B (bias gradient). */
int  csb_test_gemm_nt(const uint16_t* A, const uint16_t* B, float* C, float* colsum, int M, int N, int Kr, int splits, void* stream);
/* the same with an explicit tile policy: cg = 1 single CTAs, 2 CTA pairs (cta_group::2, needs N % 128 == 0), 0 = the engine's choice;
 * *m_tiles_out = bias-gradient partial rows written per split.  splits < 0: |splits| split slots with the uneven geometry the
 * engine uses when the last n-block has half width (N = k * 256 + 128): that block takes ceil(|splits| / 2) longer splits and its
 * CTAs zero the slots it leaves unused. */
int  csb_test_gemm_nt_cg(const uint16_t* A, const uint16_t* B, float* C, float* colsum, int M, int N, int Kr, int splits, int cg,
                         int* m_tiles_out, void* stream);

/* ---- misc ---------------------------------------------------------------------------------------------- */
int         csb_version(void);
const char* csb_strerror(int code);
const char* csb_last_error(void);
int         csb_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* hbm_bytes);

#ifdef __cplusplus
}
#endif
#endif /* CLIMSIM_B200_H */
