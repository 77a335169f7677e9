/* ffi_demo.c -- the C ABI from plain C (no C++, no Python, no torch): create an MLP_v1 engine, upload parameters, run one forward
 * and one training step on device buffers, read the loss back.
 *
 *   gcc -std=c99 -Wall -I include -I /usr/local/cuda/include examples/ffi_demo.c -L climsim_b200 -lclimsim_b200 \
 *       -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/climsim_b200 -o /tmp/ffi_demo && /tmp/ffi_demo
 *
 * This is what a binding in any host language does (INTEGRATION.md shows the ctypes version); it needs a B200 to run and is
 * compiled (not run) by tests/test_abi_cpu.py to keep the header honest C. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "climsim_b200.h"

#define CHECK(call)                                                                              \
  do {                                                                                           \
    int rc_ = (call);                                                                            \
    if (rc_ != CSB_OK) {                                                                         \
      fprintf(stderr, "%s -> %d (%s): %s\n", #call, rc_, csb_strerror(rc_), csb_last_error());   \
      return 1;                                                                                  \
    }                                                                                            \
  } while (0)

int main(void) {
  const int B = 4096, IN = 124, OUT = 128;
  static const int units[7] = {768, 640, 512, 640, 640, 128, 128};
  csb_mlp_cfg cfg;
  csb_mlp* h = NULL;
  size_t n_params, i;
  float *params, *x_host, *y_host, *x = NULL, *y = NULL, *pred = NULL, *loss = NULL, loss_host = 0.f;
  int l;

  memset(&cfg, 0, sizeof cfg);
  cfg.in_dim = IN;
  cfg.n_layers = 7;
  for (l = 0; l < 7; ++l) {
    cfg.units[l] = units[l];
    cfg.act[l] = l < 6 ? CSB_ACT_LEAKYRELU : CSB_ACT_NONE;
    cfg.alpha[l] = 0.15f;
  }
  cfg.head_relu_from = 120; /* [120 linear | 8 ReLU] output head */
  cfg.dtype = CSB_BF16;
  cfg.loss = CSB_LOSS_MSE;
  cfg.max_batch = B;
  CHECK(csb_mlp_create(&cfg, &h));

  n_params = csb_mlp_param_count(h);
  params = (float*)malloc(n_params * sizeof(float));
  for (i = 0; i < n_params; ++i) params[i] = 0.02f * ((float)rand() / (float)RAND_MAX - 0.5f);
  CHECK(csb_mlp_set_params(h, params));

  x_host = (float*)malloc((size_t)B * IN * sizeof(float));
  y_host = (float*)malloc((size_t)B * OUT * sizeof(float));
  for (i = 0; i < (size_t)B * IN; ++i) x_host[i] = 0.4f * ((float)rand() / (float)RAND_MAX - 0.5f);
  for (i = 0; i < (size_t)B * OUT; ++i) y_host[i] = 0.2f * ((float)rand() / (float)RAND_MAX - 0.5f);
  if (cudaMalloc((void**)&x, (size_t)B * IN * 4) || cudaMalloc((void**)&y, (size_t)B * OUT * 4) ||
      cudaMalloc((void**)&pred, (size_t)B * OUT * 4) || cudaMalloc((void**)&loss, 4)) {
    fprintf(stderr, "cudaMalloc failed\n");
    return 1;
  }
  cudaMemcpy(x, x_host, (size_t)B * IN * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(y, y_host, (size_t)B * OUT * 4, cudaMemcpyHostToDevice);

  CHECK(csb_mlp_forward(h, x, pred, B, 0, NULL));
  for (l = 0; l < 3; ++l) {
    CHECK(csb_mlp_train_step(h, x, y, B, 0.f, CSB_TRAIN_FUSED_OPT, loss, NULL));
    CHECK(csb_mlp_apply_opt(h, CSB_OPT_ADAM_KERAS, 1e-3f, 0.9f, 0.999f, 1e-7f, 0.f, NULL));
    cudaMemcpy(&loss_host, loss, 4, cudaMemcpyDeviceToHost);
    printf("step %d: loss %.6f\n", l, loss_host);
  }
  CHECK(csb_mlp_destroy(h));
  cudaFree(x); cudaFree(y); cudaFree(pred); cudaFree(loss);
  free(params); free(x_host); free(y_host);
  return 0;
}
