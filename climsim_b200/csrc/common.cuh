// common.cuh -- shared definitions for the climsim_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/climsim_b200.h"

namespace csb {

// ---- error plumbing -------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);

#define CSB_CUDA_CHECK(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      ::csb::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CSB_ECUDA;                                                                            \
    }                                                                                              \
  } while (0)

#define CSB_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      ::csb::set_last_error(__VA_ARGS__);       \
      return (code);                            \
    }                                           \
  } while (0)

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- activations (Keras layer semantics; see oracle/models.py::activation) -------------------------------------
__device__ __forceinline__ float act_fwd(int act, float alpha, float z) {
  switch (act) {
    case CSB_ACT_RELU: return z > 0.f ? z : 0.f;
    case CSB_ACT_LEAKYRELU: return z > 0.f ? z : alpha * z;
    case CSB_ACT_ELU: return z > 0.f ? z : expm1f(z);
    default: return z;
  }
}
// derivative expressed through the saved OUTPUT a = act(z) (TF ReluGrad / LeakyReluGrad / EluGrad do the same):
//   relu: a > 0 ? 1 : 0      leaky: a > 0 ? 1 : alpha      elu(alpha=1): a > 0 ? 1 : a + 1      (a == 0 -> z == 0)
__device__ __forceinline__ float act_bwd_from_out(int act, float alpha, float a) {
  switch (act) {
    case CSB_ACT_RELU: return a > 0.f ? 1.f : 0.f;
    case CSB_ACT_LEAKYRELU: return a > 0.f ? 1.f : alpha;
    case CSB_ACT_ELU: return a > 0.f ? 1.f : a + 1.f;
    default: return 1.f;
  }
}

// ---- programmatic dependent launch ------------------------------------------------------------------------------
// Every kernel of the training step is launched with cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may
// start (barrier setup, TMEM allocation, descriptor prefetch) while the previous kernel drains, and pdl_wait() blocks
// until that kernel has completed and its memory is visible.  pdl_wait() precedes the first global-memory access of
// every such kernel, which also makes completion transitive along the stream.  Both are no-ops in a normal launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

}  // namespace csb
