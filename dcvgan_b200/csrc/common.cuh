// Shared helpers for libdcvgan_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include "../../include/dcvgan_b200.h"

namespace dcv {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define DCV_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      dcv::set_error(__VA_ARGS__);        \
      return -1;                          \
    }                                     \
  } while (0)

#define DCV_CUDA(expr)                                                            \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      dcv::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));            \
      return -2;                                                                  \
    }                                                                             \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == DCV_ACT_LEAKY) return v > 0.f ? v : v * slope;
  if (act == DCV_ACT_TANH) return tanhf(v);
  return v;
}
// derivative expressed through the activated output a (sign(a) == sign(pre-activation) for slope >= 0)
__device__ __forceinline__ float act_grad_from_out(float a, int act, float slope) {
  if (act == DCV_ACT_LEAKY) return a > 0.f ? 1.f : slope;
  if (act == DCV_ACT_TANH) return 1.f - a * a;
  return 1.f;
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// A training iteration is ~390 dependent launches.  Every kernel of this library is launched through launch_k() and starts
// with pdl_wait() (griddepcontrol.wait: returns when the preceding kernel has completed and flushed) before it touches
// global memory, followed by pdl_trigger() (griddepcontrol.launch_dependents), so that with the programmatic-stream-
// serialization launch attribute the NEXT kernel's blocks could be scheduled while this one's last wave drains.
// MEASURED (r2d, mug-depth, batch 32, CUDA-graph replay): 11.21 ms per iteration with the attribute, 10.81 ms without -
// the early-launched blocks cost more than the hidden launch latency.  The attribute is therefore OFF by default
// (dcv_set_tuning("pdl", 1) enables it for experiments); without it the two device instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);      // errors are picked up by check_launch()
}

}  // namespace dcv
