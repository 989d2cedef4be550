// Shared helpers for libte_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/te_b200.h"

namespace te {

void set_error(const char* fmt, ...);

#define TE_CHECK_ARG(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      te::set_error(__VA_ARGS__);        \
      return TE_ERR_INVALID;             \
    }                                    \
  } while (0)

#define TE_CHECK_CUDA(expr)                                                             \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      te::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                    __LINE__);                                                          \
      return TE_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

// Checks the launch itself (configuration errors); never synchronises.
#define TE_CHECK_LAUNCH()                                                               \
  do {                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      te::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),         \
                    __FILE__, __LINE__);                                                \
      return TE_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// accumulate type: float for everything except double
template <typename T> struct Acc { using type = float; };
template <> struct Acc<double> { using type = double; };

template <typename T> __device__ __forceinline__ typename Acc<T>::type to_acc(T v) {
  return static_cast<typename Acc<T>::type>(v);
}
template <> __device__ __forceinline__ float to_acc<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <> __device__ __forceinline__ float to_acc<__half>(__half v) { return __half2float(v); }

template <typename T, typename A> __device__ __forceinline__ T from_acc(A v) {
  return static_cast<T>(v);
}
template <> __device__ __forceinline__ __nv_bfloat16 from_acc<__nv_bfloat16, float>(float v) {
  return __float2bfloat16_rn(v);
}
template <> __device__ __forceinline__ __half from_acc<__half, float>(float v) {
  return __float2half_rn(v);
}

// 16-byte vector of T
template <typename T> struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
  union {
    uint4 raw;
    T v[N];
  };
};

inline int grid_for(int64_t work_items, int threads, int max_waves = 16) {
  int64_t blocks = (work_items + threads - 1) / threads;
  int64_t cap = static_cast<int64_t>(kNumSMs) * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace te
