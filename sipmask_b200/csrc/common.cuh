// Shared helpers for the sipmask_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/sipmask_b200.h"

namespace smb {

void set_error(const char* fmt, ...);

#define SMB_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      smb::set_error(__VA_ARGS__);               \
      return SMB_EINVAL;                         \
    }                                            \
  } while (0)

#define SMB_CUDA_OK(expr)                                                            \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      smb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SMB_ECUDA;                                                              \
    }                                                                                \
  } while (0)

#define SMB_LAUNCH_OK(name)                                                          \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      smb::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));       \
      return SMB_ECUDA;                                                              \
    }                                                                                \
  } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// GroupNorm statistics are accumulated as 64-bit fixed point (deterministic integer atomics)
constexpr float kGnSumScale = 1048576.0f;   // 2^20
constexpr float kGnSqScale = 65536.0f;      // 2^16

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

}  // namespace smb
