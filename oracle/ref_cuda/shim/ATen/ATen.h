// TEST INFRASTRUCTURE ONLY - minimal stand-in for <ATen/ATen.h> so that the reference's CUDA sources
// (crop_split_cuda_kernel.cu, deform_conv_cuda_kernel.cu; written against PyTorch 1.1) compile with plain nvcc,
// from where they lie under /root/reference, without PyTorch.  No arithmetic lives here: at::Tensor is a raw
// device pointer, the dispatch macro instantiates the reference's lambda for scalar_t = float.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
namespace at {
struct Half {};
struct Tensor {
  void* ptr;
  Tensor() : ptr(nullptr) {}
  explicit Tensor(const void* p) : ptr(const_cast<void*>(p)) {}
  template <typename T> T* data() const { return static_cast<T*>(ptr); }
  template <typename T> T* data_ptr() const { return static_cast<T*>(ptr); }
  int type() const { return 0; }
  int scalar_type() const { return 0; }
};
}  // namespace at
#define AT_DISPATCH_FLOATING_TYPES_AND_HALF(TYPE, NAME, ...) \
  {                                                          \
    (void)(TYPE);                                            \
    using scalar_t = float;                                  \
    __VA_ARGS__();                                           \
  }
