// TEST INFRASTRUCTURE ONLY - stand-in for <ATen/cuda/CUDAContext.h>: the "current stream" is a settable global.
#pragma once
#include <cuda_runtime.h>
namespace at { namespace cuda {
inline cudaStream_t& ref_stream_slot() { static cudaStream_t s = 0; return s; }
inline cudaStream_t getCurrentCUDAStream() { return ref_stream_slot(); }
} }
