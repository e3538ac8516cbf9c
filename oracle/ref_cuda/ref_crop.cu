// TEST INFRASTRUCTURE ONLY.  Compiles the reference's CropSplit CUDA source from where it lies
// (REF_SRC = /root/reference/SipMask-mmdetection/mmdet/ops/crop/src/crop_split_cuda_kernel.cu, passed by
// oracle/build.py) and exports its own launcher CropSplitForward (:62-88, kernel :19-59) behind a C entry point,
// so that the GPU tests can pin both the oracle's restatement and the smb kernels against the reference kernel
// itself.  Nothing of the reference is copied into the repository.
#include REF_SRC

extern "C" int ref_crop_split_forward(const float* data, const float* rois, float* out, int H, int W, int c, int N) {
  // the caller zero-initialises `out` like CropSplitFunction.forward does (ops/crop/crop_split.py:22)
  CropSplitForward(at::Tensor(data), at::Tensor(rois), at::Tensor(out), H, W, c, N);
  return (int)cudaDeviceSynchronize();
}

extern "C" int ref_crop_split_backward(const float* top_grad, const float* rois, float* bottom_grad, int H, int W, int c, int N) {
  // the caller zero-initialises `bottom_grad` like CropSplitFunction.backward does (ops/crop/crop_split.py:35)
  CropSplitBack(at::Tensor(top_grad), at::Tensor(rois), at::Tensor(bottom_grad), H, W, c, N);
  return (int)cudaDeviceSynchronize();
}
