// TEST INFRASTRUCTURE ONLY.  Compiles the reference's CropSplitGt CUDA source from where it lies
// (REF_SRC = /root/reference/SipMask-mmdetection/mmdet/ops/crop/src/crop_split_gt_cuda_kernel.cu) and exports its launchers
// CropSplitGtForward (:51-73, kernel :19-49) and CropSplitGtBack (:106-, kernel :76-104) behind C entry points.
#include REF_SRC

extern "C" int ref_crop_split_gt_forward(const float* data, const float* rois, float* out, int H, int W, int c, int N) {
  CropSplitGtForward(at::Tensor(data), at::Tensor(rois), at::Tensor(out), H, W, c, N);     // caller zero-initialises out
  return (int)cudaDeviceSynchronize();
}

extern "C" int ref_crop_split_gt_backward(const float* top_grad, const float* rois, float* bottom_grad, int H, int W, int c, int N) {
  CropSplitGtBack(at::Tensor(top_grad), at::Tensor(rois), at::Tensor(bottom_grad), H, W, c, N);
  return (int)cudaDeviceSynchronize();
}
