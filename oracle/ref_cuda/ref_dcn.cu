// TEST INFRASTRUCTURE ONLY.  Compiles the reference's deformable-convolution CUDA source from where it lies
// (REF_SRC = /root/reference/SipMask-mmdetection/mmdet/ops/dcn/src/deform_conv_cuda_kernel.cu) and exports its
// launcher deformable_im2col (:245-277; kernel :190-243, bilinear :84-115) behind a C entry point.
//   im [B,C,H,W] fp32, offset [B,dg*2*k*k,Ho,Wo] fp32 -> col [C*k*k, B, Ho, Wo] fp32 (the layout
//   deform_conv_forward_cuda multiplies with weight.view(Cout, C*k*k), deform_conv_cuda.cpp:231-236)
#include REF_SRC

extern "C" int ref_deformable_im2col(const float* im, const float* offset, int C, int H, int W, int k, int pad,
                                     int stride, int dilation, int parallel_imgs, int dg, float* col) {
  deformable_im2col(at::Tensor(im), at::Tensor(offset), C, H, W, k, k, pad, pad, stride, stride, dilation, dilation,
                    parallel_imgs, dg, at::Tensor(col));
  return (int)cudaDeviceSynchronize();
}
