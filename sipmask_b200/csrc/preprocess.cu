// Test-time image pipeline on the device: uint8 BGR HWC image -> keep-ratio bilinear resize ->
// mean subtraction -> zero pad -> the stem's NHWC8 fp16 layout, in one kernel.
//
// Reference: Resize(keep_ratio) / Normalize(std=1, to_rgb=False) / Pad(32) / ImageToTensor on the host CPU
// (SipMask-mmdetection/mmdet/datasets/pipelines/transforms.py:97-110, 274-300, 335-363), i.e. mmcv -> cv2.resize(INTER_LINEAR),
// followed by a 12.9 MB fp32 upload.  Here the upload is the raw uint8 image (<= 4 MB) and the arithmetic is OpenCV's
// 8-bit fixed-point bilinear restated exactly (11-bit coefficients, int32 horizontal pass, the two truncating shifts of
// the vertical pass), so the stem sees bit-identical input.
#include <stdint.h>

#include <cuda_fp16.h>

#include "common.cuh"

namespace smb {

__device__ __forceinline__ uint4 pack8h(const float* f) {
  uint4 o;
  __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return o;
}

struct Coef {
  int s0, s1, a0, a1;
};

// OpenCV resize.cpp (INTER_LINEAR, 8U): fx = (float)((d + 0.5) * scale - 0.5); s = floor(fx); fx -= s; the x direction
// clamps the coefficient at the borders, the y direction only clamps the row index.
__device__ __forceinline__ Coef coef(int d, double scale, int sn, bool clamp_coef) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (clamp_coef) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= sn - 1) { f = 0.f; s = sn - 1; }
  }
  Coef c;
  c.a1 = __float2int_rn(__fmul_rn(f, 2048.f));
  c.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  c.s0 = min(max(s, 0), sn - 1);
  c.s1 = min(max(s + 1, 0), sn - 1);
  return c;
}

// out [H+6, W+8, 8] fp16: pixel (y, x) at (y+3, x+3); zeros in the halo, the padding and channels 3..7
__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ src, int sh, int sw, int pitch, int dh, int dw, double scale_y,
                                     double scale_x, float m0, float m1, float m2, __half* __restrict__ out, int H, int W) {
  const int Hp = H + 6, Wp = W + 8;
  const long long total = (long long)Hp * Wp;
  const bool identity = (sh == dh) && (sw == dw);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)blockDim.x * gridDim.x) {
    const int xp = (int)(t % Wp), yp = (int)(t / Wp);
    const int x = xp - 3, y = yp - 3;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (x >= 0 && x < dw && y >= 0 && y < dh) {
      int v[3];
      if (identity) {                                  // cv2.resize copies when the size does not change
        const uint8_t* p = src + (size_t)y * pitch + x * 3;
        v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
      } else {
        const Coef cx = coef(x, scale_x, sw, true), cy = coef(y, scale_y, sh, false);
        const uint8_t* r0 = src + (size_t)cy.s0 * pitch;
        const uint8_t* r1 = src + (size_t)cy.s1 * pitch;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int h0 = (int)r0[cx.s0 * 3 + c] * cx.a0 + (int)r0[cx.s1 * 3 + c] * cx.a1;
          const int h1 = (int)r1[cx.s0 * 3 + c] * cx.a0 + (int)r1[cx.s1 * 3 + c] * cx.a1;
          v[c] = (((cy.a0 * (h0 >> 4)) >> 16) + ((cy.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
        }
      }
      f[0] = __fsub_rn((float)v[0], m0);
      f[1] = __fsub_rn((float)v[1], m1);
      f[2] = __fsub_rn((float)v[2], m2);
    }
    *reinterpret_cast<uint4*>(out + t * 8) = pack8h(f);
  }
}

}  // namespace smb

using namespace smb;

extern "C" int smb_preprocess_u8(const uint8_t* src, int src_h, int src_w, int src_pitch_bytes, int dst_h, int dst_w,
                                 const float* host_mean3, void* out_nhwc8, int H, int W, smb_stream_t stream) {
  SMB_CHECK_ARG(src && host_mean3 && out_nhwc8, "smb_preprocess_u8: null pointer");
  SMB_CHECK_ARG(src_h > 0 && src_w > 0 && src_pitch_bytes >= 3 * src_w && dst_h > 0 && dst_w > 0 && dst_h <= H && dst_w <= W,
                "smb_preprocess_u8: bad shape (src %dx%d -> %dx%d inside %dx%d)", src_h, src_w, dst_h, dst_w, H, W);
  const double scale_y = 1.0 / ((double)dst_h / src_h), scale_x = 1.0 / ((double)dst_w / src_w);
  const long long total = (long long)(H + 6) * (W + 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  preprocess_u8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, src_h, src_w, src_pitch_bytes, dst_h, dst_w, scale_y, scale_x,
                                                                 host_mean3[0], host_mean3[1], host_mean3[2], (__half*)out_nhwc8, H, W);
  SMB_LAUNCH_OK("preprocess_u8_kernel");
  return SMB_OK;
}
