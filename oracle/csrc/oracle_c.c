/* Plain-C restatement of the reference arithmetic for the post-processing part of
 * the SipMask hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 * Compiled with -ffp-contract=off so that float expressions round like the
 * reference's separate CUDA/C++ float operations.
 *
 * Citations relative to /root/reference/SipMask-mmdetection/mmdet/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ops/nms/src/nms_kernel.cu:14-22 (devIoU, +1 legacy areas) */
static float iou_plus(const float* a, const float* b, float one) {
  float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  float width = fmaxf(right - left + one, 0.f), height = fmaxf(bottom - top + one, 0.f);
  float interS = width * height;
  float Sa = (a[2] - a[0] + one) * (a[3] - a[1] + one);
  float Sb = (b[2] - b[0] + one) * (b[3] - b[1] + one);
  return interS / (Sa + Sb - interS);
}

typedef struct { float s; int64_t i; } key_t_;
static int cmp_desc(const void* pa, const void* pb) {
  const key_t_* a = (const key_t_*)pa; const key_t_* b = (const key_t_*)pb;
  if (a->s > b->s) return -1;
  if (a->s < b->s) return 1;
  return (a->i > b->i) - (a->i < b->i);      /* stable: lower original index first */
}

/* Greedy NMS.  dets [n,5]; keep_out [n] receives ORIGINAL indices ascending
 * (ops/nms/src/nms_kernel.cu:120-138; ops/nms/src/nms_cpu.cpp:33-59).  Returns count. */
int64_t oracle_nms(const float* dets, int64_t n, float thr, int cmp_ge, int plus_one, int64_t* keep_out) {
  if (n <= 0) return 0;
  key_t_* order = (key_t_*)malloc(sizeof(key_t_) * n);
  uint8_t* sup = (uint8_t*)calloc(n, 1);
  float one = plus_one ? 1.f : 0.f;
  for (int64_t i = 0; i < n; ++i) { order[i].s = dets[i * 5 + 4]; order[i].i = i; }
  qsort(order, n, sizeof(key_t_), cmp_desc);
  for (int64_t a = 0; a < n; ++a) {
    int64_t i = order[a].i;
    if (sup[i]) continue;
    for (int64_t b = a + 1; b < n; ++b) {
      int64_t j = order[b].i;
      if (sup[j]) continue;
      float ovr = iou_plus(dets + i * 5, dets + j * 5, one);
      if (cmp_ge ? (ovr >= thr) : (ovr > thr)) sup[j] = 1;
    }
  }
  int64_t k = 0;
  for (int64_t i = 0; i < n; ++i) if (!sup[i]) keep_out[k++] = i;
  free(order); free(sup);
  return k;
}

/* Mask assembly = 4x (protos @ cof_k^T) -> sigmoid -> CropSplit, fused per output element.
 * models/anchor_heads/sipmask_head.py:609-626 + ops/crop/src/crop_split_cuda_kernel.cu:35-56.
 * protos [32,H,W] (CHW, as `feat_mask`), cofs [N,128], rois [N,4] (already in prototype
 * coordinates), out [N,H,W] (the reference's [H,W,N] permuted to [N,H,W], sipmask_head.py:627).
 * The dot product accumulates in float in channel order 0..31. */
void oracle_mask_assemble(const float* protos, const float* cofs, const float* rois,
                          int H, int W, int N, float* out) {
  for (int n = 0; n < N; ++n) {
    const float x1 = rois[n * 4 + 0], y1 = rois[n * 4 + 1], x2 = rois[n * 4 + 2], y2 = rois[n * 4 + 3];
    const float roi_w = (float)(((double)(x2 - x1) + 0.1) / 2);
    const float roi_h = (float)(((double)(y2 - y1) + 0.1) / 2);
    for (int h = 0; h < H; ++h)
      for (int w = 0; w < W; ++w) {
        float v = 0.f;
        if (((float)w >= x1) & ((float)h >= y1) & ((float)w < x2) & ((float)h < y2)) {
          int idx_w = (int)(((float)w - x1) / roi_w);
          int idx_h = (int)(((float)h - y1) / roi_h);
          int cell = idx_h * 2 + idx_w;
          if (cell < 0) cell = 0;
          if (cell > 3) cell = 3;
          const float* c = cofs + (size_t)n * 128 + cell * 32;
          float acc = 0.f;
          for (int k = 0; k < 32; ++k) acc += protos[((size_t)k * H + h) * W + w] * c[k];
          v = 1.f / (1.f + expf(-acc));
        }
        out[((size_t)n * H + h) * W + w] = v;
      }
  }
}

/* x2 bilinear upsample (align_corners=False) + threshold -> uint8.
 * sipmask_head.py:630-633 with scale_factor == 1 (integer x2).  in [N,H,W] -> out [N,2H,2W]. */
void oracle_upsample2_thresh(const float* in, int N, int H, int W, float thr, uint8_t* out) {
  const int Ho = 2 * H, Wo = 2 * W;
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < Ho; ++y) {
      float sy = ((float)y + 0.5f) * 0.5f - 0.5f; if (sy < 0.f) sy = 0.f;
      int y0 = (int)sy; int y1 = y0 + (y0 < H - 1 ? 1 : 0); float ly = sy - (float)y0, hy = 1.f - ly;
      for (int x = 0; x < Wo; ++x) {
        float sx = ((float)x + 0.5f) * 0.5f - 0.5f; if (sx < 0.f) sx = 0.f;
        int x0 = (int)sx; int x1 = x0 + (x0 < W - 1 ? 1 : 0); float lx = sx - (float)x0, hx = 1.f - lx;
        const float* p = in + (size_t)n * H * W;
        float v = hy * (hx * p[y0 * W + x0] + lx * p[y0 * W + x1]) + ly * (hx * p[y1 * W + x0] + lx * p[y1 * W + x1]);
        out[((size_t)n * Ho + y) * Wo + x] = v > thr ? 1 : 0;
      }
    }
}

/* CropSplit forward, c == 2 (ops/crop/src/crop_split_cuda_kernel.cu:19-59).
 * data [4,H,W,N], rois [N,4], out [H,W,N] (zero outside the boxes, crop_split.py:22). */
void oracle_crop_split(const float* data, const float* rois, int H, int W, int N, float* out) {
  const size_t count = (size_t)H * W * N;
  float* rw = (float*)malloc(sizeof(float) * N);
  float* rh = (float*)malloc(sizeof(float) * N);
  for (int n = 0; n < N; ++n) {
    rw[n] = (float)(((double)(rois[n * 4 + 2] - rois[n * 4 + 0]) + 0.1) / 2);
    rh[n] = (float)(((double)(rois[n * 4 + 3] - rois[n * 4 + 1]) + 0.1) / 2);
  }
  for (int ph = 0; ph < H; ++ph)
    for (int pw = 0; pw < W; ++pw) {
      const size_t base = ((size_t)ph * W + pw) * N;
      for (int n = 0; n < N; ++n) {
        const float x1 = rois[n * 4 + 0], y1 = rois[n * 4 + 1], x2 = rois[n * 4 + 2], y2 = rois[n * 4 + 3];
        float v = 0.f;
        if (((float)pw >= x1) & ((float)ph >= y1) & ((float)pw < x2) & ((float)ph < y2)) {
          int idx_w = (int)(((float)pw - x1) / rw[n]);
          int idx_h = (int)(((float)ph - y1) / rh[n]);
          int cell = idx_h * 2 + idx_w;
          if (cell < 0) cell = 0;
          if (cell > 3) cell = 3;
          v = data[(size_t)cell * count + base + n];
        }
        out[base + n] = v;
      }
    }
  free(rw); free(rh);
}
