// Fused SipMask mask assembly for sm_100a.
//
// Reference (SipMask-mmdetection/mmdet/):
//   models/anchor_heads/sipmask_head.py:609-633   4x sgemm -> 4x sigmoid -> stack -> CropSplit -> permute
//   ops/crop/src/crop_split_cuda_kernel.cu:19-59  CropSplitKernelForward
// The reference materialises [4,H,W,N] sigmoid maps (~2.5 GB of HBM traffic for 142 MB of
// algorithmic bytes).  Here one kernel reads every prototype pixel once into
// registers, and for each detection evaluates only the ONE 32-term dot product selected by the
// CropSplit cell of that pixel, and only for pixels inside the box; everything else is a zero store.
// HBM-bound: algorithmic bytes = protos (H*W*32*sizeof) + out (N*H*W*sizeof).
#include "common.cuh"

namespace smb {

struct __align__(16) BoxP {
  float x1, y1, x2, y2, roi_w, roi_h, pad0, pad1;
};

constexpr int MA_THREADS = 128;
constexpr int MA_TW = 64;   // tile width  (16 threads x 4 pixels)
constexpr int MA_TH = 8;    // tile height
constexpr int MA_NB = 64;   // detections per shared-memory chunk

template <typename PT>
__device__ __forceinline__ float to_f(PT v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename OT>
__device__ __forceinline__ void store4(OT* p, const float* o, bool vec, int nvalid);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float* o, bool vec, int nvalid) {
  if (vec) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
    for (int i = 0; i < nvalid; ++i) p[i] = o[i];
  }
}
template <>
__device__ __forceinline__ void store4<__half>(__half* p, const float* o, bool vec, int nvalid) {
  if (vec) {
    __half2 a = __floats2half2_rn(o[0], o[1]);
    __half2 b = __floats2half2_rn(o[2], o[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
  } else {
    for (int i = 0; i < nvalid; ++i) p[i] = __float2half_rn(o[i]);
  }
}

// grid: (ceil(W/64), ceil(H/8)), block 128.
template <typename PT, bool HWC, typename OT>
__global__ void __launch_bounds__(MA_THREADS) mask_assemble_kernel(
    const PT* __restrict__ protos, const float* __restrict__ cofs, const float* __restrict__ boxes,
    float sx1, float sy1, float sx2, float sy2, OT* __restrict__ out, int H, int W, int N) {
  __shared__ __align__(16) float s_cof[MA_NB * 128];
  __shared__ BoxP s_box[MA_NB];

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int h = blockIdx.y * MA_TH + ty;
  const int w0 = blockIdx.x * MA_TW + tx * 4;
  const bool row_ok = h < H;
  const int nvalid = row_ok ? max(0, min(4, W - w0)) : 0;
  const bool vec = (nvalid == 4) && ((W & 3) == 0);

  // ---- prototype pixels -> registers (fp32), read exactly once
  float P[4][32];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int k = 0; k < 32; ++k) P[p][k] = 0.f;
  if (nvalid > 0) {
    if (HWC) {
      const PT* base = protos + ((size_t)h * W + w0) * 32;
      if (sizeof(PT) == 2) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          if (p < nvalid) {
            const uint4* q = reinterpret_cast<const uint4*>(base + p * 32);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              uint4 u = __ldg(q + v);
              const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 f = __half22float2(hh[e]);
                P[p][v * 8 + e * 2] = f.x;
                P[p][v * 8 + e * 2 + 1] = f.y;
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          if (p < nvalid) {
            const float4* q = reinterpret_cast<const float4*>(base + p * 32);
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              float4 f = __ldg(q + v);
              P[p][v * 4] = f.x; P[p][v * 4 + 1] = f.y; P[p][v * 4 + 2] = f.z; P[p][v * 4 + 3] = f.w;
            }
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const PT* q = protos + ((size_t)k * H + h) * W + w0;
#pragma unroll
        for (int p = 0; p < 4; ++p)
          if (p < nvalid) P[p][k] = to_f<PT>(q[p]);
      }
    }
  }

  const float hf = (float)h;
  for (int n0 = 0; n0 < N; n0 += MA_NB) {
    const int nb = min(MA_NB, N - n0);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * 128; i += MA_THREADS) s_cof[i] = __ldg(cofs + (size_t)n0 * 128 + i);
    if (threadIdx.x < nb) {
      const float* b = boxes + (size_t)(n0 + threadIdx.x) * 4;
      BoxP bp;
      bp.x1 = b[0] * sx1; bp.y1 = b[1] * sy1; bp.x2 = b[2] * sx2; bp.y2 = b[3] * sy2;
      // (roi_x2-roi_x1+0.1)/num_cell: float difference, then double (0.1 is a double literal),
      // narrowed to float on assignment (crop_split_cuda_kernel.cu:46-47).
      bp.roi_w = (float)(((double)(bp.x2 - bp.x1) + 0.1) / 2);
      bp.roi_h = (float)(((double)(bp.y2 - bp.y1) + 0.1) / 2);
      bp.pad0 = bp.pad1 = 0.f;
      s_box[threadIdx.x] = bp;
    }
    __syncthreads();
    if (nvalid == 0) continue;
    for (int j = 0; j < nb; ++j) {
      const BoxP b = s_box[j];
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      if ((hf >= b.y1) & (hf < b.y2)) {
        const int idx_h = (int)__fdiv_rn(hf - b.y1, b.roi_h);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float wf = (float)(w0 + p);
          if ((wf >= b.x1) & (wf < b.x2)) {
            const int idx_w = (int)__fdiv_rn(wf - b.x1, b.roi_w);
            int cell = idx_h * 2 + idx_w;
            cell = min(max(cell, 0), 3);
            const float4* c4 = reinterpret_cast<const float4*>(s_cof + j * 128 + cell * 32);
            float acc = 0.f;
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              const float4 c = c4[v];
              acc = fmaf(P[p][v * 4], c.x, acc);
              acc = fmaf(P[p][v * 4 + 1], c.y, acc);
              acc = fmaf(P[p][v * 4 + 2], c.z, acc);
              acc = fmaf(P[p][v * 4 + 3], c.w, acc);
            }
            o[p] = sigmoidf_(acc);
          }
        }
      }
      store4<OT>(out + ((size_t)(n0 + j) * H + h) * W + w0, o, vec, nvalid);
    }
  }
}

// x2 bilinear (align_corners=False) + threshold, top-left paste into [N,out_h,out_w] uint8.
// sipmask_head.py:630-633,648-654.  One thread = 4 consecutive output pixels of one row.
template <typename PT>
__global__ void __launch_bounds__(256) upsample2_thresh_kernel(const PT* __restrict__ pos, uint8_t* __restrict__ out,
                                                               int N, int H, int W, int out_h, int out_w, float thr) {
  const int xq = blockIdx.x * blockDim.x + threadIdx.x;   // quad index along x
  const int y = blockIdx.y;
  const int n = blockIdx.z;
  const int x0 = xq * 4;
  if (x0 >= out_w) return;
  uint8_t r[4] = {0, 0, 0, 0};
  if (y < 2 * H) {
    float sy = ((float)y + 0.5f) * 0.5f - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    const int y0 = (int)sy;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0);
    const float ly = sy - (float)y0, hy = 1.f - ly;
    const PT* r0 = pos + ((size_t)n * H + y0) * W;
    const PT* r1 = pos + ((size_t)n * H + y1) * W;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = x0 + i;
      if (x < out_w && x < 2 * W) {
        float sx = ((float)x + 0.5f) * 0.5f - 0.5f;
        sx = sx < 0.f ? 0.f : sx;
        const int xa = (int)sx;
        const int xb = xa + (xa < W - 1 ? 1 : 0);
        const float lx = sx - (float)xa, hx = 1.f - lx;
        const float v = hy * (hx * to_f<PT>(r0[xa]) + lx * to_f<PT>(r0[xb])) +
                        ly * (hx * to_f<PT>(r1[xa]) + lx * to_f<PT>(r1[xb]));
        r[i] = v > thr ? 1 : 0;
      }
    }
  }
  uint8_t* o = out + ((size_t)n * out_h + y) * out_w + x0;
  if (x0 + 3 < out_w && ((out_w & 3) == 0)) {
    *reinterpret_cast<uchar4*>(o) = make_uchar4(r[0], r[1], r[2], r[3]);
  } else {
    for (int i = 0; i < 4 && x0 + i < out_w; ++i) o[i] = r[i];
  }
}

// Same as above but bit-packed: out_bits [N, out_h, words] uint32, bit (x & 31) of word (x >> 5), LSB first.
// One thread = one 32-pixel word (13.4 MB instead of 107 MB per 100 masks at 800x1344 -> one small D2H).
template <typename PT>
__global__ void __launch_bounds__(128) upsample2_thresh_pack_kernel(const PT* __restrict__ pos, uint32_t* __restrict__ out,
                                                                    int N, int H, int W, int out_h, int out_w, int words,
                                                                    float thr) {
  const int wq = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int n = blockIdx.z;
  if (wq >= words) return;
  uint32_t bits = 0u;
  if (y < 2 * H) {
    float sy = ((float)y + 0.5f) * 0.5f - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    const int y0 = (int)sy;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0);
    const float ly = sy - (float)y0, hy = 1.f - ly;
    const PT* r0 = pos + ((size_t)n * H + y0) * W;
    const PT* r1 = pos + ((size_t)n * H + y1) * W;
    // source columns needed by output x in [32*wq, 32*wq+31]: 16*wq-1 .. 16*wq+16
    const int c0 = 16 * wq - 1;
    float top[18], bot[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      const int c = min(max(c0 + i, 0), W - 1);
      top[i] = to_f<PT>(r0[c]);
      bot[i] = to_f<PT>(r1[c]);
    }
    // x = 2k   -> src = k - 0.25: columns (k-1, k), weights (0.25, 0.75)
    // x = 2k+1 -> src = k + 0.25: columns (k, k+1), weights (0.75, 0.25)
    // (clamped loads make the x == 0 and right-edge cases equal to PyTorch's index clamping)
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int xe = 32 * wq + 2 * m, xo = xe + 1;
      const float ve = hy * (0.25f * top[m] + 0.75f * top[m + 1]) + ly * (0.25f * bot[m] + 0.75f * bot[m + 1]);
      const float vo = hy * (0.75f * top[m + 1] + 0.25f * top[m + 2]) + ly * (0.75f * bot[m + 1] + 0.25f * bot[m + 2]);
      if (xe < out_w && xe < 2 * W && ve > thr) bits |= 1u << (2 * m);
      if (xo < out_w && xo < 2 * W && vo > thr) bits |= 1u << (2 * m + 1);
    }
  }
  out[((size_t)n * out_h + y) * words + wq] = bits;
}


// ---------------------------------------------------------------------------------------------------------------
// Fully fused mask path: prototypes -> (selected sub-region dot product -> sigmoid -> crop) -> x2 bilinear upsample
// -> threshold -> bit-pack, without ever materialising pos_masks [N,H,W] (sipmask_head.py:609-633,648-654).
// HBM traffic = prototypes once (+halo) + N*out_h*words*4 bytes of bits (13.4 MB for 100 masks at 800x1333)
// instead of 107 MB (fp32 pos) written and read again.
//
// CTA tile: 8 x 64 prototype pixels (+1 halo) in shared memory -> 16 output rows x 4 output words per detection.
// Thread (slot = tid & 63, lane4 = tid >> 6): output row 16*ty + slot/4, word 4*tx + slot%4, detections lane4, lane4+4, ...
// Words whose 2 x 18 source pixels all lie outside the detection's box are written as 0 without any arithmetic.
template <typename PT> struct PixPitch;
template <> struct PixPitch<__half> { static constexpr int kElems = 40; };   // 64 B + 16 B pad: conflict-free LDS.128
template <> struct PixPitch<float> { static constexpr int kElems = 36; };    // 128 B + 16 B pad

constexpr int MF_TH = 8, MF_TW = 64, MF_NB = 32, MF_THREADS = 256;

template <typename PT>
__device__ __forceinline__ float dot32(const PT* px, const float* cof);
template <>
__device__ __forceinline__ float dot32<__half>(const __half* px, const float* cof) {
  float acc = 0.f;
  const uint4* q = reinterpret_cast<const uint4*>(px);
  const float4* c4 = reinterpret_cast<const float4*>(cof);
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const uint4 u = q[v];
    const __half2* hh = reinterpret_cast<const __half2*>(&u);
    const float4 ca = c4[2 * v], cb = c4[2 * v + 1];
    const float2 f0 = __half22float2(hh[0]), f1 = __half22float2(hh[1]), f2 = __half22float2(hh[2]), f3 = __half22float2(hh[3]);
    acc = fmaf(f0.x, ca.x, acc); acc = fmaf(f0.y, ca.y, acc); acc = fmaf(f1.x, ca.z, acc); acc = fmaf(f1.y, ca.w, acc);
    acc = fmaf(f2.x, cb.x, acc); acc = fmaf(f2.y, cb.y, acc); acc = fmaf(f3.x, cb.z, acc); acc = fmaf(f3.y, cb.w, acc);
  }
  return acc;
}
template <>
__device__ __forceinline__ float dot32<float>(const float* px, const float* cof) {
  float acc = 0.f;
  const float4* q = reinterpret_cast<const float4*>(px);
  const float4* c4 = reinterpret_cast<const float4*>(cof);
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    const float4 a = q[v], c = c4[v];
    acc = fmaf(a.x, c.x, acc); acc = fmaf(a.y, c.y, acc); acc = fmaf(a.z, c.z, acc); acc = fmaf(a.w, c.w, acc);
  }
  return acc;
}

template <typename PT, bool HWC>
__global__ void __launch_bounds__(MF_THREADS) mask_fused_pack_kernel(
    const PT* __restrict__ protos, const float* __restrict__ cofs, const float* __restrict__ boxes, float sx1, float sy1,
    float sx2, float sy2, uint32_t* __restrict__ out, int H, int W, int N, int out_h, int out_w, int words, float thr) {
  constexpr int PP = PixPitch<PT>::kElems;
  constexpr int SR = MF_TH + 2, SC = MF_TW + 2;
  extern __shared__ __align__(16) unsigned char mf_smem[];
  PT* s_p = reinterpret_cast<PT*>(mf_smem);                                   // [SR][SC][PP]
  float* s_cof = reinterpret_cast<float*>(mf_smem + (size_t)SR * SC * PP * sizeof(PT));   // [MF_NB][128]
  BoxP* s_box = reinterpret_cast<BoxP*>(s_cof + MF_NB * 128);                 // [MF_NB]

  const int tx = blockIdx.x, ty = blockIdx.y;
  const int r_base = ty * MF_TH - 1, c_base = tx * MF_TW - 1;                  // source coords of smem (0,0)
  // ---- prototype tile (+halo, edge-replicated) -> shared memory, read exactly once per CTA
  for (int i = threadIdx.x; i < SR * SC * 4; i += MF_THREADS) {
    const int part = i & 3, pixi = i >> 2;
    const int r = pixi / SC, c = pixi - r * SC;
    const int hs = min(max(r_base + r, 0), H - 1), ws = min(max(c_base + c, 0), W - 1);
    PT* dst = s_p + (size_t)pixi * PP;
    if (sizeof(PT) == 2) {
      if (HWC) {
        reinterpret_cast<uint4*>(dst)[part] = __ldg(reinterpret_cast<const uint4*>(protos + ((size_t)hs * W + ws) * 32) + part);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[part * 8 + k] = protos[((size_t)(part * 8 + k) * H + hs) * W + ws];
      }
    } else {
      if (HWC) {
        reinterpret_cast<uint4*>(dst)[2 * part] = __ldg(reinterpret_cast<const uint4*>(protos + ((size_t)hs * W + ws) * 32) + 2 * part);
        reinterpret_cast<uint4*>(dst)[2 * part + 1] = __ldg(reinterpret_cast<const uint4*>(protos + ((size_t)hs * W + ws) * 32) + 2 * part + 1);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[part * 8 + k] = protos[((size_t)(part * 8 + k) * H + hs) * W + ws];
      }
    }
  }

  const int slot = threadIdx.x & 63, lane4 = threadIdx.x >> 6;
  const int yo = ty * 2 * MF_TH + (slot >> 2);          // output row
  const int wq = tx * (MF_TW / 16) + (slot & 3);        // output word
  const bool out_ok = (yo < out_h) && (wq < words);
  // vertical taps: yo = 2k -> rows (k-1, k) weights (.25,.75); yo = 2k+1 -> rows (k, k+1) weights (.75,.25)
  const int k = yo >> 1;
  const int ra = (yo & 1) ? k : k - 1;                   // first source row (unclamped)
  const float wy_a = (yo & 1) ? 0.75f : 0.25f, wy_b = 1.f - wy_a;
  const int ra_c = min(max(ra, 0), H - 1), rb_c = min(max(ra + 1, 0), H - 1);     // clamped rows actually read
  const int ls_a = ra - r_base, ls_b = ra + 1 - r_base;                            // smem rows (already edge-replicated)
  const int col0 = 16 * wq - 1;                                                    // first source column (unclamped)
  const bool in_rows = (yo < 2 * H);

  for (int n0 = 0; n0 < N; n0 += MF_NB) {
    const int nb = min(MF_NB, N - n0);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * 128; i += MF_THREADS) s_cof[i] = __ldg(cofs + (size_t)n0 * 128 + i);
    if (threadIdx.x < nb) {
      const float* b = boxes + (size_t)(n0 + threadIdx.x) * 4;
      BoxP bp;
      bp.x1 = b[0] * sx1; bp.y1 = b[1] * sy1; bp.x2 = b[2] * sx2; bp.y2 = b[3] * sy2;
      bp.roi_w = (float)(((double)(bp.x2 - bp.x1) + 0.1) / 2);
      bp.roi_h = (float)(((double)(bp.y2 - bp.y1) + 0.1) / 2);
      bp.pad0 = bp.pad1 = 0.f;
      s_box[threadIdx.x] = bp;
    }
    __syncthreads();
    if (!out_ok) continue;
    for (int j = lane4; j < nb; j += 4) {
      const BoxP b = s_box[j];
      uint32_t bits = 0u;
      const float fa = (float)ra_c, fb = (float)rb_c;
      const bool row_a_in = (fa >= b.y1) & (fa < b.y2), row_b_in = (fb >= b.y1) & (fb < b.y2);
      const float cl = (float)max(col0, 0), cr = (float)min(col0 + 17, W - 1);
      if (in_rows && (row_a_in | row_b_in) && (cr >= b.x1) && (cl < b.x2)) {
        const int idxh_a = (int)__fdiv_rn(fa - b.y1, b.roi_h), idxh_b = (int)__fdiv_rn(fb - b.y1, b.roi_h);
        const float* cof = s_cof + j * 128;
        // vertical blend of the two source rows for the 18 source columns
        float vcol_prev = 0.f, vcol_cur = 0.f;
#pragma unroll 1
        for (int ci = 0; ci < 18; ++ci) {
          const int wsrc = min(max(col0 + ci, 0), W - 1);
          const float wf = (float)wsrc;
          float v = 0.f;
          if ((wf >= b.x1) & (wf < b.x2)) {
            const int idx_w = (int)__fdiv_rn(wf - b.x1, b.roi_w);
            const int lc = col0 + ci - c_base;                 // smem column
            if (row_a_in) {
              const int cell = min(max(idxh_a * 2 + idx_w, 0), 3);
              v = wy_a * sigmoidf_(dot32<PT>(s_p + ((size_t)ls_a * SC + lc) * PP, cof + cell * 32));
            }
            if (row_b_in) {
              const int cell = min(max(idxh_b * 2 + idx_w, 0), 3);
              v += wy_b * sigmoidf_(dot32<PT>(s_p + ((size_t)ls_b * SC + lc) * PP, cof + cell * 32));
            }
          }
          // horizontal taps: x = 2m -> (m-1, m) weights (.25,.75); x = 2m+1 -> (m, m+1) weights (.75,.25)
          // column index ci corresponds to source col 16*wq - 1 + ci; output x = 32*wq + 2*(ci-1) + {0,1} need (ci-1, ci) / (ci, ci+1)
          if (ci >= 1) {
            // even output x = 32*wq + 2*(ci-1): uses source cols (ci-1, ci)
            const int xe = 32 * wq + 2 * (ci - 1);
            if (ci <= 16) {
              const float ve = 0.25f * vcol_cur + 0.75f * v;
              if (xe < out_w && xe < 2 * W && ve > thr) bits |= 1u << (2 * (ci - 1));
            }
            // odd output x = 32*wq + 2*(ci-2) + 1: uses source cols (ci-1, ci) with weights (.75,.25)
            if (ci >= 2) {
              const int xo = 32 * wq + 2 * (ci - 2) + 1;
              const float vo = 0.75f * vcol_cur + 0.25f * v;
              if (xo < out_w && xo < 2 * W && vo > thr) bits |= 1u << (2 * (ci - 2) + 1);
            }
          }
          vcol_prev = vcol_cur;
          vcol_cur = v;
        }
        (void)vcol_prev;
      }
      out[((size_t)(n0 + j) * out_h + yo) * words + wq] = bits;
    }
  }
}

// CropSplit operator (ops/crop/src/crop_split_cuda_kernel.cu:19-59), c == 2.
template <typename T>
__global__ void crop_split_kernel(const T* __restrict__ data, const T* __restrict__ rois, T* __restrict__ out,
                                  long long count, int H, int W, int N) {
  for (long long index = blockIdx.x * (long long)blockDim.x + threadIdx.x; index < count;
       index += (long long)blockDim.x * gridDim.x) {
    const int n = (int)(index % N);
    const int pw = (int)((index / N) % W);
    const int ph = (int)(index / N / W);
    const float x1 = to_f<T>(rois[n * 4 + 0]), y1 = to_f<T>(rois[n * 4 + 1]);
    const float x2 = to_f<T>(rois[n * 4 + 2]), y2 = to_f<T>(rois[n * 4 + 3]);
    T v = T(0.f);
    if (((float)pw >= x1) & ((float)ph >= y1) & ((float)pw < x2) & ((float)ph < y2)) {
      const float roi_w = (float)(((double)(x2 - x1) + 0.1) / 2);
      const float roi_h = (float)(((double)(y2 - y1) + 0.1) / 2);
      const int idx_w = (int)__fdiv_rn((float)pw - x1, roi_w);
      const int idx_h = (int)__fdiv_rn((float)ph - y1, roi_h);
      const int cell = min(max(idx_h * 2 + idx_w, 0), 3);
      v = data[(long long)cell * count + index];
    }
    out[index] = v;
  }
}

}  // namespace smb

using namespace smb;

extern "C" int smb_mask_assemble(const void* protos, int protos_dtype, int layout_hwc, const float* cofs,
                                 const float* boxes, const float* host_box_scale4, void* out, int out_dtype,
                                 int H, int W, int N, smb_stream_t stream) {
  SMB_CHECK_ARG(protos && cofs && boxes && out && host_box_scale4, "smb_mask_assemble: null pointer");
  SMB_CHECK_ARG(H > 0 && W > 0 && N >= 0, "smb_mask_assemble: bad shape H=%d W=%d N=%d", H, W, N);
  SMB_CHECK_ARG((protos_dtype == SMB_F32 || protos_dtype == SMB_F16) && (out_dtype == SMB_F32 || out_dtype == SMB_F16),
                "smb_mask_assemble: bad dtype");
  if (N == 0) return SMB_OK;
  dim3 grid(cdiv(W, MA_TW), cdiv(H, MA_TH)), block(MA_THREADS);
  const float s0 = host_box_scale4[0], s1 = host_box_scale4[1], s2 = host_box_scale4[2], s3 = host_box_scale4[3];
  cudaStream_t st = (cudaStream_t)stream;
#define MA_LAUNCH(PT, HWC, OT)                                                                             \
  mask_assemble_kernel<PT, HWC, OT><<<grid, block, 0, st>>>((const PT*)protos, cofs, boxes, s0, s1, s2, s3, \
                                                            (OT*)out, H, W, N)
  const int key = (protos_dtype << 2) | ((layout_hwc ? 1 : 0) << 1) | out_dtype;
  switch (key) {
    case 0: MA_LAUNCH(float, false, float); break;
    case 1: MA_LAUNCH(float, false, __half); break;
    case 2: MA_LAUNCH(float, true, float); break;
    case 3: MA_LAUNCH(float, true, __half); break;
    case 4: MA_LAUNCH(__half, false, float); break;
    case 5: MA_LAUNCH(__half, false, __half); break;
    case 6: MA_LAUNCH(__half, true, float); break;
    case 7: MA_LAUNCH(__half, true, __half); break;
  }
#undef MA_LAUNCH
  SMB_LAUNCH_OK("mask_assemble_kernel");
  return SMB_OK;
}

extern "C" int smb_mask_upsample2_threshold(const void* pos, int pos_dtype, uint8_t* out_u8, int N, int H, int W,
                                            int out_h, int out_w, float thr, smb_stream_t stream) {
  SMB_CHECK_ARG(pos && out_u8, "smb_mask_upsample2_threshold: null pointer");
  SMB_CHECK_ARG(N >= 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0 && out_h <= 65535 && N <= 65535,
                "smb_mask_upsample2_threshold: bad shape");
  if (N == 0) return SMB_OK;
  dim3 block(256), grid(cdiv(cdiv(out_w, 4), 256), out_h, N);
  cudaStream_t st = (cudaStream_t)stream;
  if (pos_dtype == SMB_F32)
    upsample2_thresh_kernel<float><<<grid, block, 0, st>>>((const float*)pos, out_u8, N, H, W, out_h, out_w, thr);
  else
    upsample2_thresh_kernel<__half><<<grid, block, 0, st>>>((const __half*)pos, out_u8, N, H, W, out_h, out_w, thr);
  SMB_LAUNCH_OK("upsample2_thresh_kernel");
  return SMB_OK;
}

extern "C" int smb_mask_upsample2_threshold_pack(const void* pos, int pos_dtype, uint32_t* out_bits, int N, int H, int W,
                                                 int out_h, int out_w, float thr, smb_stream_t stream) {
  SMB_CHECK_ARG(pos && out_bits, "smb_mask_upsample2_threshold_pack: null pointer");
  SMB_CHECK_ARG(N >= 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0 && out_h <= 65535 && N <= 65535,
                "smb_mask_upsample2_threshold_pack: bad shape");
  if (N == 0) return SMB_OK;
  const int words = cdiv(out_w, 32);
  dim3 block(128), grid(cdiv(words, 128), out_h, N);
  cudaStream_t st = (cudaStream_t)stream;
  if (pos_dtype == SMB_F32)
    upsample2_thresh_pack_kernel<float><<<grid, block, 0, st>>>((const float*)pos, out_bits, N, H, W, out_h, out_w, words, thr);
  else
    upsample2_thresh_pack_kernel<__half><<<grid, block, 0, st>>>((const __half*)pos, out_bits, N, H, W, out_h, out_w, words, thr);
  SMB_LAUNCH_OK("upsample2_thresh_pack_kernel");
  return SMB_OK;
}

extern "C" int smb_mask_assemble_pack(const void* protos, int protos_dtype, int layout_hwc, const float* cofs,
                                      const float* boxes, const float* host_box_scale4, uint32_t* out_bits, int H, int W, int N,
                                      int out_h, int out_w, float thr, smb_stream_t stream) {
  SMB_CHECK_ARG(protos && cofs && boxes && out_bits && host_box_scale4, "smb_mask_assemble_pack: null pointer");
  SMB_CHECK_ARG(H > 0 && W > 0 && N >= 0 && out_h > 0 && out_w > 0, "smb_mask_assemble_pack: bad shape");
  SMB_CHECK_ARG(protos_dtype == SMB_F32 || protos_dtype == SMB_F16, "smb_mask_assemble_pack: bad dtype");
  if (N == 0) return SMB_OK;
  const int words = cdiv(out_w, 32);
  // tiles must cover every output row/word of the [out_h, words] frame (rows beyond 2H are written as zeros)
  const int rows_src = max(cdiv(out_h, 2), 1), cols_src = max(cdiv(words * 16, 1), 1);
  dim3 grid(cdiv(cols_src, MF_TW), cdiv(rows_src, MF_TH)), block(MF_THREADS);
  const float s0 = host_box_scale4[0], s1 = host_box_scale4[1], s2 = host_box_scale4[2], s3 = host_box_scale4[3];
  cudaStream_t st = (cudaStream_t)stream;
  const size_t esz = protos_dtype == SMB_F16 ? 2 : 4;
  const int pp = protos_dtype == SMB_F16 ? PixPitch<__half>::kElems : PixPitch<float>::kElems;
  const size_t smem = (size_t)(MF_TH + 2) * (MF_TW + 2) * pp * esz + MF_NB * 128 * sizeof(float) + MF_NB * sizeof(BoxP);
  static bool attr_done = false;
  if (!attr_done) {
    SMB_CUDA_OK(cudaFuncSetAttribute(mask_fused_pack_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    SMB_CUDA_OK(cudaFuncSetAttribute(mask_fused_pack_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    SMB_CUDA_OK(cudaFuncSetAttribute(mask_fused_pack_kernel<__half, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    SMB_CUDA_OK(cudaFuncSetAttribute(mask_fused_pack_kernel<__half, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    attr_done = true;
  }
#define MF_LAUNCH(PT, HWC)                                                                                          \
  mask_fused_pack_kernel<PT, HWC><<<grid, block, smem, st>>>((const PT*)protos, cofs, boxes, s0, s1, s2, s3, out_bits, H, W, \
                                                           N, out_h, out_w, words, thr)
  if (protos_dtype == SMB_F16) {
    if (layout_hwc) MF_LAUNCH(__half, true); else MF_LAUNCH(__half, false);
  } else {
    if (layout_hwc) MF_LAUNCH(float, true); else MF_LAUNCH(float, false);
  }
#undef MF_LAUNCH
  SMB_LAUNCH_OK("mask_fused_pack_kernel");
  return SMB_OK;
}

extern "C" int smb_crop_split_forward(const void* data, const void* rois, void* out, int dtype, int H, int W, int c,
                                      int N, smb_stream_t stream) {
  SMB_CHECK_ARG(data && rois && out, "smb_crop_split_forward: null pointer");
  SMB_CHECK_ARG(c == 2, "smb_crop_split_forward: only c == 2 is used by SipMask (got %d)", c);
  SMB_CHECK_ARG(H > 0 && W > 0 && N >= 0, "smb_crop_split_forward: bad shape");
  const long long count = (long long)H * W * N;
  if (count == 0) return SMB_OK;
  const int blocks = (int)min((long long)148 * 16, (count + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SMB_F32)
    crop_split_kernel<float><<<blocks, 256, 0, st>>>((const float*)data, (const float*)rois, (float*)out, count, H, W, N);
  else if (dtype == SMB_F16)
    crop_split_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)data, (const __half*)rois, (__half*)out, count, H, W, N);
  else
    SMB_CHECK_ARG(false, "smb_crop_split_forward: bad dtype %d", dtype);
  SMB_LAUNCH_OK("crop_split_kernel");
  return SMB_OK;
}
