// COCO run-length encoding of the bit-packed instance masks on the device.
//
// Replaces, per detection, `masks[i].cpu().numpy()` (a synchronous 4.3 MB copy) + `mask_util.encode(np.array(im_mask[:, :,
// np.newaxis], order='F'))` (single-threaded pycocotools on the host) - SipMask-mmdetection/mmdet/models/anchor_heads/
// sipmask_head.py:645-657.  The format is pycocotools' (not vendored in the reference): run lengths over the mask in
// COLUMN-major order, starting with the number of zeros; `smb_rle_to_string` (host, plain C) produces the compressed
// ASCII string (5 data bits + continuation bit per character, counts[i] for i > 2 stored as a delta against counts[i-2]).
//
// One CTA per detection.  The masks are row-major bit planes (bit x&31 of word x>>5), the runs are column-major: every warp
// transposes 32x32-bit tiles with 32 ballots, so that lane c holds 32 vertically adjacent pixels of column 32*wq + c; a run
// boundary is a bit that differs from the pixel before it in column-major order (the bottom of the previous column for
// row 0; 0 for the very first pixel).  Pass 1 counts boundaries per column, a block scan orders them, pass 2 writes the
// boundary positions, pass 3 differences them into run lengths.
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace smb {

constexpr int RLE_THREADS = 1024;

// 32 rows x 32 columns bit transpose: lane l supplies the word of row l, lane c receives the 32 row bits of column c
__device__ __forceinline__ uint32_t transpose32(uint32_t w, int lane) {
  uint32_t mine = 0u;
#pragma unroll
  for (int b = 0; b < 32; ++b) {
    const uint32_t t = __ballot_sync(0xffffffffu, (w >> b) & 1u);
    if (lane == b) mine = t;
  }
  return mine;
}

__device__ __forceinline__ uint32_t pixel_bit(const uint32_t* m, int words, int x, int y) {
  return (m[(size_t)y * words + (x >> 5)] >> (x & 31)) & 1u;
}

__global__ void __launch_bounds__(RLE_THREADS) rle_counts_kernel(const uint32_t* __restrict__ bits, int mask_h, int words, int H,
                                                                 int W, const int* __restrict__ n_valid, uint32_t* __restrict__ counts,
                                                                 int cap, int* __restrict__ n_counts) {
  extern __shared__ int s_cnt[];                 // [wcols*32 + 1] boundaries per column, then exclusive offsets
  __shared__ int s_warp[32];
  const int det = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (n_valid && det >= *n_valid) {
    if (threadIdx.x == 0) n_counts[det] = 0;
    return;
  }
  const uint32_t* m = bits + (size_t)det * mask_h * words;
  uint32_t* out = counts + (size_t)det * cap;
  const int wcols = (W + 31) >> 5;
  const int ncol = wcols * 32;
  const int ychunks = (H + 31) >> 5;

  // the pixel before (x, 0) in column-major order
  auto above_first = [&](int x) -> uint32_t { return x == 0 ? 0u : pixel_bit(m, words, x - 1, H - 1); };

  // ---- pass 1: run boundaries per column
  for (int wq = warp; wq < wcols; wq += RLE_THREADS / 32) {
    const int x = wq * 32 + lane;
    uint32_t prev = (x < W) ? above_first(x) : 0u;
    int c = 0;
    for (int yc = 0; yc < ychunks; ++yc) {
      const int y = yc * 32 + lane;
      const uint32_t w = (y < H) ? m[(size_t)y * words + wq] : 0u;
      const uint32_t col = transpose32(w, lane);
      const int rows = min(32, H - yc * 32);
      const uint32_t valid = rows == 32 ? 0xffffffffu : ((1u << rows) - 1u);
      const uint32_t t = (col ^ ((col << 1) | prev)) & valid;
      c += __popc(t);
      prev = (col >> (rows - 1)) & 1u;
    }
    s_cnt[x] = (x < W) ? c : 0;
  }
  __syncthreads();
  // ---- exclusive scan over columns (ncol <= 2048: two elements per thread)
  const int i0 = 2 * threadIdx.x, i1 = i0 + 1;
  const int a0 = i0 < ncol ? s_cnt[i0] : 0, a1 = i1 < ncol ? s_cnt[i1] : 0;
  int v = a0 + a1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  if (lane == 31) s_warp[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int wv = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, wv, o);
      if (lane >= o) wv += u;
    }
    s_warp[lane] = wv;
  }
  __syncthreads();
  const int base = (warp ? s_warp[warp - 1] : 0) + v - a0 - a1;
  const int total = s_warp[31];
  __syncthreads();
  if (i0 < ncol) s_cnt[i0] = base;
  if (i1 < ncol) s_cnt[i1] = base + a0;
  __syncthreads();
  // ---- pass 2: boundary positions (column-major pixel index), in order
  for (int wq = warp; wq < wcols; wq += RLE_THREADS / 32) {
    const int x = wq * 32 + lane;
    uint32_t prev = (x < W) ? above_first(x) : 0u;
    int o = s_cnt[x];
    for (int yc = 0; yc < ychunks; ++yc) {
      const int y = yc * 32 + lane;
      const uint32_t w = (y < H) ? m[(size_t)y * words + wq] : 0u;
      const uint32_t col = transpose32(w, lane);
      const int rows = min(32, H - yc * 32);
      const uint32_t valid = rows == 32 ? 0xffffffffu : ((1u << rows) - 1u);
      uint32_t t = (col ^ ((col << 1) | prev)) & valid;
      prev = (col >> (rows - 1)) & 1u;
      if (x < W) {
        while (t) {
          const int r = __ffs(t) - 1;
          t &= t - 1;
          if (o < cap) out[o] = (uint32_t)(x * H + yc * 32 + r);
          ++o;
        }
      }
    }
  }
  __syncthreads();
  // ---- pass 3: positions -> run lengths, in place
  // counts[0] = pos[0] (zeros before the first one; 0 if the mask starts with a one), counts[k] = pos[k] - pos[k-1],
  // counts[T] = H*W - pos[T-1];  T == 0 -> the single run [H*W].
  const int T = min(total, cap - 1);
  const int P = H * W;
  // stripes run back to front: a stripe reads out[k0-1 .. k0+1023] and writes out[k0 .. k0+1023], so the stripes below it
  // still see positions
  for (int k0 = (T / RLE_THREADS) * RLE_THREADS; k0 >= 0; k0 -= RLE_THREADS) {
    const int k = k0 + threadIdx.x;
    uint32_t val = 0u;
    if (k <= T) {
      const uint32_t hi = (k < T) ? out[k] : (uint32_t)P;
      const uint32_t lo = (k > 0) ? out[k - 1] : 0u;
      val = hi - lo;
    }
    __syncthreads();                               // every position of this stripe (and out[k0-1]) has been read
    if (k <= T) out[k] = val;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_counts[det] = (total >= cap) ? -(total + 1) : (T + 1);
}

}  // namespace smb

using namespace smb;

extern "C" int smb_mask_rle_counts(const uint32_t* mask_bits, int N, int mask_h, int words, int H, int W, const int* n_valid,
                                   uint32_t* counts, int cap, int* n_counts, smb_stream_t stream) {
  SMB_CHECK_ARG(mask_bits && counts && n_counts, "smb_mask_rle_counts: null pointer");
  SMB_CHECK_ARG(N >= 0 && H > 0 && W > 0 && H <= mask_h && W <= words * 32 && W <= 2048 && cap >= 2,
                "smb_mask_rle_counts: bad shape (N=%d H=%d W=%d mask_h=%d words=%d cap=%d)", N, H, W, mask_h, words, cap);
  SMB_CHECK_ARG((long long)H * W < (1ll << 31), "smb_mask_rle_counts: mask too large");
  if (N == 0) return SMB_OK;
  const int wcols = (W + 31) / 32;
  const size_t smem = (size_t)(wcols * 32 + 1) * sizeof(int);
  rle_counts_kernel<<<N, RLE_THREADS, smem, (cudaStream_t)stream>>>(mask_bits, mask_h, words, H, W, n_valid, counts, cap, n_counts);
  SMB_LAUNCH_OK("rle_counts_kernel");
  return SMB_OK;
}

// pycocotools rleToString (host): returns the string length (without terminator), or -1 if `cap` is too small
extern "C" int smb_rle_to_string(const uint32_t* counts, int n, char* out, int cap) {
  int p = 0;
  for (int i = 0; i < n; ++i) {
    long x = (long)counts[i];
    if (i > 2) x -= (long)counts[i - 2];
    int more = 1;
    while (more) {
      char c = (char)(x & 0x1f);
      x >>= 5;
      more = (c & 0x10) ? (x != -1) : (x != 0);
      if (more) c |= 0x20;
      c += 48;
      if (p >= cap) return -1;
      out[p++] = c;
    }
  }
  if (p < cap) out[p] = 0;
  return p;
}
