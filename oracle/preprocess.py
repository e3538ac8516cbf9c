"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's test-time image pipeline.

Reference: `Resize(keep_ratio=True)` -> `Normalize(mean, std=1, to_rgb=False)` -> `Pad(size_divisor=32)` ->
`ImageToTensor`  (SipMask-mmdetection/mmdet/datasets/pipelines/transforms.py:97-110 `_resize_img`, :335-363 `Normalize`,
:274-300 `Pad`; configs/sipmask/sipmask_r50_caffe_fpn_gn_1x.py:60-61,77-90).

The arithmetic lives in a dependency that is NOT under /root/reference: **mmcv** (pinned in prose only: 0.4.3 for MM,
README.md:65), which forwards to OpenCV:
  * `mmcv.imrescale(img, scale, return_scale=True)`: `scale_factor = min(max_long / max(h, w), max_short / min(h, w))`,
    `new_size = (int(w * scale_factor + 0.5), int(h * scale_factor + 0.5))`, `cv2.resize(img, new_size, INTER_LINEAR)`;
  * `mmcv.imnormalize`: `(img.astype(float32) - mean) / std` (cv2.subtract / cv2.multiply, exact for std = 1);
  * `mmcv.impad_to_multiple`: zero padding on the bottom / right to the next multiple of `divisor`.
`resize_linear_u8` restates OpenCV's 8-bit INTER_LINEAR (modules/imgproc/src/resize.cpp: 11-bit fixed-point
coefficients `cvRound(w * 2048)`, horizontal pass in int32, vertical pass `((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2`;
x coefficients are clamped at the borders, y only clamps the row INDEX).  PIN: tests/test_preprocess.py compares it bit for
bit with `cv2.resize` (cv2 4.13 is in the image) - that is the function the reference pipeline calls.
"""
import numpy as np

MEAN_BGR = (102.9801, 115.9465, 122.7717)            # configs/sipmask/sipmask_r50_caffe_fpn_gn_1x.py:60-61


def rescale_size(h, w, scale=(1333, 800)):
    """mmcv.imrescale's size rule -> (new_h, new_w, scale_factor)."""
    max_long, max_short = max(scale), min(scale)
    sf = min(max_long / max(h, w), max_short / min(h, w))
    return int(h * float(sf) + 0.5), int(w * float(sf) + 0.5), sf


def _coef(dn, sn, clamp_coef):
    scale = 1.0 / (float(dn) / sn)                    # cv::resize: inv_scale = (double)dsize / ssize; scale = 1. / inv_scale
    d = np.arange(dn)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = f - s.astype(np.float32)
    if clamp_coef:                                    # x direction: the coefficient is forced to 0 outside
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= sn - 1
        f[hi] = 0
        s[hi] = sn - 1
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
    return np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1), a0, a1


def resize_linear_u8(src, dh, dw):
    """uint8 [h,w,c] -> uint8 [dh,dw,c], bit-exact cv2.resize(..., interpolation=cv2.INTER_LINEAR)."""
    sh, sw = src.shape[:2]
    if (sh, sw) == (dh, dw):
        return src.copy()
    sx, sx1, ax0, ax1 = _coef(dw, sw, True)
    sy, sy1, by0, by1 = _coef(dh, sh, False)
    S = src.astype(np.int64)
    rows = S[:, sx, :] * ax0[None, :, None] + S[:, sx1, :] * ax1[None, :, None]
    s0, s1 = rows[sy], rows[sy1]
    out = (((by0[:, None, None] * (s0 >> 4)) >> 16) + ((by1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def preprocess(img_u8, scale=(1333, 800), mean=MEAN_BGR, size_divisor=32):
    """uint8 BGR [h,w,3] -> (fp32 NCHW [1,3,H,W] normalized + zero padded, meta dict like the reference's img_meta)."""
    h, w = img_u8.shape[:2]
    nh, nw, sf = rescale_size(h, w, scale)
    r = resize_linear_u8(img_u8, nh, nw).astype(np.float32) - np.asarray(mean, np.float32)
    H = (nh + size_divisor - 1) // size_divisor * size_divisor
    W = (nw + size_divisor - 1) // size_divisor * size_divisor
    out = np.zeros((H, W, 3), np.float32)
    out[:nh, :nw] = r
    meta = dict(ori_shape=(h, w, 3), img_shape=(nh, nw, 3), pad_shape=(H, W, 3), scale_factor=sf, flip=False)
    return np.ascontiguousarray(out.transpose(2, 0, 1))[None], meta
