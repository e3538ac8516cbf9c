"""ctypes binding of oracle/csrc/oracle_c.c (built by oracle/build.py).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib', 'liboracle_c.so')
        if not os.path.exists(path):
            path = _build.build_c()
        _lib = ctypes.CDLL(path)
        _lib.oracle_nms.restype = ctypes.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def nms(dets, thr, cmp_ge=0, plus_one=1):
    d = np.ascontiguousarray(dets, dtype=np.float32)
    keep = np.zeros(max(d.shape[0], 1), dtype=np.int64)
    k = lib().oracle_nms(_p(d, ctypes.c_float), ctypes.c_int64(d.shape[0]), ctypes.c_float(thr),
                         int(cmp_ge), int(plus_one), _p(keep, ctypes.c_int64))
    return keep[:k].copy()


def mask_assemble(protos, cofs, rois):
    protos = np.ascontiguousarray(protos, np.float32)
    cofs = np.ascontiguousarray(cofs, np.float32)
    rois = np.ascontiguousarray(rois, np.float32)
    _, H, W = protos.shape
    N = cofs.shape[0]
    out = np.zeros((N, H, W), np.float32)
    lib().oracle_mask_assemble(_p(protos, ctypes.c_float), _p(cofs, ctypes.c_float), _p(rois, ctypes.c_float),
                               H, W, N, _p(out, ctypes.c_float))
    return out


def upsample2_thresh(pos, thr):
    pos = np.ascontiguousarray(pos, np.float32)
    N, H, W = pos.shape
    out = np.zeros((N, 2 * H, 2 * W), np.uint8)
    lib().oracle_upsample2_thresh(_p(pos, ctypes.c_float), N, H, W, ctypes.c_float(thr), _p(out, ctypes.c_uint8))
    return out


def crop_split(data, rois):
    data = np.ascontiguousarray(data, np.float32)
    rois = np.ascontiguousarray(rois, np.float32)
    _, H, W, N = data.shape
    out = np.zeros((H, W, N), np.float32)
    lib().oracle_crop_split(_p(data, ctypes.c_float), _p(rois, ctypes.c_float), H, W, N, _p(out, ctypes.c_float))
    return out
