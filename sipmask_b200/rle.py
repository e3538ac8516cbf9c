"""COCO run-length encoding of binary masks (column-major counts + the compressed string), numpy only.

Replaces `pycocotools.mask.encode(np.array(im_mask[:, :, np.newaxis], order='F'))[0]`
(MM/mmdet/models/anchor_heads/sipmask_head.py:655-656); pycocotools is not vendored in the reference and is not
available in the build environment.  Output dict layout matches pycocotools: {'size': [h, w], 'counts': bytes}.
"""
import numpy as np


def counts(mask):
    """mask [H,W] {0,1} -> run lengths in column-major order, starting with the number of zeros."""
    flat = np.asarray(mask, dtype=np.uint8).T.reshape(-1)
    if flat.size == 0:
        return np.zeros((0,), np.int64)
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    bounds = np.concatenate([[0], change, [flat.size]])
    runs = np.diff(bounds)
    if flat[0] != 0:
        runs = np.concatenate([[0], runs])
    return runs.astype(np.int64)


def to_string(cnts):
    """pycocotools rleToString: 5 data bits per character (+ continuation bit), deltas against counts[i-2] for i > 2."""
    out = bytearray()
    for i, c in enumerate(cnts):
        x = int(c)
        if i > 2:
            x -= int(cnts[i - 2])
        more = True
        while more:
            ch = x & 0x1f
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(ch + 48)
    return bytes(out)


def encode(mask):
    m = np.asarray(mask, dtype=np.uint8)
    return {'size': [int(m.shape[0]), int(m.shape[1])], 'counts': to_string(counts(m))}
