"""COCO RLE: the ABI's host string packer against the oracle's restatement of pycocotools' rleToString
(CPU test), and the device run-length kernel against the oracle's column-major counts (-m gpu)."""
import numpy as np
import pytest
import torch

from oracle import ops as O


def _random_masks(n, h, w, seed):
    rng = np.random.RandomState(seed)
    m = np.zeros((n, h, w), np.uint8)
    for i in range(n):
        kind = i % 5
        if kind == 0:
            continue                                              # empty mask -> single run [h*w]
        if kind == 1:
            m[i] = 1                                              # full mask -> [0, h*w]
            continue
        y0, x0 = rng.randint(0, h - 2), rng.randint(0, w - 2)
        y1, x1 = rng.randint(y0 + 1, h), rng.randint(x0 + 1, w)
        m[i, y0:y1, x0:x1] = 1
        if kind == 3:                                             # blobs with holes / ragged edges
            m[i] &= (rng.rand(h, w) > 0.3).astype(np.uint8)
        if kind == 4:                                             # touches the first pixel and the last column / row
            m[i, 0, 0] = 1
            m[i, h - 1, :] = 1
            m[i, :, w - 1] = 1
    return m


def test_rle_to_string_host_matches_oracle():
    from sipmask_b200 import ops
    rng = np.random.RandomState(0)
    cases = [np.array([5]), np.array([0, 7]), np.array([3, 1, 1000000, 2, 1, 70000]), np.array([0, 1, 1, 1, 1, 1, 1])]
    for _ in range(20):
        cases.append(rng.randint(0, 100000, size=rng.randint(1, 300)))
    for c in cases:
        assert ops.rle_to_string(c) == O.rle_to_string(c.astype(np.int64)), c[:8]


@pytest.mark.gpu
@pytest.mark.parametrize('h,w', [(800, 1333), (37, 70), (64, 32), (33, 3)])
def test_device_rle_counts_match_oracle(h, w):
    from sipmask_b200 import ops
    m = _random_masks(12, h, w, seed=h * 7 + w)
    words = (w + 31) // 32
    pad = np.zeros((m.shape[0], h + 3, words * 32), np.uint8)     # taller / wider storage than the cropped region
    pad[:, :h, :w] = m
    pad[:, h:, :] = 1                                             # garbage outside the crop must be ignored
    pad[:, :, w:] = 1
    bits = np.packbits(pad.reshape(m.shape[0], h + 3, words, 32), axis=-1, bitorder='little').view('<u4')[..., 0]
    bits_t = torch.from_numpy(bits.astype(np.int64).astype(np.uint32).view(np.int32)).cuda().contiguous()
    counts, n = ops.mask_rle_counts(bits_t, h, w, cap=max(2 * h * w // 3, 8))
    counts, n = counts.cpu().numpy().view(np.uint32), n.cpu().numpy()
    for i in range(m.shape[0]):
        want = O.rle_counts(m[i])
        assert n[i] == len(want), (i, n[i], len(want))
        assert counts[i, :n[i]].tolist() == list(want), i
    rles = ops.masks_to_rle(bits_t, h, w, m.shape[0], cap=16)     # tiny cap: exercises the overflow retry
    for i in range(m.shape[0]):
        assert rles[i]['size'] == [h, w]
        assert rles[i]['counts'] == O.rle_to_string(O.rle_counts(m[i])), i


@pytest.mark.gpu
def test_device_rle_respects_valid_count():
    from sipmask_b200 import ops
    m = _random_masks(6, 40, 50, seed=3)
    bits = np.packbits(np.pad(m, ((0, 0), (0, 0), (0, 14))).reshape(6, 40, 2, 32), axis=-1, bitorder='little').view('<u4')[..., 0]
    bits_t = torch.from_numpy(bits.view(np.int32).copy()).cuda()
    nv = torch.tensor([4], dtype=torch.int32, device='cuda')
    counts, n = ops.mask_rle_counts(bits_t, 40, 50, n_valid=nv)
    n = n.cpu().numpy()
    assert (n[4:] == 0).all() and (n[:4] > 0).all()
