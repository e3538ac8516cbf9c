"""Generate tests/golden/ref_vis_clip.npz by running the UNMODIFIED SipMask-VIS reference python on a 4-frame synthetic clip.

Run in the build container only (needs /root/reference):   python tests/golden/gen_golden_vis.py
Same conventions as gen_golden.py (separate process: the VIS tree has its own `mmdet` package): the CUDA-only natives
are bound to the oracle (DCN, CropSplit - both pinned against the reference's own kernels on the GPU by
tests/test_gpu_ref_cuda.py), pycocotools' encode keeps the binary mask, `torch.cuda.current_device()` (used only as a
device argument for two tiny helper tensors, sipmask_head.py:549-553,628) is pointed at the CPU.
The head is SipMaskHead(num_classes=41, stacked_convs=3) as in VIS/configs/sipmask/sipmask_r50_caffe_fpn_gn_1x.py with
test_cfg nms_pre=200, score_thr=0.03, max_per_img=10; frames 1-3 are perturbations of frame 0 so that the tracker
re-identifies objects, frame 3 is flagged is_first again (new video).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import collections  # noqa: E402
import collections.abc  # noqa: E402

for _n in ('Sequence', 'Mapping', 'Iterable'):        # python < 3.10 spellings used by the 2019 tree (datasets/utils.py:1)
    if not hasattr(collections, _n):
        setattr(collections, _n, getattr(collections.abc, _n))
import _ref_import  # noqa: E402

_ref_import.install('/root/reference/SipMask-VIS')

from oracle import ops as O  # noqa: E402
from sipmask_b200 import synth  # noqa: E402

SEED = 11
SIZES = [(12, 20), (6, 10), (3, 5), (2, 3), (1, 2)]
IMG_SHAPE, ORI_SHAPE, SF = (96, 160, 3), (64, 107, 3), 1.5


def clip_feats(seed=SEED, n_frames=5):
    """Deterministic clip: frame t = base + 0.15 * t * noise_t (objects persist, features drift)."""
    g = torch.Generator().manual_seed(seed + 10)
    base = [torch.randn(1, 256, h, w, generator=g) for (h, w) in SIZES]
    frames = [base]
    for t in range(1, n_frames):
        frames.append([b + 0.15 * t * torch.randn(b.shape, generator=g) for b in base])
    return frames


def main():
    import mmdet.core  # noqa: F401
    from mmdet.models.anchor_heads import sipmask_head as sh
    dc = sys.modules['mmdet.ops.dcn.deform_conv']
    cs = sys.modules['mmdet.ops.crop.crop_split']

    def forward(ctx, input, offset, weight, stride=1, padding=0, dilation=1, groups=1, deformable_groups=1, im2col_step=64):
        from torch.nn.modules.utils import _pair
        stride, padding, dilation = _pair(stride), _pair(padding), _pair(dilation)
        assert groups == 1
        return O.deform_conv(input, offset, weight, stride[0], padding[0], dilation[0], deformable_groups)

    dc.DeformConvFunction.forward = staticmethod(forward)
    dc.deform_conv = dc.DeformConvFunction.apply
    sh.DeformConv.forward.__globals__['deform_conv'] = dc.deform_conv
    cs.crop_split_cuda.crop_split_cuda_forward = lambda data, rois, out, h, w, c, n: out.copy_(O.crop_split(data, rois, c))
    sh.mask_util.encode = lambda arr: [np.array(arr[:, :, 0], order='C').copy()]
    torch.cuda.current_device = lambda: 'cpu'

    class Cfg(dict):
        __getattr__ = dict.get

    head = sh.SipMaskHead(num_classes=41, in_channels=256, stacked_convs=3, strides=[8, 16, 32, 64, 128],
                          norm_cfg=dict(type='GN', num_groups=32, requires_grad=True))
    sd = synth.head_state_dict(seed=SEED, prefix='', num_classes=41, stacked_convs=3, gn=True, cls_bias=-2.0, track=True)
    r = head.load_state_dict(sd, strict=False)
    assert not r.unexpected_keys, r
    assert all(k.startswith(('loss', 'crop')) for k in r.missing_keys), r
    head.eval()
    cfg = Cfg(nms_pre=200, min_bbox_size=0, score_thr=0.03, nms=Cfg(type='nms', iou_thr=0.5), max_per_img=10)
    out = dict(seed=np.array(SEED), sizes=np.array(SIZES), img_shape=np.array(IMG_SHAPE), ori_shape=np.array(ORI_SHAPE),
               scale_factor=np.array(SF), nms_pre=np.array(200), score_thr=np.array(0.03, np.float32), max_per_img=np.array(10))
    frames = clip_feats()
    is_first = [True, False, False, False, True]
    for t, feats in enumerate(frames):
        meta = dict(img_shape=IMG_SHAPE, ori_shape=ORI_SHAPE, scale_factor=SF, is_first=is_first[t])
        with torch.no_grad():
            outs = head(feats, feats, False)
            det, lab, segms, ids = head.get_bboxes(*outs, [meta], cfg, rescale=True)[0]
        k = det.shape[0]
        masks = np.zeros((k,) + ORI_SHAPE[:2], np.uint8)
        for i in range(k):
            if int(ids[i]) >= 0:
                masks[i] = segms[int(ids[i])]
        out['f%d_det' % t] = det.numpy().copy()      # copy: the head keeps (and later mutates) this tensor as prev_bboxes (:617)
        out['f%d_lab' % t] = lab.numpy().copy()
        out['f%d_ids' % t] = np.asarray(ids, np.int64)
        out['f%d_masks' % t] = np.packbits(masks, axis=-1)
        out['f%d_is_first' % t] = np.array(int(is_first[t]))
        if t == 0:                               # head outputs of one frame: pins the VIS head restatement (track branch incl.)
            for l in range(5):
                out['cls%d' % l], out['bbox%d' % l] = outs[0][l].numpy(), outs[1][l].numpy()
                out['ctr%d' % l], out['cof%d' % l] = outs[2][l].numpy(), outs[3][l].numpy()
            out['feat_masks'], out['track_feats'] = outs[4].numpy(), outs[5].numpy()
        print('frame', t, 'dets', k, 'ids', list(map(int, ids)), 'labels', lab.tolist())
    np.savez_compressed(os.path.join(HERE, 'ref_vis_clip.npz'), **out)


if __name__ == '__main__':
    torch.manual_seed(0)
    torch.set_num_threads(8)
    main()
