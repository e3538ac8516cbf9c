"""Generate tests/golden/*.npz by running the UNMODIFIED reference python.

Run in the build container only (needs /root/reference):
    python tests/golden/gen_golden.py

The reference modules (`ResNet`, `FPN`, `SipMaskHead.forward`, `SipMaskHead.get_bboxes`,
`multiclass_nms_idx`, `fast_nms`, `distance2bbox`) are imported through
tests/golden/_ref_import.py.  Three native entry points have no CPU build in the
reference and are bound as follows (documented in DESIGN.md §oracle):
  * `deform_conv_cuda.deform_conv_forward_cuda`  -> oracle.ops.deform_conv
    (restatement of deform_conv_cuda_kernel.cu:85-115,191-243; cross-checked against
     torchvision.ops.deform_conv2d in tests/test_oracle.py)
  * `crop_split_cuda.crop_split_cuda_forward`    -> oracle.ops.crop_split
    (restatement of crop_split_cuda_kernel.cu:19-59)
  * `nms_cpu.nms`                                -> the reference's own nms_cpu.cpp compiled
    into oracle/_ref (oracle/build.py); falls back to oracle.ops.nms(cmp_ge=True) if absent.
`pycocotools.mask.encode` is replaced by an identity that keeps the binary mask.

Everything is small (fits in a few hundred KB) so that the fixtures can be committed.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import _ref_import  # noqa: E402

_ref_import.install()

from oracle import ops as O  # noqa: E402
from sipmask_b200 import synth  # noqa: E402


def bind_natives():
    import mmdet.core  # noqa: F401
    dc = sys.modules['mmdet.ops.dcn.deform_conv']
    cs = sys.modules['mmdet.ops.crop.crop_split']
    import mmdet.core  # noqa: F401  (pulls in mmdet.ops)
    nw = sys.modules['mmdet.ops.nms.nms_wrapper']

    def deform_conv_forward_cuda(input, weight, offset, output, col, ones, kW, kH, dW, dH, padW, padH,
                                 dilW, dilH, groups, deformable_groups, im2col_step):
        assert groups == 1 and dW == dH and padW == padH and dilW == dilH
        out = O.deform_conv(input, offset, weight, dW, padW, dilW, deformable_groups)
        output.copy_(out)
        return 1

    dc.deform_conv_cuda.deform_conv_forward_cuda = deform_conv_forward_cuda
    # DeformConvFunction.forward raises on CPU tensors (deform_conv.py:46-47): bypass the
    # `is_cuda` test only, keep the rest of the function.
    def forward(ctx, input, offset, weight, stride=1, padding=0, dilation=1, groups=1,
                deformable_groups=1, im2col_step=64):
        from torch.nn.modules.utils import _pair
        stride, padding, dilation = _pair(stride), _pair(padding), _pair(dilation)
        output = input.new_empty(dc.DeformConvFunction._output_size(input, weight, padding, dilation, stride))
        cur = min(im2col_step, input.shape[0])
        deform_conv_forward_cuda(input, weight, offset, output, None, None, weight.size(3), weight.size(2),
                                 stride[1], stride[0], padding[1], padding[0], dilation[1], dilation[0],
                                 groups, deformable_groups, cur)
        return output

    dc.DeformConvFunction.forward = staticmethod(forward)
    dc.deform_conv = dc.DeformConvFunction.apply

    def crop_split_cuda_forward(data, rois, out, height, width, c, n):
        out.copy_(O.crop_split(data, rois, c))

    cs.crop_split_cuda.crop_split_cuda_forward = crop_split_cuda_forward

    ref_nms = None
    try:
        sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
        import sipmask_ref_nms_cpu as ref_nms  # built by oracle/build.py from the reference source
    except Exception as e:  # pragma: no cover
        print('WARNING: oracle/_ref nms_cpu not built (%s); using oracle.ops.nms(cmp_ge=True)' % e)

    def nms_cpu_nms(dets, thr):
        if ref_nms is not None:
            return ref_nms.nms(dets, float(thr))
        return torch.from_numpy(O.nms(dets.numpy(), thr, cmp_ge=True))

    nw.nms_cpu.nms = nms_cpu_nms
    return ref_nms is not None


def gen_head_case(name, stacked_convs, gn, ssd_flag, sizes, img_shape, scale_factor, score_thr, seed, ori_shape=None,
                  rescoring_flag=False, compact=False):
    from mmdet.models.anchor_heads import sipmask_head as sh
    import mmcv  # the shim

    class Cfg(dict):
        __getattr__ = dict.get

    head = sh.SipMaskHead(num_classes=81, in_channels=256, stacked_convs=stacked_convs, ssd_flag=ssd_flag,
                          rescoring_flag=rescoring_flag, strides=[8, 16, 32, 64, 128],
                          norm_cfg=dict(type='GN', num_groups=32, requires_grad=True) if gn else None)
    sd = synth.head_state_dict(seed=seed, prefix='', stacked_convs=stacked_convs, gn=gn, cls_bias=-2.0,
                               rescoring_flag=rescoring_flag)
    missing = head.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing
    assert all(k.startswith(('loss', 'crop')) for k in missing.missing_keys), missing
    head.eval()
    g = torch.Generator().manual_seed(seed + 10)
    feats = [torch.randn(1, 256, h, w, generator=g) for (h, w) in sizes]
    # torch>=1.5 rejects an ndarray `scale_factor` (the SSD path passes one, sipmask_head.py:630):
    # convert it to a tuple of floats, nothing else changes.
    import types
    import torch.nn.functional as TF

    def _interp(x, size=None, scale_factor=None, **kw):
        if isinstance(scale_factor, np.ndarray):
            scale_factor = tuple(float(v) for v in scale_factor)
        return TF.interpolate(x, size=size, scale_factor=scale_factor, **kw)

    fproxy = types.SimpleNamespace(**{k: getattr(TF, k) for k in dir(TF) if not k.startswith('__')})
    fproxy.interpolate = _interp
    sh.F = fproxy
    # keep the binary masks instead of RLE
    sh.mask_util.encode = lambda arr: [np.array(arr[:, :, 0], order='C').copy()]
    cfg = Cfg(nms_pre=60, min_bbox_size=0, score_thr=score_thr, nms=Cfg(type='nms', iou_thr=0.5), max_per_img=20)
    ori_shape = img_shape if ori_shape is None else ori_shape
    meta = dict(img_shape=img_shape, ori_shape=ori_shape, scale_factor=scale_factor)
    with torch.no_grad():
        outs = head(feats)
        res = head.get_bboxes(*outs, [meta], cfg, rescale=True)[0]
    det_bboxes, det_labels, cls_segms = res
    mask_scores = None
    if rescoring_flag:                                  # (cls_segms, mask_scores) tuple, sipmask_head.py:659-660
        cls_segms, ms = cls_segms
        mask_scores = np.zeros(det_bboxes.shape[0], np.float32)
        cnt = [0] * 80
        for i in range(det_bboxes.shape[0]):
            l = int(det_labels[i])
            mask_scores[i] = np.atleast_1d(ms[l])[cnt[l]]
            cnt[l] += 1
    masks = []
    # cls_segms is per class in detection order; rebuild detection-ordered mask stack
    counters = [0] * 80
    for i in range(det_bboxes.shape[0]):
        l = int(det_labels[i])
        masks.append(cls_segms[l][counters[l]])
        counters[l] += 1
    out = dict(
        det_bboxes=det_bboxes.numpy(), det_labels=det_labels.numpy(),
        masks=np.stack(masks).astype(np.uint8) if masks else np.zeros((0,) + tuple(ori_shape[:2]), np.uint8),
        img_shape=np.array(img_shape), ori_shape=np.array(ori_shape), rescoring_flag=np.array(int(rescoring_flag)), scale_factor=np.atleast_1d(np.asarray(scale_factor, np.float32)),
        stacked_convs=np.array(stacked_convs), gn=np.array(int(gn)), ssd_flag=np.array(int(ssd_flag)),
        score_thr=np.array(score_thr, np.float32), seed=np.array(seed), sizes=np.array(sizes),
        nms_pre=np.array(60), max_per_img=np.array(20),
    )
    if mask_scores is not None:
        out['mask_scores'] = mask_scores
    if compact:
        # large case: the inputs are regenerated by the test from `seed` (same torch.Generator sequence) and the head outputs
        # through the oracle head (itself pinned to 1e-4 by the small fixtures); only the reference's RESULTS are stored
        out['masks'] = np.packbits(out['masks'], axis=-1)
        out['mask_w'] = np.array(ori_shape[1])
    else:
        out['feat_masks'] = outs[4].numpy()
    for i in range(len(sizes) if not compact else 0):
        out['feat%d' % i] = feats[i].numpy()
        out['cls%d' % i] = outs[0][i].numpy()
        out['bbox%d' % i] = outs[1][i].numpy()
        out['ctr%d' % i] = outs[2][i].numpy()
        out['cof%d' % i] = outs[3][i].numpy()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **{k: np.asarray(v) for k, v in out.items()})
    print(name, 'dets', det_bboxes.shape[0], 'labels', sorted(set(det_labels.tolist()))[:10])


def gen_backbone_case(name, seed=1, depth=50, stage_with_dcn=(False, False, False, False)):
    """ResNet(depth, caffe, eval BN) + FPN.  stage_with_dcn: the `++` configs' DeformConvPack blocks
    (configs/sipmask/sipmask++_r101_caffe_fpn_ssd_6x.py:13-14, resnet.py:146-168,288-291)."""
    from mmdet.models.backbones.resnet import ResNet
    from mmdet.models.necks.fpn import FPN
    dcn = dict(type='DCN', deformable_groups=1, fallback_on_stride=False) if any(stage_with_dcn) else None
    net = ResNet(depth=depth, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                 norm_cfg=dict(type='BN', requires_grad=False), style='caffe', dcn=dcn, stage_with_dcn=tuple(stage_with_dcn))
    sd = synth.backbone_state_dict(depth, seed, prefix='', stage_with_dcn=tuple(stage_with_dcn))
    net.load_state_dict(sd, strict=True)
    net.eval()
    fpn = FPN(in_channels=[256, 512, 1024, 2048], out_channels=256, start_level=1, add_extra_convs=True,
              extra_convs_on_inputs=False, num_outs=5, relu_before_extra_convs=True)
    fpn.load_state_dict(synth.neck_state_dict(seed + 1, prefix=''), strict=True)
    fpn.eval()
    img = synth.synthetic_image(64, 96, seed=0)
    with torch.no_grad():
        c = net(img)
        p = fpn(c)
    out = dict(img=img.numpy())
    # keep only checksums + small slices of the big tensors to stay small
    for i, t in enumerate(c):
        out['c%d_slice' % i] = t[0, :8].numpy()
        out['c%d_sum' % i] = np.array([t.double().sum().item(), t.double().abs().sum().item()])
    for i, t in enumerate(p):
        out['p%d' % i] = t[0, :16].numpy()
        out['p%d_sum' % i] = np.array([t.double().sum().item(), t.double().abs().sum().item()])
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, [tuple(t.shape) for t in p])


def _list_literals(src):
    """All top-level `[...]` list literals that follow an `np.array(` in `src`, evaluated (data only)."""
    import ast
    vals, i = [], 0
    while True:
        i = src.find('np.array(', i)
        if i < 0:
            return vals
        j = i + len('np.array(')
        while src[j] in ' \n':
            j += 1
        if src[j] != '[':
            i = j
            continue
        depth, k = 0, j
        while True:
            depth += {'[': 1, ']': -1}.get(src[k], 0)
            k += 1
            if depth == 0:
                break
        vals.append(ast.literal_eval(src[j:k]))
        i = k


def _bm_vectors():
    """SipMask-benchmark/tests/test_nms.py:16-58 (5 boxes x 5 thresholds) and :60-233 (53 boxes, thr 0.5): the known-answer
    vectors are read from the reference's test file as DATA (boxes are xyxy, passed to box_nms as they are)."""
    src = open('/root/reference/SipMask-benchmark/tests/test_nms.py').read()
    a = src.index('def test_nms_cpu'); b = src.index('def test_nms1_cpu'); c = src.index('def test_nms_cuda') if 'def test_nms_cuda' in src else len(src)
    import ast
    l5 = _list_literals(src[a:b])
    inp = np.asarray(l5[0], np.float32).reshape(-1, 5)
    seg = src[a:b]
    thrs = ast.literal_eval(seg[seg.index('test_thresh =') + 13:seg.index('gt_indices')].strip())
    gts = ast.literal_eval(seg[seg.index('gt_indices =') + 12:seg.index('for thresh')].strip())
    keeps = -np.ones((len(gts), 5), np.int64)
    for i, gt in enumerate(gts):
        keeps[i, :len(gt)] = gt
    l53 = _list_literals(src[b:c])
    seg = src[b:c]
    thr53 = float(seg[seg.index('box_nms(boxes, scores,') + 22:].split(')')[0])
    out = dict(bm5_boxes_xyxy=inp[:, :4], bm5_scores=inp[:, 4], bm5_thrs=np.asarray(thrs, np.float32), bm5_keeps=keeps,
               bm53_boxes_xyxy=np.asarray(l53[0], np.float32), bm53_scores=np.asarray(l53[1], np.float32),
               bm53_gt_indices=np.asarray(l53[2], np.int64), bm53_thr=np.array(thr53, np.float32))
    assert out['bm53_boxes_xyxy'].shape == (53, 4) and out['bm53_scores'].shape == (53,)
    return out


def gen_nms_vectors():
    """Known-answer NMS vectors copied as DATA from the reference's tests:
    MM/tests/test_nms.py:17-41, MM/mmdet/ops/nms/nms_wrapper.py:25-34, BM/tests/test_nms.py:16-58."""
    out = {}
    out['mm4_dets'] = np.array([[49.1, 32.4, 51.0, 35.9, 0.9], [49.3, 32.9, 51.0, 35.3, 0.9],
                                [35.3, 11.5, 39.9, 14.5, 0.4], [35.2, 11.7, 39.7, 15.7, 0.3]], np.float32)
    out['mm4_thr'] = np.array(0.7, np.float32)
    out['mm4_num_keep'] = np.array(3)
    out['mm7_dets'] = np.array([[49.1, 32.4, 51.0, 35.9, 0.9], [49.3, 32.9, 51.0, 35.3, 0.9],
                                [49.2, 31.8, 51.0, 35.4, 0.5], [35.1, 11.5, 39.1, 15.7, 0.5],
                                [35.6, 11.8, 39.3, 14.2, 0.5], [35.3, 11.5, 39.9, 14.5, 0.4],
                                [35.2, 11.7, 39.7, 15.7, 0.3]], np.float32)
    out['mm7_thr'] = np.array(0.7, np.float32)
    out['mm7_num_keep'] = np.array(3)
    out.update(_bm_vectors())
    np.savez_compressed(os.path.join(HERE, 'nms_known_answers.npz'), **out)


if __name__ == '__main__':
    torch.manual_seed(0)
    torch.set_num_threads(8)
    have_ref_nms = bind_natives()
    print('reference nms_cpu.cpp in use:', have_ref_nms)
    gen_nms_vectors()
    sizes = [(12, 16), (6, 8), (3, 4), (2, 2), (1, 1)]
    gen_head_case('ref_head_gn4', 4, True, False, sizes, (96, 125, 3), 1.0, 0.05, seed=3)
    gen_head_case('ref_head_ssd2', 2, False, True, sizes, (96, 128, 3),
                  np.array([1.0, 1.0, 1.0, 1.0], np.float32), 0.1, seed=5)
    gen_backbone_case('ref_backbone_r50_64x96')
    gen_backbone_case('ref_backbone_r101_64x96', seed=4, depth=101)
    gen_backbone_case('ref_backbone_r50_dcn_64x96', seed=6, stage_with_dcn=(False, True, True, True))
    # scale_factor != 1 and ori_shape != img_shape (rescale=True): boxes are divided by the scale factor and the masks are
    # interpolated by 2/scale_factor into ori-image space (sipmask_head.py:587-588,623,629-633,648-654)
    gen_head_case('ref_head_gn4_sf', 4, True, False, sizes, (96, 125, 3), 1.6667, 0.05, seed=3, ori_shape=(58, 75, 3))
    gen_head_case('ref_head_ssd2_sf', 2, False, True, sizes, (96, 128, 3),
                  np.array([1.7, 1.3, 1.7, 1.3], np.float32), 0.1, seed=5, ori_shape=(74, 75, 3))
    # SipMask++ mask rescoring (rescoring_flag=True, sipmask_head.py:200-219,635-643): six unpadded stride-2 convs need
    # >= 127 mask rows -> 256 x 256 image
    sizes_r = [(32, 32), (16, 16), (8, 8), (4, 4), (2, 2)]
    gen_head_case('ref_head_ssd2_rescore', 2, False, True, sizes_r, (256, 256, 3),
                  np.array([1.0, 1.0, 1.0, 1.0], np.float32), 0.1, seed=7, rescoring_flag=True, compact=True)
