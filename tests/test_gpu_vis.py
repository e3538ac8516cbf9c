"""SipMask-VIS on the GPU: drop-in VIS head + tracker against the reference python's clip fixture, and
the full VIS engine (384x640, BASELINE config 5 shape) against the oracle."""
import os

import numpy as np
import pytest
import torch

from sipmask_b200 import synth
from test_head_api import rle_decode
from test_vis_oracle import clip_feats, vis_meta

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu().flatten(), torch.as_tensor(b).double().flatten()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


class Cfg(dict):
    __getattr__ = dict.get


def _head(g):
    from sipmask_b200.head import SipMaskVISHead
    head = SipMaskVISHead(num_classes=41, in_channels=256, stacked_convs=3, strides=[8, 16, 32, 64, 128])
    sd = synth.head_state_dict(seed=int(g['seed']), prefix='', num_classes=41, stacked_convs=3, gn=True, cls_bias=-2.0, track=True)
    head.load_state_dict(sd, strict=True)
    return head.cuda().eval(), sd


def test_vis_dropin_head_forward_matches_reference(golden_dir):
    g = dict(np.load(os.path.join(golden_dir, 'ref_vis_clip.npz')))
    head, _ = _head(g)
    feats = tuple(f.cuda() for f in clip_feats(g)[0])
    cls, box, ctr, cof, fm, tf, tf_ref = head(feats, feats, False)
    torch.cuda.synchronize()
    for l in range(5):
        assert cls[l].shape == g['cls%d' % l].shape
        assert _rel(cls[l], g['cls%d' % l]) < 2e-2 and _rel(box[l], g['bbox%d' % l]) < 2e-2 and _rel(cof[l], g['cof%d' % l]) < 3e-2, l
    assert _rel(fm.float(), g['feat_masks']) < 2e-2
    assert tf.shape == g['track_feats'].shape and _rel(tf, g['track_feats']) < 2e-2
    assert tf_ref is tf
    with pytest.raises(NotImplementedError):
        head(feats, feats, True)


def test_vis_get_bboxes_and_tracker_match_reference_clip(golden_dir):
    """Head outputs from the fp32 oracle VIS head (pinned to the reference at 1e-4) -> the product's device post-processing,
    track-feature gather and host tracker must reproduce the reference's detections, object ids and per-object masks."""
    from oracle import model as M
    g = dict(np.load(os.path.join(golden_dir, 'ref_vis_clip.npz')))
    head, sd = _head(g)
    ohead = M.SipMaskVISHead(num_classes=41, stacked_convs=3)
    ohead.load_state_dict(sd, strict=True)
    ohead.eval()
    cfg = Cfg(nms_pre=int(g['nms_pre']), score_thr=float(g['score_thr']), nms=Cfg(type='nms', iou_thr=0.5),
              max_per_img=int(g['max_per_img']))
    for t, feats in enumerate(clip_feats(g)):
        with torch.no_grad():
            outs = ohead(feats, feats, False)
        dev = [[x.cuda() for x in lst] for lst in outs[:4]] + [outs[4].cuda(), outs[5].cuda(), outs[6].cuda()]
        det, lab, obj_segms, ids = head.get_bboxes(*dev, [vis_meta(g, t)], cfg, rescale=True)[0]
        assert lab.cpu().tolist() == g['f%d_lab' % t].tolist(), t
        np.testing.assert_allclose(det.cpu().numpy(), g['f%d_det' % t], rtol=1e-4, atol=1e-3)
        assert np.asarray(ids).tolist() == g['f%d_ids' % t].tolist(), t
        ref_masks = np.unpackbits(g['f%d_masks' % t], axis=-1)[:, :, :int(g['ori_shape'][1])]
        last = {int(o): i for i, o in enumerate(np.asarray(ids).tolist()) if o >= 0}
        assert sorted(obj_segms.keys()) == sorted(last.keys())
        for oid, i in last.items():
            m, r = rle_decode(obj_segms[oid]).astype(bool), ref_masks[i].astype(bool)
            assert (np.logical_and(m, r).sum() + 1e-9) / (np.logical_or(m, r).sum() + 1e-9) >= 0.999, (t, oid)


def test_vis_engine_full_path_vs_oracle():
    """Whole VIS frame on the engine (R50-FPN, 3-conv towers, 40 classes, track branch; 384 x 640 = padded 360 x 640, config 5):
    head outputs within 2e-2 of the fp32 oracle; detections, 512-d track features and masks equal to the oracle run on the
    engine's own head outputs."""
    from oracle import model as M
    from oracle import postproc as P
    from sipmask_b200 import ops
    from sipmask_b200.engine import SipMaskEngine
    H, W = 384, 640
    img_shape, ori_shape, sf = (360, 640, 3), (360, 640, 3), 1.0
    sd = {}
    sd.update(synth.backbone_state_dict(50, 1))
    sd.update(synth.neck_state_dict(2))
    sd.update(synth.head_state_dict(3, num_classes=41, stacked_convs=3, gn=True, cls_bias=-2.5, track=True))
    cfg = dict(nms_pre=200, score_thr=0.03, nms=dict(type='nms', iou_thr=0.5), max_per_img=10)
    img = synth.synthetic_image(H, W, seed=0)
    net = M.SipMaskDetector(50, stacked_convs=3)
    net.bbox_head = M.SipMaskVISHead(num_classes=41, stacked_convs=3)
    net.load_state_dict(sd, strict=True)
    net.eval()
    with torch.no_grad():
        ref = net.bbox_head(net.extract_feat(img))
    eng = SipMaskEngine(sd, (H, W), stacked_convs=3, num_classes=41, test_cfg=cfg, img_shape=img_shape, ori_shape=ori_shape,
                        scale_factor=sf, use_graph=True, vis=True)
    out = eng.forward(img.cuda())
    torch.cuda.synchronize()
    ho = eng.head_outputs()
    for l in range(5):
        assert _rel(ho['cls'][l], ref[0][l]) < 2e-2 and _rel(ho['bbox'][l], ref[1][l]) < 2e-2 and _rel(ho['cof'][l], ref[3][l]) < 2e-2, l
    assert _rel(ho['feat_masks'].float(), ref[4]) < 2e-2 and _rel(ho['track_feats'], ref[5]) < 2e-2
    res = P.vis_get_bboxes_single([t[0].cpu() for t in ho['cls']], [t[0].cpu() for t in ho['bbox']], [t[0].cpu() for t in ho['ctr']],
                                  [t[0].cpu() for t in ho['cof']], ho['feat_masks'][0].float().cpu(), eng.strides, img_shape,
                                  ori_shape, sf, cfg, rescale=True)
    k = int(out['count'][0])
    assert k == res['det_bboxes'].shape[0] and 0 < k <= 10
    assert out['det_labels'][0, :k].cpu().tolist() == res['det_labels'].tolist()
    np.testing.assert_allclose(out['det_bboxes'][0, :k].cpu().numpy(), res['det_bboxes'].numpy(), rtol=1e-5, atol=1e-5)
    want_f = P.extract_box_feature_center(ho['track_feats'][0].cpu(), res['det_bboxes'][:, :4])
    np.testing.assert_array_equal(out['track_feats'][0, :k].cpu().numpy(), want_f.numpy())
    assert k == out['track_feats'].shape[1] or out['track_feats'][0, k:].abs().max().item() == 0
    masks = ops.unpack_mask_bits(out['mask_bits'][0, :k].cpu(), ori_shape[1]).numpy().astype(bool)
    want = np.zeros((k,) + ori_shape[:2], bool)
    m = res['masks'].astype(bool)
    want[:, :min(m.shape[1], ori_shape[0]), :min(m.shape[2], ori_shape[1])] = m[:, :ori_shape[0], :ori_shape[1]]
    iou = (np.logical_and(masks, want).sum((1, 2)) + 1e-9) / (np.logical_or(masks, want).sum((1, 2)) + 1e-9)
    assert iou.min() >= 0.999, iou
