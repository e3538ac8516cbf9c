"""SipMask-VIS: the oracle's VIS head, post-processing and tracker association against the
UNMODIFIED SipMask-VIS reference python run on a 5-frame synthetic clip (tests/golden/gen_golden_vis.py ->
ref_vis_clip.npz); and the product's host-side tracker against the same fixture."""
import os

import numpy as np
import pytest
import torch

from sipmask_b200 import synth


def clip_feats(g):
    gen = torch.Generator().manual_seed(int(g['seed']) + 10)
    base = [torch.randn(1, 256, int(h), int(w), generator=gen) for (h, w) in g['sizes']]
    frames = [base]
    for t in range(1, 5):
        frames.append([b + 0.15 * t * torch.randn(b.shape, generator=gen) for b in base])
    return frames


def vis_cfg(g):
    return dict(nms_pre=int(g['nms_pre']), score_thr=float(g['score_thr']), nms=dict(iou_thr=0.5), max_per_img=int(g['max_per_img']))


def vis_meta(g, t):
    return dict(img_shape=tuple(g['img_shape']), ori_shape=tuple(g['ori_shape']), scale_factor=float(g['scale_factor']),
                is_first=bool(g['f%d_is_first' % t]))


def test_vis_oracle_matches_reference_clip(golden_dir):
    from oracle import model as M
    from oracle import postproc as P
    g = dict(np.load(os.path.join(golden_dir, 'ref_vis_clip.npz')))
    head = M.SipMaskVISHead(num_classes=41, stacked_convs=3)
    sd = synth.head_state_dict(seed=int(g['seed']), prefix='', num_classes=41, stacked_convs=3, gn=True, cls_bias=-2.0, track=True)
    head.load_state_dict(sd, strict=True)
    head.eval()
    tracker = P.VISTracker()
    for t, feats in enumerate(clip_feats(g)):
        with torch.no_grad():
            outs = head(feats, feats, False)
        if t == 0:
            for l in range(5):
                np.testing.assert_allclose(outs[0][l].numpy(), g['cls%d' % l], rtol=1e-4, atol=1e-4)
                np.testing.assert_allclose(outs[1][l].numpy(), g['bbox%d' % l], rtol=1e-4, atol=1e-3)
                np.testing.assert_allclose(outs[3][l].numpy(), g['cof%d' % l], rtol=1e-4, atol=1e-4)
            np.testing.assert_allclose(outs[4].numpy(), g['feat_masks'], rtol=1e-4, atol=1e-4)
            np.testing.assert_allclose(outs[5].numpy(), g['track_feats'], rtol=1e-4, atol=1e-4)
        det, lab, masks, ids = P.vis_get_bboxes(outs, vis_meta(g, t), vis_cfg(g), tracker, rescale=True)
        assert lab.tolist() == g['f%d_lab' % t].tolist(), t
        np.testing.assert_allclose(det.numpy(), g['f%d_det' % t], rtol=1e-4, atol=1e-3)
        assert np.asarray(ids).tolist() == g['f%d_ids' % t].tolist(), t
        ref_masks = np.unpackbits(g['f%d_masks' % t], axis=-1)[:, :, :int(g['ori_shape'][1])]
        # the reference keeps ONE mask per object id (obj_segms dict, last detection wins): compare those
        last = {}
        for i, oid in enumerate(np.asarray(ids).tolist()):
            if oid >= 0:
                last[oid] = i
        for oid, i in last.items():
            m, r = masks[i].astype(bool), ref_masks[i].astype(bool)
            assert (np.logical_and(m, r).sum() + 1e-9) / (np.logical_or(m, r).sum() + 1e-9) >= 0.999, (t, i)


def test_host_tracker_matches_oracle_tracker():
    """sipmask_b200.tracker.Tracker (product, numpy on gathered records) == oracle.postproc.VISTracker on random sequences,
    incl. empty frames, is_first resets and many-to-one matches."""
    from oracle import postproc as P
    from sipmask_b200.tracker import Tracker
    rng = np.random.RandomState(0)
    a, b, c = P.VISTracker(), Tracker(native=False), Tracker(native=True)
    for t in range(12):
        n = int(rng.randint(0, 9))
        xy = rng.rand(n, 2) * 100
        wh = rng.rand(n, 2) * 60 + 4
        det = np.concatenate([xy, xy + wh, rng.rand(n, 1) * 0.9 + 0.05], 1).astype(np.float32)
        lab = rng.randint(0, 4, size=n).astype(np.int64)
        feats = rng.randn(n, 512).astype(np.float32) * 0.3
        first = t in (0, 7)
        ia = a.step(torch.from_numpy(det), torch.from_numpy(lab), torch.from_numpy(feats), first)
        ib = b.step(det, lab, feats, first)
        ic = c.step(det, lab, feats, first)                  # smb_track_step (host C) - the default of the product
        assert np.asarray(ia).tolist() == np.asarray(ib).tolist() == np.asarray(ic).tolist(), t
        assert c.num_objects == b.num_objects == a.prev_bboxes.shape[0]


def test_native_tracker_reproduces_reference_clip(golden_dir):
    """The product's default tracker step (smb_track_step, host C in the ABI library) on the reference clip: oracle head
    outputs -> oracle post-processing -> native association must give the reference python's object ids."""
    from oracle import model as M
    from oracle import postproc as P
    from sipmask_b200.tracker import Tracker
    g = dict(np.load(os.path.join(golden_dir, 'ref_vis_clip.npz')))
    head = M.SipMaskVISHead(num_classes=41, stacked_convs=3)
    head.load_state_dict(synth.head_state_dict(seed=int(g['seed']), prefix='', num_classes=41, stacked_convs=3, gn=True,
                                               cls_bias=-2.0, track=True), strict=True)
    head.eval()

    class Wrap(object):
        def __init__(self):
            self.t = Tracker(native=True)

        def step(self, det, lab, feats, first):
            return self.t.step(det.numpy(), lab.numpy(), feats.numpy(), first)
    w = Wrap()
    for t, feats in enumerate(clip_feats(g)):
        with torch.no_grad():
            outs = head(feats, feats, False)
        _, _, _, ids = P.vis_get_bboxes(outs, vis_meta(g, t), vis_cfg(g), w, rescale=True)
        assert np.asarray(ids).tolist() == g['f%d_ids' % t].tolist(), t
