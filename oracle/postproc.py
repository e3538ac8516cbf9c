"""Oracle restatement of SipMaskHead.get_bboxes_single (decode, NMS, mask assembly).

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.
Follows MM/mmdet/models/anchor_heads/sipmask_head.py:543-662 line by line.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import ops as O


def decode_candidates(cls_scores, bbox_preds, centernesses, cof_preds, strides, img_shape,
                      nms_pre=1000, num_classes=80):
    """Per-level sigmoid / top-k / box decode (sipmask_head.py:556-592, before `rescale`).

    Inputs are per-image CHW tensors.  Returns boxes [n,4], scores [n,C] (no bg column),
    centerness [n], cofs [n,128], flat_index [n] (index into the level-concatenated
    h*w grid, for cross-checking).  Top-k order: descending score, ties -> lower index."""
    mlvl_bboxes, mlvl_scores, mlvl_ctr, mlvl_cofs, mlvl_idx = [], [], [], [], []
    base = 0
    for cls_score, bbox_pred, cof_pred, centerness, stride in zip(
            cls_scores, bbox_preds, cof_preds, centernesses, strides):
        h, w = cls_score.shape[-2:]
        points = O.get_points_single(h, w, stride)
        scores = cls_score.permute(1, 2, 0).reshape(-1, num_classes).sigmoid()
        ctr = centerness.permute(1, 2, 0).reshape(-1).sigmoid()
        bbox_pred = bbox_pred.permute(1, 2, 0).reshape(-1, 4)
        cof_pred = cof_pred.permute(1, 2, 0).reshape(-1, 128)
        idx = torch.arange(h * w)
        if nms_pre > 0 and scores.shape[0] > nms_pre:
            max_scores, _ = (scores * ctr[:, None]).max(dim=1)
            _, topk_inds = max_scores.sort(descending=True, stable=True)
            topk_inds = topk_inds[:nms_pre]
            points = points[topk_inds, :]
            bbox_pred = bbox_pred[topk_inds, :]
            cof_pred = cof_pred[topk_inds, :]
            scores = scores[topk_inds, :]
            ctr = ctr[topk_inds]
            idx = idx[topk_inds]
        mlvl_bboxes.append(O.distance2bbox(points, bbox_pred, max_shape=img_shape))
        mlvl_cofs.append(cof_pred)
        mlvl_scores.append(scores)
        mlvl_ctr.append(ctr)
        mlvl_idx.append(idx + base)
        base += h * w
    return (torch.cat(mlvl_bboxes), torch.cat(mlvl_scores), torch.cat(mlvl_ctr),
            torch.cat(mlvl_cofs), torch.cat(mlvl_idx))


def mask_up_factors(scale_factor, ssd_flag, scale=2):
    """`scale / scale_factor` (sipmask_head.py:632) or, on the SSD path, `scale / scale_factor[3:1:-1]` = (h, w) factors
    (:630), as python floats (what torch >= 1.5 needs; the arithmetic keeps the operand types of the reference: python float
    scale_factor -> double division, numpy float32 array -> float32 division)."""
    if ssd_flag:
        sf = np.asarray(scale_factor)
        if sf.size == 4:
            return tuple(float(v) for v in (scale / sf[3:1:-1]))
        return float(scale / sf.reshape(-1)[0])
    if isinstance(scale_factor, np.ndarray):
        return float(scale / scale_factor.reshape(-1)[0])
    return float(scale / scale_factor)


def assemble_masks(feat_mask, det_cofs, det_boxes, box_scale, up=2.0, thr=0.4, upsample=True, legacy_interp=False):
    """sipmask_head.py:609-633.  feat_mask [32,H,W], det_cofs [N,128], det_boxes [N,4].

    Returns (pos_masks [N,H,W] fp32 after CropSplit, masks [N,H*up,W*up] uint8 or None)."""
    img_mask1 = feat_mask.permute(1, 2, 0)
    m = [torch.sigmoid(img_mask1 @ det_cofs[:, 32 * k:32 * (k + 1)].t()) for k in range(4)]
    pos_masks = torch.stack(m, dim=0)                                   # [4,H,W,N]
    rois = det_boxes * box_scale
    pos_masks = O.crop_split(pos_masks, rois, 2).permute(2, 0, 1)       # [N,H,W]
    if not upsample:
        return pos_masks, None
    if isinstance(up, (tuple, list)):
        sf = tuple(float(u) for u in up)
    else:
        sf = float(up)
    # F.interpolate(scale_factor=...) as the reference calls it.  PyTorch >= 1.6 maps coordinates with the GIVEN factor
    # (src = (dst + 0.5) / factor - 0.5): that is what the reference python does when run here and what the *_sf golden
    # fixtures contain.  PyTorch <= 1.5 (the versions the reference README pins) recomputed the factor from the rounded
    # output size (in / out): legacy_interp=True.  Both agree when 2 / scale_factor is an integer (scale_factor = 1).
    masks = F.interpolate(pos_masks.unsqueeze(0), scale_factor=sf, mode='bilinear', align_corners=False,
                          recompute_scale_factor=True if legacy_interp else None).squeeze(0)
    return pos_masks, (masks > thr).to(torch.uint8)


def get_bboxes_single(cls_scores, bbox_preds, centernesses, cof_preds, feat_mask, strides,
                      img_shape, ori_shape, scale_factor, cfg, rescale=False, ssd_flag=False,
                      num_classes=80, cmp_ge=False, mask_thr=0.4, head=None, legacy_interp=False, timing=None):
    """Returns dict(det_bboxes [k,5], det_labels [k], idxs_keep [k], pos_masks [k,Hm,Wm],
    masks [k,Hi,Wi] uint8 pasted to ori/img shape, mask_scores or None).  `timing`: optional dict that receives the wall
    time of the stages (decode_nms / mask_assembly = 4 x matmul + sigmoid + CropSplit + resize + threshold / paste) for the
    CPU baseline of bench.py."""
    import time
    _t0 = time.perf_counter()
    boxes, scores, ctr, cofs, _ = decode_candidates(
        cls_scores, bbox_preds, centernesses, cof_preds, strides, img_shape,
        cfg.get('nms_pre', -1), num_classes)
    sf_arr = np.atleast_1d(np.asarray(scale_factor, dtype=np.float32))
    if rescale:
        boxes = boxes / torch.from_numpy(sf_arr)                        # :587-588
    scores_bg = torch.cat([scores.new_zeros(scores.shape[0], 1), scores], dim=1)
    if not ssd_flag:
        det_bboxes, det_labels, idxs = O.multiclass_nms_idx(
            boxes, scores_bg, cfg['score_thr'], cfg['nms']['iou_thr'], cfg['max_per_img'],
            score_factors=ctr, cmp_ge=cmp_ge)
        det_cofs = cofs[idxs]
    else:
        s = scores * ctr.view(-1, 1)
        det_bboxes, det_labels, det_cofs, idxs = O.fast_nms(
            boxes, s.transpose(1, 0).contiguous(), cofs,
            iou_threshold=cfg['nms']['iou_thr'], score_thr=cfg['score_thr'])
    out = dict(det_bboxes=det_bboxes, det_labels=det_labels, idxs_keep=idxs,
               pos_masks=None, masks=None, mask_scores=None)
    _t1 = time.perf_counter()
    if timing is not None:
        timing['decode_nms'] = _t1 - _t0
    if det_bboxes.shape[0] > 0:
        scale = 2
        if rescale is None:
            sf_arr = sf_arr * 0 + 1.0                                   # :621-622 rebinds scale_factor: rois AND resize use 1
            scale_factor = sf_arr if sf_arr.size == 4 else 1.0
        sf_t = torch.from_numpy(sf_arr)
        box_scale = sf_t / scale
        up = mask_up_factors(scale_factor, ssd_flag, scale)
        pos_masks, masks = assemble_masks(feat_mask, det_cofs, det_bboxes[:, :4], box_scale, up, mask_thr,
                                          legacy_interp=legacy_interp)
        out['pos_masks'] = pos_masks
        _t2 = time.perf_counter()
        tgt = ori_shape if rescale else img_shape
        k = masks.shape[0]
        im = np.zeros((k, tgt[0], tgt[1]), dtype=np.uint8)              # :648-654
        hh = min(masks.shape[1], tgt[0])
        ww = min(masks.shape[2], tgt[1])
        im[:, :hh, :ww] = masks.numpy()[:, :hh, :ww]
        out['masks'] = im
        if timing is not None:
            timing['mask_assembly'] = _t2 - _t1
            timing['paste'] = time.perf_counter() - _t2
        if head is not None and getattr(head, 'rescoring_flag', False):  # :635-643
            pred_iou = pos_masks.unsqueeze(1)
            pred_iou = head.convs_scoring(pred_iou)
            pred_iou = F.relu(head.mask_scoring(pred_iou))
            pred_iou = F.max_pool2d(pred_iou, kernel_size=pred_iou.size()[2:]).squeeze(-1).squeeze(-1)
            pred_iou = pred_iou[range(pred_iou.size(0)), det_labels]
            out['mask_scores'] = pred_iou * det_bboxes[:, -1]
    return out


# ------------------------------------------------------------------------------------------------------------------
# SipMask-VIS (paths relative to /root/reference/SipMask-VIS/mmdet/)
# ------------------------------------------------------------------------------------------------------------------
def vis_get_bboxes_single(cls_scores, bbox_preds, centernesses, cof_preds, feat_mask, strides, img_shape, ori_shape,
                          scale_factor, cfg, rescale=False, num_classes=40, mask_thr=0.5, legacy_interp=False):
    """models/anchor_heads/sipmask_head.py:686-766: decode as the image head, scores x centerness, fast_nms with
    cfg.score_thr / cfg.max_per_img (:951-993), CropSplit rois = det * scale_factor / 2 when rescale else det / 2
    (:748-755), masks = interpolate(2 / scale_factor | 2) > 0.5 (:757-764).  Returns dict(det_bboxes, det_labels, idxs_keep,
    masks [k,Hf,Wf] uint8 - the un-pasted interpolated masks, as the reference returns them)."""
    boxes, scores, ctr, cofs, _ = decode_candidates(cls_scores, bbox_preds, centernesses, cof_preds, strides, img_shape,
                                                    cfg.get('nms_pre', -1), num_classes)
    sf = float(np.atleast_1d(np.asarray(scale_factor, dtype=np.float64))[0])
    if rescale:
        boxes = boxes / boxes.new_tensor(np.float32(scale_factor))
    s = scores * ctr.view(-1, 1)
    det_bboxes, det_labels, det_cofs, idxs = O.fast_nms(boxes, s.transpose(1, 0).contiguous(), cofs, iou_threshold=0.5,
                                                        top_k=200, score_thr=cfg['score_thr'], max_num=cfg['max_per_img'])
    out = dict(det_bboxes=det_bboxes, det_labels=det_labels, idxs_keep=idxs, masks=None, pos_masks=None)
    if det_bboxes.shape[0] > 0:
        box_scale = torch.tensor([np.float32(scale_factor) / 2.0 if rescale else 0.5], dtype=torch.float32)
        up = float(2 / scale_factor) if rescale else 2.0
        pos, masks = assemble_masks(feat_mask, det_cofs, det_bboxes[:, :4], box_scale, up, mask_thr, legacy_interp=legacy_interp)
        out['pos_masks'], out['masks'] = pos, masks.numpy()
    return out


def extract_box_feature_center(track_feats, boxes, ref_feat_stride=8):
    """:768-781: the 512-d tracking feature at the box centre, floor((x1 + x2) / 2 / 8), of track_feats [512,h,w]."""
    cx = torch.floor((boxes[:, 2] + boxes[:, 0]) / 2.0 / ref_feat_stride).long()
    cy = torch.floor((boxes[:, 3] + boxes[:, 1]) / 2.0 / ref_feat_stride).long()
    return track_feats.permute(1, 2, 0)[cy, cx, :].clone()


def bbox_overlaps(b1, b2):
    """core/bbox/geometry.py:4-63, mode='iou', is_aligned=False: legacy +1 widths."""
    lt = torch.max(b1[:, None, :2], b2[:, :2])
    rb = torch.min(b1[:, None, 2:], b2[:, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    overlap = wh[:, :, 0] * wh[:, :, 1]
    a1 = (b1[:, 2] - b1[:, 0] + 1) * (b1[:, 3] - b1[:, 1] + 1)
    a2 = (b2[:, 2] - b2[:, 0] + 1) * (b2[:, 3] - b2[:, 1] + 1)
    return overlap / (a1[:, None] + a2 - overlap)


class VISTracker(object):
    """The association state machine of SipMaskHead.get_bboxes (:612-667) and compute_comp_scores (:544-562):
    comp = log_softmax([0 | feats . prev_feats^T]) + 1.0 * log(score) + 2.0 * [0 | IoU] + 10 * [1 | label equal];
    argmax 0 -> new object id; several detections on one previous object -> the larger comp score wins, the others keep
    id -1; the winner's feature / box replace the stored ones (labels of existing objects are NOT updated)."""

    def __init__(self, match_coeff=(1.0, 2.0, 10.0)):
        self.match_coeff = match_coeff
        self.prev_roi_feats = self.prev_bboxes = self.prev_det_labels = None

    def step(self, det_bboxes, det_labels, det_roi_feats, is_first):
        n = det_bboxes.shape[0]
        if n == 0:                                             # :605-608 returns early, state untouched
            return np.zeros((0,), np.int32)
        if is_first or self.prev_bboxes is None:
            self.prev_bboxes, self.prev_roi_feats, self.prev_det_labels = det_bboxes.clone(), det_roi_feats.clone(), det_labels.clone()
            return np.arange(n)
        prod = det_roi_feats @ self.prev_roi_feats.t()
        match_score = torch.cat([prod.new_zeros(n, 1), prod], dim=1)
        match_logprob = F.log_softmax(match_score, dim=1)
        label_delta = (self.prev_det_labels == det_labels.view(-1, 1)).float()
        ious = bbox_overlaps(det_bboxes[:, :4], self.prev_bboxes[:, :4])
        ious = torch.cat([ious.new_zeros(n, 1), ious], dim=1)
        label_delta = torch.cat([label_delta.new_ones(n, 1), label_delta], dim=1)
        comp = (match_logprob + self.match_coeff[0] * torch.log(det_bboxes[:, 4].view(-1, 1)) + self.match_coeff[1] * ious
                + self.match_coeff[2] * label_delta)
        _, match_ids = torch.max(comp, dim=1)
        match_ids = match_ids.numpy().astype(np.int32)
        det_obj_ids = np.ones(n, dtype=np.int32) * (-1)
        best = np.ones(self.prev_bboxes.shape[0]) * (-100)
        for idx, mid in enumerate(match_ids):
            if mid == 0:
                det_obj_ids[idx] = self.prev_roi_feats.shape[0]
                self.prev_roi_feats = torch.cat((self.prev_roi_feats, det_roi_feats[idx][None]), dim=0)
                self.prev_bboxes = torch.cat((self.prev_bboxes, det_bboxes[idx][None]), dim=0)
                self.prev_det_labels = torch.cat((self.prev_det_labels, det_labels[idx][None]), dim=0)
            else:
                obj = mid - 1
                sc = float(comp[idx, mid])
                if sc > best[obj]:
                    det_obj_ids[idx] = obj
                    best[obj] = sc
                    self.prev_roi_feats[obj] = det_roi_feats[idx]
                    self.prev_bboxes[obj] = det_bboxes[idx]
        return det_obj_ids


def vis_get_bboxes(head_outs, img_meta, cfg, tracker, rescale=False, strides=(8, 16, 32, 64, 128), legacy_interp=False):
    """One frame through SipMaskHead.get_bboxes (:565-682): detections + masks, box-centre track features taken at
    res_det_bboxes = det * scale_factor (:609-613), association.  Returns (det_bboxes, det_labels, masks pasted into
    ori_shape [k,h,w] uint8, det_obj_ids)."""
    cls, box, ctr, cof, fm, tf, _ = head_outs
    res = vis_get_bboxes_single([t[0] for t in cls], [t[0] for t in box], [t[0] for t in ctr], [t[0] for t in cof], fm[0], strides,
                                img_meta['img_shape'], img_meta['ori_shape'], img_meta['scale_factor'], cfg, rescale,
                                legacy_interp=legacy_interp)
    det, lab = res['det_bboxes'], res['det_labels']
    if det.shape[0] == 0:
        return det, lab, np.zeros((0,) + tuple(img_meta['ori_shape'][:2]), np.uint8), np.zeros((0,), np.int32)
    rdet = det.clone()
    if rescale:
        rdet[:, :4] *= np.float32(img_meta['scale_factor'])
    feats = extract_box_feature_center(tf[0], rdet[:, :4])
    ids = tracker.step(det, lab, feats, bool(img_meta['is_first']))
    oh, ow = img_meta['ori_shape'][:2]
    m = res['masks']
    im = np.zeros((m.shape[0], oh, ow), np.uint8)                       # :669-674
    hh, ww = min(m.shape[1], oh), min(m.shape[2], ow)
    im[:, :hh, :ww] = m[:, :hh, :ww]
    return det, lab, im, ids
