"""Device-side `get_bboxes_single` (decode -> NMS -> coefficient gather -> mask assembly -> x2 upsample/threshold).

Mirrors MM/mmdet/models/anchor_heads/sipmask_head.py:543-662 with every step on the GPU and no host
synchronisation until the caller reads the fixed-shape result record:
    det_bboxes [max,5] f32, det_labels [max] i64, count i32, masks [max,H,W] u8.
"""
import numpy as np
import torch

from . import ops


def mask_up_factors(scale_factor, ssd_flag, scale=2):
    """`scale / scale_factor` (sipmask_head.py:632) or, on the SSD path, `scale / scale_factor[3:1:-1]` = (h, w) factors
    (:630) - the bilinear resize factor(s) of pos_masks, as python floats.  The division keeps the operand types of the
    reference (python float scale_factor -> double, numpy float32 array -> float32)."""
    if ssd_flag:
        sf = np.asarray(scale_factor)
        if sf.size == 4:
            return tuple(float(v) for v in (scale / sf[3:1:-1]))
        return float(scale / sf.reshape(-1)[0])
    if isinstance(scale_factor, np.ndarray):
        return float(scale / scale_factor.reshape(-1)[0])
    return float(scale / scale_factor)


def _cl(t):
    """CHW (reference layout) or HWC tensor -> channel-last contiguous fp32 [h,w,C]."""
    return t.float().permute(1, 2, 0).contiguous()


def get_bboxes_single(cls_scores, bbox_preds, centernesses, cof_preds, feat_mask, strides, img_shape, ori_shape,
                      scale_factor, cfg, rescale=False, ssd_flag=False, cmp_ge=False, mask_thr=0.4,
                      channel_last=False, feat_mask_layout='chw', box_scales=None, top_k=200, upsample=True, pack=False,
                      rescoring=None, legacy_interp=False, vis=False, track_feats=None):
    """Inputs per level: CHW tensors like the reference (channel_last=False) or [h,w,C] fp32 views.
    cfg: dict with nms_pre, score_thr, nms.iou_thr, max_per_img."""
    if not channel_last:
        cls_scores = [_cl(t) for t in cls_scores]
        bbox_preds = [_cl(t) for t in bbox_preds]
        centernesses = [_cl(t) for t in centernesses]
        cof_preds = [_cl(t) for t in cof_preds]
    nms_pre = cfg.get('nms_pre', -1)
    max_num = int(cfg.get('max_per_img', 100))
    sf = np.atleast_1d(np.asarray(scale_factor, dtype=np.float32))
    boxes, scores, ctr, loc = ops.decode_topk(cls_scores, bbox_preds, centernesses, strides, img_shape, nms_pre,
                                              scale_factor=(sf if rescale else None), box_scales=box_scales)
    iou_thr = cfg['nms']['iou_thr'] if isinstance(cfg['nms'], dict) else cfg['nms'].iou_thr
    if vis:
        # SipMask-VIS: fast_nms with cfg.score_thr / cfg.max_per_img (VIS/.../sipmask_head.py:733-734,951-993)
        det, lab, idx, cnt = ops.fast_nms(boxes, scores, ctr, iou_thr, top_k, cfg['score_thr'], max_num, return_count_tensor=True)
    elif not ssd_flag:
        det, lab, idx, cnt = ops.multiclass_nms_idx(boxes, scores, cfg['score_thr'], dict(iou_thr=iou_thr), max_num,
                                                    score_factors=ctr, has_bg_column=False, cmp_ge=cmp_ge,
                                                    return_count_tensor=True)
    else:
        max_num = 100                                    # hard-coded in fast_nms (sipmask_head.py:903)
        det, lab, idx, cnt = ops.fast_nms(boxes, scores, ctr, iou_thr, top_k, cfg['score_thr'], max_num,
                                          return_count_tensor=True)
    # coefficient gather for the kept rows only: loc -> level-concatenated location -> cof row
    cof_all = torch.cat([c.reshape(-1, c.shape[-1]) for c in cof_preds], 0) if len(cof_preds) > 1 else \
        cof_preds[0].reshape(-1, cof_preds[0].shape[-1])
    loc_kept = loc.long()[idx.clamp(min=0)]
    det_cofs = ops.gather_rows(cof_all, loc_kept, cnt, max_num)
    # rois = det * scale_factor / 2 (sipmask_head.py:621-623); scale_factor := 1 when rescale is None
    s4 = (sf if sf.size == 4 else np.repeat(sf, 4)).astype(np.float32)
    if rescale is None or (vis and not rescale):
        # MM: `scale_factor = scale_factor*0+1.0` rebinds it for rois AND resize when rescale is None (:621-622);
        # VIS: without rescale rois = det / 2 and the masks are resized by exactly 2 (VIS/...:752-762)
        s4 = np.ones(4, np.float32)
        scale_factor = s4 if sf.size == 4 else 1.0
    box_scale = s4 / 2.0
    up = mask_up_factors(scale_factor, ssd_flag)
    pos = ops.mask_assemble(feat_mask, det_cofs, det[:, :4].contiguous(), box_scale, layout=feat_mask_layout)
    out = dict(det_bboxes=det, det_labels=lab, idxs_keep=idx, count=cnt, pos_masks=pos, masks=None, mask_scores=None)
    if track_feats is not None:        # VIS: 512-d box-centre features of res_det = det * scale_factor (VIS/...:609-613,768-781)
        out['track_feats'] = ops.gather_track_feats(track_feats, det, cnt, (sf if rescale else np.ones(1, np.float32)))
    if rescoring is not None:      # SipMask++: dict(conv_w=[6], conv_b=[6], w1x1, b1x1) (sipmask_head.py:635-643)
        out['mask_scores'] = ops.mask_rescore(pos, rescoring['conv_w'], rescoring['conv_b'], rescoring['w1x1'],
                                              rescoring['b1x1'], lab, det, n_valid=cnt)
    if upsample:
        # masks = interpolate(pos_masks, scale_factor = 2 / scale_factor) > thr, pasted top-left into the ori_shape (rescale)
        # or img_shape canvas and truncated (sipmask_head.py:629-633,648-654)
        tgt = ori_shape if rescale else img_shape
        if pack:            # bit planes [max, H, ceil(W/32)] for the device RLE encoder (ops.masks_to_rle)
            out['mask_bits'] = ops.mask_resize_threshold_pack(pos, up, (int(tgt[0]), int(tgt[1])), mask_thr,
                                                              legacy_interp=legacy_interp)
            out['mask_hw'] = (int(tgt[0]), int(tgt[1]))
        else:
            out['masks'] = ops.mask_resize_threshold(pos, up, (int(tgt[0]), int(tgt[1])), mask_thr, legacy_interp=legacy_interp)
    return out
