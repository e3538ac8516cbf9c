"""The one collective of the path: an all-gather of the fixed-shape detection record.

The reference gathers *pickled python results* with two NCCL all_gathers (sizes, then padded payload;
MM/mmdet/apis/test.py:117-147).  Here every rank contributes one [max_per_img, 7] fp32 record
(x1, y1, x2, y2, score, label, valid) per image, so a single `all_gather_into_tensor` suffices and the shape is static
(CUDA-graph friendly).  Images are sharded one per GPU; there is no other communication on the path.
"""
import torch
import torch.distributed as dist


def pack_record(det_bboxes, det_labels, count):
    """det_bboxes [max,5] f32, det_labels [max] i64, count int32 tensor [1] (device) -> [max,7] f32."""
    max_num = det_bboxes.shape[0]
    valid = (torch.arange(max_num, device=det_bboxes.device) < count.to(torch.long)).to(det_bboxes.dtype)
    return torch.cat([det_bboxes, det_labels.to(det_bboxes.dtype).unsqueeze(1), valid.unsqueeze(1)], 1)


def gather_records(record, out=None):
    """record [max,7] -> [world, max,7] on every rank (one collective)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return record.unsqueeze(0)
    if out is None:
        out = record.new_empty((world,) + tuple(record.shape))
    if dist.get_backend() == 'nccl':
        dist.all_gather_into_tensor(out, record.contiguous())
    else:                                             # gloo (CPU tests)
        parts = [torch.empty_like(record) for _ in range(world)]
        dist.all_gather(parts, record.contiguous())
        out.copy_(torch.stack(parts, 0))
    return out


def unpack_records(gathered):
    """[world,max,7] -> list over ranks of (det_bboxes [k,5], det_labels [k] int64) in rank (= image) order."""
    res = []
    for r in range(gathered.shape[0]):
        rec = gathered[r]
        k = int(rec[:, 6].sum().item())
        res.append((rec[:k, :5].clone(), rec[:k, 5].to(torch.long)))
    return res
