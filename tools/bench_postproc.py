"""Micro-timings (CUDA events) of the post-processing kernels on bench-shaped synthetic inputs."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sipmask_b200 import ops, synth  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


sizes = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
strides = (8, 16, 32, 64, 128)
cls, box, ctr, cof = synth.head_level_inputs(sizes, seed=3)
cl = [c.permute(1, 2, 0).contiguous().cuda() for c in cls]
bl = [b.permute(1, 2, 0).contiguous().cuda() for b in box]
tl = [t.permute(1, 2, 0).contiguous().cuda() for t in ctr]
print('decode_topk        %8.1f us' % timeit(lambda: ops.decode_topk(cl, bl, tl, strides, (800, 1333), 1000, scale_factor=1.0)))
boxes, scores, ct, loc = ops.decode_topk(cl, bl, tl, strides, (800, 1333), 1000, scale_factor=1.0)
bg = torch.cat([scores.new_zeros(scores.shape[0], 1), scores], 1)
for thr in (0.9999, 0.5, 0.2, 0.05, 0.01):
    n_pairs = int((scores > thr).sum())
    t = timeit(lambda: ops.multiclass_nms_idx(boxes, bg, thr, dict(iou_thr=0.5), 100, score_factors=ct, return_count_tensor=True))
    print('multiclass_nms thr=%-6g pairs=%7d  %8.1f us' % (thr, n_pairs, t))
print('fast_nms           %8.1f us' % timeit(lambda: ops.fast_nms(boxes, scores, ct, 0.5, 200, 0.1, 100, return_count_tensor=True)))
N, Hm, Wm = 100, 400, 672
protos = synth.prototypes(Hm, Wm).permute(1, 2, 0).contiguous().half().cuda()
g = torch.Generator().manual_seed(0)
cofs = torch.randn(N, 128, generator=g).cuda()
cx, cy = torch.rand(N, generator=g) * 1333, torch.rand(N, generator=g) * 800
bw, bh = torch.rand(N, generator=g) * 480 + 32, torch.rand(N, generator=g) * 480 + 32
bx = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1).clamp(min=0).cuda()
pos32 = torch.empty((N, Hm, Wm), dtype=torch.float32, device='cuda')
pos16 = torch.empty((N, Hm, Wm), dtype=torch.float16, device='cuda')
print('mask_assemble f32  %8.1f us' % timeit(lambda: ops.mask_assemble(protos, cofs, bx, 0.5, layout='hwc', out=pos32)))
print('mask_assemble f16  %8.1f us' % timeit(lambda: ops.mask_assemble(protos, cofs, bx, 0.5, layout='hwc', out=pos16)))
bits = torch.empty((N, 800, 42), dtype=torch.int32, device='cuda')
print('upsample2_pack f32 %8.1f us' % timeit(lambda: ops.mask_upsample2_threshold_pack(pos32, (800, 1333), 0.4, out=bits)))
if hasattr(ops, 'mask_assemble_pack'):
    print('mask fused pack    %8.1f us' % timeit(lambda: ops.mask_assemble_pack(protos, cofs, bx, 0.5, (800, 1333), 0.4, layout='hwc', out=bits)))
