"""Launches the non-conv roofline kernels once each at the BASELINE config-A sizes (for `ncu --set full` captures):
mask_assemble (dense pos_masks), mask_assemble_pack (fused), deform_im2col_multi, gn_apply_multi, preprocess.
    ncu --set full --clock-control none --import-source on -k regex:'mask_|deform_im2col_multi|gn_apply_multi' \
        -o gpurun_out/r02_aux python tools/ncu_aux.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sipmask_b200 import conv, ops  # noqa: E402

dev = 'cuda'
H, W, IMG_W, N = 800, 1344, 1333, 100
Hm, Wm = H // 2, W // 2
g = torch.Generator().manual_seed(0)
protos = torch.relu(torch.randn(Hm, Wm, 32, generator=g)).half().to(dev)
cofs = torch.randn(N, 128, generator=g).to(dev)
cx, cy = torch.rand(N, generator=g) * IMG_W, torch.rand(N, generator=g) * H
bw, bh = torch.rand(N, generator=g) * 480 + 32, torch.rand(N, generator=g) * 480 + 32
boxes = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1).clamp(min=0).to(dev)
pos = torch.empty((N, Hm, Wm), dtype=torch.float32, device=dev)
bits = torch.empty((N, H, (IMG_W + 31) // 32), dtype=torch.int32, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
sizes = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)]
xs = [torch.randn(1, h, w, 256, generator=g).half().to(dev) for h, w in sizes]
offs = [(torch.randn(1, h, w, 72, generator=g) * 2).to(dev) for h, w in sizes]
cols = [torch.empty((1, h, w, 2304), dtype=torch.float16, device=dev) for h, w in sizes]
stats = [conv.groupnorm_stats(x) for x in xs]
gamma, beta = torch.ones(256, device=dev), torch.zeros(256, device=dev)
reps = int(os.environ.get('REPS', '2'))
for i in range(reps):
    flush.fill_(i)
    ops.mask_assemble(protos, cofs, boxes, 0.5, layout='hwc', out=pos)
    flush.fill_(i)
    ops.mask_assemble_pack(protos, cofs, boxes, 0.5, (H, IMG_W), 0.4, layout='hwc', out=bits)
    flush.fill_(i)
    conv.deform_im2col_multi(xs, offs, 4, cols)
    flush.fill_(i)
    conv.groupnorm_relu_apply_multi([x.clone() for x in xs], stats, gamma, beta)
torch.cuda.synchronize()
print('ok', int(bits.ne(0).sum()), float(pos.sum()))
