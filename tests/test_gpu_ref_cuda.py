"""Pins for the two CUDA-only reference operators, against the reference's OWN kernels.

`oracle/_ref/libsipmask_ref_cuda.so` is `CropSplitKernelForward` (MM/mmdet/ops/crop/src/crop_split_cuda_kernel.cu:19-88)
and `deformable_im2col_gpu_kernel` (MM/mmdet/ops/dcn/src/deform_conv_cuda_kernel.cu:84-277) compiled for sm_100a from
the reference sources where they lie (recipe: oracle/build.py::build_ref_cuda; wrappers oracle/ref_cuda/*.cu).
Each test compares three things on the same seeded inputs:
    reference kernel  ==  oracle restatement (oracle/ops.py)  ==  smb kernel through the C ABI
which removes the circularity of tests/golden/gen_golden.py (there the reference python's two native calls are bound to
the oracle, because neither has a CPU build).
"""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, 'oracle', '_ref', 'libsipmask_ref_cuda.so')


def _ref():
    if not os.path.exists(REF_SO):
        pytest.fail('%s missing: run oracle/build.py in the build container (needs /root/reference)' % REF_SO)
    lib = ctypes.CDLL(REF_SO)
    return lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def ref_crop_split(data, rois):
    """data [4,H,W,N] fp32 cuda, rois [N,4] fp32 cuda -> [H,W,N] (zero-initialised like ops/crop/crop_split.py:22)."""
    _, H, W, N = data.shape
    out = torch.zeros((H, W, N), dtype=torch.float32, device=data.device)
    torch.cuda.synchronize()
    rc = _ref().ref_crop_split_forward(_p(data), _p(rois), _p(out), H, W, 2, N)
    assert rc == 0, rc
    return out


def ref_deform_im2col(x, offset, dg, k=3, pad=1, stride=1, dil=1):
    """x [B,C,H,W], offset [B,dg*2*k*k,Ho,Wo] fp32 cuda -> col [C*k*k, B, Ho, Wo] fp32 (deform_conv_cuda.cpp:231-236)."""
    B, C, H, W = x.shape
    Ho = (H + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (k - 1) + 1)) // stride + 1
    col = torch.zeros((C * k * k, B, Ho, Wo), dtype=torch.float32, device=x.device)
    torch.cuda.synchronize()
    rc = _ref().ref_deformable_im2col(_p(x), _p(offset), C, H, W, k, pad, stride, dil, B, dg, _p(col))
    assert rc == 0, rc
    return col


def _rois(N, H, W, g):
    cx, cy = torch.rand(N, generator=g) * W, torch.rand(N, generator=g) * H
    bw, bh = torch.rand(N, generator=g) * W * 0.8 + 0.3, torch.rand(N, generator=g) * H * 0.8 + 0.3
    r = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
    fixed = torch.tensor([[0, 0, W, H], [-5, -7, W + 9, H + 3], [3.2, 4.1, 30.7, 35.2], [10, 10, 10.5, 10.5],
                          [0.5, 0.5, 1.5, 1.5], [20, 20, 19, 19], [W - 1, H - 1, W, H], [7, 3, 8, H - 1],
                          [4, 4, 12, 12], [4.0, 4.0, 11.9, 11.9]], dtype=torch.float32)
    r[:fixed.shape[0]] = fixed[:N]
    return r.contiguous()


@pytest.mark.parametrize('H,W,N', [(40, 56, 16), (100, 168, 37), (272, 272, 100)])
def test_crop_split_reference_kernel_vs_oracle_vs_smb(H, W, N):
    from oracle import ops as O
    from sipmask_b200 import ops
    g = torch.Generator().manual_seed(H * 7 + N)
    data = torch.rand(4, H, W, N, generator=g) + 0.01            # strictly positive: zeros mark "outside the box"
    rois = _rois(N, H, W, g)
    ref = ref_crop_split(data.cuda(), rois.cuda()).cpu().numpy()
    orc = O.crop_split(data, rois, 2).numpy()
    np.testing.assert_array_equal(orc, ref)                       # the restatement IS the reference kernel, bit for bit
    got = ops.crop_split(data.cuda(), rois.cuda()).cpu().numpy()
    np.testing.assert_array_equal(got, ref)                       # and so is the drop-in operator


@pytest.mark.parametrize('H,W,N,sf', [(100, 168, 40, 1.0), (136, 136, 100, 1.0), (75, 125, 33, 1.6666)])
def test_mask_assembly_vs_reference_pipeline(H, W, N, sf):
    """sipmask_head.py:615-627 with the reference's own CropSplit kernel:  4x (P @ cof_k^T) -> sigmoid -> stack ->
    CropSplitKernelForward -> permute, all fp32 on the GPU, versus the fused smb_mask_assemble."""
    from sipmask_b200 import ops, synth
    g = torch.Generator().manual_seed(N)
    protos = synth.prototypes(H, W, seed=N).cuda()                # [32,H,W]
    cofs = torch.randn(N, 128, generator=g).cuda()
    boxes = (_rois(N, H, W, g) * 2.0 / sf).cuda()                 # image space; rois = boxes * sf / 2
    P = protos.permute(1, 2, 0).contiguous()
    maps = torch.stack([torch.sigmoid(P @ cofs[:, 32 * k:32 * k + 32].t()) for k in range(4)], 0).contiguous()   # [4,H,W,N]
    rois = (boxes * (sf / 2.0)).contiguous()
    ref = ref_crop_split(maps, rois).permute(2, 0, 1).contiguous().cpu().numpy()
    got = ops.mask_assemble(protos, cofs, boxes, sf / 2.0, layout='chw').cpu().numpy()
    assert ((got == 0) == (ref == 0)).all()                       # crop + cell geometry identical to the reference kernel
    np.testing.assert_allclose(got, ref, atol=5e-6, rtol=0)       # 32-term fp32 dot + sigmoid: summation order only


@pytest.mark.parametrize('B,C,H,W,dg', [(1, 256, 13, 21, 4), (2, 256, 25, 42, 4), (1, 128, 17, 17, 1), (1, 64, 7, 11, 1)])
def test_deformable_im2col_reference_kernel_vs_oracle_vs_smb(B, C, H, W, dg):
    from oracle import ops as O
    from sipmask_b200 import conv
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(B, C, H, W, generator=g).half().float()       # fp16-representable so the smb (fp16 storage) path reads the same values
    off = torch.randn(B, dg * 18, H, W, generator=g) * 2.5
    off[:, :, 0, 0] = 0.0                                         # integer sampling points (bilinear weights exactly 0 / 1)
    off[:, :, -1, -1] = torch.round(off[:, :, -1, -1])
    off[:, 0::2, 1, :] = -3.7                                     # rows far outside the map (the `h_im > -1` predicate)
    ref = ref_deform_im2col(x.cuda(), off.cuda(), dg)             # [C*9, B, H, W]
    ref_bkhw = ref.permute(1, 0, 2, 3).reshape(B, C * 9, H * W).cpu()
    orc, _, _ = O.deform_im2col(x, off, 3, 3, 1, 1, 1, dg)        # [B, C*9, HW]
    # identical predicates / gather indices; the reference is compiled with FMA contraction, torch CPU is not
    np.testing.assert_allclose(orc.numpy(), ref_bkhw.numpy(), atol=2e-6, rtol=1e-6)
    assert ((orc == 0) == (ref_bkhw == 0)).float().mean().item() > 0.9999
    # smb kernel: NHWC fp16 in, [N,H,W,tap*C + c] fp16 out
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().half().cuda()
    off_nhwc = off.permute(0, 2, 3, 1).contiguous().cuda()
    col = conv.deform_im2col(x_nhwc, off_nhwc, dg).float().cpu()  # [B,H,W,9*C]
    want = ref_bkhw.view(B, C, 9, H * W).permute(0, 3, 2, 1).reshape(B, H, W, 9 * C)
    err = (col - want).abs().max().item()
    assert err <= 2e-3 * (want.abs().max().item() + 1.0), err      # one fp16 rounding of the output


def test_deform_conv_operator_vs_reference_im2col_gemm():
    """ops.DeformConv (drop-in for mmdet.ops.DeformConv, dcn/deform_conv.py:192-255) against the reference's
    im2col kernel followed by the weight GEMM (deform_conv_cuda.cpp:231-236), fp32 on the GPU."""
    from sipmask_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, C, H, W, dg, Cout = 1, 256, 25, 42, 4, 256
    x = torch.randn(B, C, H, W, generator=g).half().float().cuda()
    off = (torch.randn(B, dg * 18, H, W, generator=g) * 2.0).cuda()
    m = ops.DeformConv(C, Cout, 3, stride=1, padding=1, deformable_groups=dg).cuda()
    with torch.no_grad():
        m.weight.copy_((torch.randn(Cout, C, 3, 3, generator=g) / 48.0).half().float())
    col = ref_deform_im2col(x, off, dg)                            # [C*9, B, H, W]
    want = (m.weight.detach().view(Cout, -1).double() @ col.view(C * 9, -1).double()).view(Cout, B, H, W).permute(1, 0, 2, 3)
    got = m(x, off)
    assert got.shape == (B, Cout, H, W) and got.dtype == x.dtype
    err = (got.double() - want).abs().max().item()
    assert err <= 4e-3 * (want.abs().max().item() + 1.0), err      # fp16 column + fp16 output storage, fp32 accumulate


@pytest.mark.parametrize('H,W,N', [(40, 56, 16), (100, 168, 37)])
def test_crop_split_backward_and_gt_vs_reference_kernels(H, W, N):
    """Training-side companions: CropSplit backward and CropSplitGt forward / backward against the reference's
    own kernels (crop_split_cuda_kernel.cu:90-163, crop_split_gt_cuda_kernel.cu:19-140), bit for bit, incl. autograd."""
    from sipmask_b200 import ops
    lib = _ref()
    g = torch.Generator().manual_seed(H + N)
    rois = _rois(N, H, W, g).cuda()
    top = (torch.rand(H, W, N, generator=g) + 0.01).cuda()
    want = torch.zeros((4, H, W, N), dtype=torch.float32, device='cuda')              # crop_split.py:35 zero-initialises
    torch.cuda.synchronize()
    assert lib.ref_crop_split_backward(_p(top), _p(rois), _p(want), H, W, 2, N) == 0
    got = ops.crop_split_backward(top, rois)
    assert torch.equal(got, want)
    data = (torch.rand(4, H, W, N, generator=g) + 0.01).cuda().requires_grad_(True)
    out = ops.CropSplit(2)(data, rois)
    out.backward(top)
    assert torch.equal(data.grad, want)
    # CropSplitGt
    d = (torch.rand(H, W, N, generator=g) + 0.01).cuda()
    want_f = torch.zeros_like(d)
    torch.cuda.synchronize()
    assert lib.ref_crop_split_gt_forward(_p(d), _p(rois), _p(want_f), H, W, 2, N) == 0
    assert torch.equal(ops.CropSplitGt(2)(d, rois), want_f)
    want_b = torch.zeros_like(d)
    torch.cuda.synchronize()
    assert lib.ref_crop_split_gt_backward(_p(top), _p(rois), _p(want_b), H, W, 2, N) == 0
    dg = d.clone().requires_grad_(True)
    ops.CropSplitGt(2)(dg, rois).backward(top)
    assert torch.equal(dg.grad, want_b)
