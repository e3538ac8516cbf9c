"""SipMask++ mask rescoring against the oracle's fp32 restatement (sipmask_head.py:200-219,635-643).
Tolerance: fp32 direct convolutions with a different summation order than the oracle's library convs -> 1e-4 relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _oracle_head(seed=3):
    from oracle import model as M
    from sipmask_b200 import synth
    head = M.SipMaskHead(stacked_convs=2, gn=False, ssd_flag=True, rescoring_flag=True)
    sd = synth.head_state_dict(seed, stacked_convs=2, gn=False, rescoring_flag=True, cls_bias=-2.5, prefix='')
    head.load_state_dict(sd, strict=True)
    return head.eval(), sd


@pytest.mark.parametrize('hw', [(272, 272), (400, 672), (127, 131)])
def test_rescoring_chain_matches_oracle(hw):
    import torch.nn.functional as F
    from sipmask_b200 import ops
    head, _ = _oracle_head()
    g = torch.Generator().manual_seed(hw[0])
    N = 7
    pos = torch.rand(N, hw[0], hw[1], generator=g)
    pos[:, : hw[0] // 3] = 0                                               # cropped masks are mostly zeros
    labels = torch.randint(0, 80, (N,), generator=g)
    det = torch.rand(N, 5, generator=g)
    with torch.no_grad():
        x = head.convs_scoring(pos.unsqueeze(1))
        x = F.relu(head.mask_scoring(x))
        want = F.max_pool2d(x, kernel_size=x.shape[2:]).flatten(1)[range(N), labels] * det[:, 4]
    cw = [m.conv.weight.cuda() for m in head.convs_scoring]
    cb = [m.conv.bias.cuda() for m in head.convs_scoring]
    nv = torch.tensor([N - 2], dtype=torch.int32, device='cuda')
    got = ops.mask_rescore(pos.cuda(), cw, cb, head.mask_scoring.weight.cuda(), head.mask_scoring.bias.cuda(), labels.cuda(),
                           det.cuda(), n_valid=nv).cpu()
    assert (got[N - 2:] == 0).all()
    np.testing.assert_allclose(got[:N - 2].numpy(), want[:N - 2].numpy(), rtol=1e-4, atol=1e-6)
    assert want[:N - 2].abs().max() > 0


def test_head_api_returns_mask_scores_like_the_oracle():
    """Drop-in SipMaskHead(rescoring_flag=True, ssd_flag=True): same state_dict as the oracle head, `get_bboxes` returns
    (cls_segms, mask_scores) and the scores equal the oracle's on the head's own outputs."""
    from oracle import postproc as P
    from sipmask_b200.head import SipMaskHead
    head_o, sd = _oracle_head()
    head = SipMaskHead(num_classes=81, in_channels=256, stacked_convs=2, ssd_flag=True, rescoring_flag=True, norm_cfg=None,
                       strides=[8, 16, 32, 64, 128]).cuda()
    missing = head.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(0)
    sizes = [(32, 32), (16, 16), (8, 8), (4, 4), (2, 2)]       # 256 x 256 image: stride-2 masks 128 x 128 (>= 127 for six stride-2 convs)
    feats = [torch.randn(1, 256, h, w, generator=g) for h, w in sizes]
    outs = head(tuple(f.cuda() for f in feats))
    cfg = dict(nms_pre=200, score_thr=0.1, nms=dict(type='nms', iou_thr=0.5), max_per_img=100)
    sf = np.ones(4, dtype=np.float32)
    meta = dict(img_shape=(256, 256, 3), ori_shape=(256, 256, 3), scale_factor=sf)
    det, lab, (segms, mask_scores) = head.get_bboxes(*outs, [meta], cfg, rescale=True)[0]
    k = det.shape[0]
    assert k > 0 and sum(len(s) for s in segms) == k and sum(len(s) for s in mask_scores) == k
    res = P.get_bboxes_single([t[0].float().cpu() for t in outs[0]], [t[0].float().cpu() for t in outs[1]],
                              [t[0].float().cpu() for t in outs[2]], [t[0].float().cpu() for t in outs[3]],
                              outs[4][0].float().cpu(), head.strides, meta['img_shape'], meta['ori_shape'], sf, cfg, rescale=True,
                              ssd_flag=True, head=head_o)
    assert res['det_labels'].tolist() == lab.cpu().tolist()
    want = res['mask_scores'].detach().numpy()
    labels = lab.cpu().numpy()
    got = np.zeros(k, np.float32)
    for c in range(80):
        got[labels == c] = mask_scores[c]
    np.testing.assert_allclose(got, want, rtol=2e-3, atol=1e-5)
