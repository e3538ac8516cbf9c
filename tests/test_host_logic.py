"""Host-side logic that needs no GPU: resize specification, mask resize factors, bench workloads and thread detection."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_resize_spec_follows_torch_interpolate_sizes():
    import torch
    import torch.nn.functional as F
    from sipmask_b200 import ops
    for H, W, up in [(48, 62, 2.0), (48, 62, 2.0 / 1.6667), (50, 64, (2 / 1.3, 2 / 1.7)), (60, 41, 2.0 / 3.1), (400, 672, 2.0)]:
        x = torch.zeros(1, 1, H, W)
        sf = tuple(float(u) for u in up) if isinstance(up, tuple) else float(up)
        want = F.interpolate(x, scale_factor=sf, mode='bilinear', align_corners=False).shape[2:]
        fh, fw, ry, rx = ops.resize_spec(H, W, up)
        assert (fh, fw) == tuple(want)
        uh, uw = (up if isinstance(up, tuple) else (up, up))
        assert ry == float(np.float32(1.0 / uh)) and rx == float(np.float32(1.0 / uw))
        assert ops.resize_spec(H, W, up, legacy_interp=True) == (fh, fw, 0.0, 0.0)


def test_mask_up_factors_match_oracle():
    from oracle import postproc as P
    from sipmask_b200 import postproc as Q
    for sf, ssd in [(1.0, False), (1.6667, False), (np.float32(1.25), False), (np.array([1.7, 1.3, 1.7, 1.3], np.float32), True),
                    (np.ones(4, np.float32), True)]:
        assert P.mask_up_factors(sf, ssd) == Q.mask_up_factors(sf, ssd)
    assert Q.mask_up_factors(np.array([1.7, 1.3, 1.7, 1.3], np.float32), True) == (float(np.float32(2) / np.float32(1.3)),
                                                                                   float(np.float32(2) / np.float32(1.7)))


def test_bench_workloads_and_threads():
    import bench
    assert set(bench.WORKLOADS) == {'A', 'A101', 'B', 'C'}
    for k, wl in bench.WORKLOADS.items():
        assert wl['H'] % 32 == 0 and wl['W'] % 32 == 0 and wl['img_w'] <= wl['W']
        cfg = bench.test_cfg(wl)
        assert cfg['max_per_img'] in (10, 100) and cfg['nms']['iou_thr'] == 0.5
    n, info = bench.host_threads()
    assert 1 <= n <= info['affinity'] and (info['cgroup_quota'] is None or n <= math.floor(info['cgroup_quota'] + 1e-6) or n == 1)
    sd = bench.state_dict_for(bench.WORKLOADS['C'])
    assert 'bbox_head.sipmask_track.weight' in sd and sd['bbox_head.fcos_cls.weight'].shape[0] == 40


def test_record_log_single_process():
    import torch
    from sipmask_b200 import dist as D
    log = D.RecordLog(3, 5, 'cpu', feat_dim=4)
    for i in range(4):                                   # ring: the 4th append overwrites row 0
        log.append(torch.full((5, 5), float(i)), torch.arange(5), torch.tensor([2], dtype=torch.int32), torch.full((5, 4), float(i)))
    g = log.gather()
    assert g.shape == (1, 3, 5, 11)
    assert g[0, 0, 0, 0] == 3 and g[0, 1, 0, 0] == 1 and g[0, :, :, 6].sum() == 6 and g[0, 0, 0, 7] == 3
