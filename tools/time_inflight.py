"""Throughput with several images in flight (serving.make_engines + EnginePool), for planner / grid-cap sweeps.

usage: python tools/time_inflight.py n [max_ctas [min_tiles [head_max_ctas]]]
"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sipmask_b200 import synth  # noqa: E402
from sipmask_b200.serving import make_engines, EnginePool  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
kw = {}
BATCH = int(os.environ.get('SMB_BATCH', '1'))          # images per forward (throughput experiment: batch-B forwards)
if len(sys.argv) > 2:
    kw['max_ctas'] = int(sys.argv[2]) or None
if len(sys.argv) > 3:
    kw['min_tiles'] = int(sys.argv[3])
if len(sys.argv) > 4:
    kw['head_max_ctas'] = int(sys.argv[4])
sd = synth.detector_state_dict(50, seed=1, cls_bias=bench.CLS_BIAS)
K = 40 * n
engs = make_engines(sd, (bench.H, bench.W), in_flight=n, test_cfg=bench.TEST_CFG, img_shape=(bench.H, bench.IMG_W, 3), use_graph=True,
                    batch=BATCH, **kw)
for i, e in enumerate(engs):
    e.img.copy_(torch.cat([synth.synthetic_image(bench.H, bench.W, seed=i * BATCH + b) for b in range(BATCH)], 0).cuda())
pool = EnginePool(engs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(2):
    e0.record()
    for k in range(K):
        pool.step()
    pool.flush()
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print('in_flight=%d batch=%d %s: %.1f img/s  (%.4f ms per image)' % (n, BATCH, kw, K * BATCH / ms * 1e3, ms / K / BATCH), flush=True)
