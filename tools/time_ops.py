"""Warm per-op timing of the whole launch sequence of the bench workload: every engine op is captured alone in a CUDA
graph (REPS back-to-back copies) and replayed, so the number contains no Python / driver launch cost."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sipmask_b200 import synth  # noqa: E402
from sipmask_b200.engine import SipMaskEngine  # noqa: E402

sd = synth.detector_state_dict(50, seed=1, cls_bias=bench.CLS_BIAS)
eng = SipMaskEngine(sd, (bench.H, bench.W), test_cfg=bench.TEST_CFG, img_shape=(bench.H, bench.IMG_W, 3), use_graph=False)
eng.forward(synth.synthetic_image(bench.H, bench.W, seed=0).cuda())
torch.cuda.synchronize()
REPS = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
names = getattr(eng, 'op_names', None) or ['op%d' % i for i in range(len(eng.ops))]
tot = 0.0
rows = []
for i, op in enumerate(eng.ops):
    if op is None:
        continue
    op()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            op()
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / REPS * 1e3
    tot += us
    rows.append((names[i], us))
agg = {}
for n, us in rows:
    k = n.split(':')[0]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
print('total of per-op warm times: %.1f us over %d ops' % (tot, len(rows)))
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-28s %4d %10.1f us %5.1f%%' % (k, c, us, 100 * us / tot))
if '-v' in sys.argv:
    for n, us in rows:
        print('%-40s %8.1f' % (n, us))
