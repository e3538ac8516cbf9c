"""Run selected convolution plans of the bench workload between cudaProfilerStart/Stop (for ncu --set full).
usage: python tools/profile_conv.py 25 61 62   (indices into engine.conv_plans, see tools/time_convs.py output order)"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sipmask_b200 import synth  # noqa: E402
from sipmask_b200.engine import SipMaskEngine  # noqa: E402

idx = [int(a) for a in sys.argv[1:]]
sd = synth.detector_state_dict(50, seed=1, cls_bias=bench.CLS_BIAS)
eng = SipMaskEngine(sd, (bench.H, bench.W), test_cfg=bench.TEST_CFG, img_shape=(bench.H, bench.IMG_W, 3), use_graph=False)
eng.forward(synth.synthetic_image(bench.H, bench.W, seed=0).cuda())
for i in idx:
    for _ in range(3):
        eng.conv_plans[i].run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in idx:
    eng.conv_plans[i].run()
    print(i, eng.conv_meta[i])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
