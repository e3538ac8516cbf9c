"""One eager (un-graphed) forward of the bench workload between cudaProfilerStart/Stop, for ncu:

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm -c 3 \
      -o gpurun_out/conv python tools/profile_step.py
"""
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
img = synth.synthetic_image(bench.H, bench.W, seed=0).cuda()
for _ in range(2):
    eng.forward(img)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.forward(img)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('detections', int(eng.count[0]))
