"""Throughput with several images in flight: n independent engines (own activations, own CUDA graph), replayed round-robin
on n streams.  Each forward is still one image through the whole path; the question is how much of the idle SM time in the
latency-bound small-map layers a second in-flight image can fill.

usage: python tools/time_inflight.py [n ...]          (env SMB_CONV_MAX_CTAS / SMB_CONV_MIN_TILES / SMB_HEAD_MAX_CTAS apply)
"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sipmask_b200 import synth  # noqa: E402
from sipmask_b200.engine import SipMaskEngine  # noqa: E402

ns = [int(a) for a in sys.argv[1:]] or [1, 2, 3]
sd = synth.detector_state_dict(50, seed=1, cls_bias=bench.CLS_BIAS)
K = 120
for n in ns:
    engs, streams = [], []
    for i in range(n):
        eng = SipMaskEngine(sd, (bench.H, bench.W), test_cfg=bench.TEST_CFG, img_shape=(bench.H, bench.IMG_W, 3), use_graph=True)
        img = synth.synthetic_image(bench.H, bench.W, seed=i).cuda()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            eng.forward(img)
            eng.forward()
        st.synchronize()
        engs.append(eng)
        streams.append(st)
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e0.record(main)
        for st in streams:
            st.wait_stream(main)
        for k in range(K):
            with torch.cuda.stream(streams[k % n]):
                engs[k % n].graph.replay()
        for st in streams:
            main.wait_stream(st)
        e1.record(main)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cnt = [int(e.count[0]) for e in engs]
    print('in_flight=%d: %.1f img/s  (%.3f ms per image, counts %s)  env cap=%s min_tiles=%s head_cap=%s' % (
        n, K / ms * 1e3, ms / K, cnt, os.environ.get('SMB_CONV_MAX_CTAS'), os.environ.get('SMB_CONV_MIN_TILES'),
        os.environ.get('SMB_HEAD_MAX_CTAS')), flush=True)
    del engs
