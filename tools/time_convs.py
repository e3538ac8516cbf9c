"""Warm, back-to-back timing (CUDA events) of every convolution plan of the bench workload: shape, us, TFLOP/s."""
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
print('pairs above score_thr: %d of %d candidates x %d classes' % (int((eng.cand_scores[0] > 0.05).sum()), eng.ncand, eng.ncls))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot_w = tot_c = 0.0
print('%-44s %8s %5s %6s %3s %8s %8s %8s' % ('conv', 'M', 'N', 'K', 'k', 'warm_us', 'cold_us', 'TF/s(w)'))
REPS = 20
for plan, m in zip(eng.conv_plans, eng.conv_meta):
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    # warm: REPS back-to-back launches replayed from a CUDA graph (no Python / driver launch cost in the number)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            plan.run()
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) / REPS * 1e3
    cold = []
    for i in range(3):
        flush.fill_(i)
        e0.record()
        plan.run()
        e1.record()
        torch.cuda.synchronize()
        cold.append(e0.elapsed_time(e1) * 1e3)
    cold = sorted(cold)[1]
    tot_w += warm
    tot_c += cold
    print('%-44s %8d %5d %6d %3d %8.1f %8.1f %8.1f %s%s' % (m['name'][-44:], m['M'], m['N'], m['K'], m['k'], warm, cold,
                                                         m['flops'] / warm / 1e6, 'R' if m['res'] else '', 'G' if m['gn'] else ''))
print('sum warm(graph) %.1f us, sum cold %.1f us, %.1f GFLOP' % (tot_w, tot_c, eng.conv_flops / 1e9))
