"""Marginal cost of each op class with several images in flight: replay the step with one op class left out
(results are garbage, timing only) and compare against the full step."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sipmask_b200 import synth  # noqa: E402
from sipmask_b200.serving import make_engines  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
sd = synth.detector_state_dict(50, seed=1, cls_bias=bench.CLS_BIAS)
engs = make_engines(sd, (bench.H, bench.W), in_flight=n, test_cfg=bench.TEST_CFG, img_shape=(bench.H, bench.IMG_W, 3), use_graph=False)
for i, e in enumerate(engs):
    e.forward(synth.synthetic_image(bench.H, bench.W, seed=i).cuda())
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in engs]
names = sorted(set(x for x in engs[0].op_names if x not in ('fork', 'join')))
K = 20 * n


def run(only):
    graphs = []
    for e, st in zip(engs, streams):
        with torch.cuda.stream(st):
            e._run_ops(only=only)
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                e._run_ops(only=only)
        graphs.append(g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    for rep in range(2):
        e0.record()
        for st in streams:
            st.wait_event(e0)
        for k in range(K):
            with torch.cuda.stream(streams[k % n]):
                graphs[k % n].replay()
        for st in streams:
            main.wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3


full = run(set(names))
print('full step: %.1f us / image (%d in flight)' % (full, n))
for nm in names:
    t = run(set(names) - {nm})
    print('  without %-16s %.1f us  (marginal %.1f us, %d launches)' % (nm, t, full - t, engs[0].op_names.count(nm)), flush=True)
t = run({'conv'})
print('  conv only              %.1f us' % t)
