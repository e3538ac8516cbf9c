"""Does an HBM-bound convolution overlap with a tensor-bound one?  layer1.conv3 (1x1, K=64, +residual: 77 MB, ~0.1 PF/s) and
a tower conv (3x3, K=2304: 26 GF, ~1.3 PF/s) each replay R launches on their own stream with the persistent grid capped at
CAP CTAs: alone, two of a kind, and mixed.  If the mixed pair costs about max(alone_a, alone_b) the bounds overlap; if it costs
about the sum they do not (then the per-image conv time is the sum of per-kind times whatever the schedule)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sipmask_b200 import conv as C  # noqa: E402

dev = torch.device('cuda')
R = 20


def mk(H, W, cin, cout, k, res, cap, mt=16):
    prev = C.set_min_tiles(mt)
    w = torch.randn(cout, cin, k, k) * 0.05
    wk, _ = C.pack_weight(w, device=dev)
    x = (torch.randn(1, H, W, cin, device=dev) * 0.5).half()
    out = torch.empty(1, H, W, cout, device=dev, dtype=torch.float16)
    r = (torch.randn(1, H, W, cout, device=dev) * 0.5).half() if res else None
    p = C.ConvPlan(x, wk, out, k, 1, relu=True, bias=torch.zeros(cout, device=dev), residual=r)
    C.set_min_tiles(prev)
    if cap:
        p.set_max_ctas(cap)
    p._hold = (wk, x, out, r)
    return p


def bench(plans):
    streams = [torch.cuda.Stream() for _ in plans]
    graphs = []
    for p, st in zip(plans, streams):
        with torch.cuda.stream(st):
            p.run()
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(R):
                    p.run()
        graphs.append(g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    best = 1e9
    for rep in range(3):
        e0.record()
        for g, st in zip(graphs, streams):
            st.wait_event(e0)
            with torch.cuda.stream(st):
                g.replay()
        for st in streams:
            main.wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / R)
    return best                                   # us per "round" (one launch of every plan)


for cap in (48, 64, 74):
    a = lambda: mk(200, 336, 64, 256, 1, True, cap)       # HBM-bound   # noqa: E731
    b = lambda: mk(100, 168, 256, 256, 3, False, cap)     # tensor-bound  # noqa: E731
    ta, tb = bench([a()]), bench([b()])
    taa, tbb, tab = bench([a(), a()]), bench([b(), b()]), bench([a(), b()])
    t3 = bench([a(), b(), b()])
    print('cap %3d: hbm alone %.1f us, tensor alone %.1f us | hbm+hbm %.1f, tensor+tensor %.1f, hbm+tensor %.1f (sum %.1f, max %.1f) | '
          'hbm+2 tensor %.1f' % (cap, ta, tb, taa, tbb, tab, ta + tb, max(ta, tb), t3), flush=True)
