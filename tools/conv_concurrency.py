"""How do the conv shapes of the path behave when several images are in flight?  For each shape: S streams, each replaying a
graph of R launches of its own plan (private activations, shared weights, grid capped), vs one stream uncapped.
Reports time per conv, TF/s and the L2->SM operand bytes/s it implies (A 16 KB + B per k-block per CTA)."""
import math
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sipmask_b200 import conv as C  # noqa: E402

torch.manual_seed(0)
dev = torch.device('cuda')
# name, H, W, Cin, Cout, k, residual
SHAPES = [('layer1.conv2', 200, 336, 64, 64, 3, False), ('layer1.conv3', 200, 336, 64, 256, 1, True),
          ('layer1.conv1', 200, 336, 256, 64, 1, False), ('layer2.conv2', 100, 168, 128, 128, 3, False),
          ('layer2.conv3', 100, 168, 128, 512, 1, True), ('layer3.conv1', 50, 84, 1024, 256, 1, False),
          ('layer3.conv2', 50, 84, 256, 256, 3, False), ('layer3.conv3', 50, 84, 256, 1024, 1, True),
          ('layer4.conv2', 25, 42, 512, 512, 3, False), ('fpn/tower P3', 100, 168, 256, 256, 3, False)]
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
CAP = int(sys.argv[2]) if len(sys.argv) > 2 else 64
MT = int(sys.argv[3]) if len(sys.argv) > 3 else 16
R = 20


def bench(plans, streams):
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
    for rep in range(2):
        e0.record()
        for g, st in zip(graphs, streams):
            st.wait_event(e0)
            with torch.cuda.stream(st):
                g.replay()
                g.replay()
        for st in streams:
            main.wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * R * len(plans))          # us per conv


streams = [torch.cuda.Stream() for _ in range(S)]
print('%-14s %6s %5s %5s | serial us  TF/s | x%d cap%d us/conv  TF/s  L2->SM TB/s' % ('shape', 'M', 'N', 'K', S, CAP))
for name, H, W, cin, cout, k, res in SHAPES:
    w = (torch.randn(cout, cin, k, k) * 0.05)
    wk, _ = C.pack_weight(w, device=dev)
    bias = torch.zeros(cout, device=dev)

    def mk(cap, mt):
        prev = C.set_min_tiles(mt)
        x = (torch.randn(1, H, W, cin, device=dev) * 0.5).half()
        out = torch.empty(1, H, W, cout, device=dev, dtype=torch.float16)
        r = (torch.randn(1, H, W, cout, device=dev) * 0.5).half() if res else None
        p = C.ConvPlan(x, wk, out, k, 1, relu=True, bias=bias, residual=r)
        C.set_min_tiles(prev)
        if cap:
            p.set_max_ctas(cap)
        return p

    t1 = bench([mk(None, 48)], streams[:1])
    tS = bench([mk(CAP, MT) for _ in range(S)], streams)
    M, K = H * W, cin * k * k
    fl = 2.0 * M * cout * K
    tm = math.ceil(M / 128)
    nt = cout if cout <= 256 else 256
    for c in ([cout] if cout <= 256 else [256]) + ([128] if cout > 128 and cout % 128 == 0 else []) + ([64] if cout > 64 and cout % 64 == 0 else []):
        if tm * math.ceil(cout / c) >= MT:
            nt = c
            break
    else:
        nt = 64 if cout > 64 and cout % 64 == 0 else nt
    byt = tm * math.ceil(cout / nt) * (K // 64) * (16384 + nt * 64)
    print('%-14s %6d %5d %5d | %8.2f %6.0f | %14.2f %6.0f %8.2f   (nt=%d)' % (name, M, cout, K, t1, fl / t1 / 1e6, tS, fl / tS / 1e6,
                                                                          byt / tS / 1e6, nt), flush=True)
