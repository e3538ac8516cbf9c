"""Is the 3x3 (K = 9*C) convolution bound by the L2 -> SM operand ingress or by the MMA issue?  One tower-sized conv
(C = 256 -> 256, 100 x 168 and 200 x 336 maps) with the main loop's halves switched off:
    SMB_CONV_DEBUG=1   TMA operand pipeline only (no MMAs)        SMB_CONV_DEBUG=2   MMAs only (no TMA loads)
(both force single-CTA MMA, so the reference point is SMB_CONV_PAIR=0).  us per launch, CUDA-graph replay of 20."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sipmask_b200 import conv as C  # noqa: E402

dev = torch.device('cuda')
R = 20


def mk(H, W, cin, cout, k, env):
    old = {kk: os.environ.get(kk) for kk in env}
    os.environ.update(env)
    try:
        w = torch.randn(cout, cin, k, k) * 0.02
        wk, _ = C.pack_weight(w, device=dev)
        x = (torch.randn(1, H, W, cin, device=dev) * 0.5).half()
        out = torch.empty(1, H, W, cout, device=dev, dtype=torch.float16)
        p = C.ConvPlan(x, wk, out, k, 1, relu=True, bias=torch.zeros(cout, device=dev))
        p._hold = (wk, x, out)
        return p
    finally:
        for kk, v in old.items():
            if v is None:
                os.environ.pop(kk, None)
            else:
                os.environ[kk] = v


def bench(p):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        p.run()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(R):
                p.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / R)
    return best


for (H, W, cin, cout, k) in [(100, 168, 256, 256, 3), (200, 336, 64, 64, 3), (50, 84, 512, 512, 3), (100, 168, 512, 128, 1)]:
    fl = 2.0 * H * W * cin * cout * k * k
    print('--- %dx%d  %d -> %d  k=%d  (%.1f GFLOP)' % (H, W, cin, cout, k, fl / 1e9))
    for name, env in [('pair (product)', {}), ('single-CTA MMA', {'SMB_CONV_PAIR': '0'}),
                      ('TMA pipeline only (1)', {'SMB_CONV_DEBUG': '1'}), ('MMA only (2)', {'SMB_CONV_DEBUG': '2'}),
                      ('pair, lockstep epilogue', {'SMB_CONV_EPI_SPLIT': '0'})]:
        try:
            t = bench(mk(H, W, cin, cout, k, env))
            print('   %-28s %7.1f us  %7.1f TF/s' % (name, t, fl / t / 1e6))
        except Exception as e:  # noqa: BLE001
            print('   %-28s failed: %s' % (name, e))
