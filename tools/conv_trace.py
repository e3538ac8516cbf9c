"""Per-role clock64 trace of CTA 0 inside conv_gemm_kernel (needs the -DSMB_TRACE build: tools/build_trace_lib.sh).

    SMB_LIB_PATH=tools/_trace/libsipmask_b200_trace.so python tools/conv_trace.py <conv name substring> [cap]

Prints, for the first tiles of CTA 0, when each warp role passed its synchronisation points (cycles from the kernel's first
stamp), so the steady-state period of a tile can be attributed to a role instead of inferred from ablations.
"""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ts = torch.zeros(512, dtype=torch.int64, device='cuda')
os.environ['SMB_CONV_TS'] = hex(ts.data_ptr())
import bench  # noqa: E402
from sipmask_b200 import synth  # noqa: E402
from sipmask_b200.engine import SipMaskEngine  # noqa: E402

pat = sys.argv[1]
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sd = synth.detector_state_dict(50, seed=1, cls_bias=bench.CLS_BIAS)
eng = SipMaskEngine(sd, (bench.H, bench.W), test_cfg=bench.TEST_CFG, img_shape=(bench.H, bench.IMG_W, 3), use_graph=False)
eng.forward(synth.synthetic_image(bench.H, bench.W, seed=0).cuda())
torch.cuda.synchronize()
for i, m in enumerate(eng.conv_meta):
    if pat not in m['name']:
        continue
    pl = eng.conv_plans[i]
    if cap:
        pl.set_max_ctas(cap)
    for _ in range(5):
        pl.run()
    torch.cuda.synchronize()
    ts.zero_()
    pl.run()
    torch.cuda.synchronize()
    t = ts.cpu().tolist()
    t0 = t[0]
    r = lambda role, lt, e: (t[64 + role * 64 + lt * 4 + e] - t0) if t[64 + role * 64 + lt * 4 + e] else -1
    print('=== %s M=%d N=%d K=%d cap=%d debug=%s' % (m['name'], m['M'], m['N'], m['K'], cap, os.environ.get('SMB_CONV_DEBUG', '0')))
    print('   setup done=%d  kernel end(thread 0)=%d' % (t[1] - t0, t[11] - t0))
    print('   tile | producer: empty-ok  issued | mma: tempty-ok  full-ok  committed | epi(w2): top  tfull-ok  end | store warp: c0 sfull  c0 done  cLast sfull  cLast done')
    for lt in range(16):
        if r(2, lt, 0) < 0 and r(1, lt, 0) < 0:
            break
        print('   %4d | %8d %8d | %8d %8d %8d | %8d %8d %8d | %8d %8d %8d %8d' % (
            lt, r(0, lt, 0), r(0, lt, 1), r(1, lt, 0), r(1, lt, 1), r(1, lt, 2), r(2, lt, 0), r(2, lt, 1), r(2, lt, 2),
            r(3, lt, 0), r(3, lt, 1), r(3, lt, 2), r(3, lt, 3)))
    print('   tile 2 chunks (epilogue warp 2): slot ready | tmem half 0 loaded | half 0 staged | tmem half 1 loaded | math+sts done | fenced | arrived')
    for c in range(4):
        if r(4, c, 0) >= 0:
            print('      chunk %d: %8d %8d %8d %8d %8d %8d %8d' % (c, r(4, c, 0), r(4, c, 1), r(5, c, 1), r(5, c, 0), r(4, c, 2), r(5, c, 2), r(4, c, 3)))
    break
