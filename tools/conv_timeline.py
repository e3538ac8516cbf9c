"""clock64 timeline of CTA 0 inside conv_gemm_kernel for selected plans (profiling aid, SMB_CONV_TS)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ts = torch.zeros(64, dtype=torch.int64, device='cuda')
os.environ['SMB_CONV_TS'] = hex(ts.data_ptr())
import bench  # noqa: E402
from sipmask_b200 import synth  # noqa: E402
from sipmask_b200.engine import SipMaskEngine  # noqa: E402
sd = synth.detector_state_dict(50, seed=1, cls_bias=bench.CLS_BIAS)
eng = SipMaskEngine(sd, (bench.H, bench.W), test_cfg=bench.TEST_CFG, img_shape=(bench.H, bench.IMG_W, 3), use_graph=False)
eng.forward(synth.synthetic_image(bench.H, bench.W, seed=0).cuda())
names = ['start', 'setup done', 'prod first', 'mma loop start', 'mma kb0 full', 'mma kb8', 'mma kb16', 'mma kb32', 'mma last commit',
         'epi tfull', 'epi done', 'teardown']
for i in [int(a) for a in sys.argv[1:]]:
    for _ in range(3):
        eng.conv_plans[i].run()
    torch.cuda.synchronize()
    ts.zero_()
    eng.conv_plans[i].run()
    torch.cuda.synchronize()
    t = ts.cpu().tolist()
    m = eng.conv_meta[i]
    print('plan %d %s M=%d N=%d K=%d' % (i, m['name'], m['M'], m['N'], m['K']))
    print('   ' + '  '.join('%s=%d' % (n, (t[j] - t[0]) if t[j] else -1) for j, n in enumerate(names)))
    if t[16]:
        e = lambda j: (t[j] - t[16]) if t[j] else -1
        print('   tile#1 epilogue (warp 2, cycles from its loop top): prefetch/wait_read=%d bias-bar=%d tfull=%d' % (e(17), e(18), e(19)))
        for c in range(4):
            if t[20 + 4 * c]:
                print('      chunk %d: slot+residual=%d tmem-ld(h0)=%d fenced+arrived=%d' % (c, e(20 + 4 * c), e(21 + 4 * c),
                                                                          e(22 + 4 * c)))
        print('      tile end=%d   last chunk both halves staged=%d' % (e(36), e(41)))
