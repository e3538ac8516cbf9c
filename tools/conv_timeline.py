"""clock64 timeline of CTA 0 inside conv_gemm_kernel for selected plans (profiling aid, SMB_CONV_TS)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ts = torch.zeros(16, dtype=torch.int64, device='cuda')
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
