"""What paces the short-K 1x1 convolutions?  layer1.conv3 (K=64, N=256, + residual) and layer1 downsample (same, no residual)
at 48 CTAs and uncapped, with parts of the tile's memory traffic switched off / re-routed at plan creation:
    base                      TMA residual load + TMA store
    SMB_CONV_DEBUG=8          no output store at all
    SMB_CONV_NO_TMA_RES=1     residual via per-thread LDG (no TMA residual tile)
    SMB_CONV_NO_TMA_STORE=1   direct per-thread STG / LDG epilogue (no staging ring)
    SMB_CONV_STAGE_SETS=1     staging ring of one tile instead of two
    SMB_CONV_PAIR=0           single-CTA MMA (no cta_group::2)
Prints us per launch (CUDA-graph replay of 20 launches, best of 3)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sipmask_b200 import conv as C  # noqa: E402

dev = torch.device('cuda')
R = 20


def mk(res, cap, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        prev = C.set_min_tiles(16)
        H, W, cin, cout = 200, 336, 64, 256
        w = torch.randn(cout, cin, 1, 1) * 0.05
        wk, _ = C.pack_weight(w, device=dev)
        x = (torch.randn(1, H, W, cin, device=dev) * 0.5).half()
        out = torch.empty(1, H, W, cout, device=dev, dtype=torch.float16)
        r = (torch.randn(1, H, W, cout, device=dev) * 0.5).half() if res else None
        p = C.ConvPlan(x, wk, out, 1, 1, relu=True, bias=torch.zeros(cout, device=dev), residual=r)
        C.set_min_tiles(prev)
        if cap:
            p.set_max_ctas(cap)
        p._hold = (wk, x, out, r)
        return p
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


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


VARIANTS = [('base', {}), ('epilogue releases the accumulator untouched (256)', {'SMB_CONV_DEBUG': '256'}),
            ('chunk loop = wait + arrive only (760)', {'SMB_CONV_DEBUG': '760'}),
            ('pair mode off', {'SMB_CONV_PAIR': '0'}), ('pair off, chunk loop skipped', {'SMB_CONV_PAIR': '0', 'SMB_CONV_DEBUG': '256'}), ('no store (DEBUG=8)', {'SMB_CONV_DEBUG': '8'}), ('residual via LDG', {'SMB_CONV_NO_TMA_RES': '1'}),
            ('direct STG epilogue', {'SMB_CONV_NO_TMA_STORE': '1'}), ('one staging set', {'SMB_CONV_STAGE_SETS': '1'}),
            ('no store + residual via LDG', {'SMB_CONV_DEBUG': '8', 'SMB_CONV_NO_TMA_RES': '1'}),
            ('no TMEM loads (128)', {'SMB_CONV_DEBUG': '128'}), ('no bias/res math (64)', {'SMB_CONV_DEBUG': '64'}),
            ('no staging writes (32)', {'SMB_CONV_DEBUG': '32'}), ('no proxy fence (16)', {'SMB_CONV_DEBUG': '16'}),
            ('no tmem/math/sts/fence (240)', {'SMB_CONV_DEBUG': '240'}), ('no tmem/math/sts/fence/store (248)', {'SMB_CONV_DEBUG': '248'}),
            ('no MMA (1): TMA + epilogue only', {'SMB_CONV_DEBUG': '1'}),
            ('store lag 2', {'SMB_CONV_STORE_LAG': '2'}), ('store lag 3', {'SMB_CONV_STORE_LAG': '3'}),
            ('store lag 4', {'SMB_CONV_STORE_LAG': '4'}), ('store lag 4, no store', {'SMB_CONV_STORE_LAG': '4', 'SMB_CONV_DEBUG': '8'}),
            ('store lag 3, residual via LDG', {'SMB_CONV_STORE_LAG': '3', 'SMB_CONV_NO_TMA_RES': '1'})]
if os.environ.get('SMB_DIAG_QUICK'):
    VARIANTS = [v for v in VARIANTS if v[0] in ('base', 'no tmem/math/sts/fence/store (248)') or '256' in str(v[1]) or '760' in str(v[1]) or 'pair' in v[0]]
for res in (True, False):
    for cap in (48, None):
        print('--- layer1 %s, %s' % ('conv3 (+residual, 77 MB)' if res else 'downsample (no residual, 43 MB)', 'cap %d CTAs' % cap if cap else 'uncapped'))
        for name, env in VARIANTS:
            if not res and 'residual' in name:
                continue
            try:
                t = bench(mk(res, cap, env))
                print('   %-32s %7.1f us' % (name, t), flush=True)
            except Exception as ex:
                print('   %-32s failed: %s' % (name, str(ex)[:100]), flush=True)
