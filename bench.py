#!/usr/bin/env python
"""Benchmark of the SipMask inference hot path (BASELINE.json metric: images/sec @ 800x1333, bs=1/GPU).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU forward (oracle port)

One "step" = one pass of the whole hot path (image -> backbone/FPN -> head -> decode/NMS -> mask assembly ->
bit-packed masks) over one synthetic 800x1344 image per GPU.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = 'SipMask R50-FPN-GN 4conv, 800x1333 (padded 800x1344), bs=1/GPU, synthetic image + seeded synthetic weights'
H, W, IMG_W = 800, 1344, 1333


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0), 'fallback'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md 'clocks line')."""

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index),
                 '--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
                 'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
                 'clocks_event_reasons.sw_power_cap', '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------------- CPU reference arm
def build_oracle(threads):
    import torch
    from oracle import model as M
    from oracle import ops as O
    from sipmask_b200 import synth
    torch.set_num_threads(threads)
    O.USE_TORCHVISION_DCN = True
    O.USE_C_CROP_SPLIT = True
    net = M.SipMaskDetector(50)
    net.load_state_dict(synth.detector_state_dict(50, seed=1, cls_bias=CLS_BIAS), strict=True)
    net.eval()
    return net


ROLL = 500          # untimed pre/post-roll steps around the timed region while nvidia-smi samples clocks
CLS_BIAS = -5.0
TEST_CFG = dict(nms_pre=1000, score_thr=0.05, nms=dict(type='nms', iou_thr=0.5), max_per_img=100)


def oracle_step(net, img, rows=H):
    """The reference's forward on host cores: backbone -> FPN -> head -> get_bboxes (decode, per-class NMS with the
    compiled C oracle, dense 4x matmul/sigmoid/stack/CropSplit mask assembly, x2 upsample + threshold)."""
    import torch
    from oracle import cbind, postproc as P
    from oracle import ops as O
    with torch.no_grad():
        cls, box, ctr, cof, fm = net(img)
    real_nms = O.nms
    O.nms = lambda dets, thr, cmp_ge=False, plus_one=True: cbind.nms(dets, thr, int(cmp_ge), int(plus_one))
    try:
        res = P.get_bboxes_single([t[0] for t in cls], [t[0] for t in box], [t[0] for t in ctr], [t[0] for t in cof], fm[0],
                                  (8, 16, 32, 64, 128), (rows, IMG_W, 3), (rows, IMG_W, 3), 1.0, TEST_CFG, rescale=True)
    finally:
        O.nms = real_nms
    return res


def cpu_reference(steps, warmup, threads, budget_s=200.0):
    """Times `steps` steps after `warmup`.  A step is one full 800x1344 image when the whole run fits the time budget;
    otherwise a horizontal strip of `rows` image rows (rows/800 of an image - the path is convolutional, cost is linear
    in rows), so that the run stays bounded.  Returns (seconds per FULL image, detections, rows, step times)."""
    from sipmask_b200 import synth
    net = build_oracle(threads)
    img = synth.synthetic_image(H, W, seed=0)
    t0 = time.perf_counter()
    res = oracle_step(net, img)                       # probe = first warm-up step, always a full image
    t_probe = time.perf_counter() - t0
    rows = H
    total = (steps + max(warmup - 1, 0)) * t_probe
    if total > budget_s:
        rows = int(max(64, min(H, round(H * budget_s / total / 32.0) * 32)))
    sample = img[:, :, :rows].contiguous()
    for _ in range(max(warmup - 1, 0)):
        oracle_step(net, sample, rows)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        res = oracle_step(net, sample, rows)
        ts.append(time.perf_counter() - t0)
    sec_per_image = (sum(ts) / len(ts)) * (float(H) / rows)
    return sec_per_image, int(res['det_bboxes'].shape[0]), rows, ts


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sec, ndet, rows, ts = cpu_reference(args.steps, args.warmup, threads)
    val = 1.0 / sec
    sample = ('%d steps, each %d of 800 rows of the 800x1344 image (%.3f image) through the oracle = PyTorch CPU fp32 '
              'restatement of the reference forward incl. get_bboxes; the literal reference cannot run on CPU '
              '(DeformConv/CropSplit are CUDA-only); value = (rows/800) / mean step time' % (args.steps, rows, rows / float(H)))
    line = dict(impl='reference', metric='images/sec', value=val, unit='images/s', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=(sum(ts) / len(ts)) * 1e3, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f32', data='synthetic', config=dict(workload=WORKLOAD, detections=ndet, rows_per_step=rows),
                cpu_baseline=dict(value=val, unit='images/s', cores=threads, kind='port', sample=sample),
                e2e=dict(value=val, unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from sipmask_b200 import dist as sdist
    from sipmask_b200 import ops, synth
    from sipmask_b200.engine import SipMaskEngine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU fallback for the product path)'
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)

    sd = synth.detector_state_dict(50, seed=1, cls_bias=CLS_BIAS)
    from sipmask_b200.serving import PipelinedRunner, EnginePool, make_engines
    nfl = max(1, args.in_flight)
    # nfl images in flight per GPU: nfl engines (shared weights, private activations + CUDA graph), one image per forward
    engs = make_engines(sd, (H, W), in_flight=nfl, test_cfg=TEST_CFG, img_shape=(H, IMG_W, 3), use_graph=True, device=dev)
    eng = engs[0]
    img_host = synth.synthetic_image(H, W, seed=rank).pin_memory()        # one image per GPU (weak scaling)
    for e in engs:
        e.img.copy_(img_host, non_blocking=True)
    torch.cuda.synchronize()
    pool = EnginePool(engs)
    rec = torch.zeros((world, eng.max_num, 7), dtype=torch.float32, device=dev) if world > 1 else None

    def gather(out):
        # the single collective of the path: fixed-shape detection record, on the main stream
        sdist.gather_records(sdist.pack_record(out['det_bboxes'][0], out['det_labels'][0], out['count']), out=rec)

    gather_fn = gather if world > 1 else None

    def step():
        pool.step(gather_fn)

    def step_finish():
        pool.flush(gather_fn)

    # end-to-end through the public serving API: pinned-host image in, pinned-host record + bit-packed masks out, the
    # upload / replay / download of consecutive images overlapped on their own streams (sipmask_b200/serving.py)
    runner = PipelinedRunner(engs)
    copy_done = []

    def step_e2e():
        copy_done.append(runner.step(img_host, gather_fn))

    def e2e_finish():
        runner.flush(gather_fn)       # the timed region ends when the last result is on the host

    def timed(fn, steps, sample_clocks=False, finish=None):
        """K steps between barrier+synchronize, CUDA events, max over ranks.  nvidia-smi samples clocks every 100 ms;
        a short timed region would get no sample, so ROLL untimed steps of the same load run before and after it and
        the sampler stays on throughout (clocks.window says so)."""
        sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        if sample_clocks:
            for _ in range(ROLL):                       # fixed count: every rank must issue the same collectives
                fn()
            if finish is not None:
                finish()
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()                                # e.g. make the compute stream wait for the last downloads
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if sample_clocks:
            for _ in range(ROLL):
                fn()
            if finish is not None:
                finish()
            torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        if clocks is not None:
            clocks['window'] = 'timed region plus %d untimed steps of the identical load before and after it' % ROLL
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), clocks

    for _ in range(max(args.warmup, 3)):
        step()
    step_finish()
    torch.cuda.synchronize()
    total_ms, clocks = timed(step, args.steps, sample_clocks=True, finish=step_finish)
    ms_per_step = total_ms / args.steps
    value = world * 1000.0 / ms_per_step
    for _ in range(3):
        step_e2e()
    e2e_finish()
    e2e_ms, _ = timed(step_e2e, args.steps, finish=e2e_finish)
    e2e_value = world * 1000.0 / (e2e_ms / args.steps)
    last = runner.result(copy_done[-1])
    assert int(last['cnt'][0]) == int(eng.count[0].item())
    h2d, d2h = runner.h2d_bytes, runner.d2h_bytes
    ndet = int(eng.count[0].item())

    line = dict(metric='images/sec', value=value, unit='images/s', n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms_per_step, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f16',
                data='synthetic',
                config=dict(workload=WORKLOAD, parallelism='dp%d (one image per GPU, one all-gather of the detection record)' % world,
                            detections_per_image=ndet, cuda_graph=True, images_in_flight=nfl, batch_per_forward=1,
                            l2='no flush: one step streams ~1.3 GB of activations/masks through a 126 MB L2, so nothing but '
                               'weights (51 MB) can survive from the previous step'),
                clocks=clocks, e2e=dict(value=e2e_value, unit='images/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                                        ms_per_step=e2e_ms / args.steps),
                gpu_launches=eng.n_launch * args.steps)
    torch.cuda.synchronize()

    if rank == 0:
        pk, pk_kind = peaks()
        # ---- roofline of the dominant kernel (conv_gemm_kernel): the conv launches of one step of EVERY engine in flight,
        # captured per engine (same stream schedule as in the step) and replayed concurrently on the pool's streams
        graphs = []
        for e, st in zip(engs, pool.streams):
            with torch.cuda.stream(st):
                e._run_ops(only={'conv'})
                st.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    e._run_ops(only={'conv'})
            graphs.append(g)
        torch.cuda.synchronize()

        def conv_round():
            for g, st in zip(graphs, pool.streams):
                with torch.cuda.stream(st):
                    g.replay()

        for _ in range(3):
            conv_round()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, min(args.steps, 20))
        e0.record()
        for st in pool.streams:
            st.wait_event(e0)
        for _ in range(reps):
            conv_round()
        pool.join()
        e1.record()
        torch.cuda.synchronize()
        conv_ms = e0.elapsed_time(e1) / (reps * nfl)                      # per image
        tf = eng.conv_flops / (conv_ms * 1e-3) / 1e12
        peak_tf = float(pk.get('bf16_tflops_sustained', pk.get('bf16_tflops', 1400.0)))
        line['roofline'] = dict(bound='tensor', kernel='conv_gemm_kernel (%d launches/step)' % len(eng.conv_plans),
                                achieved=tf, peak=peak_tf, unit='TFLOP/s', frac=tf / peak_tf, traffic=None,
                                peak_source=pk_kind + ' bf16_tflops_sustained', algorithmic_gflop_per_step=eng.conv_flops / 1e9,
                                ms_per_step=conv_ms, share_of_step=conv_ms / ms_per_step,
                                note='conv launches of %d images in flight replayed concurrently; time per image' % nfl)
        del graphs
        # ---- strictly serial reference point: ONE engine tuned for a single stream, one image at a time
        if nfl > 1:
            e1s = make_engines(sd, (H, W), in_flight=1, test_cfg=TEST_CFG, img_shape=(H, IMG_W, 3), use_graph=True, device=dev)[0]
            e1s.img.copy_(img_host, non_blocking=True)
            for _ in range(5):
                e1s.forward(None)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(max(args.steps, 20)):
                e1s.forward(None)
            e1.record()
            torch.cuda.synchronize()
            sms = e0.elapsed_time(e1) / max(args.steps, 20)
            line['serial'] = dict(ms_per_step=sms, value=1000.0 / sms, unit='images/s',
                                  note='one image in flight (latency-optimal planner settings), same GPU, N=1 rank only')
            del e1s
        # ---- mask assembly (BASELINE metric part 2): HBM GB/s of the fused kernel, N = max_per_img detections
        N = eng.max_num
        Hm, Wm = eng.protos.shape[1], eng.protos.shape[2]
        gen = torch.Generator().manual_seed(0)
        cofs = torch.randn(N, 128, generator=gen).to(dev)
        cx, cy = torch.rand(N, generator=gen) * IMG_W, torch.rand(N, generator=gen) * H
        bw, bh = torch.rand(N, generator=gen) * 480 + 32, torch.rand(N, generator=gen) * 480 + 32
        boxes = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1).clamp(min=0).to(dev)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        pos = torch.empty((N, Hm, Wm), dtype=torch.float32, device=dev)
        bits = torch.empty((N, H, (IMG_W + 31) // 32), dtype=torch.int32, device=dev)

        def time_kernel(fn):
            ts = []
            for i in range(8):
                flush.fill_(i)                                 # evict L2 between timed launches
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return statistics.median(ts[2:])

        # (a) the reference-shaped output: dense pos_masks [N,Hm,Wm] fp32 (what CropSplit returns, permuted)
        ma_ms = time_kernel(lambda: ops.mask_assemble(eng.protos[0], cofs, boxes, 0.5, layout='hwc', out=pos))
        ma_bytes = Hm * Wm * 32 * 2 + N * 128 * 4 + N * 16 + N * Hm * Wm * 4
        gbs = ma_bytes / (ma_ms * 1e-3) / 1e9
        line['roofline_mask_assembly'] = dict(bound='hbm', kernel='mask_assemble_kernel', achieved=gbs, peak=float(pk['hbm_gbs']),
                                              unit='GB/s', frac=gbs / float(pk['hbm_gbs']), traffic=None, ms=ma_ms,
                                              algorithmic_bytes=ma_bytes, peak_source=pk_kind + ' hbm_gbs',
                                              note='protos fp16 HWC read once + fp32 pos_masks [100,400,672] written; L2 flushed between launches')
        # (b) the kernel the engine runs: fused assembly + x2 upsample + threshold + bit-pack (no pos_masks traffic)
        mf_ms = time_kernel(lambda: ops.mask_assemble_pack(eng.protos[0], cofs, boxes, 0.5, (H, IMG_W), 0.4, layout='hwc', out=bits))
        mf_bytes = Hm * Wm * 32 * 2 + N * 128 * 4 + N * 16 + bits.numel() * 4
        mgbs = mf_bytes / (mf_ms * 1e-3) / 1e9
        line['roofline_mask_fused'] = dict(bound='hbm', kernel='mask_fused_pack_kernel', achieved=mgbs, peak=float(pk['hbm_gbs']),
                                           unit='GB/s', frac=mgbs / float(pk['hbm_gbs']), traffic=None, ms=mf_ms,
                                           algorithmic_bytes=mf_bytes, peak_source=pk_kind + ' hbm_gbs',
                                           note='protos fp16 read once + bit-packed [100,800,42] int32 masks written')
        # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload on the host cores
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sec, _, rows, _ = cpu_reference(1, 1, threads)
            line['cpu_baseline'] = dict(value=1.0 / sec, unit='images/s', cores=threads, kind='port',
                                        sample='1 warm-up + 1 timed pass of %d/800 image rows through the oracle '
                                               '(PyTorch CPU fp32 restatement of the reference forward incl. get_bboxes)' % rows)
        else:
            line['cpu_baseline'] = None
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--in-flight', type=int, default=int(os.environ.get('SMB_IN_FLIGHT', '6')),
                    help='images in flight per GPU (independent batch-1 forwards on separate streams); 1 = strictly serial')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
