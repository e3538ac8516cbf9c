#!/bin/bash
# mask kernels, second pass: 4 CTAs / SM (64 registers), dense tile shapes 8x64 / 4x128 / 2x256, fused tile height 16 / 8.
# Times every configuration, tests every configuration, then runs bench / ncu / the engine suites with the fastest one.
mkdir -p gpurun_out
: > gpurun_out/r2x_mask_times.jsonl
run() { env "$@" timeout 120 python tools/time_masks.py >> gpurun_out/r2x_mask_times.jsonl 2>> gpurun_out/r2x_mask_times.err; }
run SMB_MASK_MMA=0
run SMB_MASK_MMA=1 SMB_MASK_TILE=0
run SMB_MASK_MMA=1 SMB_MASK_TILE=1
run SMB_MASK_MMA=1 SMB_MASK_TILE=2
run SMB_MASK_MMA=1 SMB_MASK_TILE=0 SMB_MASK_FUSED_TY=8
cat gpurun_out/r2x_mask_times.jsonl | cut -c1-260
python - <<'PY' > gpurun_out/r2x_best.env
import json
rows = [json.loads(l) for l in open('gpurun_out/r2x_mask_times.jsonl') if l.strip()]
mma = [r for r in rows if r['env'].get('SMB_MASK_MMA') == '1']
tile = min((r for r in mma if 'SMB_MASK_FUSED_TY' not in r['env']), key=lambda r: r['dense_ms'])['env']['SMB_MASK_TILE']
ty = min(mma, key=lambda r: r['fused_ms'])['env'].get('SMB_MASK_FUSED_TY', '16')
print('export SMB_MASK_MMA=1 SMB_MASK_TILE=%s SMB_MASK_FUSED_TY=%s' % (tile, ty))
PY
cat gpurun_out/r2x_best.env
for t in 0 1 2; do
  SMB_MASK_MMA=1 SMB_MASK_TILE=$t SMB_MASK_FUSED_TY=$((16 - 4 * t)) timeout 200 python -m pytest tests/test_gpu_postproc.py -k "mask" -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 > gpurun_out/r2x_pytest_mask_tile$t.txt
  echo "tile $t (fused TY $((16 - 4 * t))):"; tail -1 gpurun_out/r2x_pytest_mask_tile$t.txt
done
source gpurun_out/r2x_best.env
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2x_bench_A.json 2> gpurun_out/r2x_bench_A.err
python -c "
import json; d=json.load(open('gpurun_out/r2x_bench_A.json')); print('A', round(d['value'],1), round(d['e2e']['value'],1), d['ms_per_step'], d['roofline']['frac'], d['clocks'])
for k in ('roofline_mask_assembly','roofline_mask_fused'): print(k, d[k]['ms'], d[k]['frac'], d[k]['other_family'])
print(d['serial'], d['e2e_dropin']['ms_per_image'])"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'mask_' -f -o gpurun_out/r2x_aux python tools/ncu_aux.py > gpurun_out/r2x_ncu.log 2>&1
ncu -i gpurun_out/r2x_aux.ncu-rep --page raw --csv > gpurun_out/r2x_aux_raw.csv 2>> gpurun_out/r2x_ncu.log
python tools/ncu_summary.py gpurun_out/r2x_aux_raw.csv > gpurun_out/r2x_aux_summary.txt 2>&1; cat gpurun_out/r2x_aux_summary.txt | cut -c1-160
rm -f gpurun_out/r2x_aux.ncu-rep
timeout 500 python -m pytest tests/test_gpu_engine.py tests/test_head_api.py tests/test_gpu_vis.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r2x_pytest_engine.txt
tail -3 gpurun_out/r2x_pytest_engine.txt
