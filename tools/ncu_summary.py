"""ncu `--page raw --csv` export -> a text summary per launch and (optionally) the per-kernel DRAM traffic JSON that
bench.py reads (profiles/r02_ncu_traffic.json).

    python tools/ncu_summary.py gpurun_out/r02_aux_raw.csv [--json profiles/r02_ncu_traffic.json --steps S]
"""
import csv
import json
import sys

COLS = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
        ('sm__inst_executed_pipe_tensor.sum', 'tensor_inst'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'), ('launch__registers_per_thread', 'regs'),
        ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('smsp__inst_executed.sum', 'warp_inst')]


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    u = unit.lower()
    return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)


def to_us(v, unit):
    v = float(v.replace(',', ''))
    return v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(unit.lower(), 1)


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        d = dict(kernel=r[idx['Kernel Name']].split('(')[0].replace('void ', '').strip())
        for col, name in COLS:
            if col in idx and r[idx[col]] != '':
                if name == 'time':
                    d[name] = to_us(r[idx[col]], units[idx[col]])
                elif name in ('dram_rd', 'dram_wr'):
                    d[name] = to_bytes(r[idx[col]], units[idx[col]])
                else:
                    try:
                        d[name] = float(r[idx[col]].replace(',', ''))
                    except ValueError:
                        d[name] = r[idx[col]]
        out.append(d)
    print('%-52s %9s %9s %9s %8s %6s %6s %6s %6s %5s %6s' % ('kernel', 'time_us', 'rd_MB', 'wr_MB', 'GB/s', 'tens%', 'l2%', 'sm%',
                                                              'issue%', 'regs', 'grid'))
    for d in out:
        tot = d.get('dram_rd', 0) + d.get('dram_wr', 0)
        print('%-52s %9.2f %9.2f %9.2f %8.0f %6.1f %6.1f %6.1f %6.1f %5d %6d' % (
            d['kernel'][:52], d.get('time', 0), d.get('dram_rd', 0) / 1e6, d.get('dram_wr', 0) / 1e6,
            tot / max(d.get('time', 1e-9), 1e-9) / 1e3, d.get('tensor%', 0), d.get('l2%', 0), d.get('sm%', 0), d.get('issue%', 0),
            int(d.get('regs', 0)), int(d.get('grid', 0))))
    if '--json' in sys.argv:
        jpath = sys.argv[sys.argv.index('--json') + 1]
        steps = float(sys.argv[sys.argv.index('--steps') + 1]) if '--steps' in sys.argv else 1.0
        try:
            js = json.load(open(jpath))
        except Exception:
            js = {}
        agg = {}
        for d in out:
            a = agg.setdefault(d['kernel'], dict(n=0, bytes=0.0, rd=0.0, wr=0.0, time=0.0))
            a['n'] += 1
            a['bytes'] += d.get('dram_rd', 0) + d.get('dram_wr', 0)
            a['rd'] += d.get('dram_rd', 0)
            a['wr'] += d.get('dram_wr', 0)
            a['time'] += d.get('time', 0)
        for k, a in agg.items():
            key = ('conv_gemm_kernel' if 'conv_gemm_kernel' in k else 'mask_fused_kernel' if 'mask_fused_pack' in k else
                   'mask_assemble_kernel' if 'mask_assemble_kernel' in k else k)
            e = js.setdefault(key, dict(launches=0, dram_bytes=0.0, dram_read_bytes=0.0, dram_write_bytes=0.0, time_us=0.0))
            e['launches'] += a['n']
            e['dram_bytes'] += a['bytes']
            e['dram_read_bytes'] += a['rd']
            e['dram_write_bytes'] += a['wr']
            e['time_us'] += a['time']
            e['source'] = 'ncu --set full, %s' % path.split('/')[-1]
        for key, e in js.items():
            if 'launches' in e and e['launches']:
                e['dram_bytes_per_launch'] = e['dram_bytes'] / e['launches']
                if key == 'conv_gemm_kernel':
                    e['dram_bytes_per_step'] = e['dram_bytes'] / steps
        json.dump(js, open(jpath, 'w'), indent=1)


if __name__ == '__main__':
    main()
