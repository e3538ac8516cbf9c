"""Warp-stall / occupancy / memory view of an `ncu --page raw --csv` export, one block per launch:
    python tools/ncu_stalls.py profiles/r02_ncu_mask_mma_raw.csv > profiles/r02_ncu_mask_mma_stalls.txt
Stall reasons are `smsp__average_warps_issue_stalled_*_per_issue_active.ratio` (warps stalled on that reason per issued
instruction); only reasons >= 0.2 are listed."""
import csv
import sys

KEYS = [('gpu__time_duration.sum', 'time'), ('smsp__inst_executed.sum', 'warp instructions'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
        ('sm__warps_active.avg.per_cycle_active', 'warps active per SM'),
        ('launch__registers_per_thread', 'registers / thread'), ('launch__occupancy_limit_registers', 'CTAs / SM (registers)'),
        ('launch__occupancy_limit_shared_mem', 'CTAs / SM (shared memory)'), ('launch__grid_size', 'grid'),
        ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1 throughput %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %')]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        print(r[idx['Kernel Name']].split('(')[0].replace('void ', '').strip())
        for k, name in KEYS:
            if k in idx and r[idx[k]] != '':
                print('  %-28s %s %s' % (name, r[idx[k]], units[idx[k]]))
        stalls = []
        for h, i in idx.items():
            if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
                try:
                    v = float(r[i].replace(',', ''))
                except ValueError:
                    continue
                if v >= 0.2:
                    stalls.append((v, h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
        print('  stalled warps per issued instruction: ' + ', '.join('%s %.2f' % (n, v) for v, n in sorted(stalls, reverse=True)))
        print()


if __name__ == '__main__':
    main()
