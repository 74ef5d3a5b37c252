"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'lts__t_sectors_srcunit_tex_op_red.sum', 'smsp__inst_executed_op_global_red.sum']


def main(path, stalls=True):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[hdr.index('Kernel Name')][:70])
        for k in KEYS:
            if k in hdr:
                print(f'   {k:75s} {r[hdr.index(k)]} {units[hdr.index(k)]}')
        if stalls:
            st = []
            for i, h in enumerate(hdr):
                if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
                    try:
                        st.append((float(r[i]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
                    except ValueError:
                        pass
            st.sort(reverse=True)
            print('   stalls (warps per issue):', ', '.join(f'{n}={v:.2f}' for v, n in st[:7]))


if __name__ == '__main__':
    main(sys.argv[1])
