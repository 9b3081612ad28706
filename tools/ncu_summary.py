#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into the metrics the design doc argues from.  usage: ncu_summary.py rep [rep...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'registers/thread'),
    ('launch__occupancy_limit_registers', 'occupancy limit (regs) blocks/SM'),
    ('launch__occupancy_limit_shared_mem', 'occupancy limit (smem) blocks/SM'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active % of 64'),
    ('smsp__inst_executed.sum', 'warp instructions executed'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active', 'ALU pipe %'),
    ('sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'FMA pipe %'),
    ('sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'LSU pipe %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
    ('smsp__thread_inst_executed_per_inst_executed.ratio', 'active threads / instruction'),
    ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram throughput % of peak'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'), ('l1tex__t_sector_hit_rate.pct', 'L1 hit rate %'),
]


def main():
    for rep in sys.argv[1:]:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        print('## %s\n' % rep.split('/')[-1])
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            u = dict(zip(hdr, units))
            print('### %s\n' % d['Kernel Name'].split('(')[0])
            print('| metric | value |\n|---|---|')
            for k, name in KEYS:
                if k in d and d[k] != '':
                    print('| %s (`%s`) | %s %s |' % (name, k, d[k], u.get(k, '')))
            st = {k: float(v) for k, v in d.items() if k.startswith('smsp__average_warp') and 'issue_stalled' in k and k.endswith('.ratio') and v}
            print('\nstall cycles per issued instruction (top): ' + ', '.join(
                '%s %.2f' % (k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''), v)
                for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]))
            print()


if __name__ == '__main__':
    main()
