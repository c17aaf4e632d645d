"""Print the handful of ncu metrics that decide the roofline story, per kernel launch in a .ncu-rep (read on the CPU box)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'lts__t_sector_hit_rate.pct']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('----', r[idx['Kernel Name']][:90])
        for k in KEYS:
            if k in idx and r[idx[k]] not in ('', 'n/a'):
                print(f'    {k:88s} {r[idx[k]]} {units[idx[k]]}')
        if 'dram__bytes_read.sum' in idx:
            try:
                t = float(r[idx['gpu__time_duration.sum']].replace(',', ''))
                tu = units[idx['gpu__time_duration.sum']]
                t_s = t * {'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9, 's': 1}.get(tu, 1e-6)
                def tob(v, u): return float(v.replace(',', '')) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}.get(u, 1)
                rd = tob(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']])
                wr = tob(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
                print(f'    => DRAM traffic {(rd + wr) / 1e6:.1f} MB, {(rd + wr) / t_s / 1e9:.0f} GB/s')
            except Exception:
                pass


if __name__ == '__main__':
    main(sys.argv[1])
