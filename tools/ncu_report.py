"""Markdown table of the kernels in an `ncu --set full` report: duration, DRAM traffic, tensor-pipe / issue / MUFU activity,
registers.  Usage: python tools/ncu_report.py report.ncu-rep [more.ncu-rep ...]   (reads with `ncu -i ... --page raw --csv`)"""
import csv, subprocess, sys
COLS = [('gpu__time_duration.sum', 'us', 'duration'), ('dram__bytes_read.sum', 'MB', 'dram rd'), ('dram__bytes_write.sum', 'MB', 'dram wr'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', '%', 'tensor % elapsed'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', '%', 'tensor % active'),
        ('smsp__issue_active.avg.pct', '%', 'issue %'), ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed', '%', 'XU %'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', '%', 'smem wavefronts %'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', '%', 'dram %'), ('launch__registers_per_thread', '', 'regs'),
        ('launch__grid_size', '', 'grid')]
def scale(v, unit, want):
    v = float(v.replace(',', '')) if v not in ('', 'n/a') else float('nan')
    f = {('ns', 'us'): 1e-3, ('us', 'us'): 1, ('ms', 'us'): 1e3, ('nsecond', 'us'): 1e-3, ('usecond', 'us'): 1, ('msecond', 'us'): 1e3, ('byte', 'MB'): 1e-6, ('Kbyte', 'MB'): 1e-3, ('Mbyte', 'MB'): 1,
         ('Gbyte', 'MB'): 1e3}.get((unit, want), 1)
    return v * f
print('| kernel | ' + ' | '.join(c[2] + (f' ({c[1]})' if c[1] else '') for c in COLS) + ' |')
print('|---|' + '---|' * len(COLS))
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index('Kernel Name')
    for r in rows[2:]:
        name = r[kn].replace('dv::', '')
        name = name[:name.index('(')] if '(' in name else name
        cells = []
        for key, want, _ in COLS:
            if key in hdr:
                i = hdr.index(key)
                v = scale(r[i], units[i], want)
                cells.append(f'{v:.1f}' if want else f'{v:.0f}')
            else:
                cells.append('-')
        print(f'| `{name}` | ' + ' | '.join(cells) + ' |')
