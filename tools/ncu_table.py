"""Markdown table of the per-kernel metrics quoted in profiles/*.md from an
ncu report (`ncu --set full`): python tools/ncu_table.py report.ncu-rep|raw.csv [regex]"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ('duration (ms)', 'gpu__time_duration.sum'),
    ('dram read', 'dram__bytes_read.sum'),
    ('dram written', 'dram__bytes_write.sum'),
    ('dram throughput, % of peak', 'dram__throughput.avg.pct_of_peak_sustained_elapsed'),
    ('grid', 'launch__grid_size'),
    ('block', 'launch__block_size'),
    ('registers / thread', 'launch__registers_per_thread'),
    ('dynamic shared memory / CTA', 'launch__shared_mem_per_block_dynamic'),
    ('warps active, % of peak', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
    ('IPC', 'sm__inst_executed.avg.per_cycle_elapsed'),
    ('FMA pipe active, %', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
    ('FP64 pipe active, %', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
    ('LSU data-pipe wavefronts, % of peak', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
    ('L2 hit rate, %', 'lts__t_sector_hit_rate.pct'),
    ('shared-memory bank conflicts', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
    ('warp instructions', 'smsp__inst_executed.sum'),
]
STALLS = 'smsp__average_warps_issue_stalled_{}_per_issue_active.ratio'
STALL_NAMES = ['long_scoreboard', 'math_pipe_throttle', 'not_selected', 'barrier',
               'mio_throttle', 'lg_throttle', 'wait', 'short_scoreboard',
               'selected', 'membar', 'dispatch_stall', 'imc_miss', 'no_instruction',
               'tex_throttle', 'sleeping', 'branch_resolving', 'drain']


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    if rep.endswith('.csv'):      # already exported with --page raw --csv
        out = open(rep).read()
        out = out[out.index('"ID"'):]
    else:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                             capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    kernels = [r for r in rows[2:] if not pat or pat.search(r[col['Kernel Name']])]
    names = [re.sub(r'\(.*', '', r[col['Kernel Name']]).replace('void ', '')[:40]
             for r in kernels]
    print('| metric | ' + ' | '.join(names) + ' |')
    print('|---|' + '---|' * len(names))
    for label, key in KEYS:
        if key not in col:
            continue
        u = units[col[key]]
        cells = []
        for r in kernels:
            v = r[col[key]]
            try:
                v = f'{float(v.replace(",", "")):.4g}'
            except ValueError:
                pass
            cells.append(v)
        print(f'| {label}{" [" + u + "]" if u and u not in ("%",) else ""} | '
              + ' | '.join(cells) + ' |')
    cells = []
    for r in kernels:
        st = []
        for s in STALL_NAMES:
            k = STALLS.format(s)
            if k in col:
                try:
                    st.append((float(r[col[k]]), s))
                except ValueError:
                    pass
        st.sort(reverse=True)
        cells.append(', '.join(f'{s.replace("_", " ")} {v:.1f}' for v, s in st[:5]))
    print('| top stalls (warps per issue) | ' + ' | '.join(cells) + ' |')


if __name__ == '__main__':
    main()
