"""Summarise .ncu-rep captures (gpurun_out/) into small text/JSON files that are committed under profiles/.

    python profiles/extract_ncu.py gpurun_out/prof_x.ncu-rep profiles/r01_ncu_x.txt [json_key]
"""
import csv
import json
import os
import re
import subprocess
import sys

KEEP = re.compile(
    r'^(Kernel Name|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|'
    r'sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|'
    r'launch__(registers_per_thread|block_size|grid_size|occupancy_limit_\w+|shared_mem_per_block_dynamic)|'
    r'smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_(fma|lsu|alu|xu)\.avg\.pct_of_peak_sustained_active|'
    r'sm__pipe_fma_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|sm__pipe_tensor\w*cycles_active\.avg\.pct_of_peak_sustained_elapsed|'
    r'smsp__cycles_active\.avg|sm__cycles_elapsed\.max|smsp__inst_executed\.sum|'
    r'l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|lts__t_bytes\.sum|'
    r'smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio)$')


def main():
    rep, out = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    if rep.endswith('.csv'):  # already exported on the GPU box: ncu -i x.ncu-rep --page raw --csv > x.csv
        txt = open(rep).read()
    else:
        txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = ['# ncu --set full --clock-control none summary of %s' % os.path.basename(rep)]
    summary = {}
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if KEEP.match(h):
                lines.append('%-84s %s %s' % (h, v, u))
                d[h] = (v, u)
        lines.append('')

        def num(name):
            v, u = d[name]
            f = float(v.replace(',', ''))
            return f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1, 'ns': 1e-3, 'ms': 1e3}.get(u, 1)
        summary = {'kernel': d['Kernel Name'][0], 'dram_bytes_per_launch': num('dram__bytes_read.sum') + num('dram__bytes_write.sum'),
                   'time_us_under_ncu': num('gpu__time_duration.sum'), 'registers': d['launch__registers_per_thread'][0]}
        tp = 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'
        if tp in d:
            summary['tensor_pipe_active_pct'] = float(d[tp][0].replace(',', ''))
    with open(out, 'w') as f:
        f.write('\n'.join(lines) + '\n')
    if key:
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ncu_summary.json')
        allj = json.load(open(p)) if os.path.exists(p) else {}
        allj[key] = summary
        json.dump(allj, open(p, 'w'), indent=1, sort_keys=True)
    print('\n'.join(lines[:12]))


if __name__ == '__main__':
    main()
