#!/usr/bin/env python
"""Summarise `ncu --set full` captures (exported here with `ncu -i X.ncu-rep --page raw --csv`) into a markdown table and
into profiles/traffic.json (DRAM bytes per launch of the heaviest launch of each kernel family, read by bench.py for
roofline.traffic).

    python tools/ncu_summary.py NAME=path.ncu-rep[:algorithmic_bytes[:family]] ... --md profiles/r02_ncu.md --traffic profiles/traffic.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [('gpu__time_duration.sum', 'duration'), ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe % (active)'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe % (elapsed)'),
        ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
        ('l1tex__m_xbar2l1tex_read_bytes.sum', 'L2 -> SM bytes'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
        ('launch__registers_per_thread', 'registers'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__cluster_size', 'cluster')]
MULT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def load(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [(dict(zip(hdr, r)), dict(zip(hdr, units))) for r in rows[2:]]


def main(argv):
    md, traffic_path, specs = None, None, []
    i = 0
    while i < len(argv):
        if argv[i] == '--md':
            md = argv[i + 1]; i += 2
        elif argv[i] == '--traffic':
            traffic_path = argv[i + 1]; i += 2
        else:
            specs.append(argv[i]); i += 1
    lines, traffic = [], {}
    for spec in specs:
        name, rest = spec.split('=', 1)
        parts = rest.split(':')
        path, alg, fam = parts[0], (float(parts[1]) if len(parts) > 1 and parts[1] else None), (parts[2] if len(parts) > 2 else None)
        lines += ['## %s  (`%s`)' % (name, path), '', '| launch | ' + ' | '.join(k for _, k in KEYS) + ' |', '|---|' + '---|' * len(KEYS)]
        for d, u in load(path):
            kn = d['Kernel Name'].replace('<unnamed>::', '').split('(')[0]
            vals = ['%s %s' % (d.get(k, ''), u.get(k, '')) for k, _ in KEYS]
            lines.append('| `%s` | ' % kn[:48] + ' | '.join(vals) + ' |')
            if fam and fam not in traffic:
                by = sum(float(d[k].replace(',', '')) * MULT.get(u[k], 1) for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
                traffic[fam] = {'launch': name, 'dram_bytes': int(by), 'algorithmic_bytes': alg,
                                'duration_us': float(d['gpu__time_duration.sum'].replace(',', '')),
                                'tensor_pipe_pct_active': float(d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'].replace(',', '')),
                                'source': path}
        lines.append('')
    if md:
        open(md, 'w').write('\n'.join(lines) + '\n')
    if traffic_path:
        json.dump(traffic, open(traffic_path, 'w'), indent=1)
    print('\n'.join(lines))


if __name__ == '__main__':
    main(sys.argv[1:])
