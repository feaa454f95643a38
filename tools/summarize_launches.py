"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into
profiles/<name>.md: per-kernel share of ONE training step (the last complete one in the capture).
Step boundaries are the three layout_fwd launches that open Model.forward."""
import collections
import csv
import re
import sys


def norm(n):
    n = n.replace('(anonymous namespace)::', '').replace('<unnamed>::', '').replace('void ', '')
    n = re.sub(r'\(.*$', '', n)
    n = re.sub(r'at::native::|at::', 'at::', n)
    return n[:80]


def main(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = [(r['Kernel Name'], r['Grid Size'], float(r['Metric Value'].replace(',', ''))) for r in csv.DictReader(lines)]
    idx = [i for i, (n, _, _) in enumerate(rows) if 'layout_fwd' in n]
    starts = [i for k, i in enumerate(idx) if k == 0 or i - idx[k - 1] > 3]
    if len(starts) >= 2 and '--whole' not in sys.argv:
        a, b = starts[-2], starts[-1]
    else:                      # capture taken with bench.py --profile-step: the file IS one step
        a, b = 0, len(rows)
    step = rows[a:b]
    tot = sum(t for _, _, t in step)
    agg = collections.defaultdict(lambda: [0, 0.0])
    mine = 0.0
    for n, g, t in step:
        k = norm(n)
        agg[k][0] += 1
        agg[k][1] += t
        if not k.startswith('at::') and 'nccl' not in k.lower():
            mine += t
    with open(out, 'w') as f:
        f.write('# ncu launch list of one training step (%s)\n\n' % path.split('/')[-1])
        f.write('Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv ... python bench.py --steps 1 --warmup 3 '
                '--no-e2e --no-cpu-baseline`; launches %d..%d of the capture = one step (bs 32, 128x128).\n' % (a, b))
        f.write('Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n')
        f.write('* launches in the step: %d, summed GPU time %.2f ms, of which libsg_b200 kernels %.1f %%\n\n' % (len(step), tot / 1e6, 100 * mine / tot))
        f.write('| kernel | launches | ms | share |\n|---|---:|---:|---:|\n')
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write('| `%s` | %d | %.3f | %.1f %% |\n' % (k, c, t / 1e6, 100 * t / tot))
        f.write('\n## tensor-core launches by grid\n\n| kernel | grid | launches | ms | us/launch |\n|---|---|---:|---:|---:|\n')
        g2 = collections.defaultdict(lambda: [0, 0.0])
        for n, g, t in step:
            m = re.search(r'(conv_tc_kernel|wgrad_tc_kernel)<([0-9, ]+)>', n)
            if m:
                g2[('%s<%s>' % m.groups(), g)][0] += 1
                g2[('%s<%s>' % m.groups(), g)][1] += t
        for (k, g), (c, t) in sorted(g2.items(), key=lambda kv: -kv[1][1])[:30]:
            f.write('| `%s` | %s | %d | %.3f | %.1f |\n' % (k, g, c, t / 1e6, t / c / 1e3))
    print('wrote', out, 'step launches', len(step), 'ms', tot / 1e6)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
