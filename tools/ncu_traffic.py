"""ncu launch list of one marked step (tools/ncu_step.py) -> profiles/r02_traffic.json (DRAM
bytes and time per launch, keyed 'layer kind', stamped with the digest of the kernel sources)
and profiles/r02_launches.md (families and launch order).

  python tools/ncu_traffic.py gpurun_out/step_ncu.csv gpurun_out/step_calls.json"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from segmentation_b200 import build as B  # noqa: E402


def parse(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = {}
    order = []
    for row in csv.DictReader(lines):
        i = int(row['ID'])
        if i not in rows:
            rows[i] = {'name': row['Kernel Name'], 'grid': row.get('Grid Size', '')}
            order.append(i)
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        m = row['Metric Name']
        if m == 'gpu__time_duration.sum':
            v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'second': 1e6}.get(u, 1.0)
            rows[i]['us'] = v
        else:
            v *= {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)
            rows[i][m] = v
    return [rows[i] for i in order]


def short(name):
    name = re.sub(r'segb::', '', name).replace('void ', '')
    return name.split('(')[0]


def main():
    launches = parse(sys.argv[1])
    calls = json.load(open(sys.argv[2]))
    groups, cur = [], None
    for l in launches:
        if 'spin_kernel' in l['name'] or 'sleep' in l['name'].lower():
            if cur is not None:
                groups.append(cur)
            cur = []
        elif cur is not None:
            cur.append(l)
    assert len(groups) == len(calls), (len(groups), len(calls))
    out = {'csrc_digest': B._digest(), 'command': 'tools/ncu_step.py under ncu --metrics '
           'gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none',
           'note': 'cold-cache, serialised launches: compare shares, not absolutes',
           'launches': {}}
    table = []
    fam = {}
    for (fn, tag, family, fl, by), g in zip(calls, groups):
        key = bench.launch_key(fn, tag)
        us = sum(l.get('us', 0.0) for l in g)
        rd = sum(l.get('dram__bytes_read.sum', 0.0) for l in g)
        wr = sum(l.get('dram__bytes_write.sum', 0.0) for l in g)
        e = out['launches'].setdefault(key, {'us': 0.0, 'dram_bytes': 0.0, 'dram_read': 0.0,
                                             'dram_write': 0.0, 'kernels': [], 'family': family,
                                             'alg_bytes': 0.0, 'alg_flops': 0.0})
        e['us'] += us; e['dram_bytes'] += rd + wr; e['dram_read'] += rd; e['dram_write'] += wr
        e['alg_bytes'] += by; e['alg_flops'] += fl
        e['kernels'] += [short(l['name']) + ' ' + l['grid'] for l in g]
        table.append((key, family, ', '.join(short(l['name']) for l in g), us, rd + wr, by, fl))
        f = fam.setdefault(family, [0, 0.0, 0.0, 0.0, 0.0])
        f[0] += len(g); f[1] += us; f[2] += rd + wr; f[3] += by; f[4] += fl
    total = sum(t[3] for t in table)
    os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
    with open(os.path.join(ROOT, 'profiles', 'r02_traffic.json'), 'w') as f:
        json.dump(out, f, indent=1)
    with open(os.path.join(ROOT, 'profiles', 'r02_launches.md'), 'w') as f:
        f.write('# r02 - ncu launch list of one U-Net 256x256 bs16 train step (eager, one stream)\n\n')
        f.write('Command: `%s` (see tools/ncu_step.py); kernel sources digest `%s`.\n'
                'Per-launch times under ncu are cold-cache and serialised: read SHARES.\n\n'
                % (out['command'], out['csrc_digest'][:16]))
        f.write('%d kernel launches, sum %.1f us\n\n' % (sum(v[0] for v in fam.values()), total))
        f.write('| family | launches | us | share | DRAM MB (ncu) | algorithmic MB | TFLOP/s | alg GB/s |\n'
                '|---|---:|---:|---:|---:|---:|---:|---:|\n')
        for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
            f.write('| %s | %d | %.1f | %.1f %% | %.1f | %.1f | %.0f | %.0f |\n'
                    % (k, v[0], v[1], 100 * v[1] / total, v[2] / 1e6, v[3] / 1e6,
                       v[4] / (v[1] * 1e-6) / 1e12 if v[1] else 0, v[3] / (v[1] * 1e-6) / 1e9 if v[1] else 0))
        f.write('\n## launch order\n\n| launch | family | kernel(s) | us | DRAM MB | algorithmic MB | GFLOP |\n'
                '|---|---|---|---:|---:|---:|---:|\n')
        for key, family, ks, us, dr, by, fl in table:
            f.write('| %s | %s | `%s` | %.1f | %.1f | %.1f | %.2f |\n'
                    % (key, family, ks[:70], us, dr / 1e6, by / 1e6, fl / 1e9))
    print('wrote profiles/r02_traffic.json, profiles/r02_launches.md; total %.1f us' % total)


if __name__ == '__main__':
    main()
