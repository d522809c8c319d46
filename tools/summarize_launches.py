"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, steps):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            name = re.sub(r'\(.*$', '', r['Kernel Name'])
            name = re.sub(r'^void ', '', name)
            rows.append((name, r['Grid Size'], r['Block Size'], float(r['Metric Value']) / 1e6))
    agg = OrderedDict()
    for name, grid, block, ms in rows:
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ms
        a[2] = max(a[2], ms)
    total = sum(a[1] for a in agg.values())
    print(f'# ncu launch list: {path}\n')
    print(f'{len(rows)} launches captured ({steps} timed steps); per-launch times are cold-cache and serialised, compare SHARES.\n')
    print('| kernel | launches | total ms | share | max single ms |')
    print('|---|---:|---:|---:|---:|')
    for name, (n, ms, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{name}` | {n} | {ms:.3f} | {100 * ms / total:.1f}% | {mx:.3f} |')
    print(f'| **total** | {len(rows)} | {total:.3f} | 100% | |')
    print('\n## every launch, in order\n')
    print('| # | kernel | grid | block | ms |')
    print('|---:|---|---|---|---:|')
    for i, (name, grid, block, ms) in enumerate(rows):
        print(f'| {i} | `{name}` | {grid} | {block} | {ms:.4f} |')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '?')
