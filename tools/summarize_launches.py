#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into a
per-kernel table for ONE training step (launches between two consecutive adam_kernel launches).

    python tools/summarize_launches.py gpurun_out/launches.csv [step_index] > profiles/....md

ncu serialises kernels and runs them cold-cache: compare SHARES, not absolute times."""
import collections
import csv
import re
import sys


def load(path):
    lines = open(path).read().splitlines(True)
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = []
    for r in csv.DictReader(lines[start:]):
        try:
            rows.append((r['Kernel Name'], float(r['Metric Value'].replace(',', '')), r['Metric Unit']))
        except Exception:
            pass
    return rows


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rows = load(path)
    adam = [i for i, r in enumerate(rows) if 'adam_kernel' in r[0]]
    a, b = adam[which], adam[which + 1]
    step = rows[a + 1:b + 1]
    scale = 1e-3 if rows[0][2] == 'ns' else 1.0
    tot = sum(r[1] for r in step) * scale
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t, _ in step:
        n = re.sub(r'\(.*', '', re.sub(r'<.*', '', n)).replace('void ', '')
        agg[n][0] += 1
        agg[n][1] += t * scale
    ours = sum(t for n, (c, t) in agg.items() if 'unflow::' in n)
    print("# ncu launch list, one training step (%d launches, %.2f ms serialised, cold cache)\n" % (len(step), tot / 1e3))
    print("source: `%s`, step %d of the run; hand-written kernels (`unflow::`) = %.1f %% of the step\n"
          % (path, which, 100 * ours / tot))
    print("| share | time (us) | launches | kernel |\n|---:|---:|---:|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / tot < 0.003:
            continue
        print("| %.2f %% | %.1f | %d | `%s` |" % (100 * t / tot, t, c, n[:110]))


if __name__ == "__main__":
    main()
