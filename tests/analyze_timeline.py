"""Tool (not a test): which kernels own the step time?  Reads a CUPTI timeline written by tests/timeline_step.py (start_us dur_us stream
kernel grid) and, sweeping over time, attributes every microsecond either to the ONE kernel running alone at that moment ("exclusive": a
speed-up of that kernel shortens the step by the same amount), to the n kernels sharing it ("shared/n"), or to nobody ("idle": launch gaps,
stream joins).    python tests/analyze_timeline.py profiles/r01_timeline_step.txt"""
import collections
import re
import sys


def main(path):
    rows = []
    for ln in open(path):
        m = re.match(r'\s*([\d.]+)\s+([\d.]+)\s+(s\d+)\s+(.*?)\s+\[', ln)
        if ln.startswith('#') or not m:
            continue
        name = re.sub(r'[<(].*', '', m.group(4))
        rows.append((float(m.group(1)), float(m.group(1)) + float(m.group(2)), m.group(3), name))
    ev = sorted([(a, 1, i) for i, (a, b, s, n) in enumerate(rows)] + [(b, -1, i) for i, (a, b, s, n) in enumerate(rows)])
    active, last, idle = set(), 0.0, 0.0
    excl, shared, count = collections.Counter(), collections.Counter(), collections.Counter(r[3] for r in rows)
    for t, kind, i in ev:
        dt = t - last
        if dt > 0:
            if not active:
                idle += dt
            elif len(active) == 1:
                excl[rows[next(iter(active))][3]] += dt
            else:
                for j in active:
                    shared[rows[j][3]] += dt / len(active)
        last = t
        (active.add if kind == 1 else active.discard)(i)
    span = max(r[1] for r in rows)
    print('# %s: %d kernels, span %.1f us, idle (no kernel running) %.1f us' % (path, len(rows), span, idle))
    print('%-36s %6s %12s %12s' % ('kernel', 'calls', 'exclusive_us', 'shared/n_us'))
    for k in sorted(count, key=lambda k: -(excl[k] + shared[k])):
        print('%-36s %6d %12.1f %12.1f' % (k, count[k], excl[k], shared[k]))


if __name__ == '__main__':
    main(sys.argv[1])
