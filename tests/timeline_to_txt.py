"""Tool (not a test): chrome trace written by tests/timeline_step.py -> the text timeline tests/analyze_timeline.py reads
(one graph-replayed step: start_us dur_us stream kernel [grid]).    python tests/timeline_to_txt.py gpurun_out/timeline.json out.txt"""
import json
import sys


def main(src, dst):
    ev = [e for e in json.load(open(src))['traceEvents'] if e.get('cat') == 'kernel']
    ev.sort(key=lambda e: e['ts'])
    marks = [e['ts'] for e in ev if 'conv1_direct_kernel' in e['name']]          # once per step (WavEncoder conv1, first kernel of its stream)
    assert len(marks) >= 3, 'need three profiled steps'
    lo, hi = marks[-2] - 1.0, marks[-1] - 1.0
    step = [e for e in ev if lo <= e['ts'] < hi]
    streams = {}
    with open(dst, 'w') as f:
        f.write('# one graph-replayed train_iter_gan step, batch 128, B200 (CUPTI via tests/timeline_step.py): start_us dur_us stream kernel grid\n')
        for e in step:
            s = streams.setdefault(e['args'].get('stream'), 's%d' % len(streams))
            name = e['name'].replace('void ', '').replace('(anonymous namespace)::', '')
            f.write('%8.1f %7.1f %-3s %-50s %s\n' % (e['ts'] - lo, e['dur'], s, name[:50], e['args'].get('grid')))
    print('kernels', len(step), 'span_us', step[-1]['ts'] + step[-1]['dur'] - lo)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
