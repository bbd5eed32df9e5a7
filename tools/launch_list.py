"""Summarise an ncu launch list (gpu__time_duration.sum per launch, csv) of `bench.py --profile` into kernel shares of the LAST step.
usage: python tools/launch_list.py gpurun_out/launches.csv profiles/r01_step_launches_vN.json "<note>" """
import csv
import json
import re
import sys


def main(path, out, note):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r['Metric Name'] == 'gpu__time_duration.sum':
            rows.append((r['Kernel Name'], float(r['Metric Value'].replace(',', ''))))
    # a step = the launches between two tower inputs: split at the first kernel of a step (amax_abs / split of the input)
    starts = [i for i, (k, _) in enumerate(rows) if 'amax_abs_kernel' in k]
    if len(starts) >= 2:
        seg = rows[starts[-2]:starts[-1]] if len(rows) - starts[-1] < (starts[-1] - starts[-2]) else rows[starts[-1]:]
    else:
        seg = rows
    agg = {}
    for k, t in seg:
        name = re.sub(r'\(.*$', '', k).strip()
        name = re.sub(r'^void ', '', name)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += t
    tot = sum(t for _, t in seg)
    ks = sorted(({'kernel': k, 'launches': v[0], 'time_ns': v[1], 'share': v[1] / tot} for k, v in agg.items()), key=lambda d: -d['time_ns'])
    lib = sum(d['time_ns'] for d in ks if 'ptb::' in d['kernel']) / tot
    json.dump({'source': note, 'unit': 'ns', 'launches_in_capture': len(rows), 'step_launches': len(seg), 'step_total_ns': tot, 'share_libptb': lib,
               'kernels': ks}, open(out, 'w'), indent=1)
    print(out, 'step', tot / 1e6, 'ms', 'lib share', round(lib, 4))
    for d in ks[:8]:
        print(f"  {d['share']:.3f} {d['launches']:3d} {d['kernel'][:90]}")


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else '')
