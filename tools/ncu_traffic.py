"""Turn an ncu capture into the DRAM-traffic figures bench.py prints (`roofline.traffic`, `roofline_aggregation.traffic`).

    ncu -i capture.ncu-rep --page raw --csv --print-units base > capture.csv
    python tools/ncu_traffic.py capture.csv c2 [--steps N] [--note "..."]      # updates profiles/traffic.json

Every kernel launch in the capture is assigned to a stage of the step by its name (see STAGES); per stage the script
sums dram__bytes_read.sum + dram__bytes_write.sum over the launches and divides by the number of captured steps, and
records duration, L2 hit rate and launch count next to it. bench.py reads profiles/traffic.json -- the traffic it
reports is therefore always that of a committed capture of the CURRENT kernels (or null), never a constant in the code.
"""
import argparse
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGES = [  # first match wins
    ('score', r'score_topk_kernel'),
    ('aggregate', r'sage_fused_kernel|sage_generic_kernel|long_partial_kernel|long_reduce_kernel|collect_long_rows_kernel|'
                  r'pack_weights|weight_scale_kernel|linear_tc5_kernel|split_rows_f16_kernel|split_weights_f16_kernel|linear_tc_kernel'),
    ('embed_in', r'linear_small'),
    ('rescore', r'rescore_kernel|exact_topk_kernel|topk_merge_kernel'),
    ('prep', r'score_prep_kernel|colmean|score_band_kernel|order_\w+_kernel|permute_rows_kernel|radix_\w+_kernel|scan_kernel|csr_finish_kernel'),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('config')
    ap.add_argument('--steps', type=int, default=1, help='bench steps covered by the capture')
    ap.add_argument('--note', default='')
    ap.add_argument('--scale', type=float, default=1.0, help='capture covers 1/scale of a step (e.g. one of two identical layers)')
    ap.add_argument('--partial', action='store_true', help='capture misses launches of the step: bench.py then reports null')
    ap.add_argument('--out', default=os.path.join(ROOT, 'profiles', 'traffic.json'))
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    rows = rows[next(i for i, r in enumerate(rows) if 'Kernel Name' in r):]
    if 'Metric Name' in rows[0]:        # long format (`ncu --metrics ... --csv`): one row per (launch, metric) -> wide
        c = {h: i for i, h in enumerate(rows[0])}
        names, wide, unit = [], {}, {}
        for r in rows[1:]:
            if len(r) < len(rows[0]):
                continue
            if r[c['Metric Name']] not in names:
                names.append(r[c['Metric Name']])
            unit[r[c['Metric Name']]] = r[c['Metric Unit']]
            wide.setdefault(r[c['ID']], {'Kernel Name': r[c['Kernel Name']]})[r[c['Metric Name']]] = r[c['Metric Value']]
        rows = [['Kernel Name'] + names, [''] + [unit[n] for n in names]] + \
               [[w['Kernel Name']] + [w.get(n, '0') for n in names] for w in wide.values()]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    need = ['Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum']
    for n in need:
        if n not in col:
            sys.exit('column %s missing: export with `ncu -i X --page raw --csv --print-units base` from a --set full capture' % n)
    units = rows[1]
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'nsecond': 1e-6,
             'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}

    def val(r, name):
        u = units[col[name]].strip()
        return float(r[col[name]].replace(',', '')) * scale.get(u, 1.0)
    out = {}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[col['Kernel Name']]
        stage = next((s for s, pat in STAGES if re.search(pat, name)), None)
        if stage is None:
            continue
        e = out.setdefault(stage, {'dram_bytes': 0.0, 'kernel_ms': 0.0, 'launches': 0, 'kernels': {}})
        b = val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum')
        e['dram_bytes'] += b
        e['kernel_ms'] += val(r, 'gpu__time_duration.sum')
        e['launches'] += 1
        short = re.sub(r'\(.*', '', name).split('::')[-1][:60]
        k = e['kernels'].setdefault(short, {'launches': 0, 'dram_bytes': 0.0, 'ms': 0.0})
        k['launches'] += 1
        k['dram_bytes'] += b
        k['ms'] += val(r, 'gpu__time_duration.sum')
    for e in out.values():
        for k in ('dram_bytes', 'kernel_ms'):
            e[k] = e[k] / a.steps * a.scale
        e['launches'] //= a.steps
        e['source'] = os.path.basename(a.csv)
        e['scaled_by'] = a.scale
        e['partial'] = bool(a.partial)
        e['note'] = a.note or 'ncu --set full --clock-control none, per step; durations are serialised cold-cache replays'
    table = json.load(open(a.out)) if os.path.exists(a.out) else {}
    table[a.config] = out
    json.dump(table, open(a.out, 'w'), indent=1, sort_keys=True)
    for s, e in out.items():
        print('%-10s %8.3f GB DRAM  %8.3f ms  %d launches' % (s, e['dram_bytes'] / 1e9, e['kernel_ms'], e['launches']))


if __name__ == '__main__':
    main()
