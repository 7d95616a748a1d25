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
    ('prep', r'score_prep_kernel|colmean|score_band_kernel'),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('config')
    ap.add_argument('--steps', type=int, default=1, help='bench steps covered by the capture')
    ap.add_argument('--note', default='')
    ap.add_argument('--out', default=os.path.join(ROOT, 'profiles', 'traffic.json'))
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
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
            e[k] /= a.steps
        e['launches'] //= a.steps
        e['source'] = os.path.basename(a.csv)
        e['note'] = a.note or 'ncu --set full --clock-control none, per step; durations are serialised cold-cache replays'
    table = json.load(open(a.out)) if os.path.exists(a.out) else {}
    table[a.config] = out
    json.dump(table, open(a.out, 'w'), indent=1, sort_keys=True)
    for s, e in out.items():
        print('%-10s %8.3f GB DRAM  %8.3f ms  %d launches' % (s, e['dram_bytes'] / 1e9, e['kernel_ms'], e['launches']))


if __name__ == '__main__':
    main()
