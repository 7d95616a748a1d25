"""Render profiles/r02_bench_*.json (one bench.py JSON line each) as the markdown table of profiles/r02_results.md."""
import glob, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for p in sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r02_bench_*.json'))):
    d = json.loads(open(p).read().strip().splitlines()[-1])
    m = re.match(r'r02_bench_(c\d)_n(\d)(.*)\.json', os.path.basename(p))
    st = d['stages_ms']
    lay = d.get('layouts') or {}
    other = [k for k in lay if k not in ('single',)][1:] if len(lay) > 1 else []
    rows.append((m.group(1), int(m.group(2)), m.group(3).strip('_'), d, st, lay, other))
fam = lambda tag: 'idorder' if 'idorder' in tag else ''   # runs before the ordered item sweep scale against their own 1-GPU run
base = {(c, f): next((r[3]['value'] for r in rows if r[0] == c and r[1] == 1 and fam(r[2]) == f), None)
        for c in set(r[0] for r in rows) for f in ('', 'idorder')}
print('| config | GPUs | users/s (resident) | users/s (e2e) | ms/step | aggregate ms (kernels / all-gather) | score ms | tensor frac | agg. alg. GB/s (frac) | overflow (pass 2 / exact) | other layout ms | verified | SM MHz |')
print('|---|---|---|---|---|---|---|---|---|---|---|---|---|')
for c, n, tag, d, st, lay, other in sorted(rows, key=lambda r: (r[0], fam(r[2]), r[1])):
    sp = ' (%.2fx)' % (d['value'] / base[(c, fam(tag))]) if base.get((c, fam(tag))) else ''
    ra, r = d.get('roofline_aggregation') or {}, d.get('roofline') or {}
    aggs = '%.2f' % st['aggregate_ms'] if st.get('aggregate_ms') else '-'
    if st.get('aggregate_kernels_ms') is not None:
        aggs += ' (%.2f / %.2f)' % (st['aggregate_kernels_ms'], st['aggregate_allgather_ms'])
    oth = ', '.join('%s %.1f' % (k, lay[k]['ms_per_step']) for k in other) or '-'
    ver = d.get('verified_vs_exact_fp32')
    if d.get('verified_sharded_embeddings') is not None:
        ver = '%s / emb %s' % (ver, d['verified_sharded_embeddings'])
    print('| %s%s | %d | %.2fM%s | %.2fM | %.2f | %s | %.2f | %.2f | %.0f (%.2f) | %s | %s | %s | %s |' % (
        c, (' ' + tag.replace('_', ' ')) if tag else '', n, d['value'] / 1e6, sp, d['e2e']['value'] / 1e6, d['ms_per_step'], aggs,
        st.get('score_ms') or 0, r.get('frac', 0) if r.get('bound') == 'tensor' else 0, ra.get('achieved', 0), ra.get('frac', 0),
        '/'.join(str(x) for x in d.get('overflow_users', [])), oth, ver, (d.get('clocks') or {}).get('sm_mhz')))
