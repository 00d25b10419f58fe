"""Development aid: print the essentials of bench.py JSON lines (files given on the command line; the last JSON line of each)."""
import json, sys
for path in sys.argv[1:]:
    try:
        txt = [l for l in open(path) if l.startswith('{')][-1]
        d = json.loads(txt)
    except Exception as e:
        print(path, "no line", e); continue
    print(path.split('/')[-1], "value %.2f" % d['value'], 'ms/step %.1f' % d['ms_per_step'], 'e2e', d['e2e'] and round(d['e2e']['value'] or 0, 2),
          'parity', d.get('parity') and (d['parity']['ok'], d['parity']['max_rel']), 'regimes', [(r['restarts_per_step'], round(r['value'], 1)) for r in d.get('regimes', [])],
          'roof', d['roofline']['kernel'], d['roofline']['frac'], 'whole', d['roofline']['whole_step_frac'])
    print('   ', {k: (round(v['ms_per_step'], 1), v.get('frac')) for k, v in d['roofline']['kernel_classes'].items()})
