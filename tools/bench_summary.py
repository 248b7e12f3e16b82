"""One line per bench.py JSON file: ms per step, the chained kernel's time and fraction of peak.  Usage: python tools/bench_summary.py "glob"."""
import sys, json, glob
for f in sorted(glob.glob(sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms/step %.4f' % d['ms_per_step'], 'kernel', d['roofline']['kernel'].split(';')[-1], 'frac %.3f' % d['roofline']['frac'], 'loss', d['config'].get('loss_terms'))
    except Exception as e:
        print(f, 'ERR', e)
