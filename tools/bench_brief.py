"""Print the headline fields of bench.py JSON lines: python tools/bench_brief.py gpurun_out/x.log ..."""
import json
import sys
for f in sys.argv[1:]:
    try:
        l = [x for x in open(f) if x.startswith('{')][-1]
        d = json.loads(l)
        k = d['kernels']
        print('%-40s step %.3f  k2 %.3f  k3 %.3f  k1 %.4f  e2e %.3f  alt %.3f' % (
            f.split('/')[-1], d['ms_per_step'], k['k2_refine']['ms'], k['k3_nn']['ms'], k['k1_interp']['ms'],
            d['e2e']['ms_per_step'], d.get('e2e_variants', {}).get('heads16_result8', d.get('e2e_full_records', {})).get('ms_per_step', 0)))
    except Exception as e:
        print(f, 'unreadable:', e)
