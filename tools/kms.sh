#!/bin/bash
# per-kernel milliseconds of one bench configuration (development aid): tools/kms.sh [workload] [extra bench args]
wl=${1:-global025_quarter}; shift
python bench.py --workload $wl --steps 3 --warmup 2 --no-e2e --no-cpu --no-api "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('ms/step %.2f  value %.3e  sweep frac %.4f' % (d['ms_per_step'], d['value'], d['roofline']['frac']))
print('  ' + '  '.join('%s %.2f' % (k.replace('xmhw_',''), v) for k, v in d['kernel_ms'].items()))
"
