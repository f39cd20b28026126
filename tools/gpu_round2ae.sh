#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02ae_n1.json 2> gpurun_out/bench_r02ae_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r02ae_n1.json').readline())
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'api', d['e2e']['api']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['clocks'])
print({k.replace('xmhw_',''):round(v,2) for k,v in d['kernel_ms'].items()})"
