#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02n_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02n_tests.log
tail -5 gpurun_out/r02n_tests.log
python bench.py --steps 3 --warmup 3 --no-cpu --api-profile gpurun_out/api_profile_r02n.txt > gpurun_out/bench_r02n.json 2> gpurun_out/bench_r02n.err
tail -3 gpurun_out/bench_r02n.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r02n.json').readline()); print(d['ms_per_step'], d['e2e']['value'], d['e2e']['api'])"
head -40 gpurun_out/api_profile_r02n.txt
