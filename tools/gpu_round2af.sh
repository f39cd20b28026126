#!/bin/bash
python bench.py --steps 5 --warmup 3 --no-cpu --no-api --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['clocks'])"
