#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "oisst or nosmooth or synth_daily or config4" 2>&1 | tail -2
for i in 1 2; do timeout 200 bash tools/kms.sh global025_30yr 2>&1 | cut -c1-260; done
