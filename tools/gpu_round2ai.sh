#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweeps_forced or persistent" 2>&1 | tail -3
