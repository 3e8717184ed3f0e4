#!/bin/bash
# quick row-owner check: P2-P1 / rowown parity tests, then plain timings on two duct sizes
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "p2 or rowown or generic or every" 2>&1 | tail -2
timeout 100 python tools/prof_rowown.py 24 2
timeout 100 python tools/prof_rowown.py 40 2
