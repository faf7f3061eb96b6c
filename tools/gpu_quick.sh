#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_onepass.py tests/test_gpu_jet.py tests/test_gpu_f64.py -x -q 2>&1 | tail -3
