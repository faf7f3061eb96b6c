#!/bin/bash
# one-pass kernel iteration: parity tests, then timings at config-3 / config-4 shapes
timeout 900 python -m pytest tests/test_gpu_onepass.py tests/test_gpu_reference_chain.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/onepass_bench.py cfg3 --points 4194304,33554432 --variants onepass 2>&1 | cut -c1-330
timeout 300 python tools/onepass_bench.py cfg4 --points 4194304 --variants onepass 2>&1 | cut -c1-330
