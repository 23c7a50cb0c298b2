#!/bin/bash
# quick check of the multi-stream H-solve: its parity tests, then the serialised per-class profile on 1024 x 1.5 s
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_stream or alternative or m03 or ragged" > gpurun_out/pytest_ms.log 2>&1; tail -5 gpurun_out/pytest_ms.log
SNMFNAT_DEBUG=1 timeout 300 python tools/prof_run.py 1024 1.5 2>&1 | tail -3
