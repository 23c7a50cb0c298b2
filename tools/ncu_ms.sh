#!/bin/bash
# ncu capture of the multi-stream H-solve on a small fixed workload (1024 utterances x 1.5 s), launch 60 (steady state)
mkdir -p gpurun_out
SNMFNAT_DEBUG=1 timeout 300 python tools/prof_run.py 1024 1.5 > gpurun_out/prof_ms_small.log 2>&1; tail -4 gpurun_out/prof_ms_small.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hsolve_ms -s 60 -c 1 -f -o gpurun_out/hsolve_ms python tools/prof_run.py 1024 1.5 > gpurun_out/ncu_ms.log 2>&1; tail -3 gpurun_out/ncu_ms.log
