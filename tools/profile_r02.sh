#!/bin/bash
# Round-2 profiling visit: launch list of the bench command (only this library's kernels: 600 launches after 3000 of
# them, i.e. steady-state hops of the warm-up step), full ncu captures of the two MU kernels on the fixed 1024 x 1.5 s
# workload (steady-state launch 60).
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hsolve|wsolve|gain_kernel|stft|ola_int16|frame_pcm|istft|synth|regular_fft|vector_fft" -s 3000 -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train --no-parity > gpurun_out/b_ncu.log 2>&1
tail -2 gpurun_out/r02_launches.csv | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hsolve_ms -s 60 -c 1 -f -o gpurun_out/r02_hsolve_ms python tools/prof_run.py 1024 1.5 > gpurun_out/ncu_h.log 2>&1; tail -1 gpurun_out/ncu_h.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wsolve_fast -s 60 -c 1 -f -o gpurun_out/r02_wsolve_fast python tools/prof_run.py 1024 1.5 > gpurun_out/ncu_w.log 2>&1; tail -1 gpurun_out/ncu_w.log
