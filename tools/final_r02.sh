#!/bin/bash
# Final visit of round 2: GPU tests, the default bench line, the W-solve ncu capture and the launch list of the bench command.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_default.log 2>gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wsolve_fast -s 60 -c 1 -f -o gpurun_out/r02_wsolve_fast python tools/prof_run.py 1024 1.5 > gpurun_out/ncu_w.log 2>&1; tail -1 gpurun_out/ncu_w.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"hsolve|wsolve|gain_kernel|stft|ola_int16|frame_pcm|istft|synth|regular_fft|vector_fft" -s 3000 -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train --no-parity > gpurun_out/b_ncu.log 2>&1
tail -2 gpurun_out/r02_launches.csv | cut -c1-200
