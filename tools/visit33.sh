mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hsolve_fast -s 60 -c 1 -f -o gpurun_out/hsolve_fast python tools/prof_run.py 148 2.0 > gpurun_out/prof_h.log 2>&1
