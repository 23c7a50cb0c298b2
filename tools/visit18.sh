mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hphase2 -s 2 -c 1 -f -o gpurun_out/hphase2 python tools/train_prof_run.py > gpurun_out/prof_hp2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wphase2 -s 2 -c 1 -f -o gpurun_out/wphase2 python tools/train_prof_run.py > gpurun_out/prof_wp2.log 2>&1
tail -2 gpurun_out/prof_hp2.log
