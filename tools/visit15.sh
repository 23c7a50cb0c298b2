set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hphase -s 2 -c 1 -f -o gpurun_out/hphase python tools/train_prof_run.py > gpurun_out/prof_hp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wphase -s 2 -c 1 -f -o gpurun_out/wphase python tools/train_prof_run.py > gpurun_out/prof_wp.log 2>&1
tail -3 gpurun_out/prof_hp.log
