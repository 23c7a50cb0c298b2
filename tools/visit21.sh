mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hphase2 -s 2 -c 1 -f -o gpurun_out/hphase2 python tools/train_prof_run.py > gpurun_out/prof_hp2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wphase2 -s 2 -c 1 -f -o gpurun_out/wphase2 python tools/train_prof_run.py > gpurun_out/prof_wp2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/train_launches.csv python bench.py --workload train --steps 2 --warmup 1 > gpurun_out/b_ncu.log 2>&1
timeout 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/bench_train_1p25M.log 2>&1
tail -c 600 gpurun_out/bench_train_1p25M.log
timeout 900 python bench.py --workload train --train-frames 10000000 --steps 5 --warmup 3 > gpurun_out/bench_train_10M.log 2>&1
tail -c 300 gpurun_out/bench_train_10M.log
