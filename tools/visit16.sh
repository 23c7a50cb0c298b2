mkdir -p gpurun_out
SNMFNAT_TRAIN_DEBUG=1 timeout 300 python tools/train_prof_run.py 2>&1 | grep -i "probe\|error" | head
