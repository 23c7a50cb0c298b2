#!/bin/bash
# Same-box A/B of library builds: every tools/probe_libs/lib_*.so takes the place of the in-tree library in turn and runs
# the fixed profiling workload (per-class event times, serialised) twice.  Usage: bash tools/ab_libs.sh [n_utt] [seconds]
n=${1:-1024}; s=${2:-1.5}
cp se_snmf_nat_b200/libsnmfnat.so /tmp/lib_saved.so
for rep in 1 2; do
  for l in tools/probe_libs/lib_*.so; do
    cp $l se_snmf_nat_b200/libsnmfnat.so
    echo "== $l (run $rep)"
    python tools/prof_run.py $n $s 2>&1 | tail -2
  done
done
cp /tmp/lib_saved.so se_snmf_nat_b200/libsnmfnat.so
