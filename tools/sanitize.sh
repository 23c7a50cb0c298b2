#!/bin/bash
# compute-sanitizer on a small batch (16 utterances x 0.6 s: init frames, gated hops, W-solves): memcheck, racecheck
# (shared-memory hazards inside a CTA: the st.async / mbarrier kernels), synccheck.  Summaries -> gpurun_out/sanitize_*.log
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  SNMFNAT_HSOLVE=ms timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/prof_run.py 16 0.6 > gpurun_out/sanitize_${tool}.log 2>&1
  echo "== $tool (forced multi-stream H-solve)"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" gpurun_out/sanitize_${tool}.log | head -5
done
SNMFNAT_HSOLVE=single timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/prof_run.py 16 0.6 > gpurun_out/sanitize_racecheck_single.log 2>&1
echo "== racecheck (per-stream H-solve)"; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/sanitize_racecheck_single.log | head -5
