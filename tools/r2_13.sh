#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2o}
timeout 900 ncu --kernel-name regex:k_pip_build_fill --launch-skip 2 --launch-count 1 --set full --import-source on --clock-control none \
  -o gpurun_out/${T}_build_fill -f python tools/exp_pip2.py --reps 2 --tag ncu > gpurun_out/${T}_ncu.log 2>&1
ls -la gpurun_out/${T}_build_fill.ncu-rep
