#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2i}
export GPL_HULL_FUSED=${GPL_HULL_FUSED:-0}
timeout 900 ncu --kernel-name regex:k_hull_fast --launch-skip 1 --launch-count 1 --set full --import-source on --clock-control none \
  -o gpurun_out/${T}_hull_fast -f python bench.py --workload c5 --points 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_ncu.log 2>&1
ls -la gpurun_out/${T}_hull_fast.ncu-rep
