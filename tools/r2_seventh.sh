#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=${1:-r2h}
timeout 900 python -m pytest tests/test_gpu_hull.py -q -x > gpurun_out/${T}_pytest_hull.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_hull.log
tail -4 gpurun_out/${T}_pytest_hull.log
for cfg in "GPL_HULL_FUSED=1" "GPL_HULL_FUSED=0" "GPL_HULL_FUSED=1 GPL_HULL_MINB=4" "GPL_HULL_FUSED=0 GPL_HULL_MINB=4"; do
  echo "== $cfg"; env $cfg timeout 600 python bench.py --workload c5 --points 3000000 --steps 3 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/${T}_c5.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c5 3M polys: ms/step',round(d['ms_per_step'],2), d['verify'])"
done
timeout 600 ncu --kernel-name regex:k_hull_fast --launch-skip 1 --launch-count 1 --clock-control none \
  --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio \
  python bench.py --workload c5 --points 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | grep -E "k_hull_fast|gpu__time|inst_executed|issue_active|warps_active|bank_conflicts|local_op|scoreboard" | head -20
